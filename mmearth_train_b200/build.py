"""Builds ``lib/libmpmae.so`` for sm_100a with nvcc (cross-compiles without a GPU)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "csrc", "mpmae.cu")
OUT = os.path.join(HERE, "lib", "libmpmae.so")
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-shared",
              "-Xcompiler", "-fPIC", "-Xptxas", "-v"]


STAMP = OUT + ".srchash"


def _source_hash(flags) -> str:
    """Hash of everything the library is built from (sources, header, flags): staleness is decided by content, not by
    mtimes, which a copy of the tree (the snapshot sent to a GPU box) does not preserve."""
    h = hashlib.sha256(" ".join(flags).encode())
    deps = sorted(os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc")))
    deps.append(os.path.join(HERE, "..", "include", "mpmae.h"))
    for d in deps:
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def _flags():
    flags = list(NVCC_FLAGS)
    if os.environ.get("MPMAE_BUILD_KNOBS"):      # timing experiments of tools/dbg_sweep.py (never in the shipped library)
        flags.append("-DMPMAE_TC_KNOBS=1")
    return flags


def _stale() -> bool:
    if not os.path.isfile(OUT) or not os.path.isfile(STAMP):
        return True
    with open(STAMP) as f:
        return f.read().strip() != _source_hash(_flags())


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return OUT
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    flags = _flags()
    cmd = [nvcc] + flags + ["-o", OUT, SRC]
    r = subprocess.run(cmd, capture_output=True, text=True)
    log = os.path.join(HERE, "lib", "build.log")
    with open(log, "w") as f:
        f.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed ({r.returncode}); see {log}")
    with open(STAMP, "w") as f:
        f.write(_source_hash(flags) + "\n")
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
