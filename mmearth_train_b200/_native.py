"""ctypes binding of the C-ABI library ``lib/libmpmae.so`` (``include/mpmae.h``).

The library is the product: there is no Python / PyTorch fallback.  If it has not been built
(``python -c "import __graft_entry__ as g; g.build()"`` or ``mmearth_train_b200.build.build()``)
importing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

MAX_MOD = 16
PIXEL_CONTINUOUS, PIXEL_CATEGORICAL, IMAGE_CATEGORICAL, IMAGE_CONTINUOUS = 0, 1, 2, 3
STAGE_MASK, STAGE_ENCODER, STAGE_DECODER, STAGE_LOSS = 1, 2, 4, 8
BWD_LOSS, BWD_DECODER, BWD_ENCODER = 0, 1, 2

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPMAE_LIB") or os.path.join(_HERE, "lib", "libmpmae.so")   # MPMAE_LIB: experiment builds (tools/)


class Cfg(C.Structure):
    _fields_ = [
        ("batch", C.c_int32), ("img_size", C.c_int32), ("patch_size", C.c_int32), ("in_chans", C.c_int32),
        ("depths", C.c_int32 * 4), ("dims", C.c_int32 * 4), ("dec_dim", C.c_int32), ("dec_depth", C.c_int32),
        ("mask_ratio", C.c_float), ("loss_aggr", C.c_int32), ("n_mod", C.c_int32),
        ("mod_kind", C.c_int32 * MAX_MOD), ("mod_chans", C.c_int32 * MAX_MOD), ("mod_norm_pix", C.c_int32 * MAX_MOD),
        ("gemm_backend", C.c_int32),
    ]


class IO(C.Structure):
    _fields_ = [
        ("params", C.c_void_p), ("grads", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("noise", C.c_void_p), ("s2_input", C.c_void_p), ("targets", C.c_void_p * MAX_MOD),
        ("mask", C.c_void_p), ("pred_pixel", C.c_void_p), ("pred_image", C.c_void_p), ("losses", C.c_void_p),
        ("grad_out", C.c_void_p), ("flags", C.c_void_p),
    ]


class GemmDesc(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("a", "b", "bias", "resid", "aux", "aux2", "kg", "out", "out2", "colsum", "colsum2",
                                          "scratch")] + [("M", C.c_int64), ("N", C.c_int32), ("K", C.c_int32),
                                                         ("group_rows", C.c_int32), ("a_gelu", C.c_int32)] + \
               [(n, C.c_void_p) for n in ("a_scale", "acc_scale", "grn_gsq", "grn_gamma", "grn_nx", "grn_scale", "grn_denom")] + \
               [("grn_eps", C.c_float)]


RAW_MAX_BANDS = 16


class RawDesc(C.Structure):
    _fields_ = [("src", C.c_void_p), ("out", C.c_void_p), ("l2a", C.c_void_p), ("lut", C.c_void_p), ("inner", C.c_int64),
                ("B", C.c_int32), ("src_bands", C.c_int32), ("n_bands", C.c_int32), ("src_type", C.c_int32),
                ("out_int64", C.c_int32), ("has_nodata", C.c_int32), ("normalize", C.c_int32), ("nodata", C.c_double),
                ("band", C.c_int32 * RAW_MAX_BANDS), ("mean", (C.c_double * RAW_MAX_BANDS) * 2),
                ("std", (C.c_double * RAW_MAX_BANDS) * 2)]


def gemm_scratch_floats(N: int, K: int) -> int:
    """floats of scratch ``mpmae_gemm_rows`` / ``mpmae_gemm_epi`` need for the split weight (backends 1 and 3)"""
    return 2 * N * (((K + 31) // 32) * 32)


class NativeError(RuntimeError):
    pass


def _load():
    if not os.path.isfile(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA library has not been built.  Run `python -c \"import "
            "__graft_entry__ as g; g.build()\"` at the repository root (needs nvcc).  There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    P, I32, I64, F = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    sig = {
        "mpmae_last_error": (C.c_char_p, []),
        "mpmae_version": (C.c_int, []),
        "mpmae_plan_create": (C.c_int, [C.POINTER(Cfg), C.POINTER(P)]),
        "mpmae_plan_destroy": (None, [P]),
        "mpmae_param_total": (I64, [P]),
        "mpmae_param_count": (I32, [P]),
        "mpmae_param_info": (C.c_int, [P, I32, C.c_char_p, I32, C.POINTER(I64), C.POINTER(I32), C.POINTER(I64)]),
        "mpmae_param_decay": (I32, [P, I32]),
        "mpmae_visible_patches": (I32, [P]),
        "mpmae_workspace_bytes": (C.c_size_t, [P]),
        "mpmae_pred_pixel_cols": (I32, [P]),
        "mpmae_pred_image_cols": (I32, [P]),
        "mpmae_pred_col_offset": (I32, [P, I32]),
        "mpmae_tap_info": (C.c_int, [P, C.c_char_p, C.POINTER(I64), C.POINTER(I64), C.POINTER(I64)]),
        "mpmae_tap_count": (I32, [P]),
        "mpmae_tap_name": (C.c_int, [P, I32, C.c_char_p, I32]),
        "mpmae_launch_count": (I32, [P, I32]),
        "mpmae_profile_begin": (C.c_int, [P]),
        "mpmae_profile_report": (C.c_int, [P, C.c_char_p, I32]),
        "mpmae_forward": (C.c_int, [P, C.POINTER(IO), P]),
        "mpmae_forward_encoder": (C.c_int, [P, C.POINTER(IO), P]),
        "mpmae_forward_stages": (C.c_int, [P, C.POINTER(IO), I32, P]),
        "mpmae_backward": (C.c_int, [P, C.POINTER(IO), P]),
        "mpmae_backward_part": (C.c_int, [P, C.POINTER(IO), I32, P]),
        "mpmae_backward_part_range": (C.c_int, [P, I32, C.POINTER(I64), C.POINTER(I64)]),
        "mpmae_encoder_features": (C.c_int, [P, C.POINTER(IO), P, P]),
        "mpmae_gemm_rows": (C.c_int, [I32, P, P, P, P, I64, I32, I32, P, P]),
        "mpmae_gemm_epi": (C.c_int, [I32, I32, C.POINTER(GemmDesc), P]),
        "mpmae_gemm_wgrad": (C.c_int, [I32, P, P, P, I64, I32, I32, P]),
        "mpmae_gemm_wgrad_act": (C.c_int, [I32, P, P, P, I64, I32, I32, I32, P]),
        "mpmae_raw_transform": (C.c_int, [C.POINTER(RawDesc), P]),
        "mpmae_dense_im2col": (C.c_int, [P, P, I32, I32, I32, I32, I32, I32, I32, I32, P]),
        "mpmae_ln_rows": (C.c_int, [P, P, P, P, I64, I32, C.c_float, I32, P]),
        "mpmae_dense_dwconv": (C.c_int, [P, P, P, P, I32, I32, I32, I32, I32, I32, I32, I32, C.c_float, P]),
        "mpmae_grn_apply": (C.c_int, [P, P, P, P, P, I64, I32, I32, C.c_float, P, P]),
        "mpmae_backward_step": (C.c_int, [P, C.POINTER(IO), I32, P, P, P, P]),
        "mpmae_adamw_step": (C.c_int, [P, P, P, P, P, I64, F, F, F, F, F, I64, F, P]),
        "mpmae_adamw_step_dev": (C.c_int, [P, P, P, P, P, I64, F, F, F, F, F, P, P]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError here = header / library mismatch: fail loudly
        fn.restype, fn.argtypes = res, args
    return lib


lib = _load()
EXPORTS = ["mpmae_last_error", "mpmae_version", "mpmae_plan_create", "mpmae_plan_destroy", "mpmae_param_total",
           "mpmae_param_count", "mpmae_param_info", "mpmae_param_decay", "mpmae_visible_patches",
           "mpmae_workspace_bytes", "mpmae_pred_pixel_cols", "mpmae_pred_image_cols", "mpmae_pred_col_offset",
           "mpmae_tap_info", "mpmae_tap_count", "mpmae_tap_name", "mpmae_launch_count", "mpmae_profile_begin", "mpmae_profile_report", "mpmae_forward",
           "mpmae_forward_encoder", "mpmae_forward_stages", "mpmae_backward", "mpmae_backward_part", "mpmae_backward_part_range", "mpmae_backward_step", "mpmae_encoder_features", "mpmae_gemm_rows", "mpmae_gemm_epi", "mpmae_gemm_wgrad", "mpmae_gemm_wgrad_act", "mpmae_raw_transform", "mpmae_dense_im2col", "mpmae_ln_rows", "mpmae_dense_dwconv", "mpmae_grn_apply", "mpmae_adamw_step", "mpmae_adamw_step_dev"]


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise NativeError(f"{what}: mpmae error {rc}: {lib.mpmae_last_error().decode()}")


class Plan:
    """Owning wrapper of ``mpmae_plan``: parameter layout, workspace layout and launch plan."""

    def __init__(self, cfg: Cfg):
        self.cfg = cfg
        h = C.c_void_p()
        check(lib.mpmae_plan_create(C.byref(cfg), C.byref(h)), "mpmae_plan_create")
        self.handle = h

    def __del__(self):
        h = getattr(self, "handle", None)
        if h and lib is not None:          # lib is None during interpreter shutdown
            lib.mpmae_plan_destroy(h)
            self.handle = None

    def params(self):
        """[(name, shape, offset, decay)] -- reference state-dict names (heads as ``pred_dict.#<i>``)."""
        out = []
        name = C.create_string_buffer(192)
        shape = (C.c_int64 * 4)()
        nd, off = C.c_int32(), C.c_int64()
        for i in range(lib.mpmae_param_count(self.handle)):
            check(lib.mpmae_param_info(self.handle, i, name, 192, shape, C.byref(nd), C.byref(off)), "param_info")
            out.append((name.value.decode(), tuple(shape[: nd.value]), off.value, lib.mpmae_param_decay(self.handle, i)))
        return out

    @property
    def param_total(self) -> int:
        return lib.mpmae_param_total(self.handle)

    @property
    def workspace_bytes(self) -> int:
        return lib.mpmae_workspace_bytes(self.handle)

    @property
    def visible(self) -> int:
        return lib.mpmae_visible_patches(self.handle)

    @property
    def npix(self) -> int:
        return lib.mpmae_pred_pixel_cols(self.handle)

    @property
    def nimg(self) -> int:
        return lib.mpmae_pred_image_cols(self.handle)

    def col_offset(self, mod: int) -> int:
        return lib.mpmae_pred_col_offset(self.handle, mod)

    def tap(self, name: str):
        off, rows, cols = C.c_int64(), C.c_int64(), C.c_int64()
        check(lib.mpmae_tap_info(self.handle, name.encode(), C.byref(off), C.byref(rows), C.byref(cols)), "tap_info")
        return off.value, rows.value, cols.value

    def tap_names(self):
        buf = C.create_string_buffer(128)
        out = []
        for i in range(lib.mpmae_tap_count(self.handle)):
            check(lib.mpmae_tap_name(self.handle, i, buf, 128), "tap_name")
            out.append(buf.value.decode())
        return out

    def profile_begin(self) -> None:
        check(lib.mpmae_profile_begin(self.handle), "profile_begin")

    def profile_report(self):
        """[(name, launches, ms, alg_bytes, alg_flops)] since profile_begin; stops profiling."""
        buf = C.create_string_buffer(1 << 16)
        check(lib.mpmae_profile_report(self.handle, buf, len(buf)), "profile_report")
        rows = []
        for line in buf.value.decode().strip().split("\n")[1:]:
            n, c, ms, b, f = line.split(",")
            rows.append((n, int(c), float(ms), float(b), float(f)))
        return rows

    def backward_ranges(self):
        """[(lo, hi)] float ranges of the flat gradient buffer completed by backward parts 0, 1, 2."""
        out = []
        for part in range(3):
            lo, hi = C.c_int64(), C.c_int64()
            check(lib.mpmae_backward_part_range(self.handle, part, C.byref(lo), C.byref(hi)), "backward_part_range")
            out.append((lo.value, hi.value))
        return out

    def launches(self, backward: bool) -> int:
        return lib.mpmae_launch_count(self.handle, 1 if backward else 0)
