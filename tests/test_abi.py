"""The C-ABI library: builds for sm_100a, loads without a GPU, exports every symbol include/mpmae.h declares,
reports the reference's parameter names, and rejects bad configurations by return code (no compute calls)."""
import ctypes as C
import os
import re

import pytest

from tests import golden_util as gu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    src = open(os.path.join(ROOT, "include", "mpmae.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mpmae_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(native_lib):
    raw = C.CDLL(native_lib.LIB_PATH)
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(raw, n), f"{n} declared in include/mpmae.h but not exported"
    assert sorted(native_lib.EXPORTS) == names, "python binding and header disagree"
    assert native_lib.lib.mpmae_version() >= 100


def _cfg(nat, model_dims=((2, 2, 6, 2), (40, 80, 160, 320)), batch=2, img=56, patch=8, mods=None):
    from mmearth_train_b200.fcmae import modality_kind, N_CLASSES
    full = {"sentinel2": 12, "sentinel1": 8, "aster": 2, "era5": 12, "dynamic_world": 1, "canopy_height_eth": 2, "lat": 2,
            "lon": 2, "biome": 1, "eco_region": 1, "month": 2, "esa_worldcover": 1}
    mods = mods or list(full)
    c = nat.Cfg()
    c.batch, c.img_size, c.patch_size, c.in_chans = batch, img, patch, 12
    for i in range(4):
        c.depths[i], c.dims[i] = model_dims[0][i], model_dims[1][i]
    c.dec_dim, c.dec_depth, c.mask_ratio, c.loss_aggr, c.n_mod = 512, 1, 0.6, 1, len(mods)
    for i, m in enumerate(mods):
        c.mod_kind[i], c.mod_chans[i] = modality_kind(m), N_CLASSES.get(m, full[m])
        c.mod_norm_pix[i] = int(m == "sentinel2")
    return c


def test_plan_layout_matches_reference_state_dict(native_lib):
    nat = native_lib
    z, meta = gu.load("atto_p8_all_unc")
    mods = meta["modalities"]
    plan = nat.Plan(_cfg(nat, mods=mods))
    assert plan.visible == 19                                   # int(49 * 0.4), models/fcmae.py:216-217
    assert plan.npix == 2816 and plan.nimg == 878
    ref_keys = meta["state_keys"]
    total = 0
    for name, shape, off, decay in plan.params():
        if name.startswith("decoder."):
            key = "decoder_dict.sentinel2." + name[len("decoder."):]
        elif name.startswith("pred_dict.#"):
            idx, leaf = name[len("pred_dict.#"):].split(".")
            key = f"pred_dict.{mods[int(idx)]}.{leaf}"
        else:
            key = name
        assert list(shape) == ref_keys[key], (name, shape, ref_keys[key])
        n = 1
        for s in shape:
            n *= s
        total += n
        assert 0 <= off and off + n <= plan.param_total
        is_bias = key.endswith(".bias")
        assert decay == int(len(shape) > 1 and not is_bias)     # timm rule used at main_pretrain.py:312-319
    assert total == 7580674
    assert plan.workspace_bytes > 0 and plan.workspace_bytes % 256 == 0


def test_plan_rejects_bad_configs(native_lib):
    nat = native_lib
    for kw, code in ((dict(patch=12), -1), (dict(patch=32, img=224), -2), (dict(batch=0), -1),
                     (dict(model_dims=((2, 2, 6, 2), (42, 80, 160, 320))), -2)):
        h = C.c_void_p()
        c = _cfg(nat, **kw)
        assert nat.lib.mpmae_plan_create(C.byref(c), C.byref(h)) == code, kw
        assert nat.lib.mpmae_last_error()
    with pytest.raises(nat.NativeError):
        nat.Plan(_cfg(nat, batch=-3))


def test_visible_patch_count_follows_python_int(native_lib):
    nat = native_lib
    for mr in (0.6, 0.75, 0.5, 0.9, 0.25):
        c = _cfg(nat)
        c.mask_ratio = mr
        assert nat.Plan(c).visible == int(49 * (1 - mr)), mr


def test_entry_points_validate_their_arguments_before_touching_the_device(native_lib):
    """Argument errors are reported as negative codes + a message, never as a crash or an exception (no GPU needed: every
    check below fails before the first CUDA call)."""
    nat = native_lib
    plan = nat.Plan(_cfg(nat))
    io = nat.IO()                                            # all pointers null
    st = C.c_void_p(0)
    for fn, args in ((nat.lib.mpmae_forward, (plan.handle, C.byref(io), st)),
                     (nat.lib.mpmae_forward_encoder, (plan.handle, C.byref(io), st)),
                     (nat.lib.mpmae_backward, (plan.handle, C.byref(io), st)),
                     (nat.lib.mpmae_forward_stages, (plan.handle, C.byref(io), nat.STAGE_MASK | nat.STAGE_LOSS, st))):
        assert fn(*args) == -1 and b"null" in nat.lib.mpmae_last_error()
    for stages in (0, -1, 16, 255):                          # not a MPMAE_STAGE_* mask
        assert nat.lib.mpmae_forward_stages(plan.handle, C.byref(io), stages, st) == -1
        assert b"stages" in nat.lib.mpmae_last_error()
    assert nat.lib.mpmae_backward_part(plan.handle, C.byref(io), 3, st) < 0
    assert nat.lib.mpmae_adamw_step(None, None, None, None, None, 16, 1e-3, 0.9, 0.95, 1e-8, 0.05, 1, 1.0, st) == -1
    assert nat.lib.mpmae_adamw_step_dev(None, None, None, None, None, 16, 1e-3, 0.9, 0.95, 1e-8, 0.05, None, st) == -1
    lo, hi = C.c_int64(), C.c_int64()
    assert nat.lib.mpmae_backward_part_range(plan.handle, 5, C.byref(lo), C.byref(hi)) < 0
    assert nat.lib.mpmae_tap_info(plan.handle, b"no.such.tap", C.byref(lo), C.byref(hi), C.byref(C.c_int64())) < 0
    with pytest.raises(nat.NativeError, match="no.such.tap"):
        plan.tap("no.such.tap")


def test_gemm_scratch_size_follows_the_header(native_lib):
    """include/mpmae.h: backends 1 and 3 need scratch of 2*N*ceil32(K) floats (the split weight is padded to 32-column groups:
    [N][ceil(K / 32)][32 bf16 high parts | 32 bf16 remainders] behind the fp32 copy)."""
    hdr = open(os.path.join(ROOT, "include", "mpmae.h")).read()
    assert "2*N*ceil32(K)" in hdr
    for N, K in ((160, 40), (40, 160), (80, 320), (2816, 512), (8, 8), (100, 72)):
        need_fp32 = N * K                                  # Wf
        need_pair = N * ((K + 31) // 32) * 64 // 2         # bf16 pair array, in floats
        got = native_lib.gemm_scratch_floats(N, K)
        assert got == 2 * N * (((K + 31) // 32) * 32)
        assert got >= need_fp32 + need_pair
