"""Host side of the drop-in (mmearth_train_b200/fcmae.py) -- everything that does not need a GPU."""
import pytest
import torch

from oracle import fcmae_oracle as fo
from tests import golden_util as gu


def _model(native_lib, **kw):
    import mmearth_train_b200 as mp
    args = fo.make_args(kw.pop("out_modalities", None), kw.get("loss_aggr", "uncertainty"))
    args.loss_aggr = kw.pop("loss_aggr", "uncertainty")
    lf = mp.UncertaintyWeightingStrategy(len(args.out_modalities)) if args.loss_aggr == "uncertainty" else None
    return mp.convnextv2_atto(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True,
                              patch_size=kw.pop("patch_size", 8), img_size=kw.pop("img_size", 56), args=args, loss_fn=lf)


def test_state_dict_surface_and_aliases(native_lib):
    m = _model(native_lib)
    z, meta = gu.load("atto_p8_all_unc")
    sd = m.state_dict()
    assert {k: list(v.shape) for k, v in sd.items()} == meta["state_keys"]
    # the 12 decoders are one block (models/fcmae.py:119-121,137,145): same storage under every alias
    a = sd["decoder_dict.sentinel2.0.pwconv1.weight"]
    b = sd["decoder_dict.esa_worldcover.0.pwconv1.weight"]
    assert a.data_ptr() == b.data_ptr()
    assert sum(p.numel() for n, p in m.named_parameters() if n != "_ddp_token") == 7580674


def test_parameters_are_views_of_one_flat_buffer(native_lib):
    m = _model(native_lib)
    flat = m.flat_params
    lo, hi = flat.data_ptr(), flat.data_ptr() + flat.numel() * 4
    for n, p in m.named_parameters():
        if n == "_ddp_token":
            continue
        assert lo <= p.data_ptr() < hi, n
    orc = fo.build_oracle()
    fo.init_like_reference(orc, seed=3)
    m.load_state_dict(orc.state_dict())
    off, numel, shape = m._param_slices[5]
    assert torch.equal(flat[off:off + numel].view(shape), m._param_list[5].detach())
    sd = m.state_dict()
    for k, v in orc.state_dict().items():
        assert torch.equal(sd[k], v), k
    # dtype conversion is refused, float32 round trip keeps the views
    m2 = m.to(torch.float32)
    assert m2._param_list[0].data_ptr() == m2.flat_params.data_ptr()
    with pytest.raises(TypeError):
        m.half()


def test_weight_decay_mask_follows_timm_rule(native_lib):
    m = _model(native_lib)
    mask = m.decay_mask()
    for (name, shape, off, decay), (_o, numel, _s) in zip(m._layout, m._param_slices):
        expect = int(len(shape) > 1 and not name.endswith(".bias"))
        assert int(mask[off]) == expect and int(mask[off + numel - 1]) == expect, name
    # ME biases are [1, C] but named *.bias -> no decay; GRN gamma/beta [1, 4C] ARE decayed
    names = {n: d for n, _s, _o, d in m._layout}
    assert names["encoder.stages.0.0.dwconv.bias"] == 0 and names["encoder.stages.0.0.grn.gamma"] == 1


def test_no_cpu_fallback(native_lib):
    m = _model(native_lib)
    batch = fo.synthetic_batch(2, 56)
    with pytest.raises(RuntimeError, match="CUDA"):
        m(batch, mask_ratio=0.6)


def test_constructor_errors_mirror_reference_scope(native_lib):
    import mmearth_train_b200 as mp
    args = fo.make_args()
    with pytest.raises(NotImplementedError):
        mp.convnextv2_atto(args=args, loss_fn=mp.UncertaintyWeightingStrategy(12), sparse=False, img_size=56, patch_size=8)
    with pytest.raises(ValueError):
        mp.convnextv2_atto(args=args, loss_fn=None, img_size=56, patch_size=8)


def test_gen_random_mask_and_patchify_match_oracle(native_lib):
    m = _model(native_lib)
    x = torch.randn(3, 12, 56, 56)
    torch.manual_seed(7)
    mask = m.gen_random_mask(x, 0.6)
    torch.manual_seed(7)
    noise = torch.randn(3, 49)
    assert torch.equal(mask, fo.OracleFCMAE.mask_from_noise(noise, 0.6))
    assert mask.sum(1).tolist() == [30.0] * 3
    orc = fo.build_oracle()
    t = torch.randn(2, 8, 56, 56)
    assert torch.equal(m.patchify(t, "sentinel1"), orc.patchify(t, 8))


def test_product_synthetic_generator_equals_the_oracles_and_the_reference_modalities():
    """bench.py and the tools draw their batches / args from mmearth_train_b200.synthetic (the product never imports the
    oracle); same seed -> bit-identical tensors, same args fields.  With /root/reference present the band tables are also
    checked against MODALITIES.py."""
    from mmearth_train_b200 import synthetic as syn
    for outs, nan_frac in ((None, 0.05), (["sentinel2"], 0.0), (["era5", "biome", "lat"], 0.1)):
        a = syn.synthetic_batch(3, 56, outs, seed=77, nan_frac=nan_frac)
        b = fo.synthetic_batch(3, 56, outs, seed=77, nan_frac=nan_frac)
        assert list(a) == list(b)
        for k in a:
            assert a[k].dtype == b[k].dtype and torch.equal(torch.nan_to_num(a[k].double(), nan=-7.0),
                                                            torch.nan_to_num(b[k].double(), nan=-7.0)), k
        assert vars(syn.make_args(outs, "unweighted")) == vars(fo.make_args(outs, "unweighted"))
    from oracle import ref_harness
    if ref_harness.reference_available():
        M = ref_harness.load_reference().MODALITIES
        assert list(M.OUT_MODALITIES) == syn.ALL_OUT
        assert M.INP_MODALITIES["sentinel2"] == syn.S2_BANDS
        full = {k: len(v) for k, v in M.MODALITIES_FULL.items()}
        assert {k: full[k] for k in syn.FULL_BANDS} == syn.FULL_BANDS      # (the reference also lists three S2 mask products)


@pytest.mark.parametrize("name,patch,img", [("convnextv2_femto", 8, 56), ("convnextv2_pico", 8, 56), ("convnextv2_nano", 16, 112),
                                            ("convnextv2_tiny", 8, 56), ("convnextv2_base", 8, 56)])
def test_every_supported_factory_has_the_oracles_state_dict_and_decay_rule(native_lib, name, patch, img):
    """Plan creation is host-only: parameter names / shapes of every factory the kernels accept equal the oracle's state dict
    (whose surface is pinned to the reference's by the fixtures), and the per-parameter decay flag is timm's rule
    (ndim <= 1 or *.bias -> no decay; main_pretrain.py:312-319)."""
    import mmearth_train_b200 as mp
    args = fo.make_args(None, "uncertainty")
    m = getattr(mp, name)(mask_ratio=0.6, decoder_depth=1, decoder_embed_dim=512, norm_pix_loss=True, patch_size=patch,
                          img_size=img, args=args, loss_fn=mp.UncertaintyWeightingStrategy(12))
    orc = fo.build_oracle(model=name, img_size=img, patch_size=patch)
    want = {k: tuple(v.shape) for k, v in orc.state_dict().items()}
    got = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    assert got == want
    mask = m.decay_mask()
    unique = {}
    for n, p in m.named_parameters():
        if n != "_ddp_token":
            unique.setdefault(p.data_ptr(), (n, p))
    for n, p in unique.values():
        off = (p.data_ptr() - m.flat_params.data_ptr()) // 4
        decayed = bool(mask[off]) and bool(mask[off + p.numel() - 1])
        assert decayed == (not (p.ndim <= 1 or n.endswith(".bias"))), n
    assert sum(p.numel() for _, p in unique.values()) == sum(p.numel() for p in {id(q): q for q in orc.parameters()}.values())
    plan = m._plan(64)
    assert plan.workspace_bytes > 0 and plan.launches(False) == 0        # launch counts are filled in by the first call
