"""Drop-in for the reference's dense ``models.convnextv2.ConvNeXtV2`` (``models/convnextv2.py:59-207``): the network that
finetuning and linear probing run on a pretrained encoder (``hubconf.py:77-93``), INFERENCE forward on the B200.

Same constructor arguments, state-dict keys and shapes as the reference (so ``checkpoint.to_dense_state_dict`` of a
pretraining checkpoint, the reference's ``remap_checkpoint_keys``, loads with ``load_state_dict``), same ``forward_features``
/ ``forward``.  The geometry is NOT the masked encoder's: the dense stem convolution is un-padded (56 -> 54 -> 27 -> 13 -> 6)
and GRN is per sample (``models/norm_layers.py:33-44``), so activations are plain channels-last rows ``[B*H*W, C]`` and every
layer is a native call through the C ABI: un-padded 3x3 and 2x2/s2 convolutions as im2col + tcgen05 GEMM, pointwise convolutions
as tcgen05 GEMMs with the fused GELU / sum h^2 epilogue, depthwise 7x7 + LayerNorm and the GRN apply as dense row kernels
(``csrc/dense_ops.cuh``).  LayerNorm affines in front of a linear map are folded into its weights.  Gradients are not
implemented (``forward`` runs under ``no_grad``; training the dense network is outside the pretraining hot path, SURVEY.md
section 8 f4); the pooled ``[B, C]`` tail (final LayerNorm + linear head) is three tiny torch calls.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _native as nat


def _p(t: Optional[torch.Tensor]):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class ConvNeXtV2(nn.Module):
    def __init__(self, patch_size: int = 8, img_size: int = 56, in_chans: int = 12, num_classes: int = 1000,
                 depths: List[int] = None, dims: List[int] = None, drop_path_rate: float = 0.0, head_init_scale: float = 1.0,
                 use_orig_stem: bool = False, args=None, gemm_backend: int = 3):
        super().__init__()
        if use_orig_stem:
            raise NotImplementedError("use_orig_stem=True is not on the pretraining path (convnextv2.py:99-111)")
        self.depths = list(depths) if depths is not None else [3, 3, 9, 3]
        dims = list(dims) if dims is not None else [96, 192, 384, 768]
        if any(d % 8 != 0 for d in dims) or dims[3] * 4 > 8192:
            raise NotImplementedError(f"dims {dims}: channel counts must be multiples of 8 (tcgen05 GEMM K granularity)")
        self.dims, self.img_size, self.patch_size, self.in_chans = dims, img_size, patch_size, in_chans
        self.num_classes, self.gemm_backend = num_classes, gemm_backend
        k = patch_size // 8
        tn = lambda *shape: nn.Parameter(torch.nn.init.trunc_normal_(torch.empty(*shape), std=0.02))   # convnextv2.py:154-158
        zeros, ones = (lambda n: nn.Parameter(torch.zeros(n))), (lambda n: nn.Parameter(torch.ones(n)))

        def node(**params):
            m = nn.Module()
            for name, p in params.items():
                m.register_parameter(name, p)
            return m

        self.initial_conv = nn.ModuleList([node(weight=tn(dims[0], in_chans, 3, 3), bias=zeros(dims[0])),
                                           node(weight=ones(dims[0]), bias=zeros(dims[0]))])
        self.stem = nn.ModuleList([node(weight=tn(dims[0], 1, k, k), bias=zeros(dims[0])), node(weight=ones(dims[0]), bias=zeros(dims[0]))])
        self.downsample_layers = nn.ModuleList([
            nn.ModuleList([node(weight=ones(dims[i]), bias=zeros(dims[i])), node(weight=tn(dims[i + 1], dims[i], 2, 2), bias=zeros(dims[i + 1]))])
            for i in range(3)])
        self.stages = nn.ModuleList()
        for i in range(4):
            Cc = dims[i]
            blocks = nn.ModuleList()
            for _ in range(self.depths[i]):
                b = nn.Module()
                b.dwconv = node(weight=tn(Cc, 1, 7, 7), bias=zeros(Cc))
                b.norm = node(weight=ones(Cc), bias=zeros(Cc))
                b.pwconv1 = node(weight=tn(4 * Cc, Cc), bias=zeros(4 * Cc))
                b.grn = node(gamma=nn.Parameter(torch.zeros(1, 1, 1, 4 * Cc)), beta=nn.Parameter(torch.zeros(1, 1, 1, 4 * Cc)))
                b.pwconv2 = node(weight=tn(Cc, 4 * Cc), bias=zeros(Cc))
                blocks.append(b)
            self.stages.append(blocks)
        self.norm = nn.LayerNorm(dims[-1], eps=1e-6)
        self.head = nn.Linear(dims[-1], num_classes)
        with torch.no_grad():
            torch.nn.init.trunc_normal_(self.head.weight, std=0.02)
            self.head.bias.zero_()
            self.head.weight.mul_(head_init_scale)
            self.head.bias.mul_(head_init_scale)
        self._folded = None       # weights as the kernels consume them, rebuilt when a parameter changes

    # ------------------------------------------------------------------ parameter preparation
    def _signature(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def _prepare(self):
        sig = self._signature()
        if self._folded is not None and self._folded["sig"] == sig:
            return self._folded
        f = {"sig": sig, "blocks": [], "down": []}
        w = self.initial_conv[0].weight
        kk = w.shape[1] * 9
        f["ic_kpad"] = (kk + 7) // 8 * 8
        f["ic_w"] = F.pad(w.reshape(w.shape[0], kk), (0, f["ic_kpad"] - kk)).contiguous()
        for i in range(3):                       # LayerNorm affine folded into the 2x2 / stride-2 convolution that follows
            ln, conv = self.downsample_layers[i]
            wf = (conv.weight * ln.weight[None, :, None, None]).reshape(conv.weight.shape[0], -1).contiguous()
            bf = conv.bias + (conv.weight * ln.bias[None, :, None, None]).sum(dim=(1, 2, 3))
            f["down"].append((wf, bf.contiguous()))
        for i in range(4):
            for b in self.stages[i]:             # LayerNorm affine folded into pwconv1
                w1 = (b.pwconv1.weight * b.norm.weight[None, :]).contiguous()
                b1 = (b.pwconv1.bias + b.pwconv1.weight @ b.norm.bias).contiguous()
                f["blocks"].append((w1, b1, b.grn.gamma.reshape(-1).contiguous(), b.grn.beta.reshape(-1).contiguous()))
        self._folded = f
        return f

    # ------------------------------------------------------------------ native calls
    def _gemm(self, mode, a, w, bias, out, resid=None, out2=None, colsum=None, group_rows=0):
        d = nat.GemmDesc()
        M, K = a.shape
        N = w.shape[0]
        scratch = torch.empty(nat.gemm_scratch_floats(N, K), device=a.device)
        for name, t in dict(a=a, b=w, bias=bias, resid=resid, out=out, out2=out2, colsum=colsum, scratch=scratch).items():
            if t is not None:
                setattr(d, name, t.data_ptr())
        d.M, d.N, d.K, d.group_rows = M, N, K, group_rows
        backend = self.gemm_backend if M >= 64 else 0          # tiny products stay on the fp32 SIMT tiles
        if mode == 1 and 0 < group_rows < 32:
            backend = 0
        stream = torch.cuda.current_stream(a.device).cuda_stream
        nat.check(nat.lib.mpmae_gemm_epi(mode, backend, C.byref(d), C.c_void_p(stream)), "mpmae_gemm_epi")
        return out

    @torch.no_grad()
    def _feature_map(self, x: torch.Tensor):
        """[B, in, H, W] -> (rows [B*Hf*Wf, C3] channels-last, Hf, Wf)."""
        if not x.is_cuda:
            raise RuntimeError("the native dense ConvNeXtV2 runs on CUDA (sm_100a) tensors only; there is no CPU path")
        if x.device != self.head.weight.device:
            raise RuntimeError("model and input are on different devices")
        f = self._prepare()
        dev, dims, k = x.device, self.dims, self.patch_size // 8
        B, Cin, H, W = x.shape
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        new = lambda *s: torch.empty(*s, device=dev)
        with torch.cuda.device(dev):
            x = x.float().contiguous()
            # initial conv: 3x3, stride 1, NO padding (convnextv2.py:113-117) = im2col + GEMM, then LayerNorm + GELU
            H1, W1 = H - 2, W - 2
            cols = new(B * H1 * W1, f["ic_kpad"])
            nat.check(nat.lib.mpmae_dense_im2col(_p(x), _p(cols), B, Cin, H, W, 3, 1, f["ic_kpad"], 1, st), "dense_im2col")
            y = self._gemm(0, cols, f["ic_w"], self.initial_conv[0].bias, new(B * H1 * W1, dims[0]))
            t = new(B * H1 * W1, dims[0])
            nat.check(nat.lib.mpmae_ln_rows(_p(y), _p(self.initial_conv[1].weight), _p(self.initial_conv[1].bias), _p(t),
                                            B * H1 * W1, dims[0], 1e-6, 1, st), "ln_rows")
            # stem: depthwise k x k, stride k, padding k // 2 (convnextv2.py:119-130), then LayerNorm
            Hs, Ws = (H1 + 2 * (k // 2) - k) // k + 1, (W1 + 2 * (k // 2) - k) // k + 1
            y = new(B * Hs * Ws, dims[0])
            nat.check(nat.lib.mpmae_dense_dwconv(_p(t), _p(self.stem[0].weight), _p(self.stem[0].bias), _p(y), B, H1, W1, dims[0], k, k,
                                                 k // 2, 0, 0.0, st), "dense_dwconv")
            cur = new(B * Hs * Ws, dims[0])
            nat.check(nat.lib.mpmae_ln_rows(_p(y), _p(self.stem[1].weight), _p(self.stem[1].bias), _p(cur), B * Hs * Ws, dims[0],
                                            1e-6, 0, st), "ln_rows")
            Hc, Wc, bi = Hs, Ws, 0
            for i in range(4):
                Cc = dims[i]
                if i > 0:                          # LayerNorm (affine folded) + 2x2 / stride-2 convolution (floor)
                    Cp = dims[i - 1]
                    xh = new(B * Hc * Wc, Cp)
                    nat.check(nat.lib.mpmae_ln_rows(_p(cur), None, None, _p(xh), B * Hc * Wc, Cp, 1e-6, 0, st), "ln_rows")
                    Hn, Wn = Hc // 2, Wc // 2
                    cols = new(B * Hn * Wn, 4 * Cp)
                    nat.check(nat.lib.mpmae_dense_im2col(_p(xh), _p(cols), B, Cp, Hc, Wc, 2, 2, 4 * Cp, 0, st), "dense_im2col")
                    wf, bf = f["down"][i - 1]
                    cur = self._gemm(0, cols, wf, bf, new(B * Hn * Wn, Cc))
                    Hc, Wc = Hn, Wn
                R = B * Hc * Wc
                for b in self.stages[i]:           # block, convnextv2.py:42-55
                    w1, b1, gamma, beta = f["blocks"][bi]
                    bi += 1
                    vhat = new(R, Cc)
                    nat.check(nat.lib.mpmae_dense_dwconv(_p(cur), _p(b.dwconv.weight), _p(b.dwconv.bias), _p(vhat), B, Hc, Wc, Cc, 7, 1,
                                                         3, 1, 1e-6, st), "dense_dwconv")
                    a, h, gsq = new(R, 4 * Cc), new(R, 4 * Cc), torch.zeros(B, 4 * Cc, device=dev)
                    self._gemm(1, vhat, w1, b1, a, out2=h, colsum=gsq, group_rows=Hc * Wc)
                    g = a                                                   # the pre-activation is not needed at inference
                    scratch = new(2 * B * 4 * Cc + B)
                    nat.check(nat.lib.mpmae_grn_apply(_p(h), _p(gsq), _p(gamma), _p(beta), _p(g), R, 4 * Cc, Hc * Wc, 1e-4, _p(scratch),
                                                      st), "grn_apply")
                    cur = self._gemm(0, g, b.pwconv2.weight, b.pwconv2.bias, new(R, Cc), resid=cur)
        return cur, Hc, Wc

    # ------------------------------------------------------------------ reference surface
    @torch.no_grad()
    def forward_features(self, x: torch.Tensor) -> torch.Tensor:
        """``models/convnextv2.py:160-172``: global average pooling + final LayerNorm -> ``[B, dims[-1]]``."""
        rows, Hf, Wf = self._feature_map(x)
        return self.norm(rows.view(x.shape[0], Hf * Wf, -1).mean(1))

    @torch.no_grad()
    def feature_map(self, x: torch.Tensor) -> torch.Tensor:
        """The last stage's feature map ``[B, dims[-1], Hf, Wf]`` (what a dense-prediction head would consume)."""
        rows, Hf, Wf = self._feature_map(x)
        return rows.view(x.shape[0], Hf, Wf, -1).permute(0, 3, 1, 2)

    def forward(self, x: torch.Tensor, mask: torch.Tensor = None) -> torch.Tensor:
        """``models/convnextv2.py:183-207`` without a mask: logits.  (With a mask the reference runs its dense CPU stand-in
        for the sparse encoder; the masked encoder here is ``mmearth_train_b200.FCMAE``.)"""
        if mask is not None:
            raise NotImplementedError("masked forward: use mmearth_train_b200.FCMAE (the sparse encoder)")
        return self.head(self.forward_features(x))

    def upsample_mask(self, mask: torch.Tensor, scale: int) -> torch.Tensor:
        """``models/convnextv2.py:174-181``."""
        assert len(mask.shape) == 2
        p = int(mask.shape[1] ** 0.5)
        return mask.reshape(-1, p, p).repeat_interleave(scale, dim=1).repeat_interleave(scale, dim=2)


# ---------------------------------------------------------------------- factories, models/convnextv2.py:210-247
def convnextv2_atto(**kwargs):
    return ConvNeXtV2(depths=[2, 2, 6, 2], dims=[40, 80, 160, 320], **kwargs)


def convnextv2_femto(**kwargs):
    return ConvNeXtV2(depths=[2, 2, 6, 2], dims=[48, 96, 192, 384], **kwargs)


def convnext_pico(**kwargs):
    return ConvNeXtV2(depths=[2, 2, 6, 2], dims=[64, 128, 256, 512], **kwargs)


def convnextv2_nano(**kwargs):
    return ConvNeXtV2(depths=[2, 2, 8, 2], dims=[80, 160, 320, 640], **kwargs)


def convnextv2_tiny(**kwargs):
    return ConvNeXtV2(depths=[3, 3, 9, 3], dims=[96, 192, 384, 768], **kwargs)


def convnextv2_base(**kwargs):
    return ConvNeXtV2(depths=[3, 3, 27, 3], dims=[128, 256, 512, 1024], **kwargs)
