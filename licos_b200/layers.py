"""Host-side mirrors of compressai.layers.GDN / compressai.ops and the
fused layer runner that turns an ``nn.Sequential`` of conv / GDN / ReLU modules into C-ABI launches.

Module, parameter and buffer names follow CompressAI so that ``state_dict()`` keys are interchangeable
(``g_a.1.beta``, ``g_a.1.beta_reparam.lower_bound.bound``, ...; /root/reference/licos/federation_utils.py:47-53
does arithmetic on every key, eval_script.py:70-71 loads them).
"""
from __future__ import annotations

import weakref
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib, ops


class _LowerBoundFn(torch.autograd.Function):
    """max(x, bound) whose gradient also passes when it would move x back above the bound."""

    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, g):
        x, bound = ctx.saved_tensors
        keep = (x >= bound) | (g < 0)
        return keep.type(g.dtype) * g, None


class LowerBound(nn.Module):
    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x: Tensor) -> Tensor:
        return _LowerBoundFn.apply(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    pedestal: Tensor

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        self.register_buffer("pedestal", torch.Tensor([self.reparam_offset ** 2]))
        self.lower_bound = LowerBound((self.minimum + self.reparam_offset ** 2) ** 0.5)

    def init(self, x: Tensor) -> Tensor:
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x: Tensor) -> Tensor:
        return self.lower_bound(x) ** 2 - self.pedestal


class GDN(nn.Module):
    """Generalised divisive normalisation.  Inside a :class:`FusedSequential` (no-grad) it never runs as a
    module: it becomes the epilogue of the preceding conv kernel."""

    def __init__(self, in_channels: int, inverse: bool = False, beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=float(beta_min))
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(in_channels)))
        self.gamma_reparam = NonNegativeParametrizer()
        self.gamma = nn.Parameter(self.gamma_reparam.init(float(gamma_init) * torch.eye(in_channels)))

    def forward(self, x: Tensor) -> Tensor:
        # differentiable torch path (training); the inference path is the fused kernel epilogue
        C = x.size(1)
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma).reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


def conv(in_channels: int, out_channels: int, kernel_size: int = 5, stride: int = 2) -> nn.Conv2d:
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(in_channels: int, out_channels: int, kernel_size: int = 5, stride: int = 2) -> nn.ConvTranspose2d:
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


# ---------------------------------------------------------------------------------------------
# fused runner
# ---------------------------------------------------------------------------------------------

def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


def _conv_kind(m: nn.Module) -> int:
    if isinstance(m, nn.ConvTranspose2d):
        if (_pair(m.kernel_size), _pair(m.stride), _pair(m.padding), _pair(m.output_padding), _pair(m.dilation),
                m.groups) == ((5, 5), (2, 2), (2, 2), (1, 1), (1, 1), 1):
            return _lib.DECONV_5X5_S2
    elif isinstance(m, nn.Conv2d):
        key = (_pair(m.kernel_size), _pair(m.stride), _pair(m.padding), _pair(m.dilation), m.groups, m.padding_mode)
        if key == ((5, 5), (2, 2), (2, 2), (1, 1), 1, "zeros"):
            return _lib.CONV_5X5_S2
        if key == ((3, 3), (1, 1), (1, 1), (1, 1), 1, "zeros"):
            return _lib.CONV_3X3_S1
    raise NotImplementedError(f"licos_b200 has no kernel for layer {m!r}")


class _Packed:
    __slots__ = ("key", "tensors")

    def __init__(self):
        self.key, self.tensors = None, None


def _version_key(*params: Optional[Tensor]):
    return tuple((p.data_ptr(), p._version, tuple(p.shape)) if p is not None else None for p in params)


class FusedSequential(nn.Sequential):
    """``nn.Sequential`` with CompressAI's indexing / item-assignment surface (LICOS swaps ``g_a[0]`` and
    ``g_s[6]``, /root/reference/licos/model_utils.py:31-45).  With autograd off it runs as fused sm_100a
    kernels (one launch per conv layer, GDN/IGDN/ReLU in the epilogue); with autograd on it is the ordinary
    differentiable module chain."""

    def __init__(self, *args):
        super().__init__(*args)
        self._packed_cache = weakref.WeakKeyDictionary()

    def _eager_forward(self, x: Tensor) -> Tensor:
        return super().forward(x)

    def forward(self, x: Tensor, take_abs: bool = False, nhwc: Optional[Tensor] = None) -> Tensor:
        """``nhwc``: optional bf16 (B, H, W, C) copy of ``x`` that a fused producer already wrote
        (``EntropyBottleneck.forward_fused``); saves the layout-conversion launch on the fused path."""
        if not x.is_cuda:
            raise RuntimeError("licos_b200: g_a / g_s / h_a / h_s need CUDA tensors on a B200 (no CPU path exists)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            return self._eager_forward(torch.abs(x) if take_abs else x)
        return self.fused_forward(x, take_abs=take_abs, nhwc=nhwc)

    # -- caches of kernel-layout parameters, rebuilt when a parameter's version or storage changes --
    def _packed_weight(self, m: nn.Module, kind: int, in_layout: int):
        ent = self._packed_cache.setdefault(m, {})
        slot = ent.setdefault(("w", in_layout), _Packed())
        key = _version_key(m.weight, m.bias)
        if slot.key != key:
            out_c = m.out_channels
            in_c = m.in_channels
            w = m.weight.detach()
            packed = ops.pack_conv_weight(w.contiguous(), kind, out_c, in_c, in_layout)
            bias = None if m.bias is None else m.bias.detach().contiguous()
            slot.key, slot.tensors = key, (packed, bias)
        return slot.tensors

    def _packed_gdn(self, g: GDN):
        ent = self._packed_cache.setdefault(g, {})
        slot = ent.setdefault("gdn", _Packed())
        key = _version_key(g.beta, g.gamma)
        if slot.key != key:
            slot.key = key
            slot.tensors = ops.gdn_pack(
                g.beta.detach().contiguous(), g.gamma.detach().contiguous(),
                float(g.beta_reparam.lower_bound.bound), float(g.gamma_reparam.lower_bound.bound),
                float(g.beta_reparam.pedestal))
        return slot.tensors

    def fused_forward(self, x: Tensor, take_abs: bool = False, nhwc: Optional[Tensor] = None) -> Tensor:
        """x: fp32 (B, C, H, W) on a B200.  Returns fp32 NCHW like the module chain would."""
        if not x.is_cuda:
            raise RuntimeError("licos_b200: the fused path needs CUDA tensors (no CPU fallback exists)")
        mods = list(self)
        steps = []
        i = 0
        while i < len(mods):
            m = mods[i]
            kind = _conv_kind(m)
            epi, gdn = _lib.EPI_NONE, None
            if i + 1 < len(mods):
                nxt = mods[i + 1]
                if isinstance(nxt, GDN):
                    epi, gdn = (_lib.EPI_IGDN if nxt.inverse else _lib.EPI_GDN), nxt
                    i += 1
                elif isinstance(nxt, nn.ReLU):
                    epi = _lib.EPI_RELU
                    i += 1
            steps.append((m, kind, epi, gdn))
            i += 1
        if not steps:
            return x

        cur = x.contiguous()
        layout = _lib.LAYOUT_NCHW_F32
        first = steps[0][0]
        direct_first = (steps[0][1] == _lib.CONV_5X5_S2 and first.in_channels <= 16 and not take_abs)
        if not direct_first:
            if nhwc is not None and not take_abs:
                if nhwc.dtype != torch.bfloat16 or tuple(nhwc.shape) != (x.shape[0], x.shape[2], x.shape[3], x.shape[1]):
                    raise ValueError("nhwc must be the bf16 (B, H, W, C) copy of x")
                cur = nhwc.contiguous()
            else:
                cur = ops.nchw_to_nhwc_bf16(cur, take_abs=take_abs)
            layout = _lib.LAYOUT_NHWC_BF16
        for n, (m, kind, epi, gdn) in enumerate(steps):
            last = n == len(steps) - 1
            out_layout = _lib.LAYOUT_NCHW_F32 if last else _lib.LAYOUT_NHWC_BF16
            packed, bias = self._packed_weight(m, kind, layout)
            beta = gamma = None
            if gdn is not None:
                if gdn.beta.numel() != m.out_channels:
                    raise ValueError("GDN channel count does not match the preceding convolution")
                beta, gamma = self._packed_gdn(gdn)
                if bias is None:
                    bias = torch.zeros(m.out_channels, dtype=torch.float32, device=cur.device)
            cur = ops.conv_forward(cur, kind=kind, epilogue=epi, in_layout=layout, out_layout=out_layout,
                                   in_c=m.in_channels, out_c=m.out_channels, weight=packed, bias=bias,
                                   beta=beta, gamma=gamma)
            layout = out_layout
        return cur
