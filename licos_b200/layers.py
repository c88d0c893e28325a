"""Host-side mirrors of compressai.layers.GDN / compressai.ops and the
fused layer runner that turns an ``nn.Sequential`` of conv / GDN / ReLU modules into C-ABI launches.

Module, parameter and buffer names follow CompressAI so that ``state_dict()`` keys are interchangeable
(``g_a.1.beta``, ``g_a.1.beta_reparam.lower_bound.bound``, ...; /root/reference/licos/federation_utils.py:47-53
does arithmetic on every key, eval_script.py:70-71 loads them).
"""
from __future__ import annotations

import os
import weakref
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib, ops
from . import torch_ops as T


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
        # the same two constants as host floats (fp32-rounded like the buffers): reading a device buffer would be a
        # host sync in every training step, and is illegal under CUDA-graph capture
        self.pedestal_f = float(torch.tensor(self.reparam_offset ** 2, dtype=torch.float32))
        self.bound_f = float(torch.tensor((self.minimum + self.reparam_offset ** 2) ** 0.5, dtype=torch.float32))

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


_UNFUSED_GDN_BWD = bool(int(os.environ.get("LICOS_UNFUSED_GDN_BWD", "0")))  # development: GDN backward as separate passes


class _Packed:
    __slots__ = ("key", "tensors")

    def __init__(self):
        self.key, self.tensors = None, None


def _version_key(*params: Optional[Tensor]):
    return tuple((p.data_ptr(), p._version, tuple(p.shape)) if p is not None else None for p in params)


def _capturing() -> bool:
    """Under CUDA-graph capture every tensor derived from a parameter is rebuilt inside the graph (the optimizer step
    of a replay changes the parameters without touching their version counters)."""
    return ops.capture_rebuild()


class FusedSequential(nn.Sequential):
    """``nn.Sequential`` with CompressAI's indexing / item-assignment surface (LICOS swaps ``g_a[0]`` and
    ``g_s[6]``, /root/reference/licos/model_utils.py:31-45).  With autograd off it runs as fused sm_100a
    kernels (one launch per conv layer, GDN/IGDN/ReLU in the epilogue); with autograd on it is the ordinary
    differentiable module chain."""

    def __init__(self, *args):
        super().__init__(*args)
        self._packed_cache = weakref.WeakKeyDictionary()

    def __getstate__(self):
        # kernel-layout weight copies are a cache keyed on live modules: never part of a pickle (torch.save(net))
        state = self.__dict__.copy()
        state.pop("_packed_cache", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        self._packed_cache = weakref.WeakKeyDictionary()

    def __deepcopy__(self, memo):
        import copy

        new = self.__class__.__new__(self.__class__)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            if k != "_packed_cache":
                new.__dict__[k] = copy.deepcopy(v, memo)
        new._packed_cache = weakref.WeakKeyDictionary()
        return new

    def forward(self, x: Tensor, take_abs: bool = False, nhwc: Optional[Tensor] = None, *, int_max: int = 0,
                requant8: bool = False, out_dtype: Optional[torch.dtype] = None, out_max: int = 0) -> Tensor:
        """``nhwc``: optional bf16 (B, H, W, C) copy of ``x`` that a fused producer already wrote
        (``EntropyBottleneck.forward_fused``); saves the layout-conversion launch on the fused path.

        Integer tiles (inference only): ``x`` may be uint8, or 12-bit digital numbers in uint16 / int16 storage; the first
        layer scales them by ``1 / int_max`` (default 255 / 4095; ``requant8`` adds the reference's 8-bit step,
        raw_image_folder.py:192-196) while it builds its patches -- bit-identical to feeding the fp32 tensor.
        ``out_dtype`` = torch.uint8 / torch.uint16 makes the last layer write ``round(clamp(x_hat, 0, 1) * out_max)``."""
        if not x.is_cuda:
            raise RuntimeError("licos_b200: g_a / g_s / h_a / h_s need CUDA tensors on a B200 (no CPU path exists)")
        if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            if x.dtype != torch.float32 or out_dtype not in (None, torch.float32):
                raise NotImplementedError("licos_b200: integer tiles are an inference path; train on float32 tensors")
            xin = torch.abs(x) if take_abs else x
            if not self._native_backward_ok(xin):
                # no second backend: a shape the backward kernels do not take is an error, not a detour through cuDNN
                raise NotImplementedError(
                    f"licos_b200: no native backward for input shape {tuple(xin.shape)} through {self.__class__.__name__}: "
                    "every stride-2 layer needs even sizes (the reference trains on 256x256 patches, "
                    "cfg/default_cfg.toml:29) and hidden channel counts that are multiples of 64")
            return self.train_forward(xin)
        return self.fused_forward(x, take_abs=take_abs, nhwc=nhwc, int_max=int_max, requant8=requant8,
                                  out_dtype=out_dtype, out_max=out_max)

    # -- caches of kernel-layout parameters, rebuilt when a parameter's version or storage changes --
    def train(self, mode: bool = True):
        # Entering or leaving training invalidates every kernel-layout copy of the parameters: fused / foreach
        # optimizers update parameters in place WITHOUT bumping the version counters the caches are keyed on.
        self._packed_cache.clear()
        return super().train(mode)

    def _packed_weight(self, m: nn.Module, kind: int, in_layout: int, force: bool = False):
        ent = self._packed_cache.setdefault(m, {})
        slot = ent.setdefault(("w", in_layout), _Packed())
        key = _version_key(m.weight, m.bias)
        if slot.key != key or force or _capturing():
            out_c = m.out_channels
            in_c = m.in_channels
            w = m.weight.detach()
            packed = ops.pack_conv_weight(w.contiguous(), kind, out_c, in_c, in_layout)
            bias = None if m.bias is None else m.bias.detach().contiguous()
            slot.key, slot.tensors = key, (packed, bias)
        return slot.tensors

    def _packed_gdn(self, g: GDN, force: bool = False):
        ent = self._packed_cache.setdefault(g, {})
        slot = ent.setdefault("gdn", _Packed())
        key = _version_key(g.beta, g.gamma)
        if slot.key != key or force or _capturing():
            slot.key = key
            slot.tensors = ops.gdn_pack(
                g.beta.detach().contiguous(), g.gamma.detach().contiguous(),
                g.beta_reparam.bound_f, g.gamma_reparam.bound_f, g.beta_reparam.pedestal_f)
        return slot.tensors

    def _cached(self, owner: nn.Module, name, params, build, force: bool = False):
        ent = self._packed_cache.setdefault(owner, {})
        slot = ent.setdefault(name, _Packed())
        key = _version_key(*params)
        if slot.key != key or force or _capturing():
            slot.key, slot.tensors = key, build()
        return slot.tensors

    def _steps(self):
        """The chain as (conv module, kernel kind, fused epilogue, GDN module | None) launches: a GDN / IGDN / ReLU that
        follows a conv becomes that conv's epilogue."""
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
        return steps

    def _native_backward_ok(self, x: Tensor) -> bool:
        """The native backward covers the shapes the training loop uses: every stride-2 layer sees even sizes (so that
        the data gradient of a conv is exactly the transposed conv) and channel counts the kernels take."""
        if x.dim() != 4 or x.dtype != torch.float32 or len(self) == 0:
            return False
        try:
            steps = self._steps()
        except NotImplementedError:
            return False
        h, w = x.shape[2], x.shape[3]
        for n, (m, kind, epi, gdn) in enumerate(steps):
            first_direct = n == 0 and kind == _lib.CONV_5X5_S2 and m.in_channels <= 16
            narrow_last = n == len(steps) - 1 and kind == _lib.DECONV_5X5_S2 and m.out_channels <= 4
            if not first_direct and m.in_channels % 64 != 0:
                return False
            if not narrow_last and m.out_channels % 64 != 0:
                return False
            if narrow_last and (epi != _lib.EPI_NONE or m.in_channels > 256):
                return False
            if first_direct and m.out_channels not in (64, 128, 192):
                return False
            if m.in_channels > 512 or m.out_channels > 512:
                return False
            if kind == _lib.CONV_5X5_S2:
                if h % 2 or w % 2:
                    return False
                h, w = h // 2, w // 2
            elif kind == _lib.DECONV_5X5_S2:
                h, w = 2 * h, 2 * w
        return True

    def train_forward(self, x: Tensor) -> Tensor:
        """Differentiable forward on the sm_100a kernels: forward, data gradients and weight gradients all native."""
        steps = self._steps()
        params = []
        for m, kind, epi, gdn in steps:
            params += [m.weight, m.bias]
            if gdn is not None:
                params += [gdn.beta, gdn.gamma]
        return _ChainFn.apply(self, steps, x, *params)

    def fused_forward(self, x: Tensor, take_abs: bool = False, nhwc: Optional[Tensor] = None, int_max: int = 0,
                      requant8: bool = False, out_dtype: Optional[torch.dtype] = None, out_max: int = 0) -> Tensor:
        """x: fp32 (or integer pixel) (B, C, H, W) on a B200.  Returns fp32 NCHW like the module chain would (or integer
        pixels when ``out_dtype`` asks for them)."""
        if not x.is_cuda:
            raise RuntimeError("licos_b200: the fused path needs CUDA tensors (no CPU fallback exists)")
        steps = self._steps()
        if not steps:
            return x

        cur = x.contiguous()
        layout = ops.pixel_layout_of(cur, requant8)
        first = steps[0][0]
        direct_first = (steps[0][1] == _lib.CONV_5X5_S2 and first.in_channels <= 16 and not take_abs)
        if layout != _lib.LAYOUT_NCHW_F32 and not (direct_first and first.in_channels in (1, 3) and (
                cur.shape[3] * cur.element_size()) % 16 == 0 and cur.data_ptr() % 16 == 0):
            # shapes the pipelined first layer does not take (13 bands, rows that are not 16-byte multiples): scale to fp32
            # with the standalone kernel (licos_raw_dn_to_unit, same values) and continue on the fp32 path
            cur = ops.pixels_to_unit(cur, int_max, requant8)
            layout = _lib.LAYOUT_NCHW_F32
        last_layout = {None: _lib.LAYOUT_NCHW_F32, torch.float32: _lib.LAYOUT_NCHW_F32, torch.uint8: _lib.LAYOUT_NCHW_U8,
                       torch.uint16: _lib.LAYOUT_NCHW_U16}.get(out_dtype)
        if last_layout is None:
            raise TypeError(f"out_dtype must be float32, uint8 or uint16, got {out_dtype}")
        if last_layout != _lib.LAYOUT_NCHW_F32:
            lm, lk = steps[-1][0], steps[-1][1]
            if not (lk == _lib.DECONV_5X5_S2 and lm.out_channels <= 4 and lm.in_channels % 64 == 0 and lm.in_channels <= 256
                    and steps[-1][2] == _lib.EPI_NONE):
                raise NotImplementedError("licos_b200: integer pixel output needs a synthesis transform that ends in a "
                                          "5x5 stride-2 transposed conv to <= 4 bands")
        if not direct_first:
            if nhwc is not None and not take_abs:
                if nhwc.dtype != torch.bfloat16 or tuple(nhwc.shape) != (x.shape[0], x.shape[2], x.shape[3], x.shape[1]):
                    raise ValueError("nhwc must be the bf16 (B, H, W, C) copy of x")
                cur = nhwc.contiguous()
            else:
                cur = T.nchw_to_nhwc_bf16(cur, take_abs=take_abs)
            layout = _lib.LAYOUT_NHWC_BF16
        for n, (m, kind, epi, gdn) in enumerate(steps):
            last = n == len(steps) - 1
            out_layout = last_layout if last else _lib.LAYOUT_NHWC_BF16
            # integer pixel layouts share the fp32 first-layer weight packing
            w_layout = _lib.LAYOUT_NCHW_F32 if layout != _lib.LAYOUT_NHWC_BF16 else layout
            packed, bias = self._packed_weight(m, kind, w_layout)
            beta = gamma = None
            if gdn is not None:
                if gdn.beta.numel() != m.out_channels:
                    raise ValueError("GDN channel count does not match the preceding convolution")
                beta, gamma = self._packed_gdn(gdn)
                if bias is None:
                    bias = torch.zeros(m.out_channels, dtype=torch.float32, device=cur.device)
            lim = int_max if (n == 0 and layout not in (_lib.LAYOUT_NCHW_F32, _lib.LAYOUT_NHWC_BF16)) else (
                out_max if (last and out_layout != _lib.LAYOUT_NCHW_F32) else 0)
            cur = T.conv_forward(cur, kind=kind, epilogue=epi, in_layout=layout, out_layout=out_layout,
                                   in_c=m.in_channels, out_c=m.out_channels, weight=packed, bias=bias,
                                   beta=beta, gamma=gamma, int_max=lim)
            layout = out_layout
        return cur


# ---------------------------------------------------------------------------------------------
# native training path: one autograd node per transform (g_a / g_s / h_a / h_s)
# ---------------------------------------------------------------------------------------------

class _ChainFn(torch.autograd.Function):
    """forward: conv + bias + GDN / IGDN in one launch on the engine, which also writes the pre-activation for backward;
    backward: data gradients on the engine (conv <-> transposed conv with the same weight), weight / gamma gradients on
    ``licos_conv_wgrad``, bias / beta gradients by column sums.  Activations and their gradients are bf16 NHWC."""

    @staticmethod
    def forward(ctx, seq: "FusedSequential", steps, x: Tensor, *params):
        if not x.is_cuda:
            raise RuntimeError("licos_b200: training needs CUDA tensors on a B200 (no CPU path exists)")
        L = _lib
        dev = x.device
        cur = x.detach().contiguous()
        layout = L.LAYOUT_NCHW_F32
        first = steps[0][0]
        if not (steps[0][1] == L.CONV_5X5_S2 and first.in_channels <= 16):
            cur = T.nchw_to_nhwc_bf16(cur)
            layout = L.LAYOUT_NHWC_BF16
        saved = []
        # Every operand re-pack of the chain (forward layouts, GDN parameters, the data-gradient layouts backward needs)
        # is issued up front on a side stream: a dozen ~5 us launches that now run beside the first layers instead of
        # between them (under graph capture: a parallel branch of the graph).  Layer n waits for its own event only.
        main = torch.cuda.current_stream(dev)
        side = _pack_stream(dev)
        side.wait_stream(main)
        pre, lay = [], layout
        with torch.cuda.stream(side):
            for n, (m, kind, epi, gdn) in enumerate(steps):
                # the training path never trusts the version-keyed caches (see FusedSequential.train)
                ent = {"w": seq._packed_weight(m, kind, lay, force=True),
                       "gdn": None if gdn is None else seq._packed_gdn(gdn, force=True)}
                ent["ready"] = side.record_event()
                pre.append(ent)
                lay = L.LAYOUT_NHWC_BF16
            for n, (m, kind, epi, gdn) in enumerate(steps):
                pre[n]["wd"] = _dgrad_weight(m, kind, n == len(steps) - 1) if (n > 0 or x.requires_grad) else None
            packs_done = side.record_event()
        for ent in pre:
            for t in (*ent["w"], *(ent["gdn"] or ()), ent["wd"]):
                if t is not None and t.is_cuda:
                    t.record_stream(main)
        for n, (m, kind, epi, gdn) in enumerate(steps):
            last = n == len(steps) - 1
            out_layout = L.LAYOUT_NCHW_F32 if last else L.LAYOUT_NHWC_BF16
            main.wait_event(pre[n]["ready"])
            packed, bias = pre[n]["w"]
            rec = {"in": cur, "in_layout": layout, "wd": pre[n]["wd"]}
            if gdn is not None:
                C = m.out_channels
                if bias is None:
                    bias = torch.zeros(C, dtype=torch.float32, device=dev)
                beta_hat, gamma_hat = pre[n]["gdn"]
                rec["gdn_packed"] = (beta_hat, gamma_hat)
                B, _, OH, OW = _out_shape(kind, cur, layout)
                v = torch.empty((B, OH, OW, C), dtype=torch.bfloat16, device=dev)
                try:
                    # one launch: the fused conv + GDN epilogue also writes its input v = conv + bias for the backward pass
                    cur = T.conv_forward(cur, kind=kind, epilogue=epi, in_layout=layout, out_layout=out_layout,
                                         in_c=m.in_channels, out_c=C, weight=packed, bias=bias, beta=beta_hat, gamma=gamma_hat,
                                         pre_act=v)
                except NotImplementedError:
                    # shapes the fused kernels refuse (LICOS_ERR_UNSUPPORTED): conv with v kept, then the GDN as a 1x1 layer
                    v = T.conv_forward(cur, kind=kind, epilogue=L.EPI_NONE, in_layout=layout, out_layout=L.LAYOUT_NHWC_BF16,
                                       in_c=m.in_channels, out_c=C, weight=packed, bias=bias)
                    cur = T.conv_forward(v, kind=L.CONV_1X1, epilogue=epi, in_layout=L.LAYOUT_NHWC_BF16, out_layout=out_layout,
                                         in_c=C, out_c=C, weight=_identity_1x1(C, dev), bias=_zeros(C, dev), beta=beta_hat,
                                         gamma=gamma_hat)
                rec["v"] = v
            else:
                cur = T.conv_forward(cur, kind=kind, epilogue=epi, in_layout=layout, out_layout=out_layout,
                                       in_c=m.in_channels, out_c=m.out_channels, weight=packed, bias=bias)
                if epi == L.EPI_RELU:
                    rec["y"] = cur
            layout = out_layout
            saved.append(rec)
        main.wait_event(packs_done)  # join: backward (and the end of a graph capture) finds every pack finished
        ctx.seq, ctx.steps, ctx.saved = seq, steps, saved
        # The chain's OUTPUT doubles as the ReLU mask of its last layer (h_s): keep it through save_for_backward so that an
        # in-place edit by the caller (clamp_, mul_) trips autograd's version check instead of silently corrupting the mask.
        # Inner activations never leave this function, so they stay plain references.
        ctx.out_is_mask = bool(saved) and "y" in saved[-1] and saved[-1]["y"] is cur
        if ctx.out_is_mask:
            ctx.save_for_backward(cur)
        ctx.x_needs_grad = x.requires_grad
        ctx.param_needs = [p is not None and p.requires_grad for p in params]
        return cur

    @staticmethod
    def backward(ctx, g_out: Tensor):
        L = _lib
        seq, steps, saved = ctx.seq, ctx.steps, ctx.saved
        if ctx.out_is_mask:
            saved[-1]["y"] = ctx.saved_tensors[0]  # raises if the caller modified the output in place
        dev = g_out.device
        # one zeroed fp32 arena for everything the kernels accumulate into (weight / gamma gradients, channel sums)
        need = 0
        for m, kind, epi, gdn in steps:
            taps = {L.CONV_3X3_S1: 9}.get(kind, 25)
            need += taps * max(m.in_channels, 64) * max(m.out_channels, 64) * 2 + m.out_channels
            if gdn is not None:
                need += m.out_channels * m.out_channels + m.out_channels
        arena = torch.zeros(need, dtype=torch.float32, device=dev)
        used = [0]

        def take(n):
            v = arena[used[0]:used[0] + n]
            used[0] += (n + 3) // 4 * 4  # keep 16-byte alignment for the vector red.add
            return v

        grads = []  # per step, in order: weight, bias, [beta, gamma]
        g = g_out.contiguous()
        g_layout = L.LAYOUT_NCHW_F32
        for n in range(len(steps) - 1, -1, -1):
            m, kind, epi, gdn = steps[n]
            rec = saved[n]
            a_in, in_layout = rec["in"], rec["in_layout"]
            narrow = g_layout == L.LAYOUT_NCHW_F32 and kind == L.DECONV_5X5_S2 and m.out_channels <= 4
            need_dgrad = n > 0 or ctx.x_needs_grad
            Co, Ci = m.out_channels, m.in_channels
            if narrow:
                # g_s[6]: C -> C_img transposed conv, gradient arrives as fp32 NCHW
                dw = _edge_wgrad(a_in, g, take)[:, :Co * 25].reshape(Ci, Co, 5, 5)
                db = g.sum(dim=(0, 2, 3)) if m.bias is not None else None
                grads.append([dw, db])
                if need_dgrad:
                    wd = rec["wd"]
                    g = T.conv_forward(g, kind=L.CONV_5X5_S2, epilogue=L.EPI_NONE, in_layout=L.LAYOUT_NCHW_F32,
                                         out_layout=L.LAYOUT_NHWC_BF16, in_c=Co, out_c=Ci, weight=wd, bias=None)
                    g_layout = L.LAYOUT_NHWC_BF16
                continue
            if g_layout == L.LAYOUT_NCHW_F32:
                if epi == L.EPI_RELU:
                    g = g * (rec["y"] > 0).to(g.dtype)  # last layer of h_s: fp32 NCHW output
                g = T.nchw_to_nhwc_bf16(g)
                g_layout = L.LAYOUT_NHWC_BF16
            elif epi == L.EPI_RELU:
                g = ops.relu_bwd(rec["y"], g)
            d_beta = d_gamma = db = None
            if gdn is not None:
                C = Co
                v = rec["v"]
                beta_hat, gamma_hat = rec["gdn_packed"]  # fp32 [C], bf16 [C][C] == the packed 1x1 weight of the norm mix
                d_beta_hat, d_gamma_hat = take(C), take(C * C)
                db = take(C) if m.bias is not None else None
                if C == 128 and not _UNFUSED_GDN_BWD:
                    g = T.gdn_backward(v, g, gamma_hat, beta_hat, gdn.inverse, d_gamma_hat, d_beta_hat, db)
                else:
                    gamma_hat_t = gamma_hat.t().contiguous()
                    x2 = ops.square_bf16(v)
                    norm = T.conv_forward(x2, kind=L.CONV_1X1, epilogue=L.EPI_NONE, in_layout=L.LAYOUT_NHWC_BF16,
                                            out_layout=L.LAYOUT_NHWC_BF16, in_c=C, out_c=C, weight=gamma_hat, bias=beta_hat)
                    d_norm, d_direct = ops.gdn_bwd_mid(v, g, norm, gdn.inverse, sum_out=d_beta_hat)
                    t = T.conv_forward(d_norm, kind=L.CONV_1X1, epilogue=L.EPI_NONE, in_layout=L.LAYOUT_NHWC_BF16,
                                         out_layout=L.LAYOUT_NHWC_BF16, in_c=C, out_c=C, weight=gamma_hat_t, bias=None)
                    g = ops.gdn_bwd_out(v, t, d_direct, sum_out=db)
                    T.conv_wgrad(d_norm, x2, L.CONV_1X1, out=d_gamma_hat)
                d_beta, d_gamma = ops.gdn_param_grad(
                    gdn.beta.detach(), gdn.gamma.detach(), d_beta_hat, d_gamma_hat.view(C, C),
                    gdn.beta_reparam.bound_f, gdn.gamma_reparam.bound_f)
            elif m.bias is not None:
                db = ops.colsum_bf16(g, acc=take(Co))
            # g is now the gradient with respect to conv + bias
            if in_layout == L.LAYOUT_NCHW_F32:  # g_a[0]: image in, K = 25 C_in
                dw = _edge_wgrad(g, a_in, take)[:, :Ci * 25].reshape(Co, Ci, 5, 5)
            elif kind == L.DECONV_5X5_S2:
                dw = T.conv_wgrad(a_in, g, kind, out=take(25 * Ci * Co)).permute(1, 2, 0).reshape(Ci, Co, 5, 5)
            else:
                k = 3 if kind == L.CONV_3X3_S1 else 5
                dw = T.conv_wgrad(g, a_in, kind, out=take(k * k * Co * Ci)).permute(1, 2, 0).reshape(Co, Ci, k, k)
            grads.append([dw, db] + ([d_beta, d_gamma] if gdn is not None else []))
            if need_dgrad:
                out_layout = L.LAYOUT_NHWC_BF16 if n > 0 else L.LAYOUT_NCHW_F32
                dk = {L.CONV_5X5_S2: L.DECONV_5X5_S2, L.DECONV_5X5_S2: L.CONV_5X5_S2}.get(kind, L.CONV_3X3_S1)
                wd = rec["wd"]
                g = T.conv_forward(g, kind=dk, epilogue=L.EPI_NONE, in_layout=L.LAYOUT_NHWC_BF16, out_layout=out_layout,
                                     in_c=Co, out_c=Ci, weight=wd, bias=None)
                g_layout = out_layout
        grads.reverse()
        flat = [t for sg in grads for t in sg]
        flat = [t if need else None for t, need in zip(flat, ctx.param_needs)]
        gx = g if ctx.x_needs_grad else None
        return (None, None, gx, *flat)


_CONST_CACHE = {}


_PACK_STREAMS = {}


def _pack_stream(dev) -> "torch.cuda.Stream":
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _PACK_STREAMS:
        _PACK_STREAMS[key] = torch.cuda.Stream(device=key)
    return _PACK_STREAMS[key]


def _dgrad_weight(m: nn.Module, kind: int, last: bool) -> Tensor:
    """The layer's weight packed for its DATA gradient: conv 5x5 s2 <-> transposed conv with the same (O, I, 5, 5) tensor,
    conv 3x3 with the flipped, transposed one; g_s's narrow last layer (``last``) takes the fp32-NCHW-input layout."""
    L = _lib
    w, Co, Ci = m.weight.detach(), m.out_channels, m.in_channels
    if kind == L.DECONV_5X5_S2 and last and Co <= 4:
        return ops.pack_conv_weight(w.contiguous(), L.CONV_5X5_S2, Ci, Co, L.LAYOUT_NCHW_F32)
    if kind == L.CONV_5X5_S2:
        return ops.pack_conv_weight(w.contiguous(), L.DECONV_5X5_S2, Ci, Co, L.LAYOUT_NHWC_BF16)
    if kind == L.DECONV_5X5_S2:
        return ops.pack_conv_weight(w.contiguous(), L.CONV_5X5_S2, Ci, Co, L.LAYOUT_NHWC_BF16)
    return ops.pack_conv_weight(w.flip(2, 3).transpose(0, 1).contiguous(), L.CONV_3X3_S1, Ci, Co, L.LAYOUT_NHWC_BF16)


def _edge_wgrad(small: Tensor, image: Tensor, take) -> Tensor:
    """fp32 [Cs][k_pad] weight gradient of a 5x5 stride-2 edge layer (g_a[0], g_s[6]): ``small`` is the bf16 NHWC tensor on
    the many-channel side, ``image`` the fp32 NCHW one on the band side.  One fused launch where the kernel is built for
    the shape (1 / 3 bands); otherwise the patch matrix goes through HBM (im2col + a 1x1 weight gradient)."""
    cs, kp = small.shape[-1], ops.im2col_kpad(image.shape[1])
    out = take(cs * kp)
    try:
        return T.conv_wgrad_image(small, image, out)
    except NotImplementedError:
        return T.conv_wgrad(small, ops.im2col5x5s2(image), _lib.CONV_1X1, out=out)[0]


def _out_shape(kind: int, x: Tensor, layout: int):
    """(B, -, OH, OW) of a conv_forward launch of this kind (ops.conv_forward's own rule)."""
    if layout == _lib.LAYOUT_NHWC_BF16:
        B, H, W = x.shape[0], x.shape[1], x.shape[2]
    else:
        B, H, W = x.shape[0], x.shape[2], x.shape[3]
    if kind == _lib.CONV_5X5_S2:
        return B, None, (H + 1) // 2, (W + 1) // 2
    if kind == _lib.DECONV_5X5_S2:
        return B, None, 2 * H, 2 * W
    return B, None, H, W


def _identity_1x1(C: int, dev) -> Tensor:
    key = ("eye", C, str(dev))
    if key not in _CONST_CACHE:
        _CONST_CACHE[key] = ops.pack_conv_weight(torch.eye(C, device=dev).reshape(C, C, 1, 1).contiguous(), _lib.CONV_1X1, C, C,
                                                 _lib.LAYOUT_NHWC_BF16)
    return _CONST_CACHE[key]


def _zeros(C: int, dev) -> Tensor:
    key = ("zeros", C, str(dev))
    if key not in _CONST_CACHE:
        _CONST_CACHE[key] = torch.zeros(C, dtype=torch.float32, device=dev)
    return _CONST_CACHE[key]
