"""Tensor-level wrappers over the C ABI: translate torch tensors <-> raw pointers and status codes <->
exceptions.  torch is used for device memory and streams only; no torch op computes anything here.

Every device op requires CUDA tensors; there is no CPU path (the CPU restatement lives in oracle/ and is
test infrastructure).
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import ConvArgs, EbFusedArgs, EbParams, EbRawPtrs, WgradArgs, check, lib


def _stream() -> int:
    # raw cudaStream_t of torch's current stream; the Python-level torch.cuda.current_stream() costs ~8 us per call,
    # which at ~80 launches per training step was a quarter of the eager loop's host time
    try:
        return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())
    except AttributeError:  # private bindings moved: fall back to the public API
        return torch.cuda.current_stream().cuda_stream


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


_FROZEN_CAPTURE = False  # set by graphs.GraphedForward: parameters are constants of the graph being captured


def capture_rebuild() -> bool:
    """True while a CUDA graph is being captured whose replays may see CHANGED parameters (a training step: the optimizer
    updates them in place without touching version counters), so everything derived from a parameter -- packed weights,
    GDN tables, the bottleneck's tables -- must be rebuilt inside the graph instead of being taken from a cache."""
    return torch.cuda.is_available() and torch.cuda.is_current_stream_capturing() and not _FROZEN_CAPTURE


def _need_cuda(*tensors: Optional[torch.Tensor]) -> None:
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("licos_b200 kernels need CUDA tensors: there is no CPU path in this package")
        if not t.is_contiguous():
            raise ValueError("licos_b200 kernels need contiguous tensors")


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        raise TypeError(f"expected float32, got {t.dtype}")
    return t


# ---------------------------------------------------------------------------------------------
# layout
# ---------------------------------------------------------------------------------------------

def nchw_to_nhwc_bf16(x: torch.Tensor, take_abs: bool = False) -> torch.Tensor:
    """fp32 (B, C, H, W) -> bf16 tensor of logical shape (B, H, W, C)."""
    _need_cuda(_f32(x))
    B, C, H, W = x.shape
    out = torch.empty((B, H, W, C), dtype=torch.bfloat16, device=x.device)
    check(lib.licos_nchw_f32_to_nhwc_bf16(x.data_ptr(), out.data_ptr(), B, C, H * W, int(take_abs), _stream()),
          "nchw_to_nhwc_bf16")
    return out


def nhwc_bf16_to_nchw(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(x)
    if x.dtype != torch.bfloat16:
        raise TypeError("expected bfloat16")
    B, H, W, C = x.shape
    out = torch.empty((B, C, H, W), dtype=torch.float32, device=x.device)
    check(lib.licos_nhwc_bf16_to_nchw_f32(x.data_ptr(), out.data_ptr(), B, C, H * W, _stream()), "nhwc_bf16_to_nchw")
    return out


# ---------------------------------------------------------------------------------------------
# convolutions
# ---------------------------------------------------------------------------------------------

def pack_conv_weight(w: torch.Tensor, kind: int, out_c: int, in_c: int, in_layout: int) -> torch.Tensor:
    _need_cuda(_f32(w))
    nbytes = lib.licos_packed_weight_bytes(kind, out_c, in_c, in_layout)
    check(int(nbytes), "packed_weight_bytes")
    packed = torch.empty(int(nbytes) // 2, dtype=torch.bfloat16, device=w.device)
    check(lib.licos_pack_conv_weight(w.data_ptr(), kind, out_c, in_c, in_layout, packed.data_ptr(), _stream()),
          "pack_conv_weight")
    return packed


def gdn_pack(beta: torch.Tensor, gamma: torch.Tensor, beta_bound: float, gamma_bound: float, pedestal: float):
    _need_cuda(_f32(beta), _f32(gamma))
    C = beta.numel()
    beta_hat = torch.empty(C, dtype=torch.float32, device=beta.device)
    gamma_hat = torch.empty((C, C), dtype=torch.bfloat16, device=beta.device)
    check(lib.licos_gdn_pack(beta.data_ptr(), gamma.data_ptr(), C, beta_bound, gamma_bound, pedestal,
                             beta_hat.data_ptr(), gamma_hat.data_ptr(), _stream()), "gdn_pack")
    return beta_hat, gamma_hat


_PIXEL_LAYOUT_DTYPES = {
    _lib.LAYOUT_NCHW_F32: (torch.float32,), _lib.LAYOUT_NCHW_U8: (torch.uint8,),
    _lib.LAYOUT_NCHW_U16: (torch.uint16, torch.int16), _lib.LAYOUT_NCHW_U16_Q8: (torch.uint16, torch.int16),
}


def pixel_layout_of(x: torch.Tensor, requant8: bool = False) -> int:
    """The LICOS_LAYOUT_* of an image tensor (B, C, H, W): fp32 in [0, 1], or integer tiles (uint8; 12-bit DNs in
    uint16 / int16 storage, optionally with the reference's 8-bit re-quantisation, raw_image_folder.py:193-196)."""
    if x.dtype == torch.float32:
        return _lib.LAYOUT_NCHW_F32
    if x.dtype == torch.uint8:
        return _lib.LAYOUT_NCHW_U8
    if x.dtype in (torch.uint16, torch.int16):
        return _lib.LAYOUT_NCHW_U16_Q8 if requant8 else _lib.LAYOUT_NCHW_U16
    raise TypeError(f"image tensors are float32, uint8 or 16-bit integers, got {x.dtype}")


def conv_forward(x: torch.Tensor, *, kind: int, epilogue: int, in_layout: int, out_layout: int, in_c: int,
                 out_c: int, weight: torch.Tensor, bias: Optional[torch.Tensor], beta: Optional[torch.Tensor] = None,
                 gamma: Optional[torch.Tensor] = None, sm_count: int = 0, int_max: int = 0,
                 pre_act: Optional[torch.Tensor] = None) -> torch.Tensor:
    """One conv / deconv layer with its fused epilogue.  x is fp32 / integer NCHW or bf16 NHWC (see in_layout);
    out_layout may be an integer pixel layout for the model's last layer (int_max = full-scale value, 0 = default)."""
    _need_cuda(x, weight, bias, beta, gamma)
    if in_layout in _PIXEL_LAYOUT_DTYPES:
        B, C, H, W = x.shape
        if x.dtype not in _PIXEL_LAYOUT_DTYPES[in_layout]:
            raise TypeError(f"layout {in_layout} expects {_PIXEL_LAYOUT_DTYPES[in_layout]}, got {x.dtype}")
    else:
        B, H, W, C = x.shape
        if x.dtype != torch.bfloat16:
            raise TypeError("NHWC input must be bfloat16")
    if C != in_c:
        raise ValueError(f"input has {C} channels, layer expects {in_c}")
    if kind == _lib.CONV_5X5_S2:
        OH, OW = (H + 1) // 2, (W + 1) // 2
    elif kind == _lib.DECONV_5X5_S2:
        OH, OW = 2 * H, 2 * W
    else:
        OH, OW = H, W
    if out_layout == _lib.LAYOUT_NCHW_F32:
        out = torch.empty((B, out_c, OH, OW), dtype=torch.float32, device=x.device)
    elif out_layout == _lib.LAYOUT_NCHW_U8:
        out = torch.empty((B, out_c, OH, OW), dtype=torch.uint8, device=x.device)
    elif out_layout == _lib.LAYOUT_NCHW_U16:
        out = torch.empty((B, out_c, OH, OW), dtype=torch.uint16, device=x.device)
    else:
        out = torch.empty((B, OH, OW, out_c), dtype=torch.bfloat16, device=x.device)
    a = ConvArgs()
    a.kind, a.epilogue, a.batch, a.in_h, a.in_w, a.in_c, a.out_c = kind, epilogue, B, H, W, in_c, out_c
    a.in_layout, a.out_layout = in_layout, out_layout
    a.in_, a.out, a.weight = x.data_ptr(), out.data_ptr(), weight.data_ptr()
    a.bias, a.beta, a.gamma = _ptr(bias), _ptr(beta), _ptr(gamma)
    a.sm_count = sm_count
    a.int_max = int(int_max)
    if pre_act is not None:
        _need_cuda(pre_act)
        if pre_act.dtype != torch.bfloat16 or tuple(pre_act.shape) != (B, OH, OW, out_c):
            raise ValueError("pre_act must be a bf16 (B, out_h, out_w, out_c) tensor")
        a.pre_act = pre_act.data_ptr()
    ws = None
    nws = 0
    if in_layout == _lib.LAYOUT_NCHW_F32:  # only the explicit-im2col first-layer fallback needs scratch
        nws = int(lib.licos_conv_workspace_bytes(ctypes.byref(a)))
        check(nws, "conv_workspace_bytes")
    if nws > 0:
        ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
        a.workspace, a.workspace_bytes = ws.data_ptr(), nws
    check(lib.licos_conv_forward(ctypes.byref(a), _stream()), "conv_forward")
    return out


# ---------------------------------------------------------------------------------------------
# backward pass of the transforms
# ---------------------------------------------------------------------------------------------

def _bf16(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.bfloat16:
        raise TypeError(f"expected bfloat16, got {t.dtype}")
    return t


def conv_wgrad(small: torch.Tensor, big: torch.Tensor, kind: int, sm_count: int = 0,
               out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [taps][small_c][big_c] = sum over pixels of small (x) big shifted by the tap; both bf16 NHWC.
    Conv2d: small = output gradient, big = layer input; ConvTranspose2d: small = layer input, big = output gradient.
    ``out``: optional ZEROED fp32 buffer of taps * small_c * big_c elements (the kernel accumulates into it)."""
    _need_cuda(_bf16(small), _bf16(big), out)
    B, h, w, cs = small.shape
    B2, H, W, cb = big.shape
    if B != B2:
        raise ValueError("batch mismatch")
    taps = {_lib.CONV_5X5_S2: 25, _lib.DECONV_5X5_S2: 25, _lib.CONV_3X3_S1: 9, _lib.CONV_1X1: 1}[kind]
    if out is None:
        out = torch.zeros((taps, cs, cb), dtype=torch.float32, device=small.device)
    else:
        if out.numel() != taps * cs * cb or out.dtype != torch.float32:
            raise ValueError("out must hold taps * small_c * big_c float32 values")
        out = out.view(taps, cs, cb)
    a = WgradArgs()
    a.kind, a.batch, a.h, a.w, a.big_h, a.big_w, a.small_c, a.big_c = kind, B, h, w, H, W, cs, cb
    a.small_t, a.big_t, a.out, a.sm_count = small.data_ptr(), big.data_ptr(), out.data_ptr(), sm_count
    check(lib.licos_conv_wgrad(ctypes.byref(a), _stream()), "conv_wgrad")
    return out


def im2col_kpad(channels: int) -> int:
    return int(lib.licos_im2col5x5s2_kpad(channels))


def conv_wgrad_image(small: torch.Tensor, image: torch.Tensor, out: torch.Tensor, sm_count: int = 0) -> torch.Tensor:
    """Weight gradient of the 5x5 stride-2 edge layers with the im2col tile built in shared memory:
    out[cs][k] += sum_pixels small[b][oh][ow][cs] * image[b][c][2 oh + kh - 2][2 ow + kw - 2], k = (c*5 + kh)*5 + kw.
    small bf16 (B, ceil(H/2), ceil(W/2), Cs), image fp32 (B, C, H, W), out ZEROED fp32 with Cs * k_pad elements.
    Raises NotImplementedError for shapes the fused kernel is not built for (use im2col5x5s2 + conv_wgrad)."""
    _need_cuda(_bf16(small), _f32(image), _f32(out))
    B, C, H, W = image.shape
    if tuple(small.shape[:3]) != (B, (H + 1) // 2, (W + 1) // 2):
        raise ValueError("small must be (B, ceil(H/2), ceil(W/2), Cs)")
    cs, kp = small.shape[-1], im2col_kpad(C)
    if out.numel() != cs * kp:
        raise ValueError("out must hold small_c * k_pad float32 values")
    check(lib.licos_conv_wgrad_image(small.data_ptr(), image.data_ptr(), B, C, H, W, cs, out.data_ptr(), sm_count, _stream()),
          "conv_wgrad_image")
    return out.view(cs, kp)


def gdn_backward(x: torch.Tensor, g: torch.Tensor, gamma_hat: torch.Tensor, beta_hat: torch.Tensor, inverse: bool,
                 d_gamma_hat: torch.Tensor, d_beta_hat: torch.Tensor, d_bias: Optional[torch.Tensor], sm_count: int = 0):
    """Fused GDN / IGDN backward (128 channels): returns dx, accumulates into the zeroed fp32 d_gamma_hat [C][C],
    d_beta_hat [C] and d_bias [C] (optional).  Raises NotImplementedError for other widths (use the unfused pieces)."""
    _need_cuda(_bf16(x), _bf16(g), _bf16(gamma_hat), _f32(beta_hat), _f32(d_gamma_hat), _f32(d_beta_hat), d_bias)
    C = x.shape[-1]
    dx = torch.empty_like(x)
    check(lib.licos_gdn_backward(x.data_ptr(), g.data_ptr(), gamma_hat.data_ptr(), beta_hat.data_ptr(), int(inverse),
                                 x.numel() // C, C, dx.data_ptr(), d_gamma_hat.data_ptr(), d_beta_hat.data_ptr(), _ptr(d_bias),
                                 sm_count, _stream()), "gdn_backward")
    return dx


def square_bf16(x: torch.Tensor) -> torch.Tensor:
    _need_cuda(_bf16(x))
    out = torch.empty_like(x)
    check(lib.licos_square_bf16(x.data_ptr(), out.data_ptr(), x.numel(), _stream()), "square_bf16")
    return out


def gdn_bwd_mid(x: torch.Tensor, g: torch.Tensor, norm: torch.Tensor, inverse: bool, sum_out: Optional[torch.Tensor] = None):
    """(d_norm, d_direct); ``sum_out`` (zeroed fp32 [C]) receives the per-channel sums of d_norm."""
    _need_cuda(_bf16(x), _bf16(g), _bf16(norm), sum_out)
    d_norm, d_direct = torch.empty_like(x), torch.empty_like(x)
    check(lib.licos_gdn_bwd_mid(x.data_ptr(), g.data_ptr(), norm.data_ptr(), int(inverse), x.numel(), x.shape[-1],
                                d_norm.data_ptr(), d_direct.data_ptr(), _ptr(sum_out), _stream()), "gdn_bwd_mid")
    return d_norm, d_direct


def gdn_bwd_out(x: torch.Tensor, t: torch.Tensor, d_direct: torch.Tensor, sum_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dx = d_direct + 2 x t, written over d_direct; ``sum_out`` (zeroed fp32 [C]) receives its per-channel sums."""
    _need_cuda(_bf16(x), _bf16(t), _bf16(d_direct), sum_out)
    check(lib.licos_gdn_bwd_out(x.data_ptr(), t.data_ptr(), d_direct.data_ptr(), x.numel(), x.shape[-1], d_direct.data_ptr(),
                                _ptr(sum_out), _stream()), "gdn_bwd_out")
    return d_direct


def gdn_param_grad(beta: torch.Tensor, gamma: torch.Tensor, d_beta_hat: torch.Tensor, d_gamma_hat: torch.Tensor,
                   beta_bound: float, gamma_bound: float):
    """Gradients of the raw GDN parameters from those of the reparametrised ones (LowerBound's rule)."""
    _need_cuda(_f32(beta), _f32(gamma), _f32(d_beta_hat), _f32(d_gamma_hat))
    d_beta, d_gamma = torch.empty_like(beta), torch.empty_like(gamma)
    check(lib.licos_gdn_param_grad(beta.data_ptr(), gamma.data_ptr(), d_beta_hat.data_ptr(), d_gamma_hat.data_ptr(),
                                   beta.numel(), beta_bound, gamma_bound, d_beta.data_ptr(), d_gamma.data_ptr(), _stream()),
          "gdn_param_grad")
    return d_beta, d_gamma


def relu_bwd(y: torch.Tensor, g: torch.Tensor) -> torch.Tensor:
    _need_cuda(_bf16(y), _bf16(g))
    dx = torch.empty_like(g)
    check(lib.licos_relu_bwd(y.data_ptr(), g.data_ptr(), y.numel(), dx.data_ptr(), _stream()), "relu_bwd")
    return dx


def colsum_bf16(x: torch.Tensor, acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """fp32 [C] = sum over all leading dimensions of a bf16 (..., C) tensor (``acc``: optional zeroed buffer)."""
    _need_cuda(_bf16(x), acc)
    C = x.shape[-1]
    if acc is None:
        acc = torch.zeros(C, dtype=torch.float32, device=x.device)
    check(lib.licos_colsum_bf16(x.data_ptr(), x.numel() // C, C, acc.data_ptr(), _stream()), "colsum_bf16")
    return acc


def im2col5x5s2(x: torch.Tensor) -> torch.Tensor:
    """fp32 (B, C, H, W) -> bf16 (B, ceil(H/2), ceil(W/2), k_pad) patch matrix, k = (c*5 + kh)*5 + kw."""
    _need_cuda(_f32(x))
    B, C, H, W = x.shape
    kp = int(lib.licos_im2col5x5s2_kpad(C))
    rows = torch.empty((B, (H + 1) // 2, (W + 1) // 2, kp), dtype=torch.bfloat16, device=x.device)
    check(lib.licos_im2col5x5s2(x.data_ptr(), B, C, H, W, rows.data_ptr(), _stream()), "im2col5x5s2")
    return rows


# ---------------------------------------------------------------------------------------------
# entropy bottleneck
# ---------------------------------------------------------------------------------------------

class EbPacked:
    """Device-side parameter block of one EntropyBottleneck (see licos_eb_params)."""

    def __init__(self, packed: torch.Tensor, medians: torch.Tensor, widths: Sequence[int], form: int,
                 likelihood_bound: float):
        _need_cuda(_f32(packed), _f32(medians))
        self.packed, self.medians = packed, medians
        self.p = EbParams()
        self.p.channels = packed.shape[0]
        self.p.n_layers = len(widths) - 1
        for i, w in enumerate(widths):
            self.p.widths[i] = int(w)
        self.p.params_per_channel = packed.shape[1]
        self.p.packed, self.p.medians = packed.data_ptr(), medians.data_ptr()
        self.p.form = form
        self.p.likelihood_bound = float(likelihood_bound)
        self.lut = None  # [C][257] eval-mode likelihood table, built on first use (a function of the parameters only)

    def eval_lut(self) -> torch.Tensor:
        if self.lut is None or capture_rebuild():
            lut = torch.empty(int(lib.licos_eb_lut_floats(self.p.channels)), dtype=torch.float32, device=self.packed.device)
            check(lib.licos_eb_build_lut(ctypes.byref(self.p), lut.data_ptr(), _stream()), "eb_build_lut")
            self.lut = lut
        return self.lut


def eb_forward_eval(ebp: EbPacked, x: torch.Tensor):
    _need_cuda(_f32(x))
    B, C = x.shape[0], x.shape[1]
    if C != ebp.p.channels:
        raise ValueError("channel mismatch")
    hw = x.numel() // max(B * C, 1)
    y_hat, lik = torch.empty_like(x), torch.empty_like(x)
    if x.numel() == 0:
        return y_hat, lik
    lut = torch.empty(int(lib.licos_eb_lut_floats(C)), dtype=torch.float32, device=x.device)
    check(lib.licos_eb_forward_eval(ctypes.byref(ebp.p), x.data_ptr(), B, hw, lut.data_ptr(), y_hat.data_ptr(),
                                    lik.data_ptr(), _stream()), "eb_forward_eval")
    return y_hat, lik


def eb_forward_eval_fused(ebp: EbPacked, x: torch.Tensor, want_symbols: bool = False, want_nhwc: bool = False,
                          want_symbols_i16: bool = False, sum_ln: Optional[torch.Tensor] = None, want_float: bool = True):
    """One pass: (y_hat, likelihoods, symbols | None, bf16 NHWC y_hat | None[, int16 symbols | None]).  ``sum_ln`` (a
    one-element float64 device tensor) accumulates sum(ln(likelihoods)).  ``want_float=False`` skips y_hat / likelihoods
    (codec loop: symbols, the NHWC copy and the rate term are all it needs).  Shapes the tiled kernel does not take
    (hw % 4 != 0, odd channel counts) go through the separate kernels."""
    _need_cuda(_f32(x), sum_ln)
    B, C = x.shape[0], x.shape[1]
    if C != ebp.p.channels:
        raise ValueError("channel mismatch")
    hw = x.numel() // max(B * C, 1)
    if x.numel() == 0 or hw % 4 != 0 or C % 2 != 0 or C > 512 or x.dim() != 4:
        y_hat, lik = eb_forward_eval(ebp, x)
        sym = eb_symbols(x, ebp.medians) if (want_symbols or want_symbols_i16) else None
        if sum_ln is not None and x.numel():
            sum_log(lik, sum_ln)
        res = (y_hat, lik, sym if want_symbols else None,
               (nchw_to_nhwc_bf16(y_hat) if want_nhwc and x.dim() == 4 and x.numel() else None))
        if want_symbols_i16:
            res += (sym.clamp(-32768, 32767).to(torch.int16),)
        return res
    x = x.contiguous()
    y_hat, lik = (torch.empty_like(x), torch.empty_like(x)) if want_float else (None, None)
    sym = torch.empty(x.shape, dtype=torch.int32, device=x.device) if want_symbols else None
    sym16 = torch.empty(x.shape, dtype=torch.int16, device=x.device) if want_symbols_i16 else None
    nhwc = torch.empty((B, x.shape[2], x.shape[3], C), dtype=torch.bfloat16, device=x.device) if want_nhwc else None
    a = EbFusedArgs()
    a.x, a.batch, a.hw = x.data_ptr(), B, hw
    a.lut, a.lut_ready = ebp.eval_lut().data_ptr(), 1
    a.y_hat, a.lik, a.symbols, a.symbols_i16 = _ptr(y_hat), _ptr(lik), _ptr(sym), _ptr(sym16)
    a.y_hat_nhwc_bf16, a.sum_ln = _ptr(nhwc), _ptr(sum_ln)
    check(lib.licos_eb_eval_fused(ctypes.byref(ebp.p), ctypes.byref(a), _stream()), "eb_eval_fused")
    res = (y_hat, lik, sym, nhwc)
    return res + (sym16,) if want_symbols_i16 else res


def eb_forward_noise(ebp: EbPacked, x: torch.Tensor, noise: Optional[torch.Tensor], seed: int = 0):
    _need_cuda(_f32(x), noise)
    B, C = x.shape[0], x.shape[1]
    if C != ebp.p.channels:
        raise ValueError("channel mismatch")
    hw = x.numel() // max(B * C, 1)
    y_hat, lik = torch.empty_like(x), torch.empty_like(x)
    if x.numel() == 0:
        return y_hat, lik
    check(lib.licos_eb_forward_noise(ctypes.byref(ebp.p), x.data_ptr(), _ptr(noise), seed & (2 ** 64 - 1), B, hw,
                                     y_hat.data_ptr(), lik.data_ptr(), _stream()), "eb_forward_noise")
    return y_hat, lik


def eb_backward(ebp: EbPacked, y_hat: torch.Tensor, g_lik: Optional[torch.Tensor], g_yhat: Optional[torch.Tensor]):
    """Backward of the training-mode bottleneck: (d_x, d_packed [C][params_per_channel])."""
    _need_cuda(_f32(y_hat), g_lik, g_yhat)
    B, C = y_hat.shape[0], y_hat.shape[1]
    hw = y_hat.numel() // max(B * C, 1)
    d_x = torch.empty_like(y_hat)
    d_packed = torch.zeros_like(ebp.packed)
    check(lib.licos_eb_backward(ctypes.byref(ebp.p), y_hat.data_ptr(), _ptr(g_lik), _ptr(g_yhat), B, hw, d_x.data_ptr(),
                                d_packed.data_ptr(), _stream()), "eb_backward")
    return d_x, d_packed


def _eb_raw_ptrs(matrices, biases, factors) -> EbRawPtrs:
    r = EbRawPtrs()
    for i, t in enumerate(matrices):
        r.matrix[i] = t.data_ptr()
    for i, t in enumerate(biases):
        r.bias[i] = t.data_ptr()
    for i, t in enumerate(factors):
        r.factor[i] = t.data_ptr()
    return r


def _eb_widths(widths: Sequence[int]):
    return (ctypes.c_int * len(widths))(*[int(w) for w in widths])


def eb_pack_params(matrices, biases, factors, widths: Sequence[int]) -> torch.Tensor:
    """[C][params_per_channel] block (softplus(_matrix_i), _bias_i, tanh(_factor_i))_i in one launch."""
    _need_cuda(*[_f32(t) for t in (*matrices, *biases, *factors)])
    C = matrices[0].shape[0]
    ppc = sum(t[0].numel() for t in (*matrices, *biases, *factors))
    packed = torch.empty((C, ppc), dtype=torch.float32, device=matrices[0].device)
    raw = _eb_raw_ptrs(matrices, biases, factors)
    check(lib.licos_eb_pack_params(ctypes.byref(raw), C, len(widths) - 1, _eb_widths(widths), packed.data_ptr(), _stream()),
          "eb_pack_params")
    return packed


def eb_param_grads(matrices, biases, factors, d_packed: torch.Tensor, widths: Sequence[int]):
    """Raw-parameter gradients from licos_eb_backward's d_packed: (d_matrices, d_biases, d_factors)."""
    _need_cuda(_f32(d_packed), *[_f32(t) for t in (*matrices, *biases, *factors)])
    C = matrices[0].shape[0]
    gm, gb, gf = [torch.empty_like(t) for t in matrices], [torch.empty_like(t) for t in biases], [torch.empty_like(t) for t in factors]
    raw, out = _eb_raw_ptrs(matrices, biases, factors), _eb_raw_ptrs(gm, gb, gf)
    check(lib.licos_eb_param_grads(ctypes.byref(raw), d_packed.data_ptr(), C, len(widths) - 1, _eb_widths(widths),
                                   ctypes.byref(out), _stream()), "eb_param_grads")
    return gm, gb, gf


def eb_aux_loss(packed: torch.Tensor, widths: Sequence[int], quantiles: torch.Tensor, target: torch.Tensor):
    """(loss [1], d_quantiles like quantiles) of EntropyBottleneck.loss() with the density parameters held constant."""
    _need_cuda(_f32(packed), _f32(quantiles), _f32(target))
    C = packed.shape[0]
    loss = torch.zeros(1, dtype=torch.float32, device=packed.device)
    d_q = torch.empty_like(quantiles)
    check(lib.licos_eb_aux_loss(packed.data_ptr(), C, len(widths) - 1, _eb_widths(widths), quantiles.data_ptr(), target.data_ptr(),
                                loss.data_ptr(), d_q.data_ptr(), _stream()), "eb_aux_loss")
    return loss, d_q


def eb_symbols(x: torch.Tensor, medians: torch.Tensor, want_indexes: bool = False):
    _need_cuda(_f32(x), _f32(medians))
    B, C = x.shape[0], x.shape[1]
    hw = x.numel() // max(B * C, 1)
    sym = torch.empty(x.shape, dtype=torch.int32, device=x.device)
    idx = torch.empty(x.shape, dtype=torch.int32, device=x.device) if want_indexes else None
    if x.numel() == 0:
        return (sym, idx) if want_indexes else sym
    check(lib.licos_eb_symbols(x.data_ptr(), medians.data_ptr(), B, C, hw, sym.data_ptr(), _ptr(idx), _stream()),
          "eb_symbols")
    return (sym, idx) if want_indexes else sym


def eb_dequantize(sym: torch.Tensor, medians: torch.Tensor) -> torch.Tensor:
    _need_cuda(sym, _f32(medians))
    if sym.dtype != torch.int32:
        raise TypeError("symbols must be int32")
    B, C = sym.shape[0], sym.shape[1]
    hw = sym.numel() // max(B * C, 1)
    out = torch.empty(sym.shape, dtype=torch.float32, device=sym.device)
    if sym.numel() == 0:
        return out
    check(lib.licos_eb_dequantize(sym.data_ptr(), medians.data_ptr(), B, C, hw, out.data_ptr(), _stream()),
          "eb_dequantize")
    return out


# ---------------------------------------------------------------------------------------------
# gaussian conditional
# ---------------------------------------------------------------------------------------------

def gc_forward(y, scales, means=None, noise=None, *, training=False, scale_bound=0.11, likelihood_bound=1e-9,
               seed: int = 0):
    _need_cuda(_f32(y), _f32(scales), means, noise)
    if scales.shape != y.shape or (means is not None and means.shape != y.shape):
        raise ValueError("scales / means must have the shape of the input")
    y_hat, lik = torch.empty_like(y), torch.empty_like(y)
    if y.numel() == 0:
        return y_hat, lik
    check(lib.licos_gc_forward(y.data_ptr(), scales.data_ptr(), _ptr(means), _ptr(noise), seed & (2 ** 64 - 1),
                               y.numel(), int(training), scale_bound, likelihood_bound, y_hat.data_ptr(),
                               lik.data_ptr(), _stream()), "gc_forward")
    return y_hat, lik


def gc_backward(y_hat: torch.Tensor, scales: torch.Tensor, means: Optional[torch.Tensor], g_lik: Optional[torch.Tensor],
                g_yhat: Optional[torch.Tensor], scale_bound: float, likelihood_bound: float, want_means: bool = False):
    """Backward of the training-mode GaussianConditional: (d_y, d_scales, d_means | None)."""
    _need_cuda(_f32(y_hat), _f32(scales), means, g_lik, g_yhat)
    d_y, d_s = torch.empty_like(y_hat), torch.empty_like(y_hat)
    d_m = torch.empty_like(y_hat) if (want_means and means is not None) else None
    check(lib.licos_gc_backward(y_hat.data_ptr(), scales.data_ptr(), _ptr(means), _ptr(g_lik), _ptr(g_yhat), y_hat.numel(),
                                scale_bound, likelihood_bound, d_y.data_ptr(), d_s.data_ptr(), _ptr(d_m), _stream()), "gc_backward")
    return d_y, d_s, d_m


def gc_build_indexes(scales: torch.Tensor, table: torch.Tensor, scale_bound: float) -> torch.Tensor:
    _need_cuda(_f32(scales), _f32(table))
    idx = torch.empty(scales.shape, dtype=torch.int32, device=scales.device)
    check(lib.licos_gc_build_indexes(scales.data_ptr(), scales.numel(), table.data_ptr(), table.numel(), scale_bound,
                                     idx.data_ptr(), _stream()), "gc_build_indexes")
    return idx


def gc_symbols(y: torch.Tensor, means: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(_f32(y), means)
    sym = torch.empty(y.shape, dtype=torch.int32, device=y.device)
    check(lib.licos_gc_symbols(y.data_ptr(), _ptr(means), y.numel(), sym.data_ptr(), _stream()), "gc_symbols")
    return sym


# ---------------------------------------------------------------------------------------------
# reductions
# ---------------------------------------------------------------------------------------------

def sum_log(lik: torch.Tensor, acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    """acc (1-element float64 device tensor) += sum(ln(lik))."""
    _need_cuda(_f32(lik))
    if acc is None:
        acc = torch.zeros(1, dtype=torch.float64, device=lik.device)
    check(lib.licos_sum_log(lik.data_ptr(), lik.numel(), acc.data_ptr(), _stream()), "sum_log")
    return acc


def sum_sq_err(a: torch.Tensor, b: torch.Tensor, acc: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(_f32(a), _f32(b))
    if a.shape != b.shape:
        raise ValueError("shape mismatch")
    if acc is None:
        acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    check(lib.licos_sum_sq_err(a.data_ptr(), b.data_ptr(), a.numel(), acc.data_ptr(), _stream()), "sum_sq_err")
    return acc


def scaled_reciprocal(lik: torch.Tensor, coef: float, g: torch.Tensor) -> torch.Tensor:
    """coef * g / lik with g a one-element fp32 device tensor (no host sync)."""
    _need_cuda(_f32(lik), _f32(g))
    out = torch.empty_like(lik)
    check(lib.licos_scaled_reciprocal(lik.data_ptr(), lik.numel(), coef, g.data_ptr(), out.data_ptr(), _stream()),
          "scaled_reciprocal")
    return out


def scaled_diff(a: torch.Tensor, b: torch.Tensor, coef: float, g: torch.Tensor) -> torch.Tensor:
    _need_cuda(_f32(a), _f32(b), _f32(g))
    out = torch.empty_like(a)
    check(lib.licos_scaled_diff(a.data_ptr(), b.data_ptr(), a.numel(), coef, g.data_ptr(), out.data_ptr(), _stream()),
          "scaled_diff")
    return out


def raw_dn_to_unit(dn: torch.Tensor, dn_max: int = 4095, use_full_range: bool = False) -> torch.Tensor:
    """raw_image_folder.py:192-196: 12-bit digital numbers (uint16 / int16 storage, any shape) -> fp32 in [0, 1]."""
    _need_cuda(dn)
    if dn.dtype not in (torch.uint16, torch.int16):
        raise ValueError("raw digital numbers must be 16-bit integers")
    dn = dn.contiguous()
    out = torch.empty(dn.shape, dtype=torch.float32, device=dn.device)
    check(lib.licos_raw_dn_to_unit(dn.data_ptr(), dn.numel(), int(dn_max), 0 if use_full_range else 1, out.data_ptr(),
                                   _stream()), "raw_dn_to_unit")
    return out


def pixels_to_unit(x: torch.Tensor, int_max: int = 0, requant8: bool = False) -> torch.Tensor:
    """Integer tiles -> fp32 in [0, 1] with the standalone kernel: the values the fused first layer derives on the fly."""
    if x.dtype == torch.uint8:
        return raw_dn_to_unit(x.to(torch.int16), int_max or 255, use_full_range=True)
    return raw_dn_to_unit(x, int_max or 4095, use_full_range=not requant8)


_MSSSIM_WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def ms_ssim(a: torch.Tensor, b: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    """pytorch_msssim.ms_ssim(a, b, data_range) on the device (eval_utils.py:159-169): 0-dim fp32 tensor."""
    _need_cuda(_f32(a), _f32(b))
    if a.shape != b.shape or a.dim() != 4:
        raise ValueError("ms_ssim expects two (B, C, H, W) tensors of the same shape")
    B, C, H, W = a.shape
    if min(H, W) <= (11 - 1) * 2 ** 4:
        raise AssertionError("Image size should be larger than 160 due to the 4 downsamplings in ms-ssim")
    coords = torch.arange(11, dtype=torch.float32) - 5
    win = torch.exp(-(coords ** 2) / (2 * 1.5 ** 2))
    win = (win / win.sum()).contiguous()
    c1, c2 = (0.01 * data_range) ** 2, (0.03 * data_range) ** 2
    x, y = a.contiguous(), b.contiguous()
    bc = B * C
    terms = []
    for level in range(5):
        h, w = x.shape[-2:]
        ws = torch.empty(int(lib.licos_msssim_workspace_floats(bc, h, w)), dtype=torch.float32, device=a.device)
        sums = torch.zeros(bc, 2, dtype=torch.float64, device=a.device)
        check(lib.licos_msssim_level(x.data_ptr(), y.data_ptr(), bc, h, w, win.data_ptr(), c1, c2, ws.data_ptr(),
                                     sums.data_ptr(), _stream()), "msssim_level")
        means = sums / float((h - 10) * (w - 10))
        terms.append(torch.relu(means[:, 1] if level < 4 else means[:, 0]))
        if level < 4:
            ph, pw = h % 2, w % 2
            ho, wo = (h + 2 * ph - 2) // 2 + 1, (w + 2 * pw - 2) // 2 + 1
            nx = torch.empty((B, C, ho, wo), dtype=torch.float32, device=a.device)
            ny = torch.empty_like(nx)
            check(lib.licos_avgpool2(x.data_ptr(), bc, h, w, nx.data_ptr(), _stream()), "avgpool2")
            check(lib.licos_avgpool2(y.data_ptr(), bc, h, w, ny.data_ptr(), _stream()), "avgpool2")
            x, y = nx, ny
    stack = torch.stack(terms, dim=0)                                  # (5, B*C) float64
    wts = torch.tensor(_MSSSIM_WEIGHTS, dtype=torch.float64, device=a.device).view(-1, 1)
    return torch.prod(stack ** wts, dim=0).mean().float()


def weighted_sum2(a: torch.Tensor, b: torch.Tensor, wa: float, wb: float, out: Optional[torch.Tensor] = None):
    _need_cuda(_f32(a), _f32(b), out)
    if out is None:
        out = torch.empty_like(a)
    check(lib.licos_weighted_sum2(a.data_ptr(), b.data_ptr(), wa, wb, a.numel(), out.data_ptr(), _stream()),
          "weighted_sum2")
    return out


def scale_inplace(buf: torch.Tensor, w: float) -> torch.Tensor:
    _need_cuda(_f32(buf))
    check(lib.licos_scale_inplace(buf.data_ptr(), w, buf.numel(), _stream()), "scale_inplace")
    return buf


# ---------------------------------------------------------------------------------------------
# host integer path
# ---------------------------------------------------------------------------------------------

def pmf_to_quantized_cdf(pmf, precision: int = 16) -> np.ndarray:
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32))
    cdf = np.zeros(p.size + 1, dtype=np.uint32)
    rc = lib.licos_pmf_to_quantized_cdf(p.ctypes.data, p.size, precision, cdf.ctypes.data)
    if rc == -5:
        raise ValueError("Invalid `pmf`: negative, non-finite or all-zero")
    check(rc, "pmf_to_quantized_cdf")
    return cdf


def _i32(a) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32))


def rans_encode(symbols, indexes, cdfs, cdf_sizes, offsets) -> bytes:
    symbols, indexes = _i32(symbols).reshape(-1), _i32(indexes).reshape(-1)
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    if symbols.size != indexes.size or cdfs.ndim != 2:
        raise ValueError("bad symbols / indexes / cdfs")
    cap = 4 * (symbols.size * 12 + 16)
    out = np.empty(cap, dtype=np.uint8)
    n = lib.licos_rans_encode(symbols.ctypes.data, indexes.ctypes.data, symbols.size, cdfs.ctypes.data,
                              cdfs.shape[0], cdfs.shape[1], cdf_sizes.ctypes.data, offsets.ctypes.data,
                              out.ctypes.data, cap)
    check(int(n), "rans_encode")
    return out[:n].tobytes()


def rans_decode(encoded: bytes, indexes, cdfs, cdf_sizes, offsets) -> np.ndarray:
    indexes = _i32(indexes).reshape(-1)
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    enc = np.frombuffer(encoded, dtype=np.uint8)
    out = np.empty(indexes.size, dtype=np.int32)
    check(lib.licos_rans_decode(enc.ctypes.data, enc.size, indexes.ctypes.data, indexes.size, cdfs.ctypes.data,
                                cdfs.shape[0], cdfs.shape[1], cdf_sizes.ctypes.data, offsets.ctypes.data,
                                out.ctypes.data), "rans_decode")
    return out


def rans_encode_batch(symbols: np.ndarray, indexes: np.ndarray, cdfs, cdf_sizes, offsets, threads: int = 0):
    """symbols: (B, n) int32.  indexes: (B, n) or (n,) int32 (shared by every image).  -> list[bytes]."""
    symbols = _i32(symbols)
    B, n = symbols.shape
    indexes = _i32(indexes)
    stride = 0 if indexes.ndim == 1 else n
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    out_stride = 4 * (n * 12 + 16)
    # worst case is generous; shrink for big batches by encoding in slices
    out = np.empty((B, out_stride), dtype=np.uint8)
    sizes = np.zeros(B, dtype=np.int64)
    check(lib.licos_rans_encode_batch(symbols.ctypes.data, indexes.ctypes.data, B, n, stride, cdfs.ctypes.data,
                                      cdfs.shape[0], cdfs.shape[1], cdf_sizes.ctypes.data, offsets.ctypes.data,
                                      out.ctypes.data, out_stride, sizes.ctypes.data, threads), "rans_encode_batch")
    return [out[i, : sizes[i]].tobytes() for i in range(B)]


def rans_encode_device(symbols: torch.Tensor, indexes: Optional[torch.Tensor], n_spatial: int, cdfs: torch.Tensor,
                       cdf_sizes: torch.Tensor, offsets: torch.Tensor):
    """Device-resident rANS encoder.  symbols: int32 (B, ...) on the GPU; indexes: None (index = position // n_spatial),
    an int32 tensor shaped like symbols, or like one image (shared).  cdfs / cdf_sizes / offsets: int32 device tensors.
    Returns list[bytes], or None when a stream did not fit the scratch row (callers then use the host coder)."""
    _need_cuda(symbols, indexes, cdfs, cdf_sizes, offsets)
    B = symbols.shape[0]
    if B == 0:
        return []
    sym = symbols.reshape(B, -1).contiguous()
    n = sym.shape[1]
    idx, stride = None, 0
    if indexes is not None:
        idx = indexes.to(torch.int32).contiguous()
        stride = n if idx.numel() == sym.numel() else 0
        if idx.numel() not in (sym.numel(), n):
            raise ValueError("indexes must match symbols or one image of symbols")
    cdfs, cdf_sizes, offsets = cdfs.contiguous(), cdf_sizes.contiguous(), offsets.contiguous()
    dev = sym.device
    cap = n + 1024                                     # words per image: 32 bits per symbol + slack
    work = torch.empty((B, cap), dtype=torch.int32, device=dev)
    rcp = torch.empty(cdfs.numel(), dtype=torch.int64, device=dev)
    lengths = torch.empty(B, dtype=torch.int32, device=dev)
    check(lib.licos_rans_encode_device(sym.data_ptr(), _ptr(idx), stride, B, n, int(n_spatial), cdfs.data_ptr(),
                                       cdfs.shape[0], cdfs.shape[1], cdf_sizes.data_ptr(), offsets.data_ptr(),
                                       rcp.data_ptr(), work.data_ptr(), cap, lengths.data_ptr(), _stream()),
          "rans_encode_device")
    offs = torch.cumsum(lengths.clamp(min=0).to(torch.int64), 0)
    word_offsets = offs - lengths.clamp(min=0)
    host = torch.stack([lengths.to(torch.int64), word_offsets]).cpu()   # one small D2H (and the sync point)
    lens, starts = host[0].tolist(), host[1].tolist()
    if min(lens) < 0:
        return None
    total = int(starts[-1] + lens[-1])
    packed = torch.empty(total, dtype=torch.int32, device=dev)
    check(lib.licos_rans_pack_device(work.data_ptr(), cap, lengths.data_ptr(), word_offsets.data_ptr(), B,
                                     packed.data_ptr(), _stream()), "rans_pack_device")
    stage = _pinned_words(total)
    stage[:total].copy_(packed, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    view = memoryview(stage.numpy()).cast("B")
    return [view[4 * s: 4 * (s + l)].tobytes() for s, l in zip(starts, lens)]


_PINNED = {}


def _pinned_words(n: int) -> torch.Tensor:
    """A cached pinned int32 staging buffer of at least n words (the byte strings leave the device through it)."""
    buf = _PINNED.get("w")
    if buf is None or buf.numel() < n:
        buf = torch.empty(max(n, 1 << 20), dtype=torch.int32).pin_memory()
        _PINNED["w"] = buf
    return buf


def rans_decode_device(strings, indexes: Optional[torch.Tensor], n: int, n_spatial: int, cdfs: torch.Tensor,
                       cdf_sizes: torch.Tensor, offsets: torch.Tensor):
    """Device-resident rANS decoder: list[bytes] -> int32 (B, n) symbols on the GPU (one H2D copy of the strings).
    Returns None when a stream is not a whole number of 32-bit words or fails to decode (callers use the host coder)."""
    _need_cuda(indexes, cdfs, cdf_sizes, offsets)
    B = len(strings)
    dev = cdfs.device
    if B == 0:
        return torch.empty((0, n), dtype=torch.int32, device=dev)
    sizes = [len(s) for s in strings]
    if any(sz % 4 or sz < 8 for sz in sizes):
        return None
    n_words = np.asarray([sz // 4 for sz in sizes], dtype=np.int32)
    offs = np.zeros(B, dtype=np.int64)
    np.cumsum(n_words[:-1], out=offs[1:])
    total = int(n_words.sum())
    stage = _pinned_words(total)
    dst = memoryview(stage.numpy()).cast("B")
    at = 0
    for st in strings:  # straight into pinned memory: one host copy, then one async H2D
        dst[at: at + len(st)] = st
        at += len(st)
    packed = stage[:total].to(dev, non_blocking=True)
    offs_d, nw_d = torch.from_numpy(offs).to(dev), torch.from_numpy(n_words).to(dev)
    idx, stride = None, 0
    if indexes is not None:
        idx = indexes.to(torch.int32).contiguous()
        stride = n if idx.numel() == B * n else 0
        if idx.numel() not in (B * n, n):
            raise ValueError("indexes must have n entries per image (or n shared entries)")
    out = torch.empty((B, n), dtype=torch.int32, device=dev)
    status = torch.empty(B, dtype=torch.int32, device=dev)
    cdfs, cdf_sizes, offsets = cdfs.contiguous(), cdf_sizes.contiguous(), offsets.contiguous()
    check(lib.licos_rans_decode_device(packed.data_ptr(), offs_d.data_ptr(), nw_d.data_ptr(), _ptr(idx), stride, B, n,
                                       int(n_spatial), cdfs.data_ptr(), cdfs.shape[0], cdfs.shape[1], cdf_sizes.data_ptr(),
                                       offsets.data_ptr(), out.data_ptr(), status.data_ptr(), _stream()), "rans_decode_device")
    if int(status.min().item()) < 0:
        return None
    return out


def rans_decode_batch(strings, indexes: np.ndarray, n: int, cdfs, cdf_sizes, offsets, threads: int = 0) -> np.ndarray:
    B = len(strings)
    indexes = _i32(indexes)
    stride = 0 if indexes.ndim == 1 else n
    cdfs, cdf_sizes, offsets = _i32(cdfs), _i32(cdf_sizes), _i32(offsets)
    bufs = [np.frombuffer(s, dtype=np.uint8) for s in strings]
    ptrs = (ctypes.c_void_p * B)(*[b.ctypes.data for b in bufs])
    sizes = np.asarray([b.size for b in bufs], dtype=np.int64)
    out = np.empty((B, n), dtype=np.int32)
    check(lib.licos_rans_decode_batch(ctypes.cast(ptrs, ctypes.c_void_p), sizes.ctypes.data, indexes.ctypes.data, B, n,
                                      stride, cdfs.ctypes.data, cdfs.shape[0], cdfs.shape[1], cdf_sizes.ctypes.data,
                                      offsets.ctypes.data, out.ctypes.data, threads), "rans_decode_batch")
    return out
