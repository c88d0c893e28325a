"""The torch custom-op layer over the C ABI ("thin TORCH_LIBRARY wrappers translate tensors <->
pointers and status <-> exceptions"): every hot-path kernel is registered as ``torch.ops.licos_b200.<op>`` with a schema,
a CUDA implementation (``ops.py`` -> ctypes -> ``liblicos_b200.so``) and a Meta implementation (shapes / dtypes only, so
the ops trace and export).  The module code (layers.py, entropy_models.py, losses.py) calls the kernels through this
layer; ``ops.py`` + ``_lib.py`` stay the raw binding that tests/test_abi.py checks against include/licos_b200.h.

There is no CPU implementation registered: calling an op with CPU tensors raises (NotImplementedError from the
dispatcher) -- this package has no CPU path.

The functions at the bottom keep ``ops.py``'s call signatures (an ``EbPacked`` parameter block, optional outputs as None)
and unpack them into the flat tensor / int / float arguments a schema can carry.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import Tensor

from . import _lib, ops

_LIB = torch.library.Library("licos_b200", "DEF")
_EMPTY = {}


def _register(schema: str, cuda_impl, meta_impl=None) -> None:
    _LIB.define(schema)
    name = schema.split("(", 1)[0]
    _LIB.impl(name, cuda_impl, "CUDA")
    if meta_impl is not None:
        _LIB.impl(name, meta_impl, "Meta")


def _none(t: Optional[Tensor]) -> Optional[Tensor]:
    return None if t is None or t.numel() == 0 else t


def _empty(ref: Tensor, dtype=None) -> Tensor:
    return torch.empty(0, dtype=dtype or ref.dtype, device=ref.device)


def _ebp(packed: Tensor, medians: Tensor, widths: Sequence[int], form: int, bound: float, lut: Optional[Tensor] = None) -> ops.EbPacked:
    e = ops.EbPacked(packed, medians, list(widths), int(form), float(bound))
    e.lut = _none(lut)
    return e


# ---------------------------------------------------------------------------------------------------------------------
# convolutions
# ---------------------------------------------------------------------------------------------------------------------
def _conv_out_shape(x, kind, in_layout, out_layout, out_c):
    if in_layout == _lib.LAYOUT_NHWC_BF16:
        B, H, W, _ = x.shape
    else:
        B, _, H, W = x.shape
    if kind == _lib.CONV_5X5_S2:
        OH, OW = (H + 1) // 2, (W + 1) // 2
    elif kind == _lib.DECONV_5X5_S2:
        OH, OW = 2 * H, 2 * W
    else:
        OH, OW = H, W
    if out_layout == _lib.LAYOUT_NHWC_BF16:
        return (B, OH, OW, out_c), torch.bfloat16
    return (B, out_c, OH, OW), {_lib.LAYOUT_NCHW_U8: torch.uint8, _lib.LAYOUT_NCHW_U16: torch.uint16}.get(out_layout, torch.float32)


def _conv_forward_cuda(x, kind, epilogue, in_layout, out_layout, in_c, out_c, weight, bias, beta, gamma, int_max=0, pre_act=None):
    return ops.conv_forward(x, kind=kind, epilogue=epilogue, in_layout=in_layout, out_layout=out_layout, in_c=in_c, out_c=out_c,
                            weight=weight, bias=bias, beta=beta, gamma=gamma, int_max=int_max, pre_act=pre_act)


def _conv_forward_meta(x, kind, epilogue, in_layout, out_layout, in_c, out_c, weight, bias, beta, gamma, int_max=0, pre_act=None):
    shape, dtype = _conv_out_shape(x, kind, in_layout, out_layout, out_c)
    return torch.empty(shape, dtype=dtype, device=x.device)


_register("conv_forward(Tensor x, int kind, int epilogue, int in_layout, int out_layout, int in_c, int out_c, Tensor weight, "
          "Tensor? bias, Tensor? beta, Tensor? gamma, int int_max=0, Tensor(a!)? pre_act=None) -> Tensor", _conv_forward_cuda, _conv_forward_meta)

def _wgrad(small, big, kind, out):
    ops.conv_wgrad(small, big, kind, out=out)
    return out


_register("conv_wgrad(Tensor small, Tensor big, int kind, Tensor(a!) out) -> Tensor(a!)", _wgrad,
          lambda small, big, kind, out: out)


def _wgrad_image(small, image, out):
    ops.conv_wgrad_image(small, image, out)
    return out


_register("conv_wgrad_image(Tensor small, Tensor image, Tensor(a!) out) -> Tensor(a!)", _wgrad_image,
          lambda small, image, out: out)


_register("gdn_backward(Tensor x, Tensor g, Tensor gamma_hat, Tensor beta_hat, bool inverse, Tensor(a!) d_gamma_hat, "
          "Tensor(b!) d_beta_hat, Tensor(c!)? d_bias) -> Tensor",
          lambda x, g, gh, bh, inv, dgh, dbh, db: ops.gdn_backward(x, g, gh, bh, inv, dgh, dbh, db),
          lambda x, g, gh, bh, inv, dgh, dbh, db: torch.empty_like(x))

_register("nchw_to_nhwc_bf16(Tensor x, bool take_abs=False) -> Tensor",
          lambda x, take_abs=False: ops.nchw_to_nhwc_bf16(x, take_abs=take_abs),
          lambda x, take_abs=False: torch.empty((x.shape[0], x.shape[2], x.shape[3], x.shape[1]), dtype=torch.bfloat16, device=x.device))

# ---------------------------------------------------------------------------------------------------------------------
# entropy bottleneck
# ---------------------------------------------------------------------------------------------------------------------
_EB = "Tensor packed, Tensor medians, int[] widths, int form, float bound"


def _eb_eval_cuda(x, packed, medians, widths, form, bound):
    return list(ops.eb_forward_eval(_ebp(packed, medians, widths, form, bound), x))


_register(f"eb_forward_eval(Tensor x, {_EB}) -> Tensor[]", _eb_eval_cuda,
          lambda x, *a: [torch.empty_like(x), torch.empty_like(x)])


def _eb_fused_cuda(x, packed, medians, widths, form, bound, lut, want_float, want_symbols, want_symbols_i16, want_nhwc, sum_ln):
    ebp = _ebp(packed, medians, widths, form, bound, lut)
    res = ops.eb_forward_eval_fused(ebp, x, want_symbols=want_symbols, want_nhwc=want_nhwc, want_symbols_i16=True,
                                    sum_ln=sum_ln, want_float=want_float) if want_symbols_i16 else \
        ops.eb_forward_eval_fused(ebp, x, want_symbols=want_symbols, want_nhwc=want_nhwc, sum_ln=sum_ln, want_float=want_float) + (None,)
    y_hat, lik, sym, nhwc, sym16 = res
    return [y_hat if y_hat is not None else _empty(x), lik if lik is not None else _empty(x),
            sym if sym is not None else _empty(x, torch.int32), nhwc if nhwc is not None else _empty(x, torch.bfloat16),
            sym16 if sym16 is not None else _empty(x, torch.int16)]


def _eb_fused_meta(x, packed, medians, widths, form, bound, lut, want_float, want_symbols, want_symbols_i16, want_nhwc, sum_ln):
    e = lambda dt: torch.empty(0, dtype=dt, device=x.device)
    return [torch.empty_like(x) if want_float else e(x.dtype), torch.empty_like(x) if want_float else e(x.dtype),
            torch.empty(x.shape, dtype=torch.int32, device=x.device) if want_symbols else e(torch.int32),
            torch.empty((x.shape[0], x.shape[2], x.shape[3], x.shape[1]), dtype=torch.bfloat16, device=x.device) if want_nhwc else e(torch.bfloat16),
            torch.empty(x.shape, dtype=torch.int16, device=x.device) if want_symbols_i16 else e(torch.int16)]


_register(f"eb_eval_fused(Tensor x, {_EB}, Tensor lut, bool want_float, bool want_symbols, bool want_symbols_i16, bool want_nhwc, "
          "Tensor(a!)? sum_ln) -> Tensor[]", _eb_fused_cuda, _eb_fused_meta)

_register(f"eb_build_lut({_EB}) -> Tensor",
          lambda packed, medians, widths, form, bound: _ebp(packed, medians, widths, form, bound).eval_lut(),
          lambda packed, medians, widths, form, bound: torch.empty(packed.shape[0] * (2 * _lib.EB_LUT_RADIUS + 1), dtype=torch.float32,
                                                                   device=packed.device))

_register(f"eb_forward_noise(Tensor x, {_EB}, Tensor? noise, int seed) -> Tensor[]",
          lambda x, packed, medians, widths, form, bound, noise, seed: list(
              ops.eb_forward_noise(_ebp(packed, medians, widths, form, bound), x, noise, seed)),
          lambda x, *a: [torch.empty_like(x), torch.empty_like(x)])

_register(f"eb_backward(Tensor y_hat, {_EB}, Tensor? g_lik, Tensor? g_yhat) -> Tensor[]",
          lambda y_hat, packed, medians, widths, form, bound, g_lik, g_yhat: list(
              ops.eb_backward(_ebp(packed, medians, widths, form, bound), y_hat, g_lik, g_yhat)),
          lambda y_hat, packed, *a: [torch.empty_like(y_hat), torch.empty_like(packed)])

_register("eb_symbols(Tensor x, Tensor medians) -> Tensor", lambda x, medians: ops.eb_symbols(x, medians),
          lambda x, medians: torch.empty(x.shape, dtype=torch.int32, device=x.device))
_register("eb_dequantize(Tensor symbols, Tensor medians) -> Tensor", lambda s, medians: ops.eb_dequantize(s, medians),
          lambda s, medians: torch.empty(s.shape, dtype=torch.float32, device=s.device))

# ---------------------------------------------------------------------------------------------------------------------
# gaussian conditional, reductions
# ---------------------------------------------------------------------------------------------------------------------
_register("gc_forward(Tensor y, Tensor scales, Tensor? means, Tensor? noise, bool training, float scale_bound, "
          "float likelihood_bound, int seed) -> Tensor[]",
          lambda y, scales, means, noise, training, sb, lb, seed: list(
              ops.gc_forward(y, scales, means, noise, training=training, scale_bound=sb, likelihood_bound=lb, seed=seed)),
          lambda y, *a: [torch.empty_like(y), torch.empty_like(y)])


def _gc_backward_cuda(y_hat, scales, means, g_lik, g_yhat, sb, lb, want_means):
    d_y, d_s, d_m = ops.gc_backward(y_hat, scales, means, g_lik, g_yhat, sb, lb, want_means=want_means)
    return [d_y, d_s, d_m if d_m is not None else _empty(y_hat)]


_register("gc_backward(Tensor y_hat, Tensor scales, Tensor? means, Tensor? g_lik, Tensor? g_yhat, float scale_bound, "
          "float likelihood_bound, bool want_means) -> Tensor[]", _gc_backward_cuda,
          lambda y_hat, scales, means, g_lik, g_yhat, sb, lb, wm: [torch.empty_like(y_hat), torch.empty_like(y_hat),
                                                                   torch.empty_like(y_hat) if (wm and means is not None) else _empty(y_hat)])

_register("gc_build_indexes(Tensor scales, Tensor table, float scale_bound) -> Tensor",
          lambda scales, table, sb: ops.gc_build_indexes(scales, table, sb),
          lambda scales, table, sb: torch.empty(scales.shape, dtype=torch.int32, device=scales.device))
_register("gc_symbols(Tensor y, Tensor? means) -> Tensor", lambda y, means: ops.gc_symbols(y, means),
          lambda y, means: torch.empty(y.shape, dtype=torch.int32, device=y.device))

_register("sum_log(Tensor lik, Tensor(a!) acc) -> Tensor(a!)", lambda lik, acc: ops.sum_log(lik, acc), lambda lik, acc: acc)
_register("sum_sq_err(Tensor a, Tensor b, Tensor(a!) acc) -> Tensor(a!)", lambda a, b, acc: ops.sum_sq_err(a, b, acc),
          lambda a, b, acc: acc)

T = torch.ops.licos_b200


# ---------------------------------------------------------------------------------------------------------------------
# ops.py-compatible entry points for the module code
# ---------------------------------------------------------------------------------------------------------------------
def conv_forward(x: Tensor, *, kind: int, epilogue: int, in_layout: int, out_layout: int, in_c: int, out_c: int, weight: Tensor,
                 bias: Optional[Tensor], beta: Optional[Tensor] = None, gamma: Optional[Tensor] = None, int_max: int = 0,
                 pre_act: Optional[Tensor] = None) -> Tensor:
    return T.conv_forward(x, kind, epilogue, in_layout, out_layout, in_c, out_c, weight, bias, beta, gamma, int_max, pre_act)


def conv_wgrad_image(small: Tensor, image: Tensor, out: Tensor) -> Tensor:
    T.conv_wgrad_image(small, image, out)
    return out.view(small.shape[-1], -1)


def conv_wgrad(small: Tensor, big: Tensor, kind: int, out: Tensor) -> Tensor:
    taps = {_lib.CONV_5X5_S2: 25, _lib.DECONV_5X5_S2: 25, _lib.CONV_3X3_S1: 9, _lib.CONV_1X1: 1}[kind]
    T.conv_wgrad(small, big, kind, out)
    return out.view(taps, small.shape[-1], big.shape[-1])


def gdn_backward(x, g, gamma_hat, beta_hat, inverse, d_gamma_hat, d_beta_hat, d_bias):
    return T.gdn_backward(x, g, gamma_hat, beta_hat, bool(inverse), d_gamma_hat, d_beta_hat, d_bias)


def nchw_to_nhwc_bf16(x: Tensor, take_abs: bool = False) -> Tensor:
    return T.nchw_to_nhwc_bf16(x, take_abs)


def _eb_args(ebp: ops.EbPacked):
    p = ebp.p
    return ebp.packed, ebp.medians, [int(p.widths[i]) for i in range(p.n_layers + 1)], int(p.form), float(p.likelihood_bound)


def eb_forward_eval(ebp: ops.EbPacked, x: Tensor):
    y_hat, lik = T.eb_forward_eval(x, *_eb_args(ebp))
    return y_hat, lik


def eb_forward_eval_fused(ebp: ops.EbPacked, x: Tensor, want_symbols=False, want_nhwc=False, want_symbols_i16=False,
                          sum_ln: Optional[Tensor] = None, want_float=True):
    if ebp.lut is None or ops.capture_rebuild():
        ebp.lut = T.eb_build_lut(*_eb_args(ebp))
    y_hat, lik, sym, nhwc, sym16 = T.eb_eval_fused(x, *_eb_args(ebp), ebp.lut, want_float, want_symbols, want_symbols_i16,
                                                   want_nhwc, sum_ln)
    res = (_none(y_hat), _none(lik), _none(sym), _none(nhwc))
    return res + (_none(sym16),) if want_symbols_i16 else res


def eb_forward_noise(ebp: ops.EbPacked, x: Tensor, noise: Optional[Tensor], seed: int = 0):
    y_hat, lik = T.eb_forward_noise(x, *_eb_args(ebp), noise, int(seed) & (2 ** 63 - 1))
    return y_hat, lik


def eb_backward(ebp: ops.EbPacked, y_hat: Tensor, g_lik: Optional[Tensor], g_yhat: Optional[Tensor]):
    d_x, d_packed = T.eb_backward(y_hat, *_eb_args(ebp), g_lik, g_yhat)
    return d_x, d_packed


def eb_symbols(x: Tensor, medians: Tensor) -> Tensor:
    return T.eb_symbols(x, medians)


def eb_dequantize(sym: Tensor, medians: Tensor) -> Tensor:
    return T.eb_dequantize(sym, medians)


def gc_forward(y, scales, means=None, noise=None, *, training=False, scale_bound=0.11, likelihood_bound=1e-9, seed: int = 0):
    y_hat, lik = T.gc_forward(y, scales, means, noise, bool(training), float(scale_bound), float(likelihood_bound),
                              int(seed) & (2 ** 63 - 1))
    return y_hat, lik


def gc_backward(y_hat, scales, means, g_lik, g_yhat, scale_bound, likelihood_bound, want_means=False):
    d_y, d_s, d_m = T.gc_backward(y_hat, scales, means, g_lik, g_yhat, float(scale_bound), float(likelihood_bound), bool(want_means))
    return d_y, d_s, _none(d_m)


def gc_build_indexes(scales: Tensor, table: Tensor, scale_bound: float) -> Tensor:
    return T.gc_build_indexes(scales, table, float(scale_bound))


def gc_symbols(y: Tensor, means: Optional[Tensor] = None) -> Tensor:
    return T.gc_symbols(y, means)


def sum_log(lik: Tensor, acc: Optional[Tensor] = None) -> Tensor:
    if acc is None:
        acc = torch.zeros(1, dtype=torch.float64, device=lik.device)
    return T.sum_log(lik, acc)


def sum_sq_err(a: Tensor, b: Tensor, acc: Optional[Tensor] = None) -> Tensor:
    if acc is None:
        acc = torch.zeros(1, dtype=torch.float64, device=a.device)
    return T.sum_sq_err(a, b, acc)


OPS: List[str] = ["conv_forward", "conv_wgrad", "conv_wgrad_image", "gdn_backward", "nchw_to_nhwc_bf16", "eb_forward_eval", "eb_eval_fused", "eb_build_lut",
                  "eb_forward_noise", "eb_backward", "eb_symbols", "eb_dequantize", "gc_forward", "gc_backward", "gc_build_indexes",
                  "gc_symbols", "sum_log", "sum_sq_err"]
