"""Host-side mirrors of compressai.entropy_models.{EntropyModel, EntropyBottleneck, GaussianConditional}
 on top of the C ABI.

Division of labour:
  * forward() with autograd off -> one fused CUDA pass (licos_eb_forward_* / licos_gc_forward);
  * quantize("symbols") / build_indexes / dequantize -> bit-exact integer kernels;
  * update() -> host logic exactly as upstream (torch CPU ops for the tiny PMF tables) + the C++
    licos_pmf_to_quantized_cdf; run once per model (eval_script.py:72,88);
  * compress()/decompress() -> symbols, indexes and the rANS coder stay on the GPU (csrc/rans_device.cu, bitstream-identical
    to the host coder in csrc/host_codec.cpp, which remains the fallback for streams that overflow the scratch row);
  * forward() with autograd on -> the noise (training) or rounding (eval) kernel forward, licos_eb_backward /
    licos_gc_backward back (training step, train.py:190-193).  There is no torch-expression path on the device:
    inputs the kernels do not take (non-fp32, fewer than 2 dims) raise.
"""
from __future__ import annotations

import math
import os
from typing import Optional, Tuple

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import _lib, ops
from . import torch_ops as T
from .layers import LowerBound


def _require_cuda(t: Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"licos_b200: {what} needs CUDA tensors on a B200 (this package has no CPU path)")


def _wants_grad(module: nn.Module, *tensors: Optional[Tensor]) -> bool:
    if not torch.is_grad_enabled():
        return False
    if any(t is not None and t.requires_grad for t in tensors):
        return True
    return any(p.requires_grad for p in module.parameters())


EB_NATIVE_BACKWARD_MAX_PPC = 320  # kEbMaxPpc in csrc/entropy.cu: filters up to (13, 13, 3, 3) (298 parameters per channel)


def _split_raw(eb, raw):
    """raw = (_matrix0, _bias0, _factor0, _matrix1, ...) -> (matrices, biases, factors)"""
    n_layers = len(eb.filters) + 1
    ms, bs, fs, k = [], [], [], 0
    for i in range(n_layers):
        ms.append(raw[k]); bs.append(raw[k + 1])
        k += 2
        if i < n_layers - 1:
            fs.append(raw[k])
            k += 1
    return ms, bs, fs


def _native_packed(eb, raw) -> ops.EbPacked:
    ms, bs, fs = _split_raw(eb, [t.detach().contiguous() for t in raw])
    widths = (1,) + tuple(eb.filters) + (1,)
    packed = ops.eb_pack_params(ms, bs, fs, widths)
    form = _lib.EB_FORM_PLAIN if eb.likelihood_form == "plain" else _lib.EB_FORM_STABLE
    return ops.EbPacked(packed, eb.quantiles.detach()[:, 0, 1].contiguous(), widths, form,
                        eb.likelihood_bound if eb.use_likelihood_bound else 0.0)


class _EbTrainFn(torch.autograd.Function):
    """EntropyBottleneck.forward(x, training=True) with its backward on the device: parameter block in one launch, the
    noise kernel, and a backward kernel that re-evaluates the density network per element and sweeps it in reverse
    (d x, per-channel parameter gradients reduced in the block), then one launch back to the raw parameters."""

    @staticmethod
    def forward(ctx, eb, x, noise, seed, training, *raw):
        ebp = _native_packed(eb, raw)
        if not training:
            # eval mode under autograd (e.g. a validation loss that is differentiated): rounding has zero gradient with
            # respect to x, the density parameters get theirs from the same backward kernel evaluated at y_hat
            y_hat, lik = T.eb_forward_eval(ebp, x.detach().contiguous())
        else:
            if noise is None and seed is None:
                # device-side draw from torch's generator: no host sync in the training step, legal under graph capture
                noise = torch.empty_like(x).uniform_(-0.5, 0.5)
            y_hat, lik = T.eb_forward_noise(ebp, x.detach().contiguous(), None if noise is None else noise.contiguous(),
                                              0 if seed is None else seed)
        ctx.eb, ctx.ebp, ctx.training = eb, ebp, bool(training)
        ctx.save_for_backward(y_hat, *raw)
        return y_hat, lik

    @staticmethod
    def backward(ctx, g_yhat, g_lik):
        eb = ctx.eb
        y_hat, *raw = ctx.saved_tensors
        g_yhat = None if g_yhat is None else g_yhat.contiguous()
        g_lik = None if g_lik is None else g_lik.contiguous()
        d_x, d_packed = T.eb_backward(ctx.ebp, y_hat, g_lik, g_yhat)
        if not ctx.training:
            d_x = torch.zeros_like(d_x)  # y_hat = round(x - median) + median
        ms, bs, fs = _split_raw(eb, [t.detach().contiguous() for t in raw])
        gm, gb, gf = ops.eb_param_grads(ms, bs, fs, d_packed, (1,) + tuple(eb.filters) + (1,))
        grads = []
        for i in range(len(ms)):
            grads += [gm[i], gb[i]] + ([gf[i]] if i < len(fs) else [])
        return (None, d_x, None, None, None, *grads)


class _EbAuxLossFn(torch.autograd.Function):
    """EntropyBottleneck.loss(): sum |logits_cumulative(quantiles, stop_gradient=True) - target| and its gradient with
    respect to the quantiles, two launches (parameter block, loss kernel)."""

    @staticmethod
    def forward(ctx, eb, quantiles, *raw):
        ebp = _native_packed(eb, raw)
        loss, d_q = ops.eb_aux_loss(ebp.packed, (1,) + tuple(eb.filters) + (1,), quantiles.detach().contiguous(),
                                    eb.target.detach().float().contiguous())
        ctx.save_for_backward(d_q)
        ctx.n_raw = len(raw)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (d_q,) = ctx.saved_tensors
        return (None, d_q * g) + (None,) * ctx.n_raw


class _GcTrainFn(torch.autograd.Function):
    """GaussianConditional.forward(y, scales, means, training=True): the noise kernel forward, one elementwise kernel back
    (gradients of y, the scales and, when given, the means; LowerBound's rule on both bounds)."""

    @staticmethod
    def forward(ctx, gc, y, scales, means, noise, training):
        if noise is None and training:
            noise = torch.empty_like(y).uniform_(-0.5, 0.5)  # device-side draw: no host sync, graph-capture safe
        y, scales = y.detach().contiguous(), scales.detach().contiguous()
        means = None if means is None else means.detach().contiguous()
        bound = gc.likelihood_bound if gc.use_likelihood_bound else 0.0
        y_hat, lik = T.gc_forward(y, scales, means, noise.contiguous() if training else None, training=bool(training),
                                    scale_bound=gc._scale_bound_f, likelihood_bound=bound)
        ctx.gc, ctx.bound, ctx.has_means, ctx.training = gc, bound, means is not None, bool(training)
        ctx.save_for_backward(y_hat, scales, *([means] if means is not None else []))
        return y_hat, lik

    @staticmethod
    def backward(ctx, g_yhat, g_lik):
        y_hat, scales, *rest = ctx.saved_tensors
        means = rest[0] if ctx.has_means else None
        d_y, d_s, d_m = T.gc_backward(y_hat, scales, means, None if g_lik is None else g_lik.contiguous(),
                                        None if g_yhat is None else g_yhat.contiguous(), ctx.gc._scale_bound_f, ctx.bound,
                                        want_means=ctx.has_means and ctx.needs_input_grad[3])
        if not ctx.training:
            # y_hat = round(y - means) + means: no gradient reaches y, the means only see the pass-through of y_hat
            d_y = torch.zeros_like(d_y)
            if d_m is not None:
                d_m = g_yhat.contiguous() if g_yhat is not None else torch.zeros_like(d_m)
        return None, d_y, d_s, d_m, None, None


class EntropyModel(nn.Module):
    def __init__(self, likelihood_bound: float = 1e-9, entropy_coder: Optional[str] = None,
                 entropy_coder_precision: int = 16):
        super().__init__()
        if entropy_coder not in (None, "ans"):
            raise ValueError(f'Unknown entropy coder "{entropy_coder}" (only "ans" is built in)')
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.likelihood_bound = float(likelihood_bound)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())
        self.coder_threads = 0  # 0 = all host cores
        self.device_coder = True  # compress(): rANS-encode on the GPU (same bitstream); False = host threads

    offset = property(lambda self: self._offset)
    quantized_cdf = property(lambda self: self._quantized_cdf)
    cdf_length = property(lambda self: self._cdf_length)

    # ------------------------------------------------------------------ quantisation
    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            return inputs + torch.empty_like(inputs).uniform_(-0.5, 0.5)
        if mode == "symbols":
            _require_cuda(inputs, "quantize('symbols')")
            m = None if means is None else means.expand_as(inputs).contiguous()
            return T.gc_symbols(inputs.detach().contiguous(), m)
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if means is not None:
            outputs += means
        return outputs

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None, dtype: torch.dtype = torch.float) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs

    # ------------------------------------------------------------------ tables
    def _pmf_to_cdf(self, pmf: Tensor, tail_mass: Tensor, pmf_length: Tensor, max_length: int) -> Tensor:
        pmf_np = pmf.detach().cpu().float().numpy()
        tail_np = tail_mass.detach().cpu().float().numpy().reshape(len(pmf_np), -1)
        lengths = pmf_length.cpu().numpy()
        cdf = np.zeros((len(lengths), max_length + 2), dtype=np.int32)
        for i in range(len(lengths)):
            prob = np.concatenate((pmf_np[i, : lengths[i]], tail_np[i, :1]))
            row = ops.pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, : row.size] = row.astype(np.int32)
        return torch.from_numpy(cdf)

    def _check_tables(self) -> None:
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if self._quantized_cdf.dim() != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if self._offset.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if self._cdf_length.dim() != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def _host_tables(self):
        return (self._quantized_cdf.cpu().numpy(), self._cdf_length.cpu().numpy(), self._offset.cpu().numpy())

    # ------------------------------------------------------------------ range coding
    def compress(self, inputs: Tensor, indexes: Tensor, means: Optional[Tensor] = None):
        if inputs.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_tables()
        symbols = self.quantize(inputs, "symbols", means)
        B = symbols.size(0)
        if symbols.is_cuda and self.device_coder:
            # device-resident front end: symbols and indexes never leave the GPU, only the byte strings do
            out = ops.rans_encode_device(symbols.reshape(B, -1), indexes.reshape(B, -1), 1, self._quantized_cdf,
                                         self._cdf_length, self._offset)
            if out is not None:
                return out
        sym = symbols.reshape(B, -1).cpu().numpy()
        idx = indexes.reshape(B, -1).int().cpu().numpy()
        cdf, lengths, offsets = self._host_tables()
        return ops.rans_encode_batch(sym, idx, cdf, lengths, offsets, threads=self.coder_threads)

    def decompress(self, strings, indexes: Tensor, dtype: torch.dtype = torch.float, means: Optional[Tensor] = None):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if indexes.dim() < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_tables()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, indexes.dim()):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        B = indexes.size(0)
        n = indexes[0].numel() if B else 0
        if indexes.is_cuda and self.device_coder and B:
            sym_d = ops.rans_decode_device(list(strings), indexes.reshape(B, -1), n, 1, self._quantized_cdf,
                                           self._cdf_length, self._offset)
            if sym_d is not None:
                return self.dequantize(sym_d.reshape(indexes.size()), means, dtype)
        cdf, lengths, offsets = self._host_tables()
        idx = indexes.reshape(B, -1).int().cpu().numpy()
        sym = ops.rans_decode_batch(list(strings), idx, n, cdf, lengths, offsets, threads=self.coder_threads)
        outputs = torch.from_numpy(sym).reshape(indexes.size()).to(self._quantized_cdf.device)
        return self.dequantize(outputs, means, dtype)


class EntropyBottleneck(EntropyModel):
    """Fully-factorized density (Balle et al. 2018).  ``filters`` may be any widths <= 16: LICOS uses
    ``(in_channels, in_channels, 3, 3)`` (/root/reference/licos/model_utils.py:25-29)."""

    # "plain" = CompressAI >= 1.2 (forward and update() both call the tuple-returning _likelihood: sigmoid difference);
    # "stable" = the sign-stabilised form of <= 1.1.x (forward and update() alike).  One switch drives both, as upstream:
    # the integer tables update() writes must describe the same density forward() evaluates, or strings coded by one
    # release are not decodable by the other on the channels where the two forms round differently.
    likelihood_form = "plain"

    def __init__(self, channels: int, *args, tail_mass: float = 1e-9, init_scale: float = 10,
                 filters: Tuple[int, ...] = (3, 3, 3, 3), **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        if len(self.filters) + 1 > _lib.EB_MAX_LAYERS or max(self.filters) > 16:
            raise ValueError("EntropyBottleneck: at most 7 hidden layers of width <= 16 are supported")

        widths = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / widths[i + 1]))
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(torch.full((channels, widths[i + 1], widths[i]), float(init))))
            self.register_parameter(f"_bias{i:d}", nn.Parameter(torch.empty(channels, widths[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i:d}", nn.Parameter(torch.zeros(channels, widths[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.Tensor([-self.init_scale, 0, self.init_scale]).repeat(channels, 1, 1))
        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))
        self._packed_key = None
        self._packed = None

    # accept the ParameterList naming of newer CompressAI releases
    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        for old, new in (("matrices.", "_matrix"), ("biases.", "_bias"), ("factors.", "_factor"),
                         ("_matrices.", "_matrix"), ("_biases.", "_bias"), ("_factors.", "_factor")):
            for key in [k for k in state_dict if k.startswith(prefix + old)]:
                state_dict[prefix + new + key[len(prefix + old):]] = state_dict.pop(key)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def _get_medians(self) -> Tensor:
        return self.quantiles[:, :, 1:2]

    def __getstate__(self):
        # the kernel parameter block holds a ctypes struct with raw device pointers: a cache, never part of a copy / pickle
        # (copy.deepcopy(net) for EMA / best-model snapshots, torch.save(net), DataParallel replicas)
        state = self.__dict__.copy()
        state["_packed"] = state["_packed_key"] = None
        return state

    # ------------------------------------------------------------------ host expressions (update(), CPU-side loss value)
    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                matrix, bias = matrix.detach(), bias.detach()
            logits = torch.matmul(F.softplus(matrix), logits) + bias
            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs: Tensor, stop_gradient: bool = False, form: Optional[str] = None):
        lower = self._logits_cumulative(inputs - 0.5, stop_gradient=stop_gradient)
        upper = self._logits_cumulative(inputs + 0.5, stop_gradient=stop_gradient)
        if (form or self.likelihood_form) == "plain":
            likelihood = torch.sigmoid(upper) - torch.sigmoid(lower)
        else:
            sign = -torch.sign(lower + upper).detach()
            likelihood = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))
        return likelihood, lower, upper

    def loss(self) -> Tensor:
        if (self.quantiles.is_cuda and torch.is_grad_enabled() and self.quantiles.requires_grad
                and self.quantiles.dtype == torch.float32):
            return _EbAuxLossFn.apply(self, self.quantiles, *self._params()[:-1])
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def update(self, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        device = self.quantiles.device
        host = self if device.type == "cpu" else _cpu_twin(self)
        with torch.no_grad():
            q = host.quantiles
            medians = q[:, 0, 1]
            minima = torch.clamp(torch.ceil(medians - q[:, 0, 0]).int(), min=0)
            maxima = torch.clamp(torch.ceil(q[:, 0, 2] - medians).int(), min=0)
            pmf_start = medians - minima
            pmf_length = maxima + minima + 1
            max_length = int(pmf_length.max().item())
            samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
            pmf, lower, upper = host._likelihood(samples, stop_gradient=True, form=self.likelihood_form)
            pmf = pmf[:, 0, :]
            tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
            quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._offset = (-minima).to(device)
        self._quantized_cdf = quantized_cdf.to(device)
        self._cdf_length = (pmf_length + 2).to(device)
        return True

    # ------------------------------------------------------------------ kernel parameter block
    def _params(self):
        names = []
        for i in range(len(self.filters) + 1):
            names += [f"_matrix{i:d}", f"_bias{i:d}"] + ([f"_factor{i:d}"] if i < len(self.filters) else [])
        return [getattr(self, n) for n in names] + [self.quantiles]

    def train(self, mode: bool = True):
        self._packed_key = None  # fused optimizers update parameters without bumping the versions the cache is keyed on
        return super().train(mode)

    def packed_params(self, force: bool = False) -> ops.EbPacked:
        key = tuple((p.data_ptr(), p._version) for p in self._params()) + (self.likelihood_form,)
        if key != self._packed_key or force or ops.capture_rebuild():
            with torch.no_grad():
                C = self.channels
                parts = []
                for i in range(len(self.filters) + 1):
                    parts.append(F.softplus(getattr(self, f"_matrix{i:d}")).reshape(C, -1))
                    parts.append(getattr(self, f"_bias{i:d}").reshape(C, -1))
                    if i < len(self.filters):
                        parts.append(torch.tanh(getattr(self, f"_factor{i:d}")).reshape(C, -1))
                packed = torch.cat(parts, dim=1).contiguous().float()
                medians = self.quantiles[:, 0, 1].contiguous().float()
            form = _lib.EB_FORM_PLAIN if self.likelihood_form == "plain" else _lib.EB_FORM_STABLE
            self._packed = ops.EbPacked(packed, medians, (1,) + self.filters + (1,), form,
                                        self.likelihood_bound if self.use_likelihood_bound else 0.0)
            self._packed_key = key
        return self._packed

    # ------------------------------------------------------------------ forward
    def forward(self, x: Tensor, training: Optional[bool] = None, noise: Optional[Tensor] = None,
                seed: Optional[int] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training
        _require_cuda(x, "EntropyBottleneck.forward")
        if _wants_grad(self, x):
            if x.dim() < 2 or x.dtype != torch.float32:
                raise NotImplementedError("licos_b200: EntropyBottleneck under autograd takes float32 (B, C, ...) tensors")
            ppc = sum(p[0].numel() for p in self._params()[:-1])
            if ppc > EB_NATIVE_BACKWARD_MAX_PPC:
                # said here, at forward time, not from inside autograd's backward in the middle of a training step
                raise NotImplementedError(
                    f"licos_b200: the backward kernel holds at most {EB_NATIVE_BACKWARD_MAX_PPC} density parameters per "
                    f"channel; filters={self.filters} need {ppc} (LICOS uses (c, c, 3, 3) with c in 1, 3, 13)")
            return _EbTrainFn.apply(self, x, noise, seed, training, *self._params()[:-1])
        x = x.contiguous()
        ebp = self.packed_params()
        if not training:
            return T.eb_forward_eval(ebp, x)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if noise is None else 0
        return T.eb_forward_noise(ebp, x, None if noise is None else noise.contiguous(), seed)

    def forward_fused(self, x: Tensor, want_symbols: bool = False, want_nhwc: bool = True, want_symbols_i16: bool = False,
                      sum_ln: Optional[Tensor] = None, want_float: bool = True):
        """Eval-mode forward in ONE pass over x: ``(y_hat, likelihoods, symbols | None, y_hat as bf16 NHWC | None
        [, int16 symbols])``.  ``symbols`` is what ``quantize(x, "symbols", medians)`` returns; the NHWC copy is the
        layout the fused synthesis transform reads (``FusedSequential.forward(y_hat, nhwc=...)``); ``sum_ln`` (float64
        device scalar) accumulates sum(ln(likelihoods)), the rate term of compute_bpp.  Autograd-free path only."""
        _require_cuda(x, "EntropyBottleneck.forward_fused")
        if _wants_grad(self, x):
            raise RuntimeError("forward_fused is an inference path: call it under torch.no_grad()")
        return T.eb_forward_eval_fused(self.packed_params(), x.contiguous(), want_symbols=want_symbols,
                                         want_nhwc=want_nhwc, want_symbols_i16=want_symbols_i16, sum_ln=sum_ln,
                                         want_float=want_float)

    # ------------------------------------------------------------------ integer path
    @staticmethod
    def _build_indexes(size) -> Tensor:
        dims = len(size)
        view = [1] * dims
        view[1] = -1
        return torch.arange(size[1]).view(*view).int().repeat(size[0], 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor: Tensor, n: int) -> Tensor:
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def symbols(self, x: Tensor) -> Tensor:
        """round(x - medians) as int32, on the device (the quantiser the range coder consumes)."""
        return T.eb_symbols(x.contiguous(), self.packed_params().medians)

    def compress(self, x: Tensor):
        if x.dim() < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        self._check_tables()
        B, C = x.size(0), x.size(1)
        n_spatial = int(np.prod(x.shape[2:])) if x.dim() > 2 else 1
        sym_dev = self.symbols(x)
        if self.device_coder:
            out = ops.rans_encode_device(sym_dev.reshape(B, -1), None, n_spatial, self._quantized_cdf, self._cdf_length,
                                         self._offset)
            if out is not None:
                return out
        sym = sym_dev.reshape(B, -1).cpu().numpy()
        idx = np.repeat(np.arange(C, dtype=np.int32), n_spatial)  # same index plane for every image
        cdf, lengths, offsets = self._host_tables()
        return ops.rans_encode_batch(sym, idx, cdf, lengths, offsets, threads=self.coder_threads)

    def decompress(self, strings, size):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        self._check_tables()
        C = self._quantized_cdf.size(0)
        n_spatial = int(np.prod(size)) if len(size) else 1
        _require_cuda(self.quantiles, "EntropyBottleneck.decompress")
        symbols = None
        if self.device_coder and len(strings):
            symbols = ops.rans_decode_device(list(strings), None, C * n_spatial, n_spatial, self._quantized_cdf,
                                             self._cdf_length, self._offset)
            if symbols is not None:
                symbols = symbols.reshape(len(strings), C, *size)
        if symbols is None:
            idx = np.repeat(np.arange(C, dtype=np.int32), n_spatial)
            cdf, lengths, offsets = self._host_tables()
            sym = ops.rans_decode_batch(list(strings), idx, C * n_spatial, cdf, lengths, offsets,
                                        threads=self.coder_threads)
            symbols = torch.from_numpy(sym).reshape(len(strings), C, *size).to(self.quantiles.device)
        return T.eb_dequantize(symbols.contiguous(), self.packed_params().medians)


def _cpu_twin(eb: EntropyBottleneck) -> EntropyBottleneck:
    """CPU copy of the parameters: update() evaluates the (tiny) PMF grid with the same torch CPU ops
    upstream uses, so the integer tables do not depend on where the model lives."""
    twin = EntropyBottleneck(eb.channels, filters=eb.filters, tail_mass=eb.tail_mass, init_scale=eb.init_scale)
    with torch.no_grad():
        for name, p in eb.named_parameters():
            getattr(twin, name).copy_(p.detach().cpu())
    return twin


SCALES_MIN, SCALES_MAX, SCALES_LEVELS = 0.11, 256, 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class GaussianConditional(EntropyModel):
    def __init__(self, scale_table, *args, scale_bound: float = 0.11, tail_mass: float = 1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = scale_table[0]
        if scale_bound is None or scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer("scale_table", self._prepare_scale_table(scale_table) if scale_table else torch.Tensor())
        self.register_buffer("scale_bound", torch.Tensor([float(scale_bound)]))
        self._scale_bound_f = float(torch.tensor(float(scale_bound), dtype=torch.float32))

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs: Tensor) -> Tensor:
        return 0.5 * torch.erfc(float(-(2 ** -0.5)) * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force: bool = False) -> bool:
        if self._offset.numel() > 0 and not force:
            return False
        device = self.scale_table.device
        self.scale_table = self._prepare_scale_table(scale_table).to(device)
        self.update()
        return True

    def update(self) -> None:
        device = self.scale_table.device
        table = self.scale_table.detach().cpu()
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(torch.max(pmf_length).item())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        samples_scale = table.unsqueeze(1).float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower
        tail_mass = 2 * lower[:, :1]
        self._quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(device)
        self._offset = (-pmf_center).to(device)
        self._cdf_length = (pmf_length + 2).to(device)

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        values = inputs - means if means is not None else inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((0.5 - values) / scales)
        lower = self._standardized_cumulative((-0.5 - values) / scales)
        return upper - lower

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None, noise: Optional[Tensor] = None, seed: Optional[int] = None):
        if training is None:
            training = self.training
        _require_cuda(inputs, "GaussianConditional.forward")
        if _wants_grad(self, inputs, scales, means):
            if inputs.dtype != torch.float32 or scales.shape != inputs.shape or (means is not None and means.shape != inputs.shape):
                raise NotImplementedError("licos_b200: GaussianConditional under autograd takes float32 inputs with scales "
                                          "(and means) of the same shape")
            return _GcTrainFn.apply(self, inputs, scales, means, noise, training)
        if seed is None:
            seed = int(torch.randint(0, 2 ** 62, (1,)).item()) if (training and noise is None) else 0
        m = None if means is None else means.expand_as(inputs).contiguous()
        return T.gc_forward(inputs.contiguous(), scales.contiguous(), m,
                              None if noise is None else noise.contiguous(), training=bool(training),
                              scale_bound=self._scale_bound_f,
                              likelihood_bound=self.likelihood_bound if self.use_likelihood_bound else 0.0,
                              seed=seed)

    def build_indexes(self, scales: Tensor) -> Tensor:
        _require_cuda(scales, "GaussianConditional.build_indexes")
        return T.gc_build_indexes(scales.detach().contiguous(), self.scale_table.contiguous(), self._scale_bound_f)
