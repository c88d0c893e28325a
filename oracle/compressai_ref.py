"""Pure-PyTorch CPU restatement of the CompressAI modules LICOS drives.

TEST INFRASTRUCTURE (see oracle/__init__.py).  PARITY UNPINNED.

What the reference calls and where (all line numbers in /root/reference):
  * zoo lookup + module surgery ............ licos/model_utils.py:19-45
  * model(d), criterion, aux_loss .......... licos/train.py:190-200, 286-289
  * net.forward / net.compress ............. eval_utils.py:199-201
  * net.update() ........................... eval_script.py:72, 88
  * net_aux_optimizer ...................... licos/utils.py:65-73
  * state_dict arithmetic .................. licos/federation_utils.py:47-53

The arithmetic itself is upstream CompressAI (>= 1.2.0, un-pinned;
SURVEY.md section 8a rows A1-A13); module, attribute and state_dict names
follow upstream so that the same state_dict loads into this oracle and into
``licos_b200``.

Everything here is deliberately the slow, eager, one-op-per-line form.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import scipy.stats
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import cdf_rans

# ----------------------------------------------------------------------------
# A6: LowerBound with CompressAI's gradient rule
# ----------------------------------------------------------------------------


class _LowerBoundFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x, bound)
        return torch.max(x, bound)

    @staticmethod
    def backward(ctx, grad_output):
        x, bound = ctx.saved_tensors
        pass_through = (x >= bound) | (grad_output < 0)
        return pass_through.type(grad_output.dtype) * grad_output, None


class LowerBound(nn.Module):
    """max(x, bound); gradient passes where x >= bound or it pushes x up."""

    bound: Tensor

    def __init__(self, bound: float):
        super().__init__()
        self.register_buffer("bound", torch.Tensor([float(bound)]))

    def forward(self, x):
        return _LowerBoundFn.apply(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    """A5: x -> max(x, sqrt(minimum + 2^-36))^2 - 2^-36."""

    pedestal: Tensor

    def __init__(self, minimum: float = 0, reparam_offset: float = 2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.reparam_offset = float(reparam_offset)
        pedestal = self.reparam_offset ** 2
        self.register_buffer("pedestal", torch.Tensor([pedestal]))
        bound = (self.minimum + self.reparam_offset ** 2) ** 0.5
        self.lower_bound = LowerBound(bound)

    def init(self, x: Tensor) -> Tensor:
        return torch.sqrt(torch.max(x + self.pedestal, self.pedestal))

    def forward(self, x: Tensor) -> Tensor:
        out = self.lower_bound(x)
        out = out ** 2 - self.pedestal
        return out


class GDN(nn.Module):
    """A5: y[i] = x[i] / sqrt(beta[i] + sum_j gamma[i, j] * x[j]^2) (or * sqrt)."""

    def __init__(self, in_channels: int, inverse: bool = False,
                 beta_min: float = 1e-6, gamma_init: float = 0.1):
        super().__init__()
        beta_min = float(beta_min)
        gamma_init = float(gamma_init)
        self.inverse = bool(inverse)

        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        beta = torch.ones(in_channels)
        beta = self.beta_reparam.init(beta)
        self.beta = nn.Parameter(beta)

        self.gamma_reparam = NonNegativeParametrizer()
        gamma = gamma_init * torch.eye(in_channels)
        gamma = self.gamma_reparam.init(gamma)
        self.gamma = nn.Parameter(gamma)

    def forward(self, x: Tensor) -> Tensor:
        _, C, _, _ = x.size()
        beta = self.beta_reparam(self.beta)
        gamma = self.gamma_reparam(self.gamma)
        gamma = gamma.reshape(C, C, 1, 1)
        norm = F.conv2d(x ** 2, gamma, beta)
        if self.inverse:
            norm = torch.sqrt(norm)
        else:
            norm = torch.rsqrt(norm)
        return x * norm


# ----------------------------------------------------------------------------
# A3 / A4: conv / deconv helpers
# ----------------------------------------------------------------------------


def conv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size,
                     stride=stride, padding=kernel_size // 2)


def deconv(in_channels, out_channels, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(in_channels, out_channels, kernel_size=kernel_size,
                              stride=stride, output_padding=stride - 1,
                              padding=kernel_size // 2)


# ----------------------------------------------------------------------------
# A7-A12: entropy models
# ----------------------------------------------------------------------------


def pmf_to_quantized_cdf(pmf: Tensor, precision: int = 16) -> Tensor:
    cdf = cdf_rans.pmf_to_quantized_cdf(pmf.tolist(), precision)
    return torch.IntTensor(cdf)


class EntropyModel(nn.Module):
    """A10: quantise / dequantise / CDF tables / per-image rANS calls."""

    def __init__(self, likelihood_bound: float = 1e-9,
                 entropy_coder_precision: int = 16):
        super().__init__()
        self.entropy_coder_precision = int(entropy_coder_precision)
        self.use_likelihood_bound = likelihood_bound > 0
        if self.use_likelihood_bound:
            self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.register_buffer("_offset", torch.IntTensor())
        self.register_buffer("_quantized_cdf", torch.IntTensor())
        self.register_buffer("_cdf_length", torch.IntTensor())

    @property
    def offset(self):
        return self._offset

    @property
    def quantized_cdf(self):
        return self._quantized_cdf

    @property
    def cdf_length(self):
        return self._cdf_length

    def quantize(self, inputs: Tensor, mode: str, means: Optional[Tensor] = None,
                 noise: Optional[Tensor] = None) -> Tensor:
        if mode not in ("noise", "dequantize", "symbols"):
            raise ValueError(f'Invalid quantization mode: "{mode}"')
        if mode == "noise":
            if noise is None:
                half = float(0.5)
                noise = torch.empty_like(inputs).uniform_(-half, half)
            return inputs + noise
        outputs = inputs.clone()
        if means is not None:
            outputs -= means
        outputs = torch.round(outputs)
        if mode == "dequantize":
            if means is not None:
                outputs += means
            return outputs
        assert mode == "symbols", mode
        return outputs.int()

    @staticmethod
    def dequantize(inputs: Tensor, means: Optional[Tensor] = None,
                   dtype: torch.dtype = torch.float) -> Tensor:
        if means is not None:
            outputs = inputs.type_as(means)
            outputs += means
        else:
            outputs = inputs.type(dtype)
        return outputs

    def _pmf_to_cdf(self, pmf, tail_mass, pmf_length, max_length):
        cdf = torch.zeros((len(pmf_length), max_length + 2), dtype=torch.int32)
        for i, p in enumerate(pmf):
            prob = torch.cat((p[: pmf_length[i]], tail_mass[i]), dim=0)
            _cdf = pmf_to_quantized_cdf(prob, self.entropy_coder_precision)
            cdf[i, : _cdf.size(0)] = _cdf
        return cdf

    def _check_cdf_size(self):
        if self._quantized_cdf.numel() == 0:
            raise ValueError("Uninitialized CDFs. Run update() first")
        if len(self._quantized_cdf.size()) != 2:
            raise ValueError(f"Invalid CDF size {self._quantized_cdf.size()}")

    def _check_offsets_size(self):
        if self._offset.numel() == 0:
            raise ValueError("Uninitialized offsets. Run update() first")
        if len(self._offset.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._offset.size()}")

    def _check_cdf_length(self):
        if self._cdf_length.numel() == 0:
            raise ValueError("Uninitialized CDF lengths. Run update() first")
        if len(self._cdf_length.size()) != 1:
            raise ValueError(f"Invalid offsets size {self._cdf_length.size()}")

    def compress(self, inputs, indexes, means=None):
        symbols = self.quantize(inputs, "symbols", means)
        if len(inputs.size()) < 2:
            raise ValueError("Invalid `inputs` size. Expected a tensor with at least 2 dimensions.")
        if inputs.size() != indexes.size():
            raise ValueError("`inputs` and `indexes` should have the same size.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        strings = []
        for i in range(symbols.size(0)):
            rv = cdf_rans.encode_with_indexes(
                symbols[i].reshape(-1).int().numpy(),
                indexes[i].reshape(-1).int().numpy(),
                self._quantized_cdf.numpy(),
                self._cdf_length.reshape(-1).int().numpy(),
                self._offset.reshape(-1).int().numpy(),
            )
            strings.append(rv)
        return strings

    def decompress(self, strings, indexes, dtype=torch.float, means=None):
        if not isinstance(strings, (tuple, list)):
            raise ValueError("Invalid `strings` parameter type.")
        if not len(strings) == indexes.size(0):
            raise ValueError("Invalid strings or indexes parameters")
        if len(indexes.size()) < 2:
            raise ValueError("Invalid `indexes` size. Expected a tensor with at least 2 dimensions.")
        self._check_cdf_size()
        self._check_cdf_length()
        self._check_offsets_size()
        if means is not None:
            if means.size()[:2] != indexes.size()[:2]:
                raise ValueError("Invalid means or indexes parameters")
            if means.size() != indexes.size():
                for i in range(2, len(indexes.size())):
                    if means.size(i) != 1:
                        raise ValueError("Invalid means parameters")
        cdf = self._quantized_cdf
        outputs = cdf.new_empty(indexes.size())
        for i, s in enumerate(strings):
            values = cdf_rans.decode_with_indexes(
                s,
                indexes[i].reshape(-1).int().numpy(),
                cdf.numpy(),
                self._cdf_length.reshape(-1).int().numpy(),
                self._offset.reshape(-1).int().numpy(),
            )
            outputs[i] = torch.from_numpy(values).reshape(outputs[i].size())
        outputs = self.dequantize(outputs, means, dtype)
        return outputs


class EntropyBottleneck(EntropyModel):
    """A7-A11: fully-factorized density model (Balle et al. 2018, appendix 6.1)."""

    _offset: Tensor

    # "plain": sigmoid(upper) - sigmoid(lower)           (CompressAI >= 1.2: forward AND update() both call
    #           the tuple-returning _likelihood)
    # "stable": |sigmoid(s*upper) - sigmoid(s*lower)|, s = -sign(lower + upper)
    #           (CompressAI <= 1.1.x: forward's _likelihood and the inline expression in update())
    # No upstream release mixes the two, so ONE switch drives forward and update(): the tables a release writes
    # are the tables its forward's likelihoods describe.
    likelihood_form = "plain"

    def __init__(self, channels: int, *args, tail_mass: float = 1e-9,
                 init_scale: float = 10, filters: Tuple[int, ...] = (3, 3, 3, 3),
                 **kwargs):
        super().__init__(*args, **kwargs)
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)

        filters = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        channels = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / filters[i + 1]))
            matrix = torch.Tensor(channels, filters[i + 1], filters[i])
            matrix.data.fill_(init)
            self.register_parameter(f"_matrix{i:d}", nn.Parameter(matrix))

            bias = torch.Tensor(channels, filters[i + 1], 1)
            nn.init.uniform_(bias, -0.5, 0.5)
            self.register_parameter(f"_bias{i:d}", nn.Parameter(bias))

            if i < len(self.filters):
                factor = torch.Tensor(channels, filters[i + 1], 1)
                nn.init.zeros_(factor)
                self.register_parameter(f"_factor{i:d}", nn.Parameter(factor))

        self.quantiles = nn.Parameter(torch.Tensor(channels, 1, 3))
        init = torch.Tensor([-self.init_scale, 0, self.init_scale])
        self.quantiles.data = init.repeat(self.quantiles.size(0), 1, 1)

        target = np.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.Tensor([-target, 0, target]))

    def _get_medians(self) -> Tensor:
        medians = self.quantiles[:, :, 1:2]
        return medians

    def update(self, force: bool = False) -> bool:
        # Check if we need to update the bottleneck parameters, the offsets are
        # only computed and stored when the conditonal model is update()'d.
        if self._offset.numel() > 0 and not force:
            return False

        medians = self.quantiles[:, 0, 1]

        minima = medians - self.quantiles[:, 0, 0]
        minima = torch.ceil(minima).int()
        minima = torch.clamp(minima, min=0)

        maxima = self.quantiles[:, 0, 2] - medians
        maxima = torch.ceil(maxima).int()
        maxima = torch.clamp(maxima, min=0)

        self._offset = -minima

        pmf_start = medians - minima
        pmf_length = maxima + minima + 1

        max_length = pmf_length.max().item()
        samples = torch.arange(max_length)
        samples = samples[None, :] + pmf_start[:, None, None]

        half = float(0.5)
        lower = self._logits_cumulative(samples - half, stop_gradient=True)
        upper = self._logits_cumulative(samples + half, stop_gradient=True)
        if self.likelihood_form == "plain":
            pmf = torch.sigmoid(upper) - torch.sigmoid(lower)
        else:
            sign = -torch.sign(lower + upper)
            pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

        pmf = pmf[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])

        quantized_cdf = self._pmf_to_cdf(pmf.detach(), tail_mass.detach(), pmf_length, max_length)
        self._quantized_cdf = quantized_cdf
        self._cdf_length = pmf_length + 2
        return True

    def loss(self) -> Tensor:
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        loss = torch.abs(logits - self.target).sum()
        return loss

    def _logits_cumulative(self, inputs: Tensor, stop_gradient: bool) -> Tensor:
        # TorchScript not yet working (nn.Mmodule indexing not supported)
        logits = inputs
        for i in range(len(self.filters) + 1):
            matrix = getattr(self, f"_matrix{i:d}")
            if stop_gradient:
                matrix = matrix.detach()
            logits = torch.matmul(F.softplus(matrix), logits)

            bias = getattr(self, f"_bias{i:d}")
            if stop_gradient:
                bias = bias.detach()
            logits = logits + bias

            if i < len(self.filters):
                factor = getattr(self, f"_factor{i:d}")
                if stop_gradient:
                    factor = factor.detach()
                logits = logits + torch.tanh(factor) * torch.tanh(logits)
        return logits

    def _likelihood(self, inputs: Tensor) -> Tensor:
        half = float(0.5)
        lower = self._logits_cumulative(inputs - half, stop_gradient=False)
        upper = self._logits_cumulative(inputs + half, stop_gradient=False)
        if self.likelihood_form == "plain":
            return torch.sigmoid(upper) - torch.sigmoid(lower)
        sign = -torch.sign(lower + upper)
        sign = sign.detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def forward(self, x: Tensor, training: Optional[bool] = None,
                noise: Optional[Tensor] = None) -> Tuple[Tensor, Tensor]:
        if training is None:
            training = self.training

        # x from B x C x ... to C x B x ...
        perm = np.arange(len(x.shape))
        perm[0], perm[1] = perm[1], perm[0]
        # Compute inverse permutation
        inv_perm = np.arange(len(x.shape))[np.argsort(perm)]

        x = x.permute(*perm).contiguous()
        shape = x.size()
        values = x.reshape(x.size(0), 1, -1)
        if noise is not None:
            noise = noise.permute(*perm).contiguous().reshape(x.size(0), 1, -1)

        # Add noise or quantize
        outputs = self.quantize(values, "noise" if training else "dequantize",
                                self._get_medians().detach(), noise=noise)

        likelihood = self._likelihood(outputs)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)

        # Convert back to input tensor shape
        outputs = outputs.reshape(shape)
        outputs = outputs.permute(*inv_perm).contiguous()

        likelihood = likelihood.reshape(shape)
        likelihood = likelihood.permute(*inv_perm).contiguous()
        return outputs, likelihood

    @staticmethod
    def _build_indexes(size):
        dims = len(size)
        N = size[0]
        C = size[1]
        view_dims = np.ones((dims,), dtype=np.int64)
        view_dims[1] = -1
        indexes = torch.arange(C).view(*view_dims)
        indexes = indexes.int()
        return indexes.repeat(N, 1, *size[2:])

    @staticmethod
    def _extend_ndims(tensor, n):
        return tensor.reshape(-1, *([1] * n)) if n > 0 else tensor.reshape(-1)

    def compress(self, x):
        indexes = self._build_indexes(x.size())
        medians = self._get_medians().detach()
        spatial_dims = len(x.size()) - 2
        medians = self._extend_ndims(medians, spatial_dims)
        medians = medians.expand(x.size(0), *([-1] * (spatial_dims + 1)))
        return super().compress(x, indexes, medians)

    def decompress(self, strings, size):
        output_size = (len(strings), self._quantized_cdf.size(0), *size)
        indexes = self._build_indexes(output_size).to(self._quantized_cdf.device)
        medians = self._extend_ndims(self._get_medians().detach(), len(size))
        medians = medians.expand(len(strings), *([-1] * (len(size) + 1)))
        return super().decompress(strings, indexes, medians.dtype, medians)


SCALES_MIN = 0.11
SCALES_MAX = 256
SCALES_LEVELS = 64


def get_scale_table(min=SCALES_MIN, max=SCALES_MAX, levels=SCALES_LEVELS):
    return torch.exp(torch.linspace(math.log(min), math.log(max), levels))


class GaussianConditional(EntropyModel):
    """A12: zero-mean (or given-mean) Gaussian with per-element scale."""

    def __init__(self, scale_table, *args, scale_bound: float = 0.11,
                 tail_mass: float = 1e-9, **kwargs):
        super().__init__(*args, **kwargs)
        if not isinstance(scale_table, (type(None), list, tuple)):
            raise ValueError(f'Invalid type for scale_table "{type(scale_table)}"')
        if isinstance(scale_table, (list, tuple)) and len(scale_table) < 1:
            raise ValueError(f'Invalid scale_table length "{len(scale_table)}"')
        if scale_table and (scale_table != sorted(scale_table) or any(s <= 0 for s in scale_table)):
            raise ValueError(f'Invalid scale_table "({scale_table})"')
        self.tail_mass = float(tail_mass)
        if scale_bound is None and scale_table:
            scale_bound = self.scale_table[0]
        if scale_bound <= 0:
            raise ValueError("Invalid parameters")
        self.lower_bound_scale = LowerBound(scale_bound)
        self.register_buffer(
            "scale_table",
            self._prepare_scale_table(scale_table) if scale_table else torch.Tensor(),
        )
        self.register_buffer(
            "scale_bound",
            torch.Tensor([float(scale_bound)]) if scale_bound is not None else None,
        )

    @staticmethod
    def _prepare_scale_table(scale_table):
        return torch.Tensor(tuple(float(s) for s in scale_table))

    @staticmethod
    def _standardized_cumulative(inputs: Tensor) -> Tensor:
        half = float(0.5)
        const = float(-(2 ** -0.5))
        # Using the complementary error function maximizes numerical precision.
        return half * torch.erfc(const * inputs)

    @staticmethod
    def _standardized_quantile(quantile):
        return scipy.stats.norm.ppf(quantile)

    def update_scale_table(self, scale_table, force=False):
        # Check if we need to update the gaussian conditional parameters, the
        # offsets are only computed and stored when the conditonal model is updated.
        if self._offset.numel() > 0 and not force:
            return False
        self.scale_table = self._prepare_scale_table(scale_table)
        self.update()
        return True

    def update(self):
        multiplier = -self._standardized_quantile(self.tail_mass / 2)
        pmf_center = torch.ceil(self.scale_table * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = torch.max(pmf_length).item()

        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None])
        samples_scale = self.scale_table.unsqueeze(1)
        samples = samples.float()
        samples_scale = samples_scale.float()
        upper = self._standardized_cumulative((0.5 - samples) / samples_scale)
        lower = self._standardized_cumulative((-0.5 - samples) / samples_scale)
        pmf = upper - lower

        tail_mass = 2 * lower[:, :1]

        quantized_cdf = self._pmf_to_cdf(pmf, tail_mass, pmf_length, max_length)
        self._quantized_cdf = quantized_cdf
        self._offset = -pmf_center
        self._cdf_length = pmf_length + 2

    def _likelihood(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None) -> Tensor:
        half = float(0.5)
        if means is not None:
            values = inputs - means
        else:
            values = inputs
        scales = self.lower_bound_scale(scales)
        values = torch.abs(values)
        upper = self._standardized_cumulative((half - values) / scales)
        lower = self._standardized_cumulative((-half - values) / scales)
        likelihood = upper - lower
        return likelihood

    def forward(self, inputs: Tensor, scales: Tensor, means: Optional[Tensor] = None,
                training: Optional[bool] = None, noise: Optional[Tensor] = None):
        if training is None:
            training = self.training
        outputs = self.quantize(inputs, "noise" if training else "dequantize", means, noise=noise)
        likelihood = self._likelihood(outputs, scales, means)
        if self.use_likelihood_bound:
            likelihood = self.likelihood_lower_bound(likelihood)
        return outputs, likelihood

    def build_indexes(self, scales: Tensor) -> Tensor:
        scales = self.lower_bound_scale(scales)
        indexes = scales.new_full(scales.size(), len(self.scale_table) - 1).int()
        for s in self.scale_table[:-1]:
            indexes -= (scales <= s).int()
        return indexes


# ----------------------------------------------------------------------------
# A1 / A2: models
# ----------------------------------------------------------------------------


def _update_registered_buffers(module, module_name, buffer_names, state_dict):
    """Resize empty int buffers so load_state_dict accepts updated checkpoints."""
    for name in buffer_names:
        key = f"{module_name}.{name}"
        if key not in state_dict:
            continue
        new_size = state_dict[key].size()
        cur = getattr(module, name)
        if cur is None or cur.size() != new_size:
            dtype = state_dict[key].dtype
            module.register_buffer(name, torch.empty(new_size, dtype=dtype).fill_(0))


class CompressionModel(nn.Module):
    def __init__(self):
        super().__init__()

    def aux_loss(self) -> Tensor:
        loss = sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))
        return loss

    def update(self, scale_table=None, force=False):
        if scale_table is None:
            scale_table = get_scale_table()
        updated = False
        for _, module in self.named_modules():
            if isinstance(module, EntropyBottleneck):
                updated |= module.update(force=force)
            if isinstance(module, GaussianConditional):
                updated |= module.update_scale_table(scale_table, force=force)
        return updated

    def load_state_dict(self, state_dict, strict=True):
        for name, module in self.named_modules():
            if not any(x.startswith(name) for x in state_dict.keys()):
                continue
            if isinstance(module, EntropyBottleneck):
                _update_registered_buffers(module, name, ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
            if isinstance(module, GaussianConditional):
                _update_registered_buffers(
                    module, name, ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return nn.Module.load_state_dict(self, state_dict, strict=strict)


class FactorizedPrior(CompressionModel):
    def __init__(self, N, M, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(M)
        self.g_a = nn.Sequential(
            conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M),
        )
        self.g_s = nn.Sequential(
            deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
            deconv(N, N), GDN(N, inverse=True), deconv(N, 3),
        )
        self.N = N
        self.M = M

    @property
    def downsampling_factor(self) -> int:
        return 2 ** 4

    def forward(self, x, noise=None):
        y = self.g_a(x)
        y_hat, y_likelihoods = self.entropy_bottleneck(y, noise=noise)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods}}

    def compress(self, x):
        y = self.g_a(x)
        y_strings = self.entropy_bottleneck.compress(y)
        return {"strings": [y_strings], "shape": y.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 1
        y_hat = self.entropy_bottleneck.decompress(strings[0], shape)
        x_hat = self.g_s(y_hat).clamp_(0, 1)
        return {"x_hat": x_hat}


class FactorizedPriorReLU(FactorizedPrior):
    def __init__(self, N, M, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        self.g_a = nn.Sequential(
            conv(3, N), nn.ReLU(inplace=True), conv(N, N), nn.ReLU(inplace=True),
            conv(N, N), nn.ReLU(inplace=True), conv(N, M),
        )
        self.g_s = nn.Sequential(
            deconv(M, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
            deconv(N, N), nn.ReLU(inplace=True), deconv(N, 3),
        )


class ScaleHyperprior(CompressionModel):
    def __init__(self, N, M, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = nn.Sequential(
            conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M),
        )
        self.g_s = nn.Sequential(
            deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True),
            deconv(N, N), GDN(N, inverse=True), deconv(N, 3),
        )
        self.h_a = nn.Sequential(
            conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True),
            conv(N, N), nn.ReLU(inplace=True), conv(N, N),
        )
        self.h_s = nn.Sequential(
            deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
            conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True),
        )
        self.gaussian_conditional = GaussianConditional(None)
        self.N = int(N)
        self.M = int(M)

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    def forward(self, x, noise=None):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_hat, z_likelihoods = self.entropy_bottleneck(z, noise=None if noise is None else noise["z"])
        scales_hat = self.h_s(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(
            y, scales_hat, noise=None if noise is None else noise["y"])
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}

    def compress(self, x):
        y = self.g_a(x)
        z = self.h_a(torch.abs(y))
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_strings = self.gaussian_conditional.compress(y, indexes)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_hat = self.gaussian_conditional.decompress(strings[0], indexes, z_hat.dtype)
        x_hat = self.g_s(y_hat).clamp_(0, 1)
        return {"x_hat": x_hat}


# zoo table (upstream compressai/zoo/image.py `cfgs`): quality -> (N, M)
_CFGS = {
    "bmshj2018-factorized": {1: (128, 192), 2: (128, 192), 3: (128, 192), 4: (128, 192),
                             5: (128, 192), 6: (192, 320), 7: (192, 320), 8: (192, 320)},
    "bmshj2018-factorized-relu": {1: (128, 192), 2: (128, 192), 3: (128, 192), 4: (128, 192),
                                  5: (128, 192), 6: (192, 320), 7: (192, 320), 8: (192, 320)},
    "bmshj2018-hyperprior": {1: (128, 192), 2: (128, 192), 3: (128, 192), 4: (128, 192),
                             5: (128, 192), 6: (192, 320), 7: (192, 320), 8: (192, 320)},
}
_ARCH = {
    "bmshj2018-factorized": FactorizedPrior,
    "bmshj2018-factorized-relu": FactorizedPriorReLU,
    "bmshj2018-hyperprior": ScaleHyperprior,
}


def _make(name):
    def ctor(quality, metric="mse", pretrained=False, progress=True, **kwargs):
        if metric not in ("mse", "ms-ssim"):
            raise ValueError(f'Invalid metric "{metric}"')
        if quality < 1 or quality > 8:
            raise ValueError(f'Invalid quality "{quality}", should be between (1, 8)')
        if pretrained:
            raise RuntimeError("pretrained weights need a download; not available offline")
        return _ARCH[name](*_CFGS[name][quality], **kwargs)
    return ctor


image_models = {name: _make(name) for name in _ARCH}


def get_model(model, pretrained, in_channels=3, quality=1):
    """The LICOS surgery, executed on oracle modules (licos/model_utils.py:6-49)."""
    net = image_models[model](quality=quality, pretrained=pretrained)
    if model not in _ARCH:
        raise ValueError("model: " + model + " not supported for raw data.")
    net.entropy_bottleneck = EntropyBottleneck(
        channels=net.entropy_bottleneck.channels, filters=(in_channels, in_channels, 3, 3))
    old = net.g_a[0]
    net.g_a[0] = nn.Conv2d(in_channels, old.out_channels,
                           kernel_size=(old.weight.shape[2], old.weight.shape[3]),
                           stride=old.stride, padding=old.padding)
    old = net.g_s[6]
    net.g_s[6] = nn.ConvTranspose2d(old.in_channels, in_channels,
                                    kernel_size=(old.weight.shape[2], old.weight.shape[3]),
                                    stride=old.stride, padding=old.padding,
                                    output_padding=old.output_padding)
    return net


# ----------------------------------------------------------------------------
# A13: loss and optimizer split
# ----------------------------------------------------------------------------


class RateDistortionLoss(nn.Module):
    def __init__(self, lmbda=0.01, metric="mse", return_type="all"):
        super().__init__()
        if metric != "mse":
            raise NotImplementedError(f"{metric} is not implemented in the oracle")
        self.metric = nn.MSELoss()
        self.lmbda = lmbda
        self.return_type = return_type

    def forward(self, output, target):
        N, _, H, W = target.size()
        out = {}
        num_pixels = N * H * W
        out["bpp_loss"] = sum(
            (torch.log(likelihoods).sum() / (-math.log(2) * num_pixels))
            for likelihoods in output["likelihoods"].values()
        )
        out["mse_loss"] = self.metric(output["x_hat"], target)
        distortion = 255 ** 2 * out["mse_loss"]
        out["loss"] = self.lmbda * distortion + out["bpp_loss"]
        if self.return_type == "all":
            return out
        return out[self.return_type]


def net_aux_optimizer(net: nn.Module, conf: Dict[str, Dict]) -> Dict[str, torch.optim.Optimizer]:
    parameters = {
        "net": {n for n, p in net.named_parameters() if p.requires_grad and not n.endswith(".quantiles")},
        "aux": {n for n, p in net.named_parameters() if p.requires_grad and n.endswith(".quantiles")},
    }
    params_dict = dict(net.named_parameters())
    inter = parameters["net"] & parameters["aux"]
    union = parameters["net"] | parameters["aux"]
    assert len(inter) == 0
    assert len(union) - len(params_dict.keys()) == 0

    def make(key):
        kwargs = dict(conf[key])
        typ = kwargs.pop("type")
        params = (params_dict[name] for name in sorted(parameters[key]))
        return getattr(torch.optim, typ)(params, **kwargs)

    return {key: make(key) for key in ["net", "aux"]}


def federated_average(local_sd, central_sd, loss, best_loss):
    """licos/federation_utils.py:47-53: loss-weighted two-way average of every key."""
    w_local = best_loss / (best_loss + loss)
    w_central = loss / (best_loss + loss)
    out = {}
    for key in local_sd:
        v = w_local * local_sd[key]
        v += w_central * central_sd[key]
        out[key] = v
    return out
