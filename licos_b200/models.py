"""Host-side mirrors of compressai.models.{CompressionModel, FactorizedPrior, FactorizedPriorReLU,
ScaleHyperprior} and of compressai.zoo.image_models (section 8b).

The model object keeps CompressAI's surface -- ``net(x) -> {"x_hat", "likelihoods"}``, ``compress`` /
``decompress``, ``aux_loss``, ``update``, ``state_dict`` key names, indexable ``g_a`` / ``g_s`` -- so it drops
into /root/reference/licos/model_utils.py:19-45, train.py:190-200 and eval_utils.py:199-201 unchanged.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn as nn
from torch import Tensor

from .entropy_models import EntropyBottleneck, GaussianConditional, get_scale_table
from .layers import GDN, FusedSequential, conv, deconv

__all__ = ["CompressionModel", "FactorizedPrior", "FactorizedPriorReLU", "ScaleHyperprior", "image_models",
           "model_architectures"]


def _resize_buffers(module: nn.Module, prefix: str, names, state_dict) -> None:
    """Checkpoints written after update() carry non-empty integer tables; give them room before loading."""
    for name in names:
        key = f"{prefix}.{name}"
        if key not in state_dict:
            continue
        new = state_dict[key]
        cur = getattr(module, name, None)
        if cur is None or cur.size() != new.size():
            device = cur.device if cur is not None else new.device
            module.register_buffer(name, torch.zeros(new.size(), dtype=new.dtype, device=device))


class CompressionModel(nn.Module):
    def aux_loss(self) -> Tensor:
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def update(self, scale_table=None, force: bool = False) -> bool:
        if scale_table is None:
            scale_table = get_scale_table()
        updated = False
        for module in self.modules():
            if isinstance(module, EntropyBottleneck):
                updated |= module.update(force=force)
            if isinstance(module, GaussianConditional):
                updated |= module.update_scale_table(scale_table, force=force)
        return updated

    def load_state_dict(self, state_dict, strict: bool = True, **kwargs):
        for name, module in self.named_modules():
            if not any(k.startswith(name) for k in state_dict.keys()):
                continue
            if isinstance(module, EntropyBottleneck):
                _resize_buffers(module, name, ["_quantized_cdf", "_offset", "_cdf_length"], state_dict)
            if isinstance(module, GaussianConditional):
                _resize_buffers(module, name, ["_quantized_cdf", "_offset", "_cdf_length", "scale_table"], state_dict)
        return super().load_state_dict(state_dict, strict=strict, **kwargs)


def _analysis(cin: int, N: int, M: int, relu: bool = False) -> FusedSequential:
    act = (lambda: nn.ReLU(inplace=True)) if relu else (lambda: GDN(N))
    return FusedSequential(conv(cin, N), act(), conv(N, N), act(), conv(N, N), act(), conv(N, M))


def _synthesis(cout: int, N: int, M: int, relu: bool = False) -> FusedSequential:
    act = (lambda: nn.ReLU(inplace=True)) if relu else (lambda: GDN(N, inverse=True))
    return FusedSequential(deconv(M, N), act(), deconv(N, N), act(), deconv(N, N), act(), deconv(N, cout))


class FactorizedPrior(CompressionModel):
    def __init__(self, N: int, M: int, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(M)
        self.g_a = _analysis(3, N, M)
        self.g_s = _synthesis(3, N, M)
        self.N, self.M = N, M

    @property
    def downsampling_factor(self) -> int:
        return 2 ** 4

    def forward(self, x: Tensor) -> Dict:
        y = self.g_a(x)
        eb = self.entropy_bottleneck
        if not eb.training and not torch.is_grad_enabled() and y.is_cuda:
            # inference: quantise + likelihoods + the bf16 NHWC copy g_s reads, in one pass over y
            y_hat, y_likelihoods, _, y_nhwc = eb.forward_fused(y, want_symbols=False, want_nhwc=True)
            x_hat = self.g_s(y_hat, nhwc=y_nhwc)
        else:
            y_hat, y_likelihoods = eb(y)
            x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods}}

    @torch.no_grad()
    def forward_tiles(self, x: Tensor, *, int_max: int = 0, requant8: bool = False, out_dtype=None, out_max: int = 0,
                      sum_ln=None, want_float: bool = True) -> Dict:
        """The codec loop over a batch of tiles in one call (inference): encode = g_a + quantise + likelihoods + symbols,
        decode = g_s.  ``x`` is fp32 in [0, 1] or INTEGER tiles (uint8; 12-bit DNs in uint16 / int16 storage, scaled by
        1 / ``int_max`` inside the first layer -- what raw_image_folder.py:192-196 does on the host -- ``requant8`` = its
        8-bit step).  Returns ``x_hat`` (fp32, or ``round(clamp(x_hat, 0, 1) * out_max)`` as ``out_dtype`` uint8 / uint16),
        ``likelihoods``, ``y_hat`` and the int16 ``symbols`` the range coder consumes; ``sum_ln`` (float64 device scalar)
        accumulates sum(ln(likelihoods)) so that bpp needs no second pass (eval_utils.py:172-186)."""
        y = self.g_a(x, int_max=int_max, requant8=requant8)
        y_hat, lik, _, y_nhwc, sym16 = self.entropy_bottleneck.forward_fused(
            y, want_symbols=False, want_nhwc=True, want_symbols_i16=True, sum_ln=sum_ln, want_float=want_float)
        if out_dtype is not None and out_max == 0:
            out_max = int_max or (255 if out_dtype == torch.uint8 else 4095)
        x_hat = self.g_s(y_hat if y_hat is not None else y, nhwc=y_nhwc, out_dtype=out_dtype, out_max=out_max)
        return {"x_hat": x_hat, "likelihoods": {"y": lik}, "y_hat": y_hat, "symbols": sym16}

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls(state_dict["g_a.0.weight"].size(0), state_dict["g_a.6.weight"].size(0))
        net.load_state_dict(state_dict)
        return net

    @torch.no_grad()
    def compress(self, x: Tensor) -> Dict:
        y = self.g_a(x)
        return {"strings": [self.entropy_bottleneck.compress(y)], "shape": y.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape) -> Dict:
        assert isinstance(strings, list) and len(strings) == 1
        y_hat = self.entropy_bottleneck.decompress(strings[0], shape)
        return {"x_hat": self.g_s(y_hat).clamp_(0, 1)}


class FactorizedPriorReLU(FactorizedPrior):
    def __init__(self, N: int, M: int, **kwargs):
        super().__init__(N=N, M=M, **kwargs)
        self.g_a = _analysis(3, N, M, relu=True)
        self.g_s = _synthesis(3, N, M, relu=True)


class ScaleHyperprior(CompressionModel):
    def __init__(self, N: int, M: int, **kwargs):
        super().__init__(**kwargs)
        self.entropy_bottleneck = EntropyBottleneck(N)
        self.g_a = _analysis(3, N, M)
        self.g_s = _synthesis(3, N, M)
        self.h_a = FusedSequential(conv(M, N, stride=1, kernel_size=3), nn.ReLU(inplace=True),
                                   conv(N, N), nn.ReLU(inplace=True), conv(N, N))
        self.h_s = FusedSequential(deconv(N, N), nn.ReLU(inplace=True), deconv(N, N), nn.ReLU(inplace=True),
                                   conv(N, M, stride=1, kernel_size=3), nn.ReLU(inplace=True))
        self.gaussian_conditional = GaussianConditional(None)
        self.N, self.M = int(N), int(M)

    @property
    def downsampling_factor(self) -> int:
        return 2 ** (4 + 2)

    def forward(self, x: Tensor) -> Dict:
        y = self.g_a(x)
        z = self.h_a(y, take_abs=True)
        z_hat, z_likelihoods = self.entropy_bottleneck(z)
        scales_hat = self.h_s(z_hat)
        y_hat, y_likelihoods = self.gaussian_conditional(y, scales_hat)
        x_hat = self.g_s(y_hat)
        return {"x_hat": x_hat, "likelihoods": {"y": y_likelihoods, "z": z_likelihoods}}

    @classmethod
    def from_state_dict(cls, state_dict):
        net = cls(state_dict["g_a.0.weight"].size(0), state_dict["g_a.6.weight"].size(0))
        net.load_state_dict(state_dict)
        return net

    @torch.no_grad()
    def compress(self, x: Tensor) -> Dict:
        y = self.g_a(x)
        z = self.h_a(y, take_abs=True)
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_strings = self.gaussian_conditional.compress(y, indexes)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape) -> Dict:
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        scales_hat = self.h_s(z_hat)
        indexes = self.gaussian_conditional.build_indexes(scales_hat)
        y_hat = self.gaussian_conditional.decompress(strings[0], indexes, z_hat.dtype)
        return {"x_hat": self.g_s(y_hat).clamp_(0, 1)}


# ---------------------------------------------------------------------------------------------
# zoo: quality -> (N, M), as compressai/zoo/image.py `cfgs`
# ---------------------------------------------------------------------------------------------
_NM = {q: ((128, 192) if q <= 5 else (192, 320)) for q in range(1, 9)}
model_architectures = {
    "bmshj2018-factorized": FactorizedPrior,
    "bmshj2018-factorized-relu": FactorizedPriorReLU,
    "bmshj2018-hyperprior": ScaleHyperprior,
}


def _zoo_entry(name: str):
    def build(quality, metric="mse", pretrained=False, progress=True, **kwargs):
        if metric not in ("mse", "ms-ssim"):
            raise ValueError(f'Invalid metric "{metric}"')
        if quality < 1 or quality > 8:
            raise ValueError(f'Invalid quality "{quality}", should be between (1, 8)')
        if pretrained:
            raise RuntimeError("pretrained CompressAI weights need a download; load a state_dict instead")
        return model_architectures[name](*_NM[quality], **kwargs)

    build.__name__ = name.replace("-", "_")
    return build


image_models = {name: _zoo_entry(name) for name in model_architectures}
