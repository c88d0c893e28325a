"""compressai.losses.RateDistortionLoss (/root/reference/licos/train.py:123,192) and the
bpp / PSNR metrics of /root/reference/eval_utils.py:145-186, with the reductions done by the C-ABI kernels
when autograd is off."""
from __future__ import annotations

import math

import torch
import torch.nn as nn

from . import ops
from . import torch_ops as T


class _RDTermsFn(torch.autograd.Function):
    """(bpp_loss, mse_loss) with the reductions and their backward on the device kernels (train.py:192-193)."""

    @staticmethod
    def forward(ctx, x_hat, target, num_pixels, *liks):
        acc = torch.zeros(2, dtype=torch.float64, device=target.device)
        liks = [lk.contiguous() for lk in liks]
        for lk in liks:
            T.sum_log(lk, acc[0:1])
        x_hat, target = x_hat.contiguous(), target.contiguous()
        T.sum_sq_err(x_hat, target, acc[1:2])
        ctx.save_for_backward(x_hat, target, *liks)
        ctx.num_pixels = num_pixels
        return (acc[0] / (-math.log(2) * num_pixels)).float(), (acc[1] / target.numel()).float()

    @staticmethod
    def backward(ctx, g_bpp, g_mse):
        x_hat, target, *liks = ctx.saved_tensors
        g_bpp = g_bpp.reshape(1).float().contiguous()
        g_mse = g_mse.reshape(1).float().contiguous()
        d_xhat = ops.scaled_diff(x_hat, target, 2.0 / target.numel(), g_mse) if ctx.needs_input_grad[0] else None
        d_liks = [ops.scaled_reciprocal(lk, 1.0 / (-math.log(2) * ctx.num_pixels), g_bpp) if need else None
                  for lk, need in zip(liks, ctx.needs_input_grad[3:])]
        return (d_xhat, None, None, *d_liks)


class RateDistortionLoss(nn.Module):
    def __init__(self, lmbda: float = 0.01, metric: str = "mse", return_type: str = "all"):
        super().__init__()
        if metric != "mse":
            raise NotImplementedError(f"{metric} is not implemented (LICOS trains with mse, train.py:123)")
        self.lmbda = lmbda
        self.return_type = return_type

    def forward(self, output, target):
        N, _, H, W = target.size()
        num_pixels = N * H * W
        needs_grad = torch.is_grad_enabled() and (
            output["x_hat"].requires_grad or any(v.requires_grad for v in output["likelihoods"].values()))
        out = {}
        if needs_grad:
            out["bpp_loss"], out["mse_loss"] = _RDTermsFn.apply(output["x_hat"], target, num_pixels,
                                                                *output["likelihoods"].values())
        else:
            acc = torch.zeros(2, dtype=torch.float64, device=target.device)
            for lk in output["likelihoods"].values():
                T.sum_log(lk.contiguous(), acc[0:1])
            T.sum_sq_err(output["x_hat"].contiguous(), target.contiguous(), acc[1:2])
            out["bpp_loss"] = (acc[0] / (-math.log(2) * num_pixels)).float()
            out["mse_loss"] = (acc[1] / target.numel()).float()
        distortion = 255 ** 2 * out["mse_loss"]
        out["loss"] = self.lmbda * distortion + out["bpp_loss"]
        return out if self.return_type == "all" else out[self.return_type]


@torch.no_grad()
def compute_bpp(out_net) -> float:
    """eval_utils.py:172-186"""
    size = out_net["x_hat"].size()
    num_pixels = size[0] * size[2] * size[3]
    acc = torch.zeros(1, dtype=torch.float64, device=out_net["x_hat"].device)
    for lk in out_net["likelihoods"].values():
        T.sum_log(lk.contiguous(), acc)
    return float(acc.item() / (-math.log(2) * num_pixels))


@torch.no_grad()
def compute_psnr(a, b) -> float:
    """eval_utils.py:145-156"""
    acc = T.sum_sq_err(a.contiguous(), b.contiguous())
    return -10 * math.log10(acc.item() / a.numel())


def compute_msssim(a, b) -> float:
    """/root/reference/eval_utils.py:159-169 (pytorch_msssim.ms_ssim(a, b, data_range=1.0)) on the device."""
    from . import ops

    return float(ops.ms_ssim(a.float(), b.float(), data_range=1.0).item())
