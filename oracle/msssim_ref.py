"""CPU oracle for the MS-SSIM metric (TEST INFRASTRUCTURE -- only tests/ may import this).

Restates pytorch_msssim.ms_ssim (the third-party package /root/reference/eval_utils.py:5,159-169 calls; un-pinned in
environment.yml, not installed here -> parity unpinned, like the rest of the oracle): 11-tap Gaussian window
(sigma 1.5), separable "valid" filtering per channel, K = (0.01, 0.03), five scales with weights
(0.0448, 0.2856, 0.3001, 0.2363, 0.1333), contrast-structure terms of the first four scales and SSIM of the last,
each clamped at 0 (relu), 2x2 average pooling with padding = size % 2 between scales, mean over (batch, channel).
"""
import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def gauss_window(size: int = 11, sigma: float = 1.5) -> torch.Tensor:
    coords = torch.arange(size, dtype=torch.float32) - size // 2
    g = torch.exp(-(coords ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter(x: torch.Tensor, win: torch.Tensor) -> torch.Tensor:
    C = x.shape[1]
    k = win.numel()
    out = F.conv2d(x, win.view(1, 1, k, 1).repeat(C, 1, 1, 1), groups=C)
    return F.conv2d(out, win.view(1, 1, 1, k).repeat(C, 1, 1, 1), groups=C)


def _ssim(x, y, data_range, win, K=(0.01, 0.03)):
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter(x, win), _filter(y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _filter(x * x, win) - mu1_sq
    s2 = _filter(y * y, win) - mu2_sq
    s12 = _filter(x * y, win) - mu12
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return torch.flatten(ssim_map, 2).mean(-1), torch.flatten(cs_map, 2).mean(-1)


def ms_ssim(x: torch.Tensor, y: torch.Tensor, data_range: float = 1.0) -> torch.Tensor:
    if x.shape != y.shape or x.dim() != 4:
        raise ValueError("ms_ssim expects two (B, C, H, W) tensors of the same shape")
    if min(x.shape[-2:]) <= (11 - 1) * 2 ** 4:
        raise AssertionError("Image size should be larger than 160 due to the 4 downsamplings in ms-ssim")
    win = gauss_window()
    weights = torch.tensor(WEIGHTS, dtype=x.dtype)
    mcs = []
    for level in range(5):
        ssim_pc, cs = _ssim(x, y, data_range, win)
        if level < 4:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in x.shape[2:]]
            x = F.avg_pool2d(x, kernel_size=2, padding=pad)
            y = F.avg_pool2d(y, kernel_size=2, padding=pad)
    stack = torch.stack(mcs + [torch.relu(ssim_pc)], dim=0)
    val = torch.prod(stack ** weights.view(-1, 1, 1), dim=0)
    return val.mean()
