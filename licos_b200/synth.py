"""Synthetic inputs and deterministic weight conditioning for parity tests and benchmarks.

Random-init weights give all-zero symbols, which would make the bit-exact integer checks vacuous, so after the
seeded default init the last analysis conv is scaled up, medians / quantiles / density parameters are spread,
and (hyperprior) the scale head is scaled so both clamps of the scale table are hit.  No forward pass is
needed: the constants are fixed, so the same ``state_dict`` can be rebuilt anywhere from the seed.
"""
from __future__ import annotations

import torch
import torch.nn as nn

LATENT_GAIN = {1: 120.0, 3: 70.0, 13: 35.0}


def make_input(kind: str, batch: int, seed: int = 42, device="cpu") -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    if kind == "rgb256":          # cfg 1 / 2 / 5: cfg/default_cfg.toml patch_size 256x256, seed 42
        x = torch.rand(batch, 3, 256, 256, generator=g)
    elif kind == "raw512":        # cfg 3: 12-bit single band, raw_image_folder.py:192-196 scaling
        x = torch.randint(0, 4096, (batch, 1, 512, 512), generator=g).float() / 4095
    elif kind == "merged256":     # reference "merged" raw format: 13 bands
        x = torch.rand(batch, 13, 256, 256, generator=g)
    elif kind == "rgb1024":       # cfg 4
        x = torch.rand(batch, 3, 1024, 1024, generator=g)
    else:
        raise ValueError(kind)
    return x.to(device)


@torch.no_grad()
def condition_weights(net: nn.Module, seed: int = 1234) -> nn.Module:
    g = torch.Generator().manual_seed(seed)
    in_ch = net.g_a[0].in_channels
    gain = LATENT_GAIN.get(in_ch, 70.0)
    net.g_a[6].weight.mul_(gain)
    net.g_a[6].bias.mul_(gain)

    eb = net.entropy_bottleneck
    C = eb.channels
    med = torch.rand(C, generator=g) - 0.5
    lo = med - (5 + 25 * torch.rand(C, generator=g))
    hi = med + (5 + 25 * torch.rand(C, generator=g))
    eb.quantiles.copy_(torch.stack([lo, med, hi], dim=1).reshape(C, 1, 3))
    for name, p in eb.named_parameters():
        if name.startswith("_factor"):
            p.copy_(0.5 * torch.randn(p.shape, generator=g))
        elif name.startswith("_matrix"):
            p.add_(0.3 * torch.randn(p.shape, generator=g))

    if hasattr(net, "h_s"):
        # z carries real magnitude, and sigma spans both clamps of the 0.11 .. 256 table without parking a
        # fifth of the latents on the 1e-9 likelihood floor (which would make bpp hinge on a handful of ties)
        net.h_a[4].weight.mul_(12.0)
        last = net.h_s[4]
        last.weight.mul_(10.0)
        M = last.out_channels
        last.bias.copy_(torch.exp(torch.empty(M).uniform_(-2.3, 5.75, generator=g)))
    return net
