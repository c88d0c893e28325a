"""Times licos_gdn_backward alone (default: the training step's largest layer, 32 x 128 x 128 pixels, 128 channels;
PIXELS=... for the others).  LICOS_GDN_BWD_ONE_TEAM=1 selects the round-1 one-team kernel for an A/B."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licos_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
P, C = int(os.environ.get("PIXELS", 32 * 128 * 128)), 128
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn(P, C, device=dev, generator=g).bfloat16()
gr = torch.randn(P, C, device=dev, generator=g).bfloat16()
gamma = (torch.rand(C, C, device=dev, generator=g) * 0.02 + 0.1 * torch.eye(C, device=dev)).bfloat16()
beta = torch.rand(C, device=dev, generator=g) + 0.5
for inverse in (False, True):
    dg, db, dbias = torch.zeros(C, C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    for _ in range(3):
        ops.gdn_backward(x, gr, gamma, beta, inverse, dg, db, dbias)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        ops.gdn_backward(x, gr, gamma, beta, inverse, dg, db, dbias)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"inverse={inverse}: {ms * 1e3:.1f} us per launch, {3 * P * C * 2 / ms / 1e6:.0f} GB/s of x + g + dx")
