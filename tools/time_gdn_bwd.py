"""Times licos_gdn_backward on the three GDN layer sizes of a cfg-5 training step (32 tiles): CUDA events, 20 launches."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licos_b200 import ops  # noqa: E402

dev = torch.device("cuda", 0)
C = 128
gamma = (torch.rand(C, C, device=dev) * 0.02 + 0.1 * torch.eye(C, device=dev)).to(torch.bfloat16)
beta = torch.rand(C, device=dev) + 0.5
tot = 0.0
for hw in (128, 64, 32):
    P = 32 * hw * hw
    x = torch.randn(P, C, device=dev).to(torch.bfloat16)
    g = torch.randn(P, C, device=dev).to(torch.bfloat16)
    dg, db, dbias = torch.zeros(C, C, device=dev), torch.zeros(C, device=dev), torch.zeros(C, device=dev)
    for inverse in (False, True):
        for _ in range(3):
            ops.gdn_backward(x, g, gamma, beta, inverse, dg, db, dbias)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.gdn_backward(x, g, gamma, beta, inverse, dg, db, dbias)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / 20 * 1e3
        tot += us
        print(f"{'IGDN' if inverse else 'GDN '} backward 32 x {hw}x{hw} x 128: {us:7.1f} us  {3 * P * C * 2 / us / 1e6:6.2f} TB/s (x, g in; dx out)")
print(f"sum {tot:.1f} us")
