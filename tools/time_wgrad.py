"""Times licos_conv_wgrad on the layer shapes of a cfg-5 training step (32 tiles): CUDA events, 20 launches each."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licos_b200 import _lib, ops  # noqa: E402

dev = torch.device("cuda", 0)
B = 32
shapes = [("conv 128->128 @64 (g_a[2])", _lib.CONV_5X5_S2, (64, 64, 128), (128, 128, 128)),
          ("conv 128->128 @32 (g_a[4])", _lib.CONV_5X5_S2, (32, 32, 128), (64, 64, 128)),
          ("conv 128->192 @16 (g_a[6])", _lib.CONV_5X5_S2, (16, 16, 192), (32, 32, 128)),
          ("deconv 192->128 @16 (g_s[0])", _lib.DECONV_5X5_S2, (16, 16, 192), (32, 32, 128)),
          ("first/last layer 1x1 K=128 @128", _lib.CONV_1X1, (128, 128, 128), (128, 128, 128))]
tot = 0.0
for name, kind, s, b in shapes:
    small = torch.randn(B, *s, device=dev).to(torch.bfloat16)
    big = torch.randn(B, *b, device=dev).to(torch.bfloat16)
    taps = 1 if kind == _lib.CONV_1X1 else 25
    out = torch.zeros(taps * s[2] * b[2], device=dev)
    for _ in range(3):
        ops.conv_wgrad(small, big, kind, out=out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.conv_wgrad(small, big, kind, out=out)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    flop = 2.0 * taps * s[2] * b[2] * B * s[0] * s[1]
    tot += us
    print(f"{name:36s} {us:8.1f} us  {flop / us / 1e6:7.1f} TFLOP/s")
print(f"sum {tot:.1f} us  (LICOS_WGRAD_UNITS_PER_SM={os.environ.get('LICOS_WGRAD_UNITS_PER_SM', '1')}, "
      f"LICOS_WGRAD_MIN_TILES={os.environ.get('LICOS_WGRAD_MIN_TILES', '8')})")
