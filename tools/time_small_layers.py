"""Times the small-resolution engine layers of a 32-tile training step (forward and data-gradient launches)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licos_b200 import _lib as L, ops  # noqa: E402

dev = torch.device("cuda", 0)
B = 32
cases = [("conv 128->128 in 128 (g_a[2])", L.CONV_5X5_S2, 128, 128, 128, L.LAYOUT_NHWC_BF16),
         ("conv 128->128 in 64 (g_a[4])", L.CONV_5X5_S2, 128, 128, 64, L.LAYOUT_NHWC_BF16),
         ("conv 128->192 in 32 (g_a[6], NCHW out)", L.CONV_5X5_S2, 128, 192, 32, L.LAYOUT_NCHW_F32),
         ("deconv 192->128 in 16 (g_s[0])", L.DECONV_5X5_S2, 192, 128, 16, L.LAYOUT_NHWC_BF16),
         ("deconv 128->128 in 32 (g_s[2])", L.DECONV_5X5_S2, 128, 128, 32, L.LAYOUT_NHWC_BF16),
         ("deconv 128->128 in 64 (g_s[4])", L.DECONV_5X5_S2, 128, 128, 64, L.LAYOUT_NHWC_BF16)]
tot = 0.0
for name, kind, ci, co, hw, ol in cases:
    w = torch.randn((co, ci, 5, 5) if kind == L.CONV_5X5_S2 else (ci, co, 5, 5), device=dev) * 0.05
    pk = ops.pack_conv_weight(w, kind, co, ci, L.LAYOUT_NHWC_BF16)
    x = torch.randn(B, hw, hw, ci, device=dev).to(torch.bfloat16)
    f = lambda: ops.conv_forward(x, kind=kind, epilogue=L.EPI_NONE, in_layout=L.LAYOUT_NHWC_BF16, out_layout=ol, in_c=ci,  # noqa: E731
                                 out_c=co, weight=pk, bias=None)
    for _ in range(3):
        f()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 * 1e3
    tot += us
    print(f"{name:42s} {us:7.1f} us")
print(f"sum {tot:.1f} us (LICOS_NO_SMALL_TILES={os.environ.get('LICOS_NO_SMALL_TILES', '')}, "
      f"LICOS_FORCE_NACC1={os.environ.get('LICOS_FORCE_NACC1', '')})")
