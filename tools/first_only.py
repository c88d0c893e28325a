"""Runs only the fused first layer (debug aid)."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licos_b200 import _lib, ops
dev = torch.device("cuda:0")
B, cin, cout, H, W = 2, int(os.environ.get("CIN", "3")), 128, 64, 64
x = torch.randn(B, cin, H, W).to(dev)
w = torch.randn(cout, cin, 5, 5) / (cin * 25) ** 0.5
packed = ops.pack_conv_weight(w.to(dev), _lib.CONV_5X5_S2, cout, cin, _lib.LAYOUT_NCHW_F32)
epi = int(os.environ.get("EPI", "1"))
bh = gh = None
if epi in (1, 2):
    bh, gh = ops.gdn_pack(torch.ones(cout, device=dev), (0.1 * torch.eye(cout)).sqrt().to(dev), 0.0, 0.0, 0.0)
torch.cuda.synchronize()
out = ops.conv_forward(x, kind=_lib.CONV_5X5_S2, epilogue=epi, in_layout=_lib.LAYOUT_NCHW_F32,
                       out_layout=_lib.LAYOUT_NHWC_BF16, in_c=cin, out_c=cout, weight=packed,
                       bias=torch.zeros(cout, device=dev), beta=bh, gamma=gh)
torch.cuda.synchronize()
print("ok", out.float().abs().mean().item())
