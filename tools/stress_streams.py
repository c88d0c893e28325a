"""Concurrency stress: the same layer launched from 3 streams at once (what bench.py's host-buffer leg does).
Usage: python tools/stress_streams.py <layer index | all> [batch] [iters]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licos_b200 import _lib, ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 32
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 50
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
LAYERS = [
    ("g_a0", _lib.CONV_5X5_S2, 3, 128, 256, 256, _lib.EPI_GDN, True, False),
    ("g_a2", _lib.CONV_5X5_S2, 128, 128, 128, 128, _lib.EPI_GDN, False, False),
    ("g_a4", _lib.CONV_5X5_S2, 128, 128, 64, 64, _lib.EPI_GDN, False, False),
    ("g_a6", _lib.CONV_5X5_S2, 128, 192, 32, 32, _lib.EPI_NONE, False, True),
    ("g_s0", _lib.DECONV_5X5_S2, 192, 128, 16, 16, _lib.EPI_IGDN, False, False),
    ("g_s2", _lib.DECONV_5X5_S2, 128, 128, 32, 32, _lib.EPI_IGDN, False, False),
    ("g_s4", _lib.DECONV_5X5_S2, 128, 128, 64, 64, _lib.EPI_IGDN, False, False),
    ("g_s6", _lib.DECONV_5X5_S2, 128, 3, 128, 128, _lib.EPI_NONE, False, True),
]
sel = range(len(LAYERS)) if which == "all" else [int(which)]
jobs = []
for li in sel:
    name, kind, cin, cout, H, W, epi, first, nchw = LAYERS[li]
    x = torch.randn(B, cin, H, W, generator=g)
    wshape = (cin, cout, 5, 5) if kind == _lib.DECONV_5X5_S2 else (cout, cin, 5, 5)
    w = torch.randn(wshape, generator=g) / (cin * 25) ** 0.5
    in_layout = _lib.LAYOUT_NCHW_F32 if first else _lib.LAYOUT_NHWC_BF16
    out_layout = _lib.LAYOUT_NCHW_F32 if nchw else _lib.LAYOUT_NHWC_BF16
    xd = x.to(dev) if first else x.permute(0, 2, 3, 1).contiguous().bfloat16().to(dev)
    packed = ops.pack_conv_weight(w.to(dev), kind, cout, cin, in_layout)
    bh = gh = None
    if epi in (_lib.EPI_GDN, _lib.EPI_IGDN):
        bh, gh = ops.gdn_pack(torch.ones(cout, device=dev), (0.1 * torch.eye(cout)).sqrt().to(dev), 0.0, 0.0, 0.0)
    kw = dict(kind=kind, epilogue=epi, in_layout=in_layout, out_layout=out_layout, in_c=cin, out_c=cout,
              weight=packed, bias=torch.zeros(cout, device=dev), beta=bh, gamma=gh)
    ref = ops.conv_forward(xd, **kw).float().clone()
    jobs.append((name, xd, kw, ref))
torch.cuda.synchronize()
streams = [torch.cuda.Stream(device=dev) for _ in range(3)]
bad = 0
for it in range(iters):
    outs = []
    for si, s in enumerate(streams):
        with torch.cuda.stream(s):
            for name, xd, kw, ref in jobs:
                outs.append((name, ops.conv_forward(xd, **kw), ref))
    torch.cuda.synchronize()
    if it % 10 == 0:
        for name, o, ref in outs:
            if not torch.equal(o.float(), ref):
                bad += 1
                print("MISMATCH", name, (o.float() - ref).abs().max().item())
print("stress done", which, "B", B, "bad", bad)
