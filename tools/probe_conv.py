"""Per-layer timing of the conv engine at benchmark shapes plus the in-kernel cycle probes
(licos_internal_set_conv_probe, a development symbol outside the public header).  Usage: python tools/probe_conv.py [batch]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from licos_b200 import _lib, ops  # noqa: E402

import ctypes  # noqa: E402

_probe_fn = _lib.lib.licos_internal_set_conv_probe
_probe_fn.restype, _probe_fn.argtypes = None, [ctypes.c_void_p]


def _set_probe(ptr):
    _probe_fn(ptr)


NAMES = ["pA_wait", "pB_wait", "mma_waitA", "mma_waitB", "mma_waitAcc", "mma_waitX2", "mma_total", "epi_waitAcc",
         "epi_s1", "epi_waitNorm", "epi_s2", "epi_store", "epi_total", "tiles", "pA_total", "pB_total"]

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
LAYERS = [  # kind, cin, cout, H, W (input), epilogue, first, out_nchw
    ("g_a0 3->128", _lib.CONV_5X5_S2, 3, 128, 256, 256, _lib.EPI_GDN, True, False),
    ("g_a2 128->128", _lib.CONV_5X5_S2, 128, 128, 128, 128, _lib.EPI_GDN, False, False),
    ("g_a4 128->128", _lib.CONV_5X5_S2, 128, 128, 64, 64, _lib.EPI_GDN, False, False),
    ("g_a6 128->192", _lib.CONV_5X5_S2, 128, 192, 32, 32, _lib.EPI_NONE, False, True),
    ("g_s0 192->128", _lib.DECONV_5X5_S2, 192, 128, 16, 16, _lib.EPI_IGDN, False, False),
    ("g_s2 128->128", _lib.DECONV_5X5_S2, 128, 128, 32, 32, _lib.EPI_IGDN, False, False),
    ("g_s4 128->128", _lib.DECONV_5X5_S2, 128, 128, 64, 64, _lib.EPI_IGDN, False, False),
    ("g_s6 128->3", _lib.DECONV_5X5_S2, 128, 3, 128, 128, _lib.EPI_NONE, False, True),
    ("plain 128->128 (no GDN)", _lib.CONV_5X5_S2, 128, 128, 128, 128, _lib.EPI_NONE, False, False),
]
probe = torch.zeros(16 * 160, dtype=torch.int64, device=dev)
SEL = os.environ.get("LAYERS")
if SEL:
    LAYERS = [LAYERS[int(i)] for i in SEL.split(",")]
for name, kind, cin, cout, H, W, epi, first, nchw in LAYERS:
    kk = 5
    x = torch.randn(B, cin, H, W, generator=g)
    wshape = (cin, cout, kk, kk) if kind == _lib.DECONV_5X5_S2 else (cout, cin, kk, kk)
    w = torch.randn(wshape, generator=g) / (cin * 25) ** 0.5
    bias = torch.zeros(cout)
    in_layout = _lib.LAYOUT_NCHW_F32 if first else _lib.LAYOUT_NHWC_BF16
    out_layout = _lib.LAYOUT_NCHW_F32 if nchw else _lib.LAYOUT_NHWC_BF16
    xd = x.to(dev) if first else x.permute(0, 2, 3, 1).contiguous().bfloat16().to(dev)
    packed = ops.pack_conv_weight(w.to(dev), kind, cout, cin, in_layout)
    bh = gh = None
    if epi in (_lib.EPI_GDN, _lib.EPI_IGDN):
        bh, gh = ops.gdn_pack(torch.ones(cout, device=dev), (0.1 * torch.eye(cout)).sqrt().to(dev), 0.0, 0.0, 0.0)
    kw = dict(kind=kind, epilogue=epi, in_layout=in_layout, out_layout=out_layout, in_c=cin, out_c=cout,
              weight=packed, bias=bias.to(dev), beta=bh, gamma=gh)
    for _ in range(2):
        ops.conv_forward(xd, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        ops.conv_forward(xd, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    probe.zero_()
    _set_probe(probe.data_ptr())
    ops.conv_forward(xd, **kw)
    torch.cuda.synchronize()
    _set_probe(None)
    p = probe.view(-1, 16).cpu()
    p = p[p[:, 13] > 0].double()
    tiles = p[:, 13].mean().item()
    macs = {0: 25, 1: 25, 2: 9}[kind] * cin * cout * (H // 2) * (W // 2) if kind == 0 else 25 * cin * cout * H * W
    tf = 2 * macs * B / (ms * 1e-3) / 1e12
    print(f"== {name} B={B}: {ms:.3f} ms  {tf:.1f} TFLOP/s  ctas={p.shape[0]} tiles/cta={tiles:.1f}")
    per = {n: p[:, i].mean().item() / max(tiles, 1) for i, n in enumerate(NAMES)}
    print("   cycles per tile: " + "  ".join(f"{n}={per[n]:.0f}" for n in NAMES if n != "tiles"))
