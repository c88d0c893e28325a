"""Throughput of the other BASELINE.json configurations (not the bench.py headline line):
  cfg3  raw-split: bmshj2018-factorized with the reference's 1-band surgery, 1x512x512 12-bit tiles
  cfg4  bmshj2018-hyperprior N=192 / M=320 on 3x1024x1024 crops
Usage: python tools/bench_configs.py [cfg3|cfg4|all] [batch]   -- prints one JSON line per configuration."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L  # noqa: E402
from licos_b200 import _lib, ops, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda:0")


def run(name, net, x, steps=10, warmup=5):
    net = net.to(dev).eval()
    x = x.to(dev)
    events = []
    orig = ops.conv_forward
    rec = {"on": False}

    def timed(*a, **k):
        if not rec["on"]:
            return orig(*a, **k)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(*a, **k)
        e1.record()
        events.append((e0, e1, k["kind"], k["in_c"], k["out_c"], tuple(out.shape)))
        return out

    ops.conv_forward = timed
    with torch.no_grad():
        for _ in range(warmup):
            out = net(x)
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(steps):
            out = net(x)
        t1.record()
        torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / steps
        rec["on"] = True
        for _ in range(3):
            net(x)
        torch.cuda.synchronize()
    ops.conv_forward = orig
    per = {}
    for e0, e1, kind, cin, cout, shape in events:
        key = f"{['conv5s2', 'deconv5s2', 'conv3s1'][kind]}_{cin}->{cout}_{shape[-2] if len(shape) == 4 else ''}"
        per.setdefault(key, []).append(e0.elapsed_time(e1))
    pix = x.shape[0] * x.shape[2] * x.shape[3]
    lik = out["likelihoods"]
    bpp = sum(torch.log(v).sum().item() for v in lik.values()) / (-0.6931471805599453 * pix)
    print(json.dumps({"config": name, "forward_ms": ms, "forward_mpix_s": pix / ms / 1e3, "bpp": bpp,
                      "input": list(x.shape), "per_layer_ms": {k: round(sum(v) / len(v), 4) for k, v in per.items()}}))


if which in ("cfg3", "all"):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    torch.manual_seed(42)
    net = L.get_model("bmshj2018-factorized", pretrained=False, in_channels=1, quality=1) if hasattr(L, "get_model") else None
    synth.condition_weights(net)
    run(f"cfg3 raw-split 1x512x512 x{B}", net, synth.make_input("raw512", B))
if which in ("cfg4", "all"):
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
    torch.manual_seed(42)
    net = L.image_models["bmshj2018-hyperprior"](quality=6, pretrained=False)
    synth.condition_weights(net)
    run(f"cfg4 hyperprior N=192 M=320 3x1024x1024 x{B}", net, synth.make_input("rgb1024", B))
