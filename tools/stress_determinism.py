"""Runs the same 256-tile codec step many times and checks that x_hat / symbols are bit-identical every time (a lost
cross-CTA ordering in the CTA-pair engine would show up as run-to-run differences)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L  # noqa: E402
from licos_b200 import synth  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(42)
net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False)
synth.condition_weights(net)
net = net.to(dev).eval()
x = (synth.make_input("rgb256", 256, seed=3, device=dev) * 255).round().to(torch.uint8)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
ref = None
bad = 0
with torch.no_grad():
    for i in range(n):
        out = net.forward_tiles(x, out_dtype=torch.uint8)
        cur = (out["x_hat"].clone(), out["symbols"].clone())
        if ref is None:
            ref = cur
        elif not (torch.equal(cur[0], ref[0]) and torch.equal(cur[1], ref[1])):
            bad += 1
torch.cuda.synchronize()
print(f"{n} runs, {bad} differed from the first")
sys.exit(1 if bad else 0)
