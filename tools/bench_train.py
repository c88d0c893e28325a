"""cfg 5 training step on one GPU (licos/train.py:148-212 train_one_batch): forward in train mode + rate-distortion
loss + backward + clip + Adam + aux step, 32 tiles of 3x256x256.  Times the whole step with CUDA events and, with
--profile, prints the per-kernel share from torch.profiler.  (The cuDNN-autograd baseline of the same step is timed by
bench.py's gpu_library_baseline leg, with the oracle's modules on the GPU.)  One JSON line."""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L  # noqa: E402
from licos_b200 import synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=32)
ap.add_argument("--steps", type=int, default=10)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--profile", action="store_true")
ap.add_argument("--model", default="bmshj2018-factorized")
ap.add_argument("--graph", action="store_true", help="replay the step as one CUDA graph (licos_b200.GraphedTrainStep)")
args = ap.parse_args()

dev = torch.device("cuda", 0)
torch.manual_seed(100)
net = L.image_models[args.model](quality=1, pretrained=False).to(dev).train()
crit = L.RateDistortionLoss(lmbda=1e-2)
opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
x = synth.make_input("rgb256", args.batch, seed=0, device=dev)


def train_step():
    opt["net"].zero_grad(); opt["aux"].zero_grad()
    out = net(x)
    loss = crit(out, x)
    loss["loss"].backward()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
    opt["net"].step()
    aux = net.aux_loss()
    aux.backward()
    opt["aux"].step()
    return loss["loss"]


def fwd_bwd_transforms():
    """g_a and g_s alone (forward + backward), the part the native kernels cover."""
    y = net.g_a(x)
    xh = net.g_s(y)
    (xh - x).square().mean().backward()


def timed(fn, steps, warmup):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


graphed = None
if args.graph:
    graphed = L.GraphedTrainStep(net, crit, opt, x, clip_max_norm=1.0)
    eager_step = train_step

    def train_step():  # noqa: F811
        return graphed(x)["loss"]

step_ms = timed(train_step, args.steps, args.warmup)
loss = float(train_step().detach())
tr_ms = timed(fwd_bwd_transforms, args.steps, 2)
res = {"config": f"cfg5 training step, {args.model} q1, {args.batch} x 3x256x256", "path": "native sm_100a forward + dgrad + wgrad", "step_ms": step_ms,
    "graph": bool(args.graph), "train_mpix_s": args.batch * 65536 / step_ms / 1e3, "transforms_fwd_bwd_ms": tr_ms, "loss": loss}
if args.profile:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            train_step()
        torch.cuda.synchronize()
    rows = sorted((e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"),
                  key=lambda e: -e.device_time_total)
    tot = sum(e.device_time_total for e in rows)
    res["kernels_ms_per_step"] = {e.key[:70]: round(e.device_time_total / 3e3, 4) for e in rows[:28]}
    res["kernel_sum_ms_per_step"] = tot / 3e3
    res["launches_per_step"] = sum(e.count for e in rows) / 3
print(json.dumps(res))
