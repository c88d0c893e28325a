"""cfg 5 on real GPUs (run under torchrun, one rank per GPU): a federated training step -- forward + backward +
rate-distortion loss on 32 tiles per rank (native backward kernels), Adam step -- followed by the NCCL weight merge that
replaces the reference's checkpoint-file exchange (federation_utils.py:27-85).  Checks that every rank ends with the
same weighted average and times the merge.  Prints one JSON line on rank 0."""
import json
import os
import sys
import time

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import licos_b200 as L  # noqa: E402
from licos_b200 import synth  # noqa: E402
from licos_b200.federated import FlatState, federated_average  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.manual_seed(100 + rank)
net = L.image_models["bmshj2018-factorized"](quality=1, pretrained=False).to(dev).train()
state = FlatState(net)
crit = L.RateDistortionLoss(lmbda=1e-2)
opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
x = synth.make_input("rgb256", 32, seed=rank, device=dev)

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
    return float(loss["loss"])

for _ in range(3):
    loss = train_step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    loss = train_step()
torch.cuda.synchronize()
eager_ms = (time.perf_counter() - t0) / 10 * 1e3
graphed = L.GraphedTrainStep(net, crit, opt, x, clip_max_norm=1.0)
for _ in range(3):
    graphed(x)
torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    out_terms = graphed(x)
torch.cuda.synchronize()
step_ms = (time.perf_counter() - t0) / 20 * 1e3
loss = float(out_terms["loss"])

before = state.flat.clone()
gathered = [torch.empty_like(before) for _ in range(world)]
dist.all_gather(gathered, before)
losses = torch.zeros(world, dtype=torch.float64, device=dev)
losses[rank] = loss
dist.all_reduce(losses)
inv = 1.0 / losses
w = (inv / inv.sum()).tolist()
expect = sum(wi * g.double() for wi, g in zip(w, gathered)).float()
federated_average(state, loss)
torch.cuda.synchronize()
err = (state.flat - expect).abs().max().item()
same = [torch.empty_like(state.flat) for _ in range(world)]
dist.all_gather(same, state.flat)
identical = all(torch.equal(same[0], s) for s in same)
# time the merge alone
for _ in range(3):
    federated_average(state, loss)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
dist.barrier()
e0.record()
for _ in range(20):
    federated_average(state, loss)
e1.record()
torch.cuda.synchronize()
merge_us = e0.elapsed_time(e1) / 20 * 1e3
out = net.eval()(x[:2]) if True else None  # the fused kernels pick up the merged weights (cache keyed on versions)
if rank == 0:
    print(json.dumps({"config": "cfg5 federated step", "world": world, "tiles_per_rank": 32, "train_step_ms": step_ms,
                      "train_mpix_s": 32 * 65536 * world / step_ms / 1e3, "loss": loss, "merge_bytes": state.numel * 4,
                      "merge_us": merge_us, "merge_max_abs_err_vs_weighted_mean": err, "ranks_identical": identical,
                      "eager_loop_step_ms": eager_ms,
                      "train_path": "native sm_100a kernels (forward, dgrad, wgrad, GDN / bottleneck backward), "
                                    "train_step_ms = CUDA-graph replay (GraphedTrainStep), eager_loop_step_ms = Python loop"}))
dist.barrier()
dist.destroy_process_group()
