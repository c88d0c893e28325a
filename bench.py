#!/usr/bin/env python
"""Benchmark of the learned-codec hot path (BASELINE.json: encode/decode MPix/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic tiles that is already resident in HBM:
    encode: y = g_a(x); y_hat, likelihoods = entropy_bottleneck(y); symbols = round(y - median)
    decode: x_hat = g_s(y_hat)
Workload at N = 1: BASELINE.json configs[1] (bmshj2018-factorized q1, N=128 / M=192, 256 tiles of 3x256x256).
N > 1 (torchrun, one rank per GPU): every rank codes its own 256 tiles -- tiles are independent, so there is
no data-path collective (weak scaling).  Prints ONE JSON line on rank 0.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "bmshj2018-factorized"
QUALITY = 1
TILE = (3, 256, 256)
# 2 x MACs of conv + deconv + GDN for one 3x256x256 tile (5.528 GFLOP each way)
FLOP_PER_TILE_ENCODE = 2 * 2764.05e6
FLOP_PER_TILE_DECODE = 2 * 2764.05e6
# The dominant kernel (conv_igemm_kernel) runs g_a[2], g_a[4], g_a[6], g_s[0], g_s[2], g_s[4] with their fused GDN / IGDN:
# MACs per tile = conv + gamma GEMM = (1677.7 + 67.1) + (419.4 + 16.8) + 157.3 + (157.3 + 16.8) + (419.4 + 67.1) + (1677.7 + 268.4) M
FLOP_PER_TILE_ENGINE = 2 * 4945.0e6
# HBM-bound kernels: algorithmic bytes per tile (every tensor read once, written once)
BYTES_PER_TILE_FIRST = 3 * 256 * 256 * 4 + 128 * 128 * 128 * 2   # g_a[0]: fp32 NCHW x in, bf16 NHWC out
BYTES_PER_TILE_LAST = 128 * 128 * 128 * 2 + 3 * 256 * 256 * 4    # g_s[6]: bf16 NHWC in, fp32 NCHW x_hat out
BYTES_PER_TILE_EB = 192 * 16 * 16 * 18                           # y in; y_hat, likelihoods, int32 symbols, bf16 NHWC y_hat out


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="tiles per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=16, help="tiles per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=32)
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE.json configs[4])")
    ap.add_argument("--min-warmup", type=int, default=5, help="the caching allocator needs ~5 steps to settle")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "MEASURED_PEAKS.json (bf16_tflops_sustained: kernel timed inside a long step)"}
    return {"hbm_gbs": 6650.0, "tflops": 1400.0, "source": "fallback (B200_PROFILING.md)"}


# ---------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the oracle restatement on the host cores
# ---------------------------------------------------------------------------------------------

def cpu_reference(batch: int, steps: int, warmup: int, state_dict=None):
    """Times the oracle (pure PyTorch CPU restatement of the CompressAI path) on `batch` tiles per step."""
    import torch
    from licos_b200 import synth
    from oracle import compressai_ref as R  # allowed here only: cpu_baseline / --impl reference legs

    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    ref = R.image_models[MODEL](quality=QUALITY)
    if state_dict is None:
        synth.condition_weights(ref)
    else:
        ref.load_state_dict(state_dict)
    ref.eval()
    x = synth.make_input("rgb256", batch)
    med = ref.entropy_bottleneck.quantiles[:, 0, 1].detach().reshape(1, -1, 1, 1)
    times = []
    with torch.no_grad():
        for i in range(warmup + steps):
            t0 = time.perf_counter()
            y = ref.g_a(x)
            y_hat, lik = ref.entropy_bottleneck(y)
            sym = ref.entropy_bottleneck.quantize(y, "symbols", med)
            x_hat = ref.g_s(y_hat)
            dt = time.perf_counter() - t0
            if i >= warmup:
                times.append(dt)
    pix = batch * TILE[1] * TILE[2]
    med_t = statistics.median(times)
    del sym, lik, x_hat
    return {"value": pix / med_t / 1e6, "unit": "MPix/s", "cores": threads, "kind": "port",
            "sample": f"{batch} tiles of 3x256x256 per step, median of {steps} steps after {warmup} warm-ups, "
                      f"torch {torch.__version__} CPU fp32, {threads} threads",
            "ms_per_step": med_t * 1e3, "best_mpix_s": pix / min(times) / 1e6}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, min(args.steps, 10))
    warmup = max(1, min(args.warmup, 2))
    r = cpu_reference(args.ref_batch, steps, warmup)
    line = {
        "impl": "reference", "metric": "encode+decode throughput", "value": r["value"], "unit": "MPix/s",
        "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} q{QUALITY} encode+decode, {args.ref_batch}-tile sample of the "
                               f"256 x 3x256x256 batch (CPU reference arm: oracle port of the CompressAI path)"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------------------------
# clock sampling
# ---------------------------------------------------------------------------------------------

class ClockSampler:
    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _loop(self):
        nv = self.nv
        names = {"hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
                 "hw_power_brake": 0x80}
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if mask & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            self._stop.wait(0.05)

    def start(self):
        if self.nv is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self):
        self._stop.set()
        if self._thread:
            self._thread.join()
        if not self.samples:
            try:
                out = subprocess.check_output(
                    ["nvidia-smi", "-i", str(self.index), "--query-gpu=clocks.sm,clocks.max.sm",
                     "--format=csv,noheader,nounits"], text=True).strip().split(",")
                return {"sm_mhz": float(out[0]), "sm_max_mhz": float(out[1]), "reasons": ["unsampled"]}
            except Exception:
                return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        return {"sm_mhz": statistics.median(self.samples), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------
# B200 arm
# ---------------------------------------------------------------------------------------------

def bind_to_gpu_numa_node(torch, local: int):
    """Multi-GPU runs: pin this rank (and therefore the pinned host buffers it first-touches) to the NUMA node its GPU
    hangs off, when the host exposes one.  Without it the host-buffer leg of ranks on the far socket crosses the
    inter-socket link in both directions.  Returns the node or None; never raises."""
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:  # noqa: BLE001  (no sysfs, no such attribute, containerised cpuset: run unbound)
        return None


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    numa_node = bind_to_gpu_numa_node(torch, local) if world > 1 else None
    if world > 1:
        dist.init_process_group("nccl", device_id=device)

    import licos_b200 as L
    from licos_b200 import _lib, ops, synth

    if _lib.lib.licos_device_ok(local) != 0:
        raise SystemExit("bench.py: no sm_100 device -- the licos_b200 path has no fallback")

    # ---- model, weights, input (identical on every rank; each rank codes its own tiles) ----
    torch.manual_seed(42)
    net = L.image_models[MODEL](quality=QUALITY, pretrained=False)
    synth.condition_weights(net)
    net = net.to(device).eval()
    B = args.batch
    x = synth.make_input("rgb256", B, seed=42 + rank, device=device)
    eb = net.entropy_bottleneck

    launches = {"n": 0}
    conv_ms = {"events": None}
    orig_conv = ops.conv_forward

    def counted_conv(*a, **k):
        # one igemm launch (+ the first-layer im2col launch) per call
        launches["n"] += 2 if k.get("in_layout") == _lib.LAYOUT_NCHW_F32 else 1
        if conv_ms["events"] is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = orig_conv(*a, **k)
            e1.record()
            conv_ms["events"].append((e0, e1, k["kind"], k["in_c"], k["out_c"], tuple(out.shape)))
            return out
        return orig_conv(*a, **k)

    ops.conv_forward = counted_conv

    eb_ms = {"events": None}

    def encode(xb):
        y = net.g_a(xb)
        if eb_ms["events"] is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            y_hat, lik, sym, y_nhwc = eb.forward_fused(y, want_symbols=True, want_nhwc=True)
            e1.record()
            eb_ms["events"].append((e0, e1))
        else:
            y_hat, lik, sym, y_nhwc = eb.forward_fused(y, want_symbols=True, want_nhwc=True)
        return y, y_hat, lik, sym, y_nhwc

    def step(xb):
        y, y_hat, lik, sym, y_nhwc = encode(xb)
        x_hat = net.g_s(y_hat, nhwc=y_nhwc)
        # kernels besides the convs: EB likelihood table + the fused quantise / likelihood / symbols / NHWC pass
        launches["n"] += 2
        return y_hat, lik, sym, x_hat

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        n_warm = max(args.warmup, args.min_warmup)
        for _ in range(n_warm):
            step(x)
        sync_all()

        sampler = ClockSampler(local)
        if not os.environ.get("LICOS_BENCH_NOSAMPLER"):
            sampler.start()
        launches["n"] = 0
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        enc_ms = dec_ms = 0.0
        sync_all()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        marks = []
        t_start.record()
        for _ in range(args.steps):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            y, y_hat, lik, sym, y_nhwc = encode(x)
            e1.record()
            x_hat = net.g_s(y_hat, nhwc=y_nhwc)
            e2.record()
            launches["n"] += 2
            marks.append((e0, e1, e2))
        t_end.record()
        sync_all()
        clocks = sampler.stop()
        total_ms = t_start.elapsed_time(t_end)
        for e0, e1, e2 in marks:
            enc_ms += e0.elapsed_time(e1)
            dec_ms += e1.elapsed_time(e2)
        n_launch = launches["n"]

        # ---- per-launch durations of the dominant kernel (conv igemm), live, CUDA events on this stream ----
        conv_ms["events"] = []
        eb_ms["events"] = []
        inst_steps = 5
        for _ in range(inst_steps):
            step(x)
        torch.cuda.synchronize()
        per_layer = {}
        conv_total_ms = engine_total_ms = first_total_ms = last_total_ms = 0.0
        for e0, e1, kind, cin, cout, shape in conv_ms["events"]:
            ms = e0.elapsed_time(e1)
            conv_total_ms += ms
            if cin <= 16:
                first_total_ms += ms          # conv_first2_kernel
            elif cout <= 4:
                last_total_ms += ms           # deconv_narrow2_kernel
            else:
                engine_total_ms += ms         # conv_igemm_kernel
            key = f"{['conv5s2', 'deconv5s2', 'conv3s1'][kind]}_{cin}->{cout}_{shape[-2] if len(shape) == 4 else ''}"
            per_layer.setdefault(key, []).append(ms)
        eb_total_ms = sum(e0.elapsed_time(e1) for e0, e1 in eb_ms["events"])
        conv_ms["events"] = None
        eb_ms["events"] = None
        conv_ms_per_step = conv_total_ms / inst_steps
        engine_ms_per_step = engine_total_ms / inst_steps
        first_ms, last_ms, eb_ms_step = first_total_ms / inst_steps, last_total_ms / inst_steps, eb_total_ms / inst_steps

        # ---- end to end through the public model API with HOST buffers ----
        if args.no_e2e:
            e2e = {"ms_per_step": float("nan"), "h2d": 0, "d2h": 0}
        else:
            e2e = run_e2e(net, eb, x, args, torch, device, world)

    # max over ranks
    t = torch.tensor([total_ms, enc_ms, dec_ms, conv_ms_per_step, e2e["ms_per_step"], engine_ms_per_step, first_ms,
                      last_ms, eb_ms_step], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, enc_ms, dec_ms, conv_ms_per_step, e2e_ms, engine_ms_per_step, first_ms, last_ms, eb_ms_step = t.tolist()

    pix_per_step = B * TILE[1] * TILE[2] * world
    ms_per_step = total_ms / args.steps
    value = pix_per_step / (ms_per_step * 1e-3) / 1e6
    pk = peaks()
    achieved_tflops = B * FLOP_PER_TILE_ENGINE / (engine_ms_per_step * 1e-3) / 1e12

    def hbm(name, bytes_per_tile, ms, note):
        gbs = B * bytes_per_tile / (ms * 1e-3) / 1e9
        return {"kernel": name, "bound": "hbm", "achieved": gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / pk["hbm_gbs"], "ms_per_launch": round(ms, 4), "note": note}
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("conv_igemm_dram_bytes_per_launch")

    if rank == 0:
        train = None
        if not args.no_train:  # before the CPU baseline: its idle worker threads would slow the eager Python loop
            torch.cuda.empty_cache()
            try:  # an extra leg: it must never cost the headline line
                train = run_train_step(torch, device)
            except Exception as e:  # noqa: BLE001
                train = {"error": f"{type(e).__name__}: {e}"}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            sd = {k: v.cpu() for k, v in net.state_dict().items()}
            cpu = cpu_reference(args.ref_batch, steps=6, warmup=2, state_dict=sd)
        line = {
            "metric": "encode+decode throughput", "value": value, "unit": "MPix/s", "n_gpus": world,
            "steps": args.steps, "warmup": n_warm, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": f"{MODEL} q{QUALITY} (N=128, M=192) encode+decode of {B} synthetic 3x256x256 tiles "
                                   f"per GPU (BASELINE.json configs[1])",
                       "tiles_per_gpu": B, "l2": "inputs larger than L2 (201 MB input, >1 GB of activations per step)",
                       "weights": "random init seed 42 + synth.condition_weights"},
            "encode_mpix_s": pix_per_step / (enc_ms / args.steps * 1e-3) / 1e6,
            "decode_mpix_s": pix_per_step / (dec_ms / args.steps * 1e-3) / 1e6,
            "clocks": clocks,
            "e2e": {"value": pix_per_step / (e2e_ms * 1e-3) / 1e6, "unit": "MPix/s",
                    "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "path": "pinned host x -> model.g_a / entropy_bottleneck / g_s -> x_hat, symbols, bpp on host",
                    "rank0_numa_node": numa_node},
            "gpu_launches": n_launch,
            "roofline": {"bound": "tensor",
                         "kernel": "conv_igemm_kernel (6 launches per step: g_a[2,4,6], g_s[0,2,4], GDN/IGDN fused)",
                         "achieved": achieved_tflops, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved_tflops / pk["tflops"], "traffic": traffic,
                         "peak_source": pk["source"],
                         "algorithmic": "2 x (conv + gamma-GEMM MACs) of the 6 layers = 9.89 GFLOP per tile, summed over "
                                        "the 6 launches / their summed CUDA-event durations (an instrumented pass)",
                         "per_launch_ms": {k: round(statistics.mean(v), 4) for k, v in per_layer.items()},
                         "share_of_step": engine_ms_per_step / ms_per_step,
                         "conv_share_of_step": conv_ms_per_step / ms_per_step},
            "roofline_hbm": [
                hbm("conv_first2_kernel (g_a[0] + GDN)", BYTES_PER_TILE_FIRST, first_ms, "fp32 NCHW x in, bf16 NHWC out"),
                hbm("deconv_narrow2_kernel (g_s[6])", BYTES_PER_TILE_LAST, last_ms, "bf16 NHWC in, fp32 NCHW x_hat out"),
                hbm("eb_lut_kernel + eb_eval_tile_kernel (quantise + likelihoods + symbols + NHWC copy)", BYTES_PER_TILE_EB,
                    eb_ms_step, "18 B per latent element (two launches; 12 B for forward() alone)"),
            ],
        }
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        if train is not None:
            line["train_step"] = train
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_train_step(torch, device, steps: int = 20):
    """BASELINE.json configs[4], the per-rank part: train.py's train_one_batch body (forward in train mode, RD loss,
    backward, clip, Adam, aux step) on 32 synthetic 3x256x256 tiles, every kernel native (forward + dgrad on the conv
    engine, wgrad / GDN / bottleneck backward kernels), replayed as one CUDA graph and, for comparison, as the eager
    Python loop.  Reported beside the headline metric, not part of it."""
    import licos_b200 as L
    from licos_b200 import synth

    torch.manual_seed(100)
    net = L.image_models[MODEL](quality=QUALITY, pretrained=False).to(device).train()
    crit = L.RateDistortionLoss(lmbda=1e-2)
    opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
    tiles = 32
    x = synth.make_input("rgb256", tiles, seed=1, device=device)

    def eager():
        opt["net"].zero_grad(); opt["aux"].zero_grad()
        out = net(x)
        loss = crit(out, x)["loss"]
        loss.backward()
        torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
        opt["net"].step()
        aux = net.aux_loss()
        aux.backward()
        opt["aux"].step()
        return loss.detach()

    def timed(fn, n):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            last = fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n, last

    loss0 = float(eager())
    eager_ms, _ = timed(eager, steps)
    graphed = L.GraphedTrainStep(net, crit, opt, x, clip_max_norm=1.0)
    graph_ms, last = timed(lambda: graphed(x)["loss"], steps)
    flop = tiles * 3 * (5.528e9 * 2)  # forward + dgrad + wgrad of the 11.056 GFLOP / tile transforms
    pix = tiles * TILE[1] * TILE[2]
    return {"workload": f"{MODEL} q{QUALITY} training step (forward + RD loss + backward + clip + Adam + aux step) on "
                        f"{tiles} synthetic 3x256x256 tiles (BASELINE.json configs[4], one rank's share)",
            "ms_per_step": graph_ms, "mpix_s": pix / (graph_ms * 1e-3) / 1e6, "how": "licos_b200.GraphedTrainStep (CUDA graph replay)",
            "eager_loop_ms_per_step": eager_ms, "eager_loop_mpix_s": pix / (eager_ms * 1e-3) / 1e6,
            "tflops": flop / (graph_ms * 1e-3) / 1e12, "loss_first": loss0, "loss_last": float(last),
            "library_baseline": "same step through cuDNN autograd, measured in round 1 with LICOS_EAGER_AUTOGRAD=1 tools/bench_train.py "
                                "(profiles/r1_train_step.json): 15.2 ms"}


def run_e2e(net, eb, x_dev, args, torch, device, world):
    """Same step through the public API with host buffers: H2D of the tiles, D2H of x_hat, symbols, bpp."""
    import torch.distributed as dist
    from licos_b200 import ops

    B = x_dev.shape[0]
    x_host = x_dev.cpu().pin_memory()
    ysz = (B, eb.channels, TILE[1] // 16, TILE[2] // 16)
    xhat_host = torch.empty(x_host.shape, dtype=torch.float32).pin_memory()
    sym_host = torch.empty(ysz, dtype=torch.int32).pin_memory()
    bpp_host = torch.empty(16, dtype=torch.float64).pin_memory()
    chunk = max(1, min(args.e2e_chunk, B))
    streams = [torch.cuda.Stream(device=device) for _ in range(3)]
    steps = max(3, min(args.steps, 10))

    # Tiles stream through three CUDA streams, chunk by chunk, continuously across steps (no join between steps: a
    # deployed coder does not drain its pipeline every 256 tiles).  Every step's inputs are copied from pinned host
    # memory and its x_hat, symbols and bpp land in pinned host memory inside the timed region.
    def run(n_steps):
        accs = torch.zeros(n_steps, len(streams), dtype=torch.float64, device=device)
        main = torch.cuda.current_stream()
        for s in streams:
            s.wait_stream(main)
        k = 0
        for it in range(n_steps):
            for lo in range(0, B, chunk):
                hi = min(B, lo + chunk)
                si = k % len(streams)
                k += 1
                with torch.cuda.stream(streams[si]):
                    xb = x_host[lo:hi].to(device, non_blocking=True)
                    y = net.g_a(xb)
                    y_hat, lik, sym, y_nhwc = eb.forward_fused(y, want_symbols=True, want_nhwc=True)
                    x_hat = net.g_s(y_hat, nhwc=y_nhwc)
                    ops.sum_log(lik, accs[it, si:si + 1])  # per-(step, stream) accumulator, summed after the join
                    xhat_host[lo:hi].copy_(x_hat, non_blocking=True)
                    sym_host[lo:hi].copy_(sym, non_blocking=True)
        for s in streams:
            main.wait_stream(s)
        bpp = accs.sum(dim=1) / (-math.log(2) * B * TILE[1] * TILE[2])
        bpp_host[:n_steps].copy_(bpp, non_blocking=True)

    run(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(steps)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / steps
    return {"ms_per_step": dt * 1e3, "h2d": x_host.numel() * 4,
            "d2h": xhat_host.numel() * 4 + sym_host.numel() * 4 + 8, "bpp": float(bpp_host[0].item())}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
