#!/usr/bin/env python
"""Benchmark of the learned-codec hot path (BASELINE.json: encode/decode MPix/s).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch of synthetic 8-bit tiles that is already resident in HBM:
    encode: y = g_a(tiles); y_hat, likelihoods = entropy_bottleneck(y); symbols = round(y - median); rate term
    decode: x_hat = g_s(y_hat)
Workload at N = 1: BASELINE.json configs[1] (bmshj2018-factorized q1, N=128 / M=192, 256 tiles of 3x256x256).
N > 1 (torchrun, one rank per GPU): every rank codes its own 256 tiles -- tiles are independent, so there is no
data-path collective (weak scaling).  Prints ONE JSON line on rank 0.

The tiles are 8-bit pixels (uint8, what imagery is): the first layer scales them in its patch builders and the last layer
writes uint8 pixels, bit-identical to the fp32-in / fp32-out path on the same pixel values (tests/test_gpu_tiles.py); the
fp32 path is timed beside it (`fp32_io`).  `e2e` is the same step with HOST buffers: pinned uint8 tiles up, uint8 x_hat +
int16 symbols + bpp down, copies inside the timed region.

Extra legs in the same line (every rank takes part, max over ranks): `cfg3` (BASELINE configs[2]: 512 single-band 12-bit
512x512 tiles, STRONG scaling over the ranks), `cfg4` (configs[3]: hyperprior q6 on 3x1024x1024, one crop per GPU),
`cfg5` (configs[4]: graph-replayed training step on 32 tiles per rank + the NCCL weight merge inside the timed region).
"""
from __future__ import annotations

import argparse
import importlib.util
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
BYTES_PER_TILE_FIRST = 3 * 256 * 256 * 1 + 128 * 128 * 128 * 2   # g_a[0]: uint8 NCHW tiles in, bf16 NHWC out
BYTES_PER_TILE_LAST = 128 * 128 * 128 * 2 + 3 * 256 * 256 * 1    # g_s[6]: bf16 NHWC in, uint8 NCHW x_hat out
BYTES_PER_TILE_EB = 192 * 16 * 16 * 16                           # y in; y_hat, likelihoods, int16 symbols, bf16 NHWC y_hat out
# cfg 3 / cfg 4
FLOP_PER_TILE_CFG3 = 42.547e9    # 1x512x512 raw split, forward
FLOP_PER_IMG_CFG4 = 406.77e9     # hyperprior q6, 3x1024x1024, forward


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=256, help="tiles per GPU per step")
    ap.add_argument("--ref-batch", type=int, default=16, help="tiles per step of the CPU reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=128)
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE.json configs[4])")
    ap.add_argument("--no-legs", action="store_true", help="skip the cfg3 / cfg4 / library-baseline legs (profiling runs)")
    ap.add_argument("--min-warmup", type=int, default=5, help="the caching allocator needs ~5 steps to settle")
    return ap.parse_args()


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": p["hbm_gbs"], "tflops": p["bf16_tflops"],
                "tflops_sustained": p.get("bf16_tflops_sustained", p["bf16_tflops"]),
                "source": "MEASURED_PEAKS.json (bf16_tflops: the burst figure -- the timed region is tens of ms at full clock; "
                          "frac_sustained uses bf16_tflops_sustained)"}
    return {"hbm_gbs": 6650.0, "tflops": 1650.0, "tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


def load_synth():
    """licos_b200/synth.py by path: the weight / input recipe WITHOUT importing the package, so that the CPU reference arm
    never maps liblicos_b200.so."""
    spec = importlib.util.spec_from_file_location("licos_synth_standalone", os.path.join(ROOT, "licos_b200", "synth.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def make_tiles_u8(torch, batch: int, seed: int):
    g = torch.Generator().manual_seed(seed)
    return torch.randint(0, 256, (batch, *TILE), generator=g, dtype=torch.uint8)


# ---------------------------------------------------------------------------------------------
# CPU reference arm / baseline: the oracle restatement on the host cores
# ---------------------------------------------------------------------------------------------

def cpu_reference(batch: int, steps: int, warmup: int, state_dict=None):
    """Times the oracle (pure PyTorch CPU restatement of the CompressAI path) on `batch` tiles per step."""
    import torch
    from oracle import compressai_ref as R  # allowed here only: cpu_baseline / --impl reference legs

    synth = load_synth()
    threads = os.cpu_count() or 1
    torch.set_num_threads(threads)
    torch.manual_seed(42)
    ref = R.image_models[MODEL](quality=QUALITY)
    if state_dict is None:
        synth.condition_weights(ref)
    else:
        ref.load_state_dict(state_dict)
    ref.eval()
    x = (make_tiles_u8(torch, batch, 42).double() / 255).float()   # the same 8-bit pixels, as the reference's loaders ship them
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
            self._stop.wait(0.02)

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


class Timer:
    """CUDA-event timing of `steps` calls bracketed by a barrier + synchronize on both sides."""

    def __init__(self, torch, dist, world):
        self.torch, self.dist, self.world = torch, dist, world

    def sync(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def run(self, fn, steps: int, warmup: int) -> float:
        torch = self.torch
        for _ in range(warmup):
            fn()
        self.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        self.sync()
        return e0.elapsed_time(e1) / steps


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
    timer = Timer(torch, dist, world)

    # ---- model, weights, input (identical weights on every rank; each rank codes its own tiles) ----
    torch.manual_seed(42)
    net = L.image_models[MODEL](quality=QUALITY, pretrained=False)
    synth.condition_weights(net)
    net = net.to(device).eval()
    B = args.batch
    tiles_host = make_tiles_u8(torch, B, 42 + rank)
    tiles = tiles_host.to(device)                                   # uint8 (B, 3, 256, 256), resident in HBM
    x32 = (tiles_host.double() / 255).float().to(device)            # the same pixels as the fp32 tensor the reference ships
    eb = net.entropy_bottleneck
    rate = torch.zeros(1, dtype=torch.float64, device=device)

    launches = {"n": 0}
    conv_ms = {"events": None}
    orig_conv = ops.conv_forward

    def counted_conv(*a, **k):
        launches["n"] += 1  # one kernel per layer at these shapes (pipelined first layer, engine, narrow last layer)
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

    def encode(v):
        y = net.g_a(v)
        if eb_ms["events"] is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            res = eb.forward_fused(y, want_nhwc=True, want_symbols_i16=True, sum_ln=rate)
            e1.record()
            eb_ms["events"].append((e0, e1))
        else:
            res = eb.forward_fused(y, want_nhwc=True, want_symbols_i16=True, sum_ln=rate)
        launches["n"] += 1  # the fused quantise / likelihood / symbols / NHWC / rate pass (its table is cached per weights)
        y_hat, lik, _, y_nhwc, sym16 = res
        return y_hat, lik, sym16, y_nhwc

    def decode(y_hat, y_nhwc):
        return net.g_s(y_hat, nhwc=y_nhwc, out_dtype=torch.uint8, out_max=255)

    def step(v):
        y_hat, lik, sym16, y_nhwc = encode(v)
        return y_hat, lik, sym16, decode(y_hat, y_nhwc)

    def step_fp32(x):  # the CompressAI-facing formats end to end: fp32 tiles in, fp32 x_hat and int32 symbols out
        y = net.g_a(x)
        y_hat, lik, sym, y_nhwc = eb.forward_fused(y, want_symbols=True, want_nhwc=True)
        return y_hat, lik, sym, net.g_s(y_hat, nhwc=y_nhwc)

    with torch.no_grad():
        n_warm = max(args.warmup, args.min_warmup)
        for _ in range(n_warm):
            step(tiles)
        timer.sync()

        sampler = ClockSampler(local)
        if not os.environ.get("LICOS_BENCH_NOSAMPLER"):
            sampler.start()
        launches["n"] = 0
        enc_ms = dec_ms = 0.0
        timer.sync()
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        marks = []
        t_start.record()
        for _ in range(args.steps):
            e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
            e0.record()
            y_hat, lik, sym16, y_nhwc = encode(tiles)
            e1.record()
            x_hat = decode(y_hat, y_nhwc)
            e2.record()
            marks.append((e0, e1, e2))
        t_end.record()
        timer.sync()
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
            step(tiles)
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

        # ---- the fp32-in / fp32-out formats of the CompressAI API on the same pixels ----
        fp32_ms = timer.run(lambda: step_fp32(x32), steps=min(args.steps, 20), warmup=3)
        del x32

        # ---- end to end through the public model API with HOST buffers ----
        if args.no_e2e:
            e2e = {"ms_per_step": float("nan"), "h2d": 0, "d2h": 0, "copy_ms_per_step": float("nan")}
        else:
            e2e = run_e2e(net, tiles_host, args, torch, device, world)
        del tiles
        torch.cuda.empty_cache()

        # ---- the other BASELINE configurations, every rank taking part ----
        legs = {}
        if not args.no_legs:
            for name, fn in (("cfg3", leg_cfg3), ("cfg4", leg_cfg4)):
                try:
                    legs[name] = fn(L, synth, torch, dist, device, rank, world, timer)
                except Exception as e:  # noqa: BLE001  (an extra leg must never cost the headline line)
                    legs[name] = {"error": f"{type(e).__name__}: {e}"}
                torch.cuda.empty_cache()
    if not args.no_train:
        try:
            legs["cfg5"] = leg_cfg5(L, synth, torch, dist, device, rank, world, timer)
        except Exception as e:  # noqa: BLE001
            legs["cfg5"] = {"error": f"{type(e).__name__}: {e}"}
        torch.cuda.empty_cache()

    # max over ranks
    t = torch.tensor([total_ms, enc_ms, dec_ms, conv_ms_per_step, e2e["ms_per_step"], engine_ms_per_step, first_ms,
                      last_ms, eb_ms_step, fp32_ms, e2e["copy_ms_per_step"]], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    (total_ms, enc_ms, dec_ms, conv_ms_per_step, e2e_ms, engine_ms_per_step, first_ms, last_ms, eb_ms_step,
     fp32_ms, copy_ms) = t.tolist()

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
        cpu = lib_base = None
        if world == 1 and not args.no_legs:
            sd = {k: v.cpu() for k, v in net.state_dict().items()}
            try:
                lib_base = gpu_library_baseline(torch, device, sd, tiles_host, train=not args.no_train)
            except Exception as e:  # noqa: BLE001
                lib_base = {"error": f"{type(e).__name__}: {e}"}
            torch.cuda.empty_cache()
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
                       "tiles_per_gpu": B, "tile_format": "uint8 pixels in, uint8 x_hat + int16 symbols out (same results as the "
                                                          "fp32 formats, which `fp32_io` times)",
                       "l2": "inputs larger than L2 (50 MB of tiles, >1 GB of activations per step)",
                       "weights": "random init seed 42 + synth.condition_weights"},
            "encode_mpix_s": pix_per_step / (enc_ms / args.steps * 1e-3) / 1e6,
            "decode_mpix_s": pix_per_step / (dec_ms / args.steps * 1e-3) / 1e6,
            "fp32_io": {"ms_per_step": fp32_ms, "mpix_s": pix_per_step / (fp32_ms * 1e-3) / 1e6,
                        "what": "same step with fp32 tiles in, fp32 x_hat + int32 symbols out (the CompressAI tensor formats)"},
            "clocks": clocks,
            "e2e": {"value": pix_per_step / (e2e_ms * 1e-3) / 1e6, "unit": "MPix/s",
                    "h2d_bytes_per_step": e2e["h2d"], "d2h_bytes_per_step": e2e["d2h"],
                    "path": "pinned host uint8 tiles -> model.forward_tiles (g_a / entropy_bottleneck / g_s) -> uint8 x_hat, "
                            "int16 symbols, bpp on host",
                    "rank0_numa_node": numa_node, "bpp": e2e.get("bpp"),
                    "copy_floor": {"ms_per_step": copy_ms, "mpix_s": pix_per_step / (copy_ms * 1e-3) / 1e6,
                                   "host_gb_s_all_ranks": world * (e2e["h2d"] + e2e["d2h"]) / (copy_ms * 1e-3) / 1e9,
                                   "what": "the same host<->device copies (same streams, same chunks, all ranks at once) with no "
                                           "kernel in between: the rate the box's host memory / PCIe path allows"}},
            "gpu_launches": n_launch,
            "roofline": {"bound": "tensor",
                         "kernel": "conv_igemm_pair_kernel (6 launches per step: g_a[2,4,6], g_s[0,2,4], GDN/IGDN fused; CTA pairs, wide slabs)",
                         "achieved": achieved_tflops, "peak": pk["tflops"], "unit": "TFLOP/s",
                         "frac": achieved_tflops / pk["tflops"],
                         "frac_sustained": achieved_tflops / pk["tflops_sustained"], "traffic": traffic,
                         "peak_source": pk["source"],
                         "algorithmic": "2 x (conv + gamma-GEMM MACs) of the 6 layers = 9.89 GFLOP per tile, summed over "
                                        "the 6 launches / their summed CUDA-event durations (an instrumented pass)",
                         "per_launch_ms": {k: round(statistics.mean(v), 4) for k, v in per_layer.items()},
                         "share_of_step": engine_ms_per_step / ms_per_step,
                         "conv_share_of_step": conv_ms_per_step / ms_per_step},
            "roofline_hbm": [
                hbm("conv_first2_kernel (g_a[0] + GDN)", BYTES_PER_TILE_FIRST, first_ms, "uint8 NCHW tiles in, bf16 NHWC out"),
                hbm("deconv_narrow2_kernel (g_s[6])", BYTES_PER_TILE_LAST, last_ms, "bf16 NHWC in, uint8 NCHW x_hat out"),
                hbm("eb_eval_tile_kernel (quantise + likelihoods + int16 symbols + NHWC copy + rate term)", BYTES_PER_TILE_EB,
                    eb_ms_step, "16 B per latent element, one launch (12 B for forward() alone)"),
            ],
        }
        line.update(legs)
        if lib_base is not None:
            line["gpu_library_baseline"] = lib_base
        if cpu is not None:
            line["cpu_baseline"] = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def _max_over_ranks(torch, dist, world, device, *vals):
    t = torch.tensor(list(vals), dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def leg_cfg3(L, synth, torch, dist, device, rank, world, timer):
    """BASELINE.json configs[2]: 512 single-band 12-bit 512x512 tiles (raw split), STRONG scaling: rank r codes
    shard_range(512, r, world) of them, in chunks of 64 (= the pixel count of the headline step)."""
    from licos_b200.sharding import shard_range

    total, chunk = 512, 64
    lo, hi = shard_range(total, rank, world)
    torch.manual_seed(42)
    net = L.get_model(MODEL, False, 1, QUALITY)
    synth.condition_weights(net)
    net = net.to(device).eval()
    g = torch.Generator().manual_seed(4242)
    dn = torch.randint(0, 4096, (chunk, 1, 512, 512), generator=g, dtype=torch.int16).to(device)  # one chunk, re-used
    rate = torch.zeros(1, dtype=torch.float64, device=device)

    def step():
        for c0 in range(lo, hi, chunk):
            n = min(chunk, hi - c0)
            net.forward_tiles(dn[:n], int_max=4095, out_dtype=torch.uint16, sum_ln=rate)

    ms = timer.run(step, steps=10, warmup=3)
    (ms,) = _max_over_ranks(torch, dist, world, device, ms)
    pix = total * 512 * 512
    return {"workload": "BASELINE.json configs[2]: 512 x 1x512x512 12-bit tiles (int16 DNs in, uint16 x_hat + int16 symbols "
                        "out), sharded over the ranks", "scaling": "strong", "tiles_this_rank": hi - lo, "ms_per_step": ms,
            "mpix_s": pix / (ms * 1e-3) / 1e6, "tflops": total * FLOP_PER_TILE_CFG3 / (ms * 1e-3) / 1e12}


def leg_cfg4(L, synth, torch, dist, device, rank, world, timer):
    """BASELINE.json configs[3]: bmshj2018-hyperprior q6 (N=192, M=320) forward on 3x1024x1024 crops, one crop per GPU."""
    torch.manual_seed(42)
    net = L.image_models["bmshj2018-hyperprior"](quality=6, pretrained=False)
    synth.condition_weights(net)
    net = net.to(device).eval()
    x = synth.make_input("rgb1024", 1, seed=77 + rank, device=device)
    with torch.no_grad():
        eager_ms = timer.run(lambda: net(x), steps=10, warmup=3)
        ref = net(x)
    fwd = L.GraphedForward(net, x)  # one crop per call: ~40 launches cost more on the host than on the GPU
    ms = timer.run(lambda: fwd(x), steps=10, warmup=3)
    out = fwd(x)
    same = bool(torch.equal(out["x_hat"], ref["x_hat"]) and torch.equal(out["likelihoods"]["y"], ref["likelihoods"]["y"]))
    ms, eager_ms = _max_over_ranks(torch, dist, world, device, ms, eager_ms)
    pix = world * 1024 * 1024
    return {"workload": "BASELINE.json configs[3]: bmshj2018-hyperprior q6 forward (g_a, h_a, both entropy models, h_s, g_s) on "
                        "one 3x1024x1024 crop per GPU", "scaling": "weak", "how": "licos_b200.GraphedForward (CUDA graph replay)",
            "ms_per_step": ms, "eager_ms_per_step": eager_ms, "replay_equals_eager": same,
            "mpix_s": pix / (ms * 1e-3) / 1e6, "tflops": world * FLOP_PER_IMG_CFG4 / (ms * 1e-3) / 1e12}


def leg_cfg5(L, synth, torch, dist, device, rank, world, timer, steps: int = 20):
    """BASELINE.json configs[4]: train.py's train_one_batch body (forward in train mode, RD loss, backward, clip, Adam, aux
    step) on 32 synthetic 3x256x256 tiles per rank, every kernel native, replayed as one CUDA graph, THEN the weight merge
    over NCCL (licos_nccl_weighted_allreduce, weights 1 / loss from the device scalar) -- both inside the timed region."""
    from licos_b200.federated import FlatState, federated_average

    torch.manual_seed(100)
    net = L.image_models[MODEL](quality=QUALITY, pretrained=False).to(device).train()
    state = FlatState(net)  # before the optimizers and the graph see the parameters' storage
    crit = L.RateDistortionLoss(lmbda=1e-2)
    opt = L.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
    tiles = 32
    x = synth.make_input("rgb256", tiles, seed=1 + rank, device=device)
    graphed = L.GraphedTrainStep(net, crit, opt, x, clip_max_norm=1.0)
    out = graphed(x)
    loss0 = float(out["loss"])

    def train_only():
        return graphed(x)

    def train_and_merge():
        o = graphed(x)
        federated_average(state, o["loss"])
        return o

    step_ms = timer.run(train_only, steps=steps, warmup=3)
    res = {"workload": f"{MODEL} q{QUALITY} training step (forward + RD loss + backward + clip + Adam + aux step) on "
                       f"{tiles} synthetic 3x256x256 tiles per rank (BASELINE.json configs[4])",
           "how": "licos_b200.GraphedTrainStep (CUDA graph replay)", "train_ms_per_step": step_ms, "loss_first": loss0}
    if world > 1:
        both_ms = timer.run(train_and_merge, steps=steps, warmup=3)
        merge_only = timer.run(lambda: federated_average(state, out["loss"]), steps=steps, warmup=2)
        probe = state.flat[:: max(1, state.numel // 4096)].double().sum().reshape(1)
        gathered = [torch.zeros_like(probe) for _ in range(world)]
        dist.all_gather(gathered, probe)
        identical = all(bool(torch.equal(gathered[0], g)) for g in gathered)
        step_ms, both_ms, merge_only = _max_over_ranks(torch, dist, world, device, step_ms, both_ms, merge_only)
        res.update({"train_ms_per_step": step_ms, "ms_per_step": both_ms, "merge_ms": merge_only,
                    "merge": f"one ncclAllReduce (PreMulSum, device scalar 1 / loss) over {state.numel + 1} fp32 = "
                             f"{(state.numel + 1) * 4 / 1e6:.1f} MB + prep / normalise kernels",
                    "replicas_identical_after_merge": identical})
    else:
        res["ms_per_step"] = step_ms
        res["merge"] = "single rank: no exchange step (the merge is timed at N > 1)"
    pix = world * tiles * TILE[1] * TILE[2]
    res["mpix_s"] = pix / (res["ms_per_step"] * 1e-3) / 1e6
    res["tflops"] = world * tiles * 3 * (5.528e9 * 2) / (res["ms_per_step"] * 1e-3) / 1e12  # fwd + dgrad + wgrad
    res["loss_last"] = float(graphed(x)["loss"])
    return res


def gpu_library_baseline(torch, device, state_dict, tiles_host, train: bool):
    """What a LICOS user gets on this GPU today (SURVEY.md section 8d "the Blackwell kernel to beat"): the same modules as
    eager PyTorch on cuDNN / cuBLAS -- here the oracle restatement moved to the GPU, channels_last, cuDNN autotune, bf16
    autocast and TF32 -- timed in this run, on the same box.  Inference on the headline batch; training on cfg 5's batch."""
    from oracle import compressai_ref as R  # baseline leg (the checker's modules, run by the vendor libraries)

    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cuda.matmul.allow_tf32 = True
    ref = R.image_models[MODEL](quality=QUALITY)
    ref.load_state_dict(state_dict)
    ref = ref.to(device).to(memory_format=torch.channels_last).eval()
    x = (tiles_host.double() / 255).float().to(device).contiguous(memory_format=torch.channels_last)
    med = ref.entropy_bottleneck.quantiles[:, 0, 1].detach().reshape(1, -1, 1, 1)
    pix = x.shape[0] * TILE[1] * TILE[2]

    def infer():
        y = ref.g_a(x)
        y_hat, lik = ref.entropy_bottleneck(y.float())
        sym = ref.entropy_bottleneck.quantize(y.float(), "symbols", med)
        return ref.g_s(y_hat), lik, sym

    def timed(fn, n, warm):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    out = {"what": "oracle modules on cuda (cuDNN / cuBLAS eager), channels_last, cudnn.benchmark, same weights and tiles"}
    with torch.no_grad():
        ms_tf32 = timed(infer, 5, 3)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ms_bf16 = timed(infer, 5, 3)
    out["inference"] = {"tf32_ms_per_step": ms_tf32, "bf16_autocast_ms_per_step": ms_bf16,
                        "best_mpix_s": pix / (min(ms_tf32, ms_bf16) * 1e-3) / 1e6}
    if train:
        del x
        torch.manual_seed(100)
        net = R.image_models[MODEL](quality=QUALITY).to(device).to(memory_format=torch.channels_last).train()
        crit = R.RateDistortionLoss(lmbda=1e-2)
        opt = R.net_aux_optimizer(net, {"net": {"type": "Adam", "lr": 1e-4}, "aux": {"type": "Adam", "lr": 1e-3}})
        synth = load_synth()
        xb = synth.make_input("rgb256", 32, seed=1, device=device).contiguous(memory_format=torch.channels_last)

        def train_step(autocast):
            opt["net"].zero_grad(); opt["aux"].zero_grad()
            with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
                o = net(xb)
            loss = crit({"x_hat": o["x_hat"].float(), "likelihoods": {k: v.float() for k, v in o["likelihoods"].items()}}, xb)["loss"]
            loss.backward()
            torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)
            opt["net"].step()
            aux = net.aux_loss()
            aux.backward()
            opt["aux"].step()

        ms_t = timed(lambda: train_step(False), 5, 3)
        ms_b = timed(lambda: train_step(True), 5, 3)
        out["training"] = {"tf32_ms_per_step": ms_t, "bf16_autocast_ms_per_step": ms_b,
                           "what": "train.py:186-200 on 32 tiles through torch autograd (cuDNN dgrad / wgrad)"}
    return out


def run_e2e(net, tiles_host, args, torch, device, world):
    """Same step through the public API with host buffers: H2D of the uint8 tiles, D2H of the uint8 x_hat, the int16 symbols
    and the bpp scalar."""
    import torch.distributed as dist

    B = tiles_host.shape[0]
    eb = net.entropy_bottleneck
    x_host = tiles_host.pin_memory()
    ysz = (B, eb.channels, TILE[1] // 16, TILE[2] // 16)
    xhat_host = torch.empty(x_host.shape, dtype=torch.uint8).pin_memory()
    sym_host = torch.empty(ysz, dtype=torch.int16).pin_memory()
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
                    vb = x_host[lo:hi].to(device, non_blocking=True)
                    out = net.forward_tiles(vb, out_dtype=torch.uint8, sum_ln=accs[it, si:si + 1])
                    xhat_host[lo:hi].copy_(out["x_hat"], non_blocking=True)
                    sym_host[lo:hi].copy_(out["symbols"], non_blocking=True)
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

    # The same bytes over the same streams and chunks WITHOUT any kernel: what this box's host memory / PCIe path delivers to
    # all ranks at once.  When e2e sits on this floor (it does at 8 ranks), the bound is the host side, not the GPU.
    x_dev = torch.empty((chunk,) + tuple(x_host.shape[1:]), dtype=torch.uint8, device=device)
    sym_dev = torch.empty((chunk,) + ysz[1:], dtype=torch.int16, device=device)

    def run_copies(n_steps):
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
                    vb = x_host[lo:hi].to(device, non_blocking=True)
                    xhat_host[lo:hi].copy_(vb, non_blocking=True)
                    sym_host[lo:hi].copy_(sym_dev[:hi - lo], non_blocking=True)
        for s in streams:
            main.wait_stream(s)

    bpp0 = float(bpp_host[0].item())
    run_copies(2)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run_copies(steps)
    torch.cuda.synchronize()
    dt_copy = (time.perf_counter() - t0) / steps
    del x_dev
    return {"ms_per_step": dt * 1e3, "h2d": x_host.numel(), "d2h": xhat_host.numel() + sym_host.numel() * 2 + 8,
            "bpp": bpp0, "copy_ms_per_step": dt_copy * 1e3}


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
