"""Turns a `--set full` capture (gpurun_out/<tag>_prof.ncu-rep, tools/ncu_capture.sh) into the tracked summaries under
profiles/: <tag>_ncu_full_summary.csv (one row per captured launch, the metrics the roofline discussion uses) and
roofline_traffic.json (mean DRAM bytes per launch of the conv engine, read by bench.py for `roofline.traffic`).
    python tools/ncu_summarize.py r2            (runs here: ncu -i needs no GPU)"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
rep = os.path.join(ROOT, "gpurun_out", f"{tag}_prof.ncu-rep")
raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
WANT = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__cycles_elapsed.avg", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_uniform.sum", "smsp__inst_executed.sum"]
idx = {h: i for i, h in enumerate(hdr)}
cols = [w for w in WANT if w in idx]
out = os.path.join(ROOT, "profiles", f"{tag}_ncu_full_summary.csv")
with open(out, "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(cols)
    w.writerow([units[idx[c]] for c in cols])
    for r in data:
        w.writerow([r[idx[c]].split("(")[0].strip() if c == "Kernel Name" else r[idx[c]] for c in cols])
print(out)


def gb(r, c):
    v, u = float(r[idx[c]]), units[idx[c]]
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}[u]


eng = [r for r in data if "conv_igemm" in r[idx["Kernel Name"]]]
if eng:
    per = [gb(r, "dram__bytes_read.sum") + gb(r, "dram__bytes_write.sum") for r in eng]
    tj = {"conv_igemm_dram_bytes_per_launch": sum(per) / len(per), "algorithmic_bytes_per_launch": 590003333.3333334,
          "launches": len(per),
          "source": f"profiles/{tag}_ncu_full_summary.csv: dram__bytes_read.sum + dram__bytes_write.sum, mean of the conv engine launches "
                    "of one step (batch 256); algorithmic = every activation read once + written once (bf16 NHWC, fp32 NCHW y)"}
    with open(os.path.join(ROOT, "profiles", "roofline_traffic.json"), "w") as f:
        json.dump(tj, f, indent=1)
    print(tj)
