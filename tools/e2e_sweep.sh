#!/bin/bash
# e2e (host-buffer) leg of bench.py at several chunk sizes: python tools/e2e_sweep.sh -> gpurun_out/e2e_sweep.log
for c in 32 64 128 256; do
  python bench.py --steps 20 --warmup 5 --no-legs --no-train --no-cpu-baseline --e2e-chunk $c 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
print('chunk $c: value %.0f MPix/s (%.3f ms)  e2e %.0f MPix/s  ratio %.3f  first %.3f ms' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['value'] / d['value'], d['roofline_hbm'][0]['ms_per_launch']))"
done
