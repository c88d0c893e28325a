#!/bin/bash
# A/B of the CTA-pair engine against the single-CTA engine: conv parity tests, then the bench headline both ways.
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -5
for np in 1 0; do
  if [ $np = 1 ]; then export LICOS_NO_PAIR=1; else unset LICOS_NO_PAIR; fi
  timeout 300 python bench.py --steps 20 --warmup 5 --no-legs --no-train --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
r = d['roofline']
print('NO_PAIR=$np: %.3f ms/step %.0f MPix/s  engine %.0f TFLOP/s (frac %.3f)  clocks %s' % (d['ms_per_step'], d['value'], r['achieved'], r['frac'], d['clocks']))
print('   ', r['per_launch_ms'])"
done
