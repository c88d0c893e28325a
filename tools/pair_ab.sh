#!/bin/bash
# A/B of the engine variants: conv parity tests, then the bench headline for
#   LICOS_NO_PAIR=1 (single-CTA engine), LICOS_NO_WIDE=1 (CTA pairs, per-column-tap slabs), default (CTA pairs, wide slabs)
timeout 600 python -m pytest tests/test_gpu_conv.py -x -q 2>&1 | tail -3
for v in NO_PAIR NO_WIDE DEFAULT; do
  unset LICOS_NO_PAIR LICOS_NO_WIDE
  [ $v = NO_PAIR ] && export LICOS_NO_PAIR=1
  [ $v = NO_WIDE ] && export LICOS_NO_WIDE=1
  timeout 300 python bench.py --steps 20 --warmup 5 --no-legs --no-train --no-cpu-baseline --no-e2e 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][-1])
r = d['roofline']
print('$v: %.3f ms/step %.0f MPix/s  engine %.0f TFLOP/s (frac %.3f)  clocks %s' % (d['ms_per_step'], d['value'], r['achieved'], r['frac'], d['clocks']))
print('   ', r['per_launch_ms'])"
done
