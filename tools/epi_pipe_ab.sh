#!/bin/bash
# same box: pipelined GDN epilogue (A) vs not (B), three runs each, interleaved
L=licos_b200/lib
cp $L/liblicos_b200.so $L/pipe.so.bin
run() {
  timeout 120 python bench.py --steps 40 --warmup 8 --no-legs --no-train --no-cpu-baseline --no-e2e 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); r=d['roofline']; print('$1', round(d['ms_per_step'],4), round(r['achieved']), r['per_launch_ms'])"
}
for i in 1 2 3; do
  cp $L/nopipe.so.bin $L/liblicos_b200.so; touch $L/liblicos_b200.so; run B
  cp $L/pipe.so.bin $L/liblicos_b200.so; touch $L/liblicos_b200.so; run A
done
cp $L/pipe.so.bin $L/liblicos_b200.so
