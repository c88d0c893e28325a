#!/bin/bash
# Same-box A/B of two BUILDS of the library (compile-time knobs of the CTA-pair engine's GDN epilogue), three runs each,
# interleaved.  Prepare both here before the gpurun call:
#   (cd licos_b200/csrc && LICOS_NVCC_EXTRA=-DLICOS_NO_EPI_PIPE python -c "import build; build.build(force=True)")   # or -DLICOS_STG_RELEASE
#   cp licos_b200/lib/liblicos_b200.so licos_b200/lib/nopipe.so.bin                                                   # variant B
#   (cd licos_b200/csrc && python -c "import build; build.build(force=True)")                                        # variant A = default
# then: gpurun -- bash tools/epi_pipe_ab.sh      (the script restores variant A at the end)
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
