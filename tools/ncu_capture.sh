#!/bin/bash
# ncu evidence for profiles/ (B200_PROFILING.md recipe): launch list of a short bench run + one --set full capture of every
# kernel of one inference step.  Run on the GPU box: gpurun -- bash tools/ncu_capture.sh <tag>
TAG=${1:-r2}
CMD="python bench.py --steps 2 --warmup 1 --min-warmup 3 --no-legs --no-train --no-cpu-baseline --no-e2e"
$CMD > gpurun_out/${TAG}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${TAG}_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $CMD > gpurun_out/${TAG}_ncu1.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:"conv_igemm_pair|conv_first2|deconv_narrow2|eb_eval_tile" -s 18 -c 9 \
    -o gpurun_out/${TAG}_prof -f $CMD > gpurun_out/${TAG}_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out/${TAG}_prof.ncu-rep
