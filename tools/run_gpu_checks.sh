#!/bin/bash
# Runs every GPU test node in its own process (a CUDA fault poisons the context, so one failing kernel must
# not hide the others) and writes a digest to gpurun_out/diag.log.  Usage: tools/run_gpu_checks.sh [pytest -k expr]
mkdir -p gpurun_out
LOG=gpurun_out/diag.log
: > $LOG
nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader >> $LOG 2>&1
python -c "import os; print('host cores', os.cpu_count())" >> $LOG
NODES=$(python -m pytest tests -m gpu --collect-only -q ${1:+-k "$1"} 2>/dev/null | grep "::")
pass=0; fail=0
for n in $NODES; do
  out=$(timeout 300 python -m pytest "$n" -x -q -s -m gpu 2>&1)
  rc=$?
  if [ $rc -eq 0 ]; then pass=$((pass+1)); echo "PASS $n" >> $LOG; echo "$out" | grep -E "max\|err\||PSNR|agreement|smoke|worst cos|^loss " >> $LOG
  else fail=$((fail+1)); echo "FAIL($rc) $n" >> $LOG; echo "$out" | grep -E "^grad " | tail -40 >> $LOG; echo "$out" | tail -25 >> $LOG; fi
done
echo "SUMMARY pass=$pass fail=$fail" >> $LOG
tail -5 $LOG
