#!/bin/bash
# Mainloop vs data-movement decomposition of the conv engine: per-layer time and per-tile cycle probes with the ring slots
# loaded only once (LICOS_DBG_FLAGS bit 0: A slabs, bit 1: weights).  Output: gpurun_out/probe_sweep.log
for f in 0 1 2 3; do
  echo "#### LICOS_DBG_FLAGS=$f"
  LICOS_DBG_FLAGS=$f LAYERS=1,2,6,8 python tools/probe_conv.py 256 2>&1 | grep -v Warning
done
