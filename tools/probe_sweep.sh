#!/bin/bash
# Mainloop vs data-movement decomposition of the conv engine: per-layer time and per-tile cycle probes with the ring slots
# loaded only once (LICOS_DBG_FLAGS bit 0: A slabs, bit 1: weights), and ring-depth variants.  Output: stdout
for f in 0 3; do
  echo "#### LICOS_DBG_FLAGS=$f"
  LICOS_DBG_FLAGS=$f LAYERS=${LAYERS:-1,6,8} python tools/probe_conv.py 256 2>&1 | grep -v Warning
done
for cfg in "2 8" "3 4" "4 2" "2 5"; do
  set -- $cfg
  echo "#### LICOS_SA=$1 LICOS_SB=$2"
  LICOS_SA=$1 LICOS_SB=$2 LAYERS=${LAYERS:-1,6,8} python tools/probe_conv.py 256 2>&1 | grep -v Warning
done
