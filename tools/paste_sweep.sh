#!/bin/bash
# Paste-kernel sweeps on the TUNING library (the release library reads no environment knobs):
# CTAs per SM, zero-source size, and the elimination runs (UWCV_DEBUG_SKIP 1: no tile compute,
# 2: no band zero stores, 3: both).  tools/step_probe.py prints layout / paste / serial / overlapped ms.
run() { timeout 120 python tools/step_probe.py --variant tuning | cut -c1-400; }
for ctas in 1 2 3; do for kb in 16 32 64; do
  echo "== CTAs/SM $ctas, zero source $kb KB"; UWCV_PASTE_CTAS=$ctas UWCV_ZERO_KB=$kb run
done; done
for skip in 1 2 3; do echo "== UWCV_DEBUG_SKIP=$skip"; UWCV_DEBUG_SKIP=$skip run; done
