#!/bin/bash
# Fused-trace experiment (tuning libraries; r02, rejected -- see profiles/README.md): the stand-alone trace
# against tracer warps inside the paste kernel (UWCV_FUSED_TRACE=1), 2 / 3 / 4 tracer warps per CTA, with the
# tracer switched off (UWCV_DEBUG_SKIP=4: cost of the larger CTA alone), waiting for arena slots, and not
# waiting (UWCV_DEBUG_SKIP=8: tiles without a free slot go to the stand-alone trace).
run() { timeout 120 "$@" | cut -c1-900 || echo "TIMEOUT/FAIL: $*"; }
echo "== stand-alone trace"; run python tools/step_probe.py --variant tuning
for w in ${WARPS:-2 3 4}; do
  v="tuning,TRACER_WARPS=$w"
  echo "== $w tracer warp(s), tracer off"; UWCV_FUSED_TRACE=1 UWCV_DEBUG_SKIP=4 run python tools/step_probe.py --variant "$v"
  echo "== $w tracer warp(s), wait for slots"; UWCV_FUSED_TRACE=1 run python tools/step_probe.py --variant "$v"
  echo "== $w tracer warp(s), no wait"; UWCV_FUSED_TRACE=1 UWCV_DEBUG_SKIP=8 run python tools/step_probe.py --variant "$v"
done
