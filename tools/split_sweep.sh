#!/bin/bash
# sweep of the split pipeline's knobs (tuning build): L2 policy of the plane fill x trace lanes per warp
out=gpurun_out/split_sweep.jsonl
: > $out
for pol in 0 1 2; do for lanes in 22 12 8; do
  UWCV_FILL_POLICY=$pol UWCV_TRACE_LANES=$lanes python tools/step_probe.py --variant tuning --steps 8 >> $out 2>> gpurun_out/split_sweep.err
done; done
python - <<'PY'
import json
for l in open("gpurun_out/split_sweep.jsonl"):
    r = json.loads(l)
    print(r["knobs"], "fill", round(r["plane_fill"][0], 3), "fused_ovl", round(r["overlapped_step"], 3), "split_ovl", round(r["overlapped_split_step"], 3))
PY
