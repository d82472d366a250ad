#!/bin/bash
# Elimination experiment on the paste kernel (profiling knobs): which part keeps it below the fill ceiling?
for skip in 0 1 2 3; do
  UWCV_DEBUG_SKIP=$skip python bench.py --steps 5 --warmup 3 --no-cpu-baseline --images 32 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('skip $skip (1=tile compute, 2=band zero stores) paste_ms', round(d['kernel_ms']['paste_measure'],4), 'GB/s', round(d['roofline']['achieved'],1))"
done
