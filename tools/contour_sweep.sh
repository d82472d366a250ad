#!/bin/bash
# contour kernel time: mixed vs phased loop, lanes per warp
for mixed in 1 0; do for lanes in 0 26 32; do
  UWCV_CONTOUR_MIXED=$mixed UWCV_CONTOUR_LANES=$lanes python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('mixed $mixed lanes $lanes', d['kernel_ms'], 'step', round(d['ms_per_step'],3))"
done; done
