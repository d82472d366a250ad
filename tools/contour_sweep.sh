#!/bin/bash
# contour kernel: lanes of a warp in decreasing-work order (counting sort) vs call order; lanes per warp
for uns in 1 0; do for lanes in 0 12 16 32; do
  if [ $uns = 1 ]; then export UWCV_CONTOUR_UNSORTED=1; else unset UWCV_CONTOUR_UNSORTED; fi
  if [ $lanes = 0 ]; then unset UWCV_CONTOUR_LANES; else export UWCV_CONTOUR_LANES=$lanes; fi
  python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('unsorted $uns lanes $lanes', d['kernel_ms'], 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done; done
