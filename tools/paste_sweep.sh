#!/bin/bash
# Paste kernel: CTAs per SM (occupancy) vs achieved bandwidth
for ctas in 1 2 3; do
  UWCV_PASTE_CTAS=$ctas python bench.py --steps 5 --warmup 3 --no-cpu-baseline --images 32 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ctas/SM $ctas paste_ms', round(d['kernel_ms']['paste_measure'],4), 'GB/s', round(d['roofline']['achieved'],1))"
done
