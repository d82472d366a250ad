#!/bin/bash
# Sweep of the paste kernel's tuning knobs (zero-source size, chunk rotation) on the bench workload.
for kb in 16 32 64; do for rot in 0 1 7 13; do
  UWCV_ZERO_KB=$kb UWCV_PASTE_ROT=$rot python bench.py --steps 5 --warmup 3 --no-cpu-baseline --images 32 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('zero_kb $kb rot $rot paste_ms', round(d['kernel_ms']['paste_measure'],4), 'GB/s', round(d['roofline']['achieved'],1))"
done; done
