#!/bin/bash
# paste kernel time vs CTAs/SM and zero-source size (dynamic work claims)
for ctas in 1 2 3; do for kb in 16 32 64; do
  UWCV_PASTE_CTAS=$ctas UWCV_ZERO_KB=$kb python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('ctas $ctas zero_kb $kb', d['kernel_ms'], 'step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],3))"
done; done
for skip in 1 2 3; do
  UWCV_DEBUG_SKIP=$skip python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('debug_skip $skip (1: no tile compute, 2: no band zero stores)', d['kernel_ms'])"
done
