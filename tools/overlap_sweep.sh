#!/bin/bash
# Step time vs chunk count, with and without the paste / trace overlap
for ov in 1 0; do for ch in 1 2 4 8; do
  if [ $ov = 0 ]; then export UWCV_NO_OVERLAP=1; else unset UWCV_NO_OVERLAP; fi
  UWCV_BENCH_CHUNKS=$ch python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('overlap $ov chunks $ch ms_per_step', round(d['ms_per_step'],3))"
done; done
