#!/bin/bash
# Paste kernel: bound on TMA bulk stores in flight per issuer (UWCV_PASTE_ROT = depth, 0 = unbounded) x chunk size
for kb in 16 64; do for depth in 0 1 2 4 8; do
  UWCV_ZERO_KB=$kb UWCV_PASTE_ROT=$depth python bench.py --steps 5 --warmup 3 --no-cpu-baseline --images 32 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('chunk_kb $kb depth $depth paste_ms', round(d['kernel_ms']['paste_measure'],4), 'GB/s', round(d['roofline']['achieved'],1))"
done; done
