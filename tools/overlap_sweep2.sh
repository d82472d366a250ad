#!/bin/bash
for ctas in 2 3; do
  UWCV_PASTE_CTAS=$ctas python tools/overlap_probe.py 2>&1 | tail -7
done
