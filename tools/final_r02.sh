#!/bin/bash
# Last GPU call of round 2 (one B200): the GPU test set, the default bench line (with the library baseline of the
# same GPU: torch grid_sample paste, torchvision NMS) and a randomized parity sweep, all on the committed code.
# Outputs under gpurun_out/ (copied to profiles/r02_final_*).
set -o pipefail
O=gpurun_out
mkdir -p $O
timeout 420 python -m pytest tests -m gpu -q -x > $O/r02_final_gputests.txt 2>&1; echo "gpu tests rc=$?"; tail -3 $O/r02_final_gputests.txt
timeout 480 python bench.py > $O/r02_final_bench_n1_default.json 2> $O/r02_final_bench.err; echo "default bench rc=$?"
tail -c 600 $O/r02_final_bench.err
timeout 240 python tools/parity_sweep.py 150 77000 > $O/r02_final_sweep.json 2> $O/r02_final_sweep.err; echo "sweep rc=$?"
tail -c 400 $O/r02_final_sweep.json
