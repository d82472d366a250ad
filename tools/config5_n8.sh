#!/bin/bash
# BASELINE configs[4] on N ranks (default 8): deterministic run whose classes.csv / table must equal the
# 1-rank run over the same N * 8 images, then the timed (autotuned) run.  Outputs under gpurun_out/.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
$TR tools/config5_e2e.py --images 8 --deterministic --out gpurun_out/c5_det_n$N.csv 2> gpurun_out/c5_det_n$N.err | tail -1 > gpurun_out/r2_config5_det_n$N.json
python tools/config5_e2e.py --images $((8 * N)) --deterministic --out gpurun_out/c5_det_n1.csv 2> gpurun_out/c5_det_n1.err | tail -1 > gpurun_out/r2_config5_det_n1_same_images.json
if cmp gpurun_out/c5_det_n$N.csv gpurun_out/c5_det_n1.csv; then echo "classes.csv of $N ranks == 1 rank: SAME" > gpurun_out/r2_config5_equal.txt; else echo "classes.csv DIFFER" > gpurun_out/r2_config5_equal.txt; fi
$TR tools/config5_e2e.py --images 8 --out gpurun_out/c5_n$N.csv 2> gpurun_out/c5_n$N.err | tail -1 > gpurun_out/r2_config5_n$N.json
cat gpurun_out/r2_config5_equal.txt; cut -c1-700 gpurun_out/r2_config5_det_n$N.json gpurun_out/r2_config5_det_n1_same_images.json gpurun_out/r2_config5_n$N.json
