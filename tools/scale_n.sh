#!/bin/bash
# bench line + reference arm at N GPUs, then the host->device ceiling of the box with all N ranks copying
N=${1:-8}
bash tools/bench_n.sh $N r02
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29544 \
  tools/h2d_scaling.py default 2>/dev/null | grep "H2D" | sort > gpurun_out/r02_h2d_n$N.txt
cat gpurun_out/r02_h2d_n$N.txt | cut -c1-200
