#!/bin/bash
# 2-GPU checks: multi-GPU tests, then the bench line at N = 2 (configs[2]: 128 images per GPU)
N=${1:-2}
timeout 600 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -15
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 2> gpurun_out/r2_bench_n$N.err | tail -1 > gpurun_out/r2_bench_n$N.json
tail -5 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n$N.json"))
print({k:d[k] for k in ("value","ms_per_step","scaling","n_gpus")}, d["config"]["images_per_gpu"], d.get("weak_scaling"))
print("e2e", d["e2e"]); print("rows_only", d["rows_only"]["ms_per_step"], "kernels", d["kernel_ms"])
PY
