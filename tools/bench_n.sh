#!/bin/bash
# bench line at N GPUs under torchrun (as the driver launches it); output under gpurun_out/
N=${1:-8}; TAG=${2:-r2}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 2> gpurun_out/${TAG}_bench_n$N.err | tail -1 > gpurun_out/${TAG}_bench_n$N.json
grep -v "^\*\|OMP_NUM\|^$" gpurun_out/${TAG}_bench_n$N.err | tail -5 | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/${TAG}_bench_n$N.json"))
print({k:d[k] for k in ("value","ms_per_step","scaling","n_gpus")}, d["config"]["images_per_gpu"], d.get("weak_scaling"))
print("e2e", {k:v for k,v in d["e2e"].items() if k!="call"}); print("rows_only", d["rows_only"]["ms_per_step"], "kernels", d["kernel_ms"])
PY
