#!/bin/bash
# Round-2 evidence on one B200 (outputs under gpurun_out/, copied to profiles/ by hand):
# bench lines (our arm + reference arm), ncu launch list, ncu --set full of the paste kernel,
# the other configs.
set -o pipefail
O=gpurun_out
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r02_bench_n1.json 2> $O/r02_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "reference rc=$?"
timeout 300 python tools/bench_configs.py > $O/r02_configs.json 2> $O/r02_configs.err; echo "configs rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --images 16"
$CMD > $O/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/profile_target.py 16 > $O/plain_paste.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:paste_measure_kernel -s 3 -c 1 -o $O/r02_paste python tools/profile_target.py 16 > $O/ncu_paste.log 2>&1
echo "paste capture rc=$?"
ls -la $O/*.ncu-rep $O/r02_*
