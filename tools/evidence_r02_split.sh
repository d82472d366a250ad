#!/bin/bash
# Round-2 evidence of the split pipeline on one B200 (outputs under gpurun_out/, summaries copied to profiles/):
# bench lines (our arm + reference arm), ncu launch list of the bench command, ncu --set full of the plane-fill
# kernel (16 000 instances) and of the border-trace kernel (64 000 instances; SKIP_TRACE=1 leaves it out), raw pages
# exported as CSV.  TAG names the outputs (default r02_split).
set -o pipefail
O=gpurun_out
timeout 600 python bench.py > $O/${TAG:-r02_split}_bench_n1_default.json 2> $O/def.err; echo "default bench rc=$?"
timeout 600 python bench.py --steps 20 --warmup 5 > $O/${TAG:-r02_split}_bench_n1.json 2> $O/${TAG:-r02_split}_bench_n1.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG:-r02_split}_bench_reference.json 2> $O/${TAG:-r02_split}_bench_reference.err; echo "reference rc=$?"
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-other-configs --images 16"
$CMD > $O/plain_launches.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${TAG:-r02_split}_launches.csv $CMD > $O/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/profile_target.py 16 > $O/plain_fill.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:plane_fill_kernel -s 3 -c 1 -f -o $O/${TAG:-r02_split}_plane_fill python tools/profile_target.py 16 > $O/ncu_fill.log 2>&1
echo "fill capture rc=$?"
ncu -i $O/${TAG:-r02_split}_plane_fill.ncu-rep --page raw --csv > $O/${TAG:-r02_split}_plane_fill_raw.csv 2>/dev/null
if [ -z "$SKIP_TRACE" ]; then
python tools/profile_target.py 64 > $O/plain_trace.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:contour_measure_kernel -s 3 -c 1 -f -o $O/${TAG:-r02_split}_contour python tools/profile_target.py 64 > $O/ncu_trace.log 2>&1
echo "trace capture rc=$?"
ncu -i $O/${TAG:-r02_split}_contour.ncu-rep --page raw --csv > $O/${TAG:-r02_split}_contour_raw.csv 2>/dev/null
fi
ls -la $O/*.ncu-rep $O/${TAG:-r02_split}_* $O/*_raw.csv
