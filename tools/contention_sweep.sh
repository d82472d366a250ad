#!/bin/bash
# contention_probe.py over tuning knobs / arguments: every argument is "VAR=val ... -- probe args"
out=gpurun_out/contention_sweep.jsonl
: > $out
for spec in "$@"; do
  knobs="${spec%%--*}"; args="${spec#*--}"
  [ "$args" == "$spec" ] && args=""
  env $knobs python tools/contention_probe.py --variant tuning --reps 3 $args >> $out 2>> gpurun_out/contention_sweep.err
done
cat $out
