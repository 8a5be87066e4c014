#!/bin/bash
# Round-2 evidence: ncu --set full of every kernel family + the launch list of a plain bench loop.
# usage (under gpurun, 1 GPU): tools/run_profiles.sh [names...]
mkdir -p gpurun_out
names=${@:-c3 c2 c4 k50 mask mellinger}
for n in $names; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"twisted|masked" -s 3 -c 1 \
    -o gpurun_out/r02_ncu_$n -f python tools/ncu_target.py $n > gpurun_out/ncu_$n.log 2>&1
  tail -1 gpurun_out/ncu_$n.log
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
