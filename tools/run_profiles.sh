#!/bin/bash
# Round-2 evidence: ncu --set full of every kernel family + the launch list of a plain bench loop.
# usage (under gpurun, 1 GPU): tools/run_profiles.sh [names...]
mkdir -p gpurun_out
names=${@:-c3 c2 c4 k50 mask mellinger}
for n in $names; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:"twisted|masked" -s 3 -c 1 \
    -o gpurun_out/r02_ncu_$n -f python tools/ncu_target.py $n > gpurun_out/ncu_$n.log 2>&1
  tail -1 gpurun_out/ncu_$n.log
  # the .ncu-rep files are ~16 MB each and gpurun_out/ is capped at 64 MiB: summarise on the box, keep the text
  python tools/ncu_summary.py gpurun_out/r02_ncu_$n.ncu-rep > gpurun_out/r02_ncu_$n.txt 2>&1
  ncu -i gpurun_out/r02_ncu_$n.ncu-rep --page source --csv > gpurun_out/r02_ncu_${n}_source.csv 2>/dev/null
  python tools/ncu_stalls.py gpurun_out/r02_ncu_${n}_source.csv >> gpurun_out/r02_ncu_$n.txt 2>&1
  rm -f gpurun_out/r02_ncu_${n}_source.csv
  [ "$n" = "c3" ] || rm -f gpurun_out/r02_ncu_$n.ncu-rep
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_under_ncu.log 2>&1
tail -c 300 gpurun_out/bench_under_ncu.log
