#!/bin/bash
# compute-sanitizer over tools/sanitize_target.py (every kernel family of the shipped library).
out=${1:-gpurun_out/sanitizer.txt}
: > "$out"
for tool in memcheck racecheck synccheck; do
  echo "=== compute-sanitizer --tool $tool (library $(sha1sum mav_trajectory_generation_b200/libmtg_b200.so | cut -c1-12)) ===" >> "$out"
  extra=""
  # synccheck (CUDA 12.9) reports "Barrier error detected. Missing init" on tcgen05.alloc (UTCATOMSWS writes the TMEM
  # base address to shared memory; the kernels contain no mbarrier / SYNCS instruction at all) and kills the kernel:
  # it is run on every kernel family except the tcgen05 kernels.
  [ "$tool" = synccheck ] && extra="--no-tmem"
  timeout 900 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_target.py $extra 2>&1 | grep -v "^$" | tail -25 >> "$out"
done
