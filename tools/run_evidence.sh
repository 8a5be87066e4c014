#!/bin/bash
# Round-end evidence on ONE GPU (under gpurun): test suite, bench lines, variant / sweep tables, parity statistics.
# Outputs land in gpurun_out/ and are copied to profiles/ by hand after review.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r02_pytest_gpu.log
cat gpurun_out/r02_pytest_gpu.log
timeout 600 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 600 gpurun_out/r02_bench_n1.json
timeout 300 python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2>> gpurun_out/r02_bench_n1.err
timeout 300 python tools/k1_variants.py > gpurun_out/r02_k1_variants.log 2>&1
timeout 120 python tools/k_sweep.py > gpurun_out/r02_k_sweep.log 2>&1
timeout 120 python tools/early_refill_sweep.py > gpurun_out/r02_early_refill.log 2>&1
timeout 120 python tools/small_batch_sweep.py > gpurun_out/r02_small_batch_sweep.log 2>&1
timeout 600 python tools/parity_report.py > gpurun_out/r02_parity.log 2>&1
tail -3 gpurun_out/r02_parity.log
