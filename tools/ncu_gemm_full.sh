#!/bin/bash
# ncu --set full captures of the dominant kernels (one launch each), plus the launch list of one bench step.
set -x
ncu --set full --clock-control none --import-source on -k regex:svlora_gemm_kernel -s 2 -c 1 -o gpurun_out/r01_gemm_single_cfc python tools/gemm_bench.py --K 768 --N 3072 --act 1 --iters 2 > /dev/null 2>&1
FFM_GEMM_PAIR=1 ncu --set full --clock-control none --import-source on -k regex:svlora_gemm_pair -s 2 -c 1 -o gpurun_out/r01_gemm_pair_cproj python tools/gemm_bench.py --K 3072 --N 768 --iters 2 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:adapter_grad_kernel -s 2 -c 1 -o gpurun_out/r01_adapter_grad python tools/gemm_bench.py --K 768 --N 3072 --bwd --iters 2 > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 900 --csv --log-file gpurun_out/r01_launches_bench_step.csv python bench.py --steps 2 --warmup 3 --no-graph --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r01_launches_bench_step.csv
