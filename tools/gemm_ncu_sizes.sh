#!/bin/bash
# fixed vs variable cost of the fused GEMM: ncu durations for several T
M=gpu__time_duration.sum,gpc__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed
for T in 128 1184 6304 12608 25216 50432; do
  for shape in "--K 768 --N 3072" "--K 3072 --N 768"; do
    echo "== T=$T $shape dbg=${FFM_GEMM_DBG:-0}"
    ncu --metrics $M --clock-control none -k regex:svlora_gemm -s 3 -c 1 python tools/gemm_bench.py --T $T $shape --iters 2 2>&1 | grep -E "gpu__time|gpc__cycles|pipe_tensor"
  done
done
