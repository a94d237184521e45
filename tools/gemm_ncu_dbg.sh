#!/bin/bash
# ncu cycle counts / tensor-pipe activity of the fused GEMM for a few FFM_GEMM_DBG experiment masks
M=gpu__time_duration.sum,gpc__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg
for dbg in ${DBGS:-0 2 18 1 3}; do
  for shape in "--K 768 --N 3072" "--K 3072 --N 768"; do
    echo "== dbg=$dbg $shape"
    FFM_GEMM_DBG=$dbg ncu --metrics $M --clock-control none -k regex:svlora_gemm -s 3 -c 1 python tools/gemm_bench.py $shape --iters 2 2>&1 | grep -E "gpu__time|gpc__cycles|pipe_tensor|cuBLAS|kernel only"
  done
done
