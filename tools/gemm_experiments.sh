#!/bin/bash
# Bottleneck experiments for the fused SVLoRA GEMM (run on the B200 box): which of TMA loads / MMA issue / epilogue bounds it.
# FFM_GEMM_DBG bits: 1 no MMA, 2 no TMA loads, 4 no TMA stores, 128 no proxy fence, 256 no staging stores, 512 no staging-buffer
# reuse wait (64: epilogue phase print-out of the pair build when compiled with -DFFM_GEMM_PAIR_PROF)
out=gpurun_out/gemm_experiments.txt
: > $out
for pair in ${PAIRS:-0}; do
  for dbg in ${DBGS:-0 1 2 3 8 16 18 24 26 17}; do
    for shape in "--K 768 --N 3072" "--K 3072 --N 768"; do
      echo -n "pair=$pair dbg=$dbg $shape " >> $out
      FFM_GEMM_PAIR=$pair FFM_GEMM_DBG=$dbg timeout 120 python tools/gemm_bench.py $shape --iters 20 2>&1 | grep "kernel only" >> $out
    done
  done
done
cat $out
