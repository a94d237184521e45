// Peak throughput of legacy mma.sync.m16n8k16 (bf16 -> fp32) on sm_100a: decides whether an mma.sync attention kernel can
// compete with the library's tcgen05 one.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_sync_peak mma_sync_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int NACC>
__global__ void __launch_bounds__(256) mma_loop(float* out, int iters) {
  float acc[NACC][4];
#pragma unroll
  for (int i = 0; i < NACC; ++i) acc[i][0] = acc[i][1] = acc[i][2] = acc[i][3] = 0.f;
  unsigned a0 = threadIdx.x * 0x3f803f80u, a1 = a0 ^ 0x1111u, a2 = a0 + 7u, a3 = a1 + 9u, b0 = 0x3f803f80u, b1 = 0x3f003f00u;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                   : "+f"(acc[i][0]), "+f"(acc[i][1]), "+f"(acc[i][2]), "+f"(acc[i][3])
                   : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += acc[i][0] + acc[i][1] + acc[i][2] + acc[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
void run(int ctas_per_sm, int sms) {
  float* out;
  cudaMalloc(&out, sizeof(float) * 256 * sms * ctas_per_sm);
  const int iters = 4096;
  mma_loop<NACC><<<sms * ctas_per_sm, 256>>>(out, 16);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  mma_loop<NACC><<<sms * ctas_per_sm, 256>>>(out, iters);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  const double flop = 2.0 * 16 * 8 * 16 * NACC * (double)iters * 8 /*warps*/ * sms * ctas_per_sm;
  printf("NACC=%d warps/SM=%d: %.3f ms  %.1f TFLOP/s  (%.0f FLOP/clk/SM at 1.965 GHz)\n", NACC, 8 * ctas_per_sm, ms,
         flop / ms / 1e9, flop / (ms * 1e-3) / sms / 1.965e9);
  cudaFree(out);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  run<4>(1, sms); run<8>(1, sms); run<8>(2, sms); run<16>(1, sms); run<8>(4, sms);
  return 0;
}
