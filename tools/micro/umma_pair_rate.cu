// Micro-benchmark (B200): issue rate of tcgen05.mma.cta_group::2 (M = 256 over a CTA pair, K = 16, bf16, both operands in
// shared memory, each CTA holding its 128 rows of A and half of the N rows of B) against the single-CTA form.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fairfedmed_b200/csrc -o build/umma_pair_rate tools/micro/umma_pair_rate.cu
#include <cstdio>
#include "ffm_common.cuh"
using namespace ffm;

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(128, 1) rate_pair(int n, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const uint32_t rank = cluster_ctarank();
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  cluster_sync_all();
  if (warp == 0) { tmem_alloc_pair(&tslot, 512); tmem_relinquish_pair(); }
  for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (warp == 1 && rank == 0) {
    const uint32_t idesc = umma_idesc_bf16(256, n);
    const uint64_t ad = umma_desc_sw128(smem_u32(smem)), bd = umma_desc_sw128(smem_u32(smem + 32768));
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) umma_bf16_pair(tb, ad + 2u * k, bd + 2u * k, idesc, 1u);
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit_pair(&bar);
    __syncwarp();
    long long t1 = clock64();
    mbar_wait_cluster_uniform(&bar, 0);
    long long t2 = clock64();
    if ((tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  } else if (warp == 1) {
    mbar_wait_cluster_uniform(&bar, 0);     // the multicast commit arrives on both CTAs' barriers
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();
  if (warp == 0) tmem_dealloc_pair(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate_pair, cudaFuncAttributeMaxDynamicSharedMemorySize, 100352);
  const int iters = 2000;
  for (int n : {64, 128, 192, 208, 224, 256}) {
    long long h[2];
    for (int rep = 0; rep < 2; ++rep) {
      rate_pair<<<2, 128, 100352>>>(n, iters, d);
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("n %d: CUDA error %s\n", n, cudaGetErrorString(e)); return 1; }
      cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
    }
    printf("cta_group::2 M=256 N=%3d: %6.1f cyc/MMA (ideal %5.1f per SM: 128 rows x N / 256)\n", n, (double)h[1] / iters,
           128.0 * n / 256.0);
  }
  return 0;
}
