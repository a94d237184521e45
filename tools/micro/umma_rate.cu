// Micro-benchmark (B200): issue rate of tcgen05.mma (M = 128, K = 16, bf16) for the operand forms the attention kernels use,
// as a function of N.  One CTA, one issuing thread, `iters` back-to-back MMAs into the same accumulator, one commit.
//   forms: 0 = A smem K-major, B smem K-major     1 = A smem K-major, B smem MN-major
//          2 = A TMEM,         B smem MN-major     3 = A smem MN-major, B smem MN-major     4 = A TMEM, B smem K-major
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fairfedmed_b200/csrc -o build/umma_rate tools/micro/umma_rate.cu
#include <cstdio>
#include "ffm_common.cuh"
using namespace ffm;

template <int ROT>
__global__ void __launch_bounds__(128, 1) rate(int form, int n, int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 512); tmem_relinquish(); }
  for (int i = tid; i < 98304 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tslot;
  if (warp == 1) {
    // whole warp runs the loop, one elected lane issues (the form the kernels use: descriptors stay in uniform registers)
    uint32_t idesc = umma_idesc_bf16(128, n);
    const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
    uint64_t ad, bd;
    if (form == 3) { ad = umma_desc_sw128_mn(sa, 16384); idesc |= UMMA_IDESC_A_MN; } else ad = umma_desc_sw128(sa);
    if (form == 1 || form == 2 || form == 3) { bd = umma_desc_sw128_mn(sb, 16384); idesc |= UMMA_IDESC_B_MN; } else bd = umma_desc_sw128(sb);
    const bool ts = form == 2 || form == 4;
    const uint32_t kb = (form == 0 || form == 4) ? 2u : 0u;
    long long t0 = clock64();
    for (int i = 0; i < iters; i += 4) {
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const uint32_t dcol = tb + static_cast<uint32_t>((k % ROT) * n);      // ROT independent accumulators
          if (ts) umma_bf16_ts(dcol, tb + 384 + 8 * k, bd, idesc, 1u);
          else umma_bf16(dcol, ad + 2u * k, bd + kb * k, idesc, 1u);
        }
      }
      __syncwarp();
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    long long t1 = clock64();
    mbar_wait_uniform(&bar, 0);
    long long t2 = clock64();
    if ((tid & 31) == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 512);
}

int main() {
  long long* d;
  cudaMalloc(&d, 16);
  cudaFuncSetAttribute(rate<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100352);
  cudaFuncSetAttribute(rate<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100352);
  cudaFuncSetAttribute(rate<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100352);
  const char* names[] = {"SS  A K-major  B K-major ", "SS  A K-major  B MN-major", "TS  A TMEM     B MN-major", "SS  A MN-major B MN-major",
                         "TS  A TMEM     B K-major "};
  const int ns[] = {16, 64, 128, 208, 256};
  const int iters = 2000;
  for (int rot : {1, 2, 4})
  for (int f = 0; f < 5; ++f)
    for (int n : ns) {
      if (rot * n > 384 || (rot > 1 && (f == 1 || f == 3 || f == 4))) continue;
      long long h[2];
      for (int rep = 0; rep < 2; ++rep) {
        if (rot == 1) rate<1><<<1, 128, 100352>>>(f, n, iters, d);
        else if (rot == 2) rate<2><<<1, 128, 100352>>>(f, n, iters, d);
        else rate<4><<<1, 128, 100352>>>(f, n, iters, d);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("form %d n %d: CUDA error\n", f, n); return 1; }
        cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
      }
      printf("rot %d %s N=%3d: issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (ideal %5.1f)\n", rot, names[f], n, (double)h[0] / iters,
             (double)h[1] / iters, 128.0 * n / 256.0);
    }
  return 0;
}
