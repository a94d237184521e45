// Probe (B200): which shared-memory / TMEM operand layouts tcgen05.mma accepts for the attention kernels.
//   0: A K-major SW128,  B K-major SW128            (harness sanity: the layout the fused GEMM uses)
//   1: A K-major SW128,  B MN-major SW128 (N = 64)  (P·V: V stored [key rows][d contiguous])
//   2: A MN-major SW128 (M = 128 = two 64-wide column blocks, LBO = block stride), B K-major
//   3: A MN-major, LBO / SBO swapped                 (in case the descriptor fields mean the opposite)
//   4: A from TMEM (bf16 pairs packed along K in 32-bit columns, tcgen05.st 32x32b), B K-major
//   5: A from TMEM, B MN-major
// Each variant computes D[128, 64] = A[128, K] * B[64, K]^T with K = 48 (three K = 16 steps) and prints the max error.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fairfedmed_b200/csrc -o build/umma_layout_probe tools/micro/umma_layout_probe.cu
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "ffm_common.cuh"
using namespace ffm;

constexpr int M = 128, N = 64, KK = 48;

__device__ __forceinline__ uint32_t sw128(uint32_t off) { return off ^ (((off >> 7) & 7u) << 4); }

__device__ __forceinline__ uint64_t desc_generic(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((addr & 0x3FFFFu) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(layout) << 61;
  return d;
}

__device__ __forceinline__ void umma_bf16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void __launch_bounds__(128, 1) probe(const __nv_bfloat16* A, const __nv_bfloat16* B, float* out, int variant) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;              // 32 KB region for A
  uint8_t* sb = smem + 32768;      // 16 KB region for B
  __shared__ uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
  if (warp == 0) { tmem_alloc(&tslot, 128); tmem_relinquish(); }
  for (int i = tid; i < 49152 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tbase = tslot;
  const bool a_mn = (variant == 2 || variant == 3);
  const bool b_mn = (variant == 1 || variant == 5);
  const bool a_tm = (variant == 4 || variant == 5);
  // ---- A ----
  if (!a_tm) {
    for (int i = tid; i < M * KK; i += 128) {
      const int m = i / KK, k = i % KK;
      uint32_t off;
      if (!a_mn) off = m * 128 + k * 2;                                   // K-major: row m = 128 B (64 k), one k-block
      else off = (m / 64) * (KK * 128) + k * 128 + (m % 64) * 2;          // MN-major: block (m/64), row k, column m%64
      *reinterpret_cast<__nv_bfloat16*>(sa + sw128(off)) = A[i];
    }
  }
  // ---- B ----
  for (int i = tid; i < N * KK; i += 128) {
    const int n = i / KK, k = i % KK;
    const uint32_t off = b_mn ? (k * 128 + n * 2) : (n * 128 + k * 2);
    *reinterpret_cast<__nv_bfloat16*>(sb + sw128(off)) = B[i];
  }
  if (a_tm) {
    // thread = row (lane of TMEM); columns 64.. hold A packed: column 64 + j = (A[row][2j], A[row][2j+1])
    const int row = tid;
    for (int c = 0; c < KK / 16; ++c) {
      uint32_t v[8];
      for (int j = 0; j < 8; ++j) {
        const __nv_bfloat16 lo = A[row * KK + c * 16 + 2 * j], hi = A[row * KK + c * 16 + 2 * j + 1];
        v[j] = static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&lo)) |
               (static_cast<uint32_t>(*reinterpret_cast<const uint16_t*>(&hi)) << 16);
      }
      tmem_st8(tbase + ((warp * 32u) << 16) + 64 + c * 8, v);
    }
    tmem_st_wait();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    uint32_t idesc = umma_idesc_bf16(M, N);
    if (a_mn) idesc |= 1u << 15;
    if (b_mn) idesc |= 1u << 16;
    for (int s = 0; s < KK / 16; ++s) {
      uint64_t bd;
      if (b_mn) bd = desc_generic(smem_u32(sb) + s * 2048, 16, 1024, 2);   // 16 key rows per K step
      else bd = desc_generic(smem_u32(sb) + s * 32, 16, 1024, 2);
      if (a_tm) {
        umma_bf16_ts(tbase, tbase + 64 + s * 8, bd, idesc, s != 0);
      } else {
        uint64_t ad;
        if (!a_mn) ad = desc_generic(smem_u32(sa) + s * 32, 16, 1024, 2);
        else if (variant == 2) ad = desc_generic(smem_u32(sa) + s * 2048, KK * 128, 1024, 2);
        else ad = desc_generic(smem_u32(sa) + s * 2048, 1024, KK * 128, 2);
        umma_bf16(tbase, ad, bd, idesc, s != 0);
      }
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c = 0; c < N / 16; ++c) {
    uint32_t v[16];
    tmem_ld16(tbase + ((warp * 32u) << 16) + c * 16, v);
    tmem_ld_wait();
    for (int j = 0; j < 16; ++j) out[(warp * 32 + lane) * N + c * 16 + j] = __uint_as_float(v[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tbase, 128);
}

int main() {
  std::vector<__nv_bfloat16> hA(M * KK), hB(N * KK);
  std::vector<float> fA(M * KK), fB(N * KK), ref(M * N), got(M * N);
  srand(1);
  for (int i = 0; i < M * KK; ++i) { hA[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fA[i] = __bfloat162float(hA[i]); }
  for (int i = 0; i < N * KK; ++i) { hB[i] = __float2bfloat16((rand() % 2001 - 1000) / 1000.f); fB[i] = __bfloat162float(hB[i]); }
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0;
      for (int k = 0; k < KK; ++k) s += fA[m * KK + k] * fB[n * KK + k];
      ref[m * N + n] = s;
    }
  __nv_bfloat16 *dA, *dB;
  float* dO;
  cudaMalloc(&dA, hA.size() * 2); cudaMalloc(&dB, hB.size() * 2); cudaMalloc(&dO, ref.size() * 4);
  cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice);
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 51200);
  const char* names[] = {"A K-major / B K-major", "A K-major / B MN-major", "A MN-major (LBO=block, SBO=1024) / B K-major",
                         "A MN-major (LBO=1024, SBO=block) / B K-major", "A TMEM / B K-major", "A TMEM / B MN-major"};
  for (int v = 0; v < 6; ++v) {
    cudaMemset(dO, 0, ref.size() * 4);
    probe<<<1, 128, 51200>>>(dA, dB, dO, v);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("variant %d (%s): CUDA error %s\n", v, names[v], cudaGetErrorString(e)); return 1; }
    cudaMemcpy(got.data(), dO, ref.size() * 4, cudaMemcpyDeviceToHost);
    float err = 0;
    for (int i = 0; i < M * N; ++i) err = fmaxf(err, fabsf(got[i] - ref[i]));
    printf("variant %d (%s): max abs err %.5f  %s\n", v, names[v], err, err < 1e-3f ? "OK" : "MISMATCH");
  }
  return 0;
}
