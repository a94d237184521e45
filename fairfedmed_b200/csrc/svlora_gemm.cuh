// Internal declarations shared by the single-CTA and CTA-pair builds of the fused SVLoRA GEMM.
#pragma once

#include "ffm_common.cuh"

namespace ffm {

constexpr int RP = 16;              // padded adapter rank
enum : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_QUICKGELU_GRAD = 2 };

struct GemmParams {
  const float* bias;            // [N] or nullptr
  const float* s_rows;          // [n_samples, RP] fp32 (already multiplied by alpha/r)
  float* h_out;                 // [T, RP] fp32 or nullptr
  const __nv_bfloat16* aux;     // ACT_QUICKGELU_GRAD: QuickGELU'(u) saved by the forward [T, N]
  int T, K, N;
  int b_prime, num_slices;      // sample(t) = ((t / row_div) % b_prime) / num_slices
  int row_div;                  // 1: sequence-first rows [L, B', C] (reference); L: batch-first rows [B', L, C]
  int act;
  int has_pre;                  // ACT_QUICKGELU: also store QuickGELU'(u) through tm_y2
  int m_tiles, n_tiles, k_blocks;
};

struct GemmOperands {
  const void* x;        // [T, K] bf16
  const void* wmat;     // [N, K] bf16
  const void* a_side;   // [RP, K] bf16
  const void* b_side;   // [N, RP] bf16
  const float* s_rows;  // [nS, RP]
  const float* bias;    // [N] or null
  void* out;            // [T, N] bf16
  void* out_pre;        // [T, N] bf16 or null (ACT_QUICKGELU only): receives QuickGELU'(u)
  float* h_out;         // [T, RP] or null
  const void* aux;      // [T, N] bf16 (ACT_QUICKGELU_GRAD)
  int T, K, N, b_prime, num_slices, row_div, act;
};

// row-major bf16 matrix [rows, cols] (cols contiguous) -> 2-D tiled map with box [box_rows, box_cols]
int make_map_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                  uint32_t box_cols, CUtensorMapSwizzle swz, bool promote);

// event profiling hooks (ffm_profile_*): call around the kernel launch
struct GemmProfileScope {
  cudaEvent_t start = nullptr, stop = nullptr;
  bool active = false;
};
int gemm_profile_begin(GemmProfileScope* sc, cudaStream_t stream);
int gemm_profile_end(GemmProfileScope* sc, int T, int K, int N, cudaStream_t stream);

// CTA-pair (cta_group::2) build, svlora_gemm_pair.cu
int launch_svlora_gemm_pair(const GemmOperands& o, cudaStream_t stream);

#ifdef __CUDACC__
// sigmoid(1.702 u) = 0.5 + 0.5 tanh(0.851 u): ONE MUFU op (tanh.approx, rel. error ~2^-11 << bf16 output rounding)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quick_gelu(float u) {
  const float hu = 0.5f * u;
  return fmaf(hu, tanh_approx(0.851f * u), hu);
}
__device__ __forceinline__ float quick_gelu_grad(float u) {
  const float s = fmaf(0.5f, tanh_approx(0.851f * u), 0.5f);
  return s * fmaf(1.702f * u, 1.0f - s, 1.0f);
}
#endif

}  // namespace ffm
