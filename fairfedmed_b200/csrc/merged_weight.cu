// Merged weight of a plain LoRA projection (scope row a4): LoRALinear.weight(x, attr) of trainers/GLP_OT_SVLoRA.py:236-240,
//     Wm[o, i] = W[o, i] + scaling * sum_j A[i, j] B[j, o]          A = lora_A.weight [in, r], B = lora_B.weight [r, out]
// which the RN50 attention pool hands to F.multi_head_attention_forward (clip/model.py:88-97) for its q / k / v / c
// projections, and its backward  dA[i, j] = scaling sum_o dWm[o, i] B[j, o],  dB[j, o] = scaling sum_i dWm[o, i] A[i, j].
// fp32 throughout (the merged weight feeds an fp32 library projection).  The reference spends a small GEMM, a transpose, a
// scale and an add (four passes over [out, in]) forward and two GEMMs plus glue backward; here the forward is ONE pass
// (read W, write Wm: HBM-bound, 8 B per element) and the backward reads dWm twice (once per factor) with fixed-order
// reductions (deterministic, no atomics).
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int MW_RMAX = 32;       // adapter rank limit (as the fused GEMM)
constexpr int MW_TI = 128;        // input-feature columns per block
constexpr int MW_TO = 32;         // output-feature rows per block
constexpr int MW_SPLIT = 16;      // partial sums of dA over output rows

__global__ void __launch_bounds__(256)
lora_merge_fwd_kernel(const float* __restrict__ W, const float* __restrict__ A, const float* __restrict__ B,
                      float* __restrict__ out, int out_f, int in_f, int r, float scaling) {
  __shared__ float A_s[MW_TI][MW_RMAX + 1];
  __shared__ __align__(16) float B_s[MW_RMAX][MW_TO];
  const int i0 = blockIdx.x * MW_TI, o0 = blockIdx.y * MW_TO;
  for (int e = threadIdx.x; e < MW_TI * MW_RMAX; e += 256) {
    const int i = e / MW_RMAX, j = e - i * MW_RMAX;
    A_s[i][j] = (j < r && i0 + i < in_f) ? __ldg(A + static_cast<size_t>(i0 + i) * r + j) : 0.f;
  }
  for (int e = threadIdx.x; e < MW_RMAX * MW_TO; e += 256) {
    const int j = e / MW_TO, o = e - j * MW_TO;
    B_s[j][o] = (j < r && o0 + o < out_f) ? scaling * __ldg(B + static_cast<size_t>(j) * out_f + o0 + o) : 0.f;
  }
  __syncthreads();
  const int i = threadIdx.x & (MW_TI - 1), ob = (threadIdx.x >> 7) * 16;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (int j = 0; j < r; ++j) {
    const float a = A_s[i][j];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 b = *reinterpret_cast<const float4*>(&B_s[j][ob + 4 * q]);
      acc[4 * q + 0] = fmaf(a, b.x, acc[4 * q + 0]);
      acc[4 * q + 1] = fmaf(a, b.y, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(a, b.z, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(a, b.w, acc[4 * q + 3]);
    }
  }
  if (i0 + i < in_f) {
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const int o = o0 + ob + k;
      if (o < out_f) {
        const size_t idx = static_cast<size_t>(o) * in_f + i0 + i;
        out[idx] = __ldg(W + idx) + acc[k];
      }
    }
  }
}

// dA partials: block (i tile, split s) sums its share of the output rows; lanes run along i (coalesced dWm reads)
__global__ void __launch_bounds__(256)
lora_merge_da_kernel(const float* __restrict__ dWm, const float* __restrict__ B, float* __restrict__ part, int out_f,
                     int in_f, int r, float scaling) {
  __shared__ __align__(16) float Bt_s[MW_TO][MW_RMAX];       // [o][j]
  const int i0 = blockIdx.x * MW_TI;
  const int chunk = (out_f + MW_SPLIT - 1) / MW_SPLIT;
  const int o_beg = blockIdx.y * chunk, o_end = min(out_f, o_beg + chunk);
  const int i = threadIdx.x & (MW_TI - 1), jb = (threadIdx.x >> 7) * 16;
  float acc[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) acc[k] = 0.f;
  for (int ot = o_beg; ot < o_end; ot += MW_TO) {
    __syncthreads();
    for (int e = threadIdx.x; e < MW_TO * MW_RMAX; e += 256) {
      const int j = e / MW_TO, o = e - j * MW_TO;            // consecutive threads -> consecutive o (coalesced B rows)
      Bt_s[o][j] = (j < r && ot + o < o_end) ? __ldg(B + static_cast<size_t>(j) * out_f + ot + o) : 0.f;
    }
    __syncthreads();
    if (i0 + i < in_f) {
      const int n = min(MW_TO, o_end - ot);
      for (int o = 0; o < n; ++o) {
        const float d = __ldg(dWm + static_cast<size_t>(ot + o) * in_f + i0 + i);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 b = *reinterpret_cast<const float4*>(&Bt_s[o][jb + 4 * q]);
          acc[4 * q + 0] = fmaf(d, b.x, acc[4 * q + 0]);
          acc[4 * q + 1] = fmaf(d, b.y, acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(d, b.z, acc[4 * q + 2]);
          acc[4 * q + 3] = fmaf(d, b.w, acc[4 * q + 3]);
        }
      }
    }
  }
  if (i0 + i < in_f) {
    float* dst = part + (static_cast<size_t>(blockIdx.y) * in_f + i0 + i) * MW_RMAX + jb;
#pragma unroll
    for (int k = 0; k < 16; ++k) dst[k] = scaling * acc[k];
  }
}

__global__ void lora_merge_da_fold_kernel(const float* __restrict__ part, float* __restrict__ dA, int in_f, int r) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= in_f * r) return;
  const int i = e / r, j = e - i * r;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MW_SPLIT; ++k) s += part[(static_cast<size_t>(k) * in_f + i) * MW_RMAX + j];
  dA[e] = s;
}

// dB: one warp per output row o; lanes stride over i, 32 accumulators (one per j), butterfly at the end
__global__ void __launch_bounds__(256)
lora_merge_db_kernel(const float* __restrict__ dWm, const float* __restrict__ A, float* __restrict__ dB, int out_f, int in_f,
                     int r, float scaling) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int o = blockIdx.x * 8 + warp;
  if (o >= out_f) return;
  float acc[MW_RMAX];
#pragma unroll
  for (int j = 0; j < MW_RMAX; ++j) acc[j] = 0.f;
  const float* drow = dWm + static_cast<size_t>(o) * in_f;
  if (r == MW_RMAX) {
    for (int i = lane; i < in_f; i += 32) {
      const float d = __ldg(drow + i);
      const float4* ar = reinterpret_cast<const float4*>(A + static_cast<size_t>(i) * MW_RMAX);
#pragma unroll
      for (int q = 0; q < MW_RMAX / 4; ++q) {
        const float4 a = __ldg(ar + q);
        acc[4 * q + 0] = fmaf(d, a.x, acc[4 * q + 0]);
        acc[4 * q + 1] = fmaf(d, a.y, acc[4 * q + 1]);
        acc[4 * q + 2] = fmaf(d, a.z, acc[4 * q + 2]);
        acc[4 * q + 3] = fmaf(d, a.w, acc[4 * q + 3]);
      }
    }
  } else {
    for (int i = lane; i < in_f; i += 32) {
      const float d = __ldg(drow + i);
      const float* ar = A + static_cast<size_t>(i) * r;
#pragma unroll
      for (int j = 0; j < MW_RMAX; ++j)
        if (j < r) acc[j] = fmaf(d, __ldg(ar + j), acc[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < MW_RMAX; ++j) {
    float v = acc[j];
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
    if (lane == 0 && j < r) dB[static_cast<size_t>(j) * out_f + o] = scaling * v;
  }
}

}  // namespace ffm

using namespace ffm;

extern "C" {

size_t ffm_lora_merged_weight_ws_bytes(int in_f) {
  return static_cast<size_t>(MW_SPLIT) * static_cast<size_t>(in_f > 0 ? in_f : 0) * MW_RMAX * sizeof(float);
}

int ffm_lora_merged_weight(const float* W, const float* A, const float* B, float* out, int out_f, int in_f, int r,
                           float scaling, cudaStream_t stream) {
  FFM_CHECK_ARG(W && A && B && out, "ffm_lora_merged_weight: null pointer argument");
  FFM_CHECK_ARG(out_f >= 1 && in_f >= 1 && r >= 1 && r <= MW_RMAX, "ffm_lora_merged_weight: rank must be 1..32");
  dim3 grid((in_f + MW_TI - 1) / MW_TI, (out_f + MW_TO - 1) / MW_TO);
  lora_merge_fwd_kernel<<<grid, 256, 0, stream>>>(W, A, B, out, out_f, in_f, r, scaling);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_lora_merged_weight_bwd(const float* dWm, const float* A, const float* B, float* dA, float* dB, float* ws,
                               size_t ws_bytes, int out_f, int in_f, int r, float scaling, cudaStream_t stream) {
  FFM_CHECK_ARG(dWm && A && B && dA && dB && ws, "ffm_lora_merged_weight_bwd: null pointer argument");
  FFM_CHECK_ARG(out_f >= 1 && in_f >= 1 && r >= 1 && r <= MW_RMAX, "ffm_lora_merged_weight_bwd: rank must be 1..32");
  FFM_CHECK_ARG(ws_bytes >= ffm_lora_merged_weight_ws_bytes(in_f), "ffm_lora_merged_weight_bwd: workspace too small");
  FFM_CHECK_ARG(r != MW_RMAX || (reinterpret_cast<uintptr_t>(A) & 15u) == 0, "ffm_lora_merged_weight_bwd: A must be 16-byte aligned");
  dim3 grid_a((in_f + MW_TI - 1) / MW_TI, MW_SPLIT);
  lora_merge_da_kernel<<<grid_a, 256, 0, stream>>>(dWm, B, ws, out_f, in_f, r, scaling);
  FFM_CHECK_CUDA(cudaGetLastError());
  lora_merge_da_fold_kernel<<<(in_f * r + 255) / 256, 256, 0, stream>>>(ws, dA, in_f, r);
  FFM_CHECK_CUDA(cudaGetLastError());
  lora_merge_db_kernel<<<(out_f + 7) / 8, 256, 0, stream>>>(dWm, A, dB, out_f, in_f, r, scaling);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(3);
  return FFM_OK;
}

}  // extern "C"
