// Merged weight of a plain LoRA projection (scope row a4): LoRALinear.weight(x, attr) of trainers/GLP_OT_SVLoRA.py:235-239,
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
  const int i = threadIdx.x & (MW_TI - 1), ob = (threadIdx.x >> 7) * 16;
  float acc[16];
  // the frozen weight first (in flight while the tiles are staged and contracted); the low-rank term accumulates on top
#pragma unroll
  for (int k = 0; k < 16; ++k) {
    const int o = o0 + ob + k;
    acc[k] = (i0 + i < in_f && o < out_f) ? __ldg(W + static_cast<size_t>(o) * in_f + i0 + i) : 0.f;
  }
  __syncthreads();
#pragma unroll 4
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
        out[static_cast<size_t>(o) * in_f + i0 + i] = acc[k];
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
      for (int ob = 0; ob < MW_TO; ob += 8) {
        float d[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)       // eight independent loads in flight (rows past the range contribute zero)
          d[u] = (ob + u < n) ? __ldg(dWm + static_cast<size_t>(ot + ob + u) * in_f + i0 + i) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float4 b = *reinterpret_cast<const float4*>(&Bt_s[ob + u][jb + 4 * q]);
            acc[4 * q + 0] = fmaf(d[u], b.x, acc[4 * q + 0]);
            acc[4 * q + 1] = fmaf(d[u], b.y, acc[4 * q + 1]);
            acc[4 * q + 2] = fmaf(d[u], b.z, acc[4 * q + 2]);
            acc[4 * q + 3] = fmaf(d[u], b.w, acc[4 * q + 3]);
          }
        }
      }
    }
  }
  if (i0 + i < in_f) {
    float* dst = part + (static_cast<size_t>(blockIdx.y) * in_f + i0 + i) * r;
#pragma unroll
    for (int k = 0; k < 16; ++k)
      if (jb + k < r) dst[jb + k] = scaling * acc[k];
  }
}

// dst[e] = sum over `splits` partials of n elements each, in index order (deterministic)
__global__ void lora_merge_fold_kernel(const float* __restrict__ part, float* __restrict__ dst, int n, int splits) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  float s = 0.f;
  for (int k = 0; k < splits; ++k) s += part[static_cast<size_t>(k) * n + e];
  dst[e] = s;
}

// dB partials: block (16 output rows, split of the input features); dWm rows and the A rows of a 128-feature chunk are
// staged in shared memory, thread (o, j pair) accumulates two outputs
constexpr int MW_DB_O = 16;
constexpr int MW_DB_I = 128;
constexpr int MW_DB_SPLIT = 4;

__global__ void __launch_bounds__(256)
lora_merge_db_kernel(const float* __restrict__ dWm, const float* __restrict__ A, float* __restrict__ part, int out_f,
                     int in_f, int r, float scaling) {
  __shared__ __align__(16) float dW_s[MW_DB_O][MW_DB_I];
  __shared__ __align__(16) float A_s[MW_DB_I][MW_RMAX];
  const int o0 = blockIdx.x * MW_DB_O;
  const int chunk = ((in_f + MW_DB_SPLIT - 1) / MW_DB_SPLIT + MW_DB_I - 1) / MW_DB_I * MW_DB_I;
  const int i_beg = blockIdx.y * chunk, i_end = min(in_f, i_beg + chunk);
  const int o = threadIdx.x >> 4, j = (threadIdx.x & 15) * 2;
  float acc0 = 0.f, acc1 = 0.f;
  for (int it = i_beg; it < i_end; it += MW_DB_I) {
    __syncthreads();
    for (int e = threadIdx.x; e < MW_DB_O * MW_DB_I; e += 256) {
      const int oo = e / MW_DB_I, ii = e - oo * MW_DB_I;
      dW_s[oo][ii] = (o0 + oo < out_f && it + ii < i_end) ? __ldg(dWm + static_cast<size_t>(o0 + oo) * in_f + it + ii) : 0.f;
    }
    for (int e = threadIdx.x; e < MW_DB_I * MW_RMAX; e += 256) {
      const int ii = e / MW_RMAX, jj = e - ii * MW_RMAX;
      A_s[ii][jj] = (jj < r && it + ii < i_end) ? __ldg(A + static_cast<size_t>(it + ii) * r + jj) : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int ii = 0; ii < MW_DB_I; ++ii) {
      const float d = dW_s[o][ii];
      const float2 a = *reinterpret_cast<const float2*>(&A_s[ii][j]);
      acc0 = fmaf(d, a.x, acc0);
      acc1 = fmaf(d, a.y, acc1);
    }
  }
  if (o0 + o < out_f) {
    float* dst = part + static_cast<size_t>(blockIdx.y) * r * out_f;
    if (j < r) dst[static_cast<size_t>(j) * out_f + o0 + o] = scaling * acc0;
    if (j + 1 < r) dst[static_cast<size_t>(j + 1) * out_f + o0 + o] = scaling * acc1;
  }
}

}  // namespace ffm

using namespace ffm;

extern "C" {

size_t ffm_lora_merged_weight_ws_bytes(int out_f, int in_f) {
  const size_t a = static_cast<size_t>(MW_SPLIT) * static_cast<size_t>(in_f > 0 ? in_f : 0) * MW_RMAX;
  const size_t b = static_cast<size_t>(MW_DB_SPLIT) * static_cast<size_t>(out_f > 0 ? out_f : 0) * MW_RMAX;
  return (a + b) * sizeof(float);
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
  FFM_CHECK_ARG(ws_bytes >= ffm_lora_merged_weight_ws_bytes(out_f, in_f),
                "ffm_lora_merged_weight_bwd: workspace too small");
  float* part_a = ws;
  float* part_b = ws + static_cast<size_t>(MW_SPLIT) * in_f * MW_RMAX;
  dim3 grid_a((in_f + MW_TI - 1) / MW_TI, MW_SPLIT);
  lora_merge_da_kernel<<<grid_a, 256, 0, stream>>>(dWm, B, part_a, out_f, in_f, r, scaling);
  FFM_CHECK_CUDA(cudaGetLastError());
  lora_merge_fold_kernel<<<(in_f * r + 255) / 256, 256, 0, stream>>>(part_a, dA, in_f * r, MW_SPLIT);
  FFM_CHECK_CUDA(cudaGetLastError());
  dim3 grid_b((out_f + MW_DB_O - 1) / MW_DB_O, MW_DB_SPLIT);
  lora_merge_db_kernel<<<grid_b, 256, 0, stream>>>(dWm, A, part_b, out_f, in_f, r, scaling);
  FFM_CHECK_CUDA(cudaGetLastError());
  lora_merge_fold_kernel<<<(out_f * r + 255) / 256, 256, 0, stream>>>(part_b, dB, out_f * r, MW_DB_SPLIT);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(4);
  return FFM_OK;
}

}  // extern "C"
