// Fused residual-add + LayerNorm for the frozen transformer blocks around the adapted MLP (sm_100a, HBM bound).
//
//   forward : s = bf16(x + res)  (res optional)      ln = (s - mean) * rstd * gamma + beta        (fp32 statistics)
//   backward: dx = rstd * (g - mean(g) - xhat * mean(g * xhat)) + d_res,   g = gamma * d_ln,  xhat = (s - mean) * rstd
//             (gamma / beta are frozen in the FairLoRA recipe: no parameter gradients)
//
// Replaces, per residual block of clip/model.py:354-357, the chain  add -> LayerNorm(fp32 statistics, :304-310)  and
// its autograd (layer_norm_backward + two adds) — 4 row-sized tensors move per call instead of 7.
// One warp per row, the row lives in registers (C = 256 * VPL, 8 bf16 per 128-bit access), two-pass variance.
// Algorithmic bytes per row: forward (2 reads + 2 writes) * C * 2, backward (3 reads + 1 write) * C * 2.
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int LN_THREADS = 256;
constexpr int LN_ROWS_PER_BLOCK = LN_THREADS / 32;

__device__ __forceinline__ void unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = __bfloat1622float2(h2[e]);
    f[2 * e] = v.x;
    f[2 * e + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 r;
  r.x = pack_bf16x2(f[0], f[1]);
  r.y = pack_bf16x2(f[2], f[3]);
  r.z = pack_bf16x2(f[4], f[5]);
  r.w = pack_bf16x2(f[6], f[7]);
  return r;
}

// Persistent rows: a warp walks rows with a grid stride, keeps ITS columns of gamma / beta in registers for the whole
// kernel (they were 12 of the 18 load instructions per row) and requests the next row before it reduces the current one.
template <int VPL>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ res,
                  const float* __restrict__ gamma, const float* __restrict__ beta, __nv_bfloat16* __restrict__ sum_out,
                  __nv_bfloat16* __restrict__ ln_out, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                  int rows, float eps) {
  constexpr int C = 256 * VPL;
  const int lane = threadIdx.x & 31;
  const int stride = gridDim.x * LN_ROWS_PER_BLOCK;
  int row = blockIdx.x * LN_ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  float gg[VPL][8], bb[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + col));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + col + 4));
    gg[i][0] = g0.x; gg[i][1] = g0.y; gg[i][2] = g0.z; gg[i][3] = g0.w;
    gg[i][4] = g1.x; gg[i][5] = g1.y; gg[i][6] = g1.z; gg[i][7] = g1.w;
    bb[i][0] = b0.x; bb[i][1] = b0.y; bb[i][2] = b0.z; bb[i][3] = b0.w;
    bb[i][4] = b1.x; bb[i][5] = b1.y; bb[i][6] = b1.z; bb[i][7] = b1.w;
  }
  uint4 nx[VPL], nr[VPL];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const size_t off = static_cast<size_t>(row) * C + (i * 32 + lane) * 8;
    nx[i] = __ldg(reinterpret_cast<const uint4*>(x + off));
    if (res != nullptr) nr[i] = __ldg(reinterpret_cast<const uint4*>(res + off));
  }
  for (; row < rows; row += stride) {
    const size_t base = static_cast<size_t>(row) * C;
    uint4 cx[VPL], cr[VPL];
#pragma unroll
    for (int i = 0; i < VPL; ++i) { cx[i] = nx[i]; cr[i] = nr[i]; }
    if (row + stride < rows) {
#pragma unroll
      for (int i = 0; i < VPL; ++i) {
        const size_t off = static_cast<size_t>(row + stride) * C + (i * 32 + lane) * 8;
        nx[i] = __ldg(reinterpret_cast<const uint4*>(x + off));
        if (res != nullptr) nr[i] = __ldg(reinterpret_cast<const uint4*>(res + off));
      }
    }
    float v[VPL][8];
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      unpack8(cx[i], v[i]);
      if (res != nullptr) {
        float r[8];
        unpack8(cr[i], r);
#pragma unroll
        for (int e = 0; e < 8; ++e) v[i][e] += r[e];
        // the residual stream is stored in bf16: normalise exactly what is stored (and what the backward will read)
        const uint4 packed = pack8(v[i]);
        *reinterpret_cast<uint4*>(sum_out + base + col) = packed;
        unpack8(packed, v[i]);
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) s += v[i][e];
    }
    const float mean = warp_sum(s) * (1.0f / C);
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < VPL; ++i)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float d = v[i][e] - mean;
        q = fmaf(d, d, q);
      }
    const float rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
    if (lane == 0) {
      mean_out[row] = mean;
      rstd_out[row] = rstd;
    }
#pragma unroll
    for (int i = 0; i < VPL; ++i) {
      const int col = (i * 32 + lane) * 8;
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = fmaf((v[i][e] - mean) * rstd, gg[i][e], bb[i][e]);
      *reinterpret_cast<uint4*>(ln_out + base + col) = pack8(o);
    }
  }
}

template <int VPL>
__global__ void __launch_bounds__(LN_THREADS)
add_ln_bwd_kernel(const __nv_bfloat16* __restrict__ d_ln, const __nv_bfloat16* __restrict__ d_res,
                  const __nv_bfloat16* __restrict__ s_in, const float* __restrict__ gamma,
                  const float* __restrict__ mean_in, const float* __restrict__ rstd_in,
                  __nv_bfloat16* __restrict__ dx, int rows) {
  constexpr int C = 256 * VPL;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * LN_ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const size_t base = static_cast<size_t>(row) * C;
  const float mean = __ldg(mean_in + row), rstd = __ldg(rstd_in + row);
  float g[VPL][8], xh[VPL][8];
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    float dy[8], sv[8];
    unpack8(__ldg(reinterpret_cast<const uint4*>(d_ln + base + col)), dy);
    unpack8(__ldg(reinterpret_cast<const uint4*>(s_in + base + col)), sv);
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + col));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + col + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      g[i][e] = dy[e] * gg[e];
      xh[i][e] = (sv[e] - mean) * rstd;
      sg += g[i][e];
      sgx = fmaf(g[i][e], xh[i][e], sgx);
    }
  }
  const float mg = warp_sum(sg) * (1.0f / C);
  const float mgx = warp_sum(sgx) * (1.0f / C);
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    float o[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = rstd * (g[i][e] - mg - xh[i][e] * mgx);
    if (d_res != nullptr) {
      float r[8];
      unpack8(__ldg(reinterpret_cast<const uint4*>(d_res + base + col)), r);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] += r[e];
    }
    *reinterpret_cast<uint4*>(dx + base + col) = pack8(o);
  }
}

}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_add_layernorm_fwd(const void* x, const void* res, const float* gamma, const float* beta, void* sum_out,
                          void* ln_out, float* mean_out, float* rstd_out, int rows, int C, float eps,
                          cudaStream_t stream) {
  FFM_CHECK_ARG(x && gamma && beta && ln_out && mean_out && rstd_out, "ffm_add_layernorm_fwd: null pointer argument");
  FFM_CHECK_ARG(res == nullptr || sum_out != nullptr, "ffm_add_layernorm_fwd: res needs sum_out");
  FFM_CHECK_ARG(rows >= 1, "ffm_add_layernorm_fwd: rows must be >= 1");
  FFM_CHECK_ARG(C % 256 == 0 && C >= 256 && C <= 1024, "ffm_add_layernorm_fwd: C (%d) must be 256, 512, 768 or 1024", C);
  int grid = (rows + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
  // persistent rows (grid-stride loop inside): one resident wave — 126 registers at C = 768 allow two CTAs per SM
  const int ctas_per_sm = C <= 512 ? 4 : (C == 768 ? 2 : 1);
  if (grid > ctas_per_sm * num_sms()) grid = ctas_per_sm * num_sms();
  const __nv_bfloat16* xp = static_cast<const __nv_bfloat16*>(x);
  const __nv_bfloat16* rp = static_cast<const __nv_bfloat16*>(res);
  __nv_bfloat16* sp = static_cast<__nv_bfloat16*>(sum_out);
  __nv_bfloat16* lp = static_cast<__nv_bfloat16*>(ln_out);
  switch (C / 256) {
    case 1: add_ln_fwd_kernel<1><<<grid, LN_THREADS, 0, stream>>>(xp, rp, gamma, beta, sp, lp, mean_out, rstd_out, rows, eps); break;
    case 2: add_ln_fwd_kernel<2><<<grid, LN_THREADS, 0, stream>>>(xp, rp, gamma, beta, sp, lp, mean_out, rstd_out, rows, eps); break;
    case 3: add_ln_fwd_kernel<3><<<grid, LN_THREADS, 0, stream>>>(xp, rp, gamma, beta, sp, lp, mean_out, rstd_out, rows, eps); break;
    default: add_ln_fwd_kernel<4><<<grid, LN_THREADS, 0, stream>>>(xp, rp, gamma, beta, sp, lp, mean_out, rstd_out, rows, eps); break;
  }
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_add_layernorm_bwd(const void* d_ln, const void* d_res, const void* s, const float* gamma, const float* mean,
                          const float* rstd, void* dx, int rows, int C, cudaStream_t stream) {
  FFM_CHECK_ARG(d_ln && s && gamma && mean && rstd && dx, "ffm_add_layernorm_bwd: null pointer argument");
  FFM_CHECK_ARG(rows >= 1, "ffm_add_layernorm_bwd: rows must be >= 1");
  FFM_CHECK_ARG(C % 256 == 0 && C >= 256 && C <= 1024, "ffm_add_layernorm_bwd: C (%d) must be 256, 512, 768 or 1024", C);
  const int grid = (rows + LN_ROWS_PER_BLOCK - 1) / LN_ROWS_PER_BLOCK;
  const __nv_bfloat16* dl = static_cast<const __nv_bfloat16*>(d_ln);
  const __nv_bfloat16* dr = static_cast<const __nv_bfloat16*>(d_res);
  const __nv_bfloat16* sp = static_cast<const __nv_bfloat16*>(s);
  __nv_bfloat16* dp = static_cast<__nv_bfloat16*>(dx);
  switch (C / 256) {
    case 1: add_ln_bwd_kernel<1><<<grid, LN_THREADS, 0, stream>>>(dl, dr, sp, gamma, mean, rstd, dp, rows); break;
    case 2: add_ln_bwd_kernel<2><<<grid, LN_THREADS, 0, stream>>>(dl, dr, sp, gamma, mean, rstd, dp, rows); break;
    case 3: add_ln_bwd_kernel<3><<<grid, LN_THREADS, 0, stream>>>(dl, dr, sp, gamma, mean, rstd, dp, rows); break;
    default: add_ln_bwd_kernel<4><<<grid, LN_THREADS, 0, stream>>>(dl, dr, sp, gamma, mean, rstd, dp, rows); break;
  }
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
