// Training-mode BatchNorm2d (+ ReLU) of the ResNet trunk (clip/model.py:18-58: bn1 / bn2 / bn3 of the bottlenecks, the stem) on
// channels-last fp32 activations, viewed as a [R = B*H*W, C] matrix.  BatchNorm's affine parameters are trainable in the RN50
// recipe (scope row a8), so forward AND backward are needed:
//   forward : bn_stats (one read: per-channel sum / sum of squares, fp32 partials per row chunk, combined in fp64 in a fixed
//             order) -> bn_finalize (mean, biased variance, rstd; running statistics updated as torch does: momentum, unbiased
//             variance) -> bn_apply (y = [relu](gamma * xhat + beta): one read, one write)
//   backward: bn_bwd_reduce (reads x, dy: g = dy * [pre-activation > 0] recomputed from x, sums g and g * xhat per channel) ->
//             bn_bwd_finalize (dgamma, dbeta, the two per-channel means) -> bn_bwd_apply (dx = gamma * rstd * (g - mean(g) -
//             xhat * mean(g xhat)): reads x, dy, writes dx)
// 8 passes over the activation instead of the 13 of library BatchNorm + separate ReLU kernels; deterministic (no atomics).
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int BN_CHUNKS = 128;          // row chunks (partial sums per channel)

// block (channel tile of 128, row chunk): partial[chunk][2][C] = { sum, sum of squares }  (or { sum g, sum g*xhat })
template <bool BWD>
__global__ void __launch_bounds__(256)
bn_reduce_kernel(const float* __restrict__ x, const float* __restrict__ dy, const float* __restrict__ gamma,
                 const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                 float* __restrict__ part, long long R, int C, int relu, int QL) {
  // 256 threads = QL channel quads (a power of two <= 32: narrow layers keep every lane busy) x 256 / QL row lanes
  __shared__ float4 red[2][256];
  const int ql = threadIdx.x % QL, rl = threadIdx.x / QL;
  const int BN_ROWS_PER_BLOCK = 256 / QL;
  const int c4 = blockIdx.x * QL + ql;                 // channel quad
  const int C4 = C >> 2;
  const long long per = (R + BN_CHUNKS - 1) / BN_CHUNKS;
  const long long r0 = blockIdx.y * per, r1 = (r0 + per < R) ? r0 + per : R;
  float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
  if (c4 < C4) {
    float4 m = s0, rs = s0, ga = s0, be = s0;
    if (BWD) {
      m = __ldg(reinterpret_cast<const float4*>(mean) + c4);
      rs = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
      ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4);
      be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
    }
    auto accumulate = [&](const float4& v, const float4& d) {
      if (!BWD) {
        s0.x += v.x; s0.y += v.y; s0.z += v.z; s0.w += v.w;
        s1.x = fmaf(v.x, v.x, s1.x); s1.y = fmaf(v.y, v.y, s1.y); s1.z = fmaf(v.z, v.z, s1.z); s1.w = fmaf(v.w, v.w, s1.w);
      } else {
        const float xh[4] = {(v.x - m.x) * rs.x, (v.y - m.y) * rs.y, (v.z - m.z) * rs.z, (v.w - m.w) * rs.w};
        const float dd[4] = {d.x, d.y, d.z, d.w};
        const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
        float g[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) g[e] = (relu && fmaf(gg[e], xh[e], bb[e]) <= 0.f) ? 0.f : dd[e];
        s0.x += g[0]; s0.y += g[1]; s0.z += g[2]; s0.w += g[3];
        s1.x = fmaf(g[0], xh[0], s1.x); s1.y = fmaf(g[1], xh[1], s1.y); s1.z = fmaf(g[2], xh[2], s1.z);
        s1.w = fmaf(g[3], xh[3], s1.w);
      }
    };
    const float4* xp = reinterpret_cast<const float4*>(x);
    const float4* dp = reinterpret_cast<const float4*>(dy);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    long long r = r0 + rl;
    // four rows per trip: the loads are issued before the first use (the accumulation order stays row order)
    for (; r + 3 * BN_ROWS_PER_BLOCK < r1; r += 4 * BN_ROWS_PER_BLOCK) {
      float4 v[4], d[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        v[k] = __ldg(xp + (r + k * BN_ROWS_PER_BLOCK) * C4 + c4);
        d[k] = BWD ? __ldg(dp + (r + k * BN_ROWS_PER_BLOCK) * C4 + c4) : zero4;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) accumulate(v[k], d[k]);
    }
    for (; r < r1; r += BN_ROWS_PER_BLOCK) accumulate(__ldg(xp + r * C4 + c4), BWD ? __ldg(dp + r * C4 + c4) : zero4);
  }
  red[0][rl * QL + ql] = s0;
  red[1][rl * QL + ql] = s1;
  __syncthreads();
  if (rl < 2 && c4 < C4) {                               // row lane 0 folds the sums, row lane 1 the second moments
    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < BN_ROWS_PER_BLOCK; ++k) {
      const float4 v = red[rl][k * QL + ql];
      t.x += v.x; t.y += v.y; t.z += v.z; t.w += v.w;
    }
    reinterpret_cast<float4*>(part + (static_cast<size_t>(blockIdx.y) * 2 + rl) * C)[c4] = t;
  }
}

// forward: mean / rstd per channel, running statistics; fused scale a = gamma * rstd and shift b = beta - mean * a
__global__ void bn_finalize_kernel(const float* __restrict__ part, const float* __restrict__ gamma,
                                   const float* __restrict__ beta, float* __restrict__ running_mean,
                                   float* __restrict__ running_var, float* __restrict__ mean_out, float* __restrict__ rstd_out,
                                   float* __restrict__ ab, long long R, int C, float momentum, float eps) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < BN_CHUNKS; ++k) {
    s += static_cast<double>(part[(static_cast<size_t>(k) * 2) * C + c]);
    q += static_cast<double>(part[(static_cast<size_t>(k) * 2 + 1) * C + c]);
  }
  const double n = static_cast<double>(R);
  const double mu = s / n;
  double var = q / n - mu * mu;
  var = var < 0.0 ? 0.0 : var;
  const float rs = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  mean_out[c] = static_cast<float>(mu);
  rstd_out[c] = rs;
  const float a = gamma[c] * rs;
  ab[c] = a;
  ab[C + c] = beta[c] - static_cast<float>(mu) * a;
  if (running_mean != nullptr) {
    const double unbiased = R > 1 ? var * n / (n - 1.0) : var;
    running_mean[c] = (1.0f - momentum) * running_mean[c] + momentum * static_cast<float>(mu);
    running_var[c] = (1.0f - momentum) * running_var[c] + momentum * static_cast<float>(unbiased);
  }
}

__global__ void __launch_bounds__(256)
bn_apply_kernel(const float4* __restrict__ x, const float* __restrict__ ab, float4* __restrict__ y, long long n4, int C4,
                int relu) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c4 = static_cast<int>(i % C4);
  const float4 a = __ldg(reinterpret_cast<const float4*>(ab) + c4), b = __ldg(reinterpret_cast<const float4*>(ab) + C4 + c4);
  const float4 v = __ldg(x + i);
  float4 o = make_float4(fmaf(a.x, v.x, b.x), fmaf(a.y, v.y, b.y), fmaf(a.z, v.z, b.z), fmaf(a.w, v.w, b.w));
  if (relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
  y[i] = o;
}

// backward: dgamma = sum g xhat, dbeta = sum g; coef = { mean g, mean g xhat } per channel
__global__ void bn_bwd_finalize_kernel(const float* __restrict__ part, float* __restrict__ dgamma, float* __restrict__ dbeta,
                                       float* __restrict__ coef, long long R, int C) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  double s = 0.0, q = 0.0;
  for (int k = 0; k < BN_CHUNKS; ++k) {
    s += static_cast<double>(part[(static_cast<size_t>(k) * 2) * C + c]);
    q += static_cast<double>(part[(static_cast<size_t>(k) * 2 + 1) * C + c]);
  }
  dbeta[c] = static_cast<float>(s);
  dgamma[c] = static_cast<float>(q);
  coef[c] = static_cast<float>(s / static_cast<double>(R));
  coef[C + c] = static_cast<float>(q / static_cast<double>(R));
}

__global__ void __launch_bounds__(256)
bn_bwd_apply_kernel(const float4* __restrict__ x, const float4* __restrict__ dy, const float* __restrict__ gamma,
                    const float* __restrict__ beta, const float* __restrict__ mean, const float* __restrict__ rstd,
                    const float* __restrict__ coef, float4* __restrict__ dx, long long n4, int C4, int relu) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const int c4 = static_cast<int>(i % C4);
  const float4 m = __ldg(reinterpret_cast<const float4*>(mean) + c4), rs = __ldg(reinterpret_cast<const float4*>(rstd) + c4);
  const float4 ga = __ldg(reinterpret_cast<const float4*>(gamma) + c4), be = __ldg(reinterpret_cast<const float4*>(beta) + c4);
  const float4 mg = __ldg(reinterpret_cast<const float4*>(coef) + c4), mgx = __ldg(reinterpret_cast<const float4*>(coef) + C4 + c4);
  const float4 v = __ldg(x + i), d = __ldg(dy + i);
  const float xv[4] = {v.x, v.y, v.z, v.w}, dd[4] = {d.x, d.y, d.z, d.w};
  const float mm[4] = {m.x, m.y, m.z, m.w}, rr[4] = {rs.x, rs.y, rs.z, rs.w};
  const float gg[4] = {ga.x, ga.y, ga.z, ga.w}, bb[4] = {be.x, be.y, be.z, be.w};
  const float a0[4] = {mg.x, mg.y, mg.z, mg.w}, a1[4] = {mgx.x, mgx.y, mgx.z, mgx.w};
  float o[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float xh = (xv[e] - mm[e]) * rr[e];
    const float g = (relu && fmaf(gg[e], xh, bb[e]) <= 0.f) ? 0.f : dd[e];
    o[e] = gg[e] * rr[e] * (g - a0[e] - xh * a1[e]);
  }
  dx[i] = make_float4(o[0], o[1], o[2], o[3]);
}

}  // namespace ffm

using namespace ffm;

extern "C" {

size_t ffm_bn_ws_bytes(int C) {
  const size_t c = static_cast<size_t>(C > 0 ? C : 0);
  return (static_cast<size_t>(BN_CHUNKS) * 2 * c + 2 * c) * sizeof(float);     // partials + { a, b } / { mean g, mean g xhat }
}

int ffm_bn_relu_fwd(const float* x, const float* gamma, const float* beta, float* running_mean, float* running_var, float* y,
                    float* mean_out, float* rstd_out, void* ws, size_t ws_bytes, int64_t R, int C, float momentum, float eps,
                    int relu, cudaStream_t stream) {
  FFM_CHECK_ARG(x && gamma && beta && y && mean_out && rstd_out && ws, "ffm_bn_relu_fwd: null pointer argument");
  FFM_CHECK_ARG(R >= 1 && C >= 4 && C % 4 == 0, "ffm_bn_relu_fwd: C must be a positive multiple of 4");
  FFM_CHECK_ARG((running_mean == nullptr) == (running_var == nullptr), "ffm_bn_relu_fwd: running statistics come in pairs");
  FFM_CHECK_ARG(ws_bytes >= ffm_bn_ws_bytes(C), "ffm_bn_relu_fwd: workspace too small");
  float* part = static_cast<float*>(ws);
  float* ab = part + static_cast<size_t>(BN_CHUNKS) * 2 * C;
  const int C4 = C / 4;
  int QL = 32;
  while (QL > C4) QL >>= 1;
  bn_reduce_kernel<false><<<dim3((C4 + QL - 1) / QL, BN_CHUNKS), 256, 0, stream>>>(x, nullptr, nullptr, nullptr, nullptr,
                                                                                  nullptr, part, R, C, 0, QL);
  FFM_CHECK_CUDA(cudaGetLastError());
  bn_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(part, gamma, beta, running_mean, running_var, mean_out, rstd_out,
                                                          ab, R, C, momentum, eps);
  FFM_CHECK_CUDA(cudaGetLastError());
  const long long n4 = static_cast<long long>(R) * C4;
  const long long blocks = (n4 + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_bn_relu_fwd: too many elements");
  bn_apply_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<const float4*>(x), ab,
                                                                     reinterpret_cast<float4*>(y), n4, C4, relu);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(3);
  return FFM_OK;
}

int ffm_bn_relu_bwd(const float* x, const float* dy, const float* gamma, const float* beta, const float* mean,
                    const float* rstd, float* dx, float* dgamma, float* dbeta, void* ws, size_t ws_bytes, int64_t R, int C,
                    int relu, cudaStream_t stream) {
  FFM_CHECK_ARG(x && dy && gamma && beta && mean && rstd && dx && dgamma && dbeta && ws, "ffm_bn_relu_bwd: null pointer argument");
  FFM_CHECK_ARG(R >= 1 && C >= 4 && C % 4 == 0, "ffm_bn_relu_bwd: C must be a positive multiple of 4");
  FFM_CHECK_ARG(ws_bytes >= ffm_bn_ws_bytes(C), "ffm_bn_relu_bwd: workspace too small");
  float* part = static_cast<float*>(ws);
  float* coef = part + static_cast<size_t>(BN_CHUNKS) * 2 * C;
  const int C4 = C / 4;
  int QL = 32;
  while (QL > C4) QL >>= 1;
  bn_reduce_kernel<true><<<dim3((C4 + QL - 1) / QL, BN_CHUNKS), 256, 0, stream>>>(x, dy, gamma, beta, mean, rstd, part, R, C,
                                                                                 relu, QL);
  FFM_CHECK_CUDA(cudaGetLastError());
  bn_bwd_finalize_kernel<<<(C + 127) / 128, 128, 0, stream>>>(part, dgamma, dbeta, coef, R, C);
  FFM_CHECK_CUDA(cudaGetLastError());
  const long long n4 = static_cast<long long>(R) * C4;
  const long long blocks = (n4 + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_bn_relu_bwd: too many elements");
  bn_bwd_apply_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<const float4*>(dy), gamma, beta, mean, rstd, coef,
      reinterpret_cast<float4*>(dx), n4, C4, relu);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(3);
  return FFM_OK;
}

}  // extern "C"
