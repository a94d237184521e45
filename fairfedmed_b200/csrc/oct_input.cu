// OCT input side of CustomCLIP.forward (trainers/GLP_OT_SVLoRA.py:681-693, scope row f3): after the trainable slice
// projection y = Conv2d(8 -> 3, 5x5)(slices / 255) every slice-image is min-max scaled over all its pixels,
//     z = (y - lo) / (hi - lo + 1e-5),   lo = amin(y), hi = amax(y) over (C, H, W),
// normalised with the CLIP mean / std and handed to the stride-P patch convolution.  The reference spends two reductions and
// six element-wise passes (plus their autograd mirrors) on [B', 3, H, W] fp32 tensors; here:
//   forward : oct_minmax_kernel (one read of y) + oct_patchify_kernel (one read of y, bf16 im2col rows written once)
//   backward: oct_input_bwd_kernel — one block per slice-image: pass 1 reduces S1 = sum g, S2 = sum g (y - lo) and the
//             number of pixels attaining lo / hi in a fixed order (deterministic), pass 2 writes
//             d_y = g / R + [y == lo] d_lo / n_lo + [y == hi] d_hi / n_hi   with g = d_patch / std, R = hi - lo + 1e-5,
//             d_lo = -S1 / R + S2 / R^2, d_hi = -S2 / R^2  (torch.amin / amax split the gradient evenly among ties).
// The gradient of the projection's weights is then the library's convolution weight-gradient on d_y (the projection itself
// stays a library convolution this round).  HBM-bound: 4 B/pixel in + 2 B out forward, 2 + 4 in + 4 out backward.
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int OCT_THREADS = 1024;

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_min, bool is_max) {
  // fixed-order tree: lanes by shuffle, then warps by the first warp
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_min ? fminf(v, w) : (is_max ? fmaxf(v, w) : v + w);
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : (is_min ? 3.0e38f : (is_max ? -3.0e38f : 0.f));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float w = __shfl_xor_sync(0xffffffffu, t, o);
      t = is_min ? fminf(t, w) : (is_max ? fmaxf(t, w) : t + w);
    }
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// one block per slice-image: lo / hi over its n contiguous floats
__global__ void __launch_bounds__(OCT_THREADS)
oct_minmax_kernel(const float* __restrict__ y, float* __restrict__ lo, float* __restrict__ hi, int n) {
  __shared__ float red[33];
  const float4* p = reinterpret_cast<const float4*>(y + static_cast<size_t>(blockIdx.x) * n);
  float mn = 3.0e38f, mx = -3.0e38f;
  for (int i = threadIdx.x; i < n / 4; i += OCT_THREADS) {
    const float4 v = __ldg(p + i);
    mn = fminf(fminf(mn, v.x), fminf(v.y, fminf(v.z, v.w)));
    mx = fmaxf(fmaxf(mx, v.x), fmaxf(v.y, fmaxf(v.z, v.w)));
  }
  mn = block_reduce(mn, red, true, false);
  mx = block_reduce(mx, red, false, true);
  if (threadIdx.x == 0) { lo[blockIdx.x] = mn; hi[blockIdx.x] = mx; }
}

// one thread = 8 consecutive pixels of one patch row (as patchify_normalize_kernel): bf16(((y - lo) / R - mean) / std)
__global__ void __launch_bounds__(256)
oct_patchify_kernel(const float* __restrict__ y, const float* __restrict__ lo, const float* __restrict__ hi,
                    __nv_bfloat16* __restrict__ out, const float* __restrict__ mean, const float* __restrict__ stdv, int C,
                    int H, int W, int P, int gw, int G, long long n_chunks) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  const int pp8 = (P * P) >> 3;
  const int row_chunks = C * pp8;
  const long long orow = i / row_chunks;
  const int k8 = static_cast<int>(i - orow * row_chunks);
  const int c = k8 / pp8;
  const int rem = k8 - c * pp8;
  const int p8 = P >> 3;
  const int py = rem / p8, px = (rem - py * p8) << 3;
  const int b = static_cast<int>(orow / G), g = static_cast<int>(orow - static_cast<long long>(b) * G);
  const int gy = g / gw, gx = g - gy * gw;
  const float* src = y + ((static_cast<size_t>(b) * C + c) * H + static_cast<size_t>(gy) * P + py) * W +
                     static_cast<size_t>(gx) * P + px;
  const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
  const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
  float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  const float l = __ldg(lo + b), R = (__ldg(hi + b) - l) + 1e-5f;
  const float m = __ldg(mean + c), s = __ldg(stdv + c);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] = __fdiv_rn(__fdiv_rn(v[e] - l, R) - m, s);   // the reference's operation order
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]);
  pk.y = pack_bf16x2(v[2], v[3]);
  pk.z = pack_bf16x2(v[4], v[5]);
  pk.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(out)[i] = pk;
}

// one block per slice-image; pixel index i -> (c, h, w) and its position in the bf16 patch rows
__global__ void __launch_bounds__(OCT_THREADS)
oct_input_bwd_kernel(const __nv_bfloat16* __restrict__ d_patches, const float* __restrict__ y,
                     const float* __restrict__ lo, const float* __restrict__ hi, const float* __restrict__ stdv,
                     float* __restrict__ d_y, int C, int H, int W, int P, int gw, int G) {
  __shared__ float red[33];
  const int b = blockIdx.x;
  const int n = C * H * W;
  const float l = lo[b], h = hi[b], R = (h - l) + 1e-5f;
  const float* yb = y + static_cast<size_t>(b) * n;
  float* dyb = d_y + static_cast<size_t>(b) * n;
  const __nv_bfloat16* dpb = d_patches + static_cast<size_t>(b) * G * C * P * P;
  const int row_len = C * P * P;

  auto grad_at = [&](int i) -> float {      // g = d_out / std at pixel i of this slice-image
    const int c = i / (H * W);
    const int r = i - c * H * W;
    const int hh = r / W, ww = r - hh * W;
    const int g = (hh / P) * gw + ww / P;
    const int k = c * P * P + (hh % P) * P + (ww % P);
    return __bfloat162float(dpb[static_cast<size_t>(g) * row_len + k]) / __ldg(stdv + c);
  };

  float s1 = 0.f, s2 = 0.f, n_lo = 0.f, n_hi = 0.f;
  for (int i = threadIdx.x; i < n; i += OCT_THREADS) {
    const float g = grad_at(i), yv = yb[i];
    s1 += g;
    s2 = fmaf(g, yv - l, s2);
    n_lo += (yv == l) ? 1.f : 0.f;
    n_hi += (yv == h) ? 1.f : 0.f;
  }
  s1 = block_reduce(s1, red, false, false);
  s2 = block_reduce(s2, red, false, false);
  n_lo = block_reduce(n_lo, red, false, false);
  n_hi = block_reduce(n_hi, red, false, false);
  const float inv_r = 1.0f / R;
  const float d_lo = (-s1 * inv_r + s2 * inv_r * inv_r) / fmaxf(n_lo, 1.f);
  const float d_hi = (-s2 * inv_r * inv_r) / fmaxf(n_hi, 1.f);
  for (int i = threadIdx.x; i < n; i += OCT_THREADS) {
    const float yv = yb[i];
    float d = grad_at(i) * inv_r;
    if (yv == l) d += d_lo;
    if (yv == h) d += d_hi;
    dyb[i] = d;
  }
}

}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_oct_minmax_patchify(const float* y, float* lo, float* hi, void* patches, const float* mean, const float* stdv,
                            int Bp, int C, int H, int W, int patch, cudaStream_t stream) {
  FFM_CHECK_ARG(y && lo && hi && patches && mean && stdv, "ffm_oct_minmax_patchify: null pointer argument");
  FFM_CHECK_ARG(Bp >= 1 && C >= 1 && patch >= 8 && patch % 8 == 0 && H % patch == 0 && W % patch == 0 && W % 4 == 0,
                "ffm_oct_minmax_patchify: patch must be a multiple of 8 dividing H and W, W a multiple of 4");
  const int n = C * H * W;
  oct_minmax_kernel<<<Bp, OCT_THREADS, 0, stream>>>(y, lo, hi, n);
  FFM_CHECK_CUDA(cudaGetLastError());
  const int gh = H / patch, gw = W / patch, G = gh * gw;
  const long long n_chunks = static_cast<long long>(Bp) * G * C * patch * patch / 8;
  const long long blocks = (n_chunks + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_oct_minmax_patchify: too many elements");
  oct_patchify_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(y, lo, hi, static_cast<__nv_bfloat16*>(patches),
                                                                         mean, stdv, C, H, W, patch, gw, G, n_chunks);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return FFM_OK;
}

int ffm_oct_input_bwd(const void* d_patches, const float* y, const float* lo, const float* hi, const float* stdv,
                      float* d_y, int Bp, int C, int H, int W, int patch, cudaStream_t stream) {
  FFM_CHECK_ARG(d_patches && y && lo && hi && stdv && d_y, "ffm_oct_input_bwd: null pointer argument");
  FFM_CHECK_ARG(Bp >= 1 && C >= 1 && patch >= 1 && H % patch == 0 && W % patch == 0, "ffm_oct_input_bwd: bad sizes");
  const int gw = W / patch, G = (H / patch) * gw;
  oct_input_bwd_kernel<<<Bp, OCT_THREADS, 0, stream>>>(static_cast<const __nv_bfloat16*>(d_patches), y, lo, hi, stdv, d_y,
                                                       C, H, W, patch, gw, G);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
