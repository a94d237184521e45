// OCT input side of CustomCLIP.forward (trainers/GLP_OT_SVLoRA.py:681-693, scope row f3): after the trainable slice
// projection y = Conv2d(8 -> 3, 5x5)(slices / 255) every slice-image is min-max scaled over all its pixels,
//     z = (y - lo) / (hi - lo + 1e-5),   lo = amin(y), hi = amax(y) over (C, H, W),
// normalised with the CLIP mean / std and handed to the stride-P patch convolution.  The reference spends two reductions and
// six element-wise passes (plus their autograd mirrors) on [B', 3, H, W] fp32 tensors; here:
//   forward : oct_minmax_kernel (one read of y) + oct_patchify_kernel (one read of y, bf16 im2col rows written once)
//   backward: oct_bwd_reduce_kernel reduces S1 = sum g, S2 = sum g (y - lo) and the number of pixels attaining lo / hi per
//             slice-image in a fixed order (deterministic), oct_bwd_apply_kernel writes
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

// ---- backward ----
// Work item = 8 consecutive pixels of one patch row (the forward's mapping): d_patches is read as one 16-byte load, y as two.
constexpr int OCT_SPLIT = 8;       // blocks per slice-image in the reduction pass

struct OctChunk {
  const float* y;       // 8 pixels of the slice-image
  size_t pix;           // their offset inside the [C, H, W] image
  int c;
};

__device__ __forceinline__ OctChunk oct_chunk(int k, int C, int H, int W, int P, int gw) {
  // k enumerates the 16-byte chunks of one image's patch rows: [g][c][py][px8]
  const int p8 = P >> 3, pp8 = (P * P) >> 3, row_chunks = C * pp8;
  const int g = k / row_chunks, k8 = k - g * row_chunks;
  const int c = k8 / pp8, rem = k8 - c * pp8;
  const int py = rem / p8, px = (rem - py * p8) << 3;
  const int gy = g / gw, gx = g - gy * gw;
  OctChunk o;
  o.c = c;
  o.pix = (static_cast<size_t>(c) * H + static_cast<size_t>(gy) * P + py) * W + static_cast<size_t>(gx) * P + px;
  o.y = nullptr;
  return o;
}

__device__ __forceinline__ void unpack_bf16x8(const uint4& raw, float (&g)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); g[2 * e] = f.x; g[2 * e + 1] = f.y; }
}

// pass 1: block (split, image) -> partial S1 = sum g, S2 = sum g (y - lo), n_lo, n_hi  (g = d_patch / std[c])
__global__ void __launch_bounds__(256)
oct_bwd_reduce_kernel(const __nv_bfloat16* __restrict__ d_patches, const float* __restrict__ y,
                      const float* __restrict__ lo, const float* __restrict__ hi, const float* __restrict__ stdv,
                      float4* __restrict__ part, int C, int H, int W, int P, int gw, int chunks_per_image) {
  __shared__ float red[33];
  const int b = blockIdx.y;
  const float l = __ldg(lo + b), h = __ldg(hi + b);
  const uint4* dp = reinterpret_cast<const uint4*>(d_patches) + static_cast<size_t>(b) * chunks_per_image;
  const float* yb = y + static_cast<size_t>(b) * C * H * W;
  const int per = (chunks_per_image + OCT_SPLIT - 1) / OCT_SPLIT;
  const int k_end = min(chunks_per_image, (static_cast<int>(blockIdx.x) + 1) * per);
  float s1 = 0.f, s2 = 0.f, n_lo = 0.f, n_hi = 0.f;
  for (int k = blockIdx.x * per + threadIdx.x; k < k_end; k += 256) {
    const OctChunk ch = oct_chunk(k, C, H, W, P, gw);
    const uint4 raw = __ldg(dp + k);
    const float4 y0 = __ldg(reinterpret_cast<const float4*>(yb + ch.pix));
    const float4 y1 = __ldg(reinterpret_cast<const float4*>(yb + ch.pix) + 1);
    const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
    float g[8];
    unpack_bf16x8(raw, g);
    const float inv_s = 1.0f / __ldg(stdv + ch.c);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float ge = g[e] * inv_s;
      s1 += ge;
      s2 = fmaf(ge, yv[e] - l, s2);
      n_lo += (yv[e] == l) ? 1.f : 0.f;
      n_hi += (yv[e] == h) ? 1.f : 0.f;
    }
  }
  s1 = block_reduce(s1, red, false, false);
  s2 = block_reduce(s2, red, false, false);
  n_lo = block_reduce(n_lo, red, false, false);
  n_hi = block_reduce(n_hi, red, false, false);
  if (threadIdx.x == 0) part[static_cast<size_t>(b) * OCT_SPLIT + blockIdx.x] = make_float4(s1, s2, n_lo, n_hi);
}

// pass 2: coefficients per slice-image (partials folded in a fixed order), then
//   d_y = g / R + [y == lo] d_lo / n_lo + [y == hi] d_hi / n_hi
__global__ void __launch_bounds__(256)
oct_bwd_apply_kernel(const __nv_bfloat16* __restrict__ d_patches, const float* __restrict__ y,
                     const float* __restrict__ lo, const float* __restrict__ hi, const float* __restrict__ stdv,
                     const float4* __restrict__ part, float* __restrict__ d_y, int C, int H, int W, int P, int gw,
                     int chunks_per_image) {
  const int b = blockIdx.y;
  const int k = blockIdx.x * 256 + threadIdx.x;
  if (k >= chunks_per_image) return;
  const float l = __ldg(lo + b), h = __ldg(hi + b);
  float s1 = 0.f, s2 = 0.f, n_lo = 0.f, n_hi = 0.f;
#pragma unroll
  for (int q = 0; q < OCT_SPLIT; ++q) {
    const float4 v = __ldg(part + static_cast<size_t>(b) * OCT_SPLIT + q);
    s1 += v.x; s2 += v.y; n_lo += v.z; n_hi += v.w;
  }
  const float inv_r = 1.0f / ((h - l) + 1e-5f);
  const float d_lo = (-s1 * inv_r + s2 * inv_r * inv_r) / fmaxf(n_lo, 1.f);
  const float d_hi = (-s2 * inv_r * inv_r) / fmaxf(n_hi, 1.f);
  const OctChunk ch = oct_chunk(k, C, H, W, P, gw);
  const size_t base = static_cast<size_t>(b) * C * H * W + ch.pix;
  const uint4 raw = __ldg(reinterpret_cast<const uint4*>(d_patches) + static_cast<size_t>(b) * chunks_per_image + k);
  const float4 y0 = __ldg(reinterpret_cast<const float4*>(y + base));
  const float4 y1 = __ldg(reinterpret_cast<const float4*>(y + base) + 1);
  const float yv[8] = {y0.x, y0.y, y0.z, y0.w, y1.x, y1.y, y1.z, y1.w};
  float g[8], o[8];
  unpack_bf16x8(raw, g);
  const float sc = inv_r / __ldg(stdv + ch.c);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    float d = g[e] * sc;
    if (yv[e] == l) d += d_lo;
    if (yv[e] == h) d += d_hi;
    o[e] = d;
  }
  reinterpret_cast<float4*>(d_y + base)[0] = make_float4(o[0], o[1], o[2], o[3]);
  reinterpret_cast<float4*>(d_y + base)[1] = make_float4(o[4], o[5], o[6], o[7]);
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

size_t ffm_oct_input_bwd_ws_bytes(int Bp) {
  return static_cast<size_t>(Bp > 0 ? Bp : 0) * OCT_SPLIT * sizeof(float4);
}

int ffm_oct_input_bwd(const void* d_patches, const float* y, const float* lo, const float* hi, const float* stdv,
                      float* d_y, void* ws, size_t ws_bytes, int Bp, int C, int H, int W, int patch,
                      cudaStream_t stream) {
  FFM_CHECK_ARG(d_patches && y && lo && hi && stdv && d_y && ws, "ffm_oct_input_bwd: null pointer argument");
  FFM_CHECK_ARG(Bp >= 1 && Bp <= 65535 && C >= 1 && patch >= 8 && patch % 8 == 0 && H % patch == 0 && W % patch == 0 &&
                    W % 4 == 0,
                "ffm_oct_input_bwd: patch must be a multiple of 8 dividing H and W, W a multiple of 4, Bp <= 65535");
  FFM_CHECK_ARG(ws_bytes >= ffm_oct_input_bwd_ws_bytes(Bp), "ffm_oct_input_bwd: workspace too small");
  const int gw = W / patch;
  const int chunks = C * H * W / 8;
  float4* part = static_cast<float4*>(ws);
  oct_bwd_reduce_kernel<<<dim3(OCT_SPLIT, Bp), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(d_patches), y, lo, hi,
                                                                  stdv, part, C, H, W, patch, gw, chunks);
  FFM_CHECK_CUDA(cudaGetLastError());
  oct_bwd_apply_kernel<<<dim3((chunks + 255) / 256, Bp), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(d_patches), y,
                                                                           lo, hi, stdv, part, d_y, C, H, W, patch, gw,
                                                                           chunks);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return FFM_OK;
}

}  // extern "C"
