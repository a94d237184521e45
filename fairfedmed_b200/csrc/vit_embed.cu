// Input side of the ViT image tower (scope row f3, sm_100a, HBM bound): everything between the raw 0..255 image and
// the first attention block that the reference spreads over ~10 element-wise launches.
//
//   patchify_normalize : image fp32 [B', C, H, W]  ->  patches bf16 [B' * G, C * P * P]
//                        out[b, gy*gw+gx, c*P*P + py*P + px] = bf16((img[b,c,gy*P+py,gx*P+px] / 255 - mean[c]) / std[c])
//                        = CustomCLIP.forward's `/255`, mean/std normalisation (trainers/GLP_OT_SVLoRA.py:679-693), the
//                        half-precision cast and the im2col of the stride-P patch convolution (clip/model.py:431-433) in
//                        ONE pass; the convolution itself is then a plain GEMM with conv1.weight.flatten(1).
//   vit_embed_ln       : patch embeddings bf16 [B' * G, C] -> residual stream x0 bf16 [B', G+1, C] and h0 = LN_1(x0)
//                        x0 = LN_pre(cat(class_embedding, patch_emb) + positional_embedding)  (clip/model.py:434-440)
//                        fused with the first block's ln_1 (:354): one warp per token row, row in registers, fp32
//                        statistics like the reference's LayerNorm subclass (:304-310).
//
// Both are forward-only: nothing upstream of the residual stream is trainable in the 2-D recipes (conv1, class /
// positional embeddings and ln_pre are frozen and the image needs no gradient), so autograd never visits them.
// Algorithmic bytes: patchify 4 + 2 bytes per pixel (B'=64: 38.5 MB in, 19.3 MB out); embed 2 reads-equivalent + 2
// writes of [B'*(G+1), C] bf16 (19.4 MB each) + the fp32 tables.
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

// one thread = 8 consecutive pixels of one patch row (32 B in, 16 B out); consecutive threads walk the output row, so
// stores are fully coalesced and loads come in 64-byte runs (one patch row = P pixels)
__global__ void __launch_bounds__(256)
patchify_normalize_kernel(const float* __restrict__ img, __nv_bfloat16* __restrict__ out,
                          const float* __restrict__ mean, const float* __restrict__ stdv, int C, int H, int W, int P,
                          int gw, int G, long long n_chunks, int div255) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_chunks) return;
  const int pp8 = (P * P) >> 3;                 // 8-pixel chunks per (patch, channel)
  const int row_chunks = C * pp8;               // chunks per output row
  const long long orow = i / row_chunks;        // b * G + g
  const int k8 = static_cast<int>(i - orow * row_chunks);
  const int c = k8 / pp8;
  const int rem = k8 - c * pp8;
  const int p8 = P >> 3;
  const int py = rem / p8, px = (rem - py * p8) << 3;
  const int b = static_cast<int>(orow / G), g = static_cast<int>(orow - static_cast<long long>(b) * G);
  const int gy = g / gw, gx = g - gy * gw;
  const float* src = img + ((static_cast<size_t>(b) * C + c) * H + static_cast<size_t>(gy) * P + py) * W +
                     static_cast<size_t>(gx) * P + px;
  const float4 v0 = __ldg(reinterpret_cast<const float4*>(src));
  const float4 v1 = __ldg(reinterpret_cast<const float4*>(src) + 1);
  float v[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
  const float m = __ldg(mean + c), s = __ldg(stdv + c);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    // IEEE divisions, in the reference's order: image / 255, then (image - mean) / std
    const float t = div255 ? __fdiv_rn(v[e], 255.0f) : v[e];
    v[e] = __fdiv_rn(t - m, s);
  }
  uint4 pk;
  pk.x = pack_bf16x2(v[0], v[1]);
  pk.y = pack_bf16x2(v[2], v[3]);
  pk.z = pack_bf16x2(v[4], v[5]);
  pk.w = pack_bf16x2(v[6], v[7]);
  reinterpret_cast<uint4*>(out)[i] = pk;
}

constexpr int EMB_THREADS = 256;
constexpr int EMB_ROWS_PER_BLOCK = EMB_THREADS / 32;

__device__ __forceinline__ void emb_unpack8(const uint4& raw, float (&f)[8]) {
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const float2 v = __bfloat1622float2(h2[e]);
    f[2 * e] = v.x;
    f[2 * e + 1] = v.y;
  }
}
__device__ __forceinline__ uint4 emb_pack8(const float (&f)[8]) {
  uint4 r;
  r.x = pack_bf16x2(f[0], f[1]);
  r.y = pack_bf16x2(f[2], f[3]);
  r.z = pack_bf16x2(f[4], f[5]);
  r.w = pack_bf16x2(f[6], f[7]);
  return r;
}
__device__ __forceinline__ void emb_load8f(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// LayerNorm of the VPL*8 values a lane holds (row of C = 256*VPL spread over the warp), fp32, two-pass variance.
template <int VPL>
__device__ __forceinline__ void warp_row_stats(const float (&v)[VPL][8], float& mean, float& rstd, float eps) {
  constexpr int C = 256 * VPL;
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[i][e];
  mean = warp_sum(s) * (1.0f / C);
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < VPL; ++i)
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = v[i][e] - mean;
      q = fmaf(d, d, q);
    }
  rstd = rsqrtf(warp_sum(q) * (1.0f / C) + eps);
}

template <int VPL>
__global__ void __launch_bounds__(EMB_THREADS)
vit_embed_ln_kernel(const __nv_bfloat16* __restrict__ patch_emb, const float* __restrict__ cls,
                    const float* __restrict__ pos, const float* __restrict__ g_pre, const float* __restrict__ b_pre,
                    const float* __restrict__ g_1, const float* __restrict__ b_1, __nv_bfloat16* __restrict__ x0,
                    __nv_bfloat16* __restrict__ h0, float* __restrict__ mean1, float* __restrict__ rstd1, int rows,
                    int G, float eps_pre, float eps_1) {
  constexpr int C = 256 * VPL;
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * EMB_ROWS_PER_BLOCK + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int b = row / (G + 1), l = row - b * (G + 1);
  float v[VPL][8];
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    float pe[8];
    if (l == 0) {
      emb_load8f(cls + col, v[i]);
    } else {
      emb_unpack8(__ldg(reinterpret_cast<const uint4*>(patch_emb + (static_cast<size_t>(b) * G + (l - 1)) * C + col)),
                  v[i]);
    }
    emb_load8f(pos + static_cast<size_t>(l) * C + col, pe);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[i][e] += pe[e];
  }
  float mean, rstd;
  warp_row_stats<VPL>(v, mean, rstd, eps_pre);
  const size_t base = static_cast<size_t>(row) * C;
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    float gg[8], bb[8];
    emb_load8f(g_pre + col, gg);
    emb_load8f(b_pre + col, bb);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[i][e] = fmaf((v[i][e] - mean) * rstd, gg[e], bb[e]);
    // the residual stream lives in bf16: ln_1 normalises exactly what is stored
    const uint4 packed = emb_pack8(v[i]);
    *reinterpret_cast<uint4*>(x0 + base + col) = packed;
    emb_unpack8(packed, v[i]);
  }
  warp_row_stats<VPL>(v, mean, rstd, eps_1);
  if (lane == 0) {
    if (mean1 != nullptr) mean1[row] = mean;
    if (rstd1 != nullptr) rstd1[row] = rstd;
  }
#pragma unroll
  for (int i = 0; i < VPL; ++i) {
    const int col = (i * 32 + lane) * 8;
    float gg[8], bb[8], o[8];
    emb_load8f(g_1 + col, gg);
    emb_load8f(b_1 + col, bb);
#pragma unroll
    for (int e = 0; e < 8; ++e) o[e] = fmaf((v[i][e] - mean) * rstd, gg[e], bb[e]);
    *reinterpret_cast<uint4*>(h0 + base + col) = emb_pack8(o);
  }
}

}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_patchify_normalize(const float* image, void* patches, const float* mean, const float* stdv, int Bp, int C,
                           int H, int W, int patch, int div255, cudaStream_t stream) {
  FFM_CHECK_ARG(image && patches && mean && stdv, "ffm_patchify_normalize: null pointer argument");
  FFM_CHECK_ARG(Bp >= 1 && C >= 1 && patch >= 8 && patch % 8 == 0, "ffm_patchify_normalize: patch must be a multiple of 8");
  FFM_CHECK_ARG(H % patch == 0 && W % patch == 0, "ffm_patchify_normalize: H, W must be multiples of the patch size");
  FFM_CHECK_ARG(W % 4 == 0, "ffm_patchify_normalize: W must be a multiple of 4 (128-bit loads)");
  const int gh = H / patch, gw = W / patch, G = gh * gw;
  const long long n_chunks = static_cast<long long>(Bp) * G * C * patch * patch / 8;
  const long long blocks = (n_chunks + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_patchify_normalize: too many elements");
  patchify_normalize_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      image, static_cast<__nv_bfloat16*>(patches), mean, stdv, C, H, W, patch, gw, G, n_chunks, div255);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_vit_embed_ln(const void* patch_emb, const float* class_embedding, const float* positional_embedding,
                     const float* ln_pre_gamma, const float* ln_pre_beta, const float* ln1_gamma, const float* ln1_beta,
                     void* x0, void* h0, float* mean1, float* rstd1, int Bp, int G, int C, float eps_pre, float eps_1,
                     cudaStream_t stream) {
  FFM_CHECK_ARG(patch_emb && class_embedding && positional_embedding && ln_pre_gamma && ln_pre_beta && ln1_gamma &&
                    ln1_beta && x0 && h0,
                "ffm_vit_embed_ln: null pointer argument");
  FFM_CHECK_ARG(Bp >= 1 && G >= 1, "ffm_vit_embed_ln: bad sizes");
  FFM_CHECK_ARG(C % 256 == 0 && C >= 256 && C <= 1024, "ffm_vit_embed_ln: C (%d) must be 256, 512, 768 or 1024", C);
  const int rows = Bp * (G + 1);
  const int grid = (rows + EMB_ROWS_PER_BLOCK - 1) / EMB_ROWS_PER_BLOCK;
  const __nv_bfloat16* pe = static_cast<const __nv_bfloat16*>(patch_emb);
  __nv_bfloat16* xo = static_cast<__nv_bfloat16*>(x0);
  __nv_bfloat16* ho = static_cast<__nv_bfloat16*>(h0);
#define FFM_EMB_LAUNCH(V)                                                                                         \
  vit_embed_ln_kernel<V><<<grid, EMB_THREADS, 0, stream>>>(pe, class_embedding, positional_embedding, ln_pre_gamma, \
                                                           ln_pre_beta, ln1_gamma, ln1_beta, xo, ho, mean1, rstd1, \
                                                           rows, G, eps_pre, eps_1)
  switch (C / 256) {
    case 1: FFM_EMB_LAUNCH(1); break;
    case 2: FFM_EMB_LAUNCH(2); break;
    case 3: FFM_EMB_LAUNCH(3); break;
    default: FFM_EMB_LAUNCH(4); break;
  }
#undef FFM_EMB_LAUNCH
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
