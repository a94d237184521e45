// Average pooling of the ResNet trunk (clip/model.py:30, :42, :108 — nn.AvgPool2d(stride) in the anti-aliased bottlenecks and
// after the stem) on channels-last fp32 activations: a frozen, parameter-free glue op between the adapted 1x1 convolutions.
// The library kernels for this layout run at 0.3 - 0.6 TB/s (avg_pool2d_backward_out_cuda_frame_nhwc: 3.3 ms per RN50 step,
// the forward 1.3 ms); these are plain 16-byte-vectorised streaming kernels (HBM-bound: 4 B in + 4/k^2 out per element).
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

// 16 bytes of activations = 4 fp32 or 8 bf16 channels, accumulated in fp32
template <bool BF16>
struct Vec16 {
  static constexpr int N = BF16 ? 8 : 4;
  __device__ static void load(const uint4* p, float (&v)[8]) {
    const uint4 r = __ldg(p);
    if (BF16) {
      const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
      for (int e = 0; e < 4; ++e) { const float2 f = __bfloat1622float2(h2[e]); v[2 * e] = f.x; v[2 * e + 1] = f.y; }
    } else {
      v[0] = __uint_as_float(r.x); v[1] = __uint_as_float(r.y); v[2] = __uint_as_float(r.z); v[3] = __uint_as_float(r.w);
    }
  }
  __device__ static void store(uint4* p, const float (&v)[8]) {
    uint4 r;
    if (BF16) {
      r.x = pack_bf16x2(v[0], v[1]); r.y = pack_bf16x2(v[2], v[3]); r.z = pack_bf16x2(v[4], v[5]); r.w = pack_bf16x2(v[6], v[7]);
    } else {
      r.x = __float_as_uint(v[0]); r.y = __float_as_uint(v[1]); r.z = __float_as_uint(v[2]); r.w = __float_as_uint(v[3]);
    }
    *p = r;
  }
};

// one thread = one 16-byte channel group of one OUTPUT pixel
template <bool BF16>
__global__ void __launch_bounds__(256)
avgpool_nhwc_fwd_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int Ho, int Wo, int CV, int W, int k,
                        long long n_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int c = static_cast<int>(i % CV);
  long long r = i / CV;
  const int wo = static_cast<int>(r % Wo);
  r /= Wo;
  const int ho = static_cast<int>(r % Ho);
  const long long b = r / Ho;
  const long long H = static_cast<long long>(Ho) * k;
  float s[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  for (int di = 0; di < k; ++di)
    for (int dj = 0; dj < k; ++dj) {
      float v[8];
      Vec16<BF16>::load(x + ((b * H + static_cast<long long>(ho) * k + di) * W + static_cast<long long>(wo) * k + dj) * CV + c, v);
#pragma unroll
      for (int e = 0; e < Vec16<BF16>::N; ++e) s[e] += v[e];
    }
  const float inv = 1.0f / static_cast<float>(k * k);
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] *= inv;
  Vec16<BF16>::store(y + i, s);
}

// one thread = one 16-byte channel group of one INPUT pixel: dx = dy[h / k, w / k] / k^2
template <bool BF16>
__global__ void __launch_bounds__(256)
avgpool_nhwc_bwd_kernel(const uint4* __restrict__ dy, uint4* __restrict__ dx, int H, int W, int CV, int k,
                        long long n_in) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  const int c = static_cast<int>(i % CV);
  long long r = i / CV;
  const int w = static_cast<int>(r % W);
  r /= W;
  const int h = static_cast<int>(r % H);
  const long long b = r / H;
  const int Ho = H / k, Wo = W / k;
  float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  Vec16<BF16>::load(dy + ((b * Ho + h / k) * Wo + w / k) * CV + c, v);
  const float inv = 1.0f / static_cast<float>(k * k);
#pragma unroll
  for (int e = 0; e < 8; ++e) v[e] *= inv;
  Vec16<BF16>::store(dx + i, v);
}

// bf16 -> fp32 widening of a contiguous buffer (the output side of the adapted 1x1 convolutions inside the fp32 ResNet trunk):
// 8 elements per thread, 16-byte load, two 16-byte stores.  The library's converting copy for this direction is not vectorised
// (92 us for 51 M elements against 47 us at the HBM roofline).
__global__ void __launch_bounds__(256)
widen_bf16_kernel(const uint4* __restrict__ in, float4* __restrict__ out, long long n8) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const uint4 r = __ldg(in + i);
  const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&r);
  const float2 a = __bfloat1622float2(h2[0]), b = __bfloat1622float2(h2[1]);
  const float2 c = __bfloat1622float2(h2[2]), d = __bfloat1622float2(h2[3]);
  out[2 * i] = make_float4(a.x, a.y, b.x, b.y);
  out[2 * i + 1] = make_float4(c.x, c.y, d.x, d.y);
}

// y = relu(a + b) (the closing `relu(out + identity)` of a bottleneck, clip/model.py:56-58) and its backward g = dy * [y > 0]
// (the same tensor is the gradient of both inputs): one pass each instead of add + ReLU / threshold kernels.
__global__ void __launch_bounds__(256)
add_relu_kernel(const float4* __restrict__ a, const float4* __restrict__ b, float4* __restrict__ y, long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 u = __ldg(a + i), v = __ldg(b + i);
  y[i] = make_float4(fmaxf(u.x + v.x, 0.f), fmaxf(u.y + v.y, 0.f), fmaxf(u.z + v.z, 0.f), fmaxf(u.w + v.w, 0.f));
}

__global__ void __launch_bounds__(256)
relu_mask_kernel(const float4* __restrict__ dy, const float4* __restrict__ y, float4* __restrict__ g, long long n4) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 d = __ldg(dy + i), v = __ldg(y + i);
  g[i] = make_float4(v.x > 0.f ? d.x : 0.f, v.y > 0.f ? d.y : 0.f, v.z > 0.f ? d.z : 0.f, v.w > 0.f ? d.w : 0.f);
}

}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_avgpool_nhwc_fwd(const void* x, void* y, int B, int H, int W, int C, int k, int elem_bf16, cudaStream_t stream) {
  FFM_CHECK_ARG(x && y, "ffm_avgpool_nhwc_fwd: null pointer argument");
  const int per = elem_bf16 ? 8 : 4;
  FFM_CHECK_ARG(B >= 1 && k >= 1 && H % k == 0 && W % k == 0 && C % per == 0 && H >= k && W >= k,
                "ffm_avgpool_nhwc_fwd: H, W must be multiples of k and C a multiple of 4 (fp32) / 8 (bf16)");
  const long long n_out = static_cast<long long>(B) * (H / k) * (W / k) * (C / per);
  const long long blocks = (n_out + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_avgpool_nhwc_fwd: too many elements");
  if (elem_bf16)
    avgpool_nhwc_fwd_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(y), H / k, W / k, C / per, W, k, n_out);
  else
    avgpool_nhwc_fwd_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const uint4*>(x), static_cast<uint4*>(y), H / k, W / k, C / per, W, k, n_out);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_avgpool_nhwc_bwd(const void* dy, void* dx, int B, int H, int W, int C, int k, int elem_bf16, cudaStream_t stream) {
  FFM_CHECK_ARG(dy && dx, "ffm_avgpool_nhwc_bwd: null pointer argument");
  const int per = elem_bf16 ? 8 : 4;
  FFM_CHECK_ARG(B >= 1 && k >= 1 && H % k == 0 && W % k == 0 && C % per == 0 && H >= k && W >= k,
                "ffm_avgpool_nhwc_bwd: H, W must be multiples of k and C a multiple of 4 (fp32) / 8 (bf16)");
  const long long n_in = static_cast<long long>(B) * H * W * (C / per);
  const long long blocks = (n_in + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_avgpool_nhwc_bwd: too many elements");
  if (elem_bf16)
    avgpool_nhwc_bwd_kernel<true><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const uint4*>(dy), static_cast<uint4*>(dx), H, W, C / per, k, n_in);
  else
    avgpool_nhwc_bwd_kernel<false><<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
        static_cast<const uint4*>(dy), static_cast<uint4*>(dx), H, W, C / per, k, n_in);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_add_relu(const float* a, const float* b, float* y, int64_t n, cudaStream_t stream) {
  FFM_CHECK_ARG(a && b && y, "ffm_add_relu: null pointer argument");
  FFM_CHECK_ARG(n >= 4 && n % 4 == 0, "ffm_add_relu: element count must be a positive multiple of 4");
  const long long n4 = n / 4, blocks = (n4 + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_add_relu: too many elements");
  add_relu_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<const float4*>(a),
                                                                     reinterpret_cast<const float4*>(b),
                                                                     reinterpret_cast<float4*>(y), n4);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_relu_mask(const float* dy, const float* y, float* g, int64_t n, cudaStream_t stream) {
  FFM_CHECK_ARG(dy && y && g, "ffm_relu_mask: null pointer argument");
  FFM_CHECK_ARG(n >= 4 && n % 4 == 0, "ffm_relu_mask: element count must be a positive multiple of 4");
  const long long n4 = n / 4, blocks = (n4 + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_relu_mask: too many elements");
  relu_mask_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(reinterpret_cast<const float4*>(dy),
                                                                      reinterpret_cast<const float4*>(y),
                                                                      reinterpret_cast<float4*>(g), n4);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_widen_bf16(const void* in, float* out, int64_t n, cudaStream_t stream) {
  FFM_CHECK_ARG(in && out, "ffm_widen_bf16: null pointer argument");
  FFM_CHECK_ARG(n >= 8 && n % 8 == 0, "ffm_widen_bf16: element count must be a positive multiple of 8");
  const long long n8 = n / 8, blocks = (n8 + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_widen_bf16: too many elements");
  widen_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(static_cast<const uint4*>(in),
                                                                       reinterpret_cast<float4*>(out), n8);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
