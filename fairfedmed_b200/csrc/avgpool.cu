// Average pooling of the ResNet trunk (clip/model.py:30, :42, :108 — nn.AvgPool2d(stride) in the anti-aliased bottlenecks and
// after the stem) on channels-last fp32 activations: a frozen, parameter-free glue op between the adapted 1x1 convolutions.
// The library kernels for this layout run at 0.3 - 0.6 TB/s (avg_pool2d_backward_out_cuda_frame_nhwc: 3.3 ms per RN50 step,
// the forward 1.3 ms); these are plain 16-byte-vectorised streaming kernels (HBM-bound: 4 B in + 4/k^2 out per element).
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

// one thread = 4 channels of one OUTPUT pixel
__global__ void __launch_bounds__(256)
avgpool_nhwc_fwd_kernel(const float4* __restrict__ x, float4* __restrict__ y, int Ho, int Wo, int C4, int W, int k,
                        long long n_out) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  const int c = static_cast<int>(i % C4);
  long long r = i / C4;
  const int wo = static_cast<int>(r % Wo);
  r /= Wo;
  const int ho = static_cast<int>(r % Ho);
  const long long b = r / Ho;
  const long long H = static_cast<long long>(Ho) * k;
  float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int di = 0; di < k; ++di)
    for (int dj = 0; dj < k; ++dj) {
      const float4 v = __ldg(x + ((b * H + static_cast<long long>(ho) * k + di) * W + static_cast<long long>(wo) * k + dj) * C4 + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  const float inv = 1.0f / static_cast<float>(k * k);
  y[i] = make_float4(s.x * inv, s.y * inv, s.z * inv, s.w * inv);
}

// one thread = 4 channels of one INPUT pixel: dx = dy[h / k, w / k] / k^2
__global__ void __launch_bounds__(256)
avgpool_nhwc_bwd_kernel(const float4* __restrict__ dy, float4* __restrict__ dx, int H, int W, int C4, int k,
                        long long n_in) {
  const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_in) return;
  const int c = static_cast<int>(i % C4);
  long long r = i / C4;
  const int w = static_cast<int>(r % W);
  r /= W;
  const int h = static_cast<int>(r % H);
  const long long b = r / H;
  const int Ho = H / k, Wo = W / k;
  const float4 v = __ldg(dy + ((b * Ho + h / k) * Wo + w / k) * C4 + c);
  const float inv = 1.0f / static_cast<float>(k * k);
  dx[i] = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
}

}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_avgpool_nhwc_fwd(const float* x, float* y, int B, int H, int W, int C, int k, cudaStream_t stream) {
  FFM_CHECK_ARG(x && y, "ffm_avgpool_nhwc_fwd: null pointer argument");
  FFM_CHECK_ARG(B >= 1 && k >= 1 && H % k == 0 && W % k == 0 && C % 4 == 0 && H >= k && W >= k,
                "ffm_avgpool_nhwc_fwd: H, W must be multiples of k and C a multiple of 4");
  const long long n_out = static_cast<long long>(B) * (H / k) * (W / k) * (C / 4);
  const long long blocks = (n_out + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_avgpool_nhwc_fwd: too many elements");
  avgpool_nhwc_fwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(y), H / k, W / k, C / 4, W, k, n_out);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_avgpool_nhwc_bwd(const float* dy, float* dx, int B, int H, int W, int C, int k, cudaStream_t stream) {
  FFM_CHECK_ARG(dy && dx, "ffm_avgpool_nhwc_bwd: null pointer argument");
  FFM_CHECK_ARG(B >= 1 && k >= 1 && H % k == 0 && W % k == 0 && C % 4 == 0 && H >= k && W >= k,
                "ffm_avgpool_nhwc_bwd: H, W must be multiples of k and C a multiple of 4");
  const long long n_in = static_cast<long long>(B) * H * W * (C / 4);
  const long long blocks = (n_in + 255) / 256;
  FFM_CHECK_ARG(blocks <= 0x7fffffffLL, "ffm_avgpool_nhwc_bwd: too many elements");
  avgpool_nhwc_bwd_kernel<<<static_cast<unsigned>(blocks), 256, 0, stream>>>(
      reinterpret_cast<const float4*>(dy), reinterpret_cast<float4*>(dx), H, W, C / 4, k, n_in);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
