// OCT slice projection (scope row f3): proj_per_3d_slice = Conv2d(dim_per_3d_slice -> 3, kernel 5, padding 2) applied to
// image / 255 (trainers/GLP_OT_SVLoRA.py:587-595, :684), the only trainable tensor in front of the ViT.  The library path
// (cuDNN) spends 7.3 ms per step in its weight-gradient kernel for this 600-weight convolution at the config-3 shape
// (256 slice-images of 8 x 224 x 224) plus 1.3 ms forward and 1 ms of layout transposes — 27 % of the OCT step.  Here:
//   forward : one block per 32 x 32 output tile; the input halo tile [Cin][36][36] sits in shared memory, a thread computes
//             4 horizontally adjacent pixels of all output channels (two 16-byte reads of the input row and five broadcast
//             reads of the weights per (c, i) for 60 FMAs) — FMA-bound, 15.4 GFLOP at config 3.
//   wgrad   : the same tiles, persistent blocks; a thread owns the five taps (c, i, 0..4) of every output channel and a group of
//             the tile's rows: one new 16-byte read of its input row and the broadcast dy quads per 60 FMAs; per-(block, row
//             group) partials are folded in order by a second kernel (deterministic, no atomics).  The input needs no
//             gradient (it is data).
// The 1/255 of `image / 255` is folded into the weights on the way into shared memory (in_scale) and into dW on the way out.
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int OC_TILE = 32;                 // output tile edge
constexpr int OC_HALO = OC_TILE + 4;        // 36 input rows / columns
constexpr int OC_ROW = 40;                  // padded row stride (floats): 16-byte aligned quads
constexpr int OC_CH = OC_HALO * OC_ROW + 4; // channel stride: +4 spreads the channels over the banks
constexpr int OC_MAX_COUT = 4;

__device__ __forceinline__ void oc_load_halo(float* x_s, const float* __restrict__ xb, int Cin, int H, int W, int h0, int w0,
                                             int tid, int nthreads) {
  const int per_c = OC_HALO * OC_HALO;
  for (int e = tid; e < Cin * per_c; e += nthreads) {
    const int c = e / per_c, r = e - c * per_c;
    const int yy = r / OC_HALO, xx = r - yy * OC_HALO;
    const int h = h0 - 2 + yy, w = w0 - 2 + xx;
    float v = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) v = __ldg(xb + (static_cast<size_t>(c) * H + h) * W + w);
    x_s[c * OC_CH + yy * OC_ROW + xx] = v;
  }
}

// y[b, o, h, w] = bias[o] + in_scale * sum_{c, i, j} w[o, c, i, j] x[b, c, h + i - 2, w + j - 2]
__global__ void __launch_bounds__(256)
oct_conv_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wgt, const float* __restrict__ bias,
                    float* __restrict__ y, int Cin, int Cout, int H, int W, int tiles_x, float in_scale) {
  extern __shared__ __align__(16) float oc_smem[];
  float* x_s = oc_smem;                                   // [Cin][36][40] (+4 per channel)
  float* w_s = oc_smem + Cin * OC_CH;                     // [Cin * 25][4]: output channel fastest, zero padded
  const int b = blockIdx.y;
  const int ty_t = blockIdx.x / tiles_x, tx_t = blockIdx.x - ty_t * tiles_x;
  const int h0 = ty_t * OC_TILE, w0 = tx_t * OC_TILE;
  for (int e = threadIdx.x; e < Cin * 25 * OC_MAX_COUT; e += 256) {
    const int k = e >> 2, o = e & 3;
    const int c = k / 25, ij = k - c * 25;
    w_s[e] = o < Cout ? in_scale * __ldg(wgt + (static_cast<size_t>(o) * Cin + c) * 25 + ij) : 0.f;
  }
  oc_load_halo(x_s, x + static_cast<size_t>(b) * Cin * H * W, Cin, H, W, h0, w0, threadIdx.x, 256);
  __syncthreads();
  const int ty = threadIdx.x >> 3, tx4 = (threadIdx.x & 7) << 2;
  float acc[OC_MAX_COUT][4];
#pragma unroll
  for (int o = 0; o < OC_MAX_COUT; ++o) {
    const float bo = o < Cout ? __ldg(bias + o) : 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) acc[o][p] = bo;
  }
  for (int c = 0; c < Cin; ++c) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const float4 a0 = *reinterpret_cast<const float4*>(x_s + c * OC_CH + (ty + i) * OC_ROW + tx4);
      const float4 a1 = *reinterpret_cast<const float4*>(x_s + c * OC_CH + (ty + i) * OC_ROW + tx4 + 4);
      const float xr[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const float4 wv = *reinterpret_cast<const float4*>(w_s + ((c * 25 + i * 5 + j) << 2));
        const float wo[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
        for (int o = 0; o < OC_MAX_COUT; ++o)
#pragma unroll
          for (int p = 0; p < 4; ++p) acc[o][p] = fmaf(wo[o], xr[p + j], acc[o][p]);
      }
    }
  }
  const int h = h0 + ty, w = w0 + tx4;
  if (h < H && w < W) {
    for (int o = 0; o < Cout; ++o) {
      float* dst = y + ((static_cast<size_t>(b) * Cout + o) * H + h) * W + w;
      if (w + 3 < W) {
        *reinterpret_cast<float4*>(dst) = make_float4(acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
      } else {
        for (int p = 0; p < 4 && w + p < W; ++p) dst[p] = acc[o][p];
      }
    }
  }
}

// partial[set][k][o]: k < Cin * 25 = weight tap (c, i, j), k == Cin * 25 = bias; set = (block, row group).
// A thread owns the five taps (c, i, 0..4) of every output channel — 15 accumulators — and one group of the tile's rows
// (rows g, g + G, ...): per 4 pixels it reads ONE new 16-byte piece of its input row (sliding 8-value window) and the dy quads
// (broadcast within the group) for 60 FMAs; `units` = Cin * 5 tap rows + 1 bias unit per group, G groups per block.
template <int COUT>
__global__ void __launch_bounds__(256)
oct_conv_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ part, int Bp, int Cin,
                      int H, int W, int tiles_x, int tiles_y, int G) {
  extern __shared__ __align__(16) float oc_smem[];
  float* x_s = oc_smem;                                   // [Cin][36][40] (+4 per channel)
  float* dy_s = oc_smem + Cin * OC_CH;                    // [COUT][32][32], zero beyond the image
  const int units = Cin * 5 + 1;
  const int g = threadIdx.x / units, u = threadIdx.x - g * units;
  const bool active = g < G;
  const bool is_bias = u == units - 1;
  const int c = u / 5, i = u - c * 5;
  const float* xrow = x_s + (is_bias ? 0 : c * OC_CH + i * OC_ROW);
  float acc[5][COUT];
#pragma unroll
  for (int j = 0; j < 5; ++j)
#pragma unroll
    for (int o = 0; o < COUT; ++o) acc[j][o] = 0.f;
  const int tiles = tiles_x * tiles_y;
  const long long total = static_cast<long long>(Bp) * tiles;
  for (long long t = blockIdx.x; t < total; t += gridDim.x) {
    const int b = static_cast<int>(t / tiles), tt = static_cast<int>(t - static_cast<long long>(b) * tiles);
    const int ty_t = tt / tiles_x, tx_t = tt - ty_t * tiles_x;
    const int h0 = ty_t * OC_TILE, w0 = tx_t * OC_TILE;
    __syncthreads();                                      // the previous tile is consumed
    oc_load_halo(x_s, x + static_cast<size_t>(b) * Cin * H * W, Cin, H, W, h0, w0, threadIdx.x, blockDim.x);
    for (int e = threadIdx.x; e < COUT * OC_TILE * OC_TILE; e += blockDim.x) {
      const int o = e >> 10, r = e & 1023;
      const int h = h0 + (r >> 5), w = w0 + (r & 31);
      dy_s[e] = (h < H && w < W) ? __ldg(dy + ((static_cast<size_t>(b) * COUT + o) * H + h) * W + w) : 0.f;
    }
    __syncthreads();
    if (!active) continue;
    for (int py = g; py < OC_TILE; py += G) {
      const float* xr = xrow + py * OC_ROW;
      float4 lo = is_bias ? make_float4(1.f, 1.f, 1.f, 1.f) : *reinterpret_cast<const float4*>(xr);
#pragma unroll
      for (int q = 0; q < OC_TILE / 4; ++q) {
        const float4 hi = is_bias ? make_float4(0.f, 0.f, 0.f, 0.f) : *reinterpret_cast<const float4*>(xr + 4 * q + 4);
        const float w8[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
#pragma unroll
        for (int o = 0; o < COUT; ++o) {
          const float4 d = *reinterpret_cast<const float4*>(dy_s + (o << 10) + py * OC_TILE + 4 * q);
#pragma unroll
          for (int j = 0; j < 5; ++j) {
            acc[j][o] = fmaf(d.x, w8[j], acc[j][o]);
            acc[j][o] = fmaf(d.y, w8[j + 1], acc[j][o]);
            acc[j][o] = fmaf(d.z, w8[j + 2], acc[j][o]);
            acc[j][o] = fmaf(d.w, w8[j + 3], acc[j][o]);
          }
        }
        lo = is_bias ? lo : hi;
      }
    }
  }
  if (active) {
    const int K = Cin * 25;
    float* base = part + (static_cast<size_t>(blockIdx.x) * G + g) * (K + 1) * OC_MAX_COUT;
    if (is_bias) {
#pragma unroll
      for (int o = 0; o < OC_MAX_COUT; ++o) base[K * OC_MAX_COUT + o] = o < COUT ? acc[0][o] : 0.f;   // tap j = 0 saw the ones
    } else {
#pragma unroll
      for (int j = 0; j < 5; ++j)
#pragma unroll
        for (int o = 0; o < OC_MAX_COUT; ++o) base[(u * 5 + j) * OC_MAX_COUT + o] = o < COUT ? acc[j][o] : 0.f;
    }
  }
}

__global__ void oct_conv_wgrad_fold_kernel(const float* __restrict__ part, float* __restrict__ dw, float* __restrict__ dbias,
                                           int nblocks, int Cin, int Cout, float in_scale) {
  const int K = Cin * 25;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;      // (k, o)
  if (e >= (K + 1) * Cout) return;
  const int k = e / Cout, o = e - k * Cout;
  float s = 0.f;
  for (int b = 0; b < nblocks; ++b) s += part[(static_cast<size_t>(b) * (K + 1) + k) * OC_MAX_COUT + o];
  if (k < K) {
    const int c = k / 25, ij = k - c * 25;
    dw[(static_cast<size_t>(o) * Cin + c) * 25 + ij] = in_scale * s;
  } else {
    dbias[o] = s;
  }
}

static int oc_wgrad_blocks() { return 2 * num_sms(); }
static int oc_wgrad_groups(int Cin) {
  const int g = 256 / (Cin * 5 + 1);
  return g < 1 ? 1 : (g > 8 ? 8 : g);
}

}  // namespace ffm

using namespace ffm;

extern "C" {

size_t ffm_oct_slice_conv_wgrad_ws_bytes(int Cin) {
  const int cin = Cin > 0 ? Cin : 1;
  return static_cast<size_t>(oc_wgrad_blocks()) * oc_wgrad_groups(cin) * (static_cast<size_t>(cin) * 25 + 1) * OC_MAX_COUT *
         sizeof(float);
}

int ffm_oct_slice_conv_fwd(const float* x, const float* w, const float* bias, float* y, int Bp, int Cin, int Cout, int H,
                           int W, float in_scale, cudaStream_t stream) {
  FFM_CHECK_ARG(x && w && bias && y, "ffm_oct_slice_conv_fwd: null pointer argument");
  FFM_CHECK_ARG(Bp >= 1 && Bp <= 65535 && Cin >= 1 && Cin <= 32 && Cout >= 1 && Cout <= OC_MAX_COUT && H >= 1 && W >= 4 &&
                    W % 4 == 0,
                "ffm_oct_slice_conv_fwd: Cin <= 32, Cout <= 4, W a multiple of 4, Bp <= 65535");
  const int tiles_x = (W + OC_TILE - 1) / OC_TILE, tiles_y = (H + OC_TILE - 1) / OC_TILE;
  const size_t smem = (static_cast<size_t>(Cin) * OC_CH + static_cast<size_t>(Cin) * 25 * OC_MAX_COUT) * sizeof(float);
  static thread_local size_t attr_smem = 0;
  if (smem > attr_smem) {
    FFM_CHECK_CUDA(cudaFuncSetAttribute(oct_conv_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    attr_smem = smem;
  }
  oct_conv_fwd_kernel<<<dim3(tiles_x * tiles_y, Bp), 256, smem, stream>>>(x, w, bias, y, Cin, Cout, H, W, tiles_x, in_scale);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_oct_slice_conv_wgrad(const float* x, const float* dy, float* dw, float* dbias, void* ws, size_t ws_bytes, int Bp,
                             int Cin, int Cout, int H, int W, float in_scale, cudaStream_t stream) {
  FFM_CHECK_ARG(x && dy && dw && dbias && ws, "ffm_oct_slice_conv_wgrad: null pointer argument");
  FFM_CHECK_ARG(Bp >= 1 && Cin >= 1 && Cin <= 32 && Cout >= 1 && Cout <= OC_MAX_COUT && H >= 1 && W >= 1,
                "ffm_oct_slice_conv_wgrad: Cin <= 32, Cout <= 4");
  static_assert(32 * 5 + 1 <= 256, "one row group of tap units must fit a 256-thread block");
  FFM_CHECK_ARG(ws_bytes >= ffm_oct_slice_conv_wgrad_ws_bytes(Cin), "ffm_oct_slice_conv_wgrad: workspace too small");
  const int tiles_x = (W + OC_TILE - 1) / OC_TILE, tiles_y = (H + OC_TILE - 1) / OC_TILE;
  const long long total = static_cast<long long>(Bp) * tiles_x * tiles_y;
  const int nblocks = static_cast<int>(total < oc_wgrad_blocks() ? total : oc_wgrad_blocks());
  const int G = oc_wgrad_groups(Cin);
  const size_t smem = (static_cast<size_t>(Cin) * OC_CH + static_cast<size_t>(Cout) * OC_TILE * OC_TILE) * sizeof(float);
  float* part = static_cast<float*>(ws);
#define FFM_OC_WGRAD(CO)                                                                                                    \
  do {                                                                                                                     \
    static thread_local size_t attr_smem = 0;                                                                              \
    if (smem > attr_smem) {                                                                                                \
      FFM_CHECK_CUDA(cudaFuncSetAttribute(oct_conv_wgrad_kernel<CO>, cudaFuncAttributeMaxDynamicSharedMemorySize,           \
                                          static_cast<int>(smem)));                                                        \
      attr_smem = smem;                                                                                                    \
    }                                                                                                                      \
    oct_conv_wgrad_kernel<CO><<<nblocks, 256, smem, stream>>>(x, dy, part, Bp, Cin, H, W, tiles_x, tiles_y, G);             \
  } while (0)
  switch (Cout) {
    case 1: FFM_OC_WGRAD(1); break;
    case 2: FFM_OC_WGRAD(2); break;
    case 3: FFM_OC_WGRAD(3); break;
    default: FFM_OC_WGRAD(4); break;
  }
#undef FFM_OC_WGRAD
  FFM_CHECK_CUDA(cudaGetLastError());
  const int n_out = (Cin * 25 + 1) * Cout;
  oct_conv_wgrad_fold_kernel<<<(n_out + 127) / 128, 128, 0, stream>>>(part, dw, dbias, nblocks * G, Cin, Cout, in_scale);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
  return FFM_OK;
}

}  // extern "C"
