// Federated aggregation kernels and the fused double-step SGD over flat fp32 parameter buffers.
//
// Reference: utils/fed_utils.py:42-100 (average_weights_EMA), :6-40 (average_weights);
// trainers/GLP_OT_SVLoRA.py:864-871 + Dassl/dassl/engine/trainer.py:333-342 (optimizer stepped twice).
//
// One client per rank: each rank pre-scales its flat buffer (ffm_fedavg_scale), the host mirror sums the
// buffers across ranks with one NCCL all-reduce over NVLink, and ffm_fedavg_epilogue applies shared-half-S and
// the EMA with the previous global weights.  All three are HBM-bound element-wise passes over a few MB
// (P = 1.11 M fp32 for ViT-B/16 r=12): coalesced grid-stride loops, grid capped at 4 CTAs per SM, the segment
// table (<= 1024 entries) staged in shared memory.
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int FA_THREADS = 256;
constexpr int FA_MAX_SEG = 1024;

struct SegTable {
  const int32_t* kind;
  const int64_t* off;
  const int64_t* len;
  int n;
};

// largest s with off[s] <= i  (segments are sorted, contiguous, cover [0, n_elem))
__device__ __forceinline__ int find_seg(const int64_t* __restrict__ off, int n, int64_t i) {
  int lo = 0, hi = n - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (off[mid] <= i) lo = mid; else hi = mid - 1;
  }
  return lo;
}

__global__ void __launch_bounds__(FA_THREADS)
fedavg_scale_kernel(const float* __restrict__ in, float* __restrict__ out, SegTable st, int64_t n_elem,
                    float w_scalar, const float* __restrict__ w_group, int r) {
  extern __shared__ int64_t off_s[];
  for (int i = threadIdx.x; i < st.n; i += blockDim.x) off_s[i] = st.off[i];
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
    const int s = find_seg(off_s, st.n, i);
    float w = w_scalar;
    if (st.kind[s] == 1) w = w_group[(i - off_s[s]) / r];
    out[i] = in[i] * w;
  }
}

__global__ void __launch_bounds__(FA_THREADS)
fedavg_epilogue_kernel(const float* __restrict__ avg, const float* __restrict__ prev, float* __restrict__ out,
                       SegTable st, int64_t n_elem, float beta_decay, int shared_half, int G, int r) {
  extern __shared__ int64_t off_s[];
  for (int i = threadIdx.x; i < st.n; i += blockDim.x) off_s[i] = st.off[i];
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const float one_minus = 1.0f - beta_decay;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n_elem; i += stride) {
    const int s = find_seg(off_s, st.n, i);
    float v = avg[i];
    if (shared_half && st.kind[s] == 1) {
      const int64_t rel = i - off_s[s];
      const int col = static_cast<int>(rel % r);
      if (col < r / 2) {
        // first half of the singular values is shared: mean over the G group rows (fed_utils.py:90-96)
        float m = 0.f;
        for (int g = 0; g < G; ++g) m += avg[off_s[s] + static_cast<int64_t>(g) * r + col];
        v = m / static_cast<float>(G);
      }
    }
    out[i] = one_minus * v + beta_decay * prev[i];
  }
}

// torch.optim.SGD (momentum, weight decay, dampening 0, no nesterov) applied n_steps times with one gradient.
__global__ void __launch_bounds__(FA_THREADS)
sgd_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ mom, int64_t n, float lr,
           float momentum, float wd, int n_steps, int first_step) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float w = param[i];
    const float g = grad[i];
    float b = mom[i];
    bool first = first_step != 0;
    for (int k = 0; k < n_steps; ++k) {
      const float d = fmaf(wd, w, g);          // g + wd * w
      b = first ? d : fmaf(momentum, b, d);    // buf = momentum * buf + d
      first = false;
      w = fmaf(-lr, b, w);
    }
    param[i] = w;
    mom[i] = b;
  }
}

// Same update with the learning rate read from device memory: a captured CUDA graph of the training step stays valid
// when the StepLR schedule changes the rate (the host rewrites one float instead of re-capturing).
__global__ void __launch_bounds__(FA_THREADS)
sgd_dev_lr_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ mom, int64_t n,
                  const float* __restrict__ lr_dev, float momentum, float wd, int n_steps) {
  const float lr = __ldg(lr_dev);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float w = param[i];
    const float g = grad[i];
    float b = mom[i];
    for (int k = 0; k < n_steps; ++k) {
      const float d = fmaf(wd, w, g);
      b = fmaf(momentum, b, d);
      w = fmaf(-lr, b, w);
    }
    param[i] = w;
    mom[i] = b;
  }
}

static int grid_for(int64_t n) {
  int64_t blocks = (n + FA_THREADS - 1) / FA_THREADS;
  const int64_t cap = static_cast<int64_t>(num_sms()) * 4;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_fedavg_scale(const float* flat_in, float* flat_out, const int32_t* seg_kind, const int64_t* seg_off,
                     const int64_t* seg_len, int n_seg, int64_t n_elem, float w_scalar, const float* w_group, int G,
                     int r, cudaStream_t stream) {
  FFM_CHECK_ARG(flat_in && flat_out && seg_kind && seg_off && seg_len, "ffm_fedavg_scale: null pointer argument");
  FFM_CHECK_ARG(n_seg >= 1 && n_seg <= FA_MAX_SEG && n_elem >= 1, "ffm_fedavg_scale: bad segment table");
  FFM_CHECK_ARG(r >= 1 && G >= 1, "ffm_fedavg_scale: bad G / r");
  SegTable st{seg_kind, seg_off, seg_len, n_seg};
  fedavg_scale_kernel<<<grid_for(n_elem), FA_THREADS, n_seg * sizeof(int64_t), stream>>>(flat_in, flat_out, st,
                                                                                          n_elem, w_scalar, w_group, r);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_fedavg_epilogue(const float* avg, const float* prev_global, float* out, const int32_t* seg_kind,
                        const int64_t* seg_off, const int64_t* seg_len, int n_seg, int64_t n_elem, float beta_decay,
                        int shared_half_s, int G, int r, cudaStream_t stream) {
  FFM_CHECK_ARG(avg && prev_global && out && seg_kind && seg_off && seg_len,
                "ffm_fedavg_epilogue: null pointer argument");
  FFM_CHECK_ARG(avg != out || !shared_half_s, "ffm_fedavg_epilogue: in-place is not allowed with shared_half_s");
  FFM_CHECK_ARG(n_seg >= 1 && n_seg <= FA_MAX_SEG && n_elem >= 1, "ffm_fedavg_epilogue: bad segment table");
  SegTable st{seg_kind, seg_off, seg_len, n_seg};
  fedavg_epilogue_kernel<<<grid_for(n_elem), FA_THREADS, n_seg * sizeof(int64_t), stream>>>(
      avg, prev_global, out, st, n_elem, beta_decay, shared_half_s, G, r);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                 float weight_decay, int n_steps, int first_step, cudaStream_t stream) {
  FFM_CHECK_ARG(param && grad && momentum_buf, "ffm_sgd_step: null pointer argument");
  FFM_CHECK_ARG(n >= 1 && n_steps >= 1, "ffm_sgd_step: bad sizes");
  sgd_kernel<<<grid_for(n), FA_THREADS, 0, stream>>>(param, grad, momentum_buf, n, lr, momentum, weight_decay,
                                                     n_steps, first_step);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_sgd_step_dev_lr(float* param, const float* grad, float* momentum_buf, int64_t n, const float* lr_dev,
                        float momentum, float weight_decay, int n_steps, cudaStream_t stream) {
  FFM_CHECK_ARG(param && grad && momentum_buf && lr_dev, "ffm_sgd_step_dev_lr: null pointer argument");
  FFM_CHECK_ARG(n >= 1 && n_steps >= 1, "ffm_sgd_step_dev_lr: bad sizes");
  sgd_dev_lr_kernel<<<grid_for(n), FA_THREADS, 0, stream>>>(param, grad, momentum_buf, n, lr_dev, momentum,
                                                            weight_decay, n_steps);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
