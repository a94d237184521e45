// Small (HBM / latency bound) kernels around the fused SVLoRA GEMM:
//   * group mixing      s_eff = pi(attr) · S (+ S_global)        FairLoRALinear.forward :453-467
//   * its transpose     dS    = pi^T · ds_eff                    (autograd of the line above)
//   * adapter gradients dA = x^T·dh, dB = (scaling z)^T·dy       (autograd of :477-478)
//   * per-sample        ds_eff[b] = sum_{t in sample b} scaling·dzu ⊙ h   (segmented by sample id)
// Reference lines are trainers/GLP_OT_SVLoRA.py in /root/reference.
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int RPS = 16;  // padded rank, must match svlora_gemm.cu

// ----------------------------------------------------------------------------------------------
// s_eff / dS
// ----------------------------------------------------------------------------------------------
// pi[b, g] = lambda if attr[b] == g else (1-lambda)/(G-1); attr == nullptr => one row of 1/G.
__global__ void seff_kernel(const long long* __restrict__ attr, const float* __restrict__ S,
                            const float* __restrict__ S_global, float* __restrict__ s_eff, int nS, int G, int r,
                            float lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nS * r) return;
  const int b = i / r, j = i - b * r;
  float acc = 0.f;
  if (attr == nullptr) {
    const float w = 1.0f / static_cast<float>(G);
    for (int g = 0; g < G; ++g) acc += w * S[g * r + j];
  } else {
    const int a = static_cast<int>(attr[b]);
    const float off = (1.0f - lambda) / static_cast<float>(G - 1);
    for (int g = 0; g < G; ++g) acc += ((g == a) ? lambda : off) * S[g * r + j];
  }
  if (S_global != nullptr) acc += S_global[j];
  s_eff[i] = acc;
}

__global__ void ds_kernel(const long long* __restrict__ attr, const float* __restrict__ ds_eff,
                          float* __restrict__ dS, float* __restrict__ dS_global, int nS, int G, int r,
                          float lambda) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (G + 1) * r) return;
  const int g = i / r, j = i - g * r;
  float acc = 0.f;
  if (g == G) {  // gradient of the optional global singular values: plain sum over samples
    if (dS_global == nullptr) return;
    for (int b = 0; b < nS; ++b) acc += ds_eff[b * r + j];
    dS_global[j] = acc;
    return;
  }
  if (attr == nullptr) {
    const float w = 1.0f / static_cast<float>(G);
    for (int b = 0; b < nS; ++b) acc += w * ds_eff[b * r + j];
  } else {
    const float off = (1.0f - lambda) / static_cast<float>(G - 1);
    for (int b = 0; b < nS; ++b) acc += ((static_cast<int>(attr[b]) == g) ? lambda : off) * ds_eff[b * r + j];
  }
  dS[i] = acc;
}

// ----------------------------------------------------------------------------------------------
// out[c, j] = sum_t M[t, c] * v[t, j],  v[t, j] = src[t, j] * s_rows[sample(t), j]
// bf16 M [T, C]; fp32 src [T, 16]; partial sums per row chunk, reduced by colsum_reduce_kernel.
// ----------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 128;
constexpr int CS_COLS = CS_THREADS * 2;  // columns per CTA (one bf16x2 per thread per row)
constexpr int CS_ROWS = 32;              // rows staged per smem batch

__global__ void __launch_bounds__(CS_THREADS)
colsum16_kernel(const __nv_bfloat16* __restrict__ M, const float* __restrict__ src,
                const float* __restrict__ s_rows, float* __restrict__ partial, int T, int C, int rows_per_chunk,
                int b_prime, int num_slices) {
  __shared__ __align__(16) float v_s[CS_ROWS][RPS];
  const int c0 = blockIdx.y * CS_COLS + threadIdx.x * 2;
  const int t_begin = blockIdx.x * rows_per_chunk;
  const int t_end = min(T, t_begin + rows_per_chunk);
  const bool col_ok = c0 < C;  // C is even (multiple of 8)
  float acc0[RPS], acc1[RPS];
#pragma unroll
  for (int j = 0; j < RPS; ++j) { acc0[j] = 0.f; acc1[j] = 0.f; }

  for (int tb = t_begin; tb < t_end; tb += CS_ROWS) {
    const int nrows = min(CS_ROWS, t_end - tb);
    __syncthreads();
    for (int i = threadIdx.x; i < nrows * RPS; i += CS_THREADS) {
      const int rr = i / RPS, j = i - rr * RPS;
      const int t = tb + rr;
      const int sample = (t % b_prime) / num_slices;
      v_s[rr][j] = src[static_cast<size_t>(t) * RPS + j] * s_rows[sample * RPS + j];
    }
    __syncthreads();
    if (col_ok) {
#pragma unroll 4
      for (int rr = 0; rr < nrows; ++rr) {
        const __nv_bfloat162 m2 =
            *reinterpret_cast<const __nv_bfloat162*>(M + static_cast<size_t>(tb + rr) * C + c0);
        const float2 m = __bfloat1622float2(m2);
        const float4* vp = reinterpret_cast<const float4*>(v_s[rr]);
#pragma unroll
        for (int j4 = 0; j4 < 4; ++j4) {
          const float4 v = vp[j4];
          acc0[4 * j4 + 0] = fmaf(m.x, v.x, acc0[4 * j4 + 0]);
          acc0[4 * j4 + 1] = fmaf(m.x, v.y, acc0[4 * j4 + 1]);
          acc0[4 * j4 + 2] = fmaf(m.x, v.z, acc0[4 * j4 + 2]);
          acc0[4 * j4 + 3] = fmaf(m.x, v.w, acc0[4 * j4 + 3]);
          acc1[4 * j4 + 0] = fmaf(m.y, v.x, acc1[4 * j4 + 0]);
          acc1[4 * j4 + 1] = fmaf(m.y, v.y, acc1[4 * j4 + 1]);
          acc1[4 * j4 + 2] = fmaf(m.y, v.z, acc1[4 * j4 + 2]);
          acc1[4 * j4 + 3] = fmaf(m.y, v.w, acc1[4 * j4 + 3]);
        }
      }
    }
  }
  if (col_ok) {
    float4* p0 = reinterpret_cast<float4*>(partial + (static_cast<size_t>(blockIdx.x) * C + c0) * RPS);
#pragma unroll
    for (int j4 = 0; j4 < 4; ++j4) {
      p0[j4] = make_float4(acc0[4 * j4], acc0[4 * j4 + 1], acc0[4 * j4 + 2], acc0[4 * j4 + 3]);
      p0[4 + j4] = make_float4(acc1[4 * j4], acc1[4 * j4 + 1], acc1[4 * j4 + 2], acc1[4 * j4 + 3]);
    }
  }
}

// out = sum over chunks of partial[chunk, c, j]; transposed==0: out[c*r + j] ([C, r]); ==1: out[j*C + c] ([r, C])
__global__ void colsum_reduce_kernel(const float* __restrict__ partial, float* __restrict__ out, int n_chunks, int C,
                                     int r, int transposed) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= C * RPS) return;
  const int c = i / RPS, j = i - c * RPS;
  if (j >= r) return;
  float acc = 0.f;
  for (int k = 0; k < n_chunks; ++k) acc += partial[(static_cast<size_t>(k) * C + c) * RPS + j];
  if (transposed) out[static_cast<size_t>(j) * C + c] = acc;
  else out[static_cast<size_t>(c) * r + j] = acc;
}

// ----------------------------------------------------------------------------------------------
// ds_eff[b, j] = sum_{t : sample(t) == b} dzu[t, j] * h[t, j] * scaling   (scaling folded by the caller
// through s_rows? no: s_eff itself is the variable, so the factor is the bare alpha/r)
// One warp per sample: lane = (row parity, component); the warp walks the sample's rows, which sit at
// stride b_prime in the sequence-first [L, B', .] layout, then folds the two row-parity halves.
// ----------------------------------------------------------------------------------------------
__global__ void dseff_kernel(const float* __restrict__ h, const float* __restrict__ dzu, float* __restrict__ ds_eff,
                             int T, int r, int nS, int b_prime, int num_slices, float scaling) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= nS) return;
  const int j = lane & 15, half = lane >> 4;
  const int L = T / b_prime;              // sequence positions
  const int rows = L * num_slices;        // rows belonging to this sample
  float acc = 0.f;
  for (int i = half; i < rows; i += 2) {
    const int l = i / num_slices, sl = i - l * num_slices;
    const size_t t = static_cast<size_t>(l) * b_prime + warp * num_slices + sl;
    acc = fmaf(dzu[t * RPS + j], h[t * RPS + j], acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  if (half == 0 && j < r) ds_eff[warp * r + j] = acc * scaling;
}

constexpr int CS_MAX_CHUNKS = 75;  // row chunks per column group (partials: chunks x C x 16 fp32)

size_t svlora_bwd_small_scratch_bytes(int T, int K, int N) {
  const int cmax = K > N ? K : N;
  (void)T;
  return static_cast<size_t>(CS_MAX_CHUNKS) * cmax * RPS * 4 + 256;
}

static int pick_chunks(int T, int C) {
  const int col_groups = (C + CS_COLS - 1) / CS_COLS;
  int chunks = (2 * num_sms() + col_groups - 1) / col_groups;
  if (chunks > CS_MAX_CHUNKS) chunks = CS_MAX_CHUNKS;
  const int max_chunks = (T + CS_ROWS - 1) / CS_ROWS;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  return chunks;
}

int launch_svlora_bwd_small(const __nv_bfloat16* x, const __nv_bfloat16* dy, const float* h, const float* dzu,
                            const float* s_rows, float* dA, float* dB, float* ds_eff, void* scratch,
                            size_t scratch_bytes, int T, int K, int N, int r, int nS, int b_prime, int num_slices,
                            float scaling, cudaStream_t stream) {
  FFM_CHECK_ARG(T % b_prime == 0, "svlora bwd: T (%d) must be a multiple of b_prime (%d)", T, b_prime);
  float* partial = static_cast<float*>(scratch);
  // dA[K, r] = x^T · (dzu ⊙ s_rows)
  {
    const int chunks = pick_chunks(T, K);
    FFM_CHECK_ARG(static_cast<size_t>(chunks) * K * RPS * 4 <= scratch_bytes, "svlora bwd: scratch too small (dA)");
    const int rows_per_chunk = (T + chunks - 1) / chunks;
    dim3 grid(chunks, (K + CS_COLS - 1) / CS_COLS);
    colsum16_kernel<<<grid, CS_THREADS, 0, stream>>>(x, dzu, s_rows, partial, T, K, rows_per_chunk, b_prime,
                                                     num_slices);
    colsum_reduce_kernel<<<(K * RPS + 255) / 256, 256, 0, stream>>>(partial, dA, chunks, K, r, 0);
  }
  // dB[r, N] = (h ⊙ s_rows)^T · dy
  {
    const int chunks = pick_chunks(T, N);
    FFM_CHECK_ARG(static_cast<size_t>(chunks) * N * RPS * 4 <= scratch_bytes, "svlora bwd: scratch too small (dB)");
    const int rows_per_chunk = (T + chunks - 1) / chunks;
    dim3 grid(chunks, (N + CS_COLS - 1) / CS_COLS);
    colsum16_kernel<<<grid, CS_THREADS, 0, stream>>>(dy, h, s_rows, partial, T, N, rows_per_chunk, b_prime,
                                                     num_slices);
    colsum_reduce_kernel<<<(N * RPS + 255) / 256, 256, 0, stream>>>(partial, dB, chunks, N, r, 1);
  }
  count_launch(4);
  // ds_eff[nS, r]: one warp per sample
  dseff_kernel<<<(nS * 32 + 127) / 128, 128, 0, stream>>>(h, dzu, ds_eff, T, r, nS, b_prime, num_slices, scaling);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_seff(const long long* attr, const float* S, const float* S_global, float* s_eff, int n_samples, int G, int r,
             float lambda, cudaStream_t stream) {
  FFM_CHECK_ARG(S && s_eff, "ffm_seff: null pointer argument");
  FFM_CHECK_ARG(n_samples >= 1 && G >= 1 && r >= 1, "ffm_seff: bad sizes");
  FFM_CHECK_ARG(attr == nullptr || G >= 2, "ffm_seff: attribute mixing needs G >= 2 (reference divides by G-1)");
  const int n = n_samples * r;
  seff_kernel<<<(n + 127) / 128, 128, 0, stream>>>(attr, S, S_global, s_eff, n_samples, G, r, lambda);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_ds(const long long* attr, const float* ds_eff, float* dS, float* dS_global, int n_samples, int G, int r,
           float lambda, cudaStream_t stream) {
  FFM_CHECK_ARG(ds_eff && dS, "ffm_ds: null pointer argument");
  FFM_CHECK_ARG(n_samples >= 1 && G >= 1 && r >= 1, "ffm_ds: bad sizes");
  const int n = (G + 1) * r;
  ds_kernel<<<(n + 127) / 128, 128, 0, stream>>>(attr, ds_eff, dS, dS_global, n_samples, G, r, lambda);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
