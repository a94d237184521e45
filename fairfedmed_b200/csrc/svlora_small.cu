// Small (HBM / latency bound) kernels around the fused SVLoRA GEMM:
//   * group mixing      s_eff = pi(attr) · S (+ S_global)        FairLoRALinear.forward :453-467
//   * its transpose     dS    = pi^T · ds_eff                    (autograd of the line above)
//   * adapter gradients dA = x^T·dh, dB = (scaling z)^T·dy       (autograd of :477-478)
//   * per-sample        ds_eff[b] = sum_{t in sample b} scaling·dzu ⊙ h   (segmented by sample id)
// Reference lines are trainers/GLP_OT_SVLoRA.py in /root/reference.
#include <stdlib.h>

#include "../../include/ffm_b200.h"
#include "svlora_gemm.cuh"

namespace ffm {

constexpr int RPS_MAX = 32;  // largest padded rank (svlora_gemm.cuh RP_MAX); kernels take the actual one (16 or 32)

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

// dS[g, j] = sum_b pi[b, g] ds_eff[b, j] (+ the plain sum for the optional global singular values).  8 threads per
// output element split the samples, a shuffle folds them (fixed order: deterministic).
constexpr int DS_SPLIT = 8;
__global__ void ds_kernel(const long long* __restrict__ attr, const float* __restrict__ ds_eff,
                          float* __restrict__ dS, float* __restrict__ dS_global, int nS, int G, int r,
                          float lambda) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = tid / DS_SPLIT, part = tid % DS_SPLIT;
  const bool valid = i < (G + 1) * r;
  const int g = valid ? i / r : 0, j = valid ? i - g * r : 0;
  float acc = 0.f;
  if (valid) {
    if (g == G) {  // gradient of the optional global singular values: plain sum over samples
      if (dS_global != nullptr)
        for (int b = part; b < nS; b += DS_SPLIT) acc += ds_eff[b * r + j];
    } else if (attr == nullptr) {
      const float w = 1.0f / static_cast<float>(G);
      for (int b = part; b < nS; b += DS_SPLIT) acc += w * ds_eff[b * r + j];
    } else {
      const float off = (1.0f - lambda) / static_cast<float>(G - 1);
      for (int b = part; b < nS; b += DS_SPLIT)
        acc += ((static_cast<int>(attr[b]) == g) ? lambda : off) * ds_eff[b * r + j];
    }
  }
#pragma unroll
  for (int o = DS_SPLIT / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (!valid || part != 0) return;
  if (g == G) {
    if (dS_global != nullptr) dS_global[j] = acc;
  } else {
    dS[i] = acc;
  }
}

// ----------------------------------------------------------------------------------------------
// Adapter gradients: two skinny contractions over the T rows in ONE launch,
//     dA[K, r]   = x^T  · dh     (dh = bf16(scaling · dzu ⊙ s_eff[sample]), side output of the dX GEMM)
//     dB[r, N]^T = dy^T · z      (z  = bf16(scaling · h   ⊙ s_eff[sample]), side output of the forward GEMM)
// i.e. out[c, j] = sum_t M[t, c] * v[t, j] with bf16 M [T, C] and bf16 v [T, 16].  M is read from HBM exactly once
// (T*C*2 bytes) and that is the bound.  CUDA-core FMAs cannot keep up (16 FMA per 2 bytes), so the products run on
// the legacy tensor path: mma.sync m16n8k16 with A = M^T and B = v, both fetched from shared memory with
// ldmatrix.trans (the tiles sit in smem exactly as they sit in HBM: rows = t).  Both tiles arrive through cp.async
// (nothing in the loop waits on a synchronous global load).
//   CTA = 8 warps, tile = 128 columns x 64 rows per stage, 3-stage ring (3 CTAs per SM); warp w owns columns 16w..16w+15 and all
//   16 ranks (two n-tiles).  blockIdx.y < groups_a: column group of x (-> dA), else of dy (-> dB).
//   Row chunks write fp32 partials; adapter_grad_finalize_kernel folds them in a fixed order (deterministic).
// ----------------------------------------------------------------------------------------------
constexpr int CS_THREADS = 256;
constexpr int CS_COLS = 128;             // columns per CTA
constexpr int CS_ROWS = 64;              // rows per stage
constexpr int CS_STAGES = 3;
constexpr int CS_MSTRIDE = CS_COLS * 2 + 16;   // 272 B: rows 16 B apart mod 128 -> conflict-free ldmatrix
template <int R> struct CsCfg {
  static constexpr int VSTRIDE = R * 2 + 16;     // operand rows padded by 16 B -> conflict-free ldmatrix
  static constexpr int STAGE_BYTES = CS_ROWS * CS_MSTRIDE + CS_ROWS * VSTRIDE;   // 20480 (R = 16) / 22528 (R = 32)
  static constexpr int SMEM_BYTES = CS_STAGES * STAGE_BYTES;                      // 61440 / 67584: three CTAs per SM
};

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma_bf16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <int R>
__global__ void __launch_bounds__(CS_THREADS)
adapter_grad_kernel(const __nv_bfloat16* __restrict__ Ma, const __nv_bfloat16* __restrict__ va,
                    float* __restrict__ partial_a, int Ca, int groups_a, const __nv_bfloat16* __restrict__ Mb,
                    const __nv_bfloat16* __restrict__ vb, float* __restrict__ partial_b, int Cb, int T,
                    int rows_per_chunk) {
  constexpr int CS_VSTRIDE = CsCfg<R>::VSTRIDE, CS_STAGE_BYTES = CsCfg<R>::STAGE_BYTES;
  constexpr int NT = R / 8;                // 8-rank n-tiles of the m16n8k16 MMA
  extern __shared__ __align__(16) uint8_t cs_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_a = static_cast<int>(blockIdx.y) < groups_a;
  const __nv_bfloat16* __restrict__ M = is_a ? Ma : Mb;
  const __nv_bfloat16* __restrict__ v = is_a ? va : vb;
  float* __restrict__ partial = is_a ? partial_a : partial_b;
  const int C = is_a ? Ca : Cb;
  const int c_base = (is_a ? blockIdx.y : blockIdx.y - groups_a) * CS_COLS;
  const int t_begin = blockIdx.x * rows_per_chunk;
  const int t_end = min(T, t_begin + rows_per_chunk);
  const int n_stages = (t_end - t_begin + CS_ROWS - 1) / CS_ROWS;

  // stage loader: M tile (16 B = 8 columns per request) and v tile (R/8 x 16 B per row), all cp.async
  auto load_stage = [&](int st_idx, int buf) {
    uint8_t* mt = cs_smem + buf * CS_STAGE_BYTES;
    uint8_t* vt = mt + CS_ROWS * CS_MSTRIDE;
    const int t0 = t_begin + st_idx * CS_ROWS;
#pragma unroll
    for (int k = 0; k < (CS_ROWS * (CS_COLS / 8)) / CS_THREADS; ++k) {     // 4 requests per thread
      const int idx = k * CS_THREADS + threadIdx.x;
      const int rr = idx / (CS_COLS / 8), ch = idx - rr * (CS_COLS / 8);
      const int t = t0 + rr, c = c_base + ch * 8;
      uint8_t* dst = mt + rr * CS_MSTRIDE + ch * 16;
      if (t < t_end && c < C) cp_async16(dst, M + static_cast<size_t>(t) * C + c);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
    if (threadIdx.x < CS_ROWS * NT) {
      const int rr = threadIdx.x / NT, hf = threadIdx.x % NT;
      const int t = t0 + rr;
      uint8_t* dst = vt + rr * CS_VSTRIDE + hf * 16;
      if (t < t_end) cp_async16(dst, v + static_cast<size_t>(t) * R + hf * 8);
      else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  };

  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;

  // prologue
#pragma unroll
  for (int s0 = 0; s0 < CS_STAGES - 1; ++s0) {
    if (s0 < n_stages) load_stage(s0, s0);
    cp_async_commit();
  }
  // ldmatrix lane roles: matrix id = lane / 8, row inside the 8x8 block = lane % 8
  const int mi = lane >> 3, r8 = lane & 7;
  const uint32_t a_lane_off = (r8 + 8 * (mi >> 1)) * CS_MSTRIDE + (warp * 16 + 8 * (mi & 1)) * 2;
  const uint32_t b_lane_off = (r8 + 8 * (mi & 1)) * CS_VSTRIDE + (8 * (mi >> 1)) * 2;

  for (int st_idx = 0; st_idx < n_stages; ++st_idx) {
    cp_async_wait<CS_STAGES - 2>();
    __syncthreads();                       // stage st_idx landed (cp.async + plain stores); previous compute done
    {
      const int nxt = st_idx + CS_STAGES - 1;
      if (nxt < n_stages) load_stage(nxt, nxt % CS_STAGES);
      cp_async_commit();
    }
    const uint32_t mt = smem_u32(cs_smem + (st_idx % CS_STAGES) * CS_STAGE_BYTES);
    const uint32_t vt = mt + CS_ROWS * CS_MSTRIDE;
#pragma unroll
    for (int ks = 0; ks < CS_ROWS / 16; ++ks) {
      uint32_t a[4];
      ldmatrix_x4_trans(mt + ks * 16 * CS_MSTRIDE + a_lane_off, a);   // A = M^T : 16 columns x 16 rows(t)
#pragma unroll
      for (int g16 = 0; g16 < R / 16; ++g16) {
        uint32_t b[4];
        ldmatrix_x4_trans(vt + ks * 16 * CS_VSTRIDE + b_lane_off + g16 * 32, b);   // B = v : 16 rows(t) x 16 ranks
        mma_bf16_16816(acc[2 * g16], a, b[0], b[1]);                  // ranks 16*g16 + 0..7
        mma_bf16_16816(acc[2 * g16 + 1], a, b[2], b[3]);              // ranks 16*g16 + 8..15
      }
    }
  }
  cp_async_wait<0>();

  // D fragment: (row g, cols 2i,2i+1) and (row g+8, cols 2i,2i+1), g = lane/4, i = lane%4; row = column of M
  const int g = lane >> 2, i2 = (lane & 3) * 2;
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    const int c = c_base + warp * 16 + g + 8 * hrow;
    if (c < C) {
      float* p0 = partial + (static_cast<size_t>(blockIdx.x) * C + c) * R;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<float2*>(p0 + 8 * n + i2) = make_float2(acc[n][2 * hrow], acc[n][2 * hrow + 1]);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// TMA build of the same contraction (default).  ncu on the cp.async build: LDGSTS writes shared memory sector by sector
// (6-way excess wavefronts), a third of the warp stalls sit on the stage barrier, DRAM 54 % busy.  Here one elected
// thread issues three bulk tensor copies per stage (two SW128 boxes of 64 rows x 64 columns of M, one [64 x R] box of
// v) into a 4-stage ring guarded by mbarriers; consumers read the swizzled boxes with the same ldmatrix.trans pattern.
// Row chunks are multiples of 16 rows, so a chunk ends on a k-step boundary and rows past the matrix are zero-filled
// by TMA: no element-wise clipping anywhere.
// ----------------------------------------------------------------------------------------------
constexpr int CT_STAGES = 4;
template <int R> struct CtCfg {
  static constexpr int M_BYTES = CS_ROWS * CS_COLS * 2;              // 16384: two 8 KB SW128 boxes
  static constexpr int V_BYTES = CS_ROWS * R * 2;                    // 2048 / 4096, rows of 2R bytes, no swizzle
  static constexpr int STAGE_BYTES = ((M_BYTES + V_BYTES + 1023) / 1024) * 1024;
  static constexpr int SMEM_BYTES = CT_STAGES * STAGE_BYTES + 1024 + 64;   // + alignment slack + barriers
};

template <int R>
__global__ void __launch_bounds__(CS_THREADS)
adapter_grad_tma_kernel(const __grid_constant__ CUtensorMap tm_ma, const __grid_constant__ CUtensorMap tm_va,
                        const __grid_constant__ CUtensorMap tm_mb, const __grid_constant__ CUtensorMap tm_vb,
                        float* __restrict__ partial_a, int Ca, int groups_a, float* __restrict__ partial_b, int Cb,
                        int T, int rows_per_chunk) {
  using Cfg = CtCfg<R>;
  constexpr int NT = R / 8;
  extern __shared__ uint8_t ct_smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ct_smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + CT_STAGES * Cfg::STAGE_BYTES);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const bool is_a = static_cast<int>(blockIdx.y) < groups_a;
  const CUtensorMap* tm_m = is_a ? &tm_ma : &tm_mb;
  const CUtensorMap* tm_v = is_a ? &tm_va : &tm_vb;
  float* __restrict__ partial = is_a ? partial_a : partial_b;
  const int C = is_a ? Ca : Cb;
  const int c_base = (is_a ? blockIdx.y : blockIdx.y - groups_a) * CS_COLS;
  const int t_begin = blockIdx.x * rows_per_chunk;
  const int t_end = min(T, t_begin + rows_per_chunk);
  const int n_stages = (t_end - t_begin + CS_ROWS - 1) / CS_ROWS;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(tm_m);
    tma_prefetch_desc(tm_v);
    for (int i = 0; i < CT_STAGES; ++i) mbar_init(&full_bar[i], 1);
    fence_mbar_init();
  }
  __syncthreads();

  auto issue = [&](int st_idx) {          // one thread: arm the barrier and start the three copies of a stage
    const int buf = st_idx % CT_STAGES;
    uint8_t* mt = smem + buf * Cfg::STAGE_BYTES;
    const int t0 = t_begin + st_idx * CS_ROWS;
    mbar_arrive_expect_tx(&full_bar[buf], Cfg::M_BYTES + Cfg::V_BYTES);
    tma_load_2d(mt, tm_m, &full_bar[buf], c_base, t0);
    tma_load_2d(mt + Cfg::M_BYTES / 2, tm_m, &full_bar[buf], c_base + 64, t0);
    tma_load_2d(mt + Cfg::M_BYTES, tm_v, &full_bar[buf], 0, t0);
  };
  if (threadIdx.x == 0)
    for (int s0 = 0; s0 < CT_STAGES && s0 < n_stages; ++s0) issue(s0);

  float acc[NT][4];
#pragma unroll
  for (int n = 0; n < NT; ++n)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[n][e] = 0.f;

  // ldmatrix lane roles: matrix id = lane / 8, row inside the 8x8 block = lane % 8.  A = M^T from the SW128 box of this
  // warp (box = warp / 4, 16 columns = two 16-byte chunks of the 128-byte row, chunk index XOR (row & 7)); B = v.
  const int mi = lane >> 3, r8 = lane & 7;
  const uint32_t wl = warp & 3;
  const uint32_t a_lane_off = (warp >> 2) * (Cfg::M_BYTES / 2) + (r8 + 8 * (mi >> 1)) * 128 +
                              (((2 * wl + (mi & 1)) ^ static_cast<uint32_t>(r8)) << 4);
  const uint32_t b_lane_off = Cfg::M_BYTES + (r8 + 8 * (mi & 1)) * (2 * R) + (8 * (mi >> 1)) * 2;

  for (int st_idx = 0; st_idx < n_stages; ++st_idx) {
    const int buf = st_idx % CT_STAGES;
    mbar_wait(&full_bar[buf], (st_idx / CT_STAGES) & 1, 900 + buf);
    const uint32_t st = smem_u32(smem + buf * Cfg::STAGE_BYTES);
    const int rows_valid = min(CS_ROWS, t_end - (t_begin + st_idx * CS_ROWS));   // multiple of 16 unless it ends at T
#pragma unroll
    for (int ks = 0; ks < CS_ROWS / 16; ++ks) {
      if (ks * 16 < rows_valid) {           // warp-uniform: rows past the chunk belong to the next CTA
        uint32_t a[4];
        ldmatrix_x4_trans(st + ks * 16 * 128 + a_lane_off, a);          // A = M^T : 16 columns x 16 rows(t)
#pragma unroll
        for (int g16 = 0; g16 < R / 16; ++g16) {
          uint32_t b[4];
          ldmatrix_x4_trans(st + ks * 16 * (2 * R) + b_lane_off + g16 * 32, b);   // B = v : 16 rows(t) x 16 ranks
          mma_bf16_16816(acc[2 * g16], a, b[0], b[1]);
          mma_bf16_16816(acc[2 * g16 + 1], a, b[2], b[3]);
        }
      }
    }
    __syncthreads();                        // every warp is done with this buffer: refill it
    if (threadIdx.x == 0 && st_idx + CT_STAGES < n_stages) issue(st_idx + CT_STAGES);
  }

  const int g = lane >> 2, i2 = (lane & 3) * 2;
#pragma unroll
  for (int hrow = 0; hrow < 2; ++hrow) {
    const int c = c_base + warp * 16 + g + 8 * hrow;
    if (c < C) {
      float* p0 = partial + (static_cast<size_t>(blockIdx.x) * C + c) * R;
#pragma unroll
      for (int n = 0; n < NT; ++n)
        *reinterpret_cast<float2*>(p0 + 8 * n + i2) = make_float2(acc[n][2 * hrow], acc[n][2 * hrow + 1]);
    }
  }
}

// ----------------------------------------------------------------------------------------------
// One launch that finishes the adapter gradients of a layer.  Block roles by blockIdx.x:
//   [0, blocks_a)            dA[c*r + j]  = sum over row chunks of partial_a[chunk, c, j]
//   [blocks_a, +blocks_b)    dB[j*N + c]  = sum over row chunks of partial_b[chunk, c, j]        (transposed store)
//   the rest, one per sample ds_eff[b, j] = scaling * sum_{t : sample(t) == b} dzu[t, j] * h[t, j]
// The last role is the segmented reduction keyed by sample id: a sample's rows sit at stride b_prime in the
// sequence-first [L, B', .] layout (contiguous blocks for batch-first rows).  16 lanes cover the 16 components of a
// row, 16 rows per pass, then a fixed-order shared-memory fold (deterministic).
// ----------------------------------------------------------------------------------------------
constexpr int FIN_THREADS = 256;

__global__ void __launch_bounds__(FIN_THREADS)
adapter_grad_finalize_kernel(const float* __restrict__ partial_a, const float* __restrict__ partial_b,
                             float* __restrict__ dA, float* __restrict__ dB, int n_chunks, int K, int N, int r,
                             int blocks_a, int blocks_b, const float* __restrict__ h, const float* __restrict__ dzu,
                             float* __restrict__ ds_eff, int T, int b_prime, int num_slices, int row_div,
                             float scaling, int RPS) {
  __shared__ __align__(16) float red[FIN_THREADS / 32][RPS_MAX];
  int blk = blockIdx.x;
  if (blk < blocks_a + blocks_b) {
    const bool is_a = blk < blocks_a;
    const float* __restrict__ partial = is_a ? partial_a : partial_b;
    const int C = is_a ? K : N;
    const int i = (is_a ? blk : blk - blocks_a) * FIN_THREADS + threadIdx.x;
    if (i >= C * RPS) return;
    const int c = i / RPS, j = i - c * RPS;
    // four independent accumulators keep four loads in flight (fixed association order: deterministic)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    const size_t stride = static_cast<size_t>(C) * RPS;
    const float* pp = partial + static_cast<size_t>(c) * RPS + j;
    int k = 0;
    for (; k + 4 <= n_chunks; k += 4) {
      a0 += pp[(k + 0) * stride];
      a1 += pp[(k + 1) * stride];
      a2 += pp[(k + 2) * stride];
      a3 += pp[(k + 3) * stride];
    }
    for (; k < n_chunks; ++k) a0 += pp[k * stride];
    const float acc = (a0 + a1) + (a2 + a3);
    if (j >= r) return;
    if (is_a) dA[static_cast<size_t>(c) * r + j] = acc;
    else dB[static_cast<size_t>(j) * N + c] = acc;
    return;
  }
  // ---- per-sample segmented reduction: RPS/4 threads per row (float4 each), FIN_THREADS*4/RPS rows per pass ----
  const int b = blk - blocks_a - blocks_b;
  const int q4 = RPS >> 2;                       // 4 (rank 16) or 8 (rank 32) float4 lanes per row
  const int j4 = threadIdx.x % q4, lane_row = threadIdx.x / q4;
  const int rows_per_pass = FIN_THREADS / q4;
  const int L = T / b_prime;
  const int rows = L * num_slices;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 4
  for (int i = lane_row; i < rows; i += rows_per_pass) {
    const int l = i / num_slices, sl = i - l * num_slices;
    // sequence-first rows: t = l*B' + column; batch-first rows (row_div = L): t = column*L + l
    const size_t col = static_cast<size_t>(b) * num_slices + sl;
    const size_t t = (row_div == 1) ? static_cast<size_t>(l) * b_prime + col : col * row_div + l;
    const float4 a = __ldg(reinterpret_cast<const float4*>(dzu + t * RPS) + j4);
    const float4 c = __ldg(reinterpret_cast<const float4*>(h + t * RPS) + j4);
    acc.x = fmaf(a.x, c.x, acc.x);
    acc.y = fmaf(a.y, c.y, acc.y);
    acc.z = fmaf(a.z, c.z, acc.z);
    acc.w = fmaf(a.w, c.w, acc.w);
  }
  // fold the rows of a warp with shuffles (lanes with equal j4 sit q4 apart), then the warps through shared memory:
  // fixed order both times, so the result is deterministic
  for (int off = q4; off < 32; off <<= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, off);
    acc.y += __shfl_xor_sync(0xffffffffu, acc.y, off);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, off);
    acc.w += __shfl_xor_sync(0xffffffffu, acc.w, off);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane < q4) *reinterpret_cast<float4*>(&red[warp][4 * lane]) = acc;
  __syncthreads();
  if (threadIdx.x < r) {
    float tot = 0.f;
#pragma unroll
    for (int k = 0; k < FIN_THREADS / 32; ++k) tot += red[k][threadIdx.x];
    ds_eff[b * r + threadIdx.x] = tot * scaling;
  }
}

constexpr int CS_MAX_CHUNKS = 64;  // row chunks (partials: chunks x C x 16 fp32)

// three CTAs fit per SM (61 KB of smem each): fill those slots in ONE wave — a ceil() here once produced 300 CTAs for
// 296 slots and the four stragglers doubled the kernel time (ncu: DRAM 45 % busy)
static int pick_chunks(int T, int K, int N, int rp) {
  const int col_groups = (K + CS_COLS - 1) / CS_COLS + (N + CS_COLS - 1) / CS_COLS;
  const int ctas_per_sm = rp <= 16 ? 3 : 2;          // TMA build: 75 KB (rank 16) / 83 KB (rank 32) of smem per CTA
  int chunks = (ctas_per_sm * num_sms()) / col_groups;
  if (chunks > CS_MAX_CHUNKS) chunks = CS_MAX_CHUNKS;
  const int max_chunks = (T + CS_ROWS - 1) / CS_ROWS;
  if (chunks > max_chunks) chunks = max_chunks;
  if (chunks < 1) chunks = 1;
  return chunks;
}

size_t svlora_bwd_small_scratch_bytes(int T, int K, int N) {
  return static_cast<size_t>(pick_chunks(T, K, N, 16)) * (static_cast<size_t>(K) + N) * RPS_MAX * 4 + 256;
}

template <int R>
static int launch_adapter_grad(const __nv_bfloat16* x, const __nv_bfloat16* dh, float* partial_a, int K, int groups_a,
                               const __nv_bfloat16* dy, const __nv_bfloat16* z, float* partial_b, int N, int T,
                               int rows_per_chunk, dim3 grid, cudaStream_t stream) {
  static thread_local int attr_dev = -1;
  int dev = 0;
  FFM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev != attr_dev) {
    FFM_CHECK_CUDA(cudaFuncSetAttribute(adapter_grad_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        CsCfg<R>::SMEM_BYTES));
    attr_dev = dev;
  }
  adapter_grad_kernel<R><<<grid, CS_THREADS, CsCfg<R>::SMEM_BYTES, stream>>>(x, dh, partial_a, K, groups_a, dy, z,
                                                                            partial_b, N, T, rows_per_chunk);
  FFM_CHECK_CUDA(cudaGetLastError());
  return FFM_OK;
}

template <int R>
static int launch_adapter_grad_tma(const __nv_bfloat16* x, const __nv_bfloat16* dh, float* partial_a, int K,
                                   int groups_a, const __nv_bfloat16* dy, const __nv_bfloat16* z, float* partial_b, int N,
                                   int T, int rows_per_chunk, dim3 grid, cudaStream_t stream) {
  using Cfg = CtCfg<R>;
  CUtensorMap tm_ma, tm_va, tm_mb, tm_vb;
  int rc;
  if ((rc = make_map_bf16(&tm_ma, x, T, K, CS_ROWS, 64, CU_TENSOR_MAP_SWIZZLE_128B, true))) return rc;
  if ((rc = make_map_bf16(&tm_va, dh, T, R, CS_ROWS, R, CU_TENSOR_MAP_SWIZZLE_NONE, false))) return rc;
  if ((rc = make_map_bf16(&tm_mb, dy, T, N, CS_ROWS, 64, CU_TENSOR_MAP_SWIZZLE_128B, true))) return rc;
  if ((rc = make_map_bf16(&tm_vb, z, T, R, CS_ROWS, R, CU_TENSOR_MAP_SWIZZLE_NONE, false))) return rc;
  static thread_local int attr_dev = -1;
  int dev = 0;
  FFM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev != attr_dev) {
    FFM_CHECK_CUDA(cudaFuncSetAttribute(adapter_grad_tma_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        Cfg::SMEM_BYTES));
    attr_dev = dev;
  }
  adapter_grad_tma_kernel<R><<<grid, CS_THREADS, Cfg::SMEM_BYTES, stream>>>(tm_ma, tm_va, tm_mb, tm_vb, partial_a, K,
                                                                           groups_a, partial_b, N, T, rows_per_chunk);
  FFM_CHECK_CUDA(cudaGetLastError());
  return FFM_OK;
}

int launch_svlora_bwd_small(const __nv_bfloat16* x, const __nv_bfloat16* dy, const float* h, const float* dzu,
                            const __nv_bfloat16* z, const __nv_bfloat16* dh, float* dA, float* dB, float* ds_eff,
                            void* scratch, size_t scratch_bytes, int T, int K, int N, int r, int rp, int nS, int b_prime,
                            int num_slices, int row_div, float scaling, cudaStream_t stream) {
  FFM_CHECK_ARG(row_div == 1 || row_div * b_prime == T, "svlora bwd: batch-first rows need row_div * b_prime == T");
  FFM_CHECK_ARG(T % b_prime == 0, "svlora bwd: T (%d) must be a multiple of b_prime (%d)", T, b_prime);
  FFM_CHECK_ARG(rp == 16 || rp == RPS_MAX, "svlora bwd: padded rank must be 16 or %d", RPS_MAX);
  int chunks = pick_chunks(T, K, N, rp);
  FFM_CHECK_ARG(static_cast<size_t>(chunks) * (static_cast<size_t>(K) + N) * rp * 4 <= scratch_bytes,
                "svlora bwd: scratch too small");
  float* partial_a = static_cast<float*>(scratch);
  float* partial_b = partial_a + static_cast<size_t>(chunks) * K * rp;
  const int groups_a = (K + CS_COLS - 1) / CS_COLS, groups_b = (N + CS_COLS - 1) / CS_COLS;
  static const bool use_cp_async = getenv("FFM_AG_CPASYNC") != nullptr;     // A/B timing of the two builds
  int rc;
  if (use_cp_async) {
    const int rows_per_chunk = (T + chunks - 1) / chunks;
    // dA[K, r] = x^T · dh   and   dB[r, N] = z^T · dy
    const dim3 grid(chunks, groups_a + groups_b);
    rc = rp == 16 ? launch_adapter_grad<16>(x, dh, partial_a, K, groups_a, dy, z, partial_b, N, T, rows_per_chunk, grid, stream)
                  : launch_adapter_grad<32>(x, dh, partial_a, K, groups_a, dy, z, partial_b, N, T, rows_per_chunk, grid, stream);
  } else {
    // TMA build: chunks end on k-step (16-row) boundaries; fewer chunks than planned may remain
    const int rows_per_chunk = (((T + chunks - 1) / chunks) + 15) & ~15;
    chunks = (T + rows_per_chunk - 1) / rows_per_chunk;
    const dim3 grid(chunks, groups_a + groups_b);
    rc = rp == 16 ? launch_adapter_grad_tma<16>(x, dh, partial_a, K, groups_a, dy, z, partial_b, N, T, rows_per_chunk, grid, stream)
                  : launch_adapter_grad_tma<32>(x, dh, partial_a, K, groups_a, dy, z, partial_b, N, T, rows_per_chunk, grid, stream);
  }
  if (rc != FFM_OK) return rc;
  const int blocks_a = (K * rp + FIN_THREADS - 1) / FIN_THREADS, blocks_b = (N * rp + FIN_THREADS - 1) / FIN_THREADS;
  adapter_grad_finalize_kernel<<<blocks_a + blocks_b + nS, FIN_THREADS, 0, stream>>>(
      partial_a, partial_b, dA, dB, chunks, K, N, r, blocks_a, blocks_b, h, dzu, ds_eff, T, b_prime, num_slices,
      row_div, scaling, rp);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(2);
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
  const int n = (G + 1) * r * DS_SPLIT;
  ds_kernel<<<(n + 127) / 128, 128, 0, stream>>>(attr, ds_eff, dS, dS_global, n_samples, G, r, lambda);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
