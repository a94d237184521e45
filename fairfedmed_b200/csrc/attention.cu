// Frozen self-attention core of ResidualAttentionBlock (clip/model.py:350-352: nn.MultiheadAttention, need_weights=False)
// for the shapes of the CLIP towers: head dim 64, sequences up to 208 tokens (197 image tokens, 77 text tokens).
// Scope row f1.  softmax(Q K^T / sqrt(64)) V per (sample, head), forward and backward, reading q/k/v straight from the
// packed in_proj output [.., 3, H, 64] and writing dq/dk/dv packed the same way (no head split / concat copies).
//
// Why not tcgen05: one head is a 197 x 197 x 64 problem.  A UMMA tile is 128 rows (256 with cta_group::2), so the
// score matrix would be padded 1.7x and the softmax between the two contractions would still run on CUDA cores out
// of TMEM; the whole head fits one CTA's shared memory (4 x 26 KB), where warp-level mma.sync m16n8k16 on 16-row
// tiles wastes 5 % (197 -> 208) and keeps P / dS in registers between the contractions (measured mma.sync peak on
// this part: 540 TFLOP/s, tools/micro/mma_sync_peak.cu).
//
// One CTA (7 warps) per (sample, head), two CTAs per SM (loads of one overlap the math of the other):
//   forward : warp w owns query tiles w and w+7 (16 rows each); for each block of 64 keys: S = Q K^T (fp32
//             accumulators), online softmax in the exp2 domain, O += P V with P re-packed to bf16 A-fragments in
//             registers.  Output tile staged through the (private) Q rows for 16-byte coalesced stores; the row
//             log-sum-exp (base 2) is saved for the backward.
//   backward: recomputes P from (Q, K, lse).  Phase A (warp owns query tiles): dQ = scale * dS K with
//             dS = P ⊙ (dO V^T - delta).  Phase B (warp owns key tiles): dV = P^T dO and dK = scale * dS^T Q with the
//             transposed scores S^T = K Q^T recomputed, so no cross-warp reduction or P / dS staging is needed (7
//             contractions instead of 5, no shared-memory round trip, no atomics: deterministic).
//             delta = rowsum(dO ⊙ O) is computed while dO is staged.
// Shared-memory tiles are [208 rows][64 bf16] with the 16-byte chunk index XOR (row & 7): conflict-free ldmatrix.
#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {
namespace att {

constexpr int HD = 64;                     // head dimension
constexpr int LP = 208;                    // padded sequence (13 tiles of 16 rows)
constexpr int WARPS = 7;                   // 13 row tiles over 7 warps: tiles w and w + 7
constexpr int THREADS = WARPS * 32;        // 224
constexpr int TILE_BYTES = LP * HD * 2;    // 26624
constexpr int FWD_SMEM = 3 * TILE_BYTES;                        // Q, K, V           79872
constexpr int BWD_SMEM = 4 * TILE_BYTES + 2 * LP * 4;           // Q, K, V, dO + lse + delta   108160
constexpr float LOG2E = 1.4426950408889634f;

__device__ __forceinline__ uint32_t sw_off(int r, int c) { return static_cast<uint32_t>(r * 128 + ((c ^ (r & 7)) << 4)); }

__device__ __forceinline__ void cp16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_commit_wait_all() {
  asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void ldsm4(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t addr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, "
      "{%0, %1, %2, %3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Stage one [L x 64] head slice (row pitch `pitch` elements) into a swizzled tile; rows >= L are zero-filled.
__device__ __forceinline__ void stage_tile(uint8_t* tile, const __nv_bfloat16* src, size_t pitch, int L) {
  for (int idx = threadIdx.x; idx < LP * 8; idx += THREADS) {
    const int r = idx >> 3, c = idx & 7;
    uint8_t* dst = tile + sw_off(r, c);
    if (r < L) cp16(dst, src + static_cast<size_t>(r) * pitch + c * 8);
    else *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// A-operand fragments (16 rows x 64 columns = 4 k-steps) of row tile `rt` from a swizzled tile.
__device__ __forceinline__ void load_a_frags(uint32_t tile, int rt, int lane, uint32_t (&f)[4][4]) {
  const int row = rt * 16 + (lane & 15);
#pragma unroll
  for (int ks = 0; ks < 4; ++ks) ldsm4(tile + sw_off(row, 2 * ks + (lane >> 4)), f[ks]);
}

struct HeadPtrs {
  const __nv_bfloat16 *q, *k, *v;   // first row of this (sample, head)
  size_t pitch;                     // elements between consecutive tokens of the same sample in qkv
  size_t opitch;                    // same for the [.., H*64] tensors (out, d_out)
  size_t ooff;                      // element offset of this (sample, head) in those tensors
};

__device__ __forceinline__ HeadPtrs head_ptrs(const __nv_bfloat16* qkv, int b, int h, int B, int L, int H,
                                              int batch_first) {
  HeadPtrs p;
  const size_t C = static_cast<size_t>(H) * HD;
  const size_t tok = batch_first ? 1 : static_cast<size_t>(B);          // rows between consecutive tokens
  const size_t row0 = batch_first ? static_cast<size_t>(b) * L : static_cast<size_t>(b);
  p.pitch = tok * 3 * C;
  p.opitch = tok * C;
  const __nv_bfloat16* base = qkv + row0 * 3 * C + static_cast<size_t>(h) * HD;
  p.q = base;
  p.k = base + C;
  p.v = base + 2 * C;
  p.ooff = row0 * C + static_cast<size_t>(h) * HD;
  return p;
}

// ------------------------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------------------------
// One block of 64 keys for one 16-row query tile.  MASKED = false is the interior case (every key valid for every row):
// no index arithmetic or selects at all, 4 instructions per score (max, fma, ex2, add).  The running maximum m is kept
// in the RAW score domain (the scale is positive), so scale and subtraction fuse into one FMA in front of ex2.
template <bool CAUSAL, bool MASKED>
__device__ __forceinline__ void fwd_block(int kb0, int k_end, int L, int row0, int row1, uint32_t ks_, uint32_t vs,
                                          int lane, int t4, const uint32_t (&qf)[4][4], float (&o)[8][4], float& m0,
                                          float& m1, float& l0, float& l1, float scale_log2) {
  float s[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
  // S = Q K^T for 64 keys (8 n-tiles), two n-tiles per ldmatrix.x4
#pragma unroll
  for (int np = 0; np < 4; ++np) {
    if (!MASKED || kb0 + np * 16 < k_end) {            // warp-uniform
      const int key = kb0 + np * 16 + ((lane >> 4) << 3) + (lane & 7);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        uint32_t kf[4];
        ldsm4(ks_ + sw_off(key, 2 * kk + ((lane >> 3) & 1)), kf);
        mma16816(s[2 * np], qf[kk], kf[0], kf[1]);
        mma16816(s[2 * np + 1], qf[kk], kf[2], kf[3]);
      }
    }
  }
  if (MASKED) {
    // keys past the sequence, past k_end (skipped sub-blocks: they are >= L, or > every row of a causal tile) and above
    // the diagonal
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int key = kb0 + nt * 8 + 2 * t4 + (e & 1);
        const int row = (e & 2) ? row1 : row0;
        if (key >= L || (CAUSAL && key > row)) s[nt][e] = -INFINITY;
      }
    }
  }
  float bm0 = fmaxf(s[0][0], s[0][1]), bm1 = fmaxf(s[0][2], s[0][3]);
#pragma unroll
  for (int nt = 1; nt < 8; ++nt) {
    bm0 = fmaxf(bm0, fmaxf(s[nt][0], s[nt][1]));
    bm1 = fmaxf(bm1, fmaxf(s[nt][2], s[nt][3]));
  }
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 1));
  bm0 = fmaxf(bm0, __shfl_xor_sync(0xffffffffu, bm0, 2));
  bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 1));
  bm1 = fmaxf(bm1, __shfl_xor_sync(0xffffffffu, bm1, 2));
  const float mn0 = fmaxf(m0, bm0), mn1 = fmaxf(m1, bm1);
  // every row sees key 0 in the first block (also under the causal mask), so mn is finite from block 0 on
  const float a0 = ex2((m0 - mn0) * scale_log2), a1 = ex2((m1 - mn1) * scale_log2);
  m0 = mn0;
  m1 = mn1;
  const float ms0 = -mn0 * scale_log2, ms1 = -mn1 * scale_log2;
  float ps0 = 0.f, ps1 = 0.f;
#pragma unroll
  for (int nt = 0; nt < 8; ++nt) {
    s[nt][0] = ex2(fmaf(s[nt][0], scale_log2, ms0));
    s[nt][1] = ex2(fmaf(s[nt][1], scale_log2, ms0));
    s[nt][2] = ex2(fmaf(s[nt][2], scale_log2, ms1));
    s[nt][3] = ex2(fmaf(s[nt][3], scale_log2, ms1));
    ps0 += s[nt][0] + s[nt][1];
    ps1 += s[nt][2] + s[nt][3];
  }
  l0 = fmaf(l0, a0, ps0);
  l1 = fmaf(l1, a1, ps1);
#pragma unroll
  for (int dn = 0; dn < 8; ++dn) {
    o[dn][0] *= a0; o[dn][1] *= a0;
    o[dn][2] *= a1; o[dn][3] *= a1;
  }
  // O += P V : P re-packed as A fragments (k = 16 keys per step), V^T fragments through ldmatrix.trans
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    if (!MASKED || kb0 + j * 16 < k_end) {
      uint32_t pa[4];
      pa[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
      pa[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
      pa[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
      pa[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
      const int key = kb0 + j * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {
        uint32_t vf[4];
        ldsm4t(vs + sw_off(key, 2 * dp + (lane >> 4)), vf);
        mma16816(o[2 * dp], pa, vf[0], vf[1]);
        mma16816(o[2 * dp + 1], pa, vf[2], vf[3]);
      }
    }
  }
}

template <bool CAUSAL>
__global__ void __launch_bounds__(THREADS, 2)
attention_fwd_kernel(const __nv_bfloat16* __restrict__ qkv, __nv_bfloat16* __restrict__ out, float* __restrict__ lse,
                     int B, int L, int H, int batch_first, float scale_log2) {
  extern __shared__ __align__(1024) uint8_t att_smem[];
  uint8_t* Qs = att_smem;
  uint8_t* Ks = Qs + TILE_BYTES;
  uint8_t* Vs = Ks + TILE_BYTES;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const HeadPtrs hp = head_ptrs(qkv, b, h, B, L, H, batch_first);
  stage_tile(Qs, hp.q, hp.pitch, L);
  stage_tile(Ks, hp.k, hp.pitch, L);
  stage_tile(Vs, hp.v, hp.pitch, L);
  cp_commit_wait_all();
  __syncthreads();

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int n_tiles = (L + 15) >> 4;
  const int Lk = n_tiles * 16;                          // keys actually visited (multiple of 16)
  const uint32_t qs = smem_u32(Qs), ks_ = smem_u32(Ks), vs = smem_u32(Vs);

  for (int rt = warp; rt < n_tiles; rt += WARPS) {
    uint32_t qf[4][4];
    load_a_frags(qs, rt, lane, qf);
    float o[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;
    const int row0 = rt * 16 + g, row1 = row0 + 8;
    const int k_end = CAUSAL ? min(Lk, rt * 16 + 16) : Lk;      // causal: keys beyond the tile's last row are masked

    for (int kb0 = 0; kb0 < k_end; kb0 += 64) {
      // interior block: all 64 keys exist and (causal) lie at or below the tile's first row
      const bool interior = (kb0 + 64 <= L) && (!CAUSAL || kb0 + 63 <= rt * 16);
      if (interior) fwd_block<CAUSAL, false>(kb0, k_end, L, row0, row1, ks_, vs, lane, t4, qf, o, m0, m1, l0, l1, scale_log2);
      else fwd_block<CAUSAL, true>(kb0, k_end, L, row0, row1, ks_, vs, lane, t4, qf, o, m0, m1, l0, l1, scale_log2);
    }
    // finish the rows: l over the quad, normalise, stage the tile through this tile's (private) Q rows
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __syncwarp();
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      *reinterpret_cast<uint32_t*>(Qs + sw_off(row0, dn) + t4 * 4) = pack_bf16x2(o[dn][0] * i0, o[dn][1] * i0);
      *reinterpret_cast<uint32_t*>(Qs + sw_off(row1, dn) + t4 * 4) = pack_bf16x2(o[dn][2] * i1, o[dn][3] * i1);
    }
    if (t4 == 0) {
      float* lp = lse + static_cast<size_t>(blockIdx.x) * L;
      if (row0 < L) lp[row0] = fmaf(m0, scale_log2, log2f(l0));
      if (row1 < L) lp[row1] = fmaf(m1, scale_log2, log2f(l1));
    }
    __syncwarp();
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = lane + 32 * i;
      const int r = rt * 16 + (idx >> 3), c = idx & 7;
      if (r < L)
        *reinterpret_cast<uint4*>(out + hp.ooff + static_cast<size_t>(r) * hp.opitch + c * 8) =
            *reinterpret_cast<const uint4*>(Qs + sw_off(r, c));
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// backward
// ------------------------------------------------------------------------------------------------------------
template <bool causal>
__global__ void __launch_bounds__(THREADS, 2)
attention_bwd_kernel(const __nv_bfloat16* __restrict__ qkv, const __nv_bfloat16* __restrict__ out,
                     const __nv_bfloat16* __restrict__ d_out, const float* __restrict__ lse,
                     __nv_bfloat16* __restrict__ d_qkv, int B, int L, int H, int batch_first, float scale,
                     float scale_log2) {
  extern __shared__ __align__(1024) uint8_t att_smem[];
  uint8_t* Qs = att_smem;
  uint8_t* Ks = Qs + TILE_BYTES;
  uint8_t* Vs = Ks + TILE_BYTES;
  uint8_t* Gs = Vs + TILE_BYTES;                                  // dO
  float* lse_s = reinterpret_cast<float*>(Gs + TILE_BYTES);
  float* del_s = lse_s + LP;
  const int b = blockIdx.x / H, h = blockIdx.x - b * H;
  const HeadPtrs hp = head_ptrs(qkv, b, h, B, L, H, batch_first);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t4 = lane & 3;

  stage_tile(Qs, hp.q, hp.pitch, L);
  stage_tile(Ks, hp.k, hp.pitch, L);
  stage_tile(Vs, hp.v, hp.pitch, L);
  // dO through registers: delta[r] = sum_d dO[r, d] * O[r, d] on the way (8 lanes share a row)
  for (int base = 0; base < LP * 8; base += THREADS) {
    const int idx = base + threadIdx.x;
    const int r = idx >> 3, c = idx & 7;
    float part = 0.f;
    if (r < LP) {
      uint4 gv = make_uint4(0u, 0u, 0u, 0u);
      if (r < L) {
        const size_t off = hp.ooff + static_cast<size_t>(r) * hp.opitch + c * 8;
        gv = __ldg(reinterpret_cast<const uint4*>(d_out + off));
        const uint4 ov = __ldg(reinterpret_cast<const uint4*>(out + off));
        const __nv_bfloat162* g2 = reinterpret_cast<const __nv_bfloat162*>(&gv);
        const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 a = __bfloat1622float2(g2[e]), c2 = __bfloat1622float2(o2[e]);
          part = fmaf(a.x, c2.x, part);
          part = fmaf(a.y, c2.y, part);
        }
      }
      *reinterpret_cast<uint4*>(Gs + sw_off(r, c)) = gv;
    }
    part += __shfl_xor_sync(0xffffffffu, part, 1);
    part += __shfl_xor_sync(0xffffffffu, part, 2);
    part += __shfl_xor_sync(0xffffffffu, part, 4);
    if (c == 0 && r < LP) del_s[r] = part;
  }
  for (int r = threadIdx.x; r < LP; r += THREADS)
    lse_s[r] = r < L ? lse[static_cast<size_t>(blockIdx.x) * L + r] : INFINITY;     // padded rows: P = exp2(-inf) = 0
  cp_commit_wait_all();
  __syncthreads();

  const int n_tiles = (L + 15) >> 4;
  const int Lk = n_tiles * 16;
  const uint32_t qs = smem_u32(Qs), ks_ = smem_u32(Ks), vs = smem_u32(Vs), gs = smem_u32(Gs);
  const size_t C = static_cast<size_t>(H) * HD;
  __nv_bfloat16* dq_base = d_qkv + (hp.q - qkv);
  __nv_bfloat16* dk_base = dq_base + C;
  __nv_bfloat16* dv_base = dq_base + 2 * C;

  // ---------------- phase A: dQ (warp owns query tiles) ----------------
  for (int rt = warp; rt < n_tiles; rt += WARPS) {
    uint32_t qf[4][4], gf[4][4];
    load_a_frags(qs, rt, lane, qf);
    load_a_frags(gs, rt, lane, gf);
    const int row0 = rt * 16 + g, row1 = row0 + 8;
    const float ls0 = -lse_s[row0], ls1 = -lse_s[row1], de0 = del_s[row0], de1 = del_s[row1];
    float dq[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
    const int k_end = causal ? min(Lk, rt * 16 + 16) : Lk;
    for (int kb0 = 0; kb0 < k_end; kb0 += 32) {
      float s[4][4], dp[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
        dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
      }
#pragma unroll
      for (int np = 0; np < 2; ++np) {
        if (kb0 + np * 16 < k_end) {
          const int key = kb0 + np * 16 + ((lane >> 4) << 3) + (lane & 7);
#pragma unroll
          for (int kk = 0; kk < 4; ++kk) {
            uint32_t f[4];
            ldsm4(ks_ + sw_off(key, 2 * kk + ((lane >> 3) & 1)), f);
            mma16816(s[2 * np], qf[kk], f[0], f[1]);
            mma16816(s[2 * np + 1], qf[kk], f[2], f[3]);
            ldsm4(vs + sw_off(key, 2 * kk + ((lane >> 3) & 1)), f);
            mma16816(dp[2 * np], gf[kk], f[0], f[1]);
            mma16816(dp[2 * np + 1], gf[kk], f[2], f[3]);
          }
        }
      }
      // dS = P ⊙ (dP - delta); padded keys have K = V = 0, so whatever P they get meets a zero K row below
#pragma unroll
      for (int nt = 0; nt < 4; ++nt) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int key = kb0 + nt * 8 + 2 * t4 + (e & 1);
          const int row = (e & 2) ? row1 : row0;
          float p = ex2(fmaf(s[nt][e], scale_log2, (e & 2) ? ls1 : ls0));
          if (causal && key > row) p = 0.f;
          s[nt][e] = p * (dp[nt][e] - ((e & 2) ? de1 : de0));
        }
      }
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        if (kb0 + j * 16 < k_end) {
          uint32_t da[4];
          da[0] = pack_bf16x2(s[2 * j][0], s[2 * j][1]);
          da[1] = pack_bf16x2(s[2 * j][2], s[2 * j][3]);
          da[2] = pack_bf16x2(s[2 * j + 1][0], s[2 * j + 1][1]);
          da[3] = pack_bf16x2(s[2 * j + 1][2], s[2 * j + 1][3]);
          const int key = kb0 + j * 16 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
          for (int dpi = 0; dpi < 4; ++dpi) {
            uint32_t f[4];
            ldsm4t(ks_ + sw_off(key, 2 * dpi + (lane >> 4)), f);
            mma16816(dq[2 * dpi], da, f[0], f[1]);
            mma16816(dq[2 * dpi + 1], da, f[2], f[3]);
          }
        }
      }
    }
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      if (row0 < L)
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<size_t>(row0) * hp.pitch + dn * 8 + 2 * t4) =
            pack_bf16x2(dq[dn][0] * scale, dq[dn][1] * scale);
      if (row1 < L)
        *reinterpret_cast<uint32_t*>(dq_base + static_cast<size_t>(row1) * hp.pitch + dn * 8 + 2 * t4) =
            pack_bf16x2(dq[dn][2] * scale, dq[dn][3] * scale);
    }
  }

  // ---------------- phase B: dK, dV (warp owns key tiles; transposed scores) ----------------
  for (int kt = warp; kt < n_tiles; kt += WARPS) {
    uint32_t kf[4][4];
    load_a_frags(ks_, kt, lane, kf);
    const int key0 = kt * 16 + g, key1 = key0 + 8;
    float dk[8][4], dv[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    }
    const int q_begin = causal ? kt * 16 : 0;             // causal: queries before this key tile never see it
    for (int qb0 = q_begin; qb0 < Lk; qb0 += 16) {
      float st[2][4], dpt[2][4];
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
      {
        const int qrow = qb0 + ((lane >> 4) << 3) + (lane & 7);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          uint32_t f[4], vfrag[4];
          ldsm4(qs + sw_off(qrow, 2 * kk + ((lane >> 3) & 1)), f);             // B = Q rows of this query block
          mma16816(st[0], kf[kk], f[0], f[1]);
          mma16816(st[1], kf[kk], f[2], f[3]);
          ldsm4(vs + sw_off(kt * 16 + (lane & 15), 2 * kk + (lane >> 4)), vfrag); // A = V rows of this key tile
          ldsm4(gs + sw_off(qrow, 2 * kk + ((lane >> 3) & 1)), f);             // B = dO rows of this query block
          mma16816(dpt[0], vfrag, f[0], f[1]);
          mma16816(dpt[1], vfrag, f[2], f[3]);
        }
      }
      // P^T and dS^T: rows = keys (key0 / key1), columns = queries
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int q0 = qb0 + nt * 8 + 2 * t4;
        const float2 lq = *reinterpret_cast<const float2*>(lse_s + q0);
        const float2 dq2 = *reinterpret_cast<const float2*>(del_s + q0);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qi = q0 + (e & 1);
          const int key = (e & 2) ? key1 : key0;
          float p = ex2(st[nt][e] * scale_log2 - ((e & 1) ? lq.y : lq.x));
          if (causal && key > qi) p = 0.f;
          st[nt][e] = p;
          dpt[nt][e] = p * (dpt[nt][e] - ((e & 1) ? dq2.y : dq2.x));
        }
      }
      uint32_t pa[4], da[4];
      pa[0] = pack_bf16x2(st[0][0], st[0][1]);   pa[1] = pack_bf16x2(st[0][2], st[0][3]);
      pa[2] = pack_bf16x2(st[1][0], st[1][1]);   pa[3] = pack_bf16x2(st[1][2], st[1][3]);
      da[0] = pack_bf16x2(dpt[0][0], dpt[0][1]); da[1] = pack_bf16x2(dpt[0][2], dpt[0][3]);
      da[2] = pack_bf16x2(dpt[1][0], dpt[1][1]); da[3] = pack_bf16x2(dpt[1][2], dpt[1][3]);
      const int qrow_t = qb0 + (lane & 7) + (((lane >> 3) & 1) << 3);
#pragma unroll
      for (int dpi = 0; dpi < 4; ++dpi) {
        uint32_t f[4];
        ldsm4t(gs + sw_off(qrow_t, 2 * dpi + (lane >> 4)), f);                  // B[k = query][n = d] = dO
        mma16816(dv[2 * dpi], pa, f[0], f[1]);
        mma16816(dv[2 * dpi + 1], pa, f[2], f[3]);
        ldsm4t(qs + sw_off(qrow_t, 2 * dpi + (lane >> 4)), f);                  // B[k = query][n = d] = Q
        mma16816(dk[2 * dpi], da, f[0], f[1]);
        mma16816(dk[2 * dpi + 1], da, f[2], f[3]);
      }
    }
#pragma unroll
    for (int dn = 0; dn < 8; ++dn) {
      const size_t c0 = dn * 8 + 2 * t4;
      if (key0 < L) {
        *reinterpret_cast<uint32_t*>(dk_base + static_cast<size_t>(key0) * hp.pitch + c0) =
            pack_bf16x2(dk[dn][0] * scale, dk[dn][1] * scale);
        *reinterpret_cast<uint32_t*>(dv_base + static_cast<size_t>(key0) * hp.pitch + c0) =
            pack_bf16x2(dv[dn][0], dv[dn][1]);
      }
      if (key1 < L) {
        *reinterpret_cast<uint32_t*>(dk_base + static_cast<size_t>(key1) * hp.pitch + c0) =
            pack_bf16x2(dk[dn][2] * scale, dk[dn][3] * scale);
        *reinterpret_cast<uint32_t*>(dv_base + static_cast<size_t>(key1) * hp.pitch + c0) =
            pack_bf16x2(dv[dn][2], dv[dn][3]);
      }
    }
  }
}

static int set_smem_once() {
  static thread_local int attr_dev = -1;
  int dev = 0;
  FFM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev != attr_dev) {
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    attr_dev = dev;
  }
  return FFM_OK;
}

}  // namespace att
}  // namespace ffm

using namespace ffm;

extern "C" {

int ffm_attention_max_len(void) { return att::LP; }

int ffm_attention_fwd(const void* qkv, void* out, float* lse, int B, int L, int H, int head_dim, int causal,
                      int batch_first, cudaStream_t stream) {
  FFM_CHECK_ARG(qkv && out && lse, "ffm_attention_fwd: null pointer argument");
  FFM_CHECK_ARG(head_dim == att::HD, "ffm_attention_fwd: head dimension must be %d", att::HD);
  FFM_CHECK_ARG(B >= 1 && H >= 1 && L >= 1 && L <= att::LP, "ffm_attention_fwd: 1 <= L <= %d", att::LP);
  int rc = att::set_smem_once();
  if (rc != FFM_OK) return rc;
  const float scale_log2 = att::LOG2E / sqrtf(static_cast<float>(att::HD));
  if (causal)
    att::attention_fwd_kernel<true><<<B * H, att::THREADS, att::FWD_SMEM, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, B, L, H, batch_first ? 1 : 0,
        scale_log2);
  else
    att::attention_fwd_kernel<false><<<B * H, att::THREADS, att::FWD_SMEM, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<__nv_bfloat16*>(out), lse, B, L, H, batch_first ? 1 : 0,
        scale_log2);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* d_qkv, int B, int L,
                      int H, int head_dim, int causal, int batch_first, cudaStream_t stream) {
  FFM_CHECK_ARG(qkv && out && d_out && lse && d_qkv, "ffm_attention_bwd: null pointer argument");
  FFM_CHECK_ARG(head_dim == att::HD, "ffm_attention_bwd: head dimension must be %d", att::HD);
  FFM_CHECK_ARG(B >= 1 && H >= 1 && L >= 1 && L <= att::LP, "ffm_attention_bwd: 1 <= L <= %d", att::LP);
  int rc = att::set_smem_once();
  if (rc != FFM_OK) return rc;
  const float scale = 1.0f / sqrtf(static_cast<float>(att::HD));
  if (causal)
    att::attention_bwd_kernel<true><<<B * H, att::THREADS, att::BWD_SMEM, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
        static_cast<const __nv_bfloat16*>(d_out), lse, static_cast<__nv_bfloat16*>(d_qkv), B, L, H, batch_first ? 1 : 0,
        scale, scale * att::LOG2E);
  else
    att::attention_bwd_kernel<false><<<B * H, att::THREADS, att::BWD_SMEM, stream>>>(
        static_cast<const __nv_bfloat16*>(qkv), static_cast<const __nv_bfloat16*>(out),
        static_cast<const __nv_bfloat16*>(d_out), lse, static_cast<__nv_bfloat16*>(d_qkv), B, L, H, batch_first ? 1 : 0,
        scale, scale * att::LOG2E);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

}  // extern "C"
