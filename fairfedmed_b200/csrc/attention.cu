// Frozen self-attention core of ResidualAttentionBlock (clip/model.py:350-352: nn.MultiheadAttention, need_weights=False)
// on Blackwell tensor cores for the shapes of the CLIP towers: head dim 64, sequences up to 208 tokens (197 image tokens,
// 77 text tokens).  Scope row f1.  softmax(Q K^T / sqrt(64)) V per (sample, head), forward and backward, reading q/k/v
// straight from the packed in_proj output [.., 3, H, 64] through 3-D TMA maps (rows past the sequence end are zero-filled
// by the hardware) and writing dq/dk/dv packed the same way (no head split / concat copies).
//
// Both kernels are persistent (one CTA per SM, heads round-robin), warp-specialised, and keep every intermediate
// (scores, probabilities, dS) out of shared memory: the contractions accumulate in TMEM, the softmax warps read the
// scores with tcgen05.ld (one thread per row: no shuffles), write the bf16 probabilities back INTO TMEM with tcgen05.st,
// and the second contraction takes them from there as its A operand (tcgen05.mma with A in TMEM).  V, dO, Q and K are
// consumed as "MN-major" B operands in place (the [token][channel] tile as TMA delivers it), so no operand is transposed.
//
//   forward : 2 query tiles of 128 rows per head (rows >= L are zero / discarded), all keys at once (N = L rounded up
//             to 16, one UMMA_N): S_t = Q_t K^T -> TMEM[256 t .. +208); softmax warpgroup t: row max, P = exp2(..) as
//             bf16 pairs over the consumed columns, row sum; O_t = P_t V (A from TMEM, K = keys) -> TMEM[256 t + 128 ..
//             +64); epilogue O / sum -> bf16 rows of `out`, base-2 log-sum-exp to `lse`.  Q / K / V of the next head
//             are prefetched into the second shared-memory stage while the current head computes.
//   backward: per key tile j (128 keys = TMEM lanes) and query half a (<= 128 query columns):
//             S^T = K_j Q_a^T and dP^T = V_j dO_a^T (TMEM regions R0 / R1); the softmax warps form P^T = exp2(S^T c -
//             lse[q]) and dS^T = P^T (dP^T - delta[q]) per element, store both as bf16 pairs into TMEM and dS^T also
//             into a swizzled shared-memory tile; dV_j += P^T dO_a and dK_j += dS^T Q_a take A from TMEM, dQ_a += dS K_j
//             takes the shared tile as an MN-major A operand.  The five accumulators (dV_j, dK_j, dQ_0, dQ_1 in 4 x 64
//             TMEM columns + the two 128-column regions = 512) never leave the chip; delta = rowsum(dO * O) is computed
//             while the operands land.  No atomics: deterministic.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <mutex>

#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

// implemented in svlora_gemm.cu
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
PFN_encodeTiled tensor_map_encode_fn();

namespace att {

constexpr int HD = 64;                      // head dimension
constexpr int LP = 208;                     // longest sequence (13 x 16)
constexpr float LOG2E = 1.4426950408889634f;
constexpr int ROW_BYTES = HD * 2;           // 128 B = one SW128 swizzle row
constexpr int Q2_BYTES = 256 * ROW_BYTES;   // two 128-row tiles           32768
constexpr int KV_BYTES = LP * ROW_BYTES;    // all keys                    26624

__device__ __forceinline__ uint32_t pack2(float a, float b) { return pack_bf16x2(a, b); }

// ---------------------------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------------------------
// A unit = (head, query tile of 128 rows).  Two score buffers in TMEM ping-pong ([0, 208) and [208, 416)), the output
// accumulator has its own 64 columns, so the tensor pipe works on S of unit n + 1 and P V of unit n - 1 while the softmax
// warps are on unit n; two softmax groups alternate units and four separate warps drain O, so a softmax group never waits
// for its own P V or epilogue.
constexpr int FWD_THREADS = 448;            // warp 0: TMA, 1: MMA (+ TMEM alloc), 2-5 / 6-9: softmax groups, 10-13: epilogue
constexpr int FWD_STAGE = Q2_BYTES + 2 * KV_BYTES;                 // 86016: Q (2 tiles), K, V of one head
constexpr int FWD_OFF_OUT = 2 * FWD_STAGE;                          // 172032: output staging tile [128][64] bf16, SW128
constexpr int FWD_OFF_STAT = FWD_OFF_OUT + 128 * ROW_BYTES;         // 188416: [2 buffers][128 rows] {1 / sum, lse2}
constexpr int FWD_OFF_BAR = FWD_OFF_STAT + 2 * 128 * 8;
constexpr int FWD_SMEM = FWD_OFF_BAR + 256 + 1024;
constexpr uint32_t FWD_O_COL = 416;

struct FwdParams {
  float* lse;
  int B, L, H, C, batch_first;
  int lk_pad;        // L rounded up to 16: UMMA N of the scores, K extent of P V
  int n_tiles;       // query tiles of 128 rows (1 or 2)
  int total;         // B * H
  float scale_log2;  // log2(e) / sqrt(head_dim)
};

template <bool CAUSAL>
__global__ void __launch_bounds__(FWD_THREADS, 1)
attention_fwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        const __grid_constant__ CUtensorMap tm_out, const FwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float2* stat = reinterpret_cast<float2*>(smem + FWD_OFF_STAT);       // [2][128]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FWD_OFF_BAR);
  uint64_t* full = bars;            // [2] operands of a head landed
  uint64_t* empty = bars + 2;       // [2] all MMAs reading the stage completed
  uint64_t* s_full = bars + 4;      // [2] scores of a unit in TMEM buffer b
  uint64_t* p_full = bars + 6;      // [2] probabilities (and row statistics) of buffer b written
  uint64_t* o_full = bars + 8;      //     P V of a unit complete
  uint64_t* e_done = bars + 9;      // [2] epilogue of a unit with parity b finished (O drained, statistics read)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 11);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const int nt = p.n_tiles;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_out);
    mbar_init(o_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 4);
      mbar_init(&e_done[i], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    const uint32_t q_bytes = static_cast<uint32_t>(128 * nt) * ROW_BYTES, kv_bytes = static_cast<uint32_t>(p.lk_pad) * ROW_BYTES;
    int it = 0;
    for (int u = blockIdx.x; u < p.total; u += gridDim.x, ++it) {
      const int s = it & 1;
      mbar_wait_uniform(&empty[s], ((it >> 1) & 1) ^ 1);
      if (elect_one()) {
        const int b = u / p.H, h = u - b * p.H;
        uint8_t* st = smem + s * FWD_STAGE;
        mbar_arrive_expect_tx(&full[s], q_bytes + 2 * kv_bytes);
        tma_load_3d(st, &tm_q, &full[s], h * HD, 0, b);
        tma_load_3d(st + Q2_BYTES, &tm_kv, &full[s], p.C + h * HD, 0, b);
        tma_load_3d(st + Q2_BYTES + KV_BYTES, &tm_kv, &full[s], 2 * p.C + h * HD, 0, b);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t idesc_s = umma_idesc_bf16(128, p.lk_pad);
    const uint32_t idesc_pv = umma_idesc_bf16(128, HD) | UMMA_IDESC_B_MN;
    const int pv_steps = p.lk_pad / 16;
    const uint32_t sb = smem_u32(smem);
    // O = P V of unit m (its head sits in stage `stage`); `last` releases the stage to the producer
    auto issue_pv = [&](uint32_t m, int stage, bool last) {
      mbar_wait_uniform(&p_full[m & 1u], (m >> 1) & 1u);
      if (m > 0) mbar_wait_uniform(&e_done[(m - 1) & 1u], ((m - 1) >> 1) & 1u);   // O of unit m - 1 drained
      tc_fence_after();
      if (elect_one()) {
        const uint32_t vb = sb + stage * FWD_STAGE + Q2_BYTES + KV_BYTES;
        const uint32_t pa = tmem_base + 208u * (m & 1u);
        for (int k = 0; k < pv_steps; ++k)
          umma_bf16_ts(tmem_base + FWD_O_COL, pa + 8 * k, umma_desc_sw128_mn(vb + k * 2048, 16), idesc_pv,
                       k != 0 ? 1u : 0u);
        umma_commit(o_full);
        if (last) umma_commit(&empty[stage]);
      }
      __syncwarp();
    };
    int it = 0;
    uint32_t n = 0;                                           // running unit counter
    int prev_stage = 0;
    bool prev_last = false;
    for (int u = blockIdx.x; u < p.total; u += gridDim.x, ++it) {
      const int s = it & 1;
      const uint32_t st = sb + s * FWD_STAGE;
      mbar_wait_uniform(&full[s], (it >> 1) & 1);
      tc_fence_after();
      for (int t = 0; t < nt; ++t, ++n) {
        // S of unit n into buffer n & 1 (its previous tenant's probabilities were consumed by P V of unit n - 2, issued
        // before this point and executed in order)
        if (elect_one()) {
          const uint64_t ad = umma_desc_sw128(st + t * (128 * ROW_BYTES));
          const uint64_t bd = umma_desc_sw128(st + Q2_BYTES);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k)
            umma_bf16(tmem_base + 208u * (n & 1u), ad + 2u * k, bd + 2u * k, idesc_s, k != 0 ? 1u : 0u);
          umma_commit(&s_full[n & 1u]);
        }
        __syncwarp();
        if (n > 0) issue_pv(n - 1, prev_stage, prev_last);   // runs under the softmax of unit n
        prev_stage = s;
        prev_last = t == nt - 1;
      }
    }
    if (n > 0) issue_pv(n - 1, prev_stage, prev_last);
  } else if (warp < 10) {
    // ============================== softmax (one thread per query row) ==============================
    const uint32_t grp = (warp - 2u) >> 2;                   // handles units with (n & 1) == grp
    const uint32_t quarter = warp & 3u;                      // TMEM lane quarter this warp may touch
    const int row = static_cast<int>(quarter * 32u + lane);
    const uint32_t taddr = tmem_base + ((quarter * 32u) << 16) + 208u * grp;
    uint32_t n = 0;
    for (int u = blockIdx.x; u < p.total; u += gridDim.x) {
      for (int t = 0; t < nt; ++t, ++n) {
        if ((n & 1u) != grp) continue;
        const int qi = t * 128 + row;
        const int ncol = CAUSAL ? (qi + 1 < p.L ? qi + 1 : p.L) : p.L;      // valid key columns of this row
        const bool warp_live = t * 128 + static_cast<int>(quarter) * 32 < p.L;   // any valid query row in this warp?
        mbar_wait(&s_full[grp], (n >> 1) & 1u, 100 + grp);
        tc_fence_after();
        float m = -3.0e38f, sum = 0.f, mc = 0.f;
        if (warp_live) {
          // pass 1: row maximum of the raw scores (the next chunk's TMEM read is in flight under the current one)
          uint32_t va[16], vb[16];
          tmem_ld16(taddr, va);
          for (int c0 = 0; c0 < p.lk_pad; c0 += 32) {
            tmem_ld_wait();
            const bool has_b = c0 + 16 < p.lk_pad;
            if (has_b) tmem_ld16(taddr + c0 + 16, vb);
            if (c0 + 16 <= ncol) {
#pragma unroll
              for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(va[j]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < ncol) m = fmaxf(m, __uint_as_float(va[j]));
            }
            if (has_b) {
              tmem_ld_wait();
              if (c0 + 32 < p.lk_pad) tmem_ld16(taddr + c0 + 32, va);
              if (c0 + 32 <= ncol) {
#pragma unroll
                for (int j = 0; j < 16; ++j) m = fmaxf(m, __uint_as_float(vb[j]));
              } else {
#pragma unroll
                for (int j = 0; j < 16; ++j)
                  if (c0 + 16 + j < ncol) m = fmaxf(m, __uint_as_float(vb[j]));
              }
            }
          }
          // pass 2: P = exp2(s c - m c) as bf16 pairs over the columns already consumed; fp32 row sum
          mc = m * p.scale_log2;
          uint32_t v[16], w[16];
          tmem_ld16(taddr, v);
          for (int c0 = 0; c0 < p.lk_pad; c0 += 16) {
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 16; ++j) w[j] = v[j];
            if (c0 + 16 < p.lk_pad) tmem_ld16(taddr + c0 + 16, v);
            float e[16];
            if (c0 + 16 <= ncol) {
#pragma unroll
              for (int j = 0; j < 16; ++j) e[j] = fast_exp2(fmaf(__uint_as_float(w[j]), p.scale_log2, -mc));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                e[j] = (c0 + j < ncol) ? fast_exp2(fmaf(__uint_as_float(w[j]), p.scale_log2, -mc)) : 0.f;
            }
            uint32_t pk[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              sum += e[2 * j] + e[2 * j + 1];
              pk[j] = pack2(e[2 * j], e[2 * j + 1]);
            }
            tmem_st8(taddr + (c0 >> 1), pk);
          }
        }
        // row statistics for the epilogue warps (their previous reader, the epilogue of unit n - 2, must be done)
        if (n >= 2) mbar_wait(&e_done[grp], ((n - 2) >> 1) & 1u, 150 + grp);
        stat[grp * 128 + row] = make_float2(warp_live ? 1.0f / sum : 0.f, mc + log2f(sum));
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[grp]);
      }
    }
  } else {
    // ============================== epilogue: O / sum -> bf16 tile -> TMA store; log-sum-exp ==============================
    const uint32_t quarter = warp & 3u;
    const int row = static_cast<int>(quarter * 32u + lane);
    const uint32_t oaddr = tmem_base + ((quarter * 32u) << 16) + FWD_O_COL;
    uint8_t* tile = smem + FWD_OFF_OUT;
    uint8_t* line = tile + row * ROW_BYTES;
    const bool storer = warp == 10 && lane == 0;
    uint32_t n = 0;
    for (int u = blockIdx.x; u < p.total; u += gridDim.x) {
      const int b = u / p.H, h = u - b * p.H;
      for (int t = 0; t < nt; ++t, ++n) {
        mbar_wait(o_full, n & 1u, 200);
        // the statistics were released on p_full by the softmax warps: acquire that barrier directly (already complete: the
        // P V contraction that o_full tracks was issued after it) instead of relying on the hand-off through the MMA warp
        mbar_wait(&p_full[n & 1u], (n >> 1) & 1u, 201);
        tc_fence_after();
        const float2 st2 = stat[(n & 1u) * 128 + row];
        const int qi = t * 128 + row;
        if (storer) tma_store_wait_read<0>();                // the previous store has finished reading the tile
        named_bar_sync(2, 128);
#pragma unroll
        for (int c0 = 0; c0 < HD; c0 += 16) {
          uint32_t v[16];
          tmem_ld16(oaddr + c0, v);
          tmem_ld_wait();
          const float inv = st2.x;
          uint4 o0, o1;
          o0.x = pack2(__uint_as_float(v[0]) * inv, __uint_as_float(v[1]) * inv);
          o0.y = pack2(__uint_as_float(v[2]) * inv, __uint_as_float(v[3]) * inv);
          o0.z = pack2(__uint_as_float(v[4]) * inv, __uint_as_float(v[5]) * inv);
          o0.w = pack2(__uint_as_float(v[6]) * inv, __uint_as_float(v[7]) * inv);
          o1.x = pack2(__uint_as_float(v[8]) * inv, __uint_as_float(v[9]) * inv);
          o1.y = pack2(__uint_as_float(v[10]) * inv, __uint_as_float(v[11]) * inv);
          o1.z = pack2(__uint_as_float(v[12]) * inv, __uint_as_float(v[13]) * inv);
          o1.w = pack2(__uint_as_float(v[14]) * inv, __uint_as_float(v[15]) * inv);
          const int ch = c0 >> 3;
          *reinterpret_cast<uint4*>(line + (((ch) ^ (row & 7)) << 4)) = o0;
          *reinterpret_cast<uint4*>(line + (((ch + 1) ^ (row & 7)) << 4)) = o1;
        }
        // O and the statistics of this unit are consumed: the next P V / the next softmax of this parity may proceed
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&e_done[n & 1u]);
        if (qi < p.L) p.lse[static_cast<size_t>(u) * p.L + qi] = st2.y;
        fence_proxy_async_smem();
        named_bar_sync(2, 128);
        if (storer) {
          tma_store_3d(&tm_out, tile, h * HD, t * 128, b);   // rows past the sequence end are clipped
          tma_store_commit();
        }
      }
    }
    if (storer) tma_store_wait_all<0>();
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------------------------
// A unit = (key tile j of 128 keys, query chunk a of <= 64 queries).  Two TMEM buffers of 128 columns (S^T | dP^T, 64
// columns each) ping-pong, so the contractions of unit n + 1 run while the element-wise warps work on unit n; two groups
// of four warps (one warp per TMEM lane quarter) share every unit — group g owns query columns [32g, 32g + 32), each
// thread one key row — which halves the S^T / dP^T -> P^T / dS^T turn-around the MMA warp waits for (93.4 -> 89.9 us
// against whole units alternating between the groups).  delta = rowsum(dO * O) and the base-2 LSE of the NEXT head are
// prepared by a dedicated warp straight from global memory while the current head computes (89.9 -> 84.3 us).
constexpr int BWD_THREADS = 352;            // warp 0: TMA, warp 1: MMA (+ TMEM alloc), warps 2-5 / 6-9: element-wise group 0 / 1, warp 10: delta
constexpr int OFF_Q = 0;
constexpr int OFF_DO = OFF_Q + KV_BYTES;             //  26624
constexpr int OFF_K = OFF_DO + KV_BYTES;             //  53248
constexpr int OFF_V = OFF_K + Q2_BYTES;              //  86016
constexpr int OFF_DS = OFF_V + Q2_BYTES;             // 118784: dS [128 keys][256 queries] bf16 = 4 column blocks, MN-major SW128
constexpr int DS_BLOCK = 128 * ROW_BYTES;            //  16384: 64 queries x 128 key rows
constexpr int OFF_ST = OFF_DS + 4 * DS_BLOCK;        // 184320: two output staging tiles [128 rows][64] bf16 (one per group)
constexpr int ST_TILE = 128 * ROW_BYTES;             //  16384
constexpr int OFF_VEC = OFF_ST + 2 * ST_TILE;        // 217088: two buffers of { lse2[208], delta[208] } (head parity)
constexpr int OFF_BAR = OFF_VEC + 4 * LP * 4;        // 220416
constexpr int BWD_SMEM = OFF_BAR + 128 + 1024;
constexpr uint32_t ACC_DV = 256, ACC_DK = 320, ACC_DQ = 384;   // TMEM columns; buffers: S^T at 128 b, dP^T at 128 b + 64

struct BwdParams {
  const __nv_bfloat16* out;
  const __nv_bfloat16* d_out;
  const float* lse;
  __nv_bfloat16* d_qkv;
  int B, L, H, C, batch_first;
  int l_pad;        // L rounded up to 16
  int nj;           // key tiles of 128 (1 or 2)
  int nq;           // query chunks of 64 (1..4)
  int total;
  float scale, scale_log2;
  int prof;         // debug: accumulate per-role wait / work cycles into g_bwd_prof (ffm_attention_bwd_profile)
};

// debug counters, 16 per CTA: see ffm_attention_bwd_profile
__device__ long long g_bwd_prof[160 * 16];
#define ATT_TIMED(acc, stmt)                          \
  do {                                                \
    if (p.prof) {                                     \
      const long long _t = clock64();                 \
      stmt;                                           \
      acc += clock64() - _t;                          \
    } else {                                          \
      stmt;                                           \
    }                                                 \
  } while (0)

__device__ __forceinline__ float4 lds128(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
  return v;
}

// TMEM columns of K step k (16 queries) of a chunk: bf16 pairs are packed over the columns their fp32 source occupied
// TMEM column of the packed bf16 block for query columns [16k, 16k + 16) of a unit: each element-wise group writes its
// two blocks over score columns it has itself already read ([0, 16) for group 0, [32, 48) for group 1)
__device__ __forceinline__ uint32_t packed_col(int k) { return static_cast<uint32_t>(8 * k + (k >= 2 ? 16 : 0)); }

// accumulator row (fp32, 64 columns from `acc`) x `sc` -> 64 bf16 into row `row` of a SW128 staging tile
__device__ __forceinline__ void stage_row64(uint32_t acc, float sc, uint8_t* tile, int row) {
  uint8_t* line = tile + row * ROW_BYTES;
#pragma unroll
  for (int c0 = 0; c0 < HD; c0 += 16) {
    uint32_t v[16];
    tmem_ld16(acc + c0, v);
    tmem_ld_wait();
    uint4 o0, o1;
    o0.x = pack2(__uint_as_float(v[0]) * sc, __uint_as_float(v[1]) * sc);
    o0.y = pack2(__uint_as_float(v[2]) * sc, __uint_as_float(v[3]) * sc);
    o0.z = pack2(__uint_as_float(v[4]) * sc, __uint_as_float(v[5]) * sc);
    o0.w = pack2(__uint_as_float(v[6]) * sc, __uint_as_float(v[7]) * sc);
    o1.x = pack2(__uint_as_float(v[8]) * sc, __uint_as_float(v[9]) * sc);
    o1.y = pack2(__uint_as_float(v[10]) * sc, __uint_as_float(v[11]) * sc);
    o1.z = pack2(__uint_as_float(v[12]) * sc, __uint_as_float(v[13]) * sc);
    o1.w = pack2(__uint_as_float(v[14]) * sc, __uint_as_float(v[15]) * sc);
    const int ch = c0 >> 3;
    *reinterpret_cast<uint4*>(line + (((ch) ^ (row & 7)) << 4)) = o0;
    *reinterpret_cast<uint4*>(line + (((ch + 1) ^ (row & 7)) << 4)) = o1;
  }
}

template <bool CAUSAL>
__global__ void __launch_bounds__(BWD_THREADS, 1)
attention_bwd_tc_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                        const __grid_constant__ CUtensorMap tm_do, const __grid_constant__ CUtensorMap tm_dqkv,
                        const BwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  float* lse2_s = reinterpret_cast<float*>(smem + OFF_VEC);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* ld_full = bars;          // operands of a head landed
  uint64_t* ld_free = bars + 1;      // every MMA of the head completed: operands may be overwritten
  uint64_t* sdp_full = bars + 2;     // [2] S^T and dP^T of a unit in TMEM buffer b
  uint64_t* pds_full = bars + 4;     // [2] P^T / dS^T of buffer b written (TMEM + shared dS block)
  uint64_t* stage_free = bars + 6;   // [2] dQ MMAs of query half h completed: its shared dS blocks may be rewritten
  uint64_t* dvk_full = bars + 8;     //     dV_j / dK_j complete
  uint64_t* dq_full = bars + 9;      // [2] dQ of query half h complete
  uint64_t* delta_full = bars + 11;  // [2] delta / lse2 buffer (head parity) written by the delta warp
  uint64_t* delta_free = bars + 13;  // [2] ... no longer read by the element-wise warps
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
  const int nj = p.nj, nq = p.nq;
  const int units = nj * nq;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    tma_prefetch_desc(&tm_do);
    tma_prefetch_desc(&tm_dqkv);
    mbar_init(ld_full, 1);
    mbar_init(ld_free, 1);
    mbar_init(dvk_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sdp_full[i], 1);
      mbar_init(&pds_full[i], 8);
      mbar_init(&stage_free[i], 1);
      mbar_init(&dq_full[i], 1);
      mbar_init(&delta_full[i], 32);
      mbar_init(&delta_free[i], 8);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ============================== TMA producer ==============================
    const uint32_t q_bytes = static_cast<uint32_t>(p.l_pad) * ROW_BYTES, k_bytes = static_cast<uint32_t>(128 * nj) * ROW_BYTES;
    int it = 0;
    long long w_free = 0;
    for (int u = blockIdx.x; u < p.total; u += gridDim.x, ++it) {
      ATT_TIMED(w_free, mbar_wait_uniform(ld_free, (it & 1) ^ 1));
      if (elect_one()) {
        const int b = u / p.H, h = u - b * p.H;
        mbar_arrive_expect_tx(ld_full, 2 * q_bytes + 2 * k_bytes);
        tma_load_3d(smem + OFF_Q, &tm_q, ld_full, h * HD, 0, b);
        tma_load_3d(smem + OFF_DO, &tm_do, ld_full, h * HD, 0, b);
        tma_load_3d(smem + OFF_K, &tm_kv, ld_full, p.C + h * HD, 0, b);
        tma_load_3d(smem + OFF_V, &tm_kv, ld_full, 2 * p.C + h * HD, 0, b);
        const int un = u + static_cast<int>(gridDim.x);      // the operand tiles are single-buffered: at least pull the
        if (un < p.total) {                                  // next head's tiles into L2 while this one computes
          const int bn = un / p.H, hn = un - bn * p.H;
          tma_prefetch_3d(&tm_q, hn * HD, 0, bn);
          tma_prefetch_3d(&tm_do, hn * HD, 0, bn);
          tma_prefetch_3d(&tm_kv, p.C + hn * HD, 0, bn);
          tma_prefetch_3d(&tm_kv, 2 * p.C + hn * HD, 0, bn);
        }
      }
      __syncwarp();
    }
    if (p.prof && lane == 0) g_bwd_prof[blockIdx.x * 16 + 3] = w_free;
  } else if (warp == 1) {
    // ============================== MMA issuer ==============================
    const uint32_t sb = smem_u32(smem);
    const long long t_begin = clock64();
    long long w_ld = 0, w_pds = 0;
    const uint32_t idesc_acc = umma_idesc_bf16(128, HD) | UMMA_IDESC_B_MN;                      // A in TMEM
    const uint32_t idesc_dq = umma_idesc_bf16(128, HD) | UMMA_IDESC_A_MN | UMMA_IDESC_B_MN;     // A = shared dS blocks

    // S^T = K_j Q_a^T and dP^T = V_j dO_a^T of unit w into buffer (n & 1)
    auto issue_sdp = [&](int w, uint32_t n) {
      const int j = w / nq, a = w - j * nq;
      const int na = (p.l_pad - 64 * a) < 64 ? (p.l_pad - 64 * a) : 64;
      const uint32_t idesc_sn = umma_idesc_bf16(128, na);
      const uint32_t buf = tmem_base + 128u * (n & 1u);
      if (elect_one()) {
        const uint64_t kd = umma_desc_sw128(sb + OFF_K + j * (128 * ROW_BYTES));
        const uint64_t vd = umma_desc_sw128(sb + OFF_V + j * (128 * ROW_BYTES));
        const uint64_t qd = umma_desc_sw128(sb + OFF_Q + a * (64 * ROW_BYTES));
        const uint64_t dd = umma_desc_sw128(sb + OFF_DO + a * (64 * ROW_BYTES));
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(buf, kd + 2u * k, qd + 2u * k, idesc_sn, k != 0 ? 1u : 0u);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) umma_bf16(buf + 64, vd + 2u * k, dd + 2u * k, idesc_sn, k != 0 ? 1u : 0u);
        umma_commit(&sdp_full[n & 1u]);
      }
      __syncwarp();
    };

    int it = 0;
    uint32_t n = 0;                                         // running unit counter (same enumeration in every role)
    for (int u = blockIdx.x; u < p.total; u += gridDim.x, ++it) {
      ATT_TIMED(w_ld, mbar_wait_uniform(ld_full, it & 1));
      tc_fence_after();
      issue_sdp(0, n);
      for (int w = 0; w < units; ++w, ++n) {
        const int j = w / nq, a = w - j * nq;
        if (w + 1 < units) issue_sdp(w + 1, n + 1);          // runs under the element-wise work of unit w
        ATT_TIMED(w_pds, mbar_wait_uniform(&pds_full[n & 1u], (n >> 1) & 1u));
        tc_fence_after();
        if (elect_one()) {
          const int na = (p.l_pad - 64 * a) < 64 ? (p.l_pad - 64 * a) : 64;
          const uint32_t buf = tmem_base + 128u * (n & 1u);
          const int ks = na / 16;
          for (int k = 0; k < ks; ++k)                       // dV_j += P^T dO_a
            umma_bf16_ts(tmem_base + ACC_DV, buf + packed_col(k),
                         umma_desc_sw128_mn(sb + OFF_DO + (64 * a + 16 * k) * ROW_BYTES, 16), idesc_acc,
                         (a | k) != 0 ? 1u : 0u);
          for (int k = 0; k < ks; ++k)                       // dK_j += dS^T Q_a
            umma_bf16_ts(tmem_base + ACC_DK, buf + 64 + packed_col(k),
                         umma_desc_sw128_mn(sb + OFF_Q + (64 * a + 16 * k) * ROW_BYTES, 16), idesc_acc,
                         (a | k) != 0 ? 1u : 0u);
          if ((a & 1) == 1 || a == nq - 1) {                 // both chunks of query half h are staged: dQ_h += dS K_j
            const int h = a >> 1;
            const int kk = ((p.L - 128 * j < 128 ? p.L - 128 * j : 128) + 15) / 16;
            for (int k = 0; k < kk; ++k)
              umma_bf16(tmem_base + ACC_DQ + 64 * h, umma_desc_sw128_mn(sb + OFF_DS + 2 * h * DS_BLOCK + k * 2048, DS_BLOCK),
                        umma_desc_sw128_mn(sb + OFF_K + (128 * j + 16 * k) * ROW_BYTES, 16), idesc_dq,
                        (j | k) != 0 ? 1u : 0u);
            umma_commit(&stage_free[h]);
            if (j == nj - 1) umma_commit(&dq_full[h]);
          }
          if (a == nq - 1) umma_commit(dvk_full);
          if (w == units - 1) umma_commit(ld_free);
        }
        __syncwarp();
      }
    }
    if (p.prof && lane == 0) {
      g_bwd_prof[blockIdx.x * 16 + 0] = clock64() - t_begin;
      g_bwd_prof[blockIdx.x * 16 + 1] = w_ld;
      g_bwd_prof[blockIdx.x * 16 + 2] = w_pds;
    }
  } else if (warp == 10) {
    // ============================== delta warp ==============================
    // delta[q] = sum_d dO[q][d] O[q][d] and lse2[q] of every head of this CTA, straight from global memory, one head ahead
    // of the element-wise warps (two buffers by head parity); +inf masks padded queries
    int it = 0;
    for (int u = blockIdx.x; u < p.total; u += gridDim.x, ++it) {
      if (it >= 2) mbar_wait_uniform(&delta_free[it & 1], ((it >> 1) - 1) & 1);
      const int b = u / p.H, h = u - b * p.H;
      float* lse_b = lse2_s + (it & 1) * 2 * LP;
      float* del_b = lse_b + LP;
      for (int r = static_cast<int>(lane); r < LP; r += 32) {
        float d = 0.f, l2 = __int_as_float(0x7f800000);
        if (r < p.L) {
          const size_t tok = p.batch_first ? (static_cast<size_t>(b) * p.L + r) : (static_cast<size_t>(r) * p.B + b);
          const uint4* orow = reinterpret_cast<const uint4*>(p.out + tok * p.C + h * HD);
          const uint4* drow = reinterpret_cast<const uint4*>(p.d_out + tok * p.C + h * HD);
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint4 ov = __ldg(orow + c);
            const uint4 dv = __ldg(drow + c);
            const __nv_bfloat162* o2 = reinterpret_cast<const __nv_bfloat162*>(&ov);
            const __nv_bfloat162* d2 = reinterpret_cast<const __nv_bfloat162*>(&dv);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 of = __bfloat1622float2(o2[e]), df = __bfloat1622float2(d2[e]);
              d = fmaf(of.x, df.x, d);
              d = fmaf(of.y, df.y, d);
            }
          }
          l2 = p.lse[static_cast<size_t>(u) * p.L + r];
        }
        del_b[r] = d;
        lse_b[r] = l2;
      }
      mbar_arrive(&delta_full[it & 1]);          // every lane releases its own rows
    }
  } else {
    // ============================== element-wise warps ==============================
    long long c_ld = 0, c_pre = 0, c_sdp = 0, c_stage = 0, c_epw = 0, c_work = 0, c_drain = 0, c_end = 0;
    const uint32_t quarter = warp & 3u;
    const uint32_t grp = (warp - 2u) >> 2;                    // handles query columns [32 grp, 32 grp + 32) of every unit
    const int row = static_cast<int>(quarter * 32u + lane);  // key row inside tile j / query row inside a half
    const int tid_s = static_cast<int>(threadIdx.x) - 64;    // 0..255
    const uint32_t lane_addr = (quarter * 32u) << 16;
    uint8_t* stage = smem + OFF_ST + grp * ST_TILE;          // this group's output staging tile
    const bool storer = quarter == 0 && lane == 0;           // issues the group's TMA stores
    // accumulator tile (128 rows x 64 fp32 at TMEM column `acc_col`) x sc -> bf16 -> staging tile -> one TMA store to
    // d_qkv[b, row0 .. row0 + 127, col0 .. col0 + 63]; rows past the sequence end are clipped by the hardware
    auto drain_tile = [&](uint32_t acc_col, float sc, int col0, int row0, int b) {
      if (storer) tma_store_wait_read<0>();                  // the previous store has finished reading the tile
      named_bar_sync(2 + grp, 128);
      stage_row64(tmem_base + lane_addr + acc_col, sc, stage, row);
      fence_proxy_async_smem();
      named_bar_sync(2 + grp, 128);
      if (storer) {
        tma_store_3d(&tm_dqkv, stage, col0, row0, b);
        tma_store_commit();
      }
    };
    int it = 0;
    uint32_t n = 0;
    for (int u = blockIdx.x; u < p.total; u += gridDim.x, ++it) {
      const int b = u / p.H, h = u - b * p.H;
      ATT_TIMED(c_ld, mbar_wait(ld_full, it & 1, 300));
      const long long t_pre = clock64();
      // delta / lse2 of this head were prepared by the delta warp (buffer it & 1)
      mbar_wait(&delta_full[it & 1], (it >> 1) & 1, 310);
      const uint32_t lse_addr = smem_u32(lse2_s + (it & 1) * 2 * LP), delta_addr = lse_addr + LP * 4;
      c_pre += clock64() - t_pre;

      for (int w = 0; w < units; ++w, ++n) {
        // BOTH groups work on every unit (half of its query columns each): the S^T / dP^T -> P^T / dS^T turn-around that
        // the MMA warp waits for is half as long as with whole units alternating between the groups
        const int j = w / nq, a = w - j * nq;
        const int na = (p.l_pad - 64 * a) < 64 ? (p.l_pad - 64 * a) : 64;
        const int key = 128 * j + row;
        const bool kvalid = key < p.L;
        const uint32_t buf = tmem_base + lane_addr + 128u * (n & 1u);
        ATT_TIMED(c_sdp, mbar_wait(&sdp_full[n & 1u], (n >> 1) & 1u, 400));
        tc_fence_after();
        if (a == 0 && j > 0) {
          // dV_{j-1} / dK_{j-1} are complete: drain them before this unit's first dV / dK MMA overwrites the accumulators
          ATT_TIMED(c_epw, mbar_wait(dvk_full, static_cast<uint32_t>(it * nj + j - 1) & 1u, 500));
          tc_fence_after();
          const long long t_dr = clock64();
          if (grp == 0) drain_tile(ACC_DV, 1.0f, 2 * p.C + h * HD, 128 * (j - 1), b);
          else drain_tile(ACC_DK, p.scale, p.C + h * HD, 128 * (j - 1), b);
          c_drain += clock64() - t_dr;
        }
        // the shared dS blocks of this query half were last read by the dQ MMAs of the previous key tile (or head)
        {
          const int done = it * nj + j - 1;
          if (done >= 0) ATT_TIMED(c_stage, mbar_wait(&stage_free[a >> 1], static_cast<uint32_t>(done) & 1u, 600));
        }
        const long long t_work = clock64();
        uint8_t* blk = smem + OFF_DS + a * DS_BLOCK + row * ROW_BYTES;
        const int c_end_col = na < 32 * static_cast<int>(grp) + 32 ? na : 32 * static_cast<int>(grp) + 32;
        for (int c0 = 32 * static_cast<int>(grp); c0 < c_end_col; c0 += 16) {
          uint32_t sv[16], dv[16];
          tmem_ld16(buf + c0, sv);
          tmem_ld16(buf + 64 + c0, dv);
          const int q0 = 64 * a + c0;
          float l2[16], dl[16];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const float4 x = lds128(lse_addr + (q0 + 4 * e) * 4), y = lds128(delta_addr + (q0 + 4 * e) * 4);
            l2[4 * e] = x.x; l2[4 * e + 1] = x.y; l2[4 * e + 2] = x.z; l2[4 * e + 3] = x.w;
            dl[4 * e] = y.x; dl[4 * e + 1] = y.y; dl[4 * e + 2] = y.z; dl[4 * e + 3] = y.w;
          }
          tmem_ld_wait();
          uint32_t pp[8], dd[8];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float pv[2], gv[2];
#pragma unroll
            for (int t = 0; t < 2; ++t) {
              const int i = 2 * e + t;
              float pr = fast_exp2(fmaf(__uint_as_float(sv[i]), p.scale_log2, -l2[i]));
              if (CAUSAL) pr = (q0 + i >= key) ? pr : 0.f;
              pv[t] = pr;
              gv[t] = pr * (__uint_as_float(dv[i]) - dl[i]);
            }
            pp[e] = kvalid ? pack2(pv[0], pv[1]) : 0u;
            dd[e] = kvalid ? pack2(gv[0], gv[1]) : 0u;
          }
          tmem_st8(buf + packed_col(c0 >> 4), pp);
          tmem_st8(buf + 64 + packed_col(c0 >> 4), dd);
          // dS (not transposed) for dQ: [key row][query contiguous]: two 16-byte chunks of this row's 128-byte line
          const int ch = c0 >> 3;
          *reinterpret_cast<uint4*>(blk + (((ch) ^ (row & 7)) << 4)) = make_uint4(dd[0], dd[1], dd[2], dd[3]);
          *reinterpret_cast<uint4*>(blk + (((ch + 1) ^ (row & 7)) << 4)) = make_uint4(dd[4], dd[5], dd[6], dd[7]);
        }
        tmem_st_wait();
        fence_proxy_async_smem();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[n & 1u]);
        c_work += clock64() - t_work;
      }
      // ---- epilogues of the head: group 0 drains dV / dK of the last key tile, group 1 both query halves of dQ ----
      if (grp == 0) {
        ATT_TIMED(c_epw, mbar_wait(dvk_full, static_cast<uint32_t>(it * nj + nj - 1) & 1u, 700));
        tc_fence_after();
        const long long t_dr = clock64();
        drain_tile(ACC_DV, 1.0f, 2 * p.C + h * HD, 128 * (nj - 1), b);
        drain_tile(ACC_DK, p.scale, p.C + h * HD, 128 * (nj - 1), b);
        c_drain += clock64() - t_dr;
      } else {
        for (int hq = 0; hq < (nq + 1) / 2; ++hq) {
          ATT_TIMED(c_epw, mbar_wait(&dq_full[hq], it & 1, 800 + hq));
          tc_fence_after();
          const long long t_dr = clock64();
          drain_tile(ACC_DQ + 64 * hq, p.scale, h * HD, 128 * hq, b);
          c_drain += clock64() - t_dr;
        }
      }
      tc_fence_before();
      ATT_TIMED(c_end, named_bar_sync(1, 256));   // accumulators drained, delta / lse2 dead: the next head may proceed
      if (lane == 0) mbar_arrive(&delta_free[it & 1]);
    }
    if (storer) tma_store_wait_all<0>();
    if (p.prof && lane == 0 && quarter == 0) {      // one warp per group reports: slots 4.. (group 0), 10.. (group 1)
      long long* o = g_bwd_prof + blockIdx.x * 16 + (grp == 0 ? 4 : 10);
      o[0] = c_ld + c_pre; o[1] = c_sdp; o[2] = c_stage + c_epw; o[3] = c_work; o[4] = c_drain; o[5] = c_end;
    }
  }
  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// ---------------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------------
// [batch][token][channel] view of a packed bf16 activation: box = {64 channels, box_rows tokens, 1 sample}
static int make_map_tokens(CUtensorMap* out, const void* ptr, int B, int L, int width, int batch_first, int box_rows) {
  PFN_encodeTiled enc = tensor_map_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled driver entry point not available");
    return FFM_ERR_CUDA;
  }
  const cuuint64_t row = static_cast<cuuint64_t>(width) * 2;
  cuuint64_t dims[3] = {static_cast<cuuint64_t>(width), static_cast<cuuint64_t>(L), static_cast<cuuint64_t>(B)};
  cuuint64_t strides[2] = {batch_first ? row : row * B, batch_first ? row * L : row};
  cuuint32_t box[3] = {HD, static_cast<cuuint32_t>(box_rows), 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled (attention, %d x %d x %d, box %d rows) failed: %d", B, L, width, box_rows, (int)r);
    return FFM_ERR_CUDA;
  }
  return FFM_OK;
}

static int set_smem_once() {
  static thread_local int attr_dev = -1;
  int dev = 0;
  FFM_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev != attr_dev) {
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_fwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, FWD_SMEM));
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
    FFM_CHECK_CUDA(cudaFuncSetAttribute(attention_bwd_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, BWD_SMEM));
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
  FFM_CHECK_ARG((reinterpret_cast<uintptr_t>(qkv) & 15u) == 0 && (reinterpret_cast<uintptr_t>(out) & 15u) == 0,
                "ffm_attention_fwd: qkv / out must be 16-byte aligned");
  int rc = att::set_smem_once();
  if (rc != FFM_OK) return rc;
  att::FwdParams p;
  p.lse = lse;
  p.B = B; p.L = L; p.H = H; p.C = H * att::HD; p.batch_first = batch_first ? 1 : 0;
  p.lk_pad = (L + 15) / 16 * 16;
  p.n_tiles = L > 128 ? 2 : 1;
  p.total = B * H;
  p.scale_log2 = att::LOG2E / sqrtf(static_cast<float>(att::HD));
  CUtensorMap tm_q, tm_kv, tm_out;
  if ((rc = att::make_map_tokens(&tm_q, qkv, B, L, 3 * p.C, p.batch_first, 128 * p.n_tiles))) return rc;
  if ((rc = att::make_map_tokens(&tm_kv, qkv, B, L, 3 * p.C, p.batch_first, p.lk_pad))) return rc;
  if ((rc = att::make_map_tokens(&tm_out, out, B, L, p.C, p.batch_first, 128))) return rc;
  const int grid = p.total < num_sms() ? p.total : num_sms();
  if (causal)
    att::attention_fwd_tc_kernel<true><<<grid, att::FWD_THREADS, att::FWD_SMEM, stream>>>(tm_q, tm_kv, tm_out, p);
  else
    att::attention_fwd_tc_kernel<false><<<grid, att::FWD_THREADS, att::FWD_SMEM, stream>>>(tm_q, tm_kv, tm_out, p);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_attention_bwd(const void* qkv, const void* out, const void* d_out, const float* lse, void* d_qkv, int B, int L,
                      int H, int head_dim, int causal, int batch_first, cudaStream_t stream) {
  FFM_CHECK_ARG(qkv && out && d_out && lse && d_qkv, "ffm_attention_bwd: null pointer argument");
  FFM_CHECK_ARG(head_dim == att::HD, "ffm_attention_bwd: head dimension must be %d", att::HD);
  FFM_CHECK_ARG(B >= 1 && H >= 1 && L >= 1 && L <= att::LP, "ffm_attention_bwd: 1 <= L <= %d", att::LP);
  FFM_CHECK_ARG(((reinterpret_cast<uintptr_t>(qkv) | reinterpret_cast<uintptr_t>(out) | reinterpret_cast<uintptr_t>(d_out) |
                  reinterpret_cast<uintptr_t>(d_qkv)) & 15u) == 0,
                "ffm_attention_bwd: device pointers must be 16-byte aligned");
  int rc = att::set_smem_once();
  if (rc != FFM_OK) return rc;
  att::BwdParams p;
  p.out = static_cast<const __nv_bfloat16*>(out);
  p.d_out = static_cast<const __nv_bfloat16*>(d_out);
  p.lse = lse;
  p.d_qkv = static_cast<__nv_bfloat16*>(d_qkv);
  p.B = B; p.L = L; p.H = H; p.C = H * att::HD; p.batch_first = batch_first ? 1 : 0;
  p.l_pad = (L + 15) / 16 * 16;
  p.nj = L > 128 ? 2 : 1;
  p.nq = (p.l_pad + 63) / 64;
  p.total = B * H;
  p.scale = 1.0f / sqrtf(static_cast<float>(att::HD));
  p.scale_log2 = p.scale * att::LOG2E;
  {
    static int prof = -1;
    if (prof < 0) { const char* e = getenv("FFM_ATT_PROF"); prof = (e != nullptr && e[0] == '1') ? 1 : 0; }
    p.prof = prof;
  }
  CUtensorMap tm_q, tm_kv, tm_do, tm_dqkv;
  if ((rc = att::make_map_tokens(&tm_q, qkv, B, L, 3 * p.C, p.batch_first, p.l_pad))) return rc;
  if ((rc = att::make_map_tokens(&tm_kv, qkv, B, L, 3 * p.C, p.batch_first, 128 * p.nj))) return rc;
  if ((rc = att::make_map_tokens(&tm_do, d_out, B, L, p.C, p.batch_first, p.l_pad))) return rc;
  if ((rc = att::make_map_tokens(&tm_dqkv, d_qkv, B, L, 3 * p.C, p.batch_first, 128))) return rc;
  const int grid = p.total < num_sms() ? p.total : num_sms();
  if (causal)
    att::attention_bwd_tc_kernel<true><<<grid, att::BWD_THREADS, att::BWD_SMEM, stream>>>(tm_q, tm_kv, tm_do, tm_dqkv, p);
  else
    att::attention_bwd_tc_kernel<false><<<grid, att::BWD_THREADS, att::BWD_SMEM, stream>>>(tm_q, tm_kv, tm_do, tm_dqkv, p);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  if (p.prof) {      // debug only: per-role cycle counters of this launch, averaged over the CTAs
    static long long h[160 * 16];
    FFM_CHECK_CUDA(cudaStreamSynchronize(stream));
    FFM_CHECK_CUDA(cudaMemcpyFromSymbol(h, att::g_bwd_prof, sizeof(h)));
    double avg[16] = {0};
    for (int c = 0; c < grid; ++c)
      for (int k = 0; k < 16; ++k) avg[k] += static_cast<double>(h[c * 16 + k]) / grid;
    fprintf(stderr, "[att bwd prof] total %.0f | mma: wait ld %.0f pds %.0f | tma: wait free %.0f | g0: ld+pre %.0f sdp %.0f "
            "stage+ep-wait %.0f work %.0f drain %.0f end %.0f | g1: ld+pre %.0f sdp %.0f stage+ep-wait %.0f work %.0f drain "
            "%.0f end %.0f\n", avg[0], avg[1], avg[2], avg[3], avg[4], avg[5], avg[6], avg[7], avg[8], avg[9], avg[10], avg[11],
            avg[12], avg[13], avg[14], avg[15]);
  }
  return FFM_OK;
}

}  // extern "C"
