// Fused SVLoRA / FairLoRA linear on Blackwell tensor cores (sm_100a).
//
// Computes, for row-major bf16 operands,
//
//     H[T,RP]  = X[T,K] · Aside[RP,K]^T                                   (fp32, side output)
//     OUT[T,N] = epi( X · Wmat[N,K]^T + bias + (H ⊙ s_rows[sample(t)]) · Bside[N,RP]^T )
//
// in ONE persistent kernel.  This single contraction serves both directions of the reference's
// FairLoRALinear (trainers/GLP_OT_SVLoRA.py:450-482):
//   forward : X=x,  Wmat=W  [out,in],  Aside=A^T (pad), Bside=B^T (pad), s_rows = scaling·s_eff
//   backward: X=dy, Wmat=W^T[in,out],  Aside=B   (pad), Bside=A   (pad), s_rows = scaling·s_eff
//             -> OUT = dx, H = dy·B^T (un-scaled dz)
//
// Structure (one CTA per SM, 384 threads, static round-robin tile scheduler):
//   warp 0   : TMA producer  — X / Wmat / Aside k-slices into a 4-stage SW128 smem ring (warp-uniform loop, one
//                              elected lane issues)
//   warp 1   : MMA issuer    — one tcgen05.mma (M=128, N=192+16, K=16) per k-step: the W tile and the
//                              Aside tile are adjacent in smem, so a single UMMA accumulates both
//                              D (192 cols) and H (16 cols) into TMEM; later one K=16 "fix-up" UMMA
//                              D += Z · Bside^T with Z = bf16(H ⊙ s_rows) staged by the epilogue warps
//   warp 2   : TMEM allocator (512 columns = 2 accumulator stages x 256)
//   warps 4-11: epilogue     — tcgen05.ld H -> scale -> Z (SW32 smem) -> signal; then tcgen05.ld D
//                              -> +bias / QuickGELU / QuickGELU' -> bf16 -> per-warp SW64 staging -> per-warp TMA
//                              store (8 independent streams, no CTA-wide barrier)
// TMEM accumulators are double buffered so tile i's epilogue overlaps tile i+1's mainloop; the
// fix-up UMMA of tile i is slotted into tile i+1's k-loop as soon as Z(i) is ready.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include "../../include/ffm_b200.h"
#include "svlora_gemm.cuh"

namespace ffm {

// ----------------------------------------------------------------------------------------------
// error string + device info (shared by the whole library)
// ----------------------------------------------------------------------------------------------
static thread_local char g_last_error[512] = "";

void set_last_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_last_error, sizeof(g_last_error), fmt, ap);
  va_end(ap);
}

static std::atomic<long long> g_launches{0};
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// event profiling of the dominant kernel (bench.py roofline): records are appended at launch time
struct GemmRecord { cudaEvent_t start, stop; int T, K, N; };
static std::mutex g_prof_mutex;
static bool g_prof_enabled = false;
static std::vector<GemmRecord> g_prof_records;

int gemm_profile_begin(GemmProfileScope* sc, cudaStream_t stream) {
  {
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    sc->active = g_prof_enabled && g_prof_records.size() < 65536;
  }
  if (!sc->active) return FFM_OK;
  FFM_CHECK_CUDA(cudaEventCreate(&sc->start));
  FFM_CHECK_CUDA(cudaEventCreate(&sc->stop));
  FFM_CHECK_CUDA(cudaEventRecord(sc->start, stream));
  return FFM_OK;
}

int gemm_profile_end(GemmProfileScope* sc, int T, int K, int N, cudaStream_t stream) {
  if (!sc->active) return FFM_OK;
  FFM_CHECK_CUDA(cudaEventRecord(sc->stop, stream));
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  g_prof_records.push_back(GemmRecord{sc->start, sc->stop, T, K, N});
  return FFM_OK;
}

int num_sms() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ----------------------------------------------------------------------------------------------
// tile configuration
// ----------------------------------------------------------------------------------------------
constexpr int BM = 128;             // rows of X per tile      (UMMA M, one TMEM lane per row)
constexpr int BN = 192;             // rows of Wmat per tile   (output columns)
constexpr int BK = 64;              // k-slice per stage = one 128-byte swizzle row of bf16
constexpr int UMMA_K = 16;
constexpr int ACC_COLS = 256;       // TMEM columns per accumulator stage (BN + padded rank used)
constexpr int TMEM_COLS = 512;

constexpr int X_TILE_BYTES = BM * BK * 2;   // 16384
constexpr int W_TILE_BYTES = BN * BK * 2;   // 24576
constexpr int OUT_STAGING_BYTES = 8 * 2 * EPI_PIECE_BYTES;  // 8 epilogue warps x 2 buffers x 2 KB = 32768
constexpr int BIAS_BYTES = 8 * (BN / 2) * 4;          //  3072: per epilogue warp, the bias of its BN/2 columns

// Everything that depends on the padded adapter rank R (16: ViT recipes r <= 16; 32: the RN50 recipe r = 32).
// R = 16: Aside rides as 16 extra accumulator columns (UMMA N = 208), Z / Bside are SW32 tiles, 4-stage ring.
// R = 32: UMMA N = 224, Z / Bside are SW64 tiles, the fix-up is two K = 16 UMMAs, 3-stage ring (smem).
template <int R>
struct GemmCfg {
  static_assert(R == 16 || R == 32, "padded adapter rank must be 16 or 32");
  static constexpr int UMMA_N_MAIN = BN + R;                 // [D | H] in one instruction
  static constexpr int STAGES = (R == 16) ? 4 : 3;
  static constexpr int A_TILE_BYTES = R * BK * 2;            // 2048 / 4096
  static constexpr int STAGE_BYTES = X_TILE_BYTES + W_TILE_BYTES + A_TILE_BYTES;   // 43008 / 45056 (1024-multiples)
  static constexpr int Z_TILE_BYTES = BM * R * 2;            // 4096 / 8192
  static constexpr int BS_TILE_BYTES = BN * R * 2;           // 6144 / 12288
  static constexpr int OFF_STAGES = 0;
  static constexpr int OFF_OUT = OFF_STAGES + STAGES * STAGE_BYTES;
  static constexpr int OFF_Z = OFF_OUT + OUT_STAGING_BYTES;
  static constexpr int OFF_BS = OFF_Z + 2 * Z_TILE_BYTES;
  static constexpr int OFF_BIAS = OFF_BS + 2 * BS_TILE_BYTES;
  static constexpr int OFF_BAR = OFF_BIAS + BIAS_BYTES;
  static constexpr int NUM_BARS = 2 * STAGES + 5 * 2;
  static constexpr int SMEM_USED = OFF_BAR + NUM_BARS * 8 + 16;
  static constexpr int SMEM_BYTES = SMEM_USED + 1024;        // slack for manual 1024-B alignment
  static_assert(STAGE_BYTES % 1024 == 0, "stage must keep 1024-B alignment of SW128 tiles");
  static_assert(OFF_OUT % 1024 == 0 && OFF_Z % 1024 == 0 && OFF_BS % 1024 == 0, "tile alignment");
  static_assert(SMEM_BYTES <= 232448, "exceeds 227 KB of shared memory per CTA");
  static_assert(UMMA_N_MAIN <= ACC_COLS && UMMA_N_MAIN % 16 == 0, "accumulator stage");
};
static_assert((X_TILE_BYTES + W_TILE_BYTES) % 1024 == 0, "Aside tile must start on a swizzle atom");
static_assert(BN % (2 * EPI_PIECE_COLS) == 0, "epilogue pieces");

constexpr int NUM_THREADS = 384;   // 4 control warps (TMA, MMA, TMEM alloc, spare) + 8 epilogue warps
constexpr int EPI_THREADS = 256;
constexpr int Z_THREADS = 128;     // epilogue threads that write the Z tile (one per tile row)


// ADAPT = false: the same pipeline for a FROZEN linear without an adapter (in_proj / out_proj of the attention blocks):
// no Aside k-slices, UMMA N = 192, no H -> Z -> fix-up chain — the epilogue starts as soon as the mainloop commits.
template <int ACT, int R, bool ADAPT = true>
__global__ void __launch_bounds__(NUM_THREADS, 1)
svlora_gemm_kernel(const __grid_constant__ CUtensorMap tm_x, const __grid_constant__ CUtensorMap tm_w,
                   const __grid_constant__ CUtensorMap tm_a, const __grid_constant__ CUtensorMap tm_b,
                   const __grid_constant__ CUtensorMap tm_y, const __grid_constant__ CUtensorMap tm_y2,
                   const GemmParams p) {
  using C = GemmCfg<R>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, Z_TILE_BYTES = C::Z_TILE_BYTES;
  constexpr int BS_TILE_BYTES = C::BS_TILE_BYTES, NUM_BARS = C::NUM_BARS;
  constexpr int UMMA_N_MAIN = ADAPT ? C::UMMA_N_MAIN : BN;
  constexpr int LOAD_BYTES = ADAPT ? STAGE_BYTES : (X_TILE_BYTES + W_TILE_BYTES);
  constexpr int OFF_STAGES = C::OFF_STAGES, OFF_OUT = C::OFF_OUT, OFF_Z = C::OFF_Z, OFF_BS = C::OFF_BS;
  constexpr int OFF_BIAS = C::OFF_BIAS, OFF_BAR = C::OFF_BAR;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full_bar = bars;                   // [STAGES] TMA -> MMA
  uint64_t* empty_bar = bars + STAGES;         // [STAGES] MMA -> TMA
  uint64_t* h_full = bars + 2 * STAGES;        // [2] mainloop done: H (and D partial) in TMEM
  uint64_t* z_full = h_full + 2;               // [2] epilogue wrote Z tile
  uint64_t* d_full = z_full + 2;               // [2] fix-up UMMA done: D final
  uint64_t* tmem_empty = d_full + 2;           // [2] epilogue drained accumulator stage
  uint64_t* bs_full = tmem_empty + 2;          // [2] Bside tile landed
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + NUM_BARS);

  const uint32_t warp = threadIdx.x >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  const int num_tiles = p.m_tiles * p.n_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_x);
    tma_prefetch_desc(&tm_w);
    if (ADAPT) {
      tma_prefetch_desc(&tm_a);
      tma_prefetch_desc(&tm_b);
    }
    tma_prefetch_desc(&tm_y);
    if (p.has_pre) tma_prefetch_desc(&tm_y2);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&h_full[i], 1);
      mbar_init(&z_full[i], Z_THREADS);
      mbar_init(&d_full[i], 1);
      mbar_init(&tmem_empty[i], EPI_THREADS / 32);
      mbar_init(&bs_full[i], 1);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // =========================== TMA producer (whole warp, one elected lane issues) ===========================
    uint32_t stage = 0, phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile - m_blk * p.n_tiles;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        mbar_wait_uniform(&empty_bar[stage], phase ^ 1u);
        if (elect_one()) {
          uint8_t* st = smem + OFF_STAGES + stage * STAGE_BYTES;
          if (p.dbg & 2) {
            mbar_arrive(&full_bar[stage]);
          } else {
            mbar_arrive_expect_tx(&full_bar[stage], LOAD_BYTES);
            tma_load_2d(st, &tm_x, &full_bar[stage], kb * BK, m_blk * BM);
            tma_load_2d(st + X_TILE_BYTES, &tm_w, &full_bar[stage], kb * BK, n_blk * BN);
            if (ADAPT) tma_load_2d(st + X_TILE_BYTES + W_TILE_BYTES, &tm_a, &full_bar[stage], kb * BK, 0);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // =========================== MMA issuer (whole warp, one elected lane issues) ===========================
    // Everything below is warp-uniform: the only per-lane decision is elect_one() around the tcgen05 instructions
    // (an `if (lane == 0)` region makes nvcc wrap every tcgen05.mma in an ELECT/R2UR waterfall loop).
    constexpr uint32_t idesc_main = umma_idesc_bf16(BM, UMMA_N_MAIN);
    constexpr uint32_t idesc_fix = umma_idesc_bf16(BM, BN);
    const uint32_t stages_base = smem_u32(smem + OFF_STAGES);
    uint32_t stage = 0, phase = 0;
    int pend = -1;
    uint32_t pend_phase = 0;

    auto fixup = [&](int s, uint32_t ph) {
      // D[s] += Z[s] (128 x R) · Bside[s]^T (R x 192)
      mbar_wait_uniform(&bs_full[s], ph);
      tc_fence_after();
      if (elect_one()) {
        const uint32_t za = smem_u32(smem + OFF_Z + s * Z_TILE_BYTES), ba = smem_u32(smem + OFF_BS + s * BS_TILE_BYTES);
        if constexpr (R == 16) {
          umma_bf16(tmem_base + s * ACC_COLS, umma_desc_sw32(za), umma_desc_sw32(ba), idesc_fix, 1u);
        } else {   // rank 32: rows of 64 B (SW64), two K = 16 steps 32 B apart inside the swizzle row
          const uint64_t zd = umma_desc_sw64(za), bd = umma_desc_sw64(ba);
          umma_bf16(tmem_base + s * ACC_COLS, zd, bd, idesc_fix, 1u);
          umma_bf16(tmem_base + s * ACC_COLS, zd + 2u, bd + 2u, idesc_fix, 1u);
        }
        umma_commit(&d_full[s]);
      }
      __syncwarp();
    };

    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int n_blk = tile % p.n_tiles;
      const int s = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      mbar_wait_uniform(&tmem_empty[s], aph ^ 1u);   // epilogue drained tile it-2 (=> Bside[s], Z[s] free)
      tc_fence_after();
      if (ADAPT) {
        if (elect_one()) {
          mbar_arrive_expect_tx(&bs_full[s], BS_TILE_BYTES);
          tma_load_2d(smem + OFF_BS + s * BS_TILE_BYTES, &tm_b, &bs_full[s], 0, n_blk * BN);
        }
        __syncwarp();
      }

      const uint32_t d_tmem = tmem_base + s * ACC_COLS;
      for (int kb = 0; kb < p.k_blocks; ++kb) {
        if (ADAPT && pend >= 0) {
          // Z of the previous tile ready?  (vote keeps the decision warp-uniform for the compiler)
          if (__all_sync(0xffffffffu, mbar_test_wait(&z_full[pend], pend_phase))) {
            fixup(pend, pend_phase);
            pend = -1;
          }
        }
        mbar_wait_uniform(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t st = stages_base + stage * STAGE_BYTES;
          const uint64_t adesc = umma_desc_sw128(st);
          const uint64_t bdesc = umma_desc_sw128(st + X_TILE_BYTES);
          if (!(p.dbg & 1)) {
#pragma unroll
            for (int k = 0; k < BK / UMMA_K; ++k) {
              // advance 32 B along K inside the 128-B swizzle row: +2 in the (>>4) address field
              umma_bf16(d_tmem, adesc + 2u * k, bdesc + 2u * k, idesc_main, (kb | k) != 0 ? 1u : 0u);
            }
          }
          umma_commit(&empty_bar[stage]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      if (!ADAPT) {
        if (elect_one()) umma_commit(&d_full[s]);     // no fix-up: the accumulator is final
        __syncwarp();
        continue;
      }
      if (elect_one()) umma_commit(&h_full[s]);
      __syncwarp();
      if (pend >= 0) {
        mbar_wait_uniform(&z_full[pend], pend_phase);
        fixup(pend, pend_phase);
      }
      pend = s;
      pend_phase = aph;
    }
    if (ADAPT && pend >= 0) {
      mbar_wait_uniform(&z_full[pend], pend_phase);
      fixup(pend, pend_phase);
    }
  } else if (warp >= 4) {
    // =========================== epilogue (8 independent warps) ===========================
    // Two warps share each TMEM lane quarter (a warp may only touch lanes 32*(warpid%4)..+31): warp w takes the even
    // 32-column pieces of the tile, warp w+4 the odd ones.  One tile row per thread.  No CTA-wide barriers: each warp
    // stages and TMA-stores its own pieces (epi_store_piece).
    const uint32_t ew = warp - 4u;             // 0..7
    const uint32_t q = warp & 3u;
    const uint32_t half = ew >> 2;
    const uint32_t row = q * 32u + lane;
    const uint32_t lane_addr = (q * 32u) << 16;
    uint8_t* stage_w = smem + OFF_OUT + ew * (2 * EPI_PIECE_BYTES);
    float* bias_w = reinterpret_cast<float*>(smem + OFF_BIAS) + ew * (BN / 2);
    uint32_t unit = 0;                         // this warp's running count of TMA stores (buffer = unit & 1)
    constexpr int PIECES = BN / (2 * EPI_PIECE_COLS);   // pieces per warp per tile
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile / p.n_tiles;
      const int n_blk = tile - m_blk * p.n_tiles;
      const int s = it & 1;
      const uint32_t aph = (it >> 1) & 1u;
      const int grow = m_blk * BM + static_cast<int>(row);   // global row
      const int n0 = n_blk * BN;
      const uint32_t acc = tmem_base + lane_addr + s * ACC_COLS;

      // bias of this warp's columns (warp-private smem, read back as broadcasts)
#pragma unroll
      for (int pc = 0; pc < PIECES; ++pc) {
        const int col = n0 + (2 * pc + static_cast<int>(half)) * EPI_PIECE_COLS + static_cast<int>(lane);
        bias_w[pc * EPI_PIECE_COLS + lane] = (p.bias != nullptr && col < p.N) ? __ldg(p.bias + col) : 0.0f;
      }

      if (ADAPT && half == 0) {
        // ---- H -> Z (SW32 K-major A operand of the fix-up UMMA) ----
        // this row's scaled singular values first: their global-load latency hides behind the wait for H
        const int grow_c = grow < p.T ? grow : (p.T - 1);
        const int sample = ((grow_c / p.row_div) % p.b_prime) / p.num_slices;
        const float4* sr = reinterpret_cast<const float4*>(p.s_rows + static_cast<size_t>(sample) * R);
        float4 svv[R / 4];
#pragma unroll
        for (int j = 0; j < R / 4; ++j) svv[j] = __ldg(sr + j);
        mbar_wait(&h_full[s], aph, 600 + s);
        tc_fence_after();
        uint32_t hv[R];
        if constexpr (R == 16) tmem_ld16(acc + BN, hv);
        else tmem_ld32(acc + BN, hv);
        tmem_ld_wait();
        float zf[R];
#pragma unroll
        for (int j = 0; j < R / 4; ++j) {
          const float4 sv = svv[j];
          zf[4 * j + 0] = __uint_as_float(hv[4 * j + 0]) * sv.x;
          zf[4 * j + 1] = __uint_as_float(hv[4 * j + 1]) * sv.y;
          zf[4 * j + 2] = __uint_as_float(hv[4 * j + 2]) * sv.z;
          zf[4 * j + 3] = __uint_as_float(hv[4 * j + 3]) * sv.w;
        }
        // K-major A operand of the fix-up UMMA: row r at r * (2R) bytes; R = 16: SW32 (16-B chunk c at c ^ ((r>>2)&1)),
        // R = 32: SW64 (chunk c at c ^ ((r>>1)&3))
        uint8_t* zrow = smem + OFF_Z + s * Z_TILE_BYTES + row * (2u * R);
        const uint32_t sw = (R == 16) ? ((row >> 2) & 1u) : ((row >> 1) & 3u);
        uint4 zc[R / 8];
#pragma unroll
        for (int c = 0; c < R / 8; ++c) {
          zc[c].x = pack_bf16x2(zf[8 * c + 0], zf[8 * c + 1]);
          zc[c].y = pack_bf16x2(zf[8 * c + 2], zf[8 * c + 3]);
          zc[c].z = pack_bf16x2(zf[8 * c + 4], zf[8 * c + 5]);
          zc[c].w = pack_bf16x2(zf[8 * c + 6], zf[8 * c + 7]);
          *reinterpret_cast<uint4*>(zrow + ((static_cast<uint32_t>(c) ^ sw) << 4)) = zc[c];
        }
        fence_proxy_async_smem();
        mbar_arrive(&z_full[s]);
        // side outputs to HBM after the signal: they are off the tile's critical chain
        if (n_blk == 0 && grow < p.T) {
          if (p.h_out != nullptr) {
            float4* ho = reinterpret_cast<float4*>(p.h_out + static_cast<size_t>(grow) * R);
#pragma unroll
            for (int j = 0; j < R / 4; ++j)
              ho[j] = make_float4(__uint_as_float(hv[4 * j]), __uint_as_float(hv[4 * j + 1]),
                                  __uint_as_float(hv[4 * j + 2]), __uint_as_float(hv[4 * j + 3]));
          }
          if (p.z_out != nullptr) {
            uint4* zo = reinterpret_cast<uint4*>(p.z_out + static_cast<size_t>(grow) * R);
#pragma unroll
            for (int c = 0; c < R / 8; ++c) zo[c] = zc[c];
          }
        }
      }
      __syncwarp();                              // bias_w visible to the whole warp

      // ---- D -> OUT ----
      EpiAux aux_cur, aux_nxt;
      epi_load_aux<ACT>(p, aux_nxt, grow, n0 + static_cast<int>(half) * EPI_PIECE_COLS);   // piece 0, before the wait
      mbar_wait(&d_full[s], aph, 700 + s);
      tc_fence_after();
#pragma unroll 1
      for (int pc = 0; pc < PIECES; ++pc) {
        aux_cur = aux_nxt;
        if (pc + 1 < PIECES)
          epi_load_aux<ACT>(p, aux_nxt, grow, n0 + (2 * (pc + 1) + static_cast<int>(half)) * EPI_PIECE_COLS);
        const int cc = (2 * pc + static_cast<int>(half)) * EPI_PIECE_COLS;   // first tile column of this piece
        uint32_t v[32];
        tmem_ld32(acc + cc, v);
        tmem_ld_wait();
        if (pc == PIECES - 1) {
          // accumulator stage fully read by this warp: hand it back to the MMA warp
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tmem_empty[s]);
        }
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) + bias_w[pc * EPI_PIECE_COLS + j];
        epi_store_piece<ACT>(p, f, aux_cur, &tm_y, &tm_y2, stage_w, unit, lane, grow, n0 + cc, m_blk * BM + static_cast<int>(q) * 32);
      }
      __syncwarp();                              // all lanes done with bias_w before the next tile overwrites it
    }
    if (lane == 0) tma_store_wait_all<0>();
  }

  __syncwarp();
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// Prep kernel: fp32 master adapter parameters -> bf16 padded side tiles + scaled singular values, for BOTH directions
// in one launch (the backward pass reuses the forward's tiles: A and B do not change in between).
//   forward  (contraction over K): a_fwd[RP, K] = A^T (A is [K, r]),  b_fwd[N, RP] = B^T (B is [r, N])
//   backward (contraction over N): a_bwd[RP, N] = B,                  b_bwd[K, RP] = A
//   s_rows[nS, RP] = scaling * s_eff[nS, r], zero padded          (RP = padded rank, 16 or 32, a kernel argument)
// Any output pointer may be null (skipped).
// ----------------------------------------------------------------------------------------------
__global__ void svlora_prep_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                   const float* __restrict__ s_eff, __nv_bfloat16* __restrict__ a_fwd,
                                   __nv_bfloat16* __restrict__ b_fwd, __nv_bfloat16* __restrict__ a_bwd,
                                   __nv_bfloat16* __restrict__ b_bwd, float* __restrict__ s_rows, int K, int N, int r,
                                   int RP, int nS, float scaling) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x;
  const int nthreads = gridDim.x * blockDim.x;
  if (a_fwd != nullptr) {
    for (int i = tid; i < RP * K; i += nthreads) {
      const int j = i / K, k = i - j * K;
      a_fwd[i] = __float2bfloat16(j < r ? A[static_cast<size_t>(k) * r + j] : 0.f);
    }
  }
  if (b_bwd != nullptr) {
    for (int i = tid; i < K * RP; i += nthreads) {
      const int k = i / RP, j = i - k * RP;
      b_bwd[i] = __float2bfloat16(j < r ? A[static_cast<size_t>(k) * r + j] : 0.f);
    }
  }
  if (b_fwd != nullptr) {
    for (int i = tid; i < N * RP; i += nthreads) {
      const int n = i / RP, j = i - n * RP;
      b_fwd[i] = __float2bfloat16(j < r ? B[static_cast<size_t>(j) * N + n] : 0.f);
    }
  }
  if (a_bwd != nullptr) {
    for (int i = tid; i < RP * N; i += nthreads) {
      const int j = i / N, n = i - j * N;
      a_bwd[i] = __float2bfloat16(j < r ? B[static_cast<size_t>(j) * N + n] : 0.f);
    }
  }
  for (int i = tid; i < nS * RP; i += nthreads) {
    const int b = i / RP, j = i - b * RP;
    s_rows[i] = (j < r) ? scaling * s_eff[static_cast<size_t>(b) * r + j] : 0.f;
  }
}

// ----------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ----------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                    CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                    CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(p);
  });
  return fn;
}

// cuTensorMapEncodeTiled is a DRIVER entry point and needs a context bound to the calling thread.  PyTorch's autograd worker
// threads only select a device; the primary context is bound lazily by the first runtime call that needs it, and a backward
// node whose first action is encoding a tensor map (ffm_frozen_linear, ffm_attention_bwd) would otherwise fail with
// CUDA_ERROR_INVALID_CONTEXT when nothing else ran on that thread before it.  Once per thread.
static void bind_context_to_thread() {
  static thread_local bool bound = false;
  if (!bound) {
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaSetDevice(dev);   // CUDA 12: initialises and binds the primary context;
    bound = true;                                                 // not a stream operation, legal during graph capture
  }
}

PFN_encodeTiled tensor_map_encode_fn() {          // shared with attention.cu
  bind_context_to_thread();
  return get_encode_fn();
}

int make_map_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                         uint32_t box_cols, CUtensorMapSwizzle swz, bool promote) {
  PFN_encodeTiled enc = tensor_map_encode_fn();
  if (enc == nullptr) {
    set_last_error("cuTensorMapEncodeTiled driver entry point not available");
    return FFM_ERR_CUDA;
  }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, swz,
                   promote ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B : CU_TENSOR_MAP_L2_PROMOTION_NONE,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_last_error("cuTensorMapEncodeTiled failed (%d) for [%llu x %llu] box [%u x %u]", (int)r,
                   (unsigned long long)rows, (unsigned long long)cols, box_rows, box_cols);
    return FFM_ERR_CUDA;
  }
  return FFM_OK;
}


// Which build runs: the CTA-pair kernel (cta_group::2 MMAs at the ideal issue rate, 0.69x the L2 -> shared-memory operand
// bytes per FLOP, 6-stage ring) for everything except the QuickGELU dual-store epilogue (c_fc forward), where both builds
// are bound by the epilogue (76.8 vs 77.9 us).  T = 12608: K=768 N=3072 plain 63.7 -> 56.3 us, dX of c_fc 62.4 -> 56.0 us,
// K=3072 N=768 59.4 -> 51.4 us.  (Until the cross-CTA arrives stopped compiling to MEMBAR.GPU — see mbar_arrive_cluster —
// the pair build lost on K = 768: 69.7 us.)  FFM_GEMM_PAIR=0 / 1 forces one build (A/B measurements).
static bool use_pair_kernel(int act) {
  static int v = -2;
  if (v == -2) {
    const char* e = getenv("FFM_GEMM_PAIR");
    v = (e == nullptr) ? -1 : (e[0] != '0');
  }
  if (v >= 0) return v != 0;
  return act != ACT_QUICKGELU;
}

// FFM_GEMM_DBG: bottleneck experiments only (1: no MMA, 2: no TMA loads, 4: no TMA stores, 64: print the pair build's epilogue
// phases when compiled with -DFFM_GEMM_PAIR_PROF, 128: no proxy fence, 256: no staging stores, 512: no staging-buffer reuse
// wait); results are garbage
int gemm_debug_mask() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("FFM_GEMM_DBG");
    v = (e == nullptr) ? 0 : atoi(e);
  }
  return v;
}

template <int R, bool ADAPT = true>
static int launch_single(const GemmOperands& o, const GemmParams& p0, cudaStream_t stream) {
  using C = GemmCfg<R>;
  CUtensorMap tm_x, tm_w, tm_a, tm_b, tm_y, tm_y2;
  int rc;
  if ((rc = make_map_bf16(&tm_x, o.x, o.T, o.K, BM, BK, CU_TENSOR_MAP_SWIZZLE_128B, true))) return rc;
  if ((rc = make_map_bf16(&tm_w, o.wmat, o.N, o.K, BN, BK, CU_TENSOR_MAP_SWIZZLE_128B, true))) return rc;
  if (ADAPT) {
    if ((rc = make_map_bf16(&tm_a, o.a_side, R, o.K, R, BK, CU_TENSOR_MAP_SWIZZLE_128B, true))) return rc;
    if ((rc = make_map_bf16(&tm_b, o.b_side, o.N, R, BN, R,
                            R == 16 ? CU_TENSOR_MAP_SWIZZLE_32B : CU_TENSOR_MAP_SWIZZLE_64B, false)))
      return rc;
  } else {
    tm_a = tm_w;      // never dereferenced by the ADAPT = false build
    tm_b = tm_w;
  }
  if ((rc = make_map_bf16(&tm_y, o.out, o.T, o.N, 32, EPI_PIECE_COLS, CU_TENSOR_MAP_SWIZZLE_64B, false))) return rc;
  const bool has_pre = p0.has_pre != 0;
  if ((rc = make_map_bf16(&tm_y2, has_pre ? o.out_pre : o.out, o.T, o.N, 32, EPI_PIECE_COLS, CU_TENSOR_MAP_SWIZZLE_64B,
                          false)))
    return rc;
  GemmParams p = p0;
  p.m_tiles = (o.T + BM - 1) / BM;
  p.n_tiles = (o.N + BN - 1) / BN;
  p.k_blocks = (o.K + BK - 1) / BK;

  // dynamic-smem opt-in of the three epilogue variants; the attribute is per device
  {
    static thread_local int attr_dev = -1;
    int dev = 0;
    FFM_CHECK_CUDA(cudaGetDevice(&dev));
    if (dev != attr_dev) {
      FFM_CHECK_CUDA(cudaFuncSetAttribute(svlora_gemm_kernel<ACT_NONE, R, ADAPT>,
                                          cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
      if (ADAPT) {
        FFM_CHECK_CUDA(cudaFuncSetAttribute(svlora_gemm_kernel<ACT_QUICKGELU, R, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
        FFM_CHECK_CUDA(cudaFuncSetAttribute(svlora_gemm_kernel<ACT_QUICKGELU_GRAD, R, true>,
                                            cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_BYTES));
      }
      attr_dev = dev;
    }
  }
  const int tiles = p.m_tiles * p.n_tiles;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  GemmProfileScope prof;
  if ((rc = gemm_profile_begin(&prof, stream))) return rc;
  if constexpr (!ADAPT) {
    svlora_gemm_kernel<ACT_NONE, R, false><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tm_x, tm_w, tm_a, tm_b, tm_y,
                                                                                         tm_y2, p);
    FFM_CHECK_CUDA(cudaGetLastError());
    count_launch();
    return gemm_profile_end(&prof, o.T, o.K, -o.N, stream);     // N < 0 in the record: launch without the adapter terms
  }
  switch (p.act) {
    case ACT_QUICKGELU:
      svlora_gemm_kernel<ACT_QUICKGELU, R><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tm_x, tm_w, tm_a, tm_b, tm_y,
                                                                                         tm_y2, p);
      break;
    case ACT_QUICKGELU_GRAD:
      svlora_gemm_kernel<ACT_QUICKGELU_GRAD, R><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tm_x, tm_w, tm_a, tm_b,
                                                                                              tm_y, tm_y2, p);
      break;
    default:
      svlora_gemm_kernel<ACT_NONE, R><<<grid, NUM_THREADS, C::SMEM_BYTES, stream>>>(tm_x, tm_w, tm_a, tm_b, tm_y, tm_y2,
                                                                                    p);
  }
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return gemm_profile_end(&prof, o.T, o.K, o.N, stream);
}

static int launch_svlora_gemm(const GemmOperands& o, cudaStream_t stream) {
  FFM_CHECK_ARG(o.T > 0 && o.K > 0 && o.N > 0, "svlora gemm: empty problem T=%d K=%d N=%d", o.T, o.K, o.N);
  FFM_CHECK_ARG(o.K % 8 == 0 && o.N % 8 == 0, "svlora gemm: K (%d) and N (%d) must be multiples of 8 (TMA 16-B strides)",
                o.K, o.N);
  FFM_CHECK_ARG(o.b_prime > 0 && o.num_slices > 0 && o.row_div > 0, "svlora gemm: b_prime/num_slices/row_div must be positive");
  FFM_CHECK_ARG(o.act != ACT_QUICKGELU_GRAD || o.aux != nullptr, "svlora gemm: ACT_QUICKGELU_GRAD needs aux");
  FFM_CHECK_ARG(o.rp == RP || o.rp == RP_MAX, "svlora gemm: padded rank must be %d or %d", RP, RP_MAX);
  const uintptr_t align_or = reinterpret_cast<uintptr_t>(o.x) | reinterpret_cast<uintptr_t>(o.wmat) |
                             reinterpret_cast<uintptr_t>(o.a_side) | reinterpret_cast<uintptr_t>(o.b_side) |
                             reinterpret_cast<uintptr_t>(o.out) | reinterpret_cast<uintptr_t>(o.out_pre) |
                             reinterpret_cast<uintptr_t>(o.s_rows) | reinterpret_cast<uintptr_t>(o.h_out) |
                             reinterpret_cast<uintptr_t>(o.z_out) | reinterpret_cast<uintptr_t>(o.aux);
  FFM_CHECK_ARG((align_or & 15u) == 0, "svlora gemm: all device pointers must be 16-byte aligned");
  if (o.rp == RP && use_pair_kernel(o.act)) return launch_svlora_gemm_pair(o, stream);

  GemmParams p;
  p.bias = o.bias;
  p.s_rows = o.s_rows;
  p.h_out = o.h_out;
  p.z_out = reinterpret_cast<__nv_bfloat16*>(o.z_out);
  p.aux = reinterpret_cast<const __nv_bfloat16*>(o.aux);
  p.T = o.T; p.K = o.K; p.N = o.N;
  p.rp = o.rp;
  p.b_prime = o.b_prime; p.num_slices = o.num_slices; p.row_div = o.row_div;
  p.act = o.act;
  p.has_pre = (o.act == ACT_QUICKGELU && o.out_pre != nullptr) ? 1 : 0;
  p.m_tiles = p.n_tiles = p.k_blocks = 0;
  p.dbg = gemm_debug_mask();
  return o.rp == RP ? launch_single<RP>(o, p, stream) : launch_single<RP_MAX>(o, p, stream);
}

// Adapter tiles prepared by svlora_prep_kernel (all offsets 256-B aligned).  The forward workspace holds the tiles of
// both directions so that the backward pass can skip its own prep launch.
struct SvloraTiles {
  __nv_bfloat16* a_fwd;   // [RP, K]
  __nv_bfloat16* b_fwd;   // [N, RP]
  float* s_rows;          // [nS, RP]
  __nv_bfloat16* a_bwd;   // [RP, N]
  __nv_bfloat16* b_bwd;   // [K, RP]
};

static size_t align256(size_t v) { return (v + 255) & ~size_t(255); }

// sized for the largest padded rank so the workspace queries need no rank argument
static size_t svlora_tiles_bytes(int K, int N, int nS) {
  return 2 * (align256(static_cast<size_t>(RP_MAX) * K * 2) + align256(static_cast<size_t>(N) * RP_MAX * 2)) +
         align256(static_cast<size_t>(nS) * RP_MAX * 4);
}

static void carve_tiles(SvloraTiles* w, void* ws, int K, int N, int nS) {
  uint8_t* p = static_cast<uint8_t*>(ws);
  w->a_fwd = reinterpret_cast<__nv_bfloat16*>(p);
  p += align256(static_cast<size_t>(RP_MAX) * K * 2);
  w->b_fwd = reinterpret_cast<__nv_bfloat16*>(p);
  p += align256(static_cast<size_t>(N) * RP_MAX * 2);
  w->s_rows = reinterpret_cast<float*>(p);
  p += align256(static_cast<size_t>(nS) * RP_MAX * 4);
  w->a_bwd = reinterpret_cast<__nv_bfloat16*>(p);
  p += align256(static_cast<size_t>(RP_MAX) * N * 2);
  w->b_bwd = reinterpret_cast<__nv_bfloat16*>(p);
}

static int padded_rank(int r) { return r <= RP ? RP : RP_MAX; }

// implemented in svlora_small.cu
int launch_svlora_bwd_small(const __nv_bfloat16* x, const __nv_bfloat16* dy, const float* h, const float* dzu,
                            const __nv_bfloat16* z, const __nv_bfloat16* dh, float* dA, float* dB, float* ds_eff,
                            void* scratch, size_t scratch_bytes, int T, int K, int N, int r, int rp, int nS, int b_prime,
                            int num_slices, int row_div, float scaling, cudaStream_t stream);
size_t svlora_bwd_small_scratch_bytes(int T, int K, int N);

}  // namespace ffm

using namespace ffm;

extern "C" {

const char* ffm_last_error(void) { return g_last_error; }

int ffm_version(void) { return 200; }   // round 2: tcgen05 attention, frozen linear, OCT / RN50 entry points

int ffm_svlora_max_rank(void) { return RP_MAX; }

int ffm_svlora_padded_rank(int r) { return (r >= 1 && r <= RP_MAX) ? padded_rank(r) : 0; }

long long ffm_launch_count(int reset) {
  return reset ? g_launches.exchange(0) : g_launches.load();
}

int ffm_profile_enable(int enable) {
  std::lock_guard<std::mutex> lk(g_prof_mutex);
  g_prof_enabled = enable != 0;
  return FFM_OK;
}

int ffm_profile_read(float* ms_host, int* tkn_host, int max_records) {
  std::vector<GemmRecord> recs;
  {
    std::lock_guard<std::mutex> lk(g_prof_mutex);
    recs.swap(g_prof_records);
  }
  int n = 0;
  for (auto& r : recs) {
    float ms = 0.f;
    cudaError_t e = cudaEventSynchronize(r.stop);
    if (e == cudaSuccess) e = cudaEventElapsedTime(&ms, r.start, r.stop);
    cudaEventDestroy(r.start);
    cudaEventDestroy(r.stop);
    if (e != cudaSuccess) {
      set_last_error("ffm_profile_read: %s", cudaGetErrorString(e));
      return FFM_ERR_CUDA;
    }
    if (n < max_records && ms_host && tkn_host) {
      ms_host[n] = ms;
      tkn_host[3 * n] = r.T; tkn_host[3 * n + 1] = r.K; tkn_host[3 * n + 2] = r.N;
      ++n;
    }
  }
  return n;
}

size_t ffm_svlora_fwd_workspace_bytes(int T, int K, int N, int n_samples) {
  (void)T;
  return svlora_tiles_bytes(K, N, n_samples);
}

size_t ffm_svlora_bwd_workspace_bytes(int T, int K, int N, int n_samples) {
  // own tiles (only used when the forward workspace is not handed over) + dzu f32 [T,16] + dh bf16 [T,16] + partials
  return svlora_tiles_bytes(K, N, n_samples) + align256(static_cast<size_t>(T) * RP_MAX * 4) +
         align256(static_cast<size_t>(T) * RP_MAX * 2) + svlora_bwd_small_scratch_bytes(T, K, N);
}

int ffm_svlora_prepare(const float* lora_a, const float* lora_b, const float* s_eff, void* workspace,
                       size_t workspace_bytes, int K, int N, int r, int n_samples, float scaling,
                       cudaStream_t stream) {
  FFM_CHECK_ARG(lora_a && lora_b && s_eff && workspace, "ffm_svlora_prepare: null pointer argument");
  FFM_CHECK_ARG(r >= 1 && r <= RP_MAX, "ffm_svlora_prepare: rank %d not in [1, %d]", r, RP_MAX);
  FFM_CHECK_ARG(n_samples >= 1 && K >= 1 && N >= 1, "ffm_svlora_prepare: bad sizes");
  FFM_CHECK_ARG(workspace_bytes >= svlora_tiles_bytes(K, N, n_samples), "ffm_svlora_prepare: workspace too small");
  SvloraTiles ws;
  carve_tiles(&ws, workspace, K, N, n_samples);
  svlora_prep_kernel<<<96, 256, 0, stream>>>(lora_a, lora_b, s_eff, ws.a_fwd, ws.b_fwd, ws.a_bwd, ws.b_bwd, ws.s_rows,
                                             K, N, r, padded_rank(r), n_samples, scaling);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch();
  return FFM_OK;
}

int ffm_svlora_fwd(const void* x, const void* w, const float* bias, const float* lora_a, const float* lora_b,
                   const float* s_eff, void* y, void* y_dact, float* h_out, void* z_out, void* workspace,
                   size_t workspace_bytes, int T, int K, int N, int r, int n_samples, int b_prime, int num_slices,
                   int row_div, float scaling, int act, cudaStream_t stream) {
  FFM_CHECK_ARG(x && w && y && workspace, "ffm_svlora_fwd: null pointer argument");
  // lora_a == lora_b == NULL: the workspace already holds the tiles (ffm_svlora_prepare, same K / N / n_samples)
  const bool prepared = lora_a == nullptr && lora_b == nullptr;
  FFM_CHECK_ARG(prepared || (lora_a && lora_b && s_eff), "ffm_svlora_fwd: lora_a, lora_b and s_eff go together");
  FFM_CHECK_ARG(row_div >= 1, "ffm_svlora_fwd: row_div must be >= 1");
  FFM_CHECK_ARG(r >= 1 && r <= RP_MAX, "ffm_svlora_fwd: rank %d not in [1, %d]", r, RP_MAX);
  FFM_CHECK_ARG(n_samples >= 1, "ffm_svlora_fwd: n_samples must be >= 1");
  FFM_CHECK_ARG(b_prime >= 1 && num_slices >= 1 && (b_prime - 1) / num_slices < n_samples,
                "ffm_svlora_fwd: sample mapping (b_prime=%d, num_slices=%d) exceeds n_samples=%d", b_prime,
                num_slices, n_samples);
  FFM_CHECK_ARG(act == ACT_NONE || act == ACT_QUICKGELU, "ffm_svlora_fwd: act must be 0 or 1");
  FFM_CHECK_ARG(workspace_bytes >= ffm_svlora_fwd_workspace_bytes(T, K, N, n_samples),
                "ffm_svlora_fwd: workspace too small");
  SvloraTiles ws;
  carve_tiles(&ws, workspace, K, N, n_samples);
  const int rp = padded_rank(r);
  if (!prepared) {
    svlora_prep_kernel<<<96, 256, 0, stream>>>(lora_a, lora_b, s_eff, ws.a_fwd, ws.b_fwd, ws.a_bwd, ws.b_bwd,
                                               ws.s_rows, K, N, r, rp, n_samples, scaling);
    FFM_CHECK_CUDA(cudaGetLastError());
    count_launch();
  }
  GemmOperands o;
  o.x = x; o.wmat = w; o.a_side = ws.a_fwd; o.b_side = ws.b_fwd; o.s_rows = ws.s_rows; o.bias = bias;
  o.out = y; o.out_pre = y_dact; o.h_out = h_out; o.z_out = z_out; o.aux = nullptr;
  o.T = T; o.K = K; o.N = N; o.b_prime = b_prime; o.num_slices = num_slices; o.row_div = row_div; o.act = act;
  o.rp = rp;
  return launch_svlora_gemm(o, stream);
}

int ffm_frozen_linear(const void* x, const void* w, const float* bias, void* y, int T, int K, int N, cudaStream_t stream) {
  FFM_CHECK_ARG(x && w && y, "ffm_frozen_linear: null pointer argument");
  FFM_CHECK_ARG(T > 0 && K > 0 && N > 0 && K % 8 == 0 && N % 8 == 0,
                "ffm_frozen_linear: T, K, N must be positive and K (%d), N (%d) multiples of 8", K, N);
  const uintptr_t align_or = reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(w) | reinterpret_cast<uintptr_t>(y);
  FFM_CHECK_ARG((align_or & 15u) == 0, "ffm_frozen_linear: device pointers must be 16-byte aligned");
  GemmOperands o;
  o.x = x; o.wmat = w; o.a_side = nullptr; o.b_side = nullptr; o.s_rows = nullptr; o.bias = bias;
  o.out = y; o.out_pre = nullptr; o.h_out = nullptr; o.z_out = nullptr; o.aux = nullptr;
  o.T = T; o.K = K; o.N = N; o.b_prime = 1; o.num_slices = 1; o.row_div = 1; o.act = ACT_NONE; o.rp = RP;
  GemmParams p;
  p.bias = bias; p.s_rows = nullptr; p.h_out = nullptr; p.z_out = nullptr; p.aux = nullptr;
  p.T = T; p.K = K; p.N = N; p.rp = RP; p.b_prime = 1; p.num_slices = 1; p.row_div = 1; p.act = ACT_NONE; p.has_pre = 0;
  p.m_tiles = p.n_tiles = p.k_blocks = 0;
  p.dbg = gemm_debug_mask();
  // FFM_FROZEN_PAIR=0 keeps the single-CTA adapter-free build (A/B measurements)
  static const bool pair_ok = [] { const char* e = getenv("FFM_FROZEN_PAIR"); return e == nullptr || e[0] != '0'; }();
  if (pair_ok) return launch_svlora_gemm_pair(o, stream);
  return launch_single<RP, false>(o, p, stream);
}

int ffm_svlora_bwd(const void* dy, const void* x, const void* w_t, const float* lora_a, const float* lora_b,
                   const float* s_eff, const float* h, const void* z, const void* fwd_workspace,
                   const void* gelu_dact, void* dx, float* d_lora_a, float* d_lora_b, float* d_s_eff, void* workspace,
                   size_t workspace_bytes, int T, int K, int N, int r, int n_samples, int b_prime, int num_slices,
                   int row_div, float scaling, cudaStream_t stream) {
  return ffm_svlora_bwd_phase(dy, x, w_t, lora_a, lora_b, s_eff, h, z, fwd_workspace, gelu_dact, dx, d_lora_a, d_lora_b,
                              d_s_eff, workspace, workspace_bytes, T, K, N, r, n_samples, b_prime, num_slices, row_div,
                              scaling, FFM_BWD_DX | FFM_BWD_PARAMS, stream);
}

int ffm_svlora_bwd_phase(const void* dy, const void* x, const void* w_t, const float* lora_a, const float* lora_b,
                         const float* s_eff, const float* h, const void* z, const void* fwd_workspace,
                         const void* gelu_dact, void* dx, float* d_lora_a, float* d_lora_b, float* d_s_eff,
                         void* workspace, size_t workspace_bytes, int T, int K, int N, int r, int n_samples,
                         int b_prime, int num_slices, int row_div, float scaling, int phases, cudaStream_t stream) {
  FFM_CHECK_ARG(dy && x && w_t && lora_a && lora_b && s_eff && h && z && dx && d_lora_a && d_lora_b && d_s_eff &&
                    workspace,
                "ffm_svlora_bwd: null pointer argument");
  FFM_CHECK_ARG((phases & ~(FFM_BWD_DX | FFM_BWD_PARAMS)) == 0 && phases != 0, "ffm_svlora_bwd_phase: bad phase mask");
  FFM_CHECK_ARG(r >= 1 && r <= RP_MAX, "ffm_svlora_bwd: rank %d not in [1, %d]", r, RP_MAX);
  FFM_CHECK_ARG(b_prime >= 1 && num_slices >= 1 && row_div >= 1 && (b_prime - 1) / num_slices < n_samples,
                "ffm_svlora_bwd: sample mapping exceeds n_samples");
  FFM_CHECK_ARG(workspace_bytes >= ffm_svlora_bwd_workspace_bytes(T, K, N, n_samples),
                "ffm_svlora_bwd: workspace too small");
  SvloraTiles ws;
  const int rp = padded_rank(r);
  uint8_t* rest = static_cast<uint8_t*>(workspace) + svlora_tiles_bytes(K, N, n_samples);
  if (fwd_workspace != nullptr) {
    // tiles prepared by ffm_svlora_fwd of the same step (A, B, s_eff unchanged since): no prep launch
    carve_tiles(&ws, const_cast<void*>(fwd_workspace), K, N, n_samples);
  } else {
    carve_tiles(&ws, workspace, K, N, n_samples);
    if (phases & FFM_BWD_DX) {
      svlora_prep_kernel<<<96, 256, 0, stream>>>(lora_a, lora_b, s_eff, nullptr, nullptr, ws.a_bwd, ws.b_bwd,
                                                 ws.s_rows, K, N, r, rp, n_samples, scaling);
      FFM_CHECK_CUDA(cudaGetLastError());
      count_launch();
    }
  }
  float* dzu = reinterpret_cast<float*>(rest);
  rest += align256(static_cast<size_t>(T) * RP_MAX * 4);
  __nv_bfloat16* dh = reinterpret_cast<__nv_bfloat16*>(rest);
  rest += align256(static_cast<size_t>(T) * RP_MAX * 2);
  const size_t scratch_bytes = workspace_bytes - static_cast<size_t>(rest - static_cast<uint8_t*>(workspace));
  // backward: contraction over N.  Aside = B ([r, N]), Bside = A ([K, r]); side outputs dzu = dy·B^T (f32) and
  // dh = bf16(dzu ⊙ s_rows), the operand of dA.
  GemmOperands o;
  o.x = dy; o.wmat = w_t; o.a_side = ws.a_bwd; o.b_side = ws.b_bwd; o.s_rows = ws.s_rows; o.bias = nullptr;
  o.out = dx; o.out_pre = nullptr; o.h_out = dzu; o.z_out = dh; o.aux = gelu_dact;
  o.T = T; o.K = N; o.N = K; o.b_prime = b_prime; o.num_slices = num_slices; o.row_div = row_div;
  o.rp = rp;
  o.act = gelu_dact != nullptr ? ACT_QUICKGELU_GRAD : ACT_NONE;
  if (phases & FFM_BWD_DX) {
    int rc = launch_svlora_gemm(o, stream);
    if (rc != FFM_OK) return rc;
  }
  if (!(phases & FFM_BWD_PARAMS)) return FFM_OK;
  return launch_svlora_bwd_small(reinterpret_cast<const __nv_bfloat16*>(x), reinterpret_cast<const __nv_bfloat16*>(dy),
                                 h, dzu, reinterpret_cast<const __nv_bfloat16*>(z), dh, d_lora_a, d_lora_b, d_s_eff,
                                 rest, scratch_bytes, T, K, N, r, rp, n_samples, b_prime, num_slices, row_div, scaling,
                                 stream);
}

}  // extern "C"
