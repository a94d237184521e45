// Internal declarations shared by the single-CTA and CTA-pair builds of the fused SVLoRA GEMM.
#pragma once

#include "ffm_common.cuh"

namespace ffm {

constexpr int RP = 16;              // padded adapter rank of the ViT recipes (r <= 16); the pair build is RP-only
constexpr int RP_MAX = 32;          // largest padded rank (RN50 recipe r = 32): single-CTA build only
enum : int { ACT_NONE = 0, ACT_QUICKGELU = 1, ACT_QUICKGELU_GRAD = 2 };

struct GemmParams {
  const float* bias;            // [N] or nullptr
  const float* s_rows;          // [n_samples, RP] fp32 (already multiplied by alpha/r)
  float* h_out;                 // [T, RP] fp32 or nullptr
  __nv_bfloat16* z_out;         // [T, RP] bf16 or nullptr: Z = bf16(H ⊙ s_rows), the operand the adapter gradients need
  const __nv_bfloat16* aux;     // ACT_QUICKGELU_GRAD: QuickGELU'(u) saved by the forward [T, N]
  int T, K, N;
  int rp;                       // padded adapter rank (16 or 32): row stride of s_rows / h_out / z_out
  int b_prime, num_slices;      // sample(t) = ((t / row_div) % b_prime) / num_slices
  int row_div;                  // 1: sequence-first rows [L, B', C] (reference); L: batch-first rows [B', L, C]
  int act;
  int has_pre;                  // ACT_QUICKGELU: also store QuickGELU'(u) through tm_y2
  int m_tiles, n_tiles, k_blocks;
  int dbg;                      // FFM_GEMM_DBG experiment mask (see gemm_debug_mask); 0 in production
};

int gemm_debug_mask();

struct GemmOperands {
  const void* x;        // [T, K] bf16
  const void* wmat;     // [N, K] bf16
  const void* a_side;   // [rp, K] bf16
  const void* b_side;   // [N, rp] bf16
  const float* s_rows;  // [nS, rp]
  const float* bias;    // [N] or null
  void* out;            // [T, N] bf16
  void* out_pre;        // [T, N] bf16 or null (ACT_QUICKGELU only): receives QuickGELU'(u)
  float* h_out;         // [T, RP] or null
  void* z_out;          // [T, RP] bf16 or null
  const void* aux;      // [T, N] bf16 (ACT_QUICKGELU_GRAD)
  int T, K, N, b_prime, num_slices, row_div, act;
  int rp;               // padded adapter rank: 16 or 32
};

// row-major bf16 matrix [rows, cols] (cols contiguous) -> 2-D tiled map with box [box_rows, box_cols]
int make_map_bf16(CUtensorMap* out, const void* ptr, uint64_t rows, uint64_t cols, uint32_t box_rows,
                  uint32_t box_cols, CUtensorMapSwizzle swz, bool promote);

// event profiling hooks (ffm_profile_*): call around the kernel launch
struct GemmProfileScope {
  cudaEvent_t start = nullptr, stop = nullptr;
  bool active = false;
};
int gemm_profile_begin(GemmProfileScope* sc, cudaStream_t stream);
int gemm_profile_end(GemmProfileScope* sc, int T, int K, int N, cudaStream_t stream);

// CTA-pair (cta_group::2) build, svlora_gemm_pair.cu
int launch_svlora_gemm_pair(const GemmOperands& o, cudaStream_t stream);

#ifdef __CUDACC__
// sigmoid(1.702 u) = 0.5 + 0.5 tanh(0.851 u): ONE MUFU op (tanh.approx, rel. error ~2^-11 << bf16 output rounding)
__device__ __forceinline__ float tanh_approx(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float quick_gelu(float u) {
  const float hu = 0.5f * u;
  return fmaf(hu, tanh_approx(0.851f * u), hu);
}
__device__ __forceinline__ float quick_gelu_grad(float u) {
  const float s = fmaf(0.5f, tanh_approx(0.851f * u), 0.5f);
  return s * fmaf(1.702f * u, 1.0f - s, 1.0f);
}

// ----------------------------------------------------------------------------------------------
// Epilogue tail shared by the single-CTA and CTA-pair builds.  Every epilogue warp is an independent stream: it owns
// 32 tile rows (lane = row), converts one 32-column piece at a time and ships it with its OWN TMA store from its OWN
// double-buffered 2 KB staging area (SW64 rows of 64 B) — no CTA-wide barriers anywhere in the epilogue.
// ----------------------------------------------------------------------------------------------
constexpr int EPI_PIECE_COLS = 32;
constexpr int EPI_PIECE_BYTES = 32 * EPI_PIECE_COLS * 2;   // 32 rows x 64 B

// f[j] (+bias already added) -> optional x saved QuickGELU' (backward) -> optional QuickGELU (+ its derivative as a
// second store) -> bf16 -> staging -> TMA store at (row0, gcol).  `unit` counts this warp's stores (buffer = unit&1).
// Saved QuickGELU'(u) values of one piece (ACT_QUICKGELU_GRAD): 64 contiguous bytes of this lane's row.  Issued one
// piece ahead by the callers so the global-load latency overlaps the TMEM read / conversion of the current piece.
struct EpiAux {
  uint4 v[4];
  bool ready;     // false: ragged / out-of-range piece, epi_store_piece falls back to guarded scalar loads
};
template <int ACT>
__device__ __forceinline__ void epi_load_aux(const GemmParams& p, EpiAux& a, int grow, int gcol) {
  a.ready = false;
  if (ACT != ACT_QUICKGELU_GRAD) return;
  if (grow < p.T && gcol + 32 <= p.N) {
    const uint4* up = reinterpret_cast<const uint4*>(p.aux + static_cast<size_t>(grow) * p.N + gcol);
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) a.v[j8] = __ldg(up + j8);
    a.ready = true;
  }
}

template <int ACT>
__device__ __forceinline__ void epi_store_piece(const GemmParams& p, float (&f)[32], const EpiAux& aux,
                                                const CUtensorMap* tm_y, const CUtensorMap* tm_y2, uint8_t* stage_w,
                                                uint32_t& unit, uint32_t lane, int grow, int gcol, int row0) {
  if (ACT == ACT_QUICKGELU_GRAD && grow < p.T) {
    if (aux.ready) {
#pragma unroll
      for (int j8 = 0; j8 < 4; ++j8) {
        const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&aux.v[j8]);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float2 uu = __bfloat1622float2(h2[e]);
          f[j8 * 8 + 2 * e] *= uu.x;
          f[j8 * 8 + 2 * e + 1] *= uu.y;
        }
      }
    } else {
      const __nv_bfloat16* up = p.aux + static_cast<size_t>(grow) * p.N + gcol;
      for (int j = 0; j < 32; ++j)
        if (gcol + j < p.N) f[j] *= __bfloat162float(up[j]);
    }
  }
  const int n_pass = (ACT == ACT_QUICKGELU && p.has_pre) ? 2 : 1;
  float g2[32];
#pragma unroll 1
  for (int pass = 0; pass < n_pass; ++pass) {
    uint8_t* ob = stage_w + (unit & 1u) * EPI_PIECE_BYTES;
    // the TMA store that last read this buffer (2 units ago) must have finished reading smem
    if (lane == 0 && !(p.dbg & 512)) tma_store_wait_read<1>();
    __syncwarp();
    if (ACT == ACT_QUICKGELU) {
      if (n_pass == 2 && pass == 0) {
        // first store of the dual store: QuickGELU'(u), all the backward pass needs (u itself is not kept)
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          const float sgm = fmaf(0.5f, tanh_approx(0.851f * f[j]), 0.5f);
          g2[j] = f[j] * sgm;                                       // QuickGELU(u), stored by the next pass
          f[j] = sgm * fmaf(1.702f * f[j], 1.0f - sgm, 1.0f);      // QuickGELU'(u)
        }
      } else if (n_pass == 2) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = g2[j];
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = quick_gelu(f[j]);
      }
    }
    // SW64: 16-B chunk k of row r lives at chunk (k ^ ((r >> 1) & 3)) of the 64-B row
    uint8_t* orow = ob + lane * 64u;
    const uint32_t sw = (lane >> 1) & 3u;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
      uint4 pk;
      pk.x = pack_bf16x2(f[j8 * 8 + 0], f[j8 * 8 + 1]);
      pk.y = pack_bf16x2(f[j8 * 8 + 2], f[j8 * 8 + 3]);
      pk.z = pack_bf16x2(f[j8 * 8 + 4], f[j8 * 8 + 5]);
      pk.w = pack_bf16x2(f[j8 * 8 + 6], f[j8 * 8 + 7]);
      if (!(p.dbg & 256)) *reinterpret_cast<uint4*>(orow + ((static_cast<uint32_t>(j8) ^ sw) << 4)) = pk;
    }
    if (!(p.dbg & 128)) fence_proxy_async_smem();
    __syncwarp();
    if (lane == 0) {
      const CUtensorMap* tm = (n_pass == 2 && pass == 0) ? tm_y2 : tm_y;
      if (!(p.dbg & 4)) tma_store_2d(tm, ob, gcol, row0);
      tma_store_commit();
    }
    ++unit;
  }
}
#endif

}  // namespace ffm
