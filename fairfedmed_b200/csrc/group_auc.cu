// Group-wise rank statistics for the fairness metrics (AUC, ES-AUC, ES-ACC, DPD/SPD, EOD, AOD).
//
// Reference: evaluation/metrics.py:340-356 (compute_auc -> sklearn roc_auc_score, macro one-vs-rest),
// :513-547 (equity_scaled_AUC), :486-511 (equity_scaled_accuracy), :248-292 (DPD / EOD / AOD inputs).
// sklearn's AUC equals the tie-aware Mann–Whitney statistic (gt + eq/2)/(P*Nn) on the stored float32
// values, so the device side only produces INTEGER counts (bit-exact by construction); the host mirror turns
// them into the reference's floating-point scores.
//
// Pipeline (one segmented sort for every (attribute, group, probability column) at once):
//   expand   : each sample is replicated per (attribute slot a', column c) into a 64-bit key
//                [ a' : 8 | group+1 : 8 | c : 1 | order-preserving f32 score : 32 | is_positive : 1 ]
//              (a' = 0 is the "overall" slot; is_positive = (label == c)),
//   sort     : LSD radix sort, 8-bit digits, 7 passes (stable; per-block digit histograms -> one scan block per digit
//              -> ranked scatter with warp match_any, digit bases scanned in shared memory), keys only,
//   prefix   : exclusive prefix count of positives over the sorted keys,
//   rank     : every positive finds the start of its tie run by galloping back from its own position (segment starts
//              come from a small table); negatives strictly below and tied negatives accumulate into 64-bit counters
//              per slot,
//   confusion: tp / fp / tn / fn of pred = argmax(prob) per slot.
#include <algorithm>

#include "../../include/ffm_b200.h"
#include "ffm_common.cuh"

namespace ffm {

constexpr int RS_THREADS = 256;
constexpr int RS_ROWS = 32;                 // rows of 32 keys per block
constexpr int RS_CHUNK = RS_ROWS * 32;      // 1024 keys per block
constexpr int RS_PASSES = 7;                // 50 key bits, 8-bit digits
constexpr unsigned long long KEY_PAD = ~0ull;

__device__ __forceinline__ uint32_t order_f32(float f) {
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

__global__ void expand_keys_kernel(const float* __restrict__ prob, const int32_t* __restrict__ label,
                                   const int32_t* __restrict__ attrs, unsigned long long* __restrict__ keys, int N,
                                   int n_attr, long long E, long long E_pad) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < E_pad; i += stride) {
    if (i >= E) { keys[i] = KEY_PAD; continue; }     // padding sorts to the very end
    const int c = static_cast<int>(i & 1);
    const long long t = i >> 1;
    const int n = static_cast<int>(t % N);
    const int a = static_cast<int>(t / N);            // 0 = overall, 1..n_attr
    int g1 = 0;
    if (a > 0) {
      g1 = attrs[static_cast<size_t>(a - 1) * N + n] + 1;
      g1 = g1 < 0 ? 0 : (g1 > 255 ? 255 : g1);
    }
    const unsigned long long pos = (label[n] == c) ? 1ull : 0ull;
    const unsigned long long sc = order_f32(prob[static_cast<size_t>(n) * 2 + c]);
    keys[i] = (static_cast<unsigned long long>(a) << 42) | (static_cast<unsigned long long>(g1) << 34) |
              (static_cast<unsigned long long>(c) << 33) | (sc << 1) | pos;
  }
}

// ---- radix sort pass: histogram -> scan -> scatter -------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS)
radix_hist_kernel(const unsigned long long* __restrict__ keys, uint32_t* __restrict__ counts, int shift,
                  int nblocks) {
  __shared__ uint32_t h[256];
  h[threadIdx.x] = 0;
  __syncthreads();
  const size_t base = static_cast<size_t>(blockIdx.x) * RS_CHUNK;
#pragma unroll
  for (int k = 0; k < RS_CHUNK / RS_THREADS; ++k) {
    const unsigned long long key = keys[base + k * RS_THREADS + threadIdx.x];
    atomicAdd(&h[(key >> shift) & 0xFF], 1u);
  }
  __syncthreads();
  counts[static_cast<size_t>(threadIdx.x) * nblocks + blockIdx.x] = h[threadIdx.x];
}

// single-block exclusive scan of `n` uint32 values (digit-major histogram table or block sums)
__global__ void __launch_bounds__(1024)
exclusive_scan_kernel(uint32_t* __restrict__ data, int n) {
  __shared__ uint32_t part[1024];
  const int per = (n + 1023) / 1024;
  const int lo = threadIdx.x * per, hi = min(n, lo + per);
  uint32_t s = 0;
  for (int i = lo; i < hi; ++i) s += data[i];
  part[threadIdx.x] = s;
  __syncthreads();
  // Hillis–Steele inclusive scan over 1024 partials
  for (int off = 1; off < 1024; off <<= 1) {
    uint32_t v = 0;
    if (threadIdx.x >= off) v = part[threadIdx.x - off];
    __syncthreads();
    part[threadIdx.x] += v;
    __syncthreads();
  }
  uint32_t run = part[threadIdx.x] - s;   // exclusive prefix of this thread's segment
  for (int i = lo; i < hi; ++i) {
    const uint32_t v = data[i];
    data[i] = run;
    run += v;
  }
}

// block-wide exclusive scan of one value per thread (RS_THREADS threads); returns the exclusive prefix, *total = block sum
__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* warp_tot, uint32_t* total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  __syncthreads();                    // warp_tot may still be read by the previous call
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t woff = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < RS_THREADS / 32; ++w) {
    const uint32_t t = warp_tot[w];
    if (w < warp) woff += t;
    tot += t;
  }
  *total = tot;
  return woff + inc - v;
}

// one block per digit: exclusive scan of the digit's per-block counts in place (coalesced chunks of RS_THREADS), the
// digit's total to totals[digit].  256 blocks instead of one: the table scan no longer serialises a pass.
__global__ void __launch_bounds__(RS_THREADS)
digit_row_scan_kernel(uint32_t* __restrict__ counts, uint32_t* __restrict__ totals, int nblocks) {
  __shared__ uint32_t warp_tot[RS_THREADS / 32];
  uint32_t* row = counts + static_cast<size_t>(blockIdx.x) * nblocks;
  uint32_t carry = 0;
  for (int base = 0; base < nblocks; base += RS_THREADS) {
    const int i = base + threadIdx.x;
    const uint32_t v = i < nblocks ? row[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(v, warp_tot, &tot);
    if (i < nblocks) row[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) totals[blockIdx.x] = carry;
}

__global__ void __launch_bounds__(RS_THREADS)
radix_scatter_kernel(const unsigned long long* __restrict__ keys_in, unsigned long long* __restrict__ keys_out,
                     const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ totals, int shift, int nblocks) {
  __shared__ uint16_t row_cnt[RS_ROWS][256];    // keys of digit d in row r, then exclusive prefix over rows
  __shared__ uint32_t digit_base[256];          // keys with a smaller digit, over the whole array
  __shared__ uint32_t warp_tot[RS_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < RS_ROWS * 256; i += RS_THREADS) (&row_cnt[0][0])[i] = 0;
  {
    uint32_t tot;
    digit_base[threadIdx.x] = block_exclusive_scan(totals[threadIdx.x], warp_tot, &tot);
  }
  __syncthreads();
  const size_t base = static_cast<size_t>(blockIdx.x) * RS_CHUNK;
  constexpr int ROWS_PER_WARP = RS_ROWS / (RS_THREADS / 32);
  unsigned long long key[ROWS_PER_WARP];
  uint32_t rank_in_row[ROWS_PER_WARP];
#pragma unroll
  for (int k = 0; k < ROWS_PER_WARP; ++k) {
    const int row = warp * ROWS_PER_WARP + k;
    key[k] = keys_in[base + row * 32 + lane];
    const uint32_t d = static_cast<uint32_t>(key[k] >> shift) & 0xFF;
    const uint32_t peers = __match_any_sync(0xffffffffu, d);
    rank_in_row[k] = __popc(peers & ((1u << lane) - 1u));
    if (rank_in_row[k] == 0) row_cnt[row][d] = static_cast<uint16_t>(__popc(peers));
  }
  __syncthreads();
  {
    // exclusive prefix over rows, one digit per thread
    const int d = threadIdx.x;
    uint32_t run = 0;
#pragma unroll 4
    for (int r = 0; r < RS_ROWS; ++r) {
      const uint32_t v = row_cnt[r][d];
      row_cnt[r][d] = static_cast<uint16_t>(run);
      run += v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < ROWS_PER_WARP; ++k) {
    const int row = warp * ROWS_PER_WARP + k;
    const uint32_t d = static_cast<uint32_t>(key[k] >> shift) & 0xFF;
    const uint32_t dst = digit_base[d] + offsets[static_cast<size_t>(d) * nblocks + blockIdx.x] + row_cnt[row][d] +
                         rank_in_row[k];
    keys_out[dst] = key[k];
  }
}

// ---- positives prefix ------------------------------------------------------------------------------
__global__ void __launch_bounds__(RS_THREADS)
pos_block_sum_kernel(const unsigned long long* __restrict__ keys, uint32_t* __restrict__ block_sums) {
  __shared__ uint32_t red[RS_THREADS / 32];
  const size_t base = static_cast<size_t>(blockIdx.x) * RS_CHUNK;
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < RS_CHUNK / RS_THREADS; ++k) {
    const unsigned long long key = keys[base + k * RS_THREADS + threadIdx.x];
    s += (key != KEY_PAD) ? static_cast<uint32_t>(key & 1ull) : 0u;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t t = 0;
    for (int w = 0; w < RS_THREADS / 32; ++w) t += red[w];
    block_sums[blockIdx.x] = t;
  }
}

// pos_before[i] = number of positives among sorted keys [0, i); thread t owns 4 consecutive keys
__global__ void __launch_bounds__(RS_THREADS)
pos_prefix_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ block_offsets,
                  uint32_t* __restrict__ pos_before) {
  __shared__ uint32_t warp_tot[RS_THREADS / 32];
  const size_t base = static_cast<size_t>(blockIdx.x) * RS_CHUNK + threadIdx.x * 4;
  uint32_t b[4];
  uint32_t s = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned long long key = keys[base + k];
    b[k] = (key != KEY_PAD) ? static_cast<uint32_t>(key & 1ull) : 0u;
    s += b[k];
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = s;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) warp_tot[warp] = inc;
  __syncthreads();
  uint32_t woff = 0;
  for (int w = 0; w < warp; ++w) woff += warp_tot[w];
  uint32_t run = block_offsets[blockIdx.x] + woff + inc - s;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    pos_before[base + k] = run;
    run += b[k];
  }
}

// ---- rank counts ------------------------------------------------------------------------------------
__device__ __forceinline__ int slot_of(int a, int g1, int max_groups) {
  return a == 0 ? 0 : 1 + (a - 1) * (max_groups + 1) + g1;
}

// first index of every (attribute slot, group, column) segment of the sorted keys: seg_start[slot * 2 + c]
__global__ void __launch_bounds__(RS_THREADS)
segment_start_kernel(const unsigned long long* __restrict__ keys, long long E, uint32_t* __restrict__ seg_start,
                     int max_groups) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < E; i += stride) {
    const unsigned long long k = keys[i];
    if (i > 0 && (keys[i - 1] >> 33) == (k >> 33)) continue;
    const int g1 = static_cast<int>((k >> 34) & 0xFF);
    if (g1 > max_groups) continue;
    seg_start[slot_of(static_cast<int>((k >> 42) & 0xFF), g1, max_groups) * 2 + static_cast<int>((k >> 33) & 1ull)] =
        static_cast<uint32_t>(i);
  }
}

// lower bound of `target` in keys[lo, hi] known to satisfy keys[hi] >= target: gallop back from hi (tie runs are short
// for real-valued scores: typically one or two probes), then bisect the bracket
__device__ __forceinline__ long long lower_bound_back(const unsigned long long* __restrict__ keys, long long lo,
                                                      long long hi, unsigned long long target) {
  long long step = 1;
  while (hi - step >= lo && keys[hi - step] >= target) {
    hi -= step;
    step <<= 1;
  }
  long long l = (hi - step >= lo) ? hi - step + 1 : lo;     // keys[l - 1] < target (or l == lo), keys[hi] >= target
  while (l < hi) {
    const long long mid = (l + hi) >> 1;
    if (keys[mid] < target) l = mid + 1; else hi = mid;
  }
  return l;
}

__global__ void __launch_bounds__(RS_THREADS)
rank_count_kernel(const unsigned long long* __restrict__ keys, const uint32_t* __restrict__ pos_before,
                  const uint32_t* __restrict__ seg_start, long long E, unsigned long long* __restrict__ counts,
                  int n_slots, int max_groups) {
  extern __shared__ unsigned long long cnt_s[];   // [n_slots, 4] : gt0, eq0, gt1, eq1
  for (int i = threadIdx.x; i < n_slots * 4; i += blockDim.x) cnt_s[i] = 0ull;
  __syncthreads();
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < E; i += stride) {
    const unsigned long long k = keys[i];
    if ((k & 1ull) == 0ull) continue;                      // negatives contribute through the positives
    const int g1 = static_cast<int>((k >> 34) & 0xFF);
    const int a = static_cast<int>((k >> 42) & 0xFF);
    if (g1 > max_groups) continue;                          // group id outside the requested range
    const int c = static_cast<int>((k >> 33) & 1ull);
    const int slot = slot_of(a, g1, max_groups);
    const long long seg_lo = seg_start[slot * 2 + c];
    const long long run_pos = lower_bound_back(keys, seg_lo, i, k);              // first positive tied with this score
    const long long run_lo = lower_bound_back(keys, seg_lo, run_pos, k & ~1ull); // first negative tied with this score
    const unsigned long long pos_between = pos_before[run_lo] - pos_before[seg_lo];
    const unsigned long long gt = static_cast<unsigned long long>(run_lo - seg_lo) - pos_between;
    const unsigned long long eq = static_cast<unsigned long long>(run_pos - run_lo);
    atomicAdd(&cnt_s[slot * 4 + c * 2 + 0], gt);
    atomicAdd(&cnt_s[slot * 4 + c * 2 + 1], eq);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_slots * 4; i += blockDim.x) {
    const unsigned long long v = cnt_s[i];
    if (v) atomicAdd(&counts[(i >> 2) * 8 + (i & 3)], v);
  }
}

__global__ void __launch_bounds__(RS_THREADS)
confusion_kernel(const float* __restrict__ prob, const int32_t* __restrict__ label, const int32_t* __restrict__ attrs,
                 unsigned long long* __restrict__ counts, int N, int n_attr, int n_slots, int max_groups) {
  extern __shared__ unsigned long long cnt_s[];   // [n_slots, 4] : tp, fp, tn, fn
  for (int i = threadIdx.x; i < n_slots * 4; i += blockDim.x) cnt_s[i] = 0ull;
  __syncthreads();
  const int stride = gridDim.x * blockDim.x;
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < N; n += stride) {
    const int pred = prob[static_cast<size_t>(n) * 2 + 1] > prob[static_cast<size_t>(n) * 2] ? 1 : 0;  // argmax, ties -> 0
    const int y = label[n];
    const int which = pred ? (y == 1 ? 0 : 1) : (y == 0 ? 2 : 3);
    atomicAdd(&cnt_s[which], 1ull);
    for (int a = 0; a < n_attr; ++a) {
      const int g1 = attrs[static_cast<size_t>(a) * N + n] + 1;
      if (g1 < 0 || g1 > max_groups) continue;
      atomicAdd(&cnt_s[slot_of(a + 1, g1, max_groups) * 4 + which], 1ull);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_slots * 4; i += blockDim.x) {
    const unsigned long long v = cnt_s[i];
    if (v) atomicAdd(&counts[(i >> 2) * 8 + 4 + (i & 3)], v);
  }
}

static size_t a256(size_t v) { return (v + 255) & ~size_t(255); }

struct AucWs {
  unsigned long long *keys_a, *keys_b;
  uint32_t *hist, *block_sums, *pos_before;
  long long E, E_pad;
  int nblocks;
};

static void auc_sizes(int N, int n_attr, long long* E, long long* E_pad, int* nblocks) {
  *E = static_cast<long long>(N) * (n_attr + 1) * 2;
  *nblocks = static_cast<int>((*E + RS_CHUNK - 1) / RS_CHUNK);
  *E_pad = static_cast<long long>(*nblocks) * RS_CHUNK;
}

}  // namespace ffm

using namespace ffm;

extern "C" {

size_t ffm_group_auc_workspace_bytes(int N, int n_attr, int max_groups) {
  long long E, E_pad;
  int nb;
  auc_sizes(N, n_attr, &E, &E_pad, &nb);
  const size_t n_slots = 1 + static_cast<size_t>(n_attr) * (max_groups + 1);
  return 2 * a256(static_cast<size_t>(E_pad) * 8) + a256(static_cast<size_t>(256) * nb * 4) +
         a256(static_cast<size_t>(nb) * 4 + 4) + a256(static_cast<size_t>(E_pad + 1) * 4) + a256(256 * 4) +
         a256(n_slots * 2 * 4);
}

int ffm_group_auc(const float* prob, const int32_t* label, const int32_t* attrs, uint64_t* counts_out,
                  void* workspace, size_t workspace_bytes, int N, int n_attr, int max_groups, cudaStream_t stream) {
  FFM_CHECK_ARG(prob && label && counts_out && workspace, "ffm_group_auc: null pointer argument");
  FFM_CHECK_ARG(n_attr == 0 || attrs != nullptr, "ffm_group_auc: attrs missing");
  FFM_CHECK_ARG(N >= 1 && n_attr >= 0 && n_attr <= 254 && max_groups >= 0 && max_groups <= 254,
                "ffm_group_auc: bad sizes N=%d n_attr=%d max_groups=%d", N, n_attr, max_groups);
  FFM_CHECK_ARG(workspace_bytes >= ffm_group_auc_workspace_bytes(N, n_attr, max_groups),
                "ffm_group_auc: workspace too small");
  long long E, E_pad;
  int nb;
  auc_sizes(N, n_attr, &E, &E_pad, &nb);
  FFM_CHECK_ARG(static_cast<long long>(256) * nb < (1ll << 31) && E_pad < (1ll << 32), "ffm_group_auc: N too large");
  uint8_t* w = static_cast<uint8_t*>(workspace);
  unsigned long long* keys_a = reinterpret_cast<unsigned long long*>(w); w += a256(static_cast<size_t>(E_pad) * 8);
  unsigned long long* keys_b = reinterpret_cast<unsigned long long*>(w); w += a256(static_cast<size_t>(E_pad) * 8);
  uint32_t* hist = reinterpret_cast<uint32_t*>(w); w += a256(static_cast<size_t>(256) * nb * 4);
  uint32_t* block_sums = reinterpret_cast<uint32_t*>(w); w += a256(static_cast<size_t>(nb) * 4 + 4);
  uint32_t* pos_before = reinterpret_cast<uint32_t*>(w); w += a256(static_cast<size_t>(E_pad + 1) * 4);
  uint32_t* totals = reinterpret_cast<uint32_t*>(w); w += a256(256 * 4);
  uint32_t* seg_start = reinterpret_cast<uint32_t*>(w);

  const int n_slots = 1 + n_attr * (max_groups + 1);
  FFM_CHECK_ARG(static_cast<size_t>(n_slots) * 32 <= 48 * 1024, "ffm_group_auc: too many (attribute, group) slots");
  FFM_CHECK_CUDA(cudaMemsetAsync(counts_out, 0, static_cast<size_t>(n_slots) * 8 * sizeof(uint64_t), stream));

  const int egrid = static_cast<int>(std::min<long long>((E_pad + 255) / 256, 4ll * num_sms()));
  expand_keys_kernel<<<egrid, 256, 0, stream>>>(prob, label, attrs, keys_a, N, n_attr, E, E_pad);
  unsigned long long *src = keys_a, *dst = keys_b;
  for (int pass = 0; pass < RS_PASSES; ++pass) {
    const int shift = pass * 8;
    radix_hist_kernel<<<nb, RS_THREADS, 0, stream>>>(src, hist, shift, nb);
    digit_row_scan_kernel<<<256, RS_THREADS, 0, stream>>>(hist, totals, nb);
    radix_scatter_kernel<<<nb, RS_THREADS, 0, stream>>>(src, dst, hist, totals, shift, nb);
    unsigned long long* t = src; src = dst; dst = t;
  }
  // the padding keys (all ones) also carry digit 0xFF in the 8th byte, so 7 passes keep them at the end
  pos_block_sum_kernel<<<nb, RS_THREADS, 0, stream>>>(src, block_sums);
  exclusive_scan_kernel<<<1, 1024, 0, stream>>>(block_sums, nb);
  pos_prefix_kernel<<<nb, RS_THREADS, 0, stream>>>(src, block_sums, pos_before);
  const int rgrid = static_cast<int>(std::min<long long>((E + 255) / 256, 8ll * num_sms()));
  segment_start_kernel<<<rgrid, RS_THREADS, 0, stream>>>(src, E, seg_start, max_groups);
  rank_count_kernel<<<rgrid, RS_THREADS, static_cast<size_t>(n_slots) * 32, stream>>>(
      src, pos_before, seg_start, E, reinterpret_cast<unsigned long long*>(counts_out), n_slots, max_groups);
  const int cgrid = std::min((N + 255) / 256, 4 * num_sms());
  confusion_kernel<<<cgrid, RS_THREADS, static_cast<size_t>(n_slots) * 32, stream>>>(
      prob, label, attrs, reinterpret_cast<unsigned long long*>(counts_out), N, n_attr, n_slots, max_groups);
  FFM_CHECK_CUDA(cudaGetLastError());
  count_launch(1 + 3 * RS_PASSES + 6);
  return FFM_OK;
}

}  // extern "C"
