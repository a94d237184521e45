// Micro-benchmark (B200): latency of the mbarrier hand-offs the GEMM pipeline is built from.
//   A: arrive -> try_wait wake ping-pong between two warps        B: same, spinning on test_wait
//   C: tcgen05.commit (no MMA pending) -> try_wait, answered by a plain arrive
//   D: ping-pong with all 32 lanes waiting (asm loop) and one elected lane arriving
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I fairfedmed_b200/csrc -o build/mbar_latency tools/micro/mbar_latency.cu
#include <cstdio>
#include "ffm_common.cuh"
using namespace ffm;

// MODE 3: ping-pong only.  WAITERS: 0 none, 1 = 8 extra warps (all lanes) blocked in mbar_wait on a third barrier,
// 2 = same but only lane 0 of each waits, 3 = all lanes in the asm wait (mbar_wait_uniform)
template <int MODE, int WAITERS = 0>
__global__ void pingpong(int iters, long long* out) {
  __shared__ uint64_t bar_a, bar_b, bar_c;
  __shared__ uint32_t tmem_slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) { mbar_init(&bar_a, 1); mbar_init(&bar_b, 1); mbar_init(&bar_c, 1); fence_mbar_init(); }
  if (warp == 2) { tmem_alloc(&tmem_slot, 32); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  long long t0 = clock64();
  if (MODE == 3) {
    if (warp == 0) {
      for (int i = 0; i < iters; ++i) {
        mbar_wait_uniform(&bar_a, i & 1);
        if (elect_one()) mbar_arrive(&bar_b);
        __syncwarp();
      }
    } else if (warp == 1) {
      for (int i = 0; i < iters; ++i) {
        if (elect_one()) mbar_arrive(&bar_a);
        __syncwarp();
        mbar_wait_uniform(&bar_b, i & 1);
      }
    }
  } else if (lane == 0) {
    if (warp == 0) {
      for (int i = 0; i < iters; ++i) {
        if (MODE == 1) { while (!mbar_test_wait(&bar_a, i & 1)) {} } else mbar_wait(&bar_a, i & 1);
        mbar_arrive(&bar_b);
      }
    } else if (warp == 1) {
      for (int i = 0; i < iters; ++i) {
        if (MODE == 2) umma_commit(&bar_a); else mbar_arrive(&bar_a);
        if (MODE == 1) { while (!mbar_test_wait(&bar_b, i & 1)) {} } else mbar_wait(&bar_b, i & 1);
      }
    }
  }
  if (WAITERS && warp >= 3) {
    if (WAITERS == 1) mbar_wait(&bar_c, 0);
    else if (WAITERS == 2) { if (lane == 0) mbar_wait(&bar_c, 0); __syncwarp(); }
    else mbar_wait_uniform(&bar_c, 0);
  }
  long long t1 = clock64();
  if (threadIdx.x == 32) { out[MODE + 4 * WAITERS] = (t1 - t0); if (WAITERS) mbar_arrive(&bar_c); }
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_slot, 32);
}

int main() {
  long long* out; cudaMallocManaged(&out, 256);
  const int iters = 20000;
  for (int rep = 0; rep < 2; ++rep) {
    pingpong<0><<<1, 96>>>(iters, out); cudaDeviceSynchronize();
    pingpong<1><<<1, 96>>>(iters, out); cudaDeviceSynchronize();
    pingpong<2><<<1, 96>>>(iters, out); cudaDeviceSynchronize();
    pingpong<3><<<1, 96>>>(iters, out); cudaDeviceSynchronize();
    pingpong<3, 1><<<1, 96 + 256>>>(iters, out); cudaDeviceSynchronize();
    pingpong<3, 2><<<1, 96 + 256>>>(iters, out); cudaDeviceSynchronize();
    pingpong<3, 3><<<1, 96 + 256>>>(iters, out); cudaDeviceSynchronize();
  }
  printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
  printf("A try_wait ping-pong      : %.1f cycles / round trip (2 hand-offs)\n", (double)out[0] / iters);
  printf("B test_wait spin ping-pong: %.1f cycles / round trip\n", (double)out[1] / iters);
  printf("C tcgen05.commit + arrive : %.1f cycles / round trip\n", (double)out[2] / iters);
  printf("D warp-wide asm wait+elect: %.1f cycles / round trip\n", (double)out[3] / iters);
  printf("E = D + 8 warps x 32 lanes blocked in the C++ try_wait/clock64 loop: %.1f cycles\n", (double)out[3 + 4] / iters);
  printf("F = D + 8 warps, lane 0 only blocked                              : %.1f cycles\n", (double)out[3 + 8] / iters);
  printf("G = D + 8 warps x 32 lanes blocked in the asm try_wait loop       : %.1f cycles\n", (double)out[3 + 12] / iters);
  return 0;
}
