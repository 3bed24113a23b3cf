// common.cuh — error plumbing and block-level primitives shared by the kernels.
//
// Numerics contract for every .cu in this directory: compiled with -fmad=false,
// so a*b+c is never contracted; every fused multiply-add the algorithms need is
// written as fma().  double '/' and sqrt() are IEEE round-to-nearest on sm_100a,
// float '/' is IEEE because --use_fast_math is not used.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "w2t.h"

namespace w2t {

void set_last_error(const char *fmt, ...);
int cuda_fail(cudaError_t e, const char *what);

#define W2T_CUDA_TRY(expr)                                   \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return ::w2t::cuda_fail(_e, #expr); \
  } while (0)

constexpr int kWarp = 32;

__device__ __forceinline__ int lane_id() { return threadIdx.x & 31; }
__device__ __forceinline__ int warp_id() { return threadIdx.x >> 5; }

// Exclusive scan of two 0/1 flags over one block-sized chunk.
// All threads of the block must call it.  scratch: 2 * (BLOCK/32) ints of shared memory.
// Returns the exclusive prefix of each flag inside the chunk and the chunk totals.
template <int BLOCK>
__device__ __forceinline__ void block_scan2(bool fa, bool fb, int *scratch, int &ex_a, int &ex_b,
                                            int &tot_a, int &tot_b) {
  constexpr int NW = BLOCK / 32;
  const unsigned ba = __ballot_sync(0xffffffffu, fa);
  const unsigned bb = __ballot_sync(0xffffffffu, fb);
  const unsigned lt = (1u << lane_id()) - 1u;
  const int w = warp_id();
  if (lane_id() == 0) {
    scratch[w] = __popc(ba);
    scratch[NW + w] = __popc(bb);
  }
  __syncthreads();
  int pa = 0, pb = 0, ta = 0, tb = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) {
    const int ca = scratch[i], cb = scratch[NW + i];
    if (i < w) { pa += ca; pb += cb; }
    ta += ca;
    tb += cb;
  }
  ex_a = pa + __popc(ba & lt);
  ex_b = pb + __popc(bb & lt);
  tot_a = ta;
  tot_b = tb;
  __syncthreads();  // scratch may be reused right away
}

}  // namespace w2t
