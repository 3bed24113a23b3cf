// micro-test: TMEM as per-warp scratch (tcgen05.alloc / st / ld, shape 32x32b): semantics and latency.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/tmem_rw scripts/micro/tmem_rw.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
               "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
               : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int WARPS>
__global__ void k(uint32_t *out, long long *cyc, int *bad) {
  __shared__ uint32_t s_base;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"((uint32_t)__cvta_generic_to_shared(&s_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t base = s_base;
  constexpr int COLS = 512 / ((WARPS + 3) / 4);
  const uint32_t mine = base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * COLS);
  if (threadIdx.x == 0) out[0] = base;
  // write: cell (column c, lane l) of warp w = w * 1e6 + c * 100 + l
  for (int c0 = 0; c0 < COLS; c0 += 8) {
    uint32_t v[8];
    for (int j = 0; j < 8; j++) v[j] = warp * 1000000 + (c0 + j) * 100 + lane;
    tm_st8(mine + c0, v);
  }
  tm_wait_st();
  __syncthreads();  // everybody has written before anybody checks (detects overlap between warps)
  int nbad = 0;
  for (int c0 = 0; c0 < COLS; c0 += 8) {
    uint32_t v[8];
    tm_ld8(mine + c0, v);
    for (int j = 0; j < 8; j++) nbad += (v[j] != (uint32_t)(warp * 1000000 + (c0 + j) * 100 + lane));
  }
  // unaligned column offset (3) and read-modify-write
  {
    uint32_t v[8];
    tm_ld8(mine + 3, v);
    for (int j = 0; j < 8; j++) nbad += (v[j] != (uint32_t)(warp * 1000000 + (3 + j) * 100 + lane));
    for (int j = 0; j < 8; j++) v[j] += 7;
    tm_st8(mine + 3, v);
    tm_wait_st();
    uint32_t w[8];
    tm_ld8(mine + 3, w);
    for (int j = 0; j < 8; j++) nbad += (w[j] != v[j]);
  }
  atomicAdd(bad, nbad);
  // latency: dependent ld -> st -> ld chain on one column block
  long long t0 = clock64();
  uint32_t v[8];
  for (int it = 0; it < 64; it++) {
    tm_ld8(mine + 16, v);
    for (int j = 0; j < 8; j++) v[j] += 1;
    tm_st8(mine + 16, v);
    tm_wait_st();
  }
  long long t1 = clock64();
  for (int it = 0; it < 64; it++) {
    tm_ld8(mine + 8 * (it & 15), v);
    out[1] += v[0] & (lane == 99);
  }
  long long t2 = clock64();
  if (lane == 0) { cyc[warp * 2] = (t1 - t0) / 64; cyc[warp * 2 + 1] = (t2 - t1) / 64; }
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
  uint32_t *out; long long *cyc; int *bad;
  cudaMalloc(&out, 64); cudaMalloc(&cyc, 16 * 8); cudaMalloc(&bad, 4);
  for (int W : {8, 4}) {
    cudaMemset(bad, 0, 4); cudaMemset(out, 0, 64);
    if (W == 8) k<8><<<148, 256>>>(out, cyc, bad); else k<4><<<148, 128>>>(out, cyc, bad);
    cudaError_t e = cudaDeviceSynchronize();
    int hb; uint32_t ho[2]; long long hc[16];
    cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc, cyc, 16 * 8, cudaMemcpyDeviceToHost);
    printf("WARPS=%d: %s base=0x%x mismatches=%d  ld+st+wait chain %lld cyc/iter, ld+wait %lld cyc/iter\n", W,
           cudaGetErrorString(e), ho[0], hb, hc[0], hc[1]);
  }
  return 0;
}
