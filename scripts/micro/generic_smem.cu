// micro-benchmark: latency of dependent loads / RMW through shared vs generic pointers vs global (L2)
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(float *gbuf, long long *out, int sel) {
  __shared__ float s[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) { s[i] = (float)((i * 7 + 1) & 4095); gbuf[i] = s[i]; }
  __syncthreads();
  float *p = sel == 0 ? s : gbuf;          // generic pointer, decided at run time
  long long t0 = 0, t1 = 0, t2 = 0, t3 = 0;
  if (threadIdx.x < 32) {
    // 1. dependent chain through a true __shared__ access
    int idx = threadIdx.x;
    t0 = clock64();
    for (int i = 0; i < 256; i++) idx = (int)s[idx];
    t1 = clock64();
    // 2. dependent chain through the generic pointer
    int idy = threadIdx.x;
    for (int i = 0; i < 256; i++) idy = (int)p[idy];
    t2 = clock64();
    // 3. read-modify-write stream through the generic pointer (like step 6)
    float acc = 0.f;
    for (int r = 0; r < 64; r++) { float v = p[r * 61 + threadIdx.x]; v = v + 1.0f; p[r * 61 + threadIdx.x] = v; acc += (v == 0.f); }
    t3 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t1; out[2] = t3 - t2; out[3] = idx + idy + (int)acc; }
  }
}
int main() {
  float *g; long long *o, h[4];
  cudaMalloc(&g, 4096 * 4); cudaMalloc(&o, 32);
  for (int sel = 0; sel < 2; sel++) {
    for (int rep = 0; rep < 2; rep++) { k<<<1, 128>>>(g, o, sel); cudaDeviceSynchronize(); }
    cudaMemcpy(h, o, 32, cudaMemcpyDeviceToHost);
    printf("%s: shared chain %.1f cyc/load, generic chain %.1f cyc/load, generic RMW row %.1f cyc/row\n",
           sel == 0 ? "generic->shared" : "generic->global", h[0] / 256.0, h[1] / 256.0, h[2] / 64.0);
  }
  return 0;
}
