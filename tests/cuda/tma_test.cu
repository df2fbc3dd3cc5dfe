// device unit test: 1-D TMA bulk copy + mbarrier helpers of pzk_r1cs.cuh
#include <cstdio>
#include <cuda_runtime.h>
#include "pzk_kernels.cuh"
#include "pzk_r1cs.cuh"
using namespace pzkd;
__global__ void k(const u64* src, u64* dst, int n_tiles) {
  extern __shared__ __align__(16) unsigned char smem[];
  const u32 sbase = (u32)__cvta_generic_to_shared(smem);
  const u32 bar0 = sbase + 2 * 4096;
  if (threadIdx.x == 0) { mbar_init(bar0, 1); mbar_init(bar0 + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  if (threadIdx.x == 0) { mbar_expect_tx(bar0, 4096); tma_load_1d(sbase, src, 2048, bar0); tma_load_1d(sbase + 2048, src + 256, 2048, bar0); }
  u32 ph0 = 0, ph1 = 0;
  for (int t = 0; t < n_tiles; t++) {
    int slot = t & 1;
    if (threadIdx.x == 0 && t + 1 < n_tiles) { mbar_expect_tx(bar0 + 8 * (slot ^ 1), 4096); tma_load_1d(sbase + (slot ^ 1) * 4096, src + (t + 1) * 512, 4096, bar0 + 8 * (slot ^ 1)); }
    if (slot == 0) { mbar_wait(bar0, ph0); ph0 ^= 1; } else { mbar_wait(bar0 + 8, ph1); ph1 ^= 1; }
    const u64* s = reinterpret_cast<const u64*>(smem + slot * 4096);
    for (int i = threadIdx.x; i < 512; i += blockDim.x) dst[t * 512 + i] = s[i] + 1;
    __syncthreads();
  }
}
int main() {
  const int T = 7, n = T * 512;
  u64 *h = new u64[n], *o = new u64[n];
  for (int i = 0; i < n; i++) h[i] = i * 3ull;
  u64 *d, *e; cudaMalloc(&d, n * 8); cudaMalloc(&e, n * 8);
  cudaMemcpy(d, h, n * 8, cudaMemcpyHostToDevice);
  k<<<1, 128, 2 * 4096 + 16>>>(d, e, T);
  cudaError_t err = cudaDeviceSynchronize();
  cudaMemcpy(o, e, n * 8, cudaMemcpyDeviceToHost);
  int bad = 0; for (int i = 0; i < n; i++) bad += o[i] != h[i] + 1;
  printf("tma test: bad=%d err=%s\n", bad, cudaGetErrorString(err));
  return bad != 0;
}
