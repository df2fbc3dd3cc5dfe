// device unit test: PTX Montgomery product == plain-C formulation on random and edge operands
#include <cstdio>
#include <cuda_runtime.h>
#include "fr_device.cuh"
using namespace pzkd;
__global__ void k(const u64* in, u64* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  u64 a[4] = {in[8*i], in[8*i+1], in[8*i+2], in[8*i+3]}, b[4] = {in[8*i+4], in[8*i+5], in[8*i+6], in[8*i+7]}, r[4], q[4];
  fr_mul(r, a, b); fr_mul_plain(q, a, b);
  for (int j = 0; j < 4; j++) { out[8*i+j] = r[j]; out[8*i+4+j] = q[j]; }
}
int main() {
  const int n = 1 << 16; u64 *h = new u64[8*n], *o = new u64[8*n];
  unsigned long long s = 88172645463325252ull;
  const u64 P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
  for (int i = 0; i < 8*n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = s; }
  for (int i = 0; i < 2*n; i++) h[4*i+3] &= 0x0fffffffffffffffull;   // < p
  // edge operands: 0, 1, p-1, all-ones low limbs
  for (int j = 0; j < 4; j++) { h[j] = 0; h[8+j] = P[j]; h[12+j] = P[j]; h[16+j] = (j==0); h[20+j] = P[j]; }
  h[8] -= 1; h[12] -= 1; h[20] -= 1;
  u64 *d, *dout; cudaMalloc(&d, 64*n); cudaMalloc(&dout, 64*n);
  cudaMemcpy(d, h, 64*n, cudaMemcpyHostToDevice);
  k<<<n/128, 128>>>(d, dout, n); cudaMemcpy(o, dout, 64*n, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int i = 0; i < n; i++) for (int j = 0; j < 4; j++) if (o[8*i+j] != o[8*i+4+j]) { bad++; break; }
  printf("mont test: %d of %d differ, err=%s\n", bad, n, cudaGetErrorString(cudaGetLastError()));
  return bad != 0;
}
