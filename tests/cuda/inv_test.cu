#include <cstdio>
#include <cuda_runtime.h>
#include "fr_device.cuh"
using namespace pzkd;
__global__ void k(const u64* in, u64* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i >= n) return;
  u64 a[4] = {in[4*i], in[4*i+1], in[4*i+2], in[4*i+3]}, r[4], chk[4];
  fr_inv(r, a);
  fr_mul(chk, r, a);  // should be R (Montgomery one) unless a == 0
  for (int j = 0; j < 4; j++) { out[8*i+j] = r[j]; out[8*i+4+j] = chk[j]; }
}
int main() {
  const int n = 4096; u64 *h = new u64[4*n], *o = new u64[8*n];
  unsigned long long s = 88172645463325252ull;
  for (int i = 0; i < 4*n; i++) { s ^= s << 13; s ^= s >> 7; s ^= s << 17; h[i] = s; }
  for (int i = 0; i < n; i++) h[4*i+3] &= 0x0fffffffffffffffull;   // < p
  for (int j = 0; j < 4; j++) { h[j] = 0; h[4+j] = j == 0; }       // 0 and 1
  u64 *d, *dout; cudaMalloc(&d, 32*n); cudaMalloc(&dout, 64*n);
  cudaMemcpy(d, h, 32*n, cudaMemcpyHostToDevice);
  k<<<n/128, 128>>>(d, dout, n); cudaMemcpy(o, dout, 64*n, cudaMemcpyDeviceToHost);
  const u64 R[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  int bad = 0;
  for (int i = 1; i < n; i++) for (int j = 0; j < 4; j++) if (o[8*i+4+j] != R[j]) { bad++; break; }
  bool zero_ok = (o[0] | o[1] | o[2] | o[3]) == 0;
  printf("inv test: bad=%d zero_ok=%d err=%s\n", bad, (int)zero_ok, cudaGetErrorString(cudaGetLastError()));
  return bad != 0 || !zero_ok;
}
