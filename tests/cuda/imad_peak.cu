// imad_peak.cu - measured denominators for the evaluator's roofline ("a dependent-free IMAD
// microbenchmark"), printed as one JSON line that bench.py reads at run time.
//
//   imad_per_s      : 32-bit multiply-adds per second of the whole GPU, 16 independent mad.lo.u32 chains per
//                     thread, 8 CTAs x 256 threads per SM - the fma-pipe integer issue rate
//   imad_wide_per_s : the same with mad.wide.u32 (32 x 32 -> 64 accumulate), the instruction the Montgomery
//                     product is built from
//   fr_mul_per_s    : Montgomery products (fr_device.cuh, 136 IMAD each) per second with operands in registers,
//                     two independent chains per thread - what eval_kernel could reach with free operands
// Every figure is the best of 5 launches timed with CUDA events; the SM clock the driver reports is printed so
// that the reader can relate the figure to 148 SM x 64 lanes x clock.
#include <cstdio>
#include <cuda_runtime.h>
#include "fr_device.cuh"
using namespace pzkd;

__global__ void __launch_bounds__(256) imad_kernel(u32* out, u32 a, u32 b, int iters) {
  u32 x[16];
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = threadIdx.x + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
  }
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s ^= x[k];
  if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mad.hi.u32: the other half of every 32 x 32 product in a mad.lo.cc / madc.hi.cc Montgomery chain
__global__ void __launch_bounds__(256) imad_hi_kernel(u32* out, u32 a, u32 b, int iters) {
  u32 x[16];
#pragma unroll
  for (int k = 0; k < 16; k++) x[k] = threadIdx.x + k;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 16; k++) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[k]) : "r"(a), "r"(b));
  }
  u32 s = 0;
#pragma unroll
  for (int k = 0; k < 16; k++) s ^= x[k];
  if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) imad_wide_kernel(u64* out, u32 a, u32 b, int iters) {
  u64 x[8];
  u32 y[8];
#pragma unroll
  for (int k = 0; k < 8; k++) { x[k] = threadIdx.x + k; y[k] = a + k; }
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"(y[k]), "r"(b));
#pragma unroll
    for (int k = 0; k < 8; k++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[k]) : "r"(b), "r"(y[k]));
  }
  u64 s = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) s ^= x[k];
  if (s == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(128, 8) frmul_kernel(u64* out, const u64* in, int iters) {
  u64 a[4], b[4], c[4];
#pragma unroll
  for (int j = 0; j < 4; j++) { a[j] = in[j] + threadIdx.x; b[j] = in[4 + j]; c[j] = in[8 + j] ^ threadIdx.x; }
  a[3] &= 0x0fffffffffffffffull; c[3] &= 0x0fffffffffffffffull;
  for (int i = 0; i < iters; i++) {
    fr_mul(a, a, b);
    fr_mul(c, c, b);
  }
  if ((a[0] ^ c[0]) == 0x12345678u) out[blockIdx.x * blockDim.x + threadIdx.x] = a[1];
}

template <typename F>
static double best_ms(F launch) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  double best = 1e30;
  for (int r = 0; r < 6; r++) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    if (r > 0 && ms < best) best = ms;  // the first launch is the warm-up
  }
  return best;
}

int main() {
  int n_sm = 0, clk_khz = 0;
  if (cudaGetDeviceCount(&n_sm) != cudaSuccess || n_sm == 0) { printf("{\"error\": \"no CUDA device\"}\n"); return 1; }
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
  u64* buf;
  cudaMalloc(&buf, 64 << 20);
  u64 h[12] = {0x1234567890abcdefull, 0x0fedcba987654321ull, 0x1111111122222222ull, 0x0333333344444444ull,
               0x5555555566666666ull, 0x7777777788888888ull, 0x99999999aaaaaaaaull, 0x0bbbbbbbccccccccull,
               0xddddddddeeeeeeeeull, 0xffffffff00000000ull, 0x1357913579135791ull, 0x0246802468024680ull};
  u64* din;
  cudaMalloc(&din, sizeof h);
  cudaMemcpy(din, h, sizeof h, cudaMemcpyHostToDevice);
  const int grid = n_sm * 8;
  const int it1 = 4096, it2 = 2048, it3 = 512;
  double ms1 = best_ms([&] { imad_kernel<<<grid, 256>>>((u32*)buf, 0x9e3779b1u, 0x7f4a7c15u, it1); });
  double ms2 = best_ms([&] { imad_wide_kernel<<<grid, 256>>>(buf, 0x9e3779b1u, 0x7f4a7c15u, it2); });
  double ms3 = best_ms([&] { frmul_kernel<<<grid, 128>>>(buf, din, it3); });
  double ms4 = best_ms([&] { imad_hi_kernel<<<grid, 256>>>((u32*)buf, 0x9e3779b1u, 0x7f4a7c15u, it1); });
  cudaError_t e = cudaDeviceSynchronize();
  double imad = (double)grid * 256 * it1 * 16 / (ms1 * 1e-3);
  double wide = (double)grid * 256 * it2 * 16 / (ms2 * 1e-3);
  double frm = (double)grid * 128 * it3 * 2 / (ms3 * 1e-3);
  printf("{\"imad_hi_per_s\": %.6e, \"imad_per_s\": %.6e, \"imad_wide_per_s\": %.6e, \"fr_mul_per_s\": %.6e, \"fr_mul_imad_per_s\": %.6e, "
         "\"n_sm\": %d, \"sm_clock_mhz_attr\": %.1f, \"imad_per_clk_per_sm\": %.2f, \"imad_wide_per_clk_per_sm\": %.2f, "
         "\"ms\": [%.3f, %.3f, %.3f], \"cuda\": \"%s\"}\n",
         (double)grid * 256 * it1 * 16 / (ms4 * 1e-3), imad, wide, frm, frm * 136.0, n_sm, clk_khz / 1e3, imad / n_sm / (clk_khz * 1e3), wide / n_sm / (clk_khz * 1e3),
         ms1, ms2, ms3, cudaGetErrorString(e));
  return e != cudaSuccess;
}
