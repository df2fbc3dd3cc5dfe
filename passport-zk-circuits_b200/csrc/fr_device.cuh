// fr_device.cuh - BN254 Fr arithmetic for sm_100a, 4 x 64-bit Montgomery limbs, R = 2^256.
//
// The multiplier works on 32-bit halves (CIOS over 8 limbs): Blackwell's integer multiplier is
// the 32-bit IMAD on the fma pipe (64 IMAD/clk/SM) and a 64-bit product is four of them, so the
// product is expressed directly as 32x32->64 multiply-accumulates (IMAD.WIDE).  One Montgomery
// product = 8x8 product terms + 8x8 reduction terms + 8 m computations = 136 IMAD (the figure
// DESIGN.md uses for the IMAD roofline).
#pragma once
#include <cstdint>

namespace pzkd {

typedef unsigned long long u64;
typedef unsigned int u32;

struct Fr { u64 v[4]; };

__device__ __constant__ u64 FR_P_C[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull,
                                         0xb85045b68181585dull, 0x30644e72e131a029ull};
#define P0 0x43e1f593f0000001ull
#define P1 0x2833e84879b97091ull
#define P2 0xb85045b68181585dull
#define P3 0x30644e72e131a029ull
#define PINV32 0xefffffffu  // -p^-1 mod 2^32

__device__ __forceinline__ u64 addc(u64 a, u64 b, u32& carry) {
  u64 r;
  asm("{\n\t.reg .u32 c;\n\tadd.cc.u64 %0, %2, %3;\n\taddc.u32 c, 0, 0;\n\tmov.u32 %1, c;\n\t}" : "=l"(r), "=r"(carry) : "l"(a), "l"(b));
  return r;
}

// r = a + b (256-bit), returns carry out
__device__ __forceinline__ u32 add256(u64* r, const u64* a, const u64* b) {
  u32 c;
  asm("add.cc.u64 %0, %5, %9;\n\t"
      "addc.cc.u64 %1, %6, %10;\n\t"
      "addc.cc.u64 %2, %7, %11;\n\t"
      "addc.cc.u64 %3, %8, %12;\n\t"
      "addc.u32 %4, 0, 0;"
      : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]), "=r"(c)
      : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]));
  return c;
}
// r = a - b (256-bit), returns borrow (1 when a < b)
__device__ __forceinline__ u32 sub256(u64* r, const u64* a, const u64* b) {
  u32 c;
  asm("sub.cc.u64 %0, %5, %9;\n\t"
      "subc.cc.u64 %1, %6, %10;\n\t"
      "subc.cc.u64 %2, %7, %11;\n\t"
      "subc.cc.u64 %3, %8, %12;\n\t"
      "subc.u32 %4, 0, 0;"
      : "=l"(r[0]), "=l"(r[1]), "=l"(r[2]), "=l"(r[3]), "=r"(c)
      : "l"(a[0]), "l"(a[1]), "l"(a[2]), "l"(a[3]), "l"(b[0]), "l"(b[1]), "l"(b[2]), "l"(b[3]));
  return c & 1;
}
__device__ __forceinline__ bool geq_p(const u64* a) {
  if (a[3] != P3) return a[3] > P3;
  if (a[2] != P2) return a[2] > P2;
  if (a[1] != P1) return a[1] > P1;
  return a[0] >= P0;
}
__device__ __forceinline__ void sub_p(u64* a) {
  const u64 p[4] = {P0, P1, P2, P3};
  sub256(a, a, p);
}
__device__ __forceinline__ void fr_add(u64* r, const u64* a, const u64* b) {
  u64 t[4];
  u32 c = add256(t, a, b);
  if (c || geq_p(t)) sub_p(t);
  r[0] = t[0]; r[1] = t[1]; r[2] = t[2]; r[3] = t[3];
}
__device__ __forceinline__ void fr_sub(u64* r, const u64* a, const u64* b) {
  u64 t[4];
  u32 br = sub256(t, a, b);
  if (br) { const u64 p[4] = {P0, P1, P2, P3}; add256(t, t, p); }
  r[0] = t[0]; r[1] = t[1]; r[2] = t[2]; r[3] = t[3];
}
__device__ __forceinline__ bool fr_is_zero(const u64* a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
__device__ __forceinline__ bool fr_eq(const u64* a, const u64* b) {
  return ((a[0] ^ b[0]) | (a[1] ^ b[1]) | (a[2] ^ b[2]) | (a[3] ^ b[3])) == 0;
}

// ---- Montgomery product on 8 x 32-bit limbs, a*b*2^-256 mod p -----------------------------
// Coarsely-integrated operand scanning with two accumulators ("even" / "odd" columns): the
// 32x32 products of one row are split by column parity so that each parity is ONE uninterrupted
// mad.lo.cc / madc.hi.cc carry chain - no per-product carry fix-ups (IADD3), which is what the
// plain C formulation compiles to (~600 SASS instructions against ~300 here).
namespace mont32 {
__device__ __forceinline__ void mul_n(u32* acc, const u32* a, u32 bi) {  // acc[j], acc[j+1] = a[j]*bi, j even
#pragma unroll
  for (int j = 0; j < 8; j += 2)
    asm("mul.lo.u32 %0, %2, %3; mul.hi.u32 %1, %2, %3;" : "=r"(acc[j]), "=r"(acc[j + 1]) : "r"(a[j]), "r"(bi));
}
__device__ __forceinline__ void cmad_n(u32* acc, const u32* a, u32 bi) {  // acc += a[even]*bi, carry out in CC
  asm("mad.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[0]), "+r"(acc[1]) : "r"(a[0]), "r"(bi));
#pragma unroll
  for (int j = 2; j < 8; j += 2)
    asm("madc.lo.cc.u32 %0, %2, %3, %0; madc.hi.cc.u32 %1, %2, %3, %1;" : "+r"(acc[j]), "+r"(acc[j + 1]) : "r"(a[j]), "r"(bi));
}
__device__ __forceinline__ void madc_n_rshift(u32* odd, const u32* a, u32 bi) {  // odd = (odd >> 64) + a[even]*bi
#pragma unroll
  for (int j = 0; j < 6; j += 2)
    asm("madc.lo.cc.u32 %0, %2, %3, %4; madc.hi.cc.u32 %1, %2, %3, %5;"
        : "=r"(odd[j]), "=r"(odd[j + 1]) : "r"(a[j]), "r"(bi), "r"(odd[j + 2]), "r"(odd[j + 3]));
  asm("madc.lo.cc.u32 %0, %2, %3, 0; madc.hi.u32 %1, %2, %3, 0;" : "=r"(odd[6]), "=r"(odd[7]) : "r"(a[6]), "r"(bi));
}
__device__ __forceinline__ void mad_n_redc(u32* even, u32* odd, const u32* a, u32 bi, const u32* p, bool first) {
  if (first) {
    mul_n(odd, a + 1, bi);
    mul_n(even, a, bi);
  } else {
    asm("add.cc.u32 %0, %0, %1;" : "+r"(even[0]) : "r"(odd[1]));
    madc_n_rshift(odd, a + 1, bi);
    cmad_n(even, a, bi);
    asm("addc.u32 %0, %0, 0;" : "+r"(odd[7]));
  }
  u32 mi = even[0] * PINV32;
  cmad_n(odd, p + 1, mi);
  cmad_n(even, p, mi);
  asm("addc.u32 %0, %0, 0;" : "+r"(odd[7]));
}
}  // namespace mont32

__device__ __forceinline__ void fr_mul(u64* r64, const u64* a64, const u64* b64) {
  const u32 p[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                    0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  u32 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    a[2 * i] = (u32)a64[i]; a[2 * i + 1] = (u32)(a64[i] >> 32);
    b[2 * i] = (u32)b64[i]; b[2 * i + 1] = (u32)(b64[i] >> 32);
  }
  u32 even[8], odd[8];
#pragma unroll
  for (int i = 0; i < 8; i += 2) {
    mont32::mad_n_redc(even, odd, a, b[i], p, i == 0);
    mont32::mad_n_redc(odd, even, a, b[i + 1], p, false);
  }
  // merge the two accumulators: even += odd >> 32
  asm("add.cc.u32 %0, %0, %1;" : "+r"(even[0]) : "r"(odd[1]));
#pragma unroll
  for (int i = 1; i < 7; i++) asm("addc.cc.u32 %0, %0, %1;" : "+r"(even[i]) : "r"(odd[i + 1]));
  asm("addc.u32 %0, %0, 0;" : "+r"(even[7]));
  u64 r[4];
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = (u64)even[2 * i] | ((u64)even[2 * i + 1] << 32);
  if (geq_p(r)) sub_p(r);  // operands < p < 2^254: the result is < 2p and fits 256 bits
  r64[0] = r[0]; r64[1] = r[1]; r64[2] = r[2]; r64[3] = r[3];
}

// reference formulation (plain C, kept for the device unit test tests/cuda/mont_test.cu)
__device__ __forceinline__ void fr_mul_plain(u64* r64, const u64* a64, const u64* b64) {
  const u32 p[8] = {0xf0000001u, 0x43e1f593u, 0x79b97091u, 0x2833e848u,
                    0x8181585du, 0xb85045b6u, 0xe131a029u, 0x30644e72u};
  u32 a[8], b[8];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    a[2 * i] = (u32)a64[i]; a[2 * i + 1] = (u32)(a64[i] >> 32);
    b[2 * i] = (u32)b64[i]; b[2 * i + 1] = (u32)(b64[i] >> 32);
  }
  u32 t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    u32 bi = b[i];
    u32 carry = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      u64 acc = (u64)a[j] * bi + t[j] + carry;
      t[j] = (u32)acc;
      carry = (u32)(acc >> 32);
    }
    u64 top = (u64)t[8] + carry;
    t[8] = (u32)top;
    t[9] = (u32)(top >> 32);
    u32 m = t[0] * PINV32;
    u64 acc = (u64)m * p[0] + t[0];
    carry = (u32)(acc >> 32);
#pragma unroll
    for (int j = 1; j < 8; j++) {
      acc = (u64)m * p[j] + t[j] + carry;
      t[j - 1] = (u32)acc;
      carry = (u32)(acc >> 32);
    }
    top = (u64)t[8] + carry;
    t[7] = (u32)top;
    t[8] = t[9] + (u32)(top >> 32);
  }
  u64 r[4];
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = (u64)t[2 * i] | ((u64)t[2 * i + 1] << 32);
  if (t[8] || geq_p(r)) sub_p(r);
  r64[0] = r[0]; r64[1] = r[1]; r64[2] = r[2]; r64[3] = r[3];
}

__device__ __forceinline__ void fr_to_mont(u64* r, const u64* a) {
  const u64 r2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  fr_mul(r, a, r2);
}
__device__ __forceinline__ void fr_from_mont(u64* r, const u64* a) {
  const u64 one[4] = {1, 0, 0, 0};
  fr_mul(r, a, one);
}
__device__ __forceinline__ void fr_neg(u64* r, const u64* a) {
  if (fr_is_zero(a)) { r[0] = r[1] = r[2] = r[3] = 0; return; }
  const u64 p[4] = {P0, P1, P2, P3};
  sub256(r, p, a);
}

// Modular inverse, Montgomery in / out, inv(0) = 0 (circom run-time semantics: x / 0 = 0).
// Binary extended Euclid on (u, v) = (a, p) with cofactors (x1, x2) mod p: about 380 shift / subtract
// steps of 256-bit integer work (no multiplications) instead of the ~380 Montgomery products of a
// Fermat ladder.  Input aR gives (aR)^-1 = a^-1 R^-1; one product with R^3 restores a^-1 R.
__device__ __forceinline__ void shr1_256(u64* a) {
  a[0] = (a[0] >> 1) | (a[1] << 63); a[1] = (a[1] >> 1) | (a[2] << 63);
  a[2] = (a[2] >> 1) | (a[3] << 63); a[3] >>= 1;
}
__device__ __forceinline__ void half_mod_p(u64* x) {  // x/2 mod p for x in [0, p)
  if (x[0] & 1) {
    const u64 p[4] = {P0, P1, P2, P3};
    u32 c = add256(x, x, p);  // < 2^255, carry is always 0 but keep it exact
    shr1_256(x);
    x[3] |= (u64)c << 63;
  } else shr1_256(x);
}
__device__ __forceinline__ bool geq256(const u64* a, const u64* b) {
  if (a[3] != b[3]) return a[3] > b[3];
  if (a[2] != b[2]) return a[2] > b[2];
  if (a[1] != b[1]) return a[1] > b[1];
  return a[0] >= b[0];
}
// ---- inversion by approximated binary GCD rounds -------------------------------------------------------
// a^-1 mod p for 0 < a < p (plain integers).  Binary GCD in 17 rounds of 31 steps, after Pornin, "Optimized
// Binary GCD for Modular Inversion" (2020): the 31 steps of a round run on 64-bit approximations of (a, b) -
// their top 33 and low 31 bits - and four small cofactors; the round then applies the 2x2 cofactor matrix once
// to the full-width (a, b), an exact division by 2^31, and to (u, v) with a Montgomery-style division by 2^31
// modulo p, so that u * y == a and v * y == b (mod p) after every round.  2 * 254 - 1 = 507 < 17 * 31 steps always
// suffice; then a == 0, b == 1 and v == y^-1.  Every step is a masked select: one instruction stream per warp
// and a fixed trip count, about 25 k instructions instead of the 70 k of the bit-serial version.
__device__ __forceinline__ void inv2_lincomb(u64* t, const u64* x, u64 f, const u64* y, u64 g) {
  // t[0..4] (two's complement, mod 2^320) = x * f + y * g; x, y unsigned 256-bit, f, g signed 64-bit
  typedef unsigned __int128 u128;
  u64 c = 0, d = 0;
  u128 acc = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const u128 p1 = (u128)x[i] * f + c;
    const u128 p2 = (u128)y[i] * g + d;
    c = (u64)(p1 >> 64); d = (u64)(p2 >> 64);
    acc += (u128)(u64)p1 + (u64)p2;
    t[i] = (u64)acc; acc >>= 64;
  }
  t[4] = (u64)acc + c + d;
  // (u64)f is f + 2^64 for a negative f: take x * 2^64 back out (same for g)
  const u64 mf = 0 - (f >> 63), mg = 0 - (g >> 63);
  u64 borrow = 0;
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const u128 s = (u128)(x[i] & mf) + (y[i] & mg) + borrow;
    const u128 dd = (u128)t[i + 1] - s;
    t[i + 1] = (u64)dd;
    borrow = (u64)(0 - (u64)(dd >> 64)) & 3;   // 0, 1 or 2 limbs borrowed
  }
}
__device__ __forceinline__ void inv2_sar31(u64* r, const u64* t) {  // r[0..3] = low 256 bits of (t >> 31), t 5 limbs
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = (t[i] >> 31) | (t[i + 1] << 33);
}
__device__ __forceinline__ void inv_mod_p_core(u64* out, const u64* y) {
  typedef unsigned __int128 u128;
  const u64 p[4] = {P0, P1, P2, P3};
  const u64 M31 = 0x7fffffffull;
  const u64 MINV = 0x6fffffffull;  // -p^-1 mod 2^31
  u64 a[4] = {y[0], y[1], y[2], y[3]}, b[4] = {P0, P1, P2, P3};
  u64 u[4] = {1, 0, 0, 0}, v[4] = {0, 0, 0, 0};
  for (int round = 0; round < 17; round++) {
    // 64-bit approximations: low 31 bits and the 33 bits below the common top bit
    u64 a_, b_;
    {
      const u64 t3 = a[3] | b[3], t2 = a[2] | b[2], t1 = a[1] | b[1];
      const int j = t3 ? 3 : t2 ? 2 : t1 ? 1 : 0;
      const u64 top = j == 3 ? t3 : j == 2 ? t2 : j == 1 ? t1 : (a[0] | b[0]);
      const int n = 64 * j + 64 - (top ? __clzll((long long)top) : 64);
      if (n <= 64) { a_ = a[0]; b_ = b[0]; }
      else {
        const int s = n - 33, idx = s >> 6, off = s & 63;
        const u64 ah = idx < 3 ? a[idx + 1] : 0, bh = idx < 3 ? b[idx + 1] : 0;
        const u64 al = a[idx], bl = b[idx];
        const u64 ta = off ? ((al >> off) | (ah << (64 - off))) : al;
        const u64 tb = off ? ((bl >> off) | (bh << (64 - off))) : bl;
        a_ = (a[0] & M31) | ((ta & 0x1ffffffffull) << 31);
        b_ = (b[0] & M31) | ((tb & 0x1ffffffffull) << 31);
      }
    }
    u64 f0 = 1, g0 = 0, f1 = 0, g1 = 1;
#pragma unroll
    for (int i = 0; i < 31; i++) {
      const u64 odd = 0 - (a_ & 1);
      const u64 sw = odd & (0 - (u64)(a_ < b_));
      u64 t = (a_ ^ b_) & sw; a_ ^= t; b_ ^= t;
      t = (f0 ^ f1) & sw; f0 ^= t; f1 ^= t;
      t = (g0 ^ g1) & sw; g0 ^= t; g1 ^= t;
      a_ -= b_ & odd; f0 -= f1 & odd; g0 -= g1 & odd;
      a_ >>= 1; f1 <<= 1; g1 <<= 1;
    }
    // (a, b) <- |(a f0 + b g0, a f1 + b g1)| / 2^31, the signs move into the cofactors
    u64 ta[5], tb[5], na[4], nb[4];
    inv2_lincomb(ta, a, f0, b, g0);
    inv2_lincomb(tb, a, f1, b, g1);
    {
      const u64 sa = 0 - (ta[4] >> 63), sb = 0 - (tb[4] >> 63);
      inv2_sar31(na, ta); inv2_sar31(nb, tb);
      // conditional negation of the 256-bit results and of the cofactors
      u64 ca = sa & 1, cb = sb & 1;
#pragma unroll
      for (int i = 0; i < 4; i++) {
        const u64 xa = (na[i] ^ sa) + ca; ca = (xa < ca) ? 1 : 0; na[i] = xa;
        const u64 xb = (nb[i] ^ sb) + cb; cb = (xb < cb) ? 1 : 0; nb[i] = xb;
      }
      f0 = (f0 ^ sa) - sa; g0 = (g0 ^ sa) - sa;
      f1 = (f1 ^ sb) - sb; g1 = (g1 ^ sb) - sb;
    }
    // (u, v) <- (u f0 + v g0, u f1 + v g1) / 2^31 mod p
    u64 tu[5], tv[5];
    inv2_lincomb(tu, u, f0, v, g0);
    inv2_lincomb(tv, u, f1, v, g1);
#pragma unroll
    for (int w = 0; w < 2; w++) {
      u64* t = w ? tv : tu;
      const u64 k = ((t[0] & M31) * MINV) & M31;
      // t += k * p  (k < 2^31)
      u128 acc = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { acc += (u128)p[i] * k + t[i]; t[i] = (u64)acc; acc >>= 64; }
      t[4] += (u64)acc;
      u64 r[4];
      inv2_sar31(r, t);
      const u64 neg = 0 - (t[4] >> 63);          // t in (-p, 2p): add p when negative
      u64 c = 0;
#pragma unroll
      for (int i = 0; i < 4; i++) { const u128 s = (u128)r[i] + (p[i] & neg) + c; r[i] = (u64)s; c = (u64)(s >> 64); }
      u64 d[4]; u64 br = 0;                       // subtract p when r >= p
#pragma unroll
      for (int i = 0; i < 4; i++) { const u128 s = (u128)r[i] - p[i] - br; d[i] = (u64)s; br = (u64)(s >> 64) & 1; }
      // after the sar the value may also carry bit 256 set (2p > 2^255 fits 256 bits, so no)
      const u64 ge = 0 - (u64)(1 - br);
      u64* dst = w ? v : u;
#pragma unroll
      for (int i = 0; i < 4; i++) dst[i] = (d[i] & ge) | (r[i] & ~ge);
    }
#pragma unroll
    for (int i = 0; i < 4; i++) { a[i] = na[i]; b[i] = nb[i]; }
  }
  out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; out[3] = v[3];
}
__device__ __noinline__ void fr_inv(u64* r, const u64* a) {
  if (fr_is_zero(a)) { r[0] = r[1] = r[2] = r[3] = 0; return; }
  // input aR gives (aR)^-1 = a^-1 R^-1; one product with R^3 restores a^-1 R
  const u64 r3[4] = {0x5e94d8e1b4bf0040ull, 0x2a489cbe1cfbb6b8ull, 0x893cc664a19fcfedull, 0x0cf8594b7fcc657cull};
  u64 t[4];
  inv_mod_p_core(t, a);
  fr_mul(r, t, r3);
}

// ---- plain 256-bit integer helpers (class N) ----------------------------------------------
__device__ __forceinline__ int cmp256(const u64* a, const u64* b) {
#pragma unroll
  for (int i = 3; i >= 0; i--) { if (a[i] < b[i]) return -1; if (a[i] > b[i]) return 1; }
  return 0;
}
__device__ __forceinline__ void shr256(u64* r, const u64* a, unsigned s) {
  u64 t[4] = {0, 0, 0, 0};
  if (s < 256) {
    unsigned ws = s >> 6, bs = s & 63;
    for (unsigned i = 0; i + ws < 4; i++) {
      t[i] = a[i + ws] >> bs;
      if (bs && i + ws + 1 < 4) t[i] |= a[i + ws + 1] << (64 - bs);
    }
  }
  r[0] = t[0]; r[1] = t[1]; r[2] = t[2]; r[3] = t[3];
}
__device__ __forceinline__ void shl256(u64* r, const u64* a, unsigned s) {
  u64 t[4] = {0, 0, 0, 0};
  if (s < 256) {
    unsigned ws = s >> 6, bs = s & 63;
    for (int i = 3; i >= (int)ws; i--) {
      t[i] = a[i - ws] << bs;
      if (bs && i - (int)ws - 1 >= 0) t[i] |= a[i - ws - 1] >> (64 - bs);
    }
  }
  r[0] = t[0]; r[1] = t[1]; r[2] = t[2]; r[3] = t[3];
}
__device__ __noinline__ void divmod256(const u64* a, const u64* b, u64* q, u64* m) {
  u64 qq[4] = {0, 0, 0, 0}, rr[4] = {0, 0, 0, 0};
  if ((b[0] | b[1] | b[2] | b[3]) != 0) {
    for (int i = 255; i >= 0; i--) {
      shl256(rr, rr, 1);
      rr[0] |= (a[i >> 6] >> (i & 63)) & 1;
      if (cmp256(rr, b) >= 0) { sub256(rr, rr, b); qq[i >> 6] |= 1ull << (i & 63); }
    }
  }
  if (q) { q[0] = qq[0]; q[1] = qq[1]; q[2] = qq[2]; q[3] = qq[3]; }
  if (m) { m[0] = rr[0]; m[1] = rr[1]; m[2] = rr[2]; m[3] = rr[3]; }
}
__device__ __forceinline__ void reduce_p(u64* a) {
  while (geq_p(a)) sub_p(a);
}
__device__ __forceinline__ bool is_neg_rep(const u64* a) {  // a > p/2
  const u64 h[4] = {0xa1f0fac9f8000000ull, 0x9419f4243cdcb848ull, 0xdc2822db40c0ac2eull, 0x183227397098d014ull};
  return cmp256(a, h) > 0;
}
__device__ __forceinline__ int scmp256(const u64* a, const u64* b) {
  bool na = is_neg_rep(a), nb = is_neg_rep(b);
  if (na != nb) return na ? -1 : 1;
  return cmp256(a, b);
}

// ---- 128 / 64 -> 64 division (Hacker's Delight divlu), hi < d required ----------------------
__device__ __forceinline__ u64 div128by64(u64 hi, u64 lo, u64 d, u64* rem) {
  const u64 b = 1ull << 32;
  int s = __clzll(d);
  d <<= s;
  u64 vn1 = d >> 32, vn0 = d & 0xffffffffull;
  u64 un32 = s ? ((hi << s) | (lo >> (64 - s))) : hi;
  u64 un10 = lo << s;
  u64 un1 = un10 >> 32, un0 = un10 & 0xffffffffull;
  u64 q1 = un32 / vn1, rhat = un32 - q1 * vn1;
  while (q1 >= b || q1 * vn0 > b * rhat + un1) { q1--; rhat += vn1; if (rhat >= b) break; }
  u64 un21 = un32 * b + un1 - q1 * d;
  u64 q0 = un21 / vn1;
  rhat = un21 - q0 * vn1;
  while (q0 >= b || q0 * vn0 > b * rhat + un0) { q0--; rhat += vn1; if (rhat >= b) break; }
  if (rem) *rem = (un21 * b + un0 - q0 * d) >> s;
  return q1 * b + q0;
}

}  // namespace pzkd
