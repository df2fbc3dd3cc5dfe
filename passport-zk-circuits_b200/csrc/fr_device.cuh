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
// t < 2p (with `carry` the bit above the 256) -> t mod p.  Straight-line: neighbouring lanes almost never agree on
// whether the subtraction is needed, so a branch would run both sides anyway and pay for the reconvergence.
__device__ __forceinline__ void cond_sub_p(u64* t, u32 carry = 0) {
  const u64 p[4] = {P0, P1, P2, P3};
  u64 u[4];
  const u32 br = sub256(u, t, p);
  const bool take = carry || !br;
  t[0] = take ? u[0] : t[0]; t[1] = take ? u[1] : t[1]; t[2] = take ? u[2] : t[2]; t[3] = take ? u[3] : t[3];
}
__device__ __forceinline__ void fr_add(u64* r, const u64* a, const u64* b) {
  u64 t[4];
  u32 c = add256(t, a, b);
  cond_sub_p(t, c);
  r[0] = t[0]; r[1] = t[1]; r[2] = t[2]; r[3] = t[3];
}
__device__ __forceinline__ void fr_sub(u64* r, const u64* a, const u64* b) {
  u64 t[4];
  const u32 br = sub256(t, a, b);
  const u64 p[4] = {br ? P0 : 0, br ? P1 : 0, br ? P2 : 0, br ? P3 : 0};
  add256(t, t, p);
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
  cond_sub_p(r);  // operands < p < 2^254: the result is < 2p and fits 256 bits
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
  cond_sub_p(r, t[8]);
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
// a^-1 mod p for 0 < a < p (plain integers, no Montgomery factor).  One iteration halves u exactly
// once: when u is odd it is first made >= v by a conditional swap of (u, v) and (x1, x2) and v is
// subtracted (both odd, so the difference is even).  Every step is a masked select, so the 32 lanes
// of a warp run one instruction stream and differ only in the trip count (about 1.4 * 254 iterations);
// the nested `while (even) halve` loops this replaces ran, per warp, the maximum over the lanes of
// every inner trip count and cost 4-6 times as many issue slots.
// Invariants: x1 * a == u and x2 * a == v (mod p); v stays odd; u == 0 at the end, v == gcd == 1.
__device__ __forceinline__ void inv_mod_p_core(u64* out, const u64* a) {
  const u64 p[4] = {P0, P1, P2, P3};
  u64 u[4] = {a[0], a[1], a[2], a[3]};
  u64 v[4] = {P0, P1, P2, P3};
  u64 x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  for (int it = 0; it < 1024 && (u[0] | u[1] | u[2] | u[3]) != 0; it++) {
    const u64 m_odd = 0 - (u[0] & 1);
    u64 d[4];
    const u64 m_sw = m_odd & (0 - (u64)sub256(d, u, v));  // borrow: u < v
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const u64 t = (u[i] ^ v[i]) & m_sw; u[i] ^= t; v[i] ^= t;
      const u64 s = (x1[i] ^ x2[i]) & m_sw; x1[i] ^= s; x2[i] ^= s;
    }
    u64 vv[4], xx[4], pm[4];
#pragma unroll
    for (int i = 0; i < 4; i++) { vv[i] = v[i] & m_odd; xx[i] = x2[i] & m_odd; }
    sub256(u, u, vv);
    const u64 m_b = 0 - (u64)sub256(x1, x1, xx);
#pragma unroll
    for (int i = 0; i < 4; i++) pm[i] = p[i] & m_b;
    add256(x1, x1, pm);
    shr1_256(u);
    const u64 m_h = 0 - (x1[0] & 1);
#pragma unroll
    for (int i = 0; i < 4; i++) pm[i] = p[i] & m_h;
    const u32 c = add256(x1, x1, pm);
    shr1_256(x1);
    x1[3] |= (u64)c << 63;
  }
  out[0] = x2[0]; out[1] = x2[1]; out[2] = x2[2]; out[3] = x2[3];
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
// 256-bit shifts by a run-time amount, on registers only: the limb move is two conditional swaps, not an indexed
// array (which the compiler would put in local memory)
__device__ __forceinline__ void shr256(u64* r, const u64* a, unsigned s) {
  u64 x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3];
  if (s >= 256) { x0 = x1 = x2 = x3 = 0; s = 0; }
  if (s & 64) { x0 = x1; x1 = x2; x2 = x3; x3 = 0; }
  if (s & 128) { x0 = x2; x1 = x3; x2 = 0; x3 = 0; }
  const unsigned bs = s & 63;
  if (bs) {
    x0 = (x0 >> bs) | (x1 << (64 - bs));
    x1 = (x1 >> bs) | (x2 << (64 - bs));
    x2 = (x2 >> bs) | (x3 << (64 - bs));
    x3 >>= bs;
  }
  r[0] = x0; r[1] = x1; r[2] = x2; r[3] = x3;
}
__device__ __forceinline__ void shl256(u64* r, const u64* a, unsigned s) {
  u64 x0 = a[0], x1 = a[1], x2 = a[2], x3 = a[3];
  if (s >= 256) { x0 = x1 = x2 = x3 = 0; s = 0; }
  if (s & 64) { x3 = x2; x2 = x1; x1 = x0; x0 = 0; }
  if (s & 128) { x3 = x1; x2 = x0; x1 = 0; x0 = 0; }
  const unsigned bs = s & 63;
  if (bs) {
    x3 = (x3 << bs) | (x2 >> (64 - bs));
    x2 = (x2 << bs) | (x1 >> (64 - bs));
    x1 = (x1 << bs) | (x0 >> (64 - bs));
    x0 <<= bs;
  }
  r[0] = x0; r[1] = x1; r[2] = x2; r[3] = x3;
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
