// pzk_kernels.cuh - the three sm_100a kernels of the hot path.
//
//   eval_kernel   : lane-per-witness interpreter of the typed linear SSA program
//                   (what circom's generated wasm does for the reference, one passport at a time:
//                   /root/reference/test/automatisationTest.js:40-50).
//                   Constraint rows are fused into the op stream (CHECK_* records); the
//                   stand-alone checker for explicit witnesses is in pzk_r1cs.cuh.
//   export_kernel : slot planes -> canonical 32-byte little-endian wires (.wtns section 2).
//
// Data layout (HBM): slot-major, lane-minor planes so that the 32 lanes of a warp touch one
// 256-byte run per limb:   U[slot][lane] (u64)   F[slot][limb][lane] (4 x u64, Montgomery).
// The op / row / term streams are identical for every lane: every warp reads them with
// warp-uniform addresses (one L1/L2 transaction per warp, served from L2 after the first CTA).
#pragma once
#include "fr_device.cuh"
#include "pzk_program.h"

namespace pzkd {

#define PZK_LANE_BLOCK 128

struct StreamCoefs {
  const PzkCoef* coefs;
  const unsigned char* coef_kind;
  const u64* coef_mag;
};

struct EvalParams {
  const uint4* ops;
  u64 n_rec;
  u64* U;
  u64* F;
  u64 L;        // lanes in the tile (planes are blocked: [lane / 128][slot][lane % 128])
  u64 n_lanes;  // active lanes in this tile
  u32 n_u_slots, n_f_slots;
  const u64* fpool;
  const u32* list;
  const u64* inputs;  // [lane][n_inputs][4], or packed records when in_table != nullptr
  u32 n_inputs;
  const uint2* in_table;  // packed inputs: per input (kind 0 = u8, 1 = u64, 2 = 32-byte field; byte offset)
  u32 in_stride;          // bytes per lane of a packed record
  u32* status;
  // fused constraint rows
  int check_rows;
  int store_all;  // 1: every value reaches its global slot (witness export requested)
  StreamCoefs sc;
  unsigned long long* first_bad;
  // fused witness digest (see "fused digest" below): 0 = off (descriptors are skipped)
  int digest;
  const ulonglong2* dig_tab;  // per-bit coefficient tables
  u64* dig_state;             // [DIG_STATE_PIECES][dig_stride] carry-save accumulators
  u64 dig_stride;
  u64 dig_lane_base;
  u32 dig_smem_off;           // byte offset of the accumulators behind the operand cache
};

// Slot planes are private to a lane and far larger than any cache, but their accesses still go through
// L1 (plain ld/st.global): routing them around it with ld/st.global.cg cost 17 % of the throughput
// (35.5 k instead of 42.6 k witnesses/s, profiles/README.md) - L1 merges the sectors of neighbouring slots.
#define PLD(ptr) (*(ptr))
#define PST(ptr, v) (*(ptr) = (v))
#define LDU(slot) PLD(Ul + (u64)(slot) * L)
#define STU(slot, v) PST(Ul + (u64)(slot) * L, (v))
// operand words: bit 31 -> shared-memory cell, else global slot (see pzk_program.h).
// Cells are addressed with explicit 32-bit shared-space addresses: `cells` is the lane's base
// (shared window offset + 8 * tid); cell c lives at cells + c * 8 * 128.  (x << 10) turns an operand
// word or a term ref into that byte offset: the flag bits above bit 21 fall off the top.
__device__ __forceinline__ u64 lds64(u32 addr) {
  u64 v;
  asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts64(u32 addr, u64 v) { asm volatile("st.shared.u64 [%0], %1;" ::"r"(addr), "l"(v) : "memory"); }
#define CELL_ADDR(x) (cells + ((u32)(x) << 10))
#define LDO(x) (((x) & PZK_OPERAND_CELL) ? lds64(CELL_ADDR(x)) : PLD(Ul + (u64)(x) * L))
// destination words: global slot always, cell when assigned
#define STD(d, v)                                                        \
  do {                                                                   \
    const u64 v__ = (v);                                                 \
    u32 d__ = (d);                                                       \
    asm volatile("" : "+r"(d__));                                        \
    rv0 = v__;                                                           \
    if (!(d__ & PZK_DST_OPTIONAL) || store_all) PST(Ul + (u64)PZK_DST_SLOT(d__) * L, v__); \
    if (PZK_DST_CELL(d__)) sts64(cells + ((PZK_DST_CELL(d__) - 1) << 10), v__); \
  } while (0)

__device__ __forceinline__ void ldF(const u64* Fl, u64 L, u32 slot, u64* v) {
  const u64* p = Fl + (u64)slot * 4 * L;
  v[0] = PLD(p); v[1] = PLD(p + L); v[2] = PLD(p + 2 * L); v[3] = PLD(p + 3 * L);
}
__device__ __forceinline__ void stF(u64* Fl, u64 L, u32 slot, const u64* v) {
  u64* p = Fl + (u64)slot * 4 * L;
  PST(p, v[0]); PST(p + L, v[1]); PST(p + 2 * L, v[2]); PST(p + 3 * L, v[3]);
}
__device__ __forceinline__ void ldFo(const u64* Fl, u64 L, u32 cells, u32 NT, u32 x, u64* v) {
  (void)NT;
  if (x & PZK_OPERAND_CELL) {
    const u32 c = cells + (x << 10);
    v[0] = lds64(c); v[1] = lds64(c + 1024); v[2] = lds64(c + 2048); v[3] = lds64(c + 3072);
  } else ldF(Fl, L, x, v);
}
__device__ __forceinline__ void stFd(u64* Fl, u64 L, u32 cells, u32 NT, u32 d, const u64* v, bool store_all) {
  (void)NT;
  if (!(d & PZK_DST_OPTIONAL) || store_all) stF(Fl, L, PZK_DST_SLOT(d), v);
  if (PZK_DST_CELL(d)) {
    const u32 c = cells + ((PZK_DST_CELL(d) - 1) << 10);
    sts64(c, v[0]); sts64(c + 1024, v[1]); sts64(c + 2048, v[2]); sts64(c + 3072, v[3]);
  }
}
__device__ __forceinline__ void ldPool(const u64* pool, u32 idx, u64* v) {
  const ulonglong2* p = reinterpret_cast<const ulonglong2*>(pool + 4 * (u64)idx);
  ulonglong2 lo = __ldg(p), hi = __ldg(p + 1);
  v[0] = lo.x; v[1] = lo.y; v[2] = hi.x; v[3] = hi.y;
}

// ---- long_div intrinsic: Knuth algorithm D, 64-bit digits ---------------------------------
// a: k+m digits, b: k digits (b[k-1] != 0, k >= 2); q: m+1 digits, r: k digits.
// Mathematically the unique quotient / remainder that
// /root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom:190-232 (long_div) produces.
// mod_inv intrinsic (PZK_MODINV): out = (a mod p)^-1 mod p for an odd prime p < 2^256 (0 when p | a),
// the value /root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom:430-465 reaches as a^(p-2).
// Binary extended GCD with the invariants x1 * a == u, x2 * a == v (mod p); p may exceed 2^255, so the
// halving step keeps the carry of x + p.
__device__ __noinline__ void modinv_device(const u32* Lst, u64* Ul, u64 L) {
  const unsigned n = Lst[0], k = Lst[1];  // k limbs of n bits (n = 64, or 32 for the 7-chunk secp224r1 verifier), k * n <= 256
  const u32* ia = Lst + 3;
  const u32* ip = ia + k;
  const u32* oo = ip + k + 1;
  u64 a[4] = {0, 0, 0, 0}, p[4] = {0, 0, 0, 0};
  for (unsigned i = 0; i < k; i++) {
    const unsigned pos = i * n;
    if (pos >= 256) break;
    const u64 av = LDU(ia[i]), pv = LDU(ip[i]);
    a[pos >> 6] |= av << (pos & 63);
    p[pos >> 6] |= pv << (pos & 63);
  }
  STU(ip[k], 0);
  u64 u[4], v[4] = {p[0], p[1], p[2], p[3]}, x1[4] = {1, 0, 0, 0}, x2[4] = {0, 0, 0, 0};
  // u = a mod p, bit-serial with the carry of the doubling kept (rr < p <= 2^256 - 1)
  {
    u64 rr[4] = {0, 0, 0, 0};
    for (int i = 255; i >= 0; i--) {
      const u32 c = add256(rr, rr, rr);
      rr[0] |= (a[i >> 6] >> (i & 63)) & 1;
      if (c || geq256(rr, p)) sub256(rr, rr, p);
    }
    u[0] = rr[0]; u[1] = rr[1]; u[2] = rr[2]; u[3] = rr[3];
  }
  u64 r[4] = {0, 0, 0, 0};
  if ((u[0] | u[1] | u[2] | u[3]) != 0 && (p[0] & 1)) {
    auto half = [&](u64* x) {
      u32 c = 0;
      if (x[0] & 1) c = add256(x, x, p);
      shr1_256(x);
      x[3] |= (u64)c << 63;
    };
    auto submod = [&](u64* x, const u64* y) {
      if (sub256(x, x, y)) add256(x, x, p);
    };
    auto is_one = [](const u64* x) { return x[0] == 1 && (x[1] | x[2] | x[3]) == 0; };
    unsigned guard = 0;
    while (!is_one(u) && !is_one(v) && guard++ < 1100) {
      while (!(u[0] & 1)) { shr1_256(u); half(x1); }
      while (!(v[0] & 1)) { shr1_256(v); half(x2); }
      if (geq256(u, v)) { sub256(u, u, v); submod(x1, x2); }
      else { sub256(v, v, u); submod(x2, x1); }
      if ((u[0] | u[1] | u[2] | u[3]) == 0 || (v[0] | v[1] | v[2] | v[3]) == 0) break;  // gcd > 1: p not prime
    }
    const bool use1 = is_one(u);
#pragma unroll
    for (int i = 0; i < 4; i++) r[i] = use1 ? x1[i] : x2[i];
    if (!is_one(u) && !is_one(v)) r[0] = r[1] = r[2] = r[3] = 0;
  }
  const u64 lmask = n >= 64 ? ~0ull : ((1ull << n) - 1);
  for (unsigned i = 0; i < k; i++) {
    const unsigned pos = i * n;
    STU(oo[i], pos < 256 ? ((r[pos >> 6] >> (pos & 63)) & lmask) : 0);
  }
}

__device__ __noinline__ u32 bigdiv_device(const u32* Lst, u64* Ul, u64 L) {
  // limbs of n = 64 bits are the digits of algorithm D; limbs of n = 32 bits (the secp224r1 verifier) are packed
  // in pairs: kl / ml are the caller's limb counts, k / m the digit counts
  const unsigned n = Lst[0], kl = Lst[1], ml = Lst[2];
  const bool half = n == 32;
  const unsigned k = half ? (kl + 1) / 2 : kl, na_l = kl + ml, n_a = half ? (na_l + 1) / 2 : na_l, m = n_a - k;
  const u32* ia = Lst + 3;
  const u32* ib = ia + na_l;
  const u32* oq = ib + kl;
  const u32* orr = oq + (ml + 1);
  auto dig_a = [&](unsigned i) -> u64 {
    if (!half) return LDU(ia[i]);
    return LDU(ia[2 * i]) | (2 * i + 1 < na_l ? LDU(ia[2 * i + 1]) << 32 : 0ull);
  };
  auto dig_b = [&](unsigned i) -> u64 {
    if (!half) return LDU(ib[i]);
    return LDU(ib[2 * i]) | (2 * i + 1 < kl ? LDU(ib[2 * i + 1]) << 32 : 0ull);
  };
  u64 un[132], vn[66];
  u32 st = 0;
  const u64 top = (kl >= 2 && k >= 2) ? dig_b(k - 1) : 0;
  if (kl < 2 || k < 2 || k > 64 || n_a > 128 || LDU(ib[kl - 1]) == 0 || top == 0) {
    for (unsigned i = 0; i <= ml; i++) STU(oq[i], 0);
    for (unsigned i = 0; i < kl; i++) STU(orr[i], 0);
    return PZK_LANE_BIGDIV_PRE;
  }
  int s = __clzll(top);
  // normalise
  for (int i = (int)k - 1; i > 0; i--) {
    u64 hi = dig_b(i), lo = dig_b(i - 1);
    vn[i] = s ? ((hi << s) | (lo >> (64 - s))) : hi;
  }
  vn[0] = dig_b(0) << s;
  un[n_a] = s ? (dig_a(n_a - 1) >> (64 - s)) : 0;
  for (int i = (int)n_a - 1; i > 0; i--) {
    u64 hi = dig_a(i), lo = dig_a(i - 1);
    un[i] = s ? ((hi << s) | (lo >> (64 - s))) : hi;
  }
  un[0] = dig_a(0) << s;
  for (int j = (int)m; j >= 0; j--) {
    // estimate the quotient digit from the top two digits
    u64 u2 = un[j + k], u1 = un[j + k - 1], u0 = un[j + k - 2];
    u64 qhat, rhat;
    bool rhat_over = false;
    if (u2 >= vn[k - 1]) {  // quotient digit would overflow: clamp to B-1
      qhat = ~0ull;
      rhat = u1 + vn[k - 1];
      rhat_over = rhat < u1;
      // u2 == vn[k-1] in valid inputs (u2 > vn[k-1] cannot happen when a < B^(m+1) * b)
      if (u2 > vn[k - 1]) st |= PZK_LANE_BIGDIV_PRE;
    } else {
      qhat = div128by64(u2, u1, vn[k - 1], &rhat);
    }
    while (!rhat_over) {
      u64 plo = qhat * vn[k - 2], phi = __umul64hi(qhat, vn[k - 2]);
      if (phi > rhat || (phi == rhat && plo > u0)) {
        qhat--;
        u64 nr = rhat + vn[k - 1];
        rhat_over = nr < rhat;
        rhat = nr;
      } else break;
    }
    // multiply and subtract
    u64 borrow = 0, carry = 0;
    for (unsigned i = 0; i < k; i++) {
      u64 plo = qhat * vn[i], phi = __umul64hi(qhat, vn[i]);
      plo += carry; phi += (plo < carry);
      carry = phi;
      u64 t = un[i + j] - plo;
      u64 b1 = un[i + j] < plo;
      u64 t2 = t - borrow;
      u64 b2 = t < borrow;
      un[i + j] = t2;
      borrow = b1 + b2;
    }
    {
      u64 t = un[j + k] - carry;
      u64 b1 = un[j + k] < carry;
      u64 t2 = t - borrow;
      u64 b2 = t < borrow;
      un[j + k] = t2;
      borrow = b1 + b2;
    }
    if (borrow) {  // add back
      qhat--;
      u64 c = 0;
      for (unsigned i = 0; i < k; i++) {
        u64 t = un[i + j] + vn[i];
        u64 c1 = t < vn[i];
        u64 t2 = t + c;
        u64 c2 = t2 < c;
        un[i + j] = t2;
        c = c1 + c2;
      }
      un[j + k] += c;
    }
    if (!half) STU(oq[j], qhat);
    else {
      // quotient limbs beyond m are zero whenever the circom function's precondition a < 2^(n (m + 1)) b holds
      if (2u * j <= ml) STU(oq[2 * j], qhat & 0xffffffffull); else if (qhat & 0xffffffffull) st |= PZK_LANE_BIGDIV_PRE;
      if (2u * j + 1 <= ml) STU(oq[2 * j + 1], qhat >> 32); else if (qhat >> 32) st |= PZK_LANE_BIGDIV_PRE;
    }
  }
  for (unsigned i = 0; i < k; i++) {
    u64 v = s ? ((un[i] >> s) | (un[i + 1] << (64 - s))) : un[i];
    if (!half) STU(orr[i], v);
    else {
      STU(orr[2 * i], v & 0xffffffffull);
      if (2 * i + 1 < kl) STU(orr[2 * i + 1], v >> 32);
    }
  }
  return st;
}


// ---- BabyJubjub base-8 ladder hints (PZK_BJJ_MUL8) ---------------------------------------------------
// The outputs of the 2n - 1 BabyjubjubAdd instances of BabyjubjubBase8Multiplication
// (/root/reference/circuits/lib/circuits/babyjubjub/curve.circom:143-171; adder formula :71-105, the (0,0)
// "no point" convention and the selection of addZeroBabyjub :19-58), evaluated in projective coordinates
// (x = X/Z, y = Y/Z; unified twisted-Edwards addition, 13 products) and normalised with ONE inversion for all
// 507 denominators (Montgomery's trick over the F plane).  Z3 = (A^2 - dCD)(A^2 + dCD) is never zero for points
// of the curve or for (0,0); should it be, the lane is flagged and its rows fail.
__device__ __noinline__ void fr_mul_call(u64* r, const u64* a, const u64* b) { fr_mul(r, a, b); }
struct BjjPt { u64 X[4], Y[4], Z[4]; };
__device__ __noinline__ void bjj_padd(BjjPt& o, const BjjPt& p, const BjjPt& q, const u64* ca, const u64* cd) {
  u64 A[4], B[4], C[4], D[4], E[4], F[4], G[4], t[4], u[4];
  fr_mul_call(A, p.Z, q.Z); fr_mul_call(B, A, A);
  fr_mul_call(C, p.X, q.X); fr_mul_call(D, p.Y, q.Y);
  fr_mul_call(E, C, D); fr_mul_call(E, E, cd);
  fr_sub(F, B, E); fr_add(G, B, E);
  fr_add(t, p.X, p.Y); fr_add(u, q.X, q.Y); fr_mul_call(t, t, u); fr_sub(t, t, C); fr_sub(t, t, D);
  fr_mul_call(u, A, F); fr_mul_call(o.X, u, t);
  fr_mul_call(t, C, ca); fr_sub(t, D, t);
  fr_mul_call(u, A, G); fr_mul_call(o.Y, u, t);
  fr_mul_call(o.Z, F, G);
}
__device__ __noinline__ u32 bjj_mul8_device(const u32* Lst, const u64* fpool, u64* Fl, u64 L) {
  const u32 n = Lst[0], nadd = 2 * n - 1;
  u64 ca[4], cd[4], bx[4], by[4], one[4], sc[4];
  ldPool(fpool, Lst[1], ca); ldPool(fpool, Lst[2], cd); ldPool(fpool, Lst[3], bx); ldPool(fpool, Lst[4], by);
  { const u64 o1[4] = {1, 0, 0, 0}; fr_to_mont(one, o1); }
  { u64 m[4]; ldF(Fl, L, Lst[5], m); fr_from_mont(sc, m); }
  const u32* out = Lst + 6;
  const u32* scr = out + 2 * nadd;
  BjjPt S, Dd, A, Q;
#pragma unroll
  for (int i = 0; i < 4; i++) { S.X[i] = S.Y[i] = 0; S.Z[i] = one[i]; }
  u64 pre[4] = {one[0], one[1], one[2], one[3]};
  u32 k = 0;
  for (u32 i = 0; i < n; i++) {
    if (i > 0) {
      bjj_padd(Dd, S, S, ca, cd);
      stF(Fl, L, out[2 * k], Dd.X); stF(Fl, L, out[2 * k + 1], Dd.Y); stF(Fl, L, scr[2 * k], Dd.Z);
      fr_mul_call(pre, pre, Dd.Z); stF(Fl, L, scr[2 * k + 1], pre);
      k++;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) { Dd.X[j] = Dd.Y[j] = 0; Dd.Z[j] = one[j]; }
    }
    const u32 bi = n - 1 - i;
    const bool bit = (sc[bi >> 6] >> (bi & 63)) & 1;
#pragma unroll
    for (int j = 0; j < 4; j++) { Q.X[j] = bit ? bx[j] : 0; Q.Y[j] = bit ? by[j] : 0; Q.Z[j] = one[j]; }
    bjj_padd(A, Dd, Q, ca, cd);
    stF(Fl, L, out[2 * k], A.X); stF(Fl, L, out[2 * k + 1], A.Y); stF(Fl, L, scr[2 * k], A.Z);
    fr_mul_call(pre, pre, A.Z); stF(Fl, L, scr[2 * k + 1], pre);
    k++;
    // addZeroBabyjub: in1 "zero" -> in2, else in2 "zero" -> in1, else the sum (zero = x coordinate 0)
    const bool zd = fr_is_zero(Dd.X), zq = !bit;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      S.X[j] = zd ? Q.X[j] : (zq ? Dd.X[j] : A.X[j]);
      S.Y[j] = zd ? Q.Y[j] : (zq ? Dd.Y[j] : A.Y[j]);
      S.Z[j] = zd ? Q.Z[j] : (zq ? Dd.Z[j] : A.Z[j]);
    }
  }
  u32 st = fr_is_zero(pre) ? PZK_LANE_HINT : 0;
  u64 inv[4];
  fr_inv(inv, pre);
  for (u32 kk = nadd; kk-- > 0;) {
    u64 z[4], zi[4], x[4], y[4];
    ldF(Fl, L, scr[2 * kk], z);
    if (kk > 0) { u64 pp[4]; ldF(Fl, L, scr[2 * kk - 1], pp); fr_mul_call(zi, inv, pp); }
    else { zi[0] = inv[0]; zi[1] = inv[1]; zi[2] = inv[2]; zi[3] = inv[3]; }
    fr_mul_call(inv, inv, z);
    ldF(Fl, L, out[2 * kk], x); ldF(Fl, L, out[2 * kk + 1], y);
    fr_mul_call(x, x, zi); fr_mul_call(y, y, zi);
    stF(Fl, L, out[2 * kk], x); stF(Fl, L, out[2 * kk + 1], y);
  }
  return st;
}

// ---- rows fused into the op stream ---------------------------------------------------------
// term k of a row lives in record (k >> 1), words (2*(k&1), 2*(k&1)+1)
__device__ __forceinline__ uint2 row_term(const uint4* recs, u32 k) {
  const uint4 w = __ldg(recs + (k >> 1));
  return (k & 1) ? make_uint2(w.z, w.w) : make_uint2(w.x, w.y);
}
__device__ __forceinline__ long long term_icoef(const u32* list, u32 ref, u32 cw) {
  if (ref == PZK_REF_ONE_LIST || (ref < PZK_REF_ONE_LIST && (ref & PZK_TERM_COEF_LIST)))
    return (long long)((u64)__ldg(list + cw) | ((u64)__ldg(list + cw + 1) << 32));
  return (long long)(int)cw;
}
// exact integer row: |A|,|B| < 2^63 and |C| < 2^126 proven by the compiler
#define TERM_U(ref) (((ref) & PZK_TERM_CELL) ? lds64(cells + ((u32)(ref) << 10)) : PLD(Ul + (u64)PZK_REF_SLOT(ref) * L))
__device__ __forceinline__ bool check_row_int(const uint4* recs, u32 na, u32 nb, u32 nc, const u32* list,
                                              const u64* Ul, u64 L, u32 cells, u32 NT) {
  long long A = 0, B = 0;
  u64 Clo = 0, Chi = 0;
  u32 k = 0;
  for (u32 i = 0; i < na; i++, k++) {
    const uint2 t = row_term(recs, k);
    long long c = term_icoef(list, t.x, t.y);
    long long v = (t.x >= PZK_REF_ONE_LIST) ? 1 : (long long)TERM_U(t.x);
    A += c * v;
  }
  for (u32 i = 0; i < nb; i++, k++) {
    const uint2 t = row_term(recs, k);
    long long c = term_icoef(list, t.x, t.y);
    long long v = (t.x >= PZK_REF_ONE_LIST) ? 1 : (long long)TERM_U(t.x);
    B += c * v;
  }
  for (u32 i = 0; i < nc; i++, k++) {
    const uint2 t = row_term(recs, k);
    long long c = term_icoef(list, t.x, t.y);
    bool one = t.x >= PZK_REF_ONE_LIST;
    u64 v = one ? 1ull : TERM_U(t.x);
    bool vsigned = !one && PZK_REF_CLS(t.x) == 1;
    // 128-bit two's complement product c * v
    u64 lo = (u64)c * v;
    u64 hi = __umul64hi((u64)c, v);
    if (c < 0) hi -= v;
    if (vsigned && (long long)v < 0) hi -= (u64)c;
    u64 nlo = Clo + lo;
    Chi += hi + (nlo < lo);
    Clo = nlo;
  }
  if (na == 0 || nb == 0) return (Clo | Chi) == 0;
  u64 plo = (u64)A * (u64)B;
  u64 phi = __umul64hi((u64)A, (u64)B);
  if (A < 0) phi -= (u64)B;
  if (B < 0) phi -= (u64)A;
  return plo == Clo && phi == Chi;
}

// One linear combination of a CHECK_F row accumulated directly in Montgomery form.
//   F wire, coefficient +-1      : modular add / sub
//   F wire, other coefficient    : one Montgomery product
//   proven bit (PZK_TERM_BIT)    : conditional add of the coefficient
//   other narrow wire            : |v| * (c R^2) -> one Montgomery product, sign applied after
__device__ __forceinline__ void lin_field(const StreamCoefs sc, const uint2* terms, u32 k0, u32 k1, const u64* Ul,
                                          const u64* Fl, u64 L, u32 cells, u64* acc) {
  acc[0] = acc[1] = acc[2] = acc[3] = 0;
  for (u32 k = k0; k < k1; k++) {
    const uint2 tw = __ldg(terms + k);
    const u32 ref = tw.x, ci = tw.y;
    const u64* cbase = reinterpret_cast<const u64*>(sc.coefs);
    if (ref == PZK_REF_ONE) { u64 c[4]; ldPool(cbase, ci * 3 + 1, c); fr_add(acc, acc, c); continue; }
    if (ref & PZK_TERM_BIT) {
      if (TERM_U(ref) & 1) { u64 c[4]; ldPool(cbase, ci * 3 + 1, c); fr_add(acc, acc, c); }
      continue;
    }
    const u32 cls = PZK_REF_CLS(ref);
    if (cls == 3) {
      // Z wire: |v| * (c R^2) -> one Montgomery product, sign applied after
      u64 z[4], m[4], c[4], r[4];
      ldFo(Fl, L, cells, 0, (ref & PZK_TERM_CELL) ? (PZK_OPERAND_CELL | (ref & 0xffffu)) : PZK_REF_SLOT(ref), z);
      const bool vneg = (long long)z[3] < 0;
      if (vneg) { const u64 zero[4] = {0, 0, 0, 0}; sub256(m, zero, z); } else { m[0] = z[0]; m[1] = z[1]; m[2] = z[2]; m[3] = z[3]; }
      ldPool(cbase, ci * 3 + 2, c);  // c * R^2
      fr_mul(r, c, m);
      if (vneg) fr_sub(acc, acc, r); else fr_add(acc, acc, r);
      continue;
    }
    if (cls < 2) {
      u64 v = TERM_U(ref);
      const bool vneg = (cls == 1) && ((long long)v < 0);
      if (vneg) v = (u64)(-(long long)v);
      u64 c[4], w[4] = {v, 0, 0, 0}, r[4];
      ldPool(cbase, ci * 3 + 2, c);  // c * R^2
      fr_mul(r, c, w);
      if (vneg) fr_sub(acc, acc, r); else fr_add(acc, acc, r);
    } else {
      u64 w[4];
      ldFo(Fl, L, cells, 0, (ref & PZK_TERM_CELL) ? (PZK_OPERAND_CELL | (ref & 0xffffu)) : PZK_REF_SLOT(ref), w);
      const u32 kind = __ldg(sc.coef_kind + ci);
      if (kind && __ldg(sc.coef_mag + ci) == 1) {
        if (kind == 1) fr_add(acc, acc, w); else fr_sub(acc, acc, w);
      } else {
        u64 c[4], r[4];
        ldPool(cbase, ci * 3 + 1, c);  // c * R
        fr_mul(r, c, w);
        fr_add(acc, acc, r);
      }
    }
  }
}

__device__ __noinline__ bool check_row_field(const StreamCoefs sc, const uint4* recs, u32 na, u32 nb, u32 nc,
                                             const u64* Ul, const u64* Fl, u64 L, u32 cells, u32 NT) {
  (void)NT;
  const uint2* terms = reinterpret_cast<const uint2*>(recs);
  u64 c[4];
  lin_field(sc, terms, na + nb, na + nb + nc, Ul, Fl, L, cells, c);
  if (na == 0 || nb == 0) return fr_is_zero(c);
  // narrow x narrow with unit coefficients (the limb products of the big-integer multipliers):
  // exact 128-bit integer product and ONE conversion instead of three field products
  if (na == 1 && nb == 1) {
    const uint2 ta = __ldg(terms), tb = __ldg(terms + 1);
    if (ta.x < PZK_REF_ONE_LIST && tb.x < PZK_REF_ONE_LIST && PZK_REF_CLS(ta.x) == 0 && PZK_REF_CLS(tb.x) == 0 &&
        !((ta.x | tb.x) & PZK_TERM_BIT)) {
      const u32 ka = __ldg(sc.coef_kind + ta.y), kb = __ldg(sc.coef_kind + tb.y);
      if (ka && kb && __ldg(sc.coef_mag + ta.y) == 1 && __ldg(sc.coef_mag + tb.y) == 1) {
        const u64 va = TERM_U(ta.x), vb = TERM_U(tb.x);
        u64 w[4] = {va * vb, __umul64hi(va, vb), 0, 0}, pf[4];
        fr_to_mont(pf, w);
        if ((ka == 2) != (kb == 2)) fr_neg(pf, pf);
        return fr_eq(pf, c);
      }
    }
  }
  u64 a[4], b[4], ab[4];
  lin_field(sc, terms, 0, na, Ul, Fl, L, cells, a);
  lin_field(sc, terms, na, na + nb, Ul, Fl, L, cells, b);
  fr_mul(ab, a, b);
  return fr_eq(ab, c);
}


// ---- packed truth table (PZK_V_LUT): one record computes 32 / 64 one-bit signals --------------------
// r bit l = T >> (x0_l | x1_l << 1 | x2_l << 2 | x3_l << 3) & 1.  T is warp-uniform, so the masks of the
// multiplexer tree are uniform values; level 1 folds x0 into the 8 pairs of table bits, levels 2..4 are
// one LOP3 (bit select) each.  XOR3 / XOR2 / AND2 - the sigma and carry-free sums of SHA - leave early.
template <typename W>
__device__ __forceinline__ W vlut_eval(u32 T, W x0, W x1, W x2, W x3) {
  if (T == 0x9696u) return x0 ^ x1 ^ x2;
  if (T == 0x6666u) return x0 ^ x1;
  if (T == 0x8888u) return x0 & x1;
  W g[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    const W m0 = (W)0 - (W)((T >> (2 * k)) & 1u), m1 = (W)0 - (W)((T >> (2 * k + 1)) & 1u);
    g[k] = m0 ^ (x0 & (m0 ^ m1));
  }
#pragma unroll
  for (int k = 0; k < 4; k++) g[k] = g[2 * k] ^ (x1 & (g[2 * k] ^ g[2 * k + 1]));
#pragma unroll
  for (int k = 0; k < 2; k++) g[k] = g[2 * k] ^ (x2 & (g[2 * k] ^ g[2 * k + 1]));
  return g[0] ^ (x3 & (g[0] ^ g[1]));
}
__device__ __forceinline__ u32 rotr32(u32 v, u32 r) { return __funnelshift_r(v, v, r); }
__device__ __forceinline__ u64 rotr64(u64 v, u32 r) { return r ? ((v >> r) | (v << (64 - r))) : v; }

// bit field of a plain 256-bit value / a word: ((v >> s) & (2^n - 1)) << k  (views, pzk_program.h)
__device__ __forceinline__ void field256(u64* r, const u64* v, unsigned s, unsigned n, unsigned k) {
  u64 t[4];
  shr256(t, v, s);
  if (n < 256) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int lo = 64 * i;
      const u64 m = (int)n >= lo + 64 ? ~0ull : ((int)n <= lo ? 0ull : ((1ull << (n - lo)) - 1));
      t[i] &= m;
    }
  }
  shl256(r, t, k);
}

// ---- fused witness digest --------------------------------------------------------------------------------
// The digest of pzk.h (sum over wires of c(i) w_i mod p) folded WHEN A VALUE IS DEFINED instead of after the
// segment from HBM (digest_kernel below keeps the entries that cannot be attached to one defining op: outputs of
// the hint intrinsics, truth-table views over several words, views wider than 64 bits).  An op with PZK_FLAG_DIG is
// followed by a descriptor the runtime filled in when it loaded the program:
//   x = rep | has_table << 4 | nbits << 8    rep: 1 = U word, 2 = signed I word, 3 = F (Montgomery), 4 = Z, 5 = N (plain)
//   y, z = 64-bit weight of the value itself (sum of c(i) over the wires that ARE this value; 0 = none)
//   w = offset of the per-bit coefficient table (all bit-field views of this word)
// Accumulators live in shared memory behind the operand cache (12 x 8 bytes per lane): accN 128-bit (narrow words,
// view tables), accM 320-bit (sum of weight x Montgomery value), accP 320-bit (sum of weight x canonical value);
// they are flushed to the per-lane carry-save state at the end of the launch, digest_finalize_kernel normalises.
#define DIG_ACC_WORDS 12
__device__ __forceinline__ void dig_acc_add128(u32 acc, u64 lo, u64 hi) {  // acc: shared address of {lo, hi}
  u64 a0 = lds64(acc), a1 = lds64(acc + 1024);
  asm("add.cc.u64 %0, %0, %2; addc.u64 %1, %1, %3;" : "+l"(a0), "+l"(a1) : "l"(lo), "l"(hi));
  sts64(acc, a0); sts64(acc + 1024, a1);
}
// a (10 x 32 bits) += c * v (8 x 32 bits): one mad.lo carry chain over the limbs, one mad.hi chain a limb higher
__device__ __forceinline__ void mad320_chain(u32* a, u32 c32, const u32* v) {
  asm("mad.lo.cc.u32 %0, %10, %11, %0;\n\t"
        "madc.lo.cc.u32 %1, %10, %12, %1;\n\t"
        "madc.lo.cc.u32 %2, %10, %13, %2;\n\t"
        "madc.lo.cc.u32 %3, %10, %14, %3;\n\t"
        "madc.lo.cc.u32 %4, %10, %15, %4;\n\t"
        "madc.lo.cc.u32 %5, %10, %16, %5;\n\t"
        "madc.lo.cc.u32 %6, %10, %17, %6;\n\t"
        "madc.lo.cc.u32 %7, %10, %18, %7;\n\t"
        "addc.cc.u32 %8, %8, 0;\n\t"
        "addc.u32 %9, %9, 0;\n\t"
        "mad.hi.cc.u32 %1, %10, %11, %1;\n\t"
        "madc.hi.cc.u32 %2, %10, %12, %2;\n\t"
        "madc.hi.cc.u32 %3, %10, %13, %3;\n\t"
        "madc.hi.cc.u32 %4, %10, %14, %4;\n\t"
        "madc.hi.cc.u32 %5, %10, %15, %5;\n\t"
        "madc.hi.cc.u32 %6, %10, %16, %6;\n\t"
        "madc.hi.cc.u32 %7, %10, %17, %7;\n\t"
        "madc.hi.cc.u32 %8, %10, %18, %8;\n\t"
        "addc.u32 %9, %9, 0;"
        : "+r"(a[0]), "+r"(a[1]), "+r"(a[2]), "+r"(a[3]), "+r"(a[4]), "+r"(a[5]), "+r"(a[6]), "+r"(a[7]), "+r"(a[8]), "+r"(a[9])
        : "r"(c32), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]));
}
// acc (5 x 64 in shared memory) += c * w (4 x 64), c < 2^64, total < 2^320.  Almost every weight is the 32-bit c(i)
// of a single wire: 32 x 32 products on the halves of the limbs (8 IMAD.WIDE) instead of 64 x 64 ones.
__device__ __noinline__ void dig_acc_mac320(u32 acc, u64 c, u64 w0, u64 w1, u64 w2, u64 w3) {
  const u64 w[4] = {w0, w1, w2, w3};
  u64 carry = 0;
  if ((c >> 32) == 0) {
    // 32-bit weight: one mad.lo carry chain over the eight 32-bit limbs, one mad.hi chain a limb higher
    const u32 c32 = (u32)c;
    u32 a[10], v[8];
#pragma unroll
    for (int j = 0; j < 5; j++) { const u64 t = lds64(acc + 1024 * j); a[2 * j] = (u32)t; a[2 * j + 1] = (u32)(t >> 32); }
#pragma unroll
    for (int j = 0; j < 4; j++) { v[2 * j] = (u32)w[j]; v[2 * j + 1] = (u32)(w[j] >> 32); }
    mad320_chain(a, c32, v);
#pragma unroll
    for (int j = 0; j < 5; j++) sts64(acc + 1024 * j, (u64)a[2 * j] | ((u64)a[2 * j + 1] << 32));
    return;
  } else {
#pragma unroll
    for (int j = 0; j < 4; j++) {
      const u64 lo = c * w[j], hi = __umul64hi(c, w[j]);
      const u64 a = lds64(acc + 1024 * j);
      u64 t = a + lo;
      const u64 c1 = t < lo;
      const u64 t2 = t + carry;
      const u64 c2 = t2 < carry;
      sts64(acc + 1024 * j, t2);
      carry = hi + c1 + c2;  // hi <= 2^64 - 2: no overflow
    }
  }
  sts64(acc + 4096, lds64(acc + 4096) + carry);
}
// Bit-field views of a word, folded four bits at a time: the table of a word holds, per nibble, the 16 sums of
// the per-bit weights (pzk_api.cu, build_digest_program), so a nibble costs one indexed 16-byte load and one 128-bit
// add instead of four loads and four predicated adds.  The lanes of a warp read up to 16 neighbouring entries (256 B).
__device__ __noinline__ void dig_fold_table(u32 accN, const ulonglong2* T, u32 nbits, u64 W0, u64 W1, u64 W2, u64 W3) {
  u64 a0 = lds64(accN), a1 = lds64(accN + 1024);
  const u32 ng = (nbits + 3) >> 2;  // nibbles with a table
#define DIG_NIBBLE(g)                                                                                        \
  do {                                                                                                       \
    const u32 off = (g) == 0 ? (w << 4) & 0xf0u : (g) == 1 ? w & 0xf0u : (w >> (4 * (g) - 4)) & 0xf0u;         \
    const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(Tb + off + 256 * (g)));                   \
    asm("add.cc.u64 %0, %0, %2; addc.u64 %1, %1, %3;" : "+l"(a0), "+l"(a1) : "l"(t.x), "l"(t.y));            \
  } while (0)
  const unsigned char* Tb = reinterpret_cast<const unsigned char*>(T);
  for (u32 h = 0; h * 8 < ng; h++, Tb += 8 * 256) {  // 32 bits of the word at a time
    const u64 W = h < 2 ? W0 : h < 4 ? W1 : h < 6 ? W2 : W3;
    const u32 w = (h & 1u) ? (u32)(W >> 32) : (u32)W;
    const u32 n = ng - h * 8;
    if (n >= 8) {
      DIG_NIBBLE(0); DIG_NIBBLE(1); DIG_NIBBLE(2); DIG_NIBBLE(3); DIG_NIBBLE(4); DIG_NIBBLE(5); DIG_NIBBLE(6); DIG_NIBBLE(7);
    } else {
      for (u32 g = 0; g < n; g++) {
        const ulonglong2 t = __ldg(reinterpret_cast<const ulonglong2*>(Tb + (((w >> (4 * g)) & 15u) << 4) + 256 * g));
        asm("add.cc.u64 %0, %0, %2; addc.u64 %1, %1, %3;" : "+l"(a0), "+l"(a1) : "l"(t.x), "l"(t.y));
      }
    }
  }
#undef DIG_NIBBLE
  sts64(accN, a0); sts64(accN + 1024, a1);
}

// one descriptor: v0..v3 = the value just defined (a U / I word in v0, or the 4 limbs of an F / Z / N value)
__device__ __noinline__ void dig_fold(u32 dacc, const ulonglong2* tab, uint4 dg, u64 v0, u64 v1, u64 v2, u64 v3) {
  const u32 repk = dg.x & 15u, nbits = dg.x >> 8;
  const u64 c = (u64)dg.y | ((u64)dg.z << 32);
  const u32 accN = dacc, accM = dacc + 2 * 1024, accP = dacc + 7 * 1024;
  if (repk == 1 || (repk == 2 && (long long)v0 >= 0)) {
    if (c) dig_acc_add128(accN, c * v0, __umul64hi(c, v0));
    if (dg.x & 16u) dig_fold_table(accN, tab + dg.w, nbits, v0, 0, 0, 0);
  } else if (repk == 2) {  // negative signed word: its canonical value is p - |v|
    const u64 pp[4] = {P0, P1, P2, P3}; u64 m[4] = {(u64)(-(long long)v0), 0, 0, 0}, w[4];
    sub256(w, pp, m);
    if (c) dig_acc_mac320(accP, c, w[0], w[1], w[2], w[3]);
  } else if (repk == 3) {
    if (c) dig_acc_mac320(accM, c, v0, v1, v2, v3);
  } else if (repk == 4) {
    if (c) {
      u64 w[4] = {v0, v1, v2, v3};
      if ((long long)v3 < 0) { const u64 pp[4] = {P0, P1, P2, P3}; add256(w, w, pp); }
      dig_acc_mac320(accP, c, w[0], w[1], w[2], w[3]);
    }
  } else if (repk == 5) {
    if (c) dig_acc_mac320(accP, c, v0, v1, v2, v3);
    if (dg.x & 16u) dig_fold_table(accN, tab + dg.w, nbits, v0, v1, v2, v3);
  }
}

// ---- Z class: exact wide integers in 256-bit two's complement (pzk_program.h) -----------------------------
// canonical residue of a Z value: v < 0 ? p + v : v
__device__ __forceinline__ void z_canonical(u64* r, const u64* z) {
  r[0] = z[0]; r[1] = z[1]; r[2] = z[2]; r[3] = z[3];
  if ((long long)z[3] < 0) { const u64 pp[4] = {P0, P1, P2, P3}; add256(r, r, pp); }
}
// Montgomery form of a Z value (|v| < 2^250 < p)
__device__ __forceinline__ void z_to_mont(u64* r, const u64* z) {
  const bool neg = (long long)z[3] < 0;
  u64 m[4] = {z[0], z[1], z[2], z[3]};
  if (neg) { const u64 zero[4] = {0, 0, 0, 0}; sub256(m, zero, z); }
  fr_to_mont(r, m);
  if (neg) fr_neg(r, r);
}
// a * b mod 2^256 on 64-bit limbs, everything in registers (no arrays whose address escapes: a __noinline__ helper with
// pointer arguments demoted the operand arrays of every Z op to local memory - 30 % of the stall samples of the
// big-integer segments, profiles/README.md).  LA / LB: 64-bit limbs the NON-NEGATIVE operands fit (0 = unknown or
// possibly negative: truncated 4 x 4, exact in two's complement whenever the true product fits).
#define Z_MAC(acc_lo, acc_hi, carry, x, y)                         \
  do {                                                             \
    const u64 pl__ = (x) * (y), ph__ = __umul64hi((x), (y));       \
    const u64 s__ = (acc_lo) + pl__;                               \
    const u64 c__ = s__ < pl__;                                    \
    (acc_lo) = s__;                                                \
    const u64 t__ = (acc_hi) + ph__;                               \
    const u64 c2__ = t__ < ph__;                                   \
    const u64 u__ = t__ + c__;                                     \
    (carry) += c2__ + (u__ < c__);                                 \
    (acc_hi) = u__;                                                \
  } while (0)
__device__ __forceinline__ void z_mul(u64& r0, u64& r1, u64& r2, u64& r3, u64 a0, u64 a1, u64 a2, u64 a3, u64 b0, u64 b1,
                                      u64 b2, u64 b3, u32 imm) {
  const u32 la = imm & 15u, lb = (imm >> 4) & 15u;
  if (la == 1 && lb == 1) {  // 64 x 64: the limb products of the multipliers
    r0 = a0 * b0; r1 = __umul64hi(a0, b0); r2 = 0; r3 = 0;
    return;
  }
  if (la != 0 && la <= 2 && lb <= 2) {  // (64 + carry bits) x (64 + carry bits): Karatsuba leaves
    u64 c0 = a0 * b0, c1 = __umul64hi(a0, b0), c2 = 0, c3 = 0, k = 0;
    Z_MAC(c1, c2, c3, a0, b1);
    Z_MAC(c1, c2, c3, a1, b0);
    Z_MAC(c2, c3, k, a1, b1);
    r0 = c0; r1 = c1; r2 = c2; r3 = c3;
    return;
  }
  // truncated 4 x 4
  u64 c0 = a0 * b0, c1 = __umul64hi(a0, b0), c2 = 0, c3 = 0, k = 0;
  Z_MAC(c1, c2, c3, a0, b1);
  Z_MAC(c1, c2, c3, a1, b0);
  Z_MAC(c2, c3, k, a0, b2);
  Z_MAC(c2, c3, k, a1, b1);
  Z_MAC(c2, c3, k, a2, b0);
  c3 += a0 * b3 + a1 * b2 + a2 * b1 + a3 * b0;
  r0 = c0; r1 = c1; r2 = c2; r3 = c3;
}

// 7 CTAs of 128 lanes per SM (72 registers): measured best of 8 / 7 / 6 / 5 on the round-2 program (284 / 298 / 295 /
// 280 k witnesses/s): one more warp-quartet of latency hiding is worth less than the spills of a 64-register budget
// the lane index, read from the special registers at each use: kept as a loop-invariant variable the compiler
// re-derives it at the top of every record
__device__ __forceinline__ u64 lane_now() {
  u32 c, n, t;
  asm volatile("mov.u32 %0, %%ctaid.x;" : "=r"(c));
  asm volatile("mov.u32 %0, %%ntid.x;" : "=r"(n));
  asm volatile("mov.u32 %0, %%tid.x;" : "=r"(t));
  return (u64)c * n + t;
}
__global__ void __launch_bounds__(128, 7) eval_kernel(EvalParams p) {
  extern __shared__ u64 cell_mem[];
  if ((u64)blockIdx.x * blockDim.x + threadIdx.x >= p.n_lanes) return;
  // Blocked planes: the 128 lanes of a CTA own one contiguous region per plane, so the hot slots of
  // a CTA span a few 2 MB pages instead of one page per slot (the flat [slot][lane] layout was
  // page-walk bound: consecutive ops touch slots that are megabytes apart).
  const u64 L = PZK_LANE_BLOCK;
  const u32 NT = blockDim.x;
  const u32 cells = (u32)__cvta_generic_to_shared(cell_mem) + threadIdx.x * 8;  // lane base in the shared window
  u64* Ul = p.U + (u64)blockIdx.x * p.n_u_slots * PZK_LANE_BLOCK + threadIdx.x;
  u64* Fl = p.F + (u64)blockIdx.x * p.n_f_slots * 4 * PZK_LANE_BLOCK + threadIdx.x;
  u32 st = 0;
  u32 bad = 0xffffffffu;  // first failing row (row ids are 32-bit record words)
  // kernel parameters are copied into registers once: taking references to the parameter block
  // would push it to local memory and turn every use into an LDL
  const bool store_all = p.store_all != 0;
  const bool check_rows = p.check_rows != 0;
  const uint4* __restrict__ ops = p.ops;
  const u64* __restrict__ fpool = p.fpool;
  const u32* __restrict__ list = p.list;
  const StreamCoefs sc = p.sc;
  const u32 n_rec = (u32)p.n_rec;
  const bool digest = p.digest != 0;
  const ulonglong2* __restrict__ dig_tab = p.dig_tab;
  const u32 dig_off = p.dig_smem_off;  // uniform: the accumulators sit behind the operand cache
#define dacc (cells + dig_off)
  if (digest) {
#pragma unroll
    for (int k = 0; k < DIG_ACC_WORDS; k++) sts64(dacc + 1024 * k, 0);
  }
  for (u32 pc = 0; pc < n_rec; pc++) {
    const uint4 w = __ldg(ops + pc);
    const u32 opc = w.x & 0xffu, flags = (w.x >> 8) & 0xffu, imm16 = w.x >> 16;
    const u32 dst = w.y, a = w.z, b = w.w;
    u64 rv0 = 0, rv1 = 0, rv2 = 0, rv3 = 0;  // the value this record defines, for the digest fold behind it
#define FETCH_EXT() const uint4 x = __ldg(ops + (++pc))
#define UBV ((flags & PZK_FLAG_B_IMM) ? (u64)b : LDO(b))
#define LDFA(v) ldFo(Fl, L, cells, NT, a, v)
#define LDFB(v) do { if (flags & PZK_FLAG_B_POOL) ldPool(fpool, b, v); else ldFo(Fl, L, cells, NT, b, v); } while (0)
// the product of a fused multiply-add that is itself a wire (PZK_FLAG_DST2): stored when witnesses are exported,
// folded into the digest through its own descriptor (in front of the sum's)
#define MULADD_PRODUCT(v)                                                                   \
  do {                                                                                      \
    if (!(x.y & PZK_DST_OPTIONAL) || store_all) stF(Fl, L, PZK_DST_SLOT(x.y), v);            \
    if (flags & PZK_FLAG_DIG2) {                                                            \
      const uint4 dg2 = __ldg(ops + (++pc));                                                \
      if (digest && (dg2.x & 15u)) dig_fold(dacc, dig_tab, dg2, (v)[0], (v)[1], (v)[2], (v)[3]); \
    }                                                                                       \
  } while (0)
#define STFD(v) do { stFd(Fl, L, cells, NT, dst, v, store_all); rv0 = (v)[0]; rv1 = (v)[1]; rv2 = (v)[2]; rv3 = (v)[3]; } while (0)
    // U_ADD / U_MUL / U_AND / U_SHR / U_SHLADD are 57 % of the ops of the passport circuits (weighted bit sums, bit
    // extraction).  The compiler marks them with PZK_FLAG_FAST; testing the flag instead of the opcode keeps this
    // exit in front of the six-level compare tree the switch compiles to, and the four results are selected,
    // not branched on.
    if (flags & PZK_FLAG_FAST) {
      const u64 x_ = LDO(a), y_ = UBV;
      const u64 sh_ = y_ >= 64 ? 0 : x_ >> y_;
      const u64 r_ = opc == PZK_U_ADD ? x_ + y_ : opc == PZK_U_MUL ? x_ * y_ : opc == PZK_U_AND ? (x_ & y_)
                   : opc == PZK_U_SHLADD ? x_ + (y_ << imm16) : sh_;
      STD(dst, r_);
      if (flags & PZK_FLAG_DIG) {
        const uint4 dg = __ldg(ops + (++pc));
        if (digest) dig_fold(dacc, dig_tab, dg, r_, 0, 0, 0);
      }
      continue;
    }
    // the two groups that make up two thirds of the remaining records leave before the compare tree of the switch
    if (opc - PZK_F_ADD <= 2u) {
      u64 va[4], vb[4], r[4];
      LDFA(va); LDFB(vb);
      if (opc == PZK_F_ADD) fr_add(r, va, vb);
      else if (opc == PZK_F_SUB) fr_sub(r, va, vb);
      else fr_mul(r, va, vb);
      STFD(r);
    } else if (opc - PZK_Z_ADD <= 2u) {
      u64 va[4], vb[4], r[4];
      // Z_MUL: a factor may be a U word read in place (PZK_FLAG_A_U / PZK_FLAG_B_U; never set on Z_ADD / Z_SUB)
      if (flags & PZK_FLAG_A_U) { va[0] = LDO(a); va[1] = va[2] = va[3] = 0; } else LDFA(va);
      if (flags & PZK_FLAG_B_U) { vb[0] = LDO(b); vb[1] = vb[2] = vb[3] = 0; } else LDFB(vb);
      if (opc == PZK_Z_ADD) add256(r, va, vb);
      else if (opc == PZK_Z_SUB) sub256(r, va, vb);
      else z_mul(r[0], r[1], r[2], r[3], va[0], va[1], va[2], va[3], vb[0], vb[1], vb[2], vb[3], imm16);
      STFD(r);
    } else switch (opc) {
      case PZK_NOP: break;
      case PZK_U_CONST: STD(dst, ((u64)b << 32) | a); break;
      case PZK_U_ADD: STD(dst, LDO(a) + UBV); break;
      case PZK_U_SHLADD: STD(dst, LDO(a) + (LDO(b) << imm16)); break;
      case PZK_U_SUB: STD(dst, LDO(a) - UBV); break;
      case PZK_U_MUL: STD(dst, LDO(a) * UBV); break;
      case PZK_U_DIV: { u64 d = UBV; STD(dst, d ? LDO(a) / d : 0); break; }
      case PZK_U_MOD: { u64 d = UBV; STD(dst, d ? LDO(a) % d : 0); break; }
      case PZK_U_SHR: { u64 d = UBV; STD(dst, d >= 64 ? 0 : LDO(a) >> d); break; }
      case PZK_U_SHL: { u64 d = UBV; STD(dst, d >= 64 ? 0 : LDO(a) << d); break; }
      case PZK_U_AND: STD(dst, LDO(a) & UBV); break;
      case PZK_U_OR: STD(dst, LDO(a) | UBV); break;
      case PZK_U_XOR: STD(dst, LDO(a) ^ UBV); break;
      case PZK_U_LT: STD(dst, (u64)(LDO(a) < UBV)); break;
      case PZK_U_LE: STD(dst, (u64)(LDO(a) <= UBV)); break;
      case PZK_U_EQ: STD(dst, (u64)(LDO(a) == UBV)); break;
      case PZK_U_NE: STD(dst, (u64)(LDO(a) != UBV)); break;
      case PZK_I_LT: STD(dst, (u64)((long long)LDO(a) < (long long)UBV)); break;
      case PZK_I_LE: STD(dst, (u64)((long long)LDO(a) <= (long long)UBV)); break;
      case PZK_U_SEL: { FETCH_EXT(); STD(dst, LDO(a) ? LDO(b) : LDO(x.x)); break; }
      case PZK_U_LUT: case PZK_U_LUTV: {
        FETCH_EXT();
        // operand j contributes bit (x.w >> 8j) & 255 of its word (0 for plain one-bit values)
        u32 idx = 0;
        if (a != PZK_OPERAND_NONE) idx |= (u32)((LDO(a) >> (x.w & 255u)) & 1);
        if (b != PZK_OPERAND_NONE) idx |= (u32)((LDO(b) >> ((x.w >> 8) & 255u)) & 1) << 1;
        if (x.x != PZK_OPERAND_NONE) idx |= (u32)((LDO(x.x) >> ((x.w >> 16) & 255u)) & 1) << 2;
        if (x.y != PZK_OPERAND_NONE) idx |= (u32)((LDO(x.y) >> (x.w >> 24)) & 1) << 3;
        if (opc == PZK_U_LUT) STD(dst, (u64)((imm16 >> idx) & 1));
        else STD(dst, (u64)__ldg(list + x.z + 2 * idx) | ((u64)__ldg(list + x.z + 2 * idx + 1) << 32));
        break;
      }
      case PZK_V_LUT: {
        FETCH_EXT();  // {c, d, rotations, lane mask low}
        const u32 rot = x.z;
        if (flags & PZK_FLAG_W64) {
          const uint4 y = __ldg(ops + (++pc));
          const u64 lanes = (u64)x.w | ((u64)y.x << 32);
          const u64 x0 = rotr64(LDO(a), rot & 255u);
          const u64 x1 = b != PZK_OPERAND_NONE ? rotr64(LDO(b), (rot >> 8) & 255u) : 0;
          const u64 x2 = x.x != PZK_OPERAND_NONE ? rotr64(LDO(x.x), (rot >> 16) & 255u) : 0;
          const u64 x3 = x.y != PZK_OPERAND_NONE ? rotr64(LDO(x.y), rot >> 24) : 0;
          STD(dst, vlut_eval<u64>(imm16, x0, x1, x2, x3) & lanes);
        } else {
          const u32 x0 = rotr32((u32)LDO(a), rot & 255u);
          const u32 x1 = b != PZK_OPERAND_NONE ? rotr32((u32)LDO(b), (rot >> 8) & 255u) : 0;
          const u32 x2 = x.x != PZK_OPERAND_NONE ? rotr32((u32)LDO(x.x), (rot >> 16) & 255u) : 0;
          const u32 x3 = x.y != PZK_OPERAND_NONE ? rotr32((u32)LDO(x.y), rot >> 24) : 0;
          STD(dst, (u64)(vlut_eval<u32>(imm16, x0, x1, x2, x3) & x.w));
        }
        break;
      }
      case PZK_U_EXTRACT: {
        const u32 s_ = imm16 & 255u, k_ = imm16 >> 8, n_ = b;
        u64 v;
        if (flags & PZK_FLAG_NBASE) { u64 t[4], r[4]; LDFA(t); field256(r, t, s_, n_, k_); v = r[0]; }
        else { v = LDO(a) >> s_; if (n_ < 64) v &= (1ull << n_) - 1; v <<= k_; }
        STD(dst, v);
        break;
      }
      case PZK_N_EXTRACT: {
        const u32 s_ = imm16 & 255u, k_ = imm16 >> 8, n_ = b;
        u64 t[4] = {0, 0, 0, 0}, r[4];
        if (flags & PZK_FLAG_NBASE) LDFA(t); else t[0] = LDO(a);
        field256(r, t, s_, n_, k_);
        STFD(r);
        break;
      }
      case PZK_CHECK_RANGE: {
        if (check_rows) {
          bool ok;
          if (flags & PZK_FLAG_NBASE) { u64 t[4], r[4]; LDFA(t); shr256(r, t, imm16); ok = (r[0] | r[1] | r[2] | r[3]) == 0; }
          else ok = imm16 >= 64 || (LDO(a) >> imm16) == 0;
          if (!ok && dst < bad) bad = dst;
        }
        break;
      }
      case PZK_F_CONST: { u64 v[4]; ldPool(fpool, a, v); STFD(v); break; }
      case PZK_F_NEG: { u64 va[4], r[4]; LDFA(va); fr_neg(r, va); STFD(r); break; }
      case PZK_F_INV: { u64 va[4], r[4]; LDFA(va); fr_inv(r, va); STFD(r); break; }
      case PZK_F_FROM_U: { u64 va[4] = {LDO(a), 0, 0, 0}, r[4]; fr_to_mont(r, va); STFD(r); break; }
      case PZK_F_FROM_I: {
        long long v = (long long)LDO(a);
        u64 va[4] = {v < 0 ? (u64)(-v) : (u64)v, 0, 0, 0}, r[4];
        fr_to_mont(r, va);
        if (v < 0) fr_neg(r, r);
        STFD(r);
        break;
      }
      case PZK_F_SEL: { FETCH_EXT(); u64 v[4]; ldFo(Fl, L, cells, NT, LDO(a) ? b : x.x, v); STFD(v); break; }
      case PZK_F_EQ: case PZK_F_NE: {
        u64 va[4], vb[4];
        LDFA(va); LDFB(vb);
        bool eq = fr_eq(va, vb);
        STD(dst, (u64)(opc == PZK_F_EQ ? eq : !eq));
        break;
      }
      case PZK_F_CSEL: { u64 v[4]; ldPool(fpool, b + (u32)LDO(a), v); STFD(v); break; }
      case PZK_N_FROM_F: {
        u64 va[4], r[4]; LDFA(va);
        if (flags & PZK_FLAG_ZSRC) z_canonical(r, va); else fr_from_mont(r, va);
        STFD(r); break;
      }
      case PZK_F_FROM_N: {
        u64 va[4], r[4]; LDFA(va);
        if (flags & PZK_FLAG_ZSRC) z_to_mont(r, va); else { reduce_p(va); fr_to_mont(r, va); }
        STFD(r); break;
      }
      case PZK_F_MULADD: {
        FETCH_EXT();
        u64 va[4], vb[4], r[4];
        LDFA(va); LDFB(vb);
        fr_mul(r, va, vb);
        if (flags & PZK_FLAG_DST2) MULADD_PRODUCT(r);
        ldFo(Fl, L, cells, NT, x.x, va);
        if (imm16 & 0x100u) fr_sub(r, va, r);          // c - a b
        else if (imm16 & 0x200u) fr_sub(r, r, va);     // a b - c
        else fr_add(r, r, va);
        STFD(r);
        break;
      }
      case PZK_Z_MULADD: {
        FETCH_EXT();
        u64 va[4], vb[4], r[4];
        if (flags & PZK_FLAG_A_U) { va[0] = LDO(a); va[1] = va[2] = va[3] = 0; } else LDFA(va);
        if (flags & PZK_FLAG_B_U) { vb[0] = LDO(b); vb[1] = vb[2] = vb[3] = 0; } else LDFB(vb);
        z_mul(r[0], r[1], r[2], r[3], va[0], va[1], va[2], va[3], vb[0], vb[1], vb[2], vb[3], imm16);
        if (flags & PZK_FLAG_DST2) MULADD_PRODUCT(r);
        ldFo(Fl, L, cells, NT, x.x, va);
        if (imm16 & 0x100u) sub256(r, va, r);
        else if (imm16 & 0x200u) sub256(r, r, va);
        else add256(r, r, va);
        STFD(r);
        break;
      }
      case PZK_Z_FROM_U: { u64 v[4] = {LDO(a), 0, 0, 0}; STFD(v); break; }
      case PZK_Z_FROM_I: { const u64 x_ = LDO(a); const u64 sx = (u64)((long long)x_ >> 63); u64 v[4] = {x_, sx, sx, sx}; STFD(v); break; }
      case PZK_Z_CONST: { u64 v[4]; ldPool(fpool, a, v); STFD(v); break; }
      case PZK_N_FROM_U: { u64 v[4] = {LDO(a), 0, 0, 0}; STFD(v); break; }
      case PZK_N_BIT: {
        u64 limb = 0;
        if (b < 256) limb = (a & PZK_OPERAND_CELL) ? lds64(CELL_ADDR(a) + ((b >> 6) << 10)) : PLD(Fl + ((u64)a * 4 + (b >> 6)) * L);
        STD(dst, (limb >> (b & 63)) & 1);
        break;
      }
      case PZK_N_LOW: STD(dst, (a & PZK_OPERAND_CELL) ? lds64(CELL_ADDR(a)) : PLD(Fl + (u64)a * 4 * L)); break;
      case PZK_N_FITS: { u64 v[4]; LDFA(v); STD(dst, (u64)((v[1] | v[2] | v[3]) == 0)); break; }
      case PZK_N_SHR: { u64 v[4], r[4]; LDFA(v); u64 d = UBV; shr256(r, v, d > 256 ? 256u : (unsigned)d); STFD(r); break; }
      case PZK_N_SHL: {
        u64 v[4], r[4] = {0, 0, 0, 0}; LDFA(v); u64 d = UBV;
        if (d < 254) { shl256(r, v, (unsigned)d); r[3] &= 0x3fffffffffffffffull; reduce_p(r); }
        STFD(r); break;
      }
      case PZK_N_AND: case PZK_N_OR: case PZK_N_XOR: {
        u64 va[4], vb[4], r[4];
        LDFA(va); LDFB(vb);
#pragma unroll
        for (int i = 0; i < 4; i++) r[i] = opc == PZK_N_AND ? (va[i] & vb[i]) : opc == PZK_N_OR ? (va[i] | vb[i]) : (va[i] ^ vb[i]);
        r[3] &= 0x3fffffffffffffffull;
        reduce_p(r);
        STFD(r);
        break;
      }
      case PZK_N_DIV: case PZK_N_MOD: {
        u64 va[4], vb[4], r[4];
        LDFA(va); LDFB(vb);
        if (opc == PZK_N_DIV) divmod256(va, vb, r, nullptr); else divmod256(va, vb, nullptr, r);
        STFD(r);
        break;
      }
      case PZK_N_SLT: case PZK_N_SLE: {
        u64 va[4], vb[4];
        LDFA(va); LDFB(vb);
        int c = scmp256(va, vb);
        STD(dst, (u64)(opc == PZK_N_SLT ? c < 0 : c <= 0));
        break;
      }
      case PZK_CHECK_I64: {
        // |A|,|B|,|A*B|,|C| < 2^63 proven at compile time: wrapping 64-bit arithmetic is exact
        if (check_rows) {
          const uint2* terms = reinterpret_cast<const uint2*>(ops + pc + 1);
          const u32 na = imm16, nab = na + (a & 0xffffu), tot = nab + (a >> 16);
          long long A = 0, B = 0, C = 0;
          u32 k = 0;
#define I64_TERM(acc)                                                                         \
  {                                                                                           \
    const uint2 t = __ldg(terms + k);                                                         \
    const long long v = (t.x >= PZK_REF_ONE_LIST) ? 1ll : (long long)TERM_U(t.x);             \
    acc += (long long)(int)t.y * v;                                                           \
  }
          for (; k < na; k++) I64_TERM(A)
          for (; k < nab; k++) I64_TERM(B)
          for (; k < tot; k++) I64_TERM(C)
#undef I64_TERM
          const bool ok = (na == 0 || nab == na) ? (C == 0) : (A * B == C);
          if (!ok && dst < bad) bad = dst;
        }
        pc += b;
        break;
      }
      case PZK_CHECK_INT: case PZK_CHECK_F: {
        const u32 row_recs = b;
        if (check_rows) {
          const uint4* recs = ops + pc + 1;
          const u32 na = imm16, nb = a & 0xffffu, nc = a >> 16;
          bool ok = (opc == PZK_CHECK_INT) ? check_row_int(recs, na, nb, nc, list, Ul, L, cells, NT)
                                           : check_row_field(sc, recs, na, nb, nc, Ul, Fl, L, cells, NT);
          if (!ok && dst < bad) bad = dst;
        }
        pc += row_recs;
        break;
      }
      case PZK_BIGDIV: st |= bigdiv_device(list + a, Ul, L); break;
      case PZK_MODINV: modinv_device(list + a, Ul, L); break;
      case PZK_BJJ_MUL8: st |= bjj_mul8_device(list + a, fpool, Fl, L); break;
      case PZK_ASSERT_NZ: if (LDO(a) == 0) st |= PZK_LANE_ASSERT; break;
      case PZK_IN_U: {
        if (p.in_table) {
          const uint2 e = __ldg(p.in_table + a);
          const unsigned char* base = reinterpret_cast<const unsigned char*>(p.inputs) + lane_now() * p.in_stride + e.y;
          u64 v = (e.x == 0) ? (u64)*base : *reinterpret_cast<const u64*>(base);
          if (imm16 < 64 && (v >> imm16) != 0) st |= PZK_LANE_INPUT_RANGE;
          STD(dst, v);
          break;
        }
        const ulonglong2* ip = reinterpret_cast<const ulonglong2*>(p.inputs + (lane_now() * p.n_inputs + a) * 4);
        ulonglong2 lo = ip[0], hi = ip[1];
        if ((lo.y | hi.x | hi.y) != 0 || (imm16 < 64 && (lo.x >> imm16) != 0)) st |= PZK_LANE_INPUT_RANGE;
        STD(dst, lo.x);
        break;
      }
      case PZK_IN_F: {
        const ulonglong2* ip = p.in_table
            ? reinterpret_cast<const ulonglong2*>(reinterpret_cast<const unsigned char*>(p.inputs) + lane_now() * p.in_stride + __ldg(p.in_table + a).y)
            : reinterpret_cast<const ulonglong2*>(p.inputs + (lane_now() * p.n_inputs + a) * 4);
        ulonglong2 lo = ip[0], hi = ip[1];
        u64 v[4] = {lo.x, lo.y, hi.x, hi.y}, r[4];
        if (geq_p(v)) { st |= PZK_LANE_INPUT_RANGE; reduce_p(v); }
        fr_to_mont(r, v);
        STFD(r);
        break;
      }
      default: st |= 0x80000000u; break;
    }
    if (flags & PZK_FLAG_DIG) {
      // the value this op just defined, from its cache cell or its slot, folded into the digest
      const uint4 dg = __ldg(ops + (++pc));
      if (digest && (dg.x & 15u)) dig_fold(dacc, dig_tab, dg, rv0, rv1, rv2, rv3);
    }
  }
  if (digest) {
    // flush the accumulators to the per-lane carry-save state (pieces of 32 bits, see digest_kernel)
    unsigned long long* stt = reinterpret_cast<unsigned long long*>(p.dig_state) + p.dig_lane_base + lane_now();
    const u64 S = p.dig_stride;
    const int piece0[DIG_ACC_WORDS] = {0, 2, 8, 10, 12, 14, 16, 18, 20, 22, 24, 26};
#pragma unroll
    for (int k = 0; k < DIG_ACC_WORDS; k++) {
      const u64 a = lds64(dacc + 1024 * k);
      const u64 lo = a & 0xffffffffull, hi = a >> 32;
      if (lo) atomicAdd(stt + (u64)piece0[k] * S, (unsigned long long)lo);
      if (hi) atomicAdd(stt + (u64)(piece0[k] + 1) * S, (unsigned long long)hi);
    }
  }
  if (bad != 0xffffffffu) {
    st |= PZK_LANE_CONSTRAINT;
    const u64 lane = lane_now();
    if (bad < p.first_bad[lane]) p.first_bad[lane] = bad;
  }
  if (st) p.status[lane_now()] |= st;
}

// ------------------------------------------------------------------------------------------
// Export: entries (wire <- slot) of one segment for a list of lanes -> canonical 32-byte wires.
#undef dacc
// grid.x covers lanes (fast, coalesced plane reads), grid.y strides over entries.
// ------------------------------------------------------------------------------------------
struct ExportParams {
  const u32* list;           // list pool (table-view descriptors)
  u32 n_u_slots, n_f_slots;  // blocked planes
  const PzkExport* entries;
  u64 n_entries;
  const u64* U;
  const u64* F;
  u64 L;
  const u64* lanes;  // tile-local lane per output row, or nullptr = identity
  const u64* rows;   // output row per entry of `lanes`, or nullptr = lane_base + index
  u64 lane_base;     // output row offset (identity mode: output row = lane_base + lane)
  u64 n_rows;        // number of output rows handled
  u64* out;          // [row][out_wires][4], or (blocked) [row / 32][out_wires][limb][row % 32]
  u64 out_wires;
  u32 wire_off;      // out index = wire - wire_off
  u32 blocked;       // 1: the R1CS stream kernel's layout (pzk_r1cs.cuh) - device-resident hand-off
};

__device__ __forceinline__ void export_store(const ExportParams& p, u64 row_out, u32 wire, const u64* w) {
  if (p.blocked) {
    u64* dst = p.out + ((row_out >> 5) * p.out_wires + (wire - p.wire_off)) * 128 + (row_out & 31);
    dst[0] = w[0]; dst[32] = w[1]; dst[64] = w[2]; dst[96] = w[3];
  } else {
    u64* dst = p.out + (row_out * p.out_wires + (wire - p.wire_off)) * 4;
    reinterpret_cast<ulonglong2*>(dst)[0] = make_ulonglong2(w[0], w[1]);
    reinterpret_cast<ulonglong2*>(dst)[1] = make_ulonglong2(w[2], w[3]);
  }
}

// one export entry of one lane -> canonical value (wire <- slot, or wire <- view of words; pzk_program.h)
__device__ __forceinline__ void export_value(const ExportParams& p, const uint4 ew, u64 lane, u64* w) {
  const u32 ref = ew.y, aux = ew.z;
  w[0] = w[1] = w[2] = w[3] = 0;
  if (ref == PZK_REF_ZERO) return;
  if (ref == PZK_REF_ONE) { w[0] = 1; return; }
  const u64* Ub = p.U + (lane / PZK_LANE_BLOCK) * p.n_u_slots * PZK_LANE_BLOCK + (lane % PZK_LANE_BLOCK);
  const u64* Fb = p.F + (lane / PZK_LANE_BLOCK) * p.n_f_slots * 4 * PZK_LANE_BLOCK + (lane % PZK_LANE_BLOCK);
  if (ref == PZK_REF_TABVIEW) {
    const u32* Lp = p.list + aux;
    const u32 n = __ldg(Lp);
    u32 idx = 0;
    for (u32 j = 0; j < n; j++) idx |= (u32)((Ub[(u64)__ldg(Lp + 1 + 2 * j) * PZK_LANE_BLOCK] >> __ldg(Lp + 2 + 2 * j)) & 1) << j;
    const long long v = (long long)((u64)__ldg(Lp + 1 + 2 * n + 2 * idx) | ((u64)__ldg(Lp + 2 + 2 * n + 2 * idx) << 32));
    if (v < 0) { const u64 pp[4] = {P0, P1, P2, P3}; u64 m[4] = {(u64)(-v), 0, 0, 0}; sub256(w, pp, m); } else w[0] = (u64)v;
    return;
  }
  const u32 cls = PZK_REF_CLS(ref), slot = PZK_REF_SLOT(ref);
  if (cls == 3) {
    const u32 s_ = aux & 255u, n_ = (aux >> 8) & 255u, k_ = (aux >> 16) & 255u;
    u64 t[4] = {0, 0, 0, 0};
    if (ref & PZK_REF_VIEW_N) ldF(Fb, PZK_LANE_BLOCK, slot, t); else t[0] = Ub[(u64)slot * PZK_LANE_BLOCK];
    field256(w, t, s_, n_, k_);
  } else if (cls == 2) {
    u64 m[4]; ldF(Fb, PZK_LANE_BLOCK, slot, m);
    if (ref & PZK_REF_Z) z_canonical(w, m); else fr_from_mont(w, m);
  }
  else {
    const u64 v = Ub[(u64)slot * PZK_LANE_BLOCK];
    if (cls == 1 && (long long)v < 0) { const u64 pp[4] = {P0, P1, P2, P3}; u64 m[4] = {(u64)(-(long long)v), 0, 0, 0}; sub256(w, pp, m); }
    else w[0] = v;
  }
}

// lanes across the threads (coalesced plane reads), grid.y strides over the entries: public signals of every
// lane, full witnesses of many selected lanes
__global__ void __launch_bounds__(128) export_kernel(ExportParams p) {
  const u64 row = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= p.n_rows) return;
  const u64 lane = p.lanes ? p.lanes[row] : row;
  const u64 row_out = p.rows ? p.rows[row] : p.lane_base + row;
  for (u64 e = blockIdx.y; e < p.n_entries; e += gridDim.y) {
    const uint4 ew = __ldg(reinterpret_cast<const uint4*>(p.entries + e));
    u64 w[4];
    export_value(p, ew, lane, w);
    export_store(p, row_out, ew.x, w);
  }
}

// entries across the threads, grid.y = selected lane: full witnesses of a few lanes in ONE launch per segment
__global__ void __launch_bounds__(128) export_rows_kernel(ExportParams p) {
  const u64 lane = p.lanes[blockIdx.y];
  const u64 row_out = p.rows ? p.rows[blockIdx.y] : p.lane_base + blockIdx.y;
  for (u64 e = (u64)blockIdx.x * blockDim.x + threadIdx.x; e < p.n_entries; e += (u64)gridDim.x * blockDim.x) {
    const uint4 ew = __ldg(reinterpret_cast<const uint4*>(p.entries + e));
    u64 w[4];
    export_value(p, ew, lane, w);
    export_store(p, row_out, ew.x, w);
  }
}

// ------------------------------------------------------------------------------------------
// The BabyJubjub ladder as a kernel of its own.  Inside eval_kernel (72 registers, points behind references of
// __noinline__ helpers) the 8 600 serial Montgomery products of PZK_BJJ_MUL8 ran at IPC 0.2 - 114 ms of a 1 065 ms
// step for ONE record.  The compiler gives the record a segment of its own; the runtime launches this kernel for it:
// same algorithm (bjj_mul8_device), everything inlined, the three points in registers, no register cap.
// ------------------------------------------------------------------------------------------
struct BjjParams {
  const u32* list;   // list pool at the record's operand list
  const u64* fpool;
  u64* F;
  u64 n_lanes;
  u32 n_f_slots;
  u32* status;
};

__device__ __forceinline__ void bjj_padd_reg(u64* oX, u64* oY, u64* oZ, const u64* pX, const u64* pY, const u64* pZ,
                                             const u64* qX, const u64* qY, const u64* qZ, const u64* ca, const u64* cd) {
  u64 A[4], B[4], C[4], D[4], E[4], F[4], G[4], t[4], u[4];
  fr_mul(A, pZ, qZ); fr_mul(B, A, A);
  fr_mul(C, pX, qX); fr_mul(D, pY, qY);
  fr_mul(E, C, D); fr_mul(E, E, cd);
  fr_sub(F, B, E); fr_add(G, B, E);
  fr_add(t, pX, pY); fr_add(u, qX, qY); fr_mul(t, t, u); fr_sub(t, t, C); fr_sub(t, t, D);
  fr_mul(u, A, F); fr_mul(oX, u, t);
  fr_mul(t, C, ca); fr_sub(t, D, t);
  fr_mul(u, A, G); fr_mul(oY, u, t);
  fr_mul(oZ, F, G);
}

__global__ void __launch_bounds__(128) bjj_kernel(BjjParams p) {
  const u64 lane = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= p.n_lanes) return;
  const u64 L = PZK_LANE_BLOCK;
  u64* Fl = p.F + (u64)blockIdx.x * p.n_f_slots * 4 * PZK_LANE_BLOCK + threadIdx.x;
  const u32* Lst = p.list;
  const u32 n = Lst[0], nadd = 2 * n - 1;
  u64 ca[4], cd[4], bx[4], by[4], one[4], sc[4];
  ldPool(p.fpool, Lst[1], ca); ldPool(p.fpool, Lst[2], cd); ldPool(p.fpool, Lst[3], bx); ldPool(p.fpool, Lst[4], by);
  { const u64 o1[4] = {1, 0, 0, 0}; fr_to_mont(one, o1); }
  { u64 m[4]; ldF(Fl, L, Lst[5], m); fr_from_mont(sc, m); }
  const u32* out = Lst + 6;
  const u32* scr = out + 2 * nadd;
  u64 SX[4] = {0, 0, 0, 0}, SY[4] = {0, 0, 0, 0}, SZ[4] = {one[0], one[1], one[2], one[3]};
  u64 pre[4] = {one[0], one[1], one[2], one[3]};
  u32 k = 0;
  for (u32 i = 0; i < n; i++) {
    u64 DX[4], DY[4], DZ[4];
    if (i > 0) {
      bjj_padd_reg(DX, DY, DZ, SX, SY, SZ, SX, SY, SZ, ca, cd);
      stF(Fl, L, out[2 * k], DX); stF(Fl, L, out[2 * k + 1], DY); stF(Fl, L, scr[2 * k], DZ);
      fr_mul(pre, pre, DZ); stF(Fl, L, scr[2 * k + 1], pre);
      k++;
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) { DX[j] = DY[j] = 0; DZ[j] = one[j]; }
    }
    const u32 bi = n - 1 - i;
    const bool bit = (sc[bi >> 6] >> (bi & 63)) & 1;
    u64 QX[4], QY[4], AX[4], AY[4], AZ[4];
#pragma unroll
    for (int j = 0; j < 4; j++) { QX[j] = bit ? bx[j] : 0; QY[j] = bit ? by[j] : 0; }
    bjj_padd_reg(AX, AY, AZ, DX, DY, DZ, QX, QY, one, ca, cd);
    stF(Fl, L, out[2 * k], AX); stF(Fl, L, out[2 * k + 1], AY); stF(Fl, L, scr[2 * k], AZ);
    fr_mul(pre, pre, AZ); stF(Fl, L, scr[2 * k + 1], pre);
    k++;
    // addZeroBabyjub: in1 "zero" -> in2, else in2 "zero" -> in1, else the sum (zero = x coordinate 0)
    const bool zd = fr_is_zero(DX), zq = !bit;
#pragma unroll
    for (int j = 0; j < 4; j++) {
      SX[j] = zd ? QX[j] : (zq ? DX[j] : AX[j]);
      SY[j] = zd ? QY[j] : (zq ? DY[j] : AY[j]);
      SZ[j] = zd ? one[j] : (zq ? DZ[j] : AZ[j]);
    }
  }
  u32 st = fr_is_zero(pre) ? PZK_LANE_HINT : 0;
  u64 inv[4];
  fr_inv(inv, pre);
  for (u32 kk = nadd; kk-- > 0;) {
    u64 z[4], zi[4], x[4], y[4];
    ldF(Fl, L, scr[2 * kk], z);
    if (kk > 0) { u64 pp[4]; ldF(Fl, L, scr[2 * kk - 1], pp); fr_mul(zi, inv, pp); }
    else { zi[0] = inv[0]; zi[1] = inv[1]; zi[2] = inv[2]; zi[3] = inv[3]; }
    fr_mul(inv, inv, z);
    ldF(Fl, L, out[2 * kk], x); ldF(Fl, L, out[2 * kk + 1], y);
    fr_mul(x, x, zi); fr_mul(y, y, zi);
    stF(Fl, L, out[2 * kk], x); stF(Fl, L, out[2 * kk + 1], y);
  }
  if (st) p.status[lane] |= st;
}

// ------------------------------------------------------------------------------------------
// Witness digest: every wire of every lane folded into one field element per lane without writing the 72 MB
// witness:      digest = sum over wires i of c(i) * w_i   mod p,     c(i) = (splitmix64(i) >> 32) | 1  (32 bits, odd)
// w_i the canonical value of wire i (.wtns section 2).  The weights make it position sensitive; a sum does not
// care in which order the wires are visited, and it is linear in every representation the evaluator keeps:
//   * a bit-field view ((W >> s) & (2^n - 1)) << k is a sum of bits of W times powers of two, so ALL the views of
//     one word (110 per word on average in registerIdentity) collapse into one table of per-bit coefficients
//     C_b = sum over views containing bit b of c(i) 2^(b - s + k), built when the program is loaded:
//     one conditional 128-bit add per bit of the word instead of a shift / mask / multiply per wire (this kernel),
//     one indexed load + add per nibble from 16-entry subset-sum tables in the evaluator's fused fold (dig_fold_table);
//   * a Montgomery value w R is accumulated as the 320-bit integer sum of c(i) * (w R) and reduced and converted
//     ONCE per lane at the end (no from_mont per wire);
//   * narrow words are 64 x 32 -> 96-bit products in a 128-bit accumulator (negative I-class values in a second).
// The export entries are compiled into 16-byte digest records per program segment (pzk_api.cu: build_digest_program).
// Partial sums of one launch are added to the per-lane state in carry-save form (32-bit pieces in 64-bit words,
// atomicAdd), so grid.y can split the records when there are few lanes; digest_finalize_kernel normalises.
// /root/reference/test/automatisationTest.js:40-50 returns the whole vector; this is its checksum.
// ------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ u32 pzk_digest_weight(u32 wire) {
  u64 z = (u64)wire + 0x9e3779b97f4a7c15ull;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
  return (u32)((z ^ (z >> 31)) >> 32) | 1u;
}

enum { DIG_WORD_U = 1, DIG_WORD_N = 2, DIG_PLAIN_U = 3, DIG_PLAIN_I = 4, DIG_PLAIN_F = 5, DIG_GENERIC = 6, DIG_CONST = 7, DIG_PLAIN_Z = 8 };
struct DigRec { u32 type_nbits; u32 slot; u32 a; u32 b; };  // type in bits 0..7, nbits in bits 8..31
#define DIG_STATE_PIECES 28  // accN 4, accNeg 4, accM 10, accP 10 (32-bit pieces, carry-save)

struct DigestParams {
  const DigRec* recs;
  u64 n_recs;
  const ulonglong2* tab;     // per-bit coefficients (lo, hi)
  const PzkExport* exports;  // for DIG_GENERIC
  ExportParams ex;           // planes, list pool
  u64 n_lanes;
  u64* state;                // [piece][state_stride]
  u64 state_stride;
  u64 lane_base;
};

__device__ __forceinline__ void add128(u64* acc, u64 lo, u64 hi) {
  asm("add.cc.u64 %0, %0, %2; addc.u64 %1, %1, %3;" : "+l"(acc[0]), "+l"(acc[1]) : "l"(lo), "l"(hi));
}
// acc (5 x 64) += c * w (4 x 64), c < 2^32
__device__ __forceinline__ void mac320(u64* acc, u32 c, const u64* w) {
  u32 a[10], v[8];
#pragma unroll
  for (int j = 0; j < 5; j++) { a[2 * j] = (u32)acc[j]; a[2 * j + 1] = (u32)(acc[j] >> 32); }
#pragma unroll
  for (int j = 0; j < 4; j++) { v[2 * j] = (u32)w[j]; v[2 * j + 1] = (u32)(w[j] >> 32); }
  mad320_chain(a, c, v);
#pragma unroll
  for (int j = 0; j < 5; j++) acc[j] = (u64)a[2 * j] | ((u64)a[2 * j + 1] << 32);
}

__global__ void __launch_bounds__(128) digest_kernel(DigestParams p) {
  const u64 lane = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= p.n_lanes) return;
  const u64 per = (p.n_recs + gridDim.y - 1) / gridDim.y;
  const u64 r0 = (u64)blockIdx.y * per, r1 = min(p.n_recs, r0 + per);
  const u64* Ub = p.ex.U + (lane / PZK_LANE_BLOCK) * p.ex.n_u_slots * PZK_LANE_BLOCK + (lane % PZK_LANE_BLOCK);
  const u64* Fb = p.ex.F + (lane / PZK_LANE_BLOCK) * p.ex.n_f_slots * 4 * PZK_LANE_BLOCK + (lane % PZK_LANE_BLOCK);
  u64 accN[2] = {0, 0}, accNeg[2] = {0, 0}, accM[5] = {0, 0, 0, 0, 0}, accP[5] = {0, 0, 0, 0, 0};
  for (u64 r = r0; r < r1; r++) {
    const uint4 rw = __ldg(reinterpret_cast<const uint4*>(p.recs + r));
    const u32 type = rw.x & 255u, nbits = rw.x >> 8;
    switch (type) {
      case DIG_WORD_U: {
        const u64 W = Ub[(u64)rw.y * PZK_LANE_BLOCK];
        const ulonglong2* T = p.tab + rw.z;
#pragma unroll 4
        for (u32 b = 0; b < nbits; b++) {
          const ulonglong2 t = __ldg(T + b);
          const u64 m = 0ull - ((W >> b) & 1ull);
          add128(accN, t.x & m, t.y & m);
        }
        break;
      }
      case DIG_WORD_N: {
        const ulonglong2* T = p.tab + rw.z;
        for (u32 j = 0; j * 64 < nbits; j++) {
          const u64 W = Fb[((u64)rw.y * 4 + j) * PZK_LANE_BLOCK];
          const u32 nb = min(64u, nbits - j * 64);
#pragma unroll 4
          for (u32 b = 0; b < nb; b++) {
            const ulonglong2 t = __ldg(T + j * 64 + b);
            const u64 m = 0ull - ((W >> b) & 1ull);
            add128(accN, t.x & m, t.y & m);
          }
        }
        break;
      }
      case DIG_PLAIN_U: {
        const u64 v = Ub[(u64)rw.y * PZK_LANE_BLOCK];
        add128(accN, (u64)rw.z * v, __umul64hi((u64)rw.z, v));
        break;
      }
      case DIG_PLAIN_I: {
        const long long v = (long long)Ub[(u64)rw.y * PZK_LANE_BLOCK];
        const u64 mag = v < 0 ? (u64)(-v) : (u64)v;
        if (v < 0) add128(accNeg, (u64)rw.z * mag, __umul64hi((u64)rw.z, mag)); else add128(accN, (u64)rw.z * mag, __umul64hi((u64)rw.z, mag));
        break;
      }
      case DIG_PLAIN_F: {
        u64 w[4];
        ldF(Fb, PZK_LANE_BLOCK, rw.y, w);
        mac320(accM, rw.z, w);
        break;
      }
      case DIG_PLAIN_Z: {
        u64 z[4], w[4];
        ldF(Fb, PZK_LANE_BLOCK, rw.y, z);
        z_canonical(w, z);
        mac320(accP, rw.z, w);
        break;
      }
      case DIG_GENERIC: {
        const uint4 ew = __ldg(reinterpret_cast<const uint4*>(p.exports + rw.z));
        if (ew.y == PZK_REF_TABVIEW) {  // a truth-table entry: a small signed integer, no 256-bit arithmetic
          const u32* Lp = p.ex.list + ew.z;
          const u32 n = __ldg(Lp);
          u32 idx = 0;
          for (u32 j = 0; j < n; j++) idx |= (u32)((Ub[(u64)__ldg(Lp + 1 + 2 * j) * PZK_LANE_BLOCK] >> __ldg(Lp + 2 + 2 * j)) & 1) << j;
          const long long v = (long long)((u64)__ldg(Lp + 1 + 2 * n + 2 * idx) | ((u64)__ldg(Lp + 2 + 2 * n + 2 * idx) << 32));
          const u64 mag = v < 0 ? (u64)(-v) : (u64)v;
          if (v < 0) add128(accNeg, (u64)rw.w * mag, __umul64hi((u64)rw.w, mag)); else add128(accN, (u64)rw.w * mag, __umul64hi((u64)rw.w, mag));
          break;
        }
        u64 w[4];
        export_value(p.ex, ew, lane, w);
        mac320(accP, rw.w, w);
        break;
      }
      case DIG_CONST: add128(accN, (u64)rw.z | ((u64)rw.w << 32), 0); break;
      default: break;
    }
  }
  unsigned long long* st = reinterpret_cast<unsigned long long*>(p.state) + p.lane_base + lane;
  const u64 S = p.state_stride;
  auto put = [&](u32 piece0, const u64* acc, int n) {
    for (int j = 0; j < n; j++) {
      const u64 lo = acc[j] & 0xffffffffull, hi = acc[j] >> 32;
      if (lo) atomicAdd(st + (u64)(piece0 + 2 * j) * S, (unsigned long long)lo);
      if (hi) atomicAdd(st + (u64)(piece0 + 2 * j + 1) * S, (unsigned long long)hi);
    }
  };
  put(0, accN, 2); put(4, accNeg, 2); put(8, accM, 5); put(18, accP, 5);
}

// pieces (value = sum piece_k 2^(32 k)) -> limbs
__device__ __forceinline__ void dig_normalise(const u64* state, u64 S, u32 piece0, int n_pieces, u64* limbs, int n_limbs) {
  for (int j = 0; j < n_limbs; j++) limbs[j] = 0;
  for (int k = 0; k < n_pieces; k++) {
    const unsigned __int128 add = (unsigned __int128)state[(u64)(piece0 + k) * S] << (32 * (k & 1));  // < 2^96
    const int j = k >> 1;
    unsigned __int128 t = (unsigned __int128)limbs[j] + (u64)add;
    limbs[j] = (u64)t;
    u64 carry = (u64)(t >> 64) + (u64)(add >> 64);
    for (int q = j + 1; q < n_limbs && carry; q++) {
      t = (unsigned __int128)limbs[q] + carry;
      limbs[q] = (u64)t;
      carry = (u64)(t >> 64);
    }
  }
}
// 320-bit integer -> residue mod p
__device__ __forceinline__ void dig_reduce320(const u64* x, u64* r) {
  u64 lo[4] = {x[0], x[1], x[2], x[3]};
  reduce_p(lo);
  const u64 hi[4] = {x[4], 0, 0, 0};
  u64 hr[4];
  fr_to_mont(hr, hi);  // x[4] * 2^256 mod p
  fr_add(r, lo, hr);
}

__global__ void digest_finalize_kernel(const u64* state, u64 S, u64 lane_base, u64 n_lanes, u64* digest) {
  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_lanes) return;
  const u64* st = state + lane_base + i;
  u64 n3[3], g3[3], m5[5], p5[5];
  dig_normalise(st, S, 0, 4, n3, 3);
  dig_normalise(st, S, 4, 4, g3, 3);
  dig_normalise(st, S, 8, 10, m5, 5);
  dig_normalise(st, S, 18, 10, p5, 5);
  u64 d[4], t[4], u[4];
  dig_reduce320(p5, d);
  dig_reduce320(m5, t);
  fr_from_mont(u, t);
  fr_add(d, d, u);
  // wire 0 is the constant 1 and has no export entry
  u64 nn[4] = {n3[0], n3[1], n3[2], 0}, gg[4] = {g3[0], g3[1], g3[2], 0};
  const u64 w0[4] = {pzk_digest_weight(0), 0, 0, 0};
  fr_add(d, d, nn);
  fr_add(d, d, w0);
  fr_sub(d, d, gg);
  u64* out = digest + (lane_base + i) * 4;
  out[0] = d[0]; out[1] = d[1]; out[2] = d[2]; out[3] = d[3];
}

}  // namespace pzkd
