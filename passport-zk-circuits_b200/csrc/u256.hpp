// u256.hpp - host-side 256-bit integers and BN254 Fr arithmetic for the circuit compiler.
// (Compile-time constant folding of circom expressions; the device has its own code.)
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <algorithm>

namespace pzk {

typedef unsigned __int128 u128;

struct U256 {
  uint64_t w[4];
  U256() { w[0] = w[1] = w[2] = w[3] = 0; }
  U256(uint64_t v) { w[0] = v; w[1] = w[2] = w[3] = 0; }
  static U256 from_limbs(uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
    U256 r; r.w[0] = a; r.w[1] = b; r.w[2] = c; r.w[3] = d; return r;
  }
  bool is_zero() const { return (w[0] | w[1] | w[2] | w[3]) == 0; }
  bool fits64() const { return (w[1] | w[2] | w[3]) == 0; }
  bool bit(unsigned i) const { return i < 256 && ((w[i >> 6] >> (i & 63)) & 1); }
  int bitlen() const {
    for (int i = 3; i >= 0; i--)
      if (w[i]) return i * 64 + 64 - __builtin_clzll(w[i]);
    return 0;
  }
  bool operator==(const U256& o) const { return memcmp(w, o.w, 32) == 0; }
  bool operator!=(const U256& o) const { return !(*this == o); }
};

inline int cmp(const U256& a, const U256& b) {
  for (int i = 3; i >= 0; i--) {
    if (a.w[i] < b.w[i]) return -1;
    if (a.w[i] > b.w[i]) return 1;
  }
  return 0;
}
inline bool operator<(const U256& a, const U256& b) { return cmp(a, b) < 0; }

inline U256 add(const U256& a, const U256& b, uint64_t* carry = nullptr) {
  U256 r; u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a.w[i] + b.w[i]; r.w[i] = (uint64_t)c; c >>= 64; }
  if (carry) *carry = (uint64_t)c;
  return r;
}
inline U256 sub(const U256& a, const U256& b, uint64_t* borrow = nullptr) {
  U256 r; uint64_t br = 0;
  for (int i = 0; i < 4; i++) {
    u128 t = (u128)a.w[i] - b.w[i] - br;
    r.w[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1;
  }
  if (borrow) *borrow = br;
  return r;
}
inline U256 shl(const U256& a, unsigned s) {
  U256 r; if (s >= 256) return r;
  unsigned ws = s >> 6, bs = s & 63;
  for (int i = 3; i >= 0; i--) {
    uint64_t v = 0;
    if (i >= (int)ws) {
      v = a.w[i - ws] << bs;
      if (bs && i - (int)ws - 1 >= 0) v |= a.w[i - ws - 1] >> (64 - bs);
    }
    r.w[i] = v;
  }
  return r;
}
inline U256 shr(const U256& a, unsigned s) {
  U256 r; if (s >= 256) return r;
  unsigned ws = s >> 6, bs = s & 63;
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    if (i + ws < 4) {
      v = a.w[i + ws] >> bs;
      if (bs && i + ws + 1 < 4) v |= a.w[i + ws + 1] << (64 - bs);
    }
    r.w[i] = v;
  }
  return r;
}
inline U256 band(const U256& a, const U256& b) { U256 r; for (int i = 0; i < 4; i++) r.w[i] = a.w[i] & b.w[i]; return r; }
inline U256 bor(const U256& a, const U256& b) { U256 r; for (int i = 0; i < 4; i++) r.w[i] = a.w[i] | b.w[i]; return r; }
inline U256 bxor(const U256& a, const U256& b) { U256 r; for (int i = 0; i < 4; i++) r.w[i] = a.w[i] ^ b.w[i]; return r; }

// 256 / 256 integer division (shift-subtract); q = a / b, r = a % b; b != 0
inline void divmod(const U256& a, const U256& b, U256& q, U256& r) {
  q = U256(); r = U256();
  if (b.fits64() && a.fits64()) { q = U256(a.w[0] / b.w[0]); r = U256(a.w[0] % b.w[0]); return; }
  int n = a.bitlen();
  for (int i = n - 1; i >= 0; i--) {
    r = shl(r, 1);
    if (a.bit(i)) r.w[0] |= 1;
    if (cmp(r, b) >= 0) { r = sub(r, b); q.w[i >> 6] |= 1ull << (i & 63); }
  }
}

// ---- BN254 scalar field -----------------------------------------------------
// p = 21888242871839275222246405745257275088548364400416034343698204186575808495617
static const U256 FR_P = U256::from_limbs(0x43e1f593f0000001ull, 0x2833e84879b97091ull,
                                          0xb85045b68181585dull, 0x30644e72e131a029ull);
static const uint64_t FR_INV = 0xc2e1f593efffffffull;  // -p^-1 mod 2^64
// R = 2^256 mod p, R2 = 2^512 mod p, R3 = 2^768 mod p
static const U256 FR_R = U256::from_limbs(0xac96341c4ffffffbull, 0x36fc76959f60cd29ull,
                                          0x666ea36f7879462eull, 0x0e0a77c19a07df2full);
static const U256 FR_R2 = U256::from_limbs(0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull,
                                           0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull);

inline U256 fr_half() { return shr(FR_P, 1); }

inline U256 fr_add(const U256& a, const U256& b) {
  uint64_t c; U256 r = add(a, b, &c);
  if (c || cmp(r, FR_P) >= 0) r = sub(r, FR_P);
  return r;
}
inline U256 fr_sub(const U256& a, const U256& b) {
  uint64_t br; U256 r = sub(a, b, &br);
  if (br) r = add(r, FR_P);
  return r;
}
inline U256 fr_neg(const U256& a) { return a.is_zero() ? a : sub(FR_P, a); }

// Montgomery product a*b*R^-1 mod p (CIOS, 64-bit limbs)
inline U256 fr_montmul(const U256& a, const U256& b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a.w[j] * b.w[i] + t[j];
      t[j] = (uint64_t)c; c >>= 64;
    }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * FR_INV;
    c = (u128)m * FR_P.w[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * FR_P.w[j] + t[j];
      t[j - 1] = (uint64_t)c; c >>= 64;
    }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  U256 r = U256::from_limbs(t[0], t[1], t[2], t[3]);
  if (t[4] || cmp(r, FR_P) >= 0) r = sub(r, FR_P);
  return r;
}
inline U256 fr_to_mont(const U256& a) { return fr_montmul(a, FR_R2); }
inline U256 fr_from_mont(const U256& a) { return fr_montmul(a, U256(1)); }
inline U256 fr_mul(const U256& a, const U256& b) { return fr_montmul(fr_montmul(a, b), FR_R2); }
inline U256 fr_pow(const U256& a, const U256& e) {
  U256 am = fr_to_mont(a), r = FR_R;
  for (int i = e.bitlen() - 1; i >= 0; i--) {
    r = fr_montmul(r, r);
    if (e.bit(i)) r = fr_montmul(r, am);
  }
  return fr_from_mont(r);
}
inline U256 fr_inv(const U256& a) {
  if (a.is_zero()) return a;
  return fr_pow(a, sub(FR_P, U256(2)));
}
inline U256 fr_reduce(const U256& a) {  // a < 2^256 -> a mod p
  U256 q, r; divmod(a, FR_P, q, r); return r;
}

inline U256 parse_number(const char* s, size_t n) {
  U256 r;
  if (n > 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) {
    for (size_t i = 2; i < n; i++) {
      char c = s[i]; unsigned d = (c <= '9') ? c - '0' : ((c | 32) - 'a' + 10);
      r = shl(r, 4); r.w[0] |= d;
    }
    return fr_reduce(r);
  }
  // decimal (may exceed p slightly in theory: reduce as we go, values here are < 2^256)
  for (size_t i = 0; i < n; i++) {
    // r = r*10 + d  (mod p)
    U256 r2 = fr_add(r, r), r4 = fr_add(r2, r2), r8 = fr_add(r4, r4);
    r = fr_add(r8, r2);
    r = fr_add(r, U256((uint64_t)(s[i] - '0')));
  }
  return r;
}

inline std::string to_dec(U256 a) {
  if (a.is_zero()) return "0";
  std::string s;
  while (!a.is_zero()) {
    // divide by 10^18
    uint64_t rem = 0;
    for (int i = 3; i >= 0; i--) {
      u128 cur = ((u128)rem << 64) | a.w[i];
      a.w[i] = (uint64_t)(cur / 1000000000000000000ull);
      rem = (uint64_t)(cur % 1000000000000000000ull);
    }
    for (int k = 0; k < 18; k++) { s.push_back('0' + rem % 10); rem /= 10; if (a.is_zero() && rem == 0) break; }
  }
  while (s.size() > 1 && s.back() == '0') s.pop_back();
  std::reverse(s.begin(), s.end());
  return s;
}

struct U256Hash {
  size_t operator()(const U256& a) const {
    uint64_t h = a.w[0] * 0x9E3779B97F4A7C15ull;
    h ^= a.w[1] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= a.w[2] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    h ^= a.w[3] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2);
    return (size_t)h;
  }
};

}  // namespace pzk
