/*
 * ORACLE (test infrastructure; never linked into or called by the product path).
 *
 * ssa_ref.c - plain-C, one-witness-at-a-time evaluator of a compiled program
 * (include/pzk_program.h) plus the constraint check, with straightforward
 * arithmetic (no Montgomery tricks beyond a textbook CIOS product, bit-serial
 * long division).  Two jobs:
 *   1. checker for the CUDA kernels: same program, same inputs, results must be
 *      bit-identical (tests/, __graft_entry__.smoke());
 *   2. the CPU baseline of bench.py ("port": the reference's own wasm calculator,
 *      /root/reference/test/automatisationTest.js:37-51, cannot run here - no node,
 *      no circom - see SURVEY.md section 8c/8d).
 * Together with oracle/circom_oracle.py (which interprets the reference's .circom
 * sources directly) it closes the loop: circom sources -> Python big-int witness
 * == compiled program evaluated here == CUDA kernels.
 *
 * Semantics restated per op from the circom operators they lower
 * (SURVEY.md section 8a footnote); long division follows
 * /root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom:190-232
 * (long_div: true quotient / remainder digits in base 2^n).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pzk_program.h"

typedef unsigned __int128 u128;

static const uint64_t P[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull,
                              0x30644e72e131a029ull};
static const uint64_t PINV = 0xc2e1f593efffffffull;
static const uint64_t R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull,
                               0x0216d0b17f4e44a5ull};
static const uint64_t ONE[4] = {1, 0, 0, 0};

typedef struct PzkRefProgram {
  PzkHeader h;
  uint8_t* blob;
  PzkSegment* segs;
  PzkOp* ops;
  uint64_t* fpool;
  PzkCoef* coefs;
  uint32_t* list;
  PzkInput* inputs;
  PzkRow* rows;
  PzkTerm* terms;
  PzkExport* exports;
  char* meta;
} PzkRefProgram;

static uint64_t align16(uint64_t x) { return (x + 15) & ~15ull; }

PzkRefProgram* pzk_ref_load(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return NULL;
  fseek(f, 0, SEEK_END);
  long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  uint8_t* blob = (uint8_t*)malloc((size_t)sz + 16);
  if (!blob || fread(blob, 1, (size_t)sz, f) != (size_t)sz) { fclose(f); free(blob); return NULL; }
  fclose(f);
  PzkRefProgram* p = (PzkRefProgram*)calloc(1, sizeof *p);
  memcpy(&p->h, blob, sizeof p->h);
  if (p->h.magic != PZK_MAGIC || p->h.version != PZK_VERSION) { free(blob); free(p); return NULL; }
  p->blob = blob;
  uint64_t pos = align16(sizeof(PzkHeader));
  p->segs = (PzkSegment*)(blob + pos); pos = align16(pos + p->h.n_segments * sizeof(PzkSegment));
  p->ops = (PzkOp*)(blob + pos); pos = align16(pos + p->h.n_op_records * sizeof(PzkOp));
  p->fpool = (uint64_t*)(blob + pos); pos = align16(pos + (uint64_t)p->h.n_fpool * 32);
  p->coefs = (PzkCoef*)(blob + pos); pos = align16(pos + (uint64_t)p->h.n_coef * sizeof(PzkCoef));
  p->list = (uint32_t*)(blob + pos); pos = align16(pos + (uint64_t)p->h.n_list * 4);
  p->inputs = (PzkInput*)(blob + pos); pos = align16(pos + (uint64_t)p->h.n_inputs * sizeof(PzkInput));
  p->rows = (PzkRow*)(blob + pos); pos = align16(pos + p->h.n_rows * sizeof(PzkRow));
  p->terms = (PzkTerm*)(blob + pos); pos = align16(pos + p->h.n_terms * sizeof(PzkTerm));
  p->exports = (PzkExport*)(blob + pos); pos = align16(pos + p->h.n_exports * sizeof(PzkExport));
  p->meta = (char*)(blob + pos);
  return p;
}
void pzk_ref_free(PzkRefProgram* p) { if (p) { free(p->blob); free(p); } }
uint32_t pzk_ref_n_wires(const PzkRefProgram* p) { return p->h.n_wires; }
uint32_t pzk_ref_n_inputs(const PzkRefProgram* p) { return p->h.n_inputs; }
uint32_t pzk_ref_n_constraints(const PzkRefProgram* p) { return p->h.n_constraints; }
uint32_t pzk_ref_n_outputs(const PzkRefProgram* p) { return p->h.n_pub_out; }
const char* pzk_ref_meta(const PzkRefProgram* p, uint64_t* len) { *len = p->h.reserved[0]; return p->meta; }

/* ---- 256-bit helpers ------------------------------------------------------ */
static int cmp4(const uint64_t* a, const uint64_t* b) {
  for (int i = 3; i >= 0; i--) { if (a[i] < b[i]) return -1; if (a[i] > b[i]) return 1; }
  return 0;
}
static uint64_t add4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  u128 c = 0;
  for (int i = 0; i < 4; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
  return (uint64_t)c;
}
static uint64_t sub4(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t br = 0;
  for (int i = 0; i < 4; i++) { u128 t = (u128)a[i] - b[i] - br; r[i] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
  return br;
}
static int is_zero4(const uint64_t* a) { return (a[0] | a[1] | a[2] | a[3]) == 0; }
static void fadd(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t[4]; uint64_t c = add4(t, a, b);
  if (c || cmp4(t, P) >= 0) sub4(t, t, P);
  memcpy(r, t, 32);
}
static void fsub(uint64_t* r, const uint64_t* a, const uint64_t* b) {
  uint64_t t[4];
  if (sub4(t, a, b)) add4(t, t, P);
  memcpy(r, t, 32);
}
static void fmul(uint64_t* r, const uint64_t* a, const uint64_t* b) { /* a*b/R mod p */
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) { c += (u128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[4] = (uint64_t)c; t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * PINV;
    c = (u128)m * P[0] + t[0]; c >>= 64;
    for (int j = 1; j < 4; j++) { c += (u128)m * P[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
    c += t[4]; t[3] = (uint64_t)c; t[4] = t[5] + (uint64_t)(c >> 64);
  }
  if (t[4] || cmp4(t, P) >= 0) sub4(t, t, P);
  memcpy(r, t, 32);
}
static void to_mont(uint64_t* r, const uint64_t* a) { fmul(r, a, R2); }
static void from_mont(uint64_t* r, const uint64_t* a) { fmul(r, a, ONE); }
static void finv(uint64_t* r, const uint64_t* a) { /* a^(p-2), Montgomery in / out; inv(0)=0 */
  uint64_t e[4], acc[4], base[4];
  static const uint64_t two[4] = {2, 0, 0, 0};
  sub4(e, P, two);
  to_mont(acc, ONE);
  memcpy(base, a, 32);
  for (int i = 0; i < 256; i++) {
    if ((e[i >> 6] >> (i & 63)) & 1) fmul(acc, acc, base);
    fmul(base, base, base);
  }
  memcpy(r, acc, 32);
}
static void shr4(uint64_t* r, const uint64_t* a, unsigned s) {
  uint64_t t[4] = {0, 0, 0, 0};
  if (s < 256) {
    unsigned ws = s >> 6, bs = s & 63;
    for (unsigned i = 0; i + ws < 4; i++) {
      t[i] = a[i + ws] >> bs;
      if (bs && i + ws + 1 < 4) t[i] |= a[i + ws + 1] << (64 - bs);
    }
  }
  memcpy(r, t, 32);
}
static void shl4(uint64_t* r, const uint64_t* a, unsigned s) {
  uint64_t t[4] = {0, 0, 0, 0};
  if (s < 256) {
    unsigned ws = s >> 6, bs = s & 63;
    for (int i = 3; i >= (int)ws; i--) {
      t[i] = a[i - ws] << bs;
      if (bs && i - (int)ws - 1 >= 0) t[i] |= a[i - ws - 1] >> (64 - bs);
    }
  }
  memcpy(r, t, 32);
}
static void divmod4(const uint64_t* a, const uint64_t* b, uint64_t* q, uint64_t* m) {
  uint64_t qq[4] = {0, 0, 0, 0}, rr[4] = {0, 0, 0, 0};
  if (!is_zero4(b))
    for (int i = 255; i >= 0; i--) {
      shl4(rr, rr, 1);
      rr[0] |= (a[i >> 6] >> (i & 63)) & 1;
      if (cmp4(rr, b) >= 0) { sub4(rr, rr, b); qq[i >> 6] |= 1ull << (i & 63); }
    }
  if (q) memcpy(q, qq, 32);
  if (m) memcpy(m, rr, 32);
}
static void reduce_p(uint64_t* a) { while (cmp4(a, P) >= 0) sub4(a, a, P); }
/* arithmetic modulo an arbitrary m < 2^256 (m may exceed 2^255: carries out of 256 bits are tracked) */
static void addmod4(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* m) { /* a, b < m */
  uint64_t t[4];
  uint64_t c = add4(t, a, b);
  if (c || cmp4(t, m) >= 0) sub4(t, t, m);
  memcpy(r, t, 32);
}
static void mod4(uint64_t* r, const uint64_t* a, const uint64_t* m) {
  uint64_t rr[4] = {0, 0, 0, 0};
  for (int i = 255; i >= 0; i--) {
    uint64_t bit[4] = {(a[i >> 6] >> (i & 63)) & 1, 0, 0, 0};
    addmod4(rr, rr, rr, m);
    if (bit[0] && cmp4(bit, m) < 0) addmod4(rr, rr, bit, m);
  }
  memcpy(r, rr, 32);
}
static void mulmod4(uint64_t* r, const uint64_t* a, const uint64_t* b, const uint64_t* m) { /* a, b < m */
  uint64_t acc[4] = {0, 0, 0, 0}, x[4];
  memcpy(x, a, 32);
  for (int i = 0; i < 256; i++) {
    if ((b[i >> 6] >> (i & 63)) & 1) addmod4(acc, acc, x, m);
    addmod4(x, x, x, m);
  }
  memcpy(r, acc, 32);
}
static int is_neg(const uint64_t* a) { /* a > p/2 */
  uint64_t h[4]; shr4(h, P, 1);
  return cmp4(a, h) > 0;
}
/* ---- Z class (pzk_program.h): exact signed integers in 256-bit two's complement ---- */
static void z_canonical(uint64_t* r, const uint64_t* z) { /* v < 0 ? p + v : v */
  memcpy(r, z, 32);
  if ((int64_t)z[3] < 0) add4(r, r, P);
}
static void z_to_mont(uint64_t* r, const uint64_t* z) {
  uint64_t m[4], zero[4] = {0, 0, 0, 0};
  int neg = (int64_t)z[3] < 0;
  if (neg) sub4(m, zero, z); else memcpy(m, z, 32);
  to_mont(r, m);
  if (neg) fsub(r, zero, r);
}
static void z_mul(uint64_t* r, const uint64_t* a, const uint64_t* b) { /* a * b mod 2^256 */
  uint64_t t[4] = {0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    uint64_t carry = 0;
    for (int j = 0; i + j < 4; j++) {
      u128 x = (u128)a[i] * b[j] + t[i + j] + carry;
      t[i + j] = (uint64_t)x; carry = (uint64_t)(x >> 64);
    }
  }
  memcpy(r, t, 32);
}

static int scmp(const uint64_t* a, const uint64_t* b) {
  int na = is_neg(a), nb = is_neg(b);
  if (na != nb) return na ? -1 : 1;
  return cmp4(a, b);
}

/* ---- long division of (k+m) n-bit limbs by k n-bit limbs, bit serial ------- */
static int bigdiv(unsigned n, unsigned k, unsigned m, const uint64_t* a, const uint64_t* b, uint64_t* q,
                  uint64_t* r) {
  /* pack into 64-bit words (n <= 64) */
  enum { W = 200 };
  uint64_t A[W] = {0}, B[W] = {0}, Q[W] = {0}, Rm[W + 1] = {0};
  unsigned abits = n * (k + m), bbits = n * k;
  if ((abits + 63) / 64 + 1 > W) return -1;
  for (unsigned i = 0; i < k + m; i++)
    for (unsigned j = 0; j < n; j++)
      if ((a[i] >> j) & 1) { unsigned pos = i * n + j; A[pos >> 6] |= 1ull << (pos & 63); }
  for (unsigned i = 0; i < k; i++)
    for (unsigned j = 0; j < n; j++)
      if ((b[i] >> j) & 1) { unsigned pos = i * n + j; B[pos >> 6] |= 1ull << (pos & 63); }
  if (b[k - 1] == 0) return 1; /* precondition of the circom function */
  unsigned bw = (bbits + 63) / 64 + 1;
  for (int i = (int)abits - 1; i >= 0; i--) {
    /* Rm = (Rm << 1) | bit */
    uint64_t carry = (A[i >> 6] >> (i & 63)) & 1;
    for (unsigned w = 0; w < bw; w++) { uint64_t nc = Rm[w] >> 63; Rm[w] = (Rm[w] << 1) | carry; carry = nc; }
    int ge = 1;
    for (int w = (int)bw - 1; w >= 0; w--) {
      if (Rm[w] > B[w]) break;
      if (Rm[w] < B[w]) { ge = 0; break; }
    }
    if (ge) {
      uint64_t br = 0;
      for (unsigned w = 0; w < bw; w++) { u128 t = (u128)Rm[w] - B[w] - br; Rm[w] = (uint64_t)t; br = (uint64_t)(t >> 64) & 1; }
      Q[i >> 6] |= 1ull << (i & 63);
    }
  }
  uint64_t mask = n == 64 ? ~0ull : ((1ull << n) - 1);
  for (unsigned i = 0; i <= m; i++) {
    uint64_t v = 0;
    for (unsigned j = 0; j < n; j++) { unsigned pos = i * n + j; v |= ((Q[pos >> 6] >> (pos & 63)) & 1) << j; }
    q[i] = v & mask;
  }
  for (unsigned i = 0; i < k; i++) {
    uint64_t v = 0;
    for (unsigned j = 0; j < n; j++) { unsigned pos = i * n + j; v |= ((Rm[pos >> 6] >> (pos & 63)) & 1) << j; }
    r[i] = v & mask;
  }
  return 0;
}

/* ---- constraint rows -------------------------------------------------------- */
static void term_value(const PzkRefProgram* p, const PzkTerm* t, const uint64_t* U, const uint64_t* F, uint64_t* out) {
  const PzkCoef* c = &p->coefs[t->coef];
  if (t->ref == PZK_REF_ONE) { memcpy(out, c->mont, 32); return; }
  uint32_t cls = PZK_REF_CLS(t->ref), slot = PZK_REF_SLOT(t->ref);
  if (cls == 2) { fmul(out, c->mont, F + 4 * (uint64_t)slot); return; }
  uint64_t v = U[slot];
  if (cls == 1 && (int64_t)v < 0) {
    uint64_t w[4] = {(uint64_t)(-(int64_t)v), 0, 0, 0}, z[4] = {0, 0, 0, 0};
    fmul(out, c->mont2, w);
    fsub(out, z, out);
    return;
  }
  uint64_t w[4] = {v, 0, 0, 0};
  fmul(out, c->mont2, w);
}
static void lin_value(const PzkRefProgram* p, const PzkTerm* t, unsigned n, const uint64_t* U, const uint64_t* F,
                      uint64_t* acc) {
  memset(acc, 0, 32);
  for (unsigned i = 0; i < n; i++) { uint64_t v[4]; term_value(p, t + i, U, F, v); fadd(acc, acc, v); }
}

/*
 * One witness.  inputs: n_inputs x 32 bytes little endian, canonical values in main-input wire order.
 * witness (may be NULL): n_wires x 32 bytes little endian canonical.
 * returns the lane status bits; *first_bad = index of the first failing constraint (or -1).
 */
uint32_t pzk_ref_witness(const PzkRefProgram* p, const uint8_t* inputs, uint8_t* witness, int64_t* first_bad,
                         int check_rows) {
  uint64_t* U = (uint64_t*)calloc((size_t)p->h.n_u_slots + 1, 8);
  uint64_t* F = (uint64_t*)calloc((size_t)p->h.n_f_slots + 1, 32);
  uint64_t cells[1100]; /* the shared-memory operand cache of the device, one lane */
  memset(cells, 0, sizeof cells);
  uint32_t status = 0;
  int64_t bad = -1;
  if (witness) { memset(witness, 0, (size_t)p->h.n_wires * 32); witness[0] = 1; }
  for (uint32_t s = 0; s < p->h.n_segments; s++) {
    const PzkSegment* sg = &p->segs[s];
    for (uint64_t pc = sg->op_off; pc < sg->op_off + sg->n_ops; pc++) {
      const PzkOp* o = &p->ops[pc];
      const PzkOpExt* x = (const PzkOpExt*)(o + 1);
      if (o->flags & PZK_FLAG_EXT) pc++;
/* operand words: bit 31 -> cache cell, else global slot; dst words: slot | (cell + 1) << 22 */
#define RDU(x) (((x) & PZK_OPERAND_CELL) ? cells[(x) & 0xffffu] : U[x])
#define RDF(x) (((x) & PZK_OPERAND_CELL) ? (cells + ((x) & 0xffffu)) : (F + 4 * (uint64_t)(x)))
#define WRU(d, v) do { uint64_t v__ = (v); U[PZK_DST_SLOT(d)] = v__; if (PZK_DST_CELL(d)) cells[PZK_DST_CELL(d) - 1] = v__; } while (0)
#define WRF(d, src) do { uint64_t t__[4]; memcpy(t__, (src), 32); memcpy(F + 4 * (uint64_t)PZK_DST_SLOT(d), t__, 32); \
                         if (PZK_DST_CELL(d)) memcpy(cells + PZK_DST_CELL(d) - 1, t__, 32); } while (0)
#define UA RDU(o->a)
#define UB ((o->flags & PZK_FLAG_B_IMM) ? (uint64_t)o->b : RDU(o->b))
#define FA RDF(o->a)
#define FB ((o->flags & PZK_FLAG_B_POOL) ? (p->fpool + 4 * (uint64_t)o->b) : RDF(o->b))
      /* Z_MUL / Z_MULADD: a factor may be a U word read in place (PZK_FLAG_A_U / PZK_FLAG_B_U) */
      uint64_t zt_a[4] = {0, 0, 0, 0}, zt_b[4] = {0, 0, 0, 0};
#define ZFA ((o->flags & PZK_FLAG_A_U) ? (zt_a[0] = RDU(o->a), (const uint64_t*)zt_a) : (const uint64_t*)FA)
#define ZFB ((o->flags & PZK_FLAG_B_U) ? (zt_b[0] = RDU(o->b), (const uint64_t*)zt_b) : (const uint64_t*)FB)
      switch (o->opc) {
        case PZK_NOP: break;
        case PZK_U_CONST: WRU(o->dst, ((uint64_t)o->b << 32) | o->a); break;
        case PZK_U_ADD: WRU(o->dst, UA + UB); break;
        case PZK_U_SHLADD: WRU(o->dst, UA + (RDU(o->b) << o->imm16)); break;
        case PZK_U_SUB: WRU(o->dst, UA - UB); break;
        case PZK_U_MUL: WRU(o->dst, UA * UB); break;
        case PZK_U_DIV: { uint64_t b = UB; WRU(o->dst, b ? UA / b : 0); break; }
        case PZK_U_MOD: { uint64_t b = UB; WRU(o->dst, b ? UA % b : 0); break; }
        case PZK_U_SHR: { uint64_t b = UB; WRU(o->dst, b >= 64 ? 0 : UA >> b); break; }
        case PZK_U_SHL: { uint64_t b = UB; WRU(o->dst, b >= 64 ? 0 : UA << b); break; }
        case PZK_U_AND: WRU(o->dst, UA & UB); break;
        case PZK_U_OR: WRU(o->dst, UA | UB); break;
        case PZK_U_XOR: WRU(o->dst, UA ^ UB); break;
        case PZK_U_LT: WRU(o->dst, UA < UB); break;
        case PZK_U_LE: WRU(o->dst, UA <= UB); break;
        case PZK_U_EQ: WRU(o->dst, UA == UB); break;
        case PZK_U_NE: WRU(o->dst, UA != UB); break;
        case PZK_I_LT: WRU(o->dst, (int64_t)UA < (int64_t)UB); break;
        case PZK_I_LE: WRU(o->dst, (int64_t)UA <= (int64_t)UB); break;
        case PZK_U_SEL: WRU(o->dst, UA ? RDU(o->b) : RDU(x->c)); break;
        case PZK_U_LUT: case PZK_U_LUTV: {
          /* operand j contributes bit (ext.f >> 8j) & 255 of its word */
          unsigned idx = 0;
          if (o->a != PZK_OPERAND_NONE) idx |= (unsigned)((RDU(o->a) >> (x->f & 255)) & 1);
          if (o->b != PZK_OPERAND_NONE) idx |= (unsigned)((RDU(o->b) >> ((x->f >> 8) & 255)) & 1) << 1;
          if (x->c != PZK_OPERAND_NONE) idx |= (unsigned)((RDU(x->c) >> ((x->f >> 16) & 255)) & 1) << 2;
          if (x->d != PZK_OPERAND_NONE) idx |= (unsigned)((RDU(x->d) >> ((x->f >> 24) & 255)) & 1) << 3;
          if (o->opc == PZK_U_LUT) WRU(o->dst, (o->imm16 >> idx) & 1);
          else WRU(o->dst, (uint64_t)p->list[x->e + 2 * idx] | ((uint64_t)p->list[x->e + 2 * idx + 1] << 32));
          break;
        }
        case PZK_V_LUT: {
          /* packed truth table, one lane at a time: bit l of operand j = bit (l + rot_j) mod w of its word */
          unsigned w = (o->flags & PZK_FLAG_W64) ? 64 : 32;
          uint64_t lanes = x->f;
          if (w == 64) { lanes |= (uint64_t)((const PzkOpExt*)(o + 2))->c << 32; pc++; }
          uint32_t refs[4] = {o->a, o->b, x->c, x->d};
          uint64_t words[4] = {0, 0, 0, 0};
          for (int j = 0; j < 4; j++) if (refs[j] != PZK_OPERAND_NONE) words[j] = RDU(refs[j]);
          uint64_t r = 0;
          for (unsigned l = 0; l < w; l++) {
            if (!((lanes >> l) & 1)) continue;
            unsigned idx = 0;
            for (int j = 0; j < 4; j++) {
              if (refs[j] == PZK_OPERAND_NONE) continue;
              unsigned rot = (x->e >> (8 * j)) & 255;
              idx |= (unsigned)((words[j] >> ((l + rot) % w)) & 1) << j;
            }
            r |= (uint64_t)((o->imm16 >> idx) & 1) << l;
          }
          WRU(o->dst, r);
          break;
        }
        case PZK_U_EXTRACT: case PZK_N_EXTRACT: {
          /* ((a >> s) & (2^n - 1)) << k on 256 bits */
          unsigned s_ = o->imm16 & 255, k_ = o->imm16 >> 8, n_ = o->b;
          uint64_t v[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0}, one[4] = {1, 0, 0, 0};
          if (o->flags & PZK_FLAG_NBASE) memcpy(v, RDF(o->a), 32); else v[0] = RDU(o->a);
          shr4(v, v, s_);
          if (n_ < 256) { shl4(m, one, n_); sub4(m, m, one); for (int i = 0; i < 4; i++) v[i] &= m[i]; }
          shl4(v, v, k_);
          if (o->opc == PZK_U_EXTRACT) WRU(o->dst, v[0]); else WRF(o->dst, v);
          break;
        }
        case PZK_CHECK_RANGE: {
          if (check_rows) {
            uint64_t v[4] = {0, 0, 0, 0};
            if (o->flags & PZK_FLAG_NBASE) memcpy(v, RDF(o->a), 32); else v[0] = RDU(o->a);
            shr4(v, v, o->imm16);
            if (!is_zero4(v)) { status |= PZK_LANE_CONSTRAINT; if (bad < 0 || (int64_t)o->dst < bad) bad = o->dst; }
          }
          break;
        }
        case PZK_F_CONST: WRF(o->dst, p->fpool + 4 * (uint64_t)o->a); break;
        case PZK_F_ADD: { uint64_t r[4]; fadd(r, FA, FB); WRF(o->dst, r); break; }
        case PZK_F_SUB: { uint64_t r[4]; fsub(r, FA, FB); WRF(o->dst, r); break; }
        case PZK_F_MUL: { uint64_t r[4]; fmul(r, FA, FB); WRF(o->dst, r); break; }
        case PZK_F_NEG: { uint64_t z[4] = {0, 0, 0, 0}, r[4]; fsub(r, z, FA); WRF(o->dst, r); break; }
        case PZK_F_INV: { uint64_t r[4]; finv(r, FA); WRF(o->dst, r); break; }
        case PZK_F_FROM_U: { uint64_t w[4] = {UA, 0, 0, 0}, r[4]; to_mont(r, w); WRF(o->dst, r); break; }
        case PZK_F_FROM_I: {
          int64_t v = (int64_t)UA;
          uint64_t w[4] = {v < 0 ? (uint64_t)(-v) : (uint64_t)v, 0, 0, 0}, z[4] = {0, 0, 0, 0}, r[4];
          to_mont(w, w);
          if (v < 0) fsub(r, z, w); else memcpy(r, w, 32);
          WRF(o->dst, r);
          break;
        }
        case PZK_F_SEL: WRF(o->dst, UA ? RDF(o->b) : RDF(x->c)); break;
        case PZK_F_EQ: WRU(o->dst, memcmp(FA, FB, 32) == 0); break;
        case PZK_F_NE: WRU(o->dst, memcmp(FA, FB, 32) != 0); break;
        case PZK_F_CSEL: WRF(o->dst, p->fpool + 4 * ((uint64_t)o->b + UA)); break;
        case PZK_N_FROM_F: { uint64_t r[4]; if (o->flags & PZK_FLAG_ZSRC) z_canonical(r, FA); else from_mont(r, FA); WRF(o->dst, r); break; }
        case PZK_F_FROM_N: {
          uint64_t t[4], r[4];
          if (o->flags & PZK_FLAG_ZSRC) z_to_mont(r, FA); else { memcpy(t, FA, 32); reduce_p(t); to_mont(r, t); }
          WRF(o->dst, r); break;
        }
        case PZK_Z_ADD: { uint64_t r[4]; add4(r, FA, FB); WRF(o->dst, r); break; }
        case PZK_Z_SUB: { uint64_t r[4]; sub4(r, FA, FB); WRF(o->dst, r); break; }
        case PZK_Z_MUL: { uint64_t r[4]; z_mul(r, ZFA, ZFB); WRF(o->dst, r); break; }
        case PZK_F_MULADD: { /* +-(a b) +- c: imm16 bit 8 negates the product, bit 9 negates c */
          uint64_t m[4], r[4];
          fmul(m, FA, FB);
          if (o->flags & PZK_FLAG_DST2) memcpy(F + 4 * (uint64_t)PZK_DST_SLOT(x->d), m, 32); /* the product is a wire */
          if (o->imm16 & 0x100) fsub(r, RDF(x->c), m); else if (o->imm16 & 0x200) fsub(r, m, RDF(x->c)); else fadd(r, m, RDF(x->c));
          WRF(o->dst, r); break;
        }
        case PZK_Z_MULADD: {
          uint64_t m[4], r[4];
          z_mul(m, ZFA, ZFB);
          if (o->flags & PZK_FLAG_DST2) memcpy(F + 4 * (uint64_t)PZK_DST_SLOT(x->d), m, 32);
          if (o->imm16 & 0x100) sub4(r, RDF(x->c), m); else if (o->imm16 & 0x200) sub4(r, m, RDF(x->c)); else add4(r, m, RDF(x->c));
          WRF(o->dst, r); break;
        }
        case PZK_Z_FROM_U: { uint64_t w[4] = {UA, 0, 0, 0}; WRF(o->dst, w); break; }
        case PZK_Z_FROM_I: { uint64_t x_ = UA, sx = (uint64_t)((int64_t)x_ >> 63); uint64_t w[4] = {x_, sx, sx, sx}; WRF(o->dst, w); break; }
        case PZK_Z_CONST: WRF(o->dst, p->fpool + 4 * (uint64_t)o->a); break;
        case PZK_N_FROM_U: { uint64_t w[4] = {UA, 0, 0, 0}; WRF(o->dst, w); break; }
        case PZK_N_BIT: WRU(o->dst, o->b < 256 ? (FA[o->b >> 6] >> (o->b & 63)) & 1 : 0); break;
        case PZK_N_LOW: WRU(o->dst, FA[0]); break;
        case PZK_N_FITS: WRU(o->dst, (FA[1] | FA[2] | FA[3]) == 0); break;
        case PZK_N_SHR: { uint64_t b = UB; uint64_t t[4]; shr4(t, FA, b > 256 ? 256 : (unsigned)b); WRF(o->dst, t); break; }
        case PZK_N_SHL: {
          uint64_t b = UB; uint64_t t[4] = {0, 0, 0, 0};
          if (b < 254) { shl4(t, FA, (unsigned)b); t[3] &= 0x3fffffffffffffffull; reduce_p(t); }
          WRF(o->dst, t); break;
        }
        case PZK_N_AND: { const uint64_t* b = FB; uint64_t t[4]; for (int i = 0; i < 4; i++) t[i] = FA[i] & b[i]; reduce_p(t); WRF(o->dst, t); break; }
        case PZK_N_OR: { const uint64_t* b = FB; uint64_t t[4]; for (int i = 0; i < 4; i++) t[i] = FA[i] | b[i]; t[3] &= 0x3fffffffffffffffull; reduce_p(t); WRF(o->dst, t); break; }
        case PZK_N_XOR: { const uint64_t* b = FB; uint64_t t[4]; for (int i = 0; i < 4; i++) t[i] = FA[i] ^ b[i]; t[3] &= 0x3fffffffffffffffull; reduce_p(t); WRF(o->dst, t); break; }
        case PZK_N_DIV: { uint64_t q[4]; divmod4(FA, FB, q, NULL); WRF(o->dst, q); break; }
        case PZK_N_MOD: { uint64_t m[4]; divmod4(FA, FB, NULL, m); WRF(o->dst, m); break; }
        case PZK_N_SLT: WRU(o->dst, scmp(FA, FB) < 0); break;
        case PZK_N_SLE: WRU(o->dst, scmp(FA, FB) <= 0); break;
        case PZK_BIGDIV: {
          const uint32_t* L = p->list + o->a;
          unsigned n = L[0], k = L[1], m = L[2];
          uint64_t a[200], b[200], q[200], r[200];
          for (unsigned i = 0; i < k + m; i++) a[i] = U[L[3 + i]];
          for (unsigned i = 0; i < k; i++) b[i] = U[L[3 + k + m + i]];
          int rc = bigdiv(n, k, m, a, b, q, r);
          if (rc) { status |= PZK_LANE_BIGDIV_PRE; memset(q, 0, sizeof q); memset(r, 0, sizeof r); }
          for (unsigned i = 0; i <= m; i++) U[L[3 + k + m + k + i]] = q[i];
          for (unsigned i = 0; i < k; i++) U[L[3 + k + m + k + m + 1 + i]] = r[i];
          break;
        }
        case PZK_MODINV: {
          /* mod_inv of bigIntFunc.circom:430-465 as the reference defines it: 0 when a == 0 (mod p),
           * else a^(p-2) mod p by square-and-multiply - deliberately NOT the extended GCD the device uses */
          const uint32_t* L = p->list + o->a;
          unsigned nb = L[0], k = L[1];   /* k limbs of nb bits */
          uint64_t a[4] = {0, 0, 0, 0}, m[4] = {0, 0, 0, 0}, e[4], r[4] = {1, 0, 0, 0}, two[4] = {2, 0, 0, 0};
          for (unsigned i = 0; i < k; i++) {
            unsigned pos = i * nb;
            a[pos >> 6] |= U[L[3 + i]] << (pos & 63);
            m[pos >> 6] |= U[L[3 + k + i]] << (pos & 63);
          }
          mod4(a, a, m);
          sub4(e, m, two);
          mod4(r, r, m);
          for (int i = 255; i >= 0; i--) {
            mulmod4(r, r, r, m);
            if ((e[i >> 6] >> (i & 63)) & 1) mulmod4(r, r, a, m);
          }
          if (!(a[0] | a[1] | a[2] | a[3])) memset(r, 0, sizeof r);
          U[L[3 + k + k]] = 0;
          for (unsigned i = 0; i < k; i++) {
            unsigned pos = i * nb;
            U[L[3 + k + k + 1 + i]] = (r[pos >> 6] >> (pos & 63)) & (nb >= 64 ? ~0ull : ((1ull << nb) - 1));
          }
          break;
        }
        case PZK_BJJ_MUL8: {
          /* hints of BabyjubjubBase8Multiplication as the reference computes them: affine BabyjubjubAdd
           * (/root/reference/circuits/lib/circuits/babyjubjub/curve.circom:71-105) with two field divisions each,
           * x / 0 = 0, chained through addZeroBabyjub's selection (:19-58) - deliberately NOT the projective
           * ladder + batched inversion of the device */
          const uint32_t* L = p->list + o->a;
          unsigned n = L[0];
          const uint64_t *ca = p->fpool + 4 * (uint64_t)L[1], *cd = p->fpool + 4 * (uint64_t)L[2];
          const uint64_t *bx = p->fpool + 4 * (uint64_t)L[3], *by = p->fpool + 4 * (uint64_t)L[4];
          const uint32_t* out = L + 6;
          uint64_t sc[4], one[4], S[2][4] = {{0}}, D[2][4], Q[2][4], A[2][4];
          from_mont(sc, F + 4 * (uint64_t)L[5]);
          to_mont(one, ONE);
          unsigned k = 0;
          for (unsigned i = 0; i < n; i++) {
            for (int pass = (i > 0 ? 0 : 1); pass < 2; pass++) {
              const uint64_t (*P1)[4] = pass == 0 ? S : D;
              const uint64_t (*P2)[4] = pass == 0 ? S : Q;
              if (pass == 1) {
                unsigned bi = n - 1 - i;
                int bit = (int)((sc[bi >> 6] >> (bi & 63)) & 1);
                memset(Q, 0, sizeof Q);
                if (bit) { memcpy(Q[0], bx, 32); memcpy(Q[1], by, 32); }
                if (i == 0) memset(D, 0, sizeof D);
              }
              uint64_t beta[4], gamma[4], delta[4], tau[4], t[4], u[4], den[4], num[4], r0[4], r1[4];
              fmul(beta, P1[0], P2[1]); fmul(gamma, P1[1], P2[0]);
              fmul(t, P1[0], ca); fsub(t, P1[1], t); fadd(u, P2[0], P2[1]); fmul(delta, t, u);
              fmul(tau, beta, gamma);
              fmul(t, tau, cd);
              fadd(den, one, t); finv(den, den); fadd(num, beta, gamma); fmul(r0, num, den);
              fsub(den, one, t); finv(den, den); fmul(u, beta, ca); fadd(num, delta, u); fsub(num, num, gamma); fmul(r1, num, den);
              memcpy(F + 4 * (uint64_t)out[2 * k], r0, 32); memcpy(F + 4 * (uint64_t)out[2 * k + 1], r1, 32);
              k++;
              if (pass == 0) { memcpy(D[0], r0, 32); memcpy(D[1], r1, 32); } else { memcpy(A[0], r0, 32); memcpy(A[1], r1, 32); }
            }
            int zd = is_zero4(D[0]), zq = is_zero4(Q[0]);
            memcpy(S, zd ? Q : (zq ? D : A), sizeof S);
          }
          break;
        }
        case PZK_CHECK_I64: case PZK_CHECK_INT: case PZK_CHECK_F: {
          /* a constraint row fused into the op stream: A.w * B.w == C.w, checked here with
           * the generic field arithmetic for both kinds (the integer fast path is a device-side
           * optimisation whose result must agree) */
          unsigned na = o->imm16, nb = o->a & 0xffffu, nc = o->a >> 16, nrec = o->b;
          if (check_rows) {
            const uint32_t* tw = (const uint32_t*)(o + 1);
            uint64_t acc[3][4];
            unsigned lens[3] = {na, nb, nc}, k = 0;
            for (int part = 0; part < 3; part++) {
              memset(acc[part], 0, 32);
              for (unsigned i = 0; i < lens[part]; i++, k++) {
                uint32_t ref = tw[2 * k], cw = tw[2 * k + 1];
                uint64_t cm[4], v[4], t[4];
                if (o->opc != PZK_CHECK_F) {
                  int64_t c;
                  if (ref == PZK_REF_ONE_LIST || (ref < PZK_REF_ONE_LIST && (ref & PZK_TERM_COEF_LIST)))
                    c = (int64_t)((uint64_t)p->list[cw] | ((uint64_t)p->list[cw + 1] << 32));
                  else c = (int32_t)cw;
                  uint64_t m[4] = {c < 0 ? (uint64_t)(-c) : (uint64_t)c, 0, 0, 0}, z[4] = {0, 0, 0, 0};
                  to_mont(cm, m);
                  if (c < 0) fsub(cm, z, cm);
                } else memcpy(cm, p->coefs[cw].mont, 32);
                if (ref >= PZK_REF_ONE_LIST) { fadd(acc[part], acc[part], cm); continue; }
                uint32_t cls = PZK_REF_CLS(ref), slot = PZK_REF_SLOT(ref);
                int in_cell = (ref & PZK_TERM_CELL) != 0;
                if (cls == 2) memcpy(v, in_cell ? (cells + (ref & 0xffffu)) : (F + 4 * (uint64_t)slot), 32);
                else if (cls == 3) z_to_mont(v, in_cell ? (cells + (ref & 0xffffu)) : (F + 4 * (uint64_t)slot));
                else {
                  uint64_t raw = in_cell ? cells[ref & 0xffffu] : U[slot];
                  int neg = cls == 1 && (int64_t)raw < 0;
                  uint64_t m[4] = {neg ? (uint64_t)(-(int64_t)raw) : raw, 0, 0, 0}, z[4] = {0, 0, 0, 0};
                  to_mont(v, m);
                  if (neg) fsub(v, z, v);
                }
                fmul(t, cm, v);
                fadd(acc[part], acc[part], t);
              }
            }
            uint64_t ab[4];
            fmul(ab, acc[0], acc[1]);
            if (memcmp(ab, acc[2], 32) != 0) {
              status |= PZK_LANE_CONSTRAINT;
              if (bad < 0 || (int64_t)o->dst < bad) bad = o->dst;
            }
          }
          pc += nrec;
          break;
        }
        case PZK_ASSERT_NZ: if (UA == 0) status |= PZK_LANE_ASSERT; break;
        case PZK_IN_U: {
          uint64_t w[4]; memcpy(w, inputs + 32 * (uint64_t)o->a, 32);
          int bits = o->imm16;
          if (w[1] | w[2] | w[3] || (bits < 64 && (w[0] >> bits))) status |= PZK_LANE_INPUT_RANGE;
          WRU(o->dst, w[0]);
          break;
        }
        case PZK_IN_F: {
          uint64_t w[4], r[4]; memcpy(w, inputs + 32 * (uint64_t)o->a, 32);
          if (cmp4(w, P) >= 0) { status |= PZK_LANE_INPUT_RANGE; reduce_p(w); }
          to_mont(r, w);
          WRF(o->dst, r);
          break;
        }
        default: fprintf(stderr, "ssa_ref: bad opcode %d\n", o->opc); abort();
      }
      if (o->flags & PZK_FLAG_DIG) pc++; /* digest descriptor of the device's fused witness digest: not an op */
      if ((o->opc == PZK_F_MULADD || o->opc == PZK_Z_MULADD) && (o->flags & PZK_FLAG_DIG2)) pc++; /* the product's */
    }
    if (check_rows) {
      for (uint64_t r = sg->row_off; r < sg->row_off + sg->n_rows; r++) {
        const PzkRow* row = &p->rows[r];
        const PzkTerm* t = p->terms + row->term_off;
        uint64_t a[4], b[4], c[4], ab[4];
        lin_value(p, t, row->na, U, F, a);
        lin_value(p, t + row->na, row->nb, U, F, b);
        lin_value(p, t + row->na + row->nb, row->nc, U, F, c);
        fmul(ab, a, b);
        if (memcmp(ab, c, 32) != 0) {
          status |= PZK_LANE_CONSTRAINT;
          if (bad < 0 || (int64_t)row->index < bad) bad = row->index;
        }
      }
    }
    if (witness) {
      for (uint64_t e = sg->exp_off; e < sg->exp_off + sg->n_exp; e++) {
        const PzkExport* ex = &p->exports[e];
        uint8_t* dst = witness + 32 * (uint64_t)ex->wire;
        if (ex->ref == PZK_REF_ZERO) continue;
        uint32_t cls = PZK_REF_CLS(ex->ref), slot = PZK_REF_SLOT(ex->ref);
        uint64_t w[4] = {0, 0, 0, 0};
        if (ex->ref == PZK_REF_TABVIEW) {
          /* truth-table function of bits of words: {n, (slot, pos) x n, 2^n x int64} in the list pool */
          const uint32_t* L = p->list + ex->aux;
          unsigned n = L[0], idx = 0;
          for (unsigned j = 0; j < n; j++) idx |= (unsigned)((U[L[1 + 2 * j]] >> L[2 + 2 * j]) & 1) << j;
          int64_t v = (int64_t)((uint64_t)L[1 + 2 * n + 2 * idx] | ((uint64_t)L[2 + 2 * n + 2 * idx] << 32));
          if (v < 0) { uint64_t m[4] = {(uint64_t)(-v), 0, 0, 0}; sub4(w, P, m); } else w[0] = (uint64_t)v;
        } else if (cls == 3) {
          /* bit-field view: ((word >> s) & (2^n - 1)) << k */
          unsigned s_ = ex->aux & 255, n_ = (ex->aux >> 8) & 255, k_ = (ex->aux >> 16) & 255;
          if (!(ex->ref & PZK_REF_VIEW_N) && s_ < 64 && n_ + k_ <= 64) {   /* the field stays inside one 64-bit word */
            uint64_t v = U[slot] >> s_;
            if (n_ < 64) v &= ((uint64_t)1 << n_) - 1;
            w[0] = v << k_;
            memcpy(dst, w, 32);
            continue;
          }
          uint64_t m[4] = {0, 0, 0, 0}, one[4] = {1, 0, 0, 0};
          if (ex->ref & PZK_REF_VIEW_N) memcpy(w, F + 4 * (uint64_t)slot, 32); else w[0] = U[slot];
          shr4(w, w, s_);
          shl4(m, one, n_); sub4(m, m, one);
          for (int i = 0; i < 4; i++) w[i] &= m[i];
          shl4(w, w, k_);
        }
        else if (cls == 2 && (ex->ref & PZK_REF_Z)) z_canonical(w, F + 4 * (uint64_t)slot);
        else if (cls == 2) from_mont(w, F + 4 * (uint64_t)slot);
        else if (cls == 1 && (int64_t)U[slot] < 0) { uint64_t m[4] = {(uint64_t)(-(int64_t)U[slot]), 0, 0, 0}; sub4(w, P, m); }
        else w[0] = U[slot];
        memcpy(dst, w, 32);
      }
    }
  }
  free(U); free(F);
  if (first_bad) *first_bad = bad;
  return status;
}

/* wtns check semantics on an explicit witness + .r1cs-equivalent rows is done in
 * oracle/formats.py with Python integers; this file only checks its own rows. */
