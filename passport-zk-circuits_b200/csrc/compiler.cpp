// compiler.cpp - see compiler.hpp.  Product code (never touches oracle/).
#include "compiler.hpp"

#include <chrono>
#include <queue>
#include <cstdarg>
#include <cstring>

namespace pzk {

static const i128 I128_ONE = 1;
static const i128 BOUND = I128_ONE << 126;
static const i128 U64_MAX_ = (I128_ONE << 64) - 1;
static const i128 I64_MIN_ = -(I128_ONE << 63);
static const i128 I64_MAX_ = (I128_ONE << 63) - 1;

static std::string fmt(const char* f, ...) {
  char buf[2048];
  va_list ap; va_start(ap, f); vsnprintf(buf, sizeof buf, f, ap); va_end(ap);
  return buf;
}

// ---- circom constant semantics (SURVEY.md section 8a footnote) ---------------------
static const U256 MASK254 = sub(shl(U256(1), 254), U256(1));
static inline bool fr_is_neg(const U256& a) { return cmp(a, fr_half()) > 0; }
static int fr_scmp(const U256& a, const U256& b) {
  bool na = fr_is_neg(a), nb = fr_is_neg(b);
  if (na != nb) return na ? -1 : 1;
  return cmp(a, b);
}
static U256 fr_shr_c(const U256& a, const U256& b);
static U256 fr_shl_c(const U256& a, const U256& b) {
  if (fr_is_neg(b)) return fr_shr_c(a, fr_neg(b));
  if (!b.fits64() || b.w[0] >= 254) return U256();
  return fr_reduce(band(shl(a, (unsigned)b.w[0]), MASK254));
}
static U256 fr_shr_c(const U256& a, const U256& b) {
  if (fr_is_neg(b)) return fr_shl_c(a, fr_neg(b));
  if (!b.fits64() || b.w[0] >= 254) return U256();
  return shr(a, (unsigned)b.w[0]);
}
static bool fold_bin(int op, const U256& a, const U256& b, U256& r) {
  switch (op) {
    case O_ADD: r = fr_add(a, b); return true;
    case O_SUB: r = fr_sub(a, b); return true;
    case O_MUL: r = fr_mul(a, b); return true;
    case O_DIV: r = b.is_zero() ? U256() : fr_mul(a, fr_inv(b)); return true;
    case O_IDIV: { if (b.is_zero()) { r = U256(); return true; } U256 q, m; divmod(a, b, q, m); r = q; return true; }
    case O_MOD: { if (b.is_zero()) { r = U256(); return true; } U256 q, m; divmod(a, b, q, m); r = m; return true; }
    case O_POW: r = fr_pow(a, b); return true;
    case O_SHL: r = fr_shl_c(a, b); return true;
    case O_SHR: r = fr_shr_c(a, b); return true;
    case O_BAND: r = fr_reduce(band(a, b)); return true;
    case O_BOR: r = fr_reduce(band(bor(a, b), MASK254)); return true;
    case O_BXOR: r = fr_reduce(band(bxor(a, b), MASK254)); return true;
    case O_EQ: r = U256(a == b); return true;
    case O_NE: r = U256(a != b); return true;
    case O_LT: r = U256(fr_scmp(a, b) < 0); return true;
    case O_GT: r = U256(fr_scmp(a, b) > 0); return true;
    case O_LE: r = U256(fr_scmp(a, b) <= 0); return true;
    case O_GE: r = U256(fr_scmp(a, b) >= 0); return true;
    case O_LAND: r = U256(!a.is_zero() && !b.is_zero()); return true;
    case O_LOR: r = U256(!a.is_zero() || !b.is_zero()); return true;
  }
  return false;
}
static U256 fold_un(int op, const U256& a) {
  switch (op) {
    case O_NEG: return fr_neg(a);
    case O_LNOT: return U256(a.is_zero());
    case O_BNOT: return fr_reduce(band(bxor(a, MASK254), MASK254));
  }
  throw CompileError("bad unary operator");
}
// small signed view of a field constant: |v| < 2^62
static bool small_signed(const U256& c, int64_t& out) {
  if (c.fits64() && c.w[0] < (1ull << 62)) { out = (int64_t)c.w[0]; return true; }
  U256 n = sub(FR_P, c);
  if (n.fits64() && n.w[0] < (1ull << 62)) { out = -(int64_t)n.w[0]; return true; }
  return false;
}
static U256 from_signed(int64_t v) { return v >= 0 ? U256((uint64_t)v) : sub(FR_P, U256((uint64_t)(-v))); }
// field element of a Z-class constant stored as a 256-bit two's complement pattern
static U256 z_field(const U256& pat) {
  if (!(pat.w[3] >> 63)) return pat;
  U256 zero;
  return sub(FR_P, sub(zero, pat));
}

// ---- linear algebra for constraints -------------------------------------------------
static void lin_addmul(Lin& dst, const Lin& src, const U256& k) {  // dst += k * src
  if (k.is_zero()) return;
  bool one = (k == U256(1));
  std::vector<std::pair<uint32_t, U256>> out;
  out.reserve(dst.t.size() + src.t.size());
  size_t i = 0, j = 0;
  while (i < dst.t.size() || j < src.t.size()) {
    if (j >= src.t.size() || (i < dst.t.size() && dst.t[i].first < src.t[j].first)) out.push_back(dst.t[i++]);
    else if (i >= dst.t.size() || src.t[j].first < dst.t[i].first) {
      U256 c = one ? src.t[j].second : fr_mul(src.t[j].second, k);
      if (!c.is_zero()) out.emplace_back(src.t[j].first, c);
      j++;
    } else {
      U256 c = fr_add(dst.t[i].second, one ? src.t[j].second : fr_mul(src.t[j].second, k));
      if (!c.is_zero()) out.emplace_back(dst.t[i].first, c);
      i++; j++;
    }
  }
  dst.t.swap(out);
  dst.k = fr_add(dst.k, one ? src.k : fr_mul(src.k, k));
}
static void lin_scale(Lin& l, const U256& k) {
  if (k == U256(1)) return;
  if (k.is_zero()) { l.t.clear(); l.k = U256(); return; }
  for (auto& t : l.t) t.second = fr_mul(t.second, k);
  l.k = fr_mul(l.k, k);
}

struct Scope { std::vector<std::pair<int, Value>> vars; };
typedef std::vector<Scope> Env;

struct Ref {
  enum K { NONE, VAR, SIG, COMP, UNKNOWN } k = NONE;
  // SIG
  Comp* comp = nullptr; Layout* lay = nullptr;
  uint32_t off = 0; std::vector<int> dims; int skind = 0; int name = -1;
  // COMP
  std::vector<int> path;
};

struct Ctx {  // one template instance being executed (or nullptr inside functions)
  Layout* lay = nullptr;
  Comp* comp = nullptr;
};

struct RowRec { uint64_t off; uint32_t na, nb, nc; bool by_def = false; int32_t line = 0, file = -1, tname = -1; };

struct Compiler::Impl {
  SourceUnit unit;
  CompileOptions opt;
  std::string main_path;
  int phase = 0;  // 0 = A, 1 = B
  std::unordered_map<std::string, Layout*> layouts;
  std::vector<Layout*> layout_list;
  Layout* main_lay = nullptr;
  std::vector<Comp*> comps;

  // SSA values (id 0 is reserved = "none")
  std::vector<uint8_t> v_cls;
  std::vector<i128> v_lo, v_hi;
  std::vector<int32_t> v_tbl;
  std::vector<uint32_t> v_convF, v_convN, v_def;
  std::vector<uint32_t> v_convZ;   // Z-class copy of a narrow value
  std::vector<uint16_t> v_zb;      // Z values: |v| < 2^v_zb
  std::vector<uint8_t> v_zneg;     // Z values: may be negative
  std::vector<uint8_t> v_const;
  std::unordered_map<uint32_t, U256> const_of;
  std::vector<Table> tables;
  std::vector<OpRec> ops;
  std::vector<uint32_t> list_pool;
  std::vector<U256> fpool;
  std::unordered_map<U256, uint32_t, U256Hash> fpool_mont, fpool_plain, constU, constF, constZ;

  // signals & constraints
  std::vector<uint32_t> sig_val;  // per signal index -> value id (0 = unassigned)
  std::vector<RowRec> rows;
  std::vector<std::pair<uint32_t, uint32_t>> terms;  // (signal index, coef index)
  std::vector<U256> coefs;
  std::unordered_map<U256, uint32_t, U256Hash> coef_idx;

  // function memo (scalar constant args)
  std::unordered_map<std::string, Value> fn_memo;
  int fn_depth = 0;
  bool returned = false;
  Value ret_val;
  // `return` under a signal-dependent condition: (condition, value) pairs of the running function;
  // execution continues (functions are pure) and the results are merged with selects at the end
  std::vector<std::vector<std::pair<SVal, Value>>> early_returns;

  CompileStats* stats = nullptr;

  // results of the back end
  std::vector<uint32_t> sig2wire;
  std::vector<uint32_t> v_slot;
  std::vector<PzkOp> out_ops;  // 16-byte records
  std::vector<PzkSegment> segs;
  std::vector<PzkRow> out_rows;
  std::vector<PzkTerm> out_terms;
  std::vector<PzkExport> out_exports;
  std::vector<PzkInput> out_inputs;
  std::vector<uint32_t> out_list;
  uint32_t n_u_slots = 0, n_f_slots = 0;
  std::vector<uint8_t> row_kind;  // per constraint: 0 run-time check, 1 alias, 2 table proof, 3 symbolic proof, 4 definitional
  uint64_t n_static_rows = 0, n_def_rows = 0, n_table_rows = 0, n_symbolic_rows = 0, n_fused = 0, n_fused_mac = 0, n_fused_mac_wire = 0, n_z_u_operands = 0, n_dig_wide = 0, n_dig_macro = 0, n_i64_rows = 0, n_int_rows = 0, n_field_rows = 0, eval_bytes = 0, check_bytes = 0, cache_hit_refs = 0, cache_miss_refs = 0;
  std::vector<uint32_t> seg_quads;
  uint32_t n_pub_out = 0, n_pub_in = 0, n_prv_in = 0;
  std::string meta_json;
  std::string o1_stats;

  const std::string& nm(int id) { return unit.names.str(id); }
  [[noreturn]] void fail(const Stmt* s, const std::string& m) {
    throw CompileError(fmt("%s:%d: %s", s ? unit.files[s->file].c_str() : "?", s ? s->line : 0, m.c_str()));
  }
  [[noreturn]] void fail(const std::string& m) { throw CompileError(m); }

  // ================================================================== values
  uint32_t new_value(int cls, i128 lo = 0, i128 hi = 0, int32_t tbl = -1) {
    uint32_t id = (uint32_t)v_cls.size();
    v_cls.push_back((uint8_t)cls); v_lo.push_back(lo); v_hi.push_back(hi); v_tbl.push_back(tbl);
    v_convF.push_back(0); v_convN.push_back(0); v_def.push_back((uint32_t)ops.size()); v_const.push_back(0);
    v_convZ.push_back(0); v_zb.push_back(0); v_zneg.push_back(0);
    return id;
  }
  // ---- batched inversion (Montgomery's trick), decided at compile time ----------------------
  // F_INV results are placeholders until some emitted op needs one of them (or the queue is
  // full); then the whole queue is materialised with ONE field inversion and 3 products per
  // element.  inv(0) = 0 is preserved: zeros are replaced by 1 inside the product chain and the
  // corresponding results forced back to 0.
  std::vector<std::pair<uint32_t, uint32_t>> pending_inv;  // (operand value, placeholder result)
  std::vector<uint8_t> v_pending;
  std::vector<OpRec> deferred;
  std::vector<uint32_t> deferred_defs;
  std::vector<std::pair<uint32_t, uint32_t>> deferred_lutv;
  uint32_t pending_lutv_off = 0;
  bool flushing = false;
  static const size_t INV_BATCH = 48;
  // macro ops (long_div, mod_inv, the BabyJubjub ladder) read and write their values through the list pool
  static bool is_macro(int opc) { return opc == PZK_BIGDIV || opc == PZK_MODINV || opc == PZK_BJJ_MUL8; }
  void macro_layout(const OpRec& o, uint32_t& op0, uint32_t& nop, uint32_t& def0, uint32_t& ndef) const {
    if (o.opc == PZK_BJJ_MUL8) { op0 = o.a + 5; nop = 1; def0 = o.a + 6; ndef = 4 * (2 * list_pool[o.a] - 1); return; }
    uint32_t k = list_pool[o.a + 1], m = list_pool[o.a + 2];
    op0 = o.a + 3; nop = k + m + k; def0 = op0 + nop; ndef = m + 1 + k;
  }
  bool op_reads_values(int opc) {
    switch (opc) { case PZK_NOP: case PZK_U_CONST: case PZK_F_CONST: case PZK_Z_CONST: case PZK_IN_U: case PZK_IN_F: case PZK_BIGDIV: case PZK_MODINV: case PZK_BJJ_MUL8: return false; }
    return true;
  }
  bool is_pending(uint32_t v) { return v != PZK_OPERAND_NONE && v < v_pending.size() && v_pending[v]; }
  uint32_t emit(int opc, uint32_t dst, uint32_t a = 0, uint32_t b = 0, int flags = 0, int imm16 = 0,
                uint32_t c = PZK_OPERAND_NONE, uint32_t d = PZK_OPERAND_NONE) {
    if (!flushing && !pending_inv.empty() && op_reads_values(opc)) {
      bool need = is_pending(a);
      bool b_is_value = !(flags & (PZK_FLAG_B_IMM | PZK_FLAG_B_POOL)) && opc != PZK_N_BIT && opc != PZK_F_CSEL;
      if (b_is_value && is_pending(b)) need = true;
      if (is_pending(c) || is_pending(d)) need = true;
      if (need) {
        // the op depends on a queued inversion: defer it (and transitively its users) so that
        // further independent inversions can join the batch
        OpRec o; o.opc = (uint8_t)opc; o.flags = (uint8_t)flags; o.imm16 = (uint16_t)imm16;
        o.dst = dst; o.a = a; o.b = b; o.c = c; o.d = d;
        deferred.push_back(o);
        if (opc == PZK_U_LUTV) deferred_lutv.push_back({(uint32_t)deferred.size() - 1, pending_lutv_off});
        if (dst && opc != PZK_ASSERT_NZ) {
          if (v_pending.size() <= dst) v_pending.resize(dst + 1024, 0);
          v_pending[dst] = 1;
          deferred_defs.push_back(dst);
        }
        if (deferred.size() > 8192) flush_inversions();
        return dst;
      }
    }
    OpRec o; o.opc = (uint8_t)opc; o.flags = (uint8_t)flags; o.imm16 = (uint16_t)imm16;
    o.dst = dst; o.a = a; o.b = b; o.c = c; o.d = d;
    if (dst && opc != PZK_ASSERT_NZ && !is_macro(opc) && dst < v_def.size()) v_def[dst] = (uint32_t)ops.size();
    ops.push_back(o);
    if (op_stats) op_tag.push_back(cur_tname);
    return dst;
  }
  bool op_stats = getenv("PZK_OP_STATS") != nullptr;
  int cur_tname = -1;
  std::vector<int> op_tag;
  uint32_t queue_inversion(uint32_t x) {
    if (is_pending(x)) flush_inversions();
    uint32_t r = new_value(CLS_F);
    if (v_pending.size() <= r) v_pending.resize(r + 1024, 0);
    v_pending[r] = 1;
    pending_inv.emplace_back(x, r);
    inv_of[r] = x;
    stats->f_inv++;
    if (pending_inv.size() >= INV_BATCH) flush_inversions();
    return r;
  }
  void flush_inversions() {
    if (pending_inv.empty()) return;
    flushing = true;
    std::vector<std::pair<uint32_t, uint32_t>> q;
    q.swap(pending_inv);
    for (auto& pr : q) v_pending[pr.second] = 0;
    size_t n = q.size();
    if (n == 1) {
      emit(PZK_F_INV, q[0].second, q[0].first);
      stats->f_inv_real++;
    } else {
      uint32_t one = pool_mont(U256(1)), zero = pool_mont(U256());
      std::vector<uint32_t> isz(n), xs(n), pre(n);
      for (size_t i = 0; i < n; i++) {
        isz[i] = new_value(CLS_U, 0, 1);
        emit(PZK_F_EQ, isz[i], q[i].first, zero, PZK_FLAG_B_POOL);
        uint32_t c1 = const_value_F(U256(1));
        xs[i] = new_value(CLS_F);
        emit(PZK_F_SEL, xs[i], isz[i], c1, PZK_FLAG_EXT, 0, q[i].first);
        if (i == 0) pre[0] = xs[0];
        else { pre[i] = new_value(CLS_F); emit(PZK_F_MUL, pre[i], pre[i - 1], xs[i]); stats->f_mul++; }
      }
      (void)one;
      uint32_t acc = new_value(CLS_F);
      emit(PZK_F_INV, acc, pre[n - 1]);
      stats->f_inv_real++;
      uint32_t c0 = const_value_F(U256());
      for (size_t i = n; i-- > 0;) {
        uint32_t inv_i;
        if (i == 0) inv_i = acc;
        else {
          inv_i = new_value(CLS_F); emit(PZK_F_MUL, inv_i, acc, pre[i - 1]); stats->f_mul++;
          uint32_t nacc = new_value(CLS_F); emit(PZK_F_MUL, nacc, acc, xs[i]); stats->f_mul++;
          acc = nacc;
        }
        emit(PZK_F_SEL, q[i].second, isz[i], c0, PZK_FLAG_EXT, 0, inv_i);
      }
    }
    // now the deferred dependants, in their original order
    std::vector<OpRec> dq; dq.swap(deferred);
    std::vector<std::pair<uint32_t, uint32_t>> lq; lq.swap(deferred_lutv);
    for (uint32_t d : deferred_defs) v_pending[d] = 0;
    deferred_defs.clear();
    size_t li = 0;
    for (size_t k = 0; k < dq.size(); k++) {
      const OpRec& o = dq[k];
      if (o.dst && o.opc != PZK_ASSERT_NZ && !is_macro(o.opc) && o.dst < v_def.size()) v_def[o.dst] = (uint32_t)ops.size();
      ops.push_back(o);
      if (op_stats) op_tag.push_back(cur_tname);
      if (li < lq.size() && lq[li].first == k) { lutv_off[(uint32_t)ops.size() - 1] = lq[li].second; li++; }
    }
    flushing = false;
  }
  uint32_t pool_mont(const U256& c) {
    auto it = fpool_mont.find(c);
    if (it != fpool_mont.end()) return it->second;
    uint32_t i = (uint32_t)fpool.size(); fpool.push_back(fr_to_mont(c)); fpool_mont[c] = i; return i;
  }
  uint32_t pool_plain(const U256& c) {
    auto it = fpool_plain.find(c);
    if (it != fpool_plain.end()) return it->second;
    uint32_t i = (uint32_t)fpool.size(); fpool.push_back(c); fpool_plain[c] = i; return i;
  }
  // constant as a narrow (U/I) value
  bool const_narrow(const U256& c, i128& v) {
    if (c.fits64()) { v = (i128)c.w[0]; return true; }
    U256 n = sub(FR_P, c);
    if (n.fits64() && n.w[0] <= (1ull << 63)) { v = -(i128)n.w[0]; return true; }
    return false;
  }
  uint32_t const_value_U(const U256& c) {  // c must be narrow
    auto it = constU.find(c);
    if (it != constU.end()) return it->second;
    i128 v; if (!const_narrow(c, v)) fail("internal: const_value_U on wide constant");
    uint64_t bits = (uint64_t)v;
    uint32_t id = new_value(v >= 0 ? CLS_U : CLS_I, v, v);
    emit(PZK_U_CONST, id, (uint32_t)bits, (uint32_t)(bits >> 32));
    v_const[id] = 1; const_of[id] = c; constU[c] = id;
    return id;
  }
  uint32_t const_value_F(const U256& c) {
    auto it = constF.find(c);
    if (it != constF.end()) return it->second;
    uint32_t id = new_value(CLS_F);
    emit(PZK_F_CONST, id, pool_mont(c));
    v_const[id] = 1; const_of[id] = c; constF[c] = id;
    return id;
  }
  uint32_t const_value(const U256& c) {
    i128 v;
    if (const_narrow(c, v)) return const_value_U(c);
    return const_value_F(c);
  }

  bool narrow_of(const SVal& s, i128& lo, i128& hi) {
    if (s.kind == 0) { i128 v; if (!const_narrow(s.c, v)) return false; lo = hi = v; return true; }
    if (s.kind != 1) return false;
    if (v_cls[s.id] == CLS_U || v_cls[s.id] == CLS_I) { lo = v_lo[s.id]; hi = v_hi[s.id]; return true; }
    return false;
  }
  bool exact_u(const SVal& s, i128& hi) {  // canonical value provably in [0, 2^64)
    if (s.kind == 0) { if (!s.c.fits64()) return false; hi = (i128)s.c.w[0]; return true; }
    if (s.kind == 1 && v_cls[s.id] == CLS_U) { hi = v_hi[s.id]; return true; }
    return false;
  }
  static int cls_for(i128 lo, i128 hi) {
    if (lo >= 0 && hi <= U64_MAX_) return CLS_U;
    if (lo >= I64_MIN_ && hi <= I64_MAX_) return CLS_I;
    return -1;
  }
  uint32_t u_operand(const SVal& s) {
    if (s.kind == 0) return const_value_U(s.c);
    return s.id;
  }
  uint32_t to_F(const SVal& s) {
    if (s.kind == 0) return const_value_F(s.c);
    uint32_t id = s.id;
    int c = v_cls[id];
    if (c == CLS_F) return id;
    if (v_convF[id]) return v_convF[id];
    uint32_t r = new_value(CLS_F);
    if (c == CLS_U) emit(PZK_F_FROM_U, r, id);
    else if (c == CLS_I) emit(PZK_F_FROM_I, r, id);
    else if (c == CLS_Z) emit(PZK_F_FROM_N, r, id, 0, PZK_FLAG_ZSRC);
    else emit(PZK_F_FROM_N, r, id);
    v_convF[id] = r;
    return r;
  }
  uint32_t to_N(const SVal& s) {
    if (s.kind == 0) {
      uint32_t r = new_value(CLS_N);
      emit(PZK_F_CONST, r, pool_plain(s.c));
      return r;
    }
    uint32_t id = s.id;
    int c = v_cls[id];
    if (c == CLS_N) return id;
    if (v_convN[id]) return v_convN[id];
    uint32_t r = new_value(CLS_N);
    if (c == CLS_U) emit(PZK_N_FROM_U, r, id);
    else if (c == CLS_F) emit(PZK_N_FROM_F, r, id);
    else if (c == CLS_Z) emit(PZK_N_FROM_F, r, id, 0, PZK_FLAG_ZSRC);
    else { uint32_t f = to_F(s); emit(PZK_N_FROM_F, r, f); }
    v_convN[id] = r;
    return r;
  }
  SVal sv(uint32_t id) { return SVal::ssa(id); }

  // ================================================================== Z class: exact wide integers
  // A value is "integral" when the compiler knows it as an integer v (not just a residue): narrow values, Z values
  // and constants of small magnitude.  bits: |v| < 2^bits; neg: v may be negative.
  static const int ZMAX = 250;
  bool int_bits(const SVal& s, int& bits, bool& neg) {
    if (s.kind == 0) {
      int bl = s.c.bitlen();
      if (bl <= ZMAX) { bits = bl; neg = false; return true; }
      U256 n = sub(FR_P, s.c);
      if (n.bitlen() <= ZMAX) { bits = n.bitlen(); neg = true; return true; }
      return false;
    }
    if (s.kind != 1) return false;
    int c = v_cls[s.id];
    if (c == CLS_U || c == CLS_I) {
      i128 a = v_lo[s.id] < 0 ? -v_lo[s.id] : v_lo[s.id], b = v_hi[s.id] < 0 ? -v_hi[s.id] : v_hi[s.id];
      bits = bitlen128(a > b ? a : b); neg = v_lo[s.id] < 0;
      return true;
    }
    if (c == CLS_Z) { bits = v_zb[s.id]; neg = v_zneg[s.id] != 0; return true; }
    return false;
  }
  // two's complement 256-bit pattern of a constant of small magnitude (as the field represents it)
  U256 z_pattern(const U256& c) {
    if (c.bitlen() <= ZMAX) return c;
    U256 n = sub(FR_P, c);          // c = -n
    U256 zero;
    return sub(zero, n);            // 2^256 - n
  }
  uint32_t to_Z(const SVal& s) {
    if (s.kind == 0) {
      auto it = constZ.find(s.c);
      if (it != constZ.end()) return it->second;
      uint32_t id = new_value(CLS_Z);
      int b; bool ng; int_bits(s, b, ng);
      v_zb[id] = (uint16_t)b; v_zneg[id] = ng;
      emit(PZK_Z_CONST, id, pool_plain(z_pattern(s.c)));
      v_const[id] = 1; const_of[id] = s.c; constZ[s.c] = id;
      return id;
    }
    uint32_t id = s.id;
    int c = v_cls[id];
    if (c == CLS_Z) return id;
    if (v_convZ[id]) return v_convZ[id];
    uint32_t r = new_value(CLS_Z);
    int b; bool ng; int_bits(s, b, ng);
    v_zb[r] = (uint16_t)b; v_zneg[r] = ng;
    emit(c == CLS_U ? PZK_Z_FROM_U : PZK_Z_FROM_I, r, id);
    v_convZ[id] = r;
    stats->z_ops++;
    return r;
  }
  SVal emit_z(int op, const SVal& a, const SVal& b, int rbits, bool rneg) {
    int ba = 0, bb = 0; bool na = false, nb = false;
    int_bits(a, ba, na); int_bits(b, bb, nb);
    uint32_t ia = to_Z(a);
    uint32_t id;
    int imm = 0;
    if (op == O_MUL && !na && !nb) {
      int la = (ba + 63) / 64, lb = (bb + 63) / 64;   // 64-bit limbs the (non-negative) operands fit
      if (la < 1) la = 1;
      if (lb < 1) lb = 1;
      if (la + lb <= 4) imm = la | (lb << 4);
    }
    const int opc = op == O_ADD ? PZK_Z_ADD : op == O_SUB ? PZK_Z_SUB : PZK_Z_MUL;
    if (b.kind == 0) {
      uint32_t pi = pool_plain(z_pattern(b.c));
      id = new_value(CLS_Z);
      emit(opc, id, ia, pi, PZK_FLAG_B_POOL, imm);
    } else {
      uint32_t ib = to_Z(b);
      id = new_value(CLS_Z);
      emit(opc, id, ia, ib, 0, imm);
    }
    v_zb[id] = (uint16_t)rbits; v_zneg[id] = rneg;
    if (op == O_MUL) stats->z_mul++; else stats->z_ops++;
    return sv(id);
  }

  // ================================================================== tables
  bool table_of(const SVal& s, Table& t, bool root_only) {
    if (s.kind == 0) { int64_t v; if (!small_signed(s.c, v)) return false; t.n = 0; t.e[0] = v; return true; }
    if (s.kind != 1) return false;
    uint32_t id = s.id;
    if (!root_only && v_tbl[id] >= 0) { t = tables[v_tbl[id]]; return true; }
    if (v_cls[id] == CLS_U && v_lo[id] >= 0 && v_hi[id] <= 1) { t.n = 1; t.sup[0] = id; t.e[0] = 0; t.e[1] = 1; return true; }
    return false;
  }
  // evaluate f over the union support of up to three tables
  bool table_combine(const Table* in, int nin, Table& out, const std::function<bool(const int64_t*, int64_t&)>& f) {
    uint32_t sup[12]; int n = 0;
    for (int k = 0; k < nin; k++)
      for (int j = 0; j < in[k].n; j++) {
        bool dup = false;
        for (int q = 0; q < n; q++) if (sup[q] == in[k].sup[j]) dup = true;
        if (!dup) { if (n >= 12) return false; sup[n++] = in[k].sup[j]; }
      }
    if (n > 4) return false;
    std::sort(sup, sup + n);
    out.n = (uint8_t)n;
    for (int j = 0; j < n; j++) out.sup[j] = sup[j];
    int pos[3][4];
    for (int k = 0; k < nin; k++)
      for (int j = 0; j < in[k].n; j++)
        for (int q = 0; q < n; q++) if (sup[q] == in[k].sup[j]) pos[k][j] = q;
    for (int idx = 0; idx < (1 << n); idx++) {
      int64_t args[3];
      for (int k = 0; k < nin; k++) {
        int sub = 0;
        for (int j = 0; j < in[k].n; j++) sub |= ((idx >> pos[k][j]) & 1) << j;
        args[k] = in[k].e[sub];
      }
      if (!f(args, out.e[idx])) return false;
    }
    return true;
  }
  bool try_table(const SVal* xs, int nx, Table& out, const std::function<bool(const int64_t*, int64_t&)>& f) {
    Table in[3];
    bool ok = true;
    for (int k = 0; k < nx && ok; k++) ok = table_of(xs[k], in[k], false);
    if (ok && table_combine(in, nx, out, f)) return true;
    ok = true;
    for (int k = 0; k < nx && ok; k++) ok = table_of(xs[k], in[k], true);
    if (ok && table_combine(in, nx, out, f)) return true;
    return false;
  }
  static bool table_is_const(const Table& t, int64_t& v) {
    for (int i = 1; i < (1 << t.n); i++) if (t.e[i] != t.e[0]) return false;
    v = t.e[0]; return true;
  }
  static bool table_bits(const Table& t) {
    for (int i = 0; i < (1 << t.n); i++) if (t.e[i] != 0 && t.e[i] != 1) return false;
    return true;
  }
  static void table_range(const Table& t, i128& lo, i128& hi) {
    lo = hi = t.e[0];
    for (int i = 1; i < (1 << t.n); i++) { if (t.e[i] < lo) lo = t.e[i]; if (t.e[i] > hi) hi = t.e[i]; }
  }
  // materialise a table-valued result as LUT / LUTV over its support
  uint32_t emit_table_value(const Table& t) {
    i128 lo, hi; table_range(t, lo, hi);
    int cls = cls_for(lo, hi);
    tables.push_back(t);
    int32_t ti = (int32_t)tables.size() - 1;
    uint32_t id = new_value(cls, lo, hi, ti);
    uint32_t s[4] = {PZK_OPERAND_NONE, PZK_OPERAND_NONE, PZK_OPERAND_NONE, PZK_OPERAND_NONE};
    for (int j = 0; j < t.n; j++) s[j] = t.sup[j];
    if (table_bits(t)) {
      int imm = 0;
      for (int i = 0; i < 16; i++) if (t.e[i & ((1 << t.n) - 1)] == 1) imm |= 1 << i;
      emit(PZK_U_LUT, id, s[0], s[1], PZK_FLAG_EXT, imm, s[2], s[3]);
    } else {
      uint32_t off = (uint32_t)list_pool.size();
      for (int i = 0; i < 16; i++) {
        uint64_t v = (uint64_t)t.e[i & ((1 << t.n) - 1)];
        list_pool.push_back((uint32_t)v); list_pool.push_back((uint32_t)(v >> 32));
      }
      pending_lutv_off = off;
      size_t before = ops.size();
      emit(PZK_U_LUTV, id, s[0], s[1], PZK_FLAG_EXT, 0, s[2], s[3]);
      if (ops.size() > before && ops.back().opc == PZK_U_LUTV && ops.back().dst == id) lutv_off[(uint32_t)ops.size() - 1] = off;
    }
    stats->lut++;
    return id;
  }
  std::unordered_map<uint32_t, uint32_t> lutv_off;  // op index -> list offset

  // ================================================================== op emission
  SVal fold_or_emit_bin(int op, const SVal& a, const SVal& b, const Stmt* at) {
    if (a.kind == 2 || b.kind == 2) return SVal::unk();
    if (a.kind == 0 && b.kind == 0) {
      U256 r;
      if (!fold_bin(op, a.c, b.c, r)) fail(at, "unsupported operator");
      return SVal::konst(r);
    }
    if (phase == 0) return SVal::unk();
    return emit_bin(op, a, b, at);
  }

  bool is_ring(int op) { return op == O_ADD || op == O_SUB || op == O_MUL; }

  // Facts about values that let `x * inv(x)` be computed without the inverse (IsZero / IsEqual,
  // circomlib comparators: `inv <-- in != 0 ? 1/in : 0; out <== -in*inv + 1`): x * inv(x) is the bit
  // (x != 0) because inv(0) = 0.  The bit is a narrow value, so everything downstream of an IsZero stays
  // in 64-bit arithmetic, and `out` no longer waits for the batched inversion.
  std::unordered_map<uint32_t, uint32_t> inv_of, neg_of, nz_of;  // result value -> operand value
  bool same_value(uint32_t x, uint32_t y) {
    if (x == y) return true;
    if (x < v_convF.size() && v_convF[x] && v_convF[x] == y) return true;
    if (y < v_convF.size() && v_convF[y] && v_convF[y] == x) return true;
    if (x < v_convZ.size() && v_convZ[x] && v_convZ[x] == y) return true;
    if (y < v_convZ.size() && v_convZ[y] && v_convZ[y] == x) return true;
    return false;
  }
  bool try_inverse_product(const SVal& p, const SVal& q, SVal& out, const Stmt* at) {
    auto it = inv_of.find(q.id);
    if (it == inv_of.end()) return false;
    uint32_t v = it->second;
    if (same_value(p.id, v)) { out = truthy(p, at); return true; }
    auto nt = neg_of.find(p.id);
    if (nt != neg_of.end() && same_value(nt->second, v)) {
      out = emit_bin(O_SUB, SVal::konst(U256()), truthy(sv(nt->second), at), at);
      return true;
    }
    return false;
  }
  SVal emit_bin(int op, SVal a, SVal b, const Stmt* at) {
    if (op == O_MUL && a.kind == 1 && b.kind == 1 && !inv_of.empty()) {
      SVal r;
      if (try_inverse_product(a, b, r, at) || try_inverse_product(b, a, r, at)) return r;
    }
    SVal r = emit_bin_core(op, a, b, at);
    if (r.kind == 1) {
      if (op == O_SUB && a.kind == 0 && a.c.is_zero() && b.kind == 1) neg_of[r.id] = b.id;
      else if (op == O_NE && b.kind == 0 && b.c.is_zero() && a.kind == 1) nz_of[r.id] = a.id;
      else if (op == O_NE && a.kind == 0 && a.c.is_zero() && b.kind == 1) nz_of[r.id] = b.id;
    }
    return r;
  }
  SVal emit_bin_core(int op, SVal a, SVal b, const Stmt* at) {
    // canonicalise > and >= into < and <=
    if (op == O_GT) { std::swap(a, b); op = O_LT; }
    else if (op == O_GE) { std::swap(a, b); op = O_LE; }
    // algebraic identities
    if (op == O_ADD) { if (a.kind == 0 && a.c.is_zero()) return b; if (b.kind == 0 && b.c.is_zero()) return a; }
    if (op == O_SUB && b.kind == 0 && b.c.is_zero()) return a;
    if (op == O_MUL) {
      if ((a.kind == 0 && a.c.is_zero()) || (b.kind == 0 && b.c.is_zero())) return SVal::konst(U256());
      if (a.kind == 0 && a.c == U256(1)) return b;
      if (b.kind == 0 && b.c == U256(1)) return a;
    }
    if (op == O_DIV && b.kind == 0) {
      if (b.c.is_zero()) return SVal::konst(U256());
      return emit_bin(O_MUL, a, SVal::konst(fr_inv(b.c)), at);
    }
    if (op == O_POW) {
      if (b.kind != 0 || !b.c.fits64() || b.c.w[0] > 64) fail(at, "** with a signal-dependent or large exponent");
      SVal r = SVal::konst(U256(1));
      for (uint64_t i = 0; i < b.c.w[0]; i++) r = (r.kind == 0 && r.c == U256(1)) ? a : emit_bin(O_MUL, r, a, at);
      return r;
    }
    // ---- exhaustive table evaluation over <= 4 root bits
    {
      SVal xs[2] = {a, b};
      Table t;
      auto f = [&](const int64_t* v, int64_t& r) {
        U256 out;
        if (!fold_bin(op, from_signed(v[0]), from_signed(v[1]), out)) return false;
        return small_signed(out, r);
      };
      if (try_table(xs, 2, t, f)) {
        int64_t cv;
        if (table_is_const(t, cv)) return SVal::konst(from_signed(cv));
        bool derived = false;
        for (int k = 0; k < 2; k++)
          if (xs[k].kind == 1 && v_tbl[xs[k].id] >= 0 && tables[v_tbl[xs[k].id]].n > 0 &&
              !(tables[v_tbl[xs[k].id]].n == 1 && tables[v_tbl[xs[k].id]].sup[0] == xs[k].id)) derived = true;
        if (!is_ring(op) || (table_bits(t) && derived)) return sv(emit_table_value(t));
        // ring op with a table: emit plain arithmetic below but remember the table
        SVal r = emit_bin_notable(op, a, b, at);
        if (r.kind == 1 && (v_cls[r.id] == CLS_U || v_cls[r.id] == CLS_I)) {
          tables.push_back(t); v_tbl[r.id] = (int32_t)tables.size() - 1;
          i128 lo, hi; table_range(t, lo, hi);
          int c = cls_for(lo, hi);
          if (c >= 0 && c <= v_cls[r.id]) { v_cls[r.id] = (uint8_t)c; v_lo[r.id] = lo; v_hi[r.id] = hi; }
        }
        return r;
      }
    }
    return emit_bin_notable(op, a, b, at);
  }

  SVal emit_u(int opc, const SVal& a, const SVal& b, int cls, i128 lo, i128 hi) {
    uint32_t id;
    uint32_t ia = u_operand(a);
    if (b.kind == 0 && b.c.fits64() && b.c.w[0] < (1ull << 32)) {
      id = new_value(cls, lo, hi);
      emit(opc, id, ia, (uint32_t)b.c.w[0], PZK_FLAG_B_IMM);
    } else {
      uint32_t ib = u_operand(b);
      id = new_value(cls, lo, hi);
      emit(opc, id, ia, ib);
    }
    stats->u_ops++;
    return sv(id);
  }
  SVal emit_f(int opc, const SVal& a, const SVal& b) {
    uint32_t ia = to_F(a);
    uint32_t id;
    if (b.kind == 0) {
      uint32_t pi = pool_mont(b.c);
      id = new_value(CLS_F);
      emit(opc, id, ia, pi, PZK_FLAG_B_POOL);
    } else {
      uint32_t ib = to_F(b);
      id = new_value(CLS_F);
      emit(opc, id, ia, ib);
    }
    if (opc == PZK_F_MUL) stats->f_mul++; else stats->f_other++;
    return sv(id);
  }
  // N-class binary op; operand b may be a pool constant (plain) or immediate shift
  SVal emit_n(int opc, const SVal& a, const SVal& b, bool b_is_shift) {
    uint32_t ia = to_N(a);
    uint32_t id;
    if (b_is_shift) {
      if (b.kind == 0) {
        if (!b.c.fits64() || b.c.w[0] > 300) fail("shift amount too large");
        id = new_value(CLS_N); emit(opc, id, ia, (uint32_t)b.c.w[0], PZK_FLAG_B_IMM);
      } else {
        i128 hi; if (!exact_u(b, hi)) fail("shift by a wide signal-dependent amount");
        id = new_value(CLS_N); emit(opc, id, ia, b.id);
      }
    } else if (b.kind == 0) {
      uint32_t pi = pool_plain(b.c);
      id = new_value(CLS_N); emit(opc, id, ia, pi, PZK_FLAG_B_POOL);
    } else {
      uint32_t ib = to_N(b);
      id = new_value(CLS_N); emit(opc, id, ia, ib);
    }
    stats->f_other++;
    return sv(id);
  }
  SVal n_low(const SVal& n, i128 hi) {
    uint32_t id = new_value(CLS_U, 0, hi);
    emit(PZK_N_LOW, id, n.id);
    return sv(id);
  }
  SVal truthy(const SVal& a, const Stmt* at) {  // -> 0/1 value
    if (a.kind == 0) return SVal::konst(U256(!a.c.is_zero()));
    i128 lo, hi;
    if (narrow_of(a, lo, hi) && lo >= 0 && hi <= 1) return a;
    return emit_bin(O_NE, a, SVal::konst(U256()), at);
  }

  SVal emit_bin_notable(int op, const SVal& a, const SVal& b, const Stmt* at) {
    i128 alo, ahi, blo, bhi;
    bool an = narrow_of(a, alo, ahi), bn = narrow_of(b, blo, bhi);
    switch (op) {
      case O_ADD: case O_SUB: case O_MUL: {
        if (an && bn) {
          i128 lo, hi; bool ok = true;
          if (op == O_ADD) { lo = alo + blo; hi = ahi + bhi; }
          else if (op == O_SUB) { lo = alo - bhi; hi = ahi - blo; }
          else {
            i128 c[4]; ok = !__builtin_mul_overflow(alo, blo, &c[0]) && !__builtin_mul_overflow(alo, bhi, &c[1]) &&
                            !__builtin_mul_overflow(ahi, blo, &c[2]) && !__builtin_mul_overflow(ahi, bhi, &c[3]);
            if (ok) { lo = hi = c[0]; for (int i = 1; i < 4; i++) { if (c[i] < lo) lo = c[i]; if (c[i] > hi) hi = c[i]; } }
          }
          if (ok && lo > -BOUND && hi < BOUND) {
            int cls = cls_for(lo, hi);
            if (cls >= 0) {
              if ((op == O_ADD || op == O_MUL) && a.kind == 0) return emit_u(op == O_ADD ? PZK_U_ADD : PZK_U_MUL, b, a, cls, lo, hi);
              return emit_u(op == O_ADD ? PZK_U_ADD : op == O_SUB ? PZK_U_SUB : PZK_U_MUL, a, b, cls, lo, hi);
            }
          }
        }
        if (opt.zclass) {
          // both operands are integers of bounded magnitude and so is the result: exact integer arithmetic in the Z
          // class (sound in the field because |result| < 2^250 < p / 2: the integer determines the residue)
          int ba, bb; bool na, nb;
          if (int_bits(a, ba, na) && int_bits(b, bb, nb)) {
            int rb = op == O_MUL ? ba + bb : (ba > bb ? ba : bb) + 1;
            bool rn = op == O_SUB ? true : (na || nb);
            if (rb <= ZMAX) {
              if ((op == O_ADD || op == O_MUL) && a.kind == 0) return emit_z(op, b, a, rb, rn);
              return emit_z(op, a, b, rb, rn);
            }
          }
        }
        if ((op == O_ADD || op == O_MUL) && a.kind == 0) return emit_f(op == O_ADD ? PZK_F_ADD : PZK_F_MUL, b, a);
        return emit_f(op == O_ADD ? PZK_F_ADD : op == O_SUB ? PZK_F_SUB : PZK_F_MUL, a, b);
      }
      case O_DIV: {
        // a / b = a * inv(b), inv(0) = 0
        uint32_t ib = to_F(b);
        uint32_t inv = queue_inversion(ib);
        if (a.kind == 0 && a.c == U256(1)) return sv(inv);
        return emit_f(PZK_F_MUL, sv(inv), a);
      }
      case O_IDIV: case O_MOD: {
        i128 ah, bh;
        if (exact_u(a, ah) && exact_u(b, bh)) {
          if (b.kind == 0 && b.c.w[0] && (b.c.w[0] & (b.c.w[0] - 1)) == 0) {
            int k = __builtin_ctzll(b.c.w[0]);
            if (op == O_IDIV) return emit_u(PZK_U_SHR, a, SVal::konst(U256((uint64_t)k)), CLS_U, 0, ah >> k);
            return emit_u(PZK_U_AND, a, SVal::konst(U256(b.c.w[0] - 1)), CLS_U, 0, std::min<i128>(ah, (i128)b.c.w[0] - 1));
          }
          if (op == O_IDIV) return emit_u(PZK_U_DIV, a, b, CLS_U, 0, ah);
          return emit_u(PZK_U_MOD, a, b, CLS_U, 0, std::min<i128>(ah, bh > 0 ? bh - 1 : 0));
        }
        // wide: power-of-two divisors become shifts / masks
        if (b.kind == 0 && !b.c.is_zero()) {
          int bl = b.c.bitlen();
          if (b.c == shl(U256(1), bl - 1)) {
            int k = bl - 1;
            if (op == O_IDIV) return shift_right(a, k, at);
            if (k <= 64) {
              SVal n = sv(to_N(a));
              SVal low = n_low(n, U64_MAX_);
              if (k == 64) return low;
              return emit_u(PZK_U_AND, low, SVal::konst(U256((1ull << k) - 1)), CLS_U, 0, ((i128)1 << k) - 1);
            }
          }
        }
        SVal r = emit_n(op == O_IDIV ? PZK_N_DIV : PZK_N_MOD, a, b, false);
        if (op == O_MOD && exact_u(b, bh)) return n_low(r, bh > 0 ? bh - 1 : 0);
        return r;
      }
      case O_SHR: {
        if (b.kind == 0 && b.c.fits64() && b.c.w[0] < 254) return shift_right(a, (int)b.c.w[0], at);
        i128 ah, bh;
        if (exact_u(a, ah) && exact_u(b, bh)) return emit_u(PZK_U_SHR, a, b, CLS_U, 0, ah);
        return emit_n(PZK_N_SHR, a, b, true);
      }
      case O_SHL: {
        i128 ah, bh;
        if (exact_u(a, ah) && b.kind == 0 && b.c.fits64() && b.c.w[0] < 64) {
          i128 hi = ah << (int)b.c.w[0];
          if (hi <= U64_MAX_) return emit_u(PZK_U_SHL, a, b, CLS_U, 0, hi);
        }
        (void)bh;
        return emit_n(PZK_N_SHL, a, b, true);
      }
      case O_BAND: case O_BOR: case O_BXOR: {
        i128 ah, bh;
        if (exact_u(a, ah) && exact_u(b, bh)) {
          i128 hi = op == O_BAND ? std::min(ah, bh) : U64_MAX_;
          if (op != O_BAND) { int bl = 0; i128 m = std::max(ah, bh); while (m) { bl++; m >>= 1; } hi = bl >= 64 ? U64_MAX_ : (((i128)1 << bl) - 1); }
          if (a.kind == 0) return emit_u(op == O_BAND ? PZK_U_AND : op == O_BOR ? PZK_U_OR : PZK_U_XOR, b, a, CLS_U, 0, hi);
          return emit_u(op == O_BAND ? PZK_U_AND : op == O_BOR ? PZK_U_OR : PZK_U_XOR, a, b, CLS_U, 0, hi);
        }
        if (op == O_BAND) {
          // (x >> k) & 1 on a wide value -> single bit extract
          const SVal* wide = nullptr; const SVal* msk = nullptr;
          if (b.kind == 0) { wide = &a; msk = &b; } else if (a.kind == 0) { wide = &b; msk = &a; }
          if (wide && msk->c.fits64()) {
            if (msk->c.w[0] == 1 && wide->kind == 1 && v_cls[wide->id] == CLS_N && !is_pending(wide->id)) {
              const OpRec& d = ops[v_def[wide->id]];
              if (d.opc == PZK_N_SHR && (d.flags & PZK_FLAG_B_IMM) && d.dst == wide->id && d.b < 256) {
                uint32_t id = new_value(CLS_U, 0, 1);
                emit(PZK_N_BIT, id, d.a, d.b, PZK_FLAG_B_IMM);
                stats->u_ops++;
                return sv(id);
              }
            }
            SVal low = n_low(sv(to_N(*wide)), U64_MAX_);
            return emit_u(PZK_U_AND, low, *msk, CLS_U, 0, (i128)msk->c.w[0]);
          }
        }
        return emit_n(op == O_BAND ? PZK_N_AND : op == O_BOR ? PZK_N_OR : PZK_N_XOR, a, b, false);
      }
      case O_LT: case O_LE: {
        i128 ah, bh;
        if (exact_u(a, ah) && exact_u(b, bh)) {
          if (a.kind == 0) {  // c < x  -> need operand a as a value
            uint32_t ia = const_value_U(a.c);
            uint32_t id = new_value(CLS_U, 0, 1);
            emit(op == O_LT ? PZK_U_LT : PZK_U_LE, id, ia, b.id);
            stats->u_ops++;
            return sv(id);
          }
          return emit_u(op == O_LT ? PZK_U_LT : PZK_U_LE, a, b, CLS_U, 0, 1);
        }
        if (an && bn && alo >= I64_MIN_ && ahi <= I64_MAX_ && blo >= I64_MIN_ && bhi <= I64_MAX_) {
          uint32_t ia = u_operand(a), ib = u_operand(b);
          uint32_t id = new_value(CLS_U, 0, 1);
          emit(op == O_LT ? PZK_I_LT : PZK_I_LE, id, ia, ib);
          stats->u_ops++;
          return sv(id);
        }
        uint32_t ia = to_N(a), ib = to_N(b);
        uint32_t id = new_value(CLS_U, 0, 1);
        emit(op == O_LT ? PZK_N_SLT : PZK_N_SLE, id, ia, ib);
        stats->f_other++;
        return sv(id);
      }
      case O_EQ: case O_NE: {
        if (an && bn) {
          bool a_u = alo >= 0, b_u = blo >= 0;
          bool compat = (a_u && b_u) || (ahi <= I64_MAX_ && bhi <= I64_MAX_);
          if (compat) {
            if (a.kind == 0) return emit_u(op == O_EQ ? PZK_U_EQ : PZK_U_NE, b, a, CLS_U, 0, 1);
            return emit_u(op == O_EQ ? PZK_U_EQ : PZK_U_NE, a, b, CLS_U, 0, 1);
          }
        }
        const SVal& x = a.kind == 0 ? b : a;
        const SVal& y = a.kind == 0 ? a : b;
        uint32_t ia = to_F(x);
        uint32_t id;
        if (y.kind == 0) { uint32_t pi = pool_mont(y.c); id = new_value(CLS_U, 0, 1); emit(op == O_EQ ? PZK_F_EQ : PZK_F_NE, id, ia, pi, PZK_FLAG_B_POOL); }
        else { uint32_t ib = to_F(y); id = new_value(CLS_U, 0, 1); emit(op == O_EQ ? PZK_F_EQ : PZK_F_NE, id, ia, ib); }
        stats->f_other++;
        return sv(id);
      }
      case O_LAND: case O_LOR: {
        SVal ta = truthy(a, at), tb = truthy(b, at);
        if (ta.kind == 0) return (op == O_LAND) ? (ta.c.is_zero() ? ta : tb) : (ta.c.is_zero() ? tb : ta);
        if (tb.kind == 0) return (op == O_LAND) ? (tb.c.is_zero() ? tb : ta) : (tb.c.is_zero() ? ta : tb);
        return emit_u(op == O_LAND ? PZK_U_AND : PZK_U_OR, ta, tb, CLS_U, 0, 1);
      }
    }
    fail(at, "unsupported operator on signals");
  }

  SVal shift_right(const SVal& a, int k, const Stmt* at) {
    (void)at;
    if (k == 0) return a;
    i128 ah;
    if (exact_u(a, ah)) {
      if (k >= 64) return SVal::konst(U256());
      return emit_u(PZK_U_SHR, a, SVal::konst(U256((uint64_t)k)), CLS_U, 0, ah >> k);
    }
    SVal r = emit_n(PZK_N_SHR, a, SVal::konst(U256((uint64_t)k)), true);
    if (254 - k <= 64) return n_low(r, (((i128)1) << (254 - k)) - 1);
    return r;
  }

  SVal emit_un(int op, const SVal& a, const Stmt* at) {
    if (a.kind == 2) return SVal::unk();
    if (a.kind == 0) return SVal::konst(fold_un(op, a.c));
    if (phase == 0) return SVal::unk();
    if (op == O_NEG) return emit_bin(O_SUB, SVal::konst(U256()), a, at);
    if (op == O_LNOT) return emit_bin(O_EQ, a, SVal::konst(U256()), at);
    // ~x = (x ^ mask) mod p
    return emit_bin(O_BXOR, a, SVal::konst(MASK254), at);
  }

  SVal emit_select(const SVal& c, const SVal& x, const SVal& y, const Stmt* at) {
    if (c.kind == 2 || x.kind == 2 || y.kind == 2) return SVal::unk();
    if (c.kind == 0) return c.c.is_zero() ? y : x;
    // (v != 0) ? inv(v) : 0  is  inv(v)  (inv(0) = 0)
    if (y.kind == 0 && y.c.is_zero() && x.kind == 1 && c.kind == 1) {
      auto ix = inv_of.find(x.id);
      auto ic = nz_of.find(c.id);
      if (ix != inv_of.end() && ic != nz_of.end() && same_value(ix->second, ic->second)) return x;
    }
    // table
    {
      SVal xs[3] = {c, x, y};
      Table t;
      auto f = [&](const int64_t* v, int64_t& r) { r = v[0] != 0 ? v[1] : v[2]; return true; };
      if (try_table(xs, 3, t, f)) {
        int64_t cv;
        if (table_is_const(t, cv)) return SVal::konst(from_signed(cv));
        return sv(emit_table_value(t));
      }
    }
    SVal tc = truthy(c, at);
    i128 xlo, xhi, ylo, yhi;
    if (narrow_of(x, xlo, xhi) && narrow_of(y, ylo, yhi)) {
      i128 lo = std::min(xlo, ylo), hi = std::max(xhi, yhi);
      int cls = cls_for(lo, hi);
      if (cls >= 0) {
        uint32_t ix = u_operand(x), iy = u_operand(y);
        uint32_t id = new_value(cls, lo, hi);
        emit(PZK_U_SEL, id, tc.id, ix, PZK_FLAG_EXT, 0, iy);
        stats->u_ops++;
        return sv(id);
      }
    }
    uint32_t ix = to_F(x), iy = to_F(y);
    uint32_t id = new_value(CLS_F);
    emit(PZK_F_SEL, id, tc.id, ix, PZK_FLAG_EXT, 0, iy);
    stats->f_other++;
    return sv(id);
  }

  // ================================================================== algebra
  AlgP alg_const(const U256& c) { AlgP a = std::make_shared<Alg>(); a->deg = 0; a->c.k = c; return a; }
  AlgP alg_sig(uint32_t sig) { AlgP a = std::make_shared<Alg>(); a->deg = 1; a->c.t.emplace_back(sig, U256(1)); return a; }
  AlgP alg_bad() { AlgP a = std::make_shared<Alg>(); a->deg = 3; return a; }
  AlgP alg_of(const SVal& s) {
    if (s.alg) return s.alg;
    if (s.kind == 0) return alg_const(s.c);
    return alg_bad();
  }
  AlgP alg_bin(int op, const SVal& a, const SVal& b) {
    AlgP x = alg_of(a), y = alg_of(b);
    if (x->deg == 3 || y->deg == 3) {
      // x * 0 style cancellations are not tracked; anything else is not quadratic
      return alg_bad();
    }
    if (x->deg == 0 && y->deg == 0) { U256 r; if (fold_bin(op, x->c.k, y->c.k, r)) return alg_const(r); return alg_bad(); }
    switch (op) {
      case O_ADD: case O_SUB: {
        if (x->deg == 2 && y->deg == 2) return alg_bad();
        AlgP r = std::make_shared<Alg>(*(x->deg == 2 ? x : (y->deg == 2 ? y : x)));
        bool used_x = !(y->deg == 2 && x->deg != 2);
        U256 k = (op == O_ADD) ? U256(1) : fr_neg(U256(1));
        if (used_x) { lin_addmul(r->c, y->c, k); r->deg = std::max(x->deg, y->deg); }
        else {  // r = copy of y (quadratic), x is linear: x +/- y
          if (op == O_SUB) { lin_scale(r->a, k); lin_scale(r->c, k); }
          lin_addmul(r->c, x->c, U256(1)); r->deg = 2;
        }
        if (r->deg == 1 && r->c.t.empty()) r->deg = 0;
        return r;
      }
      case O_MUL: {
        if (x->deg == 0 || y->deg == 0) {
          const AlgP& k = x->deg == 0 ? x : y; const AlgP& v = x->deg == 0 ? y : x;
          AlgP r = std::make_shared<Alg>(*v);
          if (r->deg == 2) lin_scale(r->a, k->c.k);
          lin_scale(r->c, k->c.k);
          if (k->c.k.is_zero()) { r->deg = 0; r->a = Lin(); r->b = Lin(); }
          return r;
        }
        if (x->deg == 1 && y->deg == 1) {
          AlgP r = std::make_shared<Alg>(); r->deg = 2; r->a = x->c; r->b = y->c; return r;
        }
        return alg_bad();
      }
      case O_DIV: {
        if (y->deg == 0 && !y->c.k.is_zero()) {
          AlgP r = std::make_shared<Alg>(*x); U256 k = fr_inv(y->c.k);
          if (r->deg == 2) lin_scale(r->a, k);
          lin_scale(r->c, k); return r;
        }
        return alg_bad();
      }
    }
    return alg_bad();
  }

  uint32_t coef_id(const U256& c) {
    auto it = coef_idx.find(c);
    if (it != coef_idx.end()) return it->second;
    uint32_t i = (uint32_t)coefs.size(); coefs.push_back(c); coef_idx[c] = i; return i;
  }
  // signal index 0xFFFFFFFF denotes the constant-one wire
  void push_lin(const Lin& l, uint32_t& n) {
    n = 0;
    if (!l.k.is_zero()) { terms.emplace_back(0xFFFFFFFFu, coef_id(l.k)); n++; }
    for (auto& t : l.t) { terms.emplace_back(t.first, coef_id(t.second)); n++; }
  }
  // constraint  e == 0  where e = a*b + c   ->  A*B = -C
  void add_constraint(const AlgP& e, const Stmt* at, bool by_def = false) {
    if (e->deg == 3) fail(at, "non-quadratic constraint");
    if (e->deg == 0) {
      if (!e->c.k.is_zero()) fail(at, "constraint is constant and false");
      return;  // circom drops trivially true constraints
    }
    RowRec r; r.off = terms.size();
    Lin negc = e->c; lin_scale(negc, fr_neg(U256(1)));
    if (e->deg == 2) { push_lin(e->a, r.na); push_lin(e->b, r.nb); }
    else { r.na = r.nb = 0; }
    push_lin(negc, r.nc);
    r.by_def = by_def;
    if (at) { r.line = at->line; r.file = at->file; }
    r.tname = cur_tname;
    rows.push_back(r);
  }

  // ================================================================== environment
  Value* find_var(Env& env, int name) {
    for (size_t i = env.size(); i-- > 0;)
      for (size_t j = env[i].vars.size(); j-- > 0;)
        if (env[i].vars[j].first == name) return &env[i].vars[j].second;
    return nullptr;
  }
  void declare_var(Env& env, int name, const Value& v) {
    for (auto& p : env.back().vars) if (p.first == name) { p.second = v; return; }
    env.back().vars.emplace_back(name, v);
  }
  static Value scalar(const SVal& s) { Value v; v.s = s; return v; }
  static Value zeros(const std::vector<int>& dims) {
    Value v;
    if (dims.empty()) return v;
    v.arr = true; v.a = std::make_shared<AVal>(); v.a->dims = dims;
    size_t n = 1; for (int d : dims) n *= (size_t)d;
    v.a->v.assign(n, SVal());
    return v;
  }
  static bool has_unk(const Value& v) {
    if (!v.arr) return v.s.kind == 2;
    for (auto& s : v.a->v) if (s.kind == 2) return true;
    return false;
  }
  static bool all_const(const Value& v) {
    if (!v.arr) return v.s.kind == 0;
    for (auto& s : v.a->v) if (s.kind != 0) return false;
    return true;
  }
  std::string value_key(const Value& v) {
    std::string k;
    if (!v.arr) { k.append((const char*)v.s.c.w, 32); return k; }
    k.push_back('[');
    for (int d : v.a->dims) { k += std::to_string(d); k.push_back(','); }
    for (auto& s : v.a->v) k.append((const char*)s.c.w, 32);
    k.push_back(']');
    return k;
  }
  int const_int(const SVal& s, const Stmt* at, const char* what) {
    if (s.kind == 2) fail(at, std::string(what) + " depends on a signal");
    if (s.kind != 0) fail(at, std::string(what) + " is not a compile-time constant");
    if (!s.c.fits64() || s.c.w[0] > 0x7fffffff) fail(at, std::string(what) + " out of range");
    return (int)s.c.w[0];
  }

  // ================================================================== layouts (phase A)
  Layout* layout_of(int tname, const std::vector<Value>& args, const Stmt* at) {
    std::string key = nm(tname);
    key.push_back('(');
    for (auto& a : args) { key += value_key(a); key.push_back(';'); }
    auto it = layouts.find(key);
    if (it != layouts.end()) return it->second;
    auto tt = unit.templates.find(tname);
    if (tt == unit.templates.end()) fail(at, "unknown template " + nm(tname));
    TemplateDef& td = tt->second;
    if (td.params.size() != args.size()) fail(at, "wrong number of arguments for template " + nm(tname));
    Layout* lay = new Layout();
    lay->tname = tname; lay->args = args;
    layout_list.push_back(lay);
    int saved = phase; phase = 0;
    Env env(1);
    for (size_t i = 0; i < args.size(); i++) declare_var(env, td.params[i], args[i]);
    Ctx ctx; ctx.lay = lay;
    bool sr = returned; returned = false;
    exec(td.body, env, &ctx);
    returned = sr;
    phase = saved;
    uint32_t off = 0;
    for (int kind : {2, 1, 0})
      for (int n : lay->order) {
        SigInfo& s = lay->sigs[n];
        if (s.kind != kind) continue;
        s.off = off;
        uint32_t cnt = 1; for (int d : s.dims) cnt *= (uint32_t)d;
        off += cnt;
        if (kind == 1) lay->n_inputs += cnt;
      }
    lay->own = off;
    for (auto& ck : lay->child_order) { Child& c = lay->children[ck]; c.rel_base = off; off += c.lay->total; }
    lay->total = off;
    layouts[key] = lay;
    return lay;
  }

  // ================================================================== references
  static uint32_t prod(const std::vector<int>& d, size_t from = 0) {
    uint32_t n = 1; for (size_t i = from; i < d.size(); i++) n *= (uint32_t)d[i]; return n;
  }
  Ref resolve(Expr* e, Env& env, Ctx* ctx, const Stmt* at) {
    Ref r;
    switch (e->k) {
      case Expr::VAR: {
        if (find_var(env, e->name)) { r.k = Ref::VAR; return r; }
        if (!ctx) fail(at, "unknown identifier " + nm(e->name));
        Layout* lay = ctx->lay;
        auto si = lay->sigs.find(e->name);
        if (si != lay->sigs.end()) {
          r.k = Ref::SIG; r.comp = ctx->comp; r.lay = lay; r.off = si->second.off; r.dims = si->second.dims;
          r.skind = si->second.kind; r.name = e->name; return r;
        }
        auto ci = lay->comp_dims.find(e->name);
        if (ci != lay->comp_dims.end()) { r.k = Ref::COMP; r.name = e->name; r.dims = ci->second; return r; }
        fail(at, "unknown identifier " + nm(e->name));
      }
      case Expr::IDX: {
        r = resolve(e->a, env, ctx, at);
        if (r.k == Ref::VAR || r.k == Ref::UNKNOWN) return r;
        Value iv = eval(e->b, env, ctx, false, at);
        int i = const_int(iv.s, at, "signal/component index");
        if (r.k == Ref::SIG) {
          if (r.dims.empty()) fail(at, "too many indices on signal " + nm(r.name));
          if (i >= r.dims[0]) fail(at, fmt("index %d out of range for signal %s (dim %d)", i, nm(r.name).c_str(), r.dims[0]));
          r.off += (uint32_t)i * prod(r.dims, 1);
          r.dims.erase(r.dims.begin());
          return r;
        }
        size_t used = r.path.size();
        if (used >= r.dims.size()) fail(at, "too many indices on component " + nm(r.name));
        if (i >= r.dims[used]) fail(at, "component index out of range for " + nm(r.name));
        r.path.push_back(i);
        return r;
      }
      case Expr::MEM: {
        Ref b = resolve(e->a, env, ctx, at);
        if (b.k != Ref::COMP || b.path.size() != b.dims.size()) fail(at, "member access on a non-component");
        if (phase == 0) { r.k = Ref::UNKNOWN; return r; }
        Comp* child = get_child(ctx->comp, b, at);
        auto si = child->lay->sigs.find(e->name);
        if (si == child->lay->sigs.end()) fail(at, "component has no signal " + nm(e->name));
        r.k = Ref::SIG; r.comp = child; r.lay = child->lay; r.off = si->second.off; r.dims = si->second.dims;
        r.skind = si->second.kind; r.name = e->name;
        return r;
      }
      default: fail(at, "expression is not assignable");
    }
  }
  static int flat_index(const Ref& c) {
    int f = 0;
    for (size_t i = 0; i < c.path.size(); i++) f = f * c.dims[i] + c.path[i];
    return f;
  }
  Comp* get_child(Comp* comp, const Ref& c, const Stmt* at) {
    auto it = comp->kids.find({c.name, flat_index(c)});
    if (it == comp->kids.end()) fail(at, "component " + nm(c.name) + " used before it was instantiated");
    return it->second;
  }

  // ================================================================== signals
  SVal read_signal_elem(Comp* comp, uint32_t off, bool alg) {
    uint32_t sig = comp->base + off;
    uint32_t vid = sig_val[sig];
    SVal s;
    if (vid == 0) s = SVal::konst(U256());
    else if (v_const[vid]) s = SVal::konst(const_of[vid]);
    else s = sv(vid);
    if (alg) s.alg = alg_sig(sig);
    return s;
  }
  Value read_signal(const Ref& r, bool alg) {
    if (phase == 0) return scalar(SVal::unk());
    if (r.dims.empty()) return scalar(read_signal_elem(r.comp, r.off, alg));
    Value v; v.arr = true; v.a = std::make_shared<AVal>(); v.a->dims = r.dims;
    uint32_t n = prod(r.dims);
    v.a->v.reserve(n);
    for (uint32_t i = 0; i < n; i++) v.a->v.push_back(read_signal_elem(r.comp, r.off + i, alg));
    return v;
  }
  void store_signal(const Ref& r, const Value& v, bool constrain, const Stmt* at) {
    uint32_t n = prod(r.dims);
    if (r.dims.empty() && v.arr) fail(at, "array assigned to scalar signal " + nm(r.name));
    if (!r.dims.empty() && (!v.arr || v.a->v.size() != n)) fail(at, "array size mismatch assigning signal " + nm(r.name));
    for (uint32_t i = 0; i < n; i++) {
      const SVal& s = r.dims.empty() ? v.s : v.a->v[i];
      uint32_t sig = r.comp->base + r.off + i;
      if (sig_val[sig] != 0) fail(at, "signal " + nm(r.name) + " assigned twice");
      uint32_t vid;
      if (s.kind == 0) vid = const_value(s.c);
      else if (s.kind == 1) {
        vid = s.id;
        if (v_cls[vid] == CLS_N) vid = to_F(s);
      } else fail(at, "internal: unknown value stored to a signal");
      sig_val[sig] = vid;
      if (constrain) {
        // rhs - sig == 0
        SVal self; self.kind = 1; self.alg = alg_sig(sig);
        AlgP e = alg_bin(O_SUB, s, self);
        add_constraint(e, at, true);
      }
    }
    if (r.skind == 1 && r.comp->parent != nullptr) {
      r.comp->pending -= n;
      if (r.comp->pending == 0) run_comp(r.comp, at);
    }
  }

  void run_comp(Comp* comp, const Stmt* at) {
    if (comp->ran) fail(at, "component executed twice");
    comp->ran = true;
    TemplateDef& td = unit.templates[comp->lay->tname];
    Env env(1);
    for (size_t i = 0; i < td.params.size(); i++) declare_var(env, td.params[i], comp->lay->args[i]);
    Ctx ctx; ctx.lay = comp->lay; ctx.comp = comp;
    bool sr = returned; returned = false;
    int saved_t = cur_tname; cur_tname = comp->lay->tname;
    bool hinted = phase == 1 && opt.intrinsics && try_bjj_hints(comp);
    exec(td.body, env, &ctx);
    if (hinted) {
      if (hint_stack.back().next != hint_stack.back().vals.size()) fail(at, "internal: BabyJubjub hint count does not match the template");
      hint_stack.pop_back();
    }
    cur_tname = saved_t;
    returned = sr;
  }

  // BabyjubjubBase8Multiplication (/root/reference/circuits/lib/circuits/babyjubjub/curve.circom:143-171): 507
  // BabyjubjubAdd instances in one dependent chain, each with two `<--` field divisions (:97,101).  The BJJ_MUL8
  // record computes all their outputs first (projective ladder + one batched inversion on the device); the template
  // body then runs as written, except that those two hints take the precomputed values instead of dividing.  Every
  // constraint of the template is still emitted and checked, so a wrong hint could only fail a row, never pass.
  struct HintCtx { std::vector<uint32_t> vals; size_t next = 0; };
  std::vector<HintCtx> hint_stack;
  int id_bjj_mul8 = -2, id_bjj_add = -2, id_sig_out = -2, id_sig_scalar = -2;
  bool try_bjj_hints(Comp* comp) {
    if (id_bjj_mul8 == -2) {
      id_bjj_mul8 = unit.names.get("BabyjubjubBase8Multiplication"); id_bjj_add = unit.names.get("BabyjubjubAdd");
      id_sig_out = unit.names.get("out"); id_sig_scalar = unit.names.get("scalar");
    }
    if (comp->lay->tname != id_bjj_mul8) return false;
    auto si = comp->lay->sigs.find(id_sig_scalar);
    if (si == comp->lay->sigs.end() || si->second.kind != 1 || !si->second.dims.empty()) return false;
    uint32_t vid = sig_val[comp->base + si->second.off];
    if (!vid || v_const[vid]) return false;
    uint32_t sf = to_F(sv(vid));
    if (is_pending(sf)) flush_inversions();
    const uint32_t nbits = 254, nadd = 2 * nbits - 1;
    // curve constants of the template family (curve.circom:84-85, get.circom:9-10)
    auto dec = [](const char* t) { return parse_number(t, strlen(t)); };
    U256 ca = dec("168700"), cd = dec("168696");
    U256 bx = dec("5299619240641551281634865583518297030282874472190772894086521144482721001553");
    U256 by = dec("16950150798460657717958625567821834550301663161624707787222815936182638968203");
    uint32_t off = (uint32_t)list_pool.size();
    list_pool.push_back(nbits);
    list_pool.push_back(pool_mont(ca)); list_pool.push_back(pool_mont(cd)); list_pool.push_back(pool_mont(bx)); list_pool.push_back(pool_mont(by));
    list_pool.push_back(sf);
    HintCtx h;
    uint32_t first = 0;
    for (uint32_t i = 0; i < 4 * nadd; i++) {
      uint32_t id = new_value(CLS_F);
      if (i == 0) first = id;
      list_pool.push_back(id);
      if (i < 2 * nadd) h.vals.push_back(id);
    }
    emit(PZK_BJJ_MUL8, first, off, 0);
    for (uint32_t i = 0; i < 4 * nadd; i++) v_def[list_pool[off + 6 + i]] = (uint32_t)ops.size() - 1;
    stats->bjj++;
    hint_stack.push_back(std::move(h));
    return true;
  }

  void instantiate(const Ref& c, Expr* rhs, Env& env, Ctx* ctx, const Stmt* at) {
    if (c.path.size() != c.dims.size()) fail(at, "component array assigned as a whole");
    if (rhs->k != Expr::CALL) fail(at, "component initialiser must be a template call");
    std::vector<Value> args;
    for (Expr* a : rhs->args) args.push_back(eval(a, env, ctx, false, at));
    for (auto& a : args) {
      if (has_unk(a)) fail(at, "template argument depends on a signal");
      if (!all_const(a)) fail(at, "template argument is not a compile-time constant");
    }
    std::pair<int, int> key{c.name, flat_index(c)};
    if (phase == 0) {
      Layout* lay = ctx->lay;
      if (lay->children.count(key)) fail(at, "component " + nm(c.name) + " instantiated twice");
      Child ch; ch.tname = rhs->name; ch.lay = layout_of(rhs->name, args, at); ch.rel_base = 0;
      lay->children[key] = ch;
      lay->child_order.push_back(key);
      return;
    }
    Comp* comp = ctx->comp;
    auto it = comp->lay->children.find(key);
    if (it == comp->lay->children.end() || it->second.tname != rhs->name) fail(at, "phase mismatch instantiating " + nm(c.name));
    Layout* sub = layout_of(rhs->name, args, at);
    if (sub != it->second.lay) fail(at, "phase mismatch (arguments) instantiating " + nm(c.name));
    Comp* child = new Comp();
    comps.push_back(child);
    child->lay = sub; child->base = comp->base + it->second.rel_base; child->pending = sub->n_inputs; child->parent = comp;
    comp->kids[key] = child;
    if (child->pending == 0) run_comp(child, at);
  }

  // ================================================================== expressions
  Value eval(Expr* e, Env& env, Ctx* ctx, bool alg, const Stmt* at) {
    switch (e->k) {
      case Expr::NUM: return scalar(SVal::konst(e->num));
      case Expr::VAR: {
        Value* v = find_var(env, e->name);
        if (v) return *v;
        Ref r = resolve(e, env, ctx, at);
        if (r.k == Ref::SIG) return read_signal(r, alg);
        fail(at, "component " + nm(e->name) + " used as a value");
      }
      case Expr::IDX: {
        // var array fast path
        Expr* base = e; int depth = 0;
        while (base->k == Expr::IDX) { base = base->a; depth++; }
        if (base->k == Expr::VAR) {
          Value* v = find_var(env, base->name);
          if (v) {
            if (!v->arr) { if (v->s.kind == 2) return scalar(SVal::unk()); fail(at, "indexing scalar variable " + nm(base->name)); }
            std::vector<int> idx(depth);
            Expr* w = e;
            for (int i = depth - 1; i >= 0; i--) {
              Value iv = eval(w->b, env, ctx, false, at);
              if (iv.s.kind == 2) return scalar(SVal::unk());
              idx[i] = const_int(iv.s, at, "array index");
              w = w->a;
            }
            AVal& a = *v->a;
            if ((size_t)depth > a.dims.size()) fail(at, "too many indices on " + nm(base->name));
            size_t off = 0;
            for (int i = 0; i < depth; i++) {
              if (idx[i] >= a.dims[i]) fail(at, fmt("index %d out of range for variable %s (dim %d)", idx[i], nm(base->name).c_str(), a.dims[i]));
              off = off * (size_t)a.dims[i] + (size_t)idx[i];
            }
            if ((size_t)depth == a.dims.size()) return scalar(a.v[off]);
            size_t stride = 1; for (size_t i = depth; i < a.dims.size(); i++) stride *= (size_t)a.dims[i];
            Value out; out.arr = true; out.a = std::make_shared<AVal>();
            out.a->dims.assign(a.dims.begin() + depth, a.dims.end());
            out.a->v.assign(a.v.begin() + off * stride, a.v.begin() + (off + 1) * stride);
            return out;
          }
        }
        Ref r = resolve(e, env, ctx, at);
        if (r.k == Ref::UNKNOWN) return scalar(SVal::unk());
        if (r.k == Ref::SIG) return read_signal(r, alg);
        fail(at, "component used as a value");
      }
      case Expr::MEM: {
        Ref r = resolve(e, env, ctx, at);
        if (r.k == Ref::UNKNOWN) return scalar(SVal::unk());
        return read_signal(r, alg);
      }
      case Expr::CALL: return call_function(e, env, ctx, at);
      case Expr::ARR: {
        Value out; out.arr = true; out.a = std::make_shared<AVal>();
        std::vector<int> sub;
        for (size_t i = 0; i < e->args.size(); i++) {
          Value x = eval(e->args[i], env, ctx, alg, at);
          if (x.arr) { if (i == 0) sub = x.a->dims; out.a->v.insert(out.a->v.end(), x.a->v.begin(), x.a->v.end()); }
          else out.a->v.push_back(x.s);
        }
        out.a->dims.push_back((int)e->args.size());
        out.a->dims.insert(out.a->dims.end(), sub.begin(), sub.end());
        return out;
      }
      case Expr::UN: {
        Value a = eval(e->a, env, ctx, alg, at);
        if (a.arr) fail(at, "unary operator on an array");
        SVal r = emit_un(e->op, a.s, at);
        if (alg && r.kind != 2) {
          if (e->op == O_NEG) r.alg = alg_bin(O_SUB, SVal::konst(U256()), a.s);
          else r.alg = (r.kind == 0) ? nullptr : alg_bad();
        }
        return scalar(r);
      }
      case Expr::BIN: {
        Value a = eval(e->a, env, ctx, alg, at);
        if (a.arr) fail(at, "binary operator on an array");
        // short circuit on constants
        if (a.s.kind == 0) {
          if (e->op == O_LAND && a.s.c.is_zero()) return scalar(SVal::konst(U256()));
          if (e->op == O_LOR && !a.s.c.is_zero()) return scalar(SVal::konst(U256(1)));
        }
        Value b = eval(e->b, env, ctx, alg, at);
        if (b.arr) fail(at, "binary operator on an array");
        SVal r = fold_or_emit_bin(e->op, a.s, b.s, at);
        if (alg && r.kind != 2) r.alg = alg_bin(e->op, a.s, b.s);
        return scalar(r);
      }
      case Expr::TERN: {
        Value c = eval(e->a, env, ctx, false, at);
        if (c.arr) fail(at, "array used as a condition");
        if (c.s.kind == 2) return scalar(SVal::unk());
        if (c.s.kind == 0) return eval(c.s.c.is_zero() ? e->c : e->b, env, ctx, alg, at);
        Value x = eval(e->b, env, ctx, false, at), y = eval(e->c, env, ctx, false, at);
        if (x.arr || y.arr) fail(at, "signal-dependent selection between arrays");
        SVal r = emit_select(c.s, x.s, y.s, at);
        if (alg) r.alg = (r.kind == 0) ? nullptr : alg_bad();
        return scalar(r);
      }
    }
    fail(at, "bad expression");
  }

  // ---- functions -----------------------------------------------------------------
  int id_long_div = -2, id_mod_inv = -2;
  Value call_function(Expr* e, Env& env, Ctx* ctx, const Stmt* at) {
    auto ft = unit.functions.find(e->name);
    if (ft == unit.functions.end()) fail(at, "unknown function " + nm(e->name));
    TemplateDef& fd = ft->second;
    std::vector<Value> args;
    bool unk = false, konst = true;
    for (Expr* a : e->args) {
      args.push_back(eval(a, env, ctx, false, at));
      if (has_unk(args.back())) unk = true;
      if (!all_const(args.back())) konst = false;
    }
    if (unk) return scalar(SVal::unk());
    if (fd.params.size() != args.size()) fail(at, "wrong number of arguments for function " + nm(e->name));
    std::string key;
    if (konst) {
      key = nm(e->name); key.push_back('(');
      for (auto& a : args) { key += value_key(a); key.push_back(';'); }
      auto it = fn_memo.find(key);
      if (it != fn_memo.end()) return it->second;
    } else if (opt.intrinsics) {
      if (id_long_div == -2) id_long_div = unit.names.get("long_div");
      Value out;
      if (e->name == id_long_div && try_bigdiv(args, out, at)) return out;
      if (id_mod_inv == -2) id_mod_inv = unit.names.get("mod_inv");
      if (e->name == id_mod_inv && try_modinv(args, out, at)) return out;
    }
    Env fenv(1);
    for (size_t i = 0; i < args.size(); i++) declare_var(fenv, fd.params[i], args[i]);
    bool sr = returned; Value sv_ret = ret_val;
    returned = false;
    if (++fn_depth > 400) fail(at, "function recursion too deep");
    early_returns.emplace_back();
    exec(fd.body, fenv, nullptr);
    fn_depth--;
    if (!returned) fail(at, "function " + nm(e->name) + " did not return");
    Value out = ret_val;
    {
      std::vector<std::pair<SVal, Value>> er;
      er.swap(early_returns.back());
      early_returns.pop_back();
      for (size_t k = er.size(); k-- > 0;) out = select_value(er[k].first, er[k].second, out, at);
    }
    returned = sr; ret_val = sv_ret;
    if (konst) fn_memo[key] = out;
    return out;
  }

  // a limb that is not provably < 2^n is narrowed at run time: its low n bits + assertions that nothing was cut off
  SVal narrow_limb_checked(const SVal& s_, uint64_t n) {
    const i128 lim = n == 64 ? U64_MAX_ : (((i128)1 << n) - 1);
    i128 hi;
    if (exact_u(s_, hi) && hi <= lim) return s_;
    if (s_.kind == 0) return SVal::unk();
    SVal low;
    if (exact_u(s_, hi)) low = s_;   // a 64-bit word wider than the limb
    else {
      SVal nn = sv(to_N(s_));
      uint32_t fits = new_value(CLS_U, 0, 1);
      emit(PZK_N_FITS, fits, nn.id);
      emit(PZK_ASSERT_NZ, 0, fits);
      low = n_low(nn, U64_MAX_);
    }
    if (n == 64) return low;
    SVal top = emit_u(PZK_U_SHR, low, SVal::konst(U256(n)), CLS_U, 0, (i128)(U64_MAX_ >> n));
    SVal none = emit_u(PZK_U_EQ, top, SVal::konst(U256()), CLS_U, 0, 1);
    emit(PZK_ASSERT_NZ, 0, none.id);
    return emit_u(PZK_U_AND, low, SVal::konst(U256((uint64_t)lim)), CLS_U, 0, lim);
  }

  // long_div(n, k, m, a, b) of /root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom:190-232
  // evaluated natively: a has k+m limbs of n bits, b has k limbs (top limb non-zero);
  // returns out[2][200] with the m+1 quotient limbs and k remainder limbs.
  bool try_bigdiv(const std::vector<Value>& args, Value& out, const Stmt* at) {
    if (args.size() != 5) return false;
    for (int i = 0; i < 3; i++) if (args[i].arr || args[i].s.kind != 0 || !args[i].s.c.fits64()) return false;
    uint64_t n = args[0].s.c.w[0], k = args[1].s.c.w[0], m = args[2].s.c.w[0];
    if ((n != 64 && n != 32) || k < 2 || k > 64 || k + m > 128 || (n == 32 && k < 3) || !args[3].arr || !args[4].arr) return false;
    if (args[3].a->v.size() < k + m || args[4].a->v.size() < k) return false;
    i128 lim = n == 64 ? U64_MAX_ : (((i128)1 << n) - 1);
    auto narrow_limb = [&](const SVal& s_) -> SVal { return narrow_limb_checked(s_, n); };
    std::vector<SVal> la(k + m), lb(k);
    for (uint64_t i = 0; i < k + m; i++) { la[i] = narrow_limb(args[3].a->v[i]); if (la[i].kind == 2) return false; }
    for (uint64_t i = 0; i < k; i++) { lb[i] = narrow_limb(args[4].a->v[i]); if (lb[i].kind == 2) return false; }
    // BIGDIV reads its operands through the list pool: none of them may still be deferred
    for (uint64_t i = 0; i < k + m; i++) if (la[i].kind == 1 && is_pending(la[i].id)) flush_inversions();
    for (uint64_t i = 0; i < k; i++) if (lb[i].kind == 1 && is_pending(lb[i].id)) flush_inversions();
    uint32_t off = (uint32_t)list_pool.size();
    list_pool.push_back((uint32_t)n); list_pool.push_back((uint32_t)k); list_pool.push_back((uint32_t)m);
    for (uint64_t i = 0; i < k + m; i++) list_pool.push_back(u_operand(la[i]));
    for (uint64_t i = 0; i < k; i++) list_pool.push_back(u_operand(lb[i]));
    out = zeros({2, 200});
    uint32_t first = 0;
    for (uint64_t i = 0; i <= m; i++) {
      uint32_t id = new_value(CLS_U, 0, lim);
      if (i == 0) first = id;
      list_pool.push_back(id);
      out.a->v[i] = sv(id);
    }
    for (uint64_t i = 0; i < k; i++) {
      uint32_t id = new_value(CLS_U, 0, lim);
      list_pool.push_back(id);
      out.a->v[200 + i] = sv(id);
    }
    emit(PZK_BIGDIV, first, off, 0);
    for (uint64_t i = 0; i < (m + 1) + k; i++) v_def[list_pool[off + 3 + (k + m) + k + i]] = (uint32_t)ops.size() - 1;
    stats->bigdiv++;
    (void)at;
    return true;
  }

  // (a * b) mod m and Miller-Rabin on 256-bit values, compile time only
  static U256 mulmod256(const U256& a, const U256& b, const U256& m) {
    U256 r, x = a;
    { U256 q, t; divmod(x, m, q, t); x = t; }
    for (int i = 0; i < 256; i++) {
      if ((b.w[i >> 6] >> (i & 63)) & 1) {
        uint64_t c; U256 t = add(r, x, &c);
        if (c || !(t < m)) t = sub(t, m);
        r = t;
      }
      uint64_t c; U256 t = add(x, x, &c);
      if (c || !(t < m)) t = sub(t, m);
      x = t;
    }
    return r;
  }
  static bool is_odd_prime(const U256& n) {
    if (!(n.w[0] & 1) || n < U256(5)) return n == U256(3);
    U256 d = sub(n, U256(1)), nm1 = d;
    int s_ = 0;
    while (!(d.w[0] & 1)) { d = shr(d, 1); s_++; }
    for (uint64_t base : {2ull, 3ull, 5ull, 7ull, 11ull, 13ull, 17ull, 19ull, 23ull, 29ull, 31ull, 37ull}) {
      U256 x(1), b(base);
      for (int i = 255; i >= 0; i--) {
        x = mulmod256(x, x, n);
        if ((d.w[i >> 6] >> (i & 63)) & 1) x = mulmod256(x, b, n);
      }
      if (x == U256(1) || x == nm1) continue;
      bool comp = true;
      for (int r = 1; r < s_ && comp; r++) { x = mulmod256(x, x, n); if (x == nm1) comp = false; }
      if (comp) return false;
    }
    return true;
  }
  std::unordered_map<U256, bool, U256Hash> prime_memo;

  // mod_inv(n, k, a, p) of /root/reference/circuits/lib/circuits/bigInt/bigIntFunc.circom:430-465 computes
  // a^(p-2) mod p by square-and-multiply (about 380 prod + long_div rounds, 100 k records unrolled).
  // When p is a compile-time constant odd prime that is the modular inverse of a (0 when p | a): one
  // MODINV record, evaluated with a binary extended GCD on the device.
  bool try_modinv(const std::vector<Value>& args, Value& out, const Stmt* at) {
    if (args.size() != 4) return false;
    for (int i = 0; i < 2; i++) if (args[i].arr || args[i].s.kind != 0 || !args[i].s.c.fits64()) return false;
    uint64_t n = args[0].s.c.w[0], k = args[1].s.c.w[0];
    if ((n != 64 && n != 32) || k < 1 || k * n > 256 || !args[2].arr || !args[3].arr) return false;
    if (args[2].a->v.size() < k || args[3].a->v.size() < k) return false;
    U256 pm;
    for (uint64_t i = 0; i < k; i++) {
      const SVal& q = args[3].a->v[i];
      if (q.kind != 0 || !q.c.fits64() || (n < 64 && (q.c.w[0] >> n))) return false;
      pm = add(pm, shl(U256(q.c.w[0]), (int)(i * n)));
    }
    auto it = prime_memo.find(pm);
    bool prime = it != prime_memo.end() ? it->second : (prime_memo[pm] = is_odd_prime(pm));
    if (!prime) return false;
    i128 lim = n == 64 ? U64_MAX_ : (((i128)1 << n) - 1);
    std::vector<SVal> la(k);
    bool all_const = true;
    for (uint64_t i = 0; i < k; i++) {
      const SVal& s_ = args[2].a->v[i];
      if (s_.kind == 2) return false;
      if (s_.kind == 0 && !s_.c.fits64()) return false;
      if (s_.kind != 0) all_const = false;
    }
    if (all_const) return false;  // the unrolled function folds at compile time
    for (uint64_t i = 0; i < k; i++) {
      const SVal& s_ = args[2].a->v[i];
      if (s_.kind == 0) { if (n < 64 && (s_.c.w[0] >> n)) return false; la[i] = s_; continue; }
      // not provably an n-bit limb (the difference limbs of long_sub): narrowed under an assertion
      la[i] = narrow_limb_checked(s_, n);
    }
    for (uint64_t i = 0; i < k; i++) if (la[i].kind == 1 && is_pending(la[i].id)) flush_inversions();
    uint32_t off = (uint32_t)list_pool.size();
    list_pool.push_back((uint32_t)n); list_pool.push_back((uint32_t)k); list_pool.push_back(0);
    for (uint64_t i = 0; i < k; i++) list_pool.push_back(u_operand(la[i]));
    for (uint64_t i = 0; i < k; i++) list_pool.push_back(u_operand(args[3].a->v[i]));
    out = zeros({200});
    uint32_t first = new_value(CLS_U, 0, 0);  // the BIGDIV layout's quotient limb: always 0 here
    list_pool.push_back(first);
    for (uint64_t i = 0; i < k; i++) {
      uint32_t id = new_value(CLS_U, 0, lim);
      list_pool.push_back(id);
      out.a->v[i] = sv(id);
    }
    emit(PZK_MODINV, first, off, 0);
    for (uint64_t i = 0; i < 1 + k; i++) v_def[list_pool[off + 3 + k + k + i]] = (uint32_t)ops.size() - 1;
    stats->modinv++;
    (void)at;
    return true;
  }

  // ================================================================== statements
  bool relevant_A(Stmt* s) {  // does the statement matter in phase A?
    if (!s) return false;
    if (s->has_decl >= 0) return s->has_decl;
    bool r = false;
    switch (s->k) {
      case Stmt::BLOCK: for (Stmt* c : s->stmts) if (relevant_A(c)) r = true; break;
      case Stmt::SIGDECL: case Stmt::VARDECL: case Stmt::COMPDECL: case Stmt::INCDEC: case Stmt::RETURN: case Stmt::ASSERT: r = true; break;
      case Stmt::ASSIGN: r = (s->op == O_ASSIGN); break;
      case Stmt::IF: r = relevant_A(s->s1) || relevant_A(s->s2); break;
      case Stmt::FOR: r = relevant_A(s->body); break;
      case Stmt::WHILE: r = true; break;
      default: r = false;
    }
    s->has_decl = r;
    return r;
  }
  void collect_assigned(Stmt* s, std::vector<int>& out) {
    if (!s) return;
    switch (s->k) {
      case Stmt::BLOCK: for (Stmt* c : s->stmts) collect_assigned(c, out); break;
      case Stmt::ASSIGN: case Stmt::INCDEC: { Expr* b = s->lhs; while (b->k == Expr::IDX) b = b->a; if (b->k == Expr::VAR) out.push_back(b->name); break; }
      case Stmt::IF: collect_assigned(s->s1, out); collect_assigned(s->s2, out); break;
      case Stmt::FOR: collect_assigned(s->s1, out); collect_assigned(s->s2, out); collect_assigned(s->body, out); break;
      case Stmt::WHILE: collect_assigned(s->body, out); break;
      case Stmt::SIGDECL: case Stmt::COMPDECL: fail(s, "declaration under a signal-dependent condition");
      default: break;
    }
  }

  void assign_var(Expr* lhs, const Value& v, Env& env, Ctx* ctx, const Stmt* at) {
    if (lhs->k == Expr::VAR) {
      Value* slot = find_var(env, lhs->name);
      if (!slot) fail(at, "assignment to unknown variable " + nm(lhs->name));
      *slot = v;
      return;
    }
    if (lhs->k != Expr::IDX) fail(at, "bad assignment target");
    Expr* base = lhs; int depth = 0;
    while (base->k == Expr::IDX) { base = base->a; depth++; }
    if (base->k != Expr::VAR) fail(at, "bad assignment target");
    Value* slot = find_var(env, base->name);
    if (!slot) fail(at, "assignment to unknown variable " + nm(base->name));
    std::vector<int> idx(depth);
    Expr* w = lhs;
    bool unk = false;
    for (int i = depth - 1; i >= 0; i--) {
      Value iv = eval(w->b, env, ctx, false, at);
      if (iv.s.kind == 2) unk = true; else idx[i] = const_int(iv.s, at, "array index");
      w = w->a;
    }
    if (!slot->arr) { if (slot->s.kind == 2) return; fail(at, "indexing scalar variable " + nm(base->name)); }
    if (unk) { *slot = scalar(SVal::unk()); return; }
    if (slot->a.use_count() > 1) slot->a = std::make_shared<AVal>(*slot->a);  // copy on write
    AVal& a = *slot->a;
    if ((size_t)depth > a.dims.size()) fail(at, "too many indices on " + nm(base->name));
    size_t off = 0;
    for (int i = 0; i < depth; i++) {
      if (idx[i] >= a.dims[i]) fail(at, fmt("index %d out of range for variable %s (dim %d)", idx[i], nm(base->name).c_str(), a.dims[i]));
      off = off * (size_t)a.dims[i] + (size_t)idx[i];
    }
    if ((size_t)depth == a.dims.size()) {
      if (v.arr) fail(at, "array assigned to an array element");
      a.v[off] = v.s;
      return;
    }
    size_t stride = 1; for (size_t i = depth; i < a.dims.size(); i++) stride *= (size_t)a.dims[i];
    if (!v.arr) fail(at, "scalar assigned to a sub-array");
    size_t n = std::min(stride, v.a->v.size());
    for (size_t i = 0; i < n; i++) a.v[off * stride + i] = v.a->v[i];
  }

  void exec_assign(Stmt* s, Env& env, Ctx* ctx) {
    int op = s->op;
    if (phase == 0 && op != O_ASSIGN) return;
    Ref target = resolve(s->lhs, env, ctx, s);
    if (target.k == Ref::VAR) {
      if (op != O_ASSIGN) fail(s, "signal operator used on a variable");
      Value v = eval(s->rhs, env, ctx, ctx != nullptr && phase == 1, s);
      assign_var(s->lhs, v, env, ctx, s);
      return;
    }
    if (target.k == Ref::COMP) {
      if (op != O_ASSIGN) fail(s, "components are assigned with =");
      instantiate(target, s->rhs, env, ctx, s);
      return;
    }
    if (target.k == Ref::UNKNOWN) return;
    if (op == O_ASSIGN) fail(s, "signals are assigned with <== or <--");
    if (op == O_WITNESS_L && phase == 1 && !hint_stack.empty() && ctx && ctx->lay->tname == id_bjj_add && target.name == id_sig_out &&
        target.dims.empty()) {
      HintCtx& h = hint_stack.back();
      if (h.next >= h.vals.size()) fail(s, "internal: more BabyjubjubAdd instances than hints");
      store_signal(target, scalar(sv(h.vals[h.next++])), false, s);
      return;
    }
    Value v = eval(s->rhs, env, ctx, op == O_CONSTRAIN_L, s);
    store_signal(target, v, op == O_CONSTRAIN_L, s);
  }

  void exec(Stmt* s, Env& env, Ctx* ctx) {
    if (returned) return;
    switch (s->k) {
      case Stmt::BLOCK: {
        env.emplace_back();
        for (Stmt* c : s->stmts) { exec(c, env, ctx); if (returned) break; }
        env.pop_back();
        return;
      }
      case Stmt::ASSIGN: exec_assign(s, env, ctx); return;
      case Stmt::VARDECL: {
        for (auto& it : s->items) {
          Value v;
          if (it.init) v = eval(it.init, env, ctx, ctx != nullptr && phase == 1, s);
          else {
            std::vector<int> dd;
            for (Expr* d : it.dims) dd.push_back(const_int(eval(d, env, ctx, false, s).s, s, "variable dimension"));
            v = zeros(dd);
          }
          declare_var(env, it.name, v);
        }
        return;
      }
      case Stmt::SIGDECL: {
        if (!ctx) fail(s, "signal declared inside a function");
        for (auto& it : s->items) {
          if (phase == 0) {
            std::vector<int> dd;
            for (Expr* d : it.dims) dd.push_back(const_int(eval(d, env, ctx, false, s).s, s, "signal dimension"));
            if (ctx->lay->sigs.count(it.name)) fail(s, "signal " + nm(it.name) + " declared twice");
            SigInfo si; si.off = 0; si.dims = dd; si.kind = s->op;
            ctx->lay->sigs[it.name] = si;
            ctx->lay->order.push_back(it.name);
          } else if (it.init) {
            Ref r; r.k = Ref::SIG; r.comp = ctx->comp; r.lay = ctx->lay;
            SigInfo& si = ctx->lay->sigs[it.name];
            r.off = si.off; r.dims = si.dims; r.skind = si.kind; r.name = it.name;
            Value v = eval(it.init, env, ctx, it.init_op == O_CONSTRAIN_L, s);
            store_signal(r, v, it.init_op == O_CONSTRAIN_L, s);
          }
        }
        return;
      }
      case Stmt::COMPDECL: {
        if (!ctx) fail(s, "component declared inside a function");
        for (auto& it : s->items) {
          if (phase == 0) {
            std::vector<int> dd;
            for (Expr* d : it.dims) dd.push_back(const_int(eval(d, env, ctx, false, s).s, s, "component dimension"));
            if (!ctx->lay->comp_dims.count(it.name)) ctx->lay->comp_dims[it.name] = dd;
          }
          if (it.init) {
            Ref c; c.k = Ref::COMP; c.name = it.name; c.dims = ctx->lay->comp_dims[it.name];
            instantiate(c, it.init, env, ctx, s);
          }
        }
        return;
      }
      case Stmt::IF: {
        Value c = eval(s->cond, env, ctx, false, s);
        if (c.arr) fail(s, "array used as a condition");
        if (c.s.kind == 2) {
          std::vector<int> names; collect_assigned(s->s1, names); collect_assigned(s->s2, names);
          for (int n : names) { Value* v = find_var(env, n); if (v) *v = scalar(SVal::unk()); }
          return;
        }
        if (c.s.kind == 1) { exec_data_if(s, c.s, env, ctx); return; }
        if (!c.s.c.is_zero()) exec(s->s1, env, ctx);
        else if (s->s2) exec(s->s2, env, ctx);
        return;
      }
      case Stmt::FOR: {
        if (phase == 0 && !relevant_A(s)) return;
        env.emplace_back();
        exec(s->s1, env, ctx);
        uint64_t guard = 0;
        while (!returned) {
          Value c = eval(s->cond, env, ctx, false, s);
          if (c.s.kind != 0) fail(s, "loop condition depends on a signal");
          if (c.s.c.is_zero()) break;
          exec(s->body, env, ctx);
          if (returned) break;
          exec(s->s2, env, ctx);
          if (++guard > 100000000ull) fail(s, "loop does not terminate");
        }
        env.pop_back();
        return;
      }
      case Stmt::WHILE: {
        uint64_t guard = 0;
        while (!returned) {
          Value c = eval(s->cond, env, ctx, false, s);
          if (c.s.kind != 0) fail(s, "while condition depends on a signal");
          if (c.s.c.is_zero()) break;
          exec(s->body, env, ctx);
          if (++guard > 100000000ull) fail(s, "loop does not terminate");
        }
        return;
      }
      case Stmt::INCDEC: {
        Value cur = eval(s->lhs, env, ctx, false, s);
        if (cur.arr) fail(s, "++/-- on an array");
        SVal r = fold_or_emit_bin(s->op > 0 ? O_ADD : O_SUB, cur.s, SVal::konst(U256(1)), s);
        assign_var(s->lhs, scalar(r), env, ctx, s);
        return;
      }
      case Stmt::CONSTR: {
        if (phase == 0) return;
        if (!ctx) fail(s, "constraint inside a function");
        Value a = eval(s->lhs, env, ctx, true, s), b = eval(s->rhs, env, ctx, true, s);
        if (a.arr != b.arr) fail(s, "constraint between an array and a scalar");
        if (!a.arr) { add_constraint(alg_bin(O_SUB, a.s, b.s), s); return; }
        if (a.a->v.size() != b.a->v.size()) fail(s, "array constraint size mismatch");
        for (size_t i = 0; i < a.a->v.size(); i++) add_constraint(alg_bin(O_SUB, a.a->v[i], b.a->v[i]), s);
        return;
      }
      case Stmt::RETURN: {
        if (ctx) fail(s, "return inside a template");
        ret_val = eval(s->rhs, env, ctx, false, s);
        returned = true;
        return;
      }
      case Stmt::ASSERT: {
        Value c = eval(s->cond, env, ctx, false, s);
        if (c.s.kind == 2) return;
        if (c.s.kind == 0) { if (c.s.c.is_zero()) fail(s, "assert failed at compile time"); return; }
        SVal t = truthy(c.s, s);
        emit(PZK_ASSERT_NZ, 0, t.id);
        return;
      }
      case Stmt::LOG: return;
    }
  }

  Value select_value(const SVal& cond, const Value& x, const Value& y, const Stmt* at) {
    if (x.arr != y.arr) fail(at, "if-conversion: an array and a scalar meet");
    if (!x.arr) return scalar(emit_select(cond, x.s, y.s, at));
    if (x.a->v.size() != y.a->v.size()) fail(at, "if-conversion: array size mismatch");
    Value out; out.arr = true; out.a = std::make_shared<AVal>(); out.a->dims = x.a->dims;
    out.a->v.resize(x.a->v.size());
    for (size_t i = 0; i < x.a->v.size(); i++) {
      const SVal& p = x.a->v[i]; const SVal& q = y.a->v[i];
      if (p.kind == q.kind && ((p.kind == 0 && p.c == q.c) || (p.kind == 1 && p.id == q.id))) out.a->v[i] = p;
      else out.a->v[i] = emit_select(cond, p, q, at);
    }
    return out;
  }

  // if (signal-dependent) { ... } else { ... } on variables only: run both arms on copies of
  // the environment and merge every changed variable with a select (if-conversion).  An arm that
  // returns (inside a function) records (condition, value) and contributes nothing to the merge.
  void exec_data_if(Stmt* s, const SVal& cond, Env& env, Ctx* ctx) {
    std::vector<int> names; collect_assigned(s->s1, names); collect_assigned(s->s2, names);
    Env e1 = env, e2 = env;
    bool r0 = returned;
    // returns recorded by nested data-dependent ifs inside an arm only happen on that arm's path:
    // their conditions get the arm's condition ANDed in
    size_t er0 = early_returns.empty() ? 0 : early_returns.back().size();
    exec(s->s1, e1, ctx);
    bool r1 = returned; Value rv1 = ret_val; returned = r0;
    size_t er1 = early_returns.empty() ? 0 : early_returns.back().size();
    if (s->s2) exec(s->s2, e2, ctx);
    bool r2 = returned; Value rv2 = ret_val; returned = r0;
    size_t er2 = early_returns.empty() ? 0 : early_returns.back().size();
    if (r1 || r2 || er2 > er0) {
      if (ctx || early_returns.empty()) fail(s, "return under a signal-dependent condition outside a function");
      SVal c1 = truthy(cond, s);
      SVal c0 = emit_bin(O_EQ, c1, SVal::konst(U256()), s);
      auto& er = early_returns.back();
      for (size_t i = er0; i < er1; i++) er[i].first = emit_bin(O_BAND, er[i].first, c1, s);
      for (size_t i = er1; i < er2; i++) er[i].first = emit_bin(O_BAND, er[i].first, c0, s);
      // earlier entries win at the merge, so an arm's own return goes after its nested ones
      if (r1) er.insert(er.begin() + er1, std::make_pair(c1, rv1));
      if (r2) er.emplace_back(c0, rv2);
    }
    if (r1 && r2) { returned = true; ret_val = rv2; return; }  // both arms left the function
    std::sort(names.begin(), names.end());
    names.erase(std::unique(names.begin(), names.end()), names.end());
    for (int n : names) {
      Value* dst = find_var(env, n);
      if (!dst) continue;  // declared inside the arm
      Value* a = find_var(e1, n); Value* b = find_var(e2, n);
      if (!a || !b) continue;
      if (r1) { *dst = *b; continue; }
      if (r2) { *dst = *a; continue; }
      *dst = select_value(cond, *a, *b, s);
    }
  }

  // ================================================================== driver
  std::vector<std::pair<std::string, std::vector<int>>> main_inputs_desc, main_outputs_desc;
  std::vector<uint32_t> main_input_sig;  // signal index per flattened input, in wire order

  void run_main() {
    if (!unit.main.present) fail("no `component main` in " + main_path);
    Expr* call = unit.main.call;
    if (call->k != Expr::CALL) fail("main must be a template instantiation");
    Env env(1);
    std::vector<Value> args;
    phase = 0;
    for (Expr* a : call->args) args.push_back(eval(a, env, nullptr, false, nullptr));
    main_lay = layout_of(call->name, args, nullptr);
    uint32_t nsig = main_lay->total;
    stats->n_signals = nsig;
    sig_val.assign(nsig, 0);
    // value 0 is reserved
    new_value(CLS_U);
    emit(PZK_NOP, 0);
    // wire order: 1 | outputs | public inputs | private inputs | rest
    sig2wire.assign(nsig, 0);
    uint32_t n_out = 0;
    for (int n : main_lay->order) { SigInfo& s = main_lay->sigs[n]; if (s.kind == 2) n_out += prod(s.dims); }
    std::set<int> pub(unit.main.publics.begin(), unit.main.publics.end());
    for (int p : pub) { auto it = main_lay->sigs.find(p); if (it == main_lay->sigs.end() || it->second.kind != 1) fail("public signal " + nm(p) + " is not a main input"); }
    uint32_t w = 1;
    for (int n : main_lay->order) { SigInfo& s = main_lay->sigs[n]; if (s.kind == 2) { for (uint32_t i = 0; i < prod(s.dims); i++) sig2wire[s.off + i] = w++; main_outputs_desc.emplace_back(nm(n), s.dims); } }
    n_pub_out = n_out;
    for (int pass = 0; pass < 2; pass++)
      for (int n : main_lay->order) {
        SigInfo& s = main_lay->sigs[n];
        if (s.kind != 1) continue;
        bool is_pub = pub.count(n) > 0;
        if ((pass == 0) != is_pub) continue;
        uint32_t cnt = prod(s.dims);
        for (uint32_t i = 0; i < cnt; i++) { sig2wire[s.off + i] = w++; main_input_sig.push_back(s.off + i); }
        main_inputs_desc.emplace_back(nm(n), s.dims);
        if (is_pub) n_pub_in += cnt; else n_prv_in += cnt;
      }
    for (uint32_t sgn = main_lay->n_inputs + n_out; sgn < nsig; sgn++) sig2wire[sgn] = w++;
    // input ops
    phase = 1;
    {
      uint32_t k = 0;
      for (auto& d : main_inputs_desc) {
        int bits = 0;
        auto it = opt.input_bits.find(d.first);
        if (it != opt.input_bits.end()) bits = it->second;
        if (bits < 0 || bits > 64) fail("declared input width must be 1..64 bits (0 = field element)");
        uint32_t cnt = prod(d.second);
        for (uint32_t i = 0; i < cnt; i++, k++) {
          uint32_t id;
          if (bits > 0) {
            i128 hi = bits == 64 ? U64_MAX_ : (((i128)1 << bits) - 1);
            id = new_value(CLS_U, 0, hi);
            emit(PZK_IN_U, id, k, 0, 0, bits);
          } else {
            id = new_value(CLS_F);
            emit(PZK_IN_F, id, k);
          }
          sig_val[main_input_sig[k]] = id;
          PzkInput pi; pi.wire = sig2wire[main_input_sig[k]]; pi.bits = (uint32_t)bits;
          out_inputs.push_back(pi);
        }
      }
    }
    Comp* main = new Comp();
    comps.push_back(main);
    main->lay = main_lay; main->base = 0; main->pending = 0; main->parent = nullptr;
    run_comp(main, nullptr);
    flush_inversions();
    stats->n_constraints = rows.size();
    stats->n_values = v_cls.size();
  }

  // ================================================================== bit-field views
  // (see include/pzk_program.h, "Views")  A value whose defining op only moves bits of one word around -
  // (x >> i) & 1, x & mask, bit * 2^i, the running sums of Num2Bits / Bits2Num / GetLastNBits
  // (/root/reference/circuits/lib/circuits/bitify/bitify.circom:10-55, int/arithmetic.circom:161-204) - is
  // described as a bit field of that word and gets no op of its own.
  std::vector<ViewD> vw;                 // per value
  std::vector<uint8_t> v_tabview;        // per value: exported as a truth-table function of bits of words
  std::vector<uint32_t> alias_of;        // per value: 0 or the value that holds the same number
  struct RangeRow { uint32_t base; uint16_t bits; };
  std::unordered_map<uint32_t, RangeRow> range_rows;  // row -> (base >> bits) == 0
  uint64_t n_range_emitted = 0;
  uint64_t n_view_rows = 0, n_range_rows = 0, n_vlut = 0, n_vlut_lanes = 0, n_view_sigs = 0, n_tabview_sigs = 0, n_extracts = 0;
  static int bitlen128(i128 v) { int n = 0; while (v > 0) { n++; v >>= 1; } return n; }
  uint32_t canon(uint32_t v) const { while (v < alias_of.size() && alias_of[v]) v = alias_of[v]; return v; }
  // bits of v in terms of a word that has (or will have) a slot
  bool src_view(uint32_t v, ViewD& d) const {
    if (v == 0 || v == PZK_OPERAND_NONE || v >= vw.size()) return false;
    if (vw[v].base) { d = vw[v]; return true; }
    int c = v_cls[v];
    if (c == CLS_U && v_lo[v] >= 0) {
      int n = bitlen128(v_hi[v]);
      if (n == 0 || n > 64) return false;
      d.base = v; d.s = 0; d.n = (uint8_t)n; d.k = 0;
      return true;
    }
    if (c == CLS_N && v_def[v] < ops.size() && ops[v_def[v]].dst == v && ops[v_def[v]].opc == PZK_N_FROM_F) {
      d.base = v; d.s = 0; d.n = 254; d.k = 0;
      return true;
    }
    return false;
  }
  static bool pow2_of(const U256& c, int& r) {
    if (c.is_zero()) return false;
    int bl = c.bitlen();
    if (!(c == shl(U256(1), bl - 1))) return false;
    r = bl - 1;
    return true;
  }
  // keep the bits of the value that lie in positions [lo, hi)
  static bool view_window(ViewD& d, int lo, int hi) {
    int t0 = std::max<int>(d.k, lo), t1 = std::min<int>(d.k + d.n, hi);
    if (t1 <= t0) return false;
    d.s = (uint8_t)(d.s + (t0 - d.k)); d.n = (uint8_t)(t1 - t0); d.k = (uint8_t)t0;
    return true;
  }
  static bool view_merge(const ViewD& x, const ViewD& y, int limit, ViewD& out) {
    if (x.base != y.base || (int)x.s - (int)x.k != (int)y.s - (int)y.k) return false;
    const ViewD& lo = x.k <= y.k ? x : y; const ViewD& hi = x.k <= y.k ? y : x;
    if (lo.k + lo.n != hi.k) return false;
    if (lo.k + lo.n + hi.n > limit || lo.n + hi.n > 254) return false;
    out = lo; out.n = (uint8_t)(lo.n + hi.n);
    return true;
  }
  bool const_operand(const OpRec& o, U256& c) const {  // operand b as a constant
    if (o.flags & PZK_FLAG_B_IMM) { c = U256((uint64_t)o.b); return true; }
    if (o.flags & PZK_FLAG_B_POOL) return false;
    if (o.b < v_const.size() && v_const[o.b] && v_cls[o.b] == CLS_U) { auto it = const_of.find(o.b); if (it != const_of.end()) { c = it->second; return true; } }
    return false;
  }
  // the view the result of op `o` is, given the views of its operands
  bool try_view(const OpRec& o, ViewD& d) const {
    U256 c; int r;
    switch (o.opc) {
      case PZK_U_AND: {
        if (!const_operand(o, c) || !c.fits64() || c.w[0] == 0) return false;
        uint64_t m = c.w[0];
        int lo = __builtin_ctzll(m), pc = __builtin_popcountll(m);
        uint64_t run = pc == 64 ? ~0ull : (((1ull << pc) - 1) << lo);
        if (m != run || !src_view(o.a, d)) return false;
        return view_window(d, lo, lo + pc);
      }
      case PZK_U_SHR: case PZK_N_SHR: {
        if (!(o.flags & PZK_FLAG_B_IMM) || !src_view(o.a, d)) return false;
        int sh = (int)o.b;
        if (sh >= 254 || !view_window(d, sh, 256)) return false;
        d.k = (uint8_t)(d.k - sh);
        return true;
      }
      case PZK_U_SHL: case PZK_U_MUL: {
        if (o.opc == PZK_U_MUL && !(o.flags & (PZK_FLAG_B_IMM | PZK_FLAG_B_POOL)) && o.a == o.b) {
          if (!src_view(o.a, d)) return false;
          return d.n == 1 && d.k == 0;  // bit * bit
        }
        if (!const_operand(o, c) || !c.fits64()) return false;
        if (o.opc == PZK_U_SHL) { if (!(o.flags & PZK_FLAG_B_IMM)) return false; r = (int)c.w[0]; }
        else if (!pow2_of(c, r)) return false;
        if (v_cls[o.dst] != CLS_U || !src_view(o.a, d)) return false;
        if (d.k + d.n + r > 64) return false;
        d.k = (uint8_t)(d.k + r);
        return true;
      }
      case PZK_U_ADD: case PZK_F_ADD: case PZK_Z_ADD: {
        if (o.flags & (PZK_FLAG_B_IMM | PZK_FLAG_B_POOL)) return false;
        ViewD x, y;
        if (!src_view(o.a, x) || !src_view(o.b, y)) return false;
        return view_merge(x, y, o.opc == PZK_U_ADD ? 64 : (o.opc == PZK_Z_ADD ? 250 : 253), d);
      }
      case PZK_Z_MUL: {
        if (!(o.flags & PZK_FLAG_B_POOL)) {
          if (o.a == o.b && src_view(o.a, d)) return d.n == 1 && d.k == 0;
          return false;
        }
        c = fpool[o.b];   // raw pattern: a power of two is non-negative
        if ((c.w[3] >> 63) || !pow2_of(c, r) || !src_view(o.a, d)) return false;
        if (d.k + d.n + r > 250) return false;
        d.k = (uint8_t)(d.k + r);
        return true;
      }
      case PZK_F_MUL: {
        if (!(o.flags & PZK_FLAG_B_POOL)) {
          if (o.a == o.b && src_view(o.a, d)) return d.n == 1 && d.k == 0;
          return false;
        }
        c = fr_from_mont(fpool[o.b]);
        if (!pow2_of(c, r) || !src_view(o.a, d)) return false;
        if (d.k + d.n + r > 253) return false;
        d.k = (uint8_t)(d.k + r);
        return true;
      }
      case PZK_N_BIT: {
        if (!src_view(o.a, d)) return false;
        int b = (int)o.b;
        if (!view_window(d, b, b + 1)) return false;
        d.k = 0;
        return true;
      }
      case PZK_N_LOW:
        if (!src_view(o.a, d)) return false;
        return view_window(d, 0, 64);
      case PZK_N_FROM_U: case PZK_F_FROM_U: case PZK_Z_FROM_U: return src_view(o.a, d);
      case PZK_F_FROM_N: case PZK_N_FROM_F:
        if (!vw[o.a].base) return false;  // only views pass through; a real F value is the base of its own bits
        d = vw[o.a];
        return d.k + d.n <= 253;
    }
    return false;
  }
  // ---- packed truth tables ------------------------------------------------------------------------
  // One-bit LUT ops whose operands are bits (W_j, q_j + l mod w) of the same words for lanes l = 0, 1, ... -
  // the 32 XOR3 / Ch / Maj instances of a SHA round, rotations included
  // (/root/reference/circuits/lib/circuits/hasher/sha2/sha256/sha256Compress.circom:63-84) - become one
  // V_LUT record computing a whole word; each member signal is then bit l of that word (a view).
  struct Place { uint32_t word; uint8_t pos; };
  struct Group {
    uint32_t word;            // value id of the result
    int n, w;                 // operands, lane width (32 / 64)
    uint16_t tbl;
    uint32_t W[4]; uint8_t q[4];  // operand words and their bit positions for lane 0
    uint64_t lanes = 0;
    uint64_t last_join = 0;
    bool open = true;
  };
  std::vector<Group> groups;
  std::unordered_map<uint32_t, uint32_t> group_of_word;      // word value -> group index
  std::unordered_map<uint32_t, Place> late_place;            // scalar bit -> (later word, position)
  std::unordered_map<uint32_t, std::vector<std::pair<uint32_t, uint8_t>>> word_comp;  // word built from scalar bits
  std::unordered_map<uint64_t, std::vector<uint32_t>> open_groups;  // hash of operand words -> group indices

  static uint16_t permute_table(uint16_t t, int n, const int* perm) {  // operand j of the new table = operand perm[j] of t
    uint16_t out = 0;
    for (int idx = 0; idx < 16; idx++) {
      int src = 0;
      for (int j = 0; j < n; j++) if ((idx >> j) & 1) src |= 1 << perm[j];
      if ((t >> src) & 1) out |= (uint16_t)(1u << idx);
    }
    return out;
  }
  static uint16_t canon_table(uint16_t t, int n) {  // replicate over the unused high index bits
    uint16_t out = 0;
    for (int idx = 0; idx < 16; idx++) if ((t >> (idx & ((1 << n) - 1))) & 1) out |= (uint16_t)(1u << idx);
    return out;
  }

  void vectorize_sweep(std::vector<uint8_t>& keep, const std::vector<uint8_t>& row_needed,
                       const std::function<void(const OpRec&, const std::function<void(uint32_t)>&)>& for_operands) {
    const size_t nops = ops.size();
    std::vector<OpRec> out;
    out.reserve(nops / 2);
    std::vector<int> out_tag;
    vw.assign(v_cls.size(), ViewD());
    v_tabview.assign(v_cls.size(), 0);
    alias_of.assign(v_cls.size(), 0);
    std::vector<uint8_t> is_real(v_cls.size(), 0);  // the value has (or will have) an op of its own in `out`
    auto grow = [&]() {
      size_t n = v_cls.size();
      if (vw.size() < n) { vw.resize(n); v_tabview.resize(n, 0); alias_of.resize(n, 0); is_real.resize(n, 0); }
    };
    int cur_tag = -1;
    auto push = [&](const OpRec& o) { out.push_back(o); if (op_stats) out_tag.push_back(cur_tag); };
    std::map<std::tuple<uint32_t, int, int, int, int>, uint32_t> mat_cse;  // (base, s, n, k, class) -> value
    std::function<uint32_t(uint32_t)> real_of;
    std::function<void(uint32_t)> close_group;
    // ---- backward hint: which table-valued values have a consumer that needs them as numbers
    std::vector<uint8_t> need_real(v_cls.size(), 0);
    for (size_t v = 0; v < row_needed.size(); v++) if (row_needed[v]) need_real[v] = 1;
    auto viewable_opc = [&](const OpRec& o) {
      switch (o.opc) {
        case PZK_U_AND: case PZK_U_SHR: case PZK_U_SHL: case PZK_N_SHR: return (o.flags & PZK_FLAG_B_IMM) != 0;
        case PZK_U_MUL: { U256 c; int r; return const_operand(o, c) && pow2_of(c, r); }
        case PZK_F_MUL: { int r; return (o.flags & PZK_FLAG_B_POOL) && pow2_of(fr_from_mont(fpool[o.b]), r); }
        case PZK_Z_MUL: { int r; return (o.flags & PZK_FLAG_B_POOL) && !(fpool[o.b].w[3] >> 63) && pow2_of(fpool[o.b], r); }
        case PZK_N_BIT: case PZK_N_LOW: case PZK_N_FROM_U: case PZK_F_FROM_U: case PZK_Z_FROM_U: return true;
      }
      return false;
    };
    auto tabled = [&](const OpRec& o) {
      if (o.dst == 0 || o.dst >= v_tbl.size() || v_tbl[o.dst] < 0) return false;
      if (v_cls[o.dst] != CLS_U && v_cls[o.dst] != CLS_I) return false;
      switch (o.opc) {
        case PZK_U_ADD: case PZK_U_SUB: case PZK_U_MUL: case PZK_U_LUT: case PZK_U_LUTV: case PZK_U_AND: case PZK_U_OR:
        case PZK_U_XOR: case PZK_U_EQ: case PZK_U_NE: case PZK_U_LT: case PZK_U_LE: case PZK_U_SHR: case PZK_U_SHL: return tables[v_tbl[o.dst]].n >= 1;
      }
      return false;
    };
    for (size_t i = nops; i-- > 0;) {
      if (!keep[i]) continue;
      const OpRec& o = ops[i];
      if (viewable_opc(o)) continue;
      if (tabled(o) && !need_real[o.dst]) continue;
      for_operands(o, [&](uint32_t v) { if (v != PZK_OPERAND_NONE && v < need_real.size()) need_real[v] = 1; });
    }
    // ---- placement of a one-bit value inside a word
    auto place_of = [&](uint32_t b, Place& p) -> bool {
      if (vw[b].base) {
        const ViewD& d = vw[b];
        if (d.n != 1 || d.k != 0 || d.s >= 64) return false;
        int bc = v_cls[d.base];
        if (bc != CLS_U) return false;
        p.word = d.base; p.pos = d.s;
        return true;
      }
      auto it = late_place.find(b);
      if (it != late_place.end()) { p = it->second; return true; }
      return false;
    };
    // ---- materialisation: the op that gives a view / table view a slot of its own
    auto word_ready = [&](uint32_t w) {
      auto g = group_of_word.find(w);
      if (g != group_of_word.end() && groups[g->second].open) close_group(g->second);
    };
    auto emit_lut_positions = [&](uint32_t dst, const Table& t) {
      OpRec o; memset(&o, 0, sizeof o);
      o.dst = dst; o.flags = PZK_FLAG_EXT;
      uint32_t s[4] = {PZK_OPERAND_NONE, PZK_OPERAND_NONE, PZK_OPERAND_NONE, PZK_OPERAND_NONE};
      uint32_t posw = 0;
      for (int j = 0; j < t.n; j++) {
        uint32_t root = t.sup[j];
        Place p;
        if (vw[root].base && place_of(root, p)) { word_ready(p.word); s[j] = p.word; posw |= (uint32_t)p.pos << (8 * j); }
        else s[j] = real_of(root);
      }
      o.a = s[0]; o.b = s[1]; o.c = s[2]; o.d = s[3]; o.f = posw;
      if (table_bits(t)) {
        int imm = 0;
        for (int k = 0; k < 16; k++) if (t.e[k & ((1 << t.n) - 1)] == 1) imm |= 1 << k;
        o.opc = PZK_U_LUT; o.imm16 = (uint16_t)imm;
      } else {
        o.opc = PZK_U_LUTV; o.e = (uint32_t)list_pool.size();
        for (int k = 0; k < 16; k++) { uint64_t v = (uint64_t)t.e[k & ((1 << t.n) - 1)]; list_pool.push_back((uint32_t)v); list_pool.push_back((uint32_t)(v >> 32)); }
      }
      push(o);
    };
    real_of = [&](uint32_t v) -> uint32_t {
      if (v == PZK_OPERAND_NONE || v == 0) return v;
      v = canon(v);
      if (is_real[v]) return v;
      if (group_of_word.count(v)) { word_ready(v); return v; }
      if (v_tabview[v]) {
        v_tabview[v] = 0; is_real[v] = 1;
        emit_lut_positions(v, tables[v_tbl[v]]);
        n_extracts++;
        return v;
      }
      if (!vw[v].base) { is_real[v] = 1; return v; }  // values defined outside the sweep's knowledge (should not happen)
      ViewD d = vw[v];
      uint32_t b = real_of(d.base);
      const int cls = v_cls[v], bcls = v_cls[b];
      const bool nbase = (bcls == CLS_N || bcls == CLS_F);
      int bw = nbase ? 254 : bitlen128(v_hi[b]);
      if (!nbase && d.s == 0 && d.k == 0 && d.n >= bw && cls == CLS_U) { alias_of[v] = b; return b; }
      auto key = std::make_tuple(b, (int)d.s, (int)d.n, (int)d.k, cls);
      auto it = mat_cse.find(key);
      if (it != mat_cse.end()) { alias_of[v] = it->second; return it->second; }
      mat_cse[key] = v;
      OpRec o; memset(&o, 0, sizeof o);
      o.dst = v; o.a = b;
      auto extract = [&](int opc, uint32_t dst) {
        OpRec x; memset(&x, 0, sizeof x);
        x.opc = (uint8_t)opc; x.dst = dst; x.a = b; x.flags = nbase ? PZK_FLAG_NBASE : 0;
        x.imm16 = (uint16_t)(d.s | (d.k << 8)); x.b = d.n;
        push(x);
      };
      n_extracts++;
      if (cls == CLS_U) {
        if (!nbase && d.k == 0 && d.s == 0 && d.n <= 32) { o.opc = PZK_U_AND; o.flags = PZK_FLAG_B_IMM; o.b = d.n == 32 ? 0xffffffffu : ((1u << d.n) - 1); push(o); }
        else if (!nbase && d.k == 0 && d.s + d.n >= bw) { o.opc = PZK_U_SHR; o.flags = PZK_FLAG_B_IMM; o.b = d.s; push(o); }
        else if (!nbase && d.s == 0 && d.n >= bw && d.k < 32) { o.opc = PZK_U_MUL; o.flags = PZK_FLAG_B_IMM; o.b = 1u << d.k; push(o); }
        else extract(PZK_U_EXTRACT, v);
      } else if (cls == CLS_N) {
        if (!nbase && d.s == 0 && d.k == 0 && d.n >= bw) { o.opc = PZK_N_FROM_U; push(o); }
        else extract(PZK_N_EXTRACT, v);
      } else if (cls == CLS_Z) {  // a non-negative plain integer is its own two's complement pattern
        if (!nbase && d.s == 0 && d.k == 0 && d.n >= bw) { o.opc = PZK_Z_FROM_U; push(o); }
        else extract(PZK_N_EXTRACT, v);
      } else {  // CLS_F: Montgomery form of the integer
        if (!nbase && d.s == 0 && d.k == 0 && d.n >= bw) { o.opc = PZK_F_FROM_U; push(o); }
        else if (d.k + d.n <= 64) {
          uint32_t t = new_value(CLS_U, 0, d.k + d.n == 64 ? U64_MAX_ : (((i128)1 << (d.k + d.n)) - 1)); grow();
          extract(PZK_U_EXTRACT, t); is_real[t] = 1;
          o.opc = PZK_F_FROM_U; o.a = t; push(o);
        } else {
          uint32_t t = new_value(CLS_N); grow();
          extract(PZK_N_EXTRACT, t); is_real[t] = 1;
          o.opc = PZK_F_FROM_N; o.a = t; push(o);
        }
      }
      is_real[v] = 1;
      return v;
    };
    close_group = [&](uint32_t gi) {
      Group& g0 = groups[gi];
      if (!g0.open) return;
      g0.open = false;
      uint32_t W[4]; for (int j = 0; j < g0.n; j++) W[j] = real_of(g0.W[j]);
      Group& g = groups[gi];  // real_of may grow `groups`
      OpRec o; memset(&o, 0, sizeof o);
      o.opc = PZK_V_LUT; o.flags = PZK_FLAG_EXT | (g.w == 64 ? PZK_FLAG_W64 : 0); o.imm16 = g.tbl; o.dst = g.word;
      uint32_t s[4] = {PZK_OPERAND_NONE, PZK_OPERAND_NONE, PZK_OPERAND_NONE, PZK_OPERAND_NONE};
      uint32_t rot = 0;
      for (int j = 0; j < g.n; j++) { s[j] = W[j]; rot |= (uint32_t)g.q[j] << (8 * j); }
      o.a = s[0]; o.b = s[1]; o.c = s[2]; o.d = s[3]; o.e = rot; o.f = (uint32_t)g.lanes; o.g = (uint32_t)(g.lanes >> 32);
      // exact interval of the word: its highest lane
      v_lo[g.word] = 0; v_hi[g.word] = (i128)g.lanes;
      push(o);
      is_real[g.word] = 1;
      n_vlut++; n_vlut_lanes += (uint64_t)__builtin_popcountll(g.lanes);
      if (getenv("PZK_DUMP_GROUPS")) fprintf(stderr, "G n=%d w=%d tbl=%04x lanes=%016llx W=%u,%u,%u q=%d,%d,%d tag=%s\n", g.n, g.w, g.tbl, (unsigned long long)g.lanes, g.W[0], g.n > 1 ? g.W[1] : 0, g.n > 2 ? g.W[2] : 0, g.q[0], g.n > 1 ? g.q[1] : -1, g.n > 2 ? g.q[2] : -1, cur_tag >= 0 ? nm(cur_tag).c_str() : "?");
    };
    auto words_key = [&](const uint32_t* W, int n, int w) {
      uint32_t s[4]; for (int j = 0; j < n; j++) s[j] = W[j];
      std::sort(s, s + n);
      uint64_t h = (uint64_t)w * 1315423911u + (uint64_t)n;
      for (int j = 0; j < n; j++) h = h * 0x9E3779B97F4A7C15ull + s[j];
      return h;
    };
    auto try_group = [&](const OpRec& o, const Table& tb, uint64_t at) -> bool {
      if (!opt.vectorize || !table_bits(tb) || tb.n < 1) return false;
      Place pl[4]; int n = tb.n;
      for (int j = 0; j < n; j++) if (!place_of(canon(tb.sup[j]), pl[j])) return false;
      uint16_t timm = 0;
      for (int k = 0; k < 16; k++) if (tb.e[k & ((1 << n) - 1)] == 1) timm |= (uint16_t)(1u << k);
      int w = 32;
      for (int j = 0; j < n; j++) if (pl[j].pos >= 32) w = 64;
      uint32_t Ws[4]; for (int j = 0; j < n; j++) Ws[j] = pl[j].word;
      const uint16_t T = timm;
      uint64_t key = words_key(Ws, n, w);
      auto& cand = open_groups[key];
      // try to join: some assignment of our operands to the group's operands with one common lane
      for (size_t ci = cand.size(); ci-- > 0;) {
        Group& g = groups[cand[ci]];
        if (!g.open) { cand.erase(cand.begin() + ci); continue; }
        if (g.n != n || g.w != w) continue;
        int perm[4] = {0, 1, 2, 3};
        std::sort(perm, perm + n);
        do {
          // our operand perm[j] plays the role of the group's operand j
          bool ok = true; int lane = -1;
          for (int j = 0; j < n && ok; j++) {
            const Place& p = pl[perm[j]];
            if (p.word != g.W[j]) { ok = false; break; }
            int l = ((int)p.pos - (int)g.q[j]) & (w - 1);
            if (lane < 0) lane = l; else if (l != lane) ok = false;
          }
          if (!ok || ((g.lanes >> lane) & 1)) continue;
          // group table indexed by group operand order: bit j of the index = our operand perm[j]
          // our table T is indexed by our operand order; T' (idx) = T(idx') with idx' bit perm[j] = idx bit j
          uint16_t Tp = 0;
          for (int idx = 0; idx < 16; idx++) {
            int src = 0;
            for (int j = 0; j < n; j++) if ((idx >> j) & 1) src |= 1 << perm[j];
            if ((T >> src) & 1) Tp |= (uint16_t)(1u << idx);
          }
          Tp = canon_table(Tp, n);
          if (Tp != g.tbl) continue;
          g.lanes |= 1ull << lane; g.last_join = at;
          vw[o.dst].base = g.word; vw[o.dst].s = (uint8_t)lane; vw[o.dst].n = 1; vw[o.dst].k = 0;
          if (__builtin_popcountll(g.lanes) == w) close_group(cand[ci]);
          return true;
        } while (std::next_permutation(perm, perm + n));
      }
      // open a new group with this op as lane 0
      Group g; g.n = n; g.w = w; g.tbl = T; g.lanes = 1; g.last_join = at;
      for (int j = 0; j < n; j++) { g.W[j] = pl[j].word; g.q[j] = pl[j].pos; }
      g.word = new_value(CLS_U, 0, w == 64 ? U64_MAX_ : (i128)0xffffffffu); grow();
      groups.push_back(g);
      uint32_t gi = (uint32_t)groups.size() - 1;
      group_of_word[g.word] = gi;
      cand.push_back(gi);
      if (cand.size() > 8) { close_group(cand.front()); cand.erase(cand.begin()); }
      vw[o.dst].base = g.word; vw[o.dst].s = 0; vw[o.dst].n = 1; vw[o.dst].k = 0;
      return true;
    };
    // ---- the sweep
    std::vector<uint32_t> stale_check;
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      OpRec o = ops[i];
      if (op_stats && i < op_tag.size()) cur_tag = op_tag[i];
      if (o.opc == PZK_U_LUTV) { auto it = lutv_off.find((uint32_t)i); if (it != lutv_off.end()) o.e = it->second; }
      const uint32_t d = o.dst;
      bool has_value = !(o.opc == PZK_NOP || o.opc == PZK_ASSERT_NZ || is_macro(o.opc));
      bool dropped = false;
      if (opt.views && has_value && d) {
        ViewD dv;
        // canonical operands first: views built on aliases keep pointing at real words
        if (try_view(o, dv)) {
          uint32_t b = canon(dv.base);
          if (v_tabview[b]) b = real_of(b);
          if (vw[b].base) {  // cannot happen: bases are words
            throw CompileError("internal: view of a view");
          }
          dv.base = b;
          vw[d] = dv; dropped = true;
        } else if (tabled(o)) {
          const Table& tb = tables[v_tbl[d]];
          if (try_group(o, tb, i)) dropped = true;
          else if (!need_real[d]) {
            // exported as a table over bits of words: every root needs an address (a word and a position)
            for (int j = 0; j < tb.n; j++) { uint32_t root = canon(tb.sup[j]); Place p; if (!(vw[root].base && place_of(root, p))) real_of(root); }
            v_tabview[d] = 1; dropped = true;
          }
        }
      }
      if (dropped) {
        if (row_needed.size() > d && row_needed[d]) real_of(d);
      } else {
        // a real op: its operands need slots
        switch (o.opc) {
          case PZK_NOP: case PZK_U_CONST: case PZK_F_CONST: case PZK_IN_U: case PZK_IN_F: break;
          case PZK_BIGDIV: case PZK_MODINV: case PZK_BJJ_MUL8: {
            uint32_t op0, nop, def0, ndef; macro_layout(o, op0, nop, def0, ndef);
            for (uint32_t j = 0; j < nop; j++) list_pool[op0 + j] = real_of(list_pool[op0 + j]);
            break;
          }
          case PZK_ASSERT_NZ: case PZK_N_BIT: case PZK_F_CSEL: o.a = real_of(o.a); break;
          case PZK_U_LUT: case PZK_U_LUTV: {
            // operands that are bits of words are read in place (bit positions in the extension record)
            uint32_t* s[4] = {&o.a, &o.b, &o.c, &o.d};
            uint32_t posw = 0;
            for (int j = 0; j < 4; j++) {
              if (*s[j] == PZK_OPERAND_NONE) continue;
              uint32_t v = canon(*s[j]);
              Place p;
              if (!is_real[v] && vw[v].base && place_of(v, p)) { word_ready(p.word); *s[j] = p.word; posw |= (uint32_t)p.pos << (8 * j); }
              else *s[j] = real_of(v);
            }
            o.f = posw;
            break;
          }
          case PZK_U_SEL: case PZK_F_SEL: o.a = real_of(o.a); o.b = real_of(o.b); o.c = real_of(o.c); break;
          default:
            o.a = real_of(o.a);
            if (!(o.flags & (PZK_FLAG_B_IMM | PZK_FLAG_B_POOL))) {
              switch (o.opc) {
                case PZK_F_NEG: case PZK_F_INV: case PZK_F_FROM_U: case PZK_F_FROM_I: case PZK_N_FROM_F: case PZK_Z_FROM_U: case PZK_Z_FROM_I:
                case PZK_F_FROM_N: case PZK_N_FROM_U: case PZK_N_LOW: case PZK_N_FITS: break;
                default: o.b = real_of(o.b);
              }
            }
        }
        push(o);
        if (has_value && d) is_real[d] = 1;
        if (is_macro(o.opc)) {
          uint32_t op0, nop, def0, ndef; macro_layout(o, op0, nop, def0, ndef);
          for (uint32_t j = 0; j < ndef; j++) is_real[list_pool[def0 + j]] = 1;
        }
        // a word assembled from scalar bits (b << k summed in ascending order): the bits get a place in it
        if (opt.vectorize && o.opc == PZK_U_ADD && !(o.flags & PZK_FLAG_B_IMM) && v_cls[d] == CLS_U) {
          const OpRec& src = ops[i];  // operands before materialisation
          std::vector<std::pair<uint32_t, uint8_t>> comp;
          uint64_t occ = 0; bool ok = true;
          for (uint32_t opnd : {src.a, src.b}) {
            uint32_t v = canon(opnd);
            auto wc = word_comp.find(v);
            if (wc != word_comp.end()) { for (auto& e : wc->second) { if ((occ >> e.second) & 1) ok = false; occ |= 1ull << e.second; comp.push_back(e); } continue; }
            // a scalar bit shifted left by k
            uint32_t bit = 0; int k = -1;
            const OpRec* def = (v < v_def.size() && v_def[v] < nops && ops[v_def[v]].dst == v) ? &ops[v_def[v]] : nullptr;
            U256 c; int r;
            if (v_hi[v] <= 1 && v_lo[v] >= 0 && v_cls[v] == CLS_U) { bit = v; k = 0; }
            else if (def && def->opc == PZK_U_MUL && const_operand(*def, c) && pow2_of(c, r) && r < 64) {
              uint32_t x = canon(def->a);
              if (v_cls[x] == CLS_U && v_lo[x] >= 0 && v_hi[x] <= 1) { bit = x; k = r; }
            }
            if (k < 0) { ok = false; break; }
            if ((occ >> k) & 1) ok = false;
            occ |= 1ull << k; comp.emplace_back(bit, (uint8_t)k);
          }
          if (ok && comp.size() >= 2) {
            for (auto& e : comp) { Place p; p.word = d; p.pos = e.second; if (!vw[e.first].base) late_place[e.first] = p; }
            word_comp[d] = std::move(comp);
          }
        }
      }
      // groups that stopped growing are closed so that their operands can die
      if ((i & 1023) == 0 && !groups.empty()) {
        for (auto& kv : open_groups) {
          auto& cand = kv.second;
          for (size_t ci = cand.size(); ci-- > 0;) {
            Group& g = groups[cand[ci]];
            if (!g.open) { cand.erase(cand.begin() + ci); continue; }
            if (i - g.last_join > 4096) { uint32_t gi = cand[ci]; cand.erase(cand.begin() + ci); close_group(gi); }
          }
        }
      }
    }
    for (size_t gi = 0; gi < groups.size(); gi++) if (groups[gi].open) close_group((uint32_t)gi);
    ops.swap(out);
    if (op_stats) op_tag.swap(out_tag);
    keep.assign(ops.size(), 1);
    lutv_off.clear();
  }

  // ================================================================== back end
  void backend();
  void build_meta();
};

// ---------------------------------------------------------------------------------------
// Back end: dead-code elimination, segmentation, liveness, slot allocation, row lowering
// ---------------------------------------------------------------------------------------
void Compiler::Impl::backend() {
  size_t nv = v_cls.size(), nops = ops.size();
  // ---- operand enumeration helper
  std::function<void(const OpRec&, const std::function<void(uint32_t)>&)> for_operands = [&](const OpRec& o, const std::function<void(uint32_t)>& f) {
    switch (o.opc) {
      case PZK_NOP: case PZK_U_CONST: case PZK_F_CONST: case PZK_Z_CONST: case PZK_IN_U: case PZK_IN_F: return;
      case PZK_BIGDIV: case PZK_MODINV: case PZK_BJJ_MUL8: {
        uint32_t op0, nop, def0, ndef; macro_layout(o, op0, nop, def0, ndef);
        for (uint32_t i = 0; i < nop; i++) f(list_pool[op0 + i]);
        return;
      }
      case PZK_ASSERT_NZ: case PZK_U_EXTRACT: case PZK_N_EXTRACT: f(o.a); return;
      case PZK_U_LUT: case PZK_U_LUTV: case PZK_V_LUT:
        if (o.a != PZK_OPERAND_NONE) f(o.a);
        if (o.b != PZK_OPERAND_NONE) f(o.b);
        if (o.c != PZK_OPERAND_NONE) f(o.c);
        if (o.d != PZK_OPERAND_NONE) f(o.d);
        return;
      case PZK_U_SEL: case PZK_F_SEL: f(o.a); f(o.b); f(o.c); return;
      case PZK_N_BIT: f(o.a); return;
      case PZK_F_CSEL: f(o.a); return;
      case PZK_F_MULADD: case PZK_Z_MULADD: f(o.a); if (!(o.flags & PZK_FLAG_B_POOL)) f(o.b); f(o.c); return;
      case PZK_Z_MUL: f(o.a); if (!(o.flags & PZK_FLAG_B_POOL)) f(o.b); return;  // bit 0 is PZK_FLAG_B_U here, not B_IMM
      default:
        f(o.a);
        if (!(o.flags & (PZK_FLAG_B_IMM | PZK_FLAG_B_POOL))) {
          switch (o.opc) {  // unary ops have no b
            case PZK_F_NEG: case PZK_F_INV: case PZK_F_FROM_U: case PZK_F_FROM_I: case PZK_N_FROM_F: case PZK_Z_FROM_U: case PZK_Z_FROM_I:
            case PZK_F_FROM_N: case PZK_N_FROM_U: case PZK_N_LOW: case PZK_N_FITS: break;
            default: f(o.b);
          }
        }
    }
  };
  auto for_defs = [&](const OpRec& o, const std::function<void(uint32_t)>& f) {
    if (o.opc == PZK_NOP || o.opc == PZK_ASSERT_NZ) return;
    if (is_macro(o.opc)) {
      uint32_t op0, nop, def0, ndef; macro_layout(o, op0, nop, def0, ndef);
      for (uint32_t i = 0; i < ndef; i++) f(list_pool[def0 + i]);
      return;
    }
    f(o.dst);
    if ((o.opc == PZK_F_MULADD || o.opc == PZK_Z_MULADD) && (o.flags & PZK_FLAG_DST2)) f(o.d);  // the product, a wire
  };
  // ---- sanity: every operand is defined by an earlier op (catches scheduling bugs of deferred ops)
  {
    std::vector<uint32_t> def_at(nv, 0xFFFFFFFFu);
    for (size_t i = 0; i < nops; i++) for_defs(ops[i], [&](uint32_t d) { if (d < nv) def_at[d] = (uint32_t)i; });
    for (size_t i = 0; i < nops; i++)
      for_operands(ops[i], [&](uint32_t v) {
        if (v == PZK_OPERAND_NONE) return;
        if (v >= nv || def_at[v] == 0xFFFFFFFFu || def_at[v] >= i)
          throw CompileError(fmt("internal: op %zu (opc %d) uses value %u defined at op %u (opc %d)", i, (int)ops[i].opc, v,
                                 v < nv ? def_at[v] : 0u, (v < nv && def_at[v] < nops) ? (int)ops[def_at[v]].opc : -1));
      });
  }
  // ---- DCE
  std::vector<uint8_t> used(nv, 0), keep(nops, 0);
  for (uint32_t v : sig_val) if (v) used[v] = 1;
  for (size_t i = nops; i-- > 0;) {
    const OpRec& o = ops[i];
    bool live = (o.opc == PZK_ASSERT_NZ) || (o.opc == PZK_IN_U) || (o.opc == PZK_IN_F);
    if (!live) for_defs(o, [&](uint32_t d) { if (used[d]) live = true; });
    if (!live) continue;
    keep[i] = 1;
    for_operands(o, [&](uint32_t v) { used[v] = 1; });
  }
  auto print_op_stats = [&](const char* tag) {
    if (!op_stats || op_tag.size() != ops.size()) return;
    std::map<std::pair<std::string, int>, uint64_t> hist;
    std::map<std::string, uint64_t> per_t;
    for (size_t i = 0; i < ops.size(); i++) if (keep[i]) { std::string t = op_tag[i] >= 0 ? nm(op_tag[i]) : "<main>"; hist[{t, ops[i].opc}]++; per_t[t]++; }
    for (auto& kv : per_t) {
      fprintf(stderr, "%s %-40s %9llu :", tag, kv.first.c_str(), (unsigned long long)kv.second);
      for (auto& h : hist) if (h.first.first == kv.first) fprintf(stderr, " %d:%llu", h.first.second, (unsigned long long)h.second);
      fprintf(stderr, "\n");
    }
  };
  print_op_stats("T");
  size_t nrows = rows.size();
  struct MT { uint32_t val; U256 c; };  // val: value id, 0xFFFFFFFF = constant one
  // merge the terms of each linear combination by SSA value: wires that alias the same value
  // (every `a <== b`) collapse, and a row whose combinations cancel completely is satisfied by
  // construction - it is proven at compile time and needs no run-time work.
  auto merge_row = [&](uint32_t r, std::vector<MT>* parts) {
    uint32_t lens[3] = {rows[r].na, rows[r].nb, rows[r].nc};
    uint64_t off = rows[r].off;
    for (int part = 0; part < 3; part++) {
      parts[part].clear();
      for (uint32_t t = 0; t < lens[part]; t++, off++) {
        uint32_t sig = terms[off].first;
        uint32_t val = 0xFFFFFFFFu;
        if (sig != 0xFFFFFFFFu) { val = sig_val[sig]; if (!val) continue; val = canon(val); }
        const U256& c = coefs[terms[off].second];
        bool found = false;
        for (auto& m : parts[part]) if (m.val == val) { m.c = fr_add(m.c, c); found = true; break; }
        if (!found) parts[part].push_back({val, c});
      }
      size_t w = 0;
      for (size_t k = 0; k < parts[part].size(); k++) if (!parts[part][k].c.is_zero()) parts[part][w++] = parts[part][k];
      parts[part].resize(w);
    }
    return parts[2].empty() && (parts[0].empty() || parts[1].empty());
  };
  std::vector<uint8_t>& row_static = row_kind;
  row_static.assign(nrows, 0);
  {
    std::vector<MT> parts[3];
    // A row whose wires are all functions of a few proven bits (truth tables over <= 8 root bits: the
    // XOR / Maj / Ch / carry rows of SHA, bit checks b * (b - 1) = 0, selectors) is an identity over
    // those bits: it is evaluated here for every assignment of the roots, with the very tables the device
    // uses to compute the wires, and discharged when it holds for all of them.
    auto const_of = [&](uint32_t id, U256& out) -> bool {
      uint32_t di = v_def[id];
      if (di >= nops) return false;
      const OpRec& o = ops[di];
      if (o.dst != id) return false;
      if (o.opc == PZK_U_CONST) {
        uint64_t raw = ((uint64_t)o.b << 32) | o.a;
        if (v_cls[id] == CLS_I && (int64_t)raw < 0) out = fr_neg(U256((uint64_t)(-(int64_t)raw)));
        else out = U256(raw);
        return true;
      }
      if (o.opc == PZK_F_CONST) { out = fr_from_mont(fpool[o.a]); return true; }
      if (o.opc == PZK_Z_CONST) { out = z_field(fpool[o.a]); return true; }
      return false;
    };
    auto prove_by_tables = [&](const std::vector<MT>* parts) -> bool {
      if (!opt.table_rows_static) return false;
      uint32_t roots[8]; int nr = 0;
      struct TV { const Table* t; Table self; i128 coef; int pos[4]; };
      std::vector<TV> tv[3];
      for (int part = 0; part < 3; part++) {
        tv[part].clear();
        for (const MT& m : parts[part]) {
          TV x; x.t = nullptr;
          i128 cv;
          if (!const_narrow(m.c, cv)) return false;
          x.coef = cv;
          if (m.val != 0xFFFFFFFFu) {
            uint32_t id = m.val;
            U256 kc; int64_t ks;
            if (const_of(id, kc)) { if (!small_signed(kc, ks)) return false; x.self.n = 0; x.self.e[0] = ks; }
            else if (v_tbl[id] >= 0) x.t = &tables[v_tbl[id]];
            else if (v_cls[id] == CLS_U && v_lo[id] >= 0 && v_hi[id] <= 1) { x.self.n = 1; x.self.sup[0] = id; x.self.e[0] = 0; x.self.e[1] = 1; }
            else return false;
          } else { x.self.n = 0; x.self.e[0] = 1; }
          tv[part].push_back(x);
        }
      }
      for (int part = 0; part < 3; part++)
        for (TV& x : tv[part]) {
          const Table* t = x.t ? x.t : &x.self;
          for (int j = 0; j < t->n; j++) {
            int q = 0;
            while (q < nr && roots[q] != t->sup[j]) q++;
            if (q == nr) { if (nr >= 8) return false; roots[nr++] = t->sup[j]; }
            x.pos[j] = q;
          }
        }
      for (int idx = 0; idx < (1 << nr); idx++) {
        i128 acc[3] = {0, 0, 0};
        for (int part = 0; part < 3; part++)
          for (const TV& x : tv[part]) {
            const Table* t = x.t ? x.t : &x.self;
            int sub = 0;
            for (int j = 0; j < t->n; j++) sub |= ((idx >> x.pos[j]) & 1) << j;
            i128 term;
            if (__builtin_mul_overflow(x.coef, (i128)t->e[sub], &term) || __builtin_add_overflow(acc[part], term, &acc[part])) return false;
          }
        i128 ab;
        if (__builtin_mul_overflow(acc[0], acc[1], &ab)) return false;
        // |values| far below p / 2: equality over the integers is equality in the field
        const i128 lim = (i128)1 << 120;
        if (ab > lim || ab < -lim || acc[2] > lim || acc[2] < -lim) return false;
        if (ab != acc[2]) return false;
      }
      return true;
    };
    // Symbolic proof from the op stream itself: every value is replaced, latest definition first, by the
    // expression its defining op computes (ADD / SUB / multiplication by a constant / conversions, in U,
    // I or F class - all exact: narrow ops only exist where the interval analysis excludes a wrap) until
    // the linear form cancels.  For A * B = C with single-wire A and B the product x * y is matched against
    // the MUL op that defines a wire of C.  This is what makes `z <== x + 2^k * y` and `z <== x * y`
    // hold for every input; rows the expansion cannot close stay run-time checks.
    auto prove_symbolic = [&](const std::vector<MT>* parts) -> bool {
      if (!opt.symbolic_rows_static) return false;
      std::map<uint32_t, U256> lin;  // value -> coefficient of (A*B - C), constant one under 0xFFFFFFFF
      auto add = [&](uint32_t v, const U256& c) {
        if (v != 0xFFFFFFFFu) { U256 k; if (const_of(v, k)) { lin[0xFFFFFFFFu] = fr_add(lin[0xFFFFFFFFu], fr_mul(c, k)); return; } }
        lin[v] = fr_add(lin[v], c);
      };
      for (const MT& m : parts[2]) add(m.val, fr_neg(m.c));
      bool quad = !parts[0].empty() && !parts[1].empty();
      uint32_t qx = 0, qy = 0; U256 qc;
      if (quad) {
        if (parts[0].size() != 1 || parts[1].size() != 1) return false;
        qx = parts[0][0].val; qy = parts[1][0].val;
        if (qx == 0xFFFFFFFFu || qy == 0xFFFFFFFFu) return false;
        qc = fr_mul(parts[0][0].c, parts[1][0].c);
      }
      auto same = [&](uint32_t x, uint32_t y) { return same_value(x, y); };
      for (int step = 0; step < 48; step++) {
        // latest-defined value with a non-zero coefficient
        uint32_t pick = 0; bool any = false;
        for (auto it = lin.begin(); it != lin.end();) {
          if (it->second.is_zero()) { it = lin.erase(it); continue; }
          if (it->first != 0xFFFFFFFFu && (!any || it->first > pick)) { pick = it->first; any = true; }
          ++it;
        }
        if (!any) { auto one = lin.find(0xFFFFFFFFu); return !quad && (one == lin.end() || one->second.is_zero()); }
        uint32_t di = v_def[pick];
        if (di >= nops || ops[di].dst != pick) return false;
        const OpRec& o = ops[di];
        U256 c = lin[pick];
        lin.erase(pick);
        const bool imm = (o.flags & PZK_FLAG_B_IMM) != 0, pool = (o.flags & PZK_FLAG_B_POOL) != 0;
        switch (o.opc) {
          case PZK_U_ADD: case PZK_U_SUB: case PZK_F_ADD: case PZK_F_SUB: case PZK_Z_ADD: case PZK_Z_SUB: {
            const bool sub_ = (o.opc == PZK_U_SUB || o.opc == PZK_F_SUB || o.opc == PZK_Z_SUB);
            const bool zop = (o.opc == PZK_Z_ADD || o.opc == PZK_Z_SUB);
            add(o.a, c);
            U256 cb = sub_ ? fr_neg(c) : c;
            if (imm) add(0xFFFFFFFFu, fr_mul(cb, U256((uint64_t)o.b)));
            else if (pool) add(0xFFFFFFFFu, fr_mul(cb, zop ? z_field(fpool[o.b]) : fr_from_mont(fpool[o.b])));
            else add(o.b, cb);
            break;
          }
          case PZK_U_MUL: case PZK_F_MUL: case PZK_Z_MUL: {
            U256 k;
            if (imm) { add(o.a, fr_mul(c, U256((uint64_t)o.b))); break; }
            if (pool) { add(o.a, fr_mul(c, o.opc == PZK_Z_MUL ? z_field(fpool[o.b]) : fr_from_mont(fpool[o.b]))); break; }
            if (const_of(o.b, k)) { add(o.a, fr_mul(c, k)); break; }
            if (const_of(o.a, k)) { add(o.b, fr_mul(c, k)); break; }
            // a genuine product: it must be the row's own A * B, with exactly the opposite coefficient
            if (!quad) return false;
            if (!((same(o.a, qx) && same(o.b, qy)) || (same(o.a, qy) && same(o.b, qx)))) return false;
            if (!(fr_add(c, qc).is_zero())) return false;
            quad = false;
            break;
          }
          case PZK_F_NEG: add(o.a, fr_neg(c)); break;
          case PZK_F_FROM_U: case PZK_F_FROM_I: case PZK_Z_FROM_U: case PZK_Z_FROM_I: add(o.a, c); break;
          case PZK_F_FROM_N: if (o.flags & PZK_FLAG_ZSRC) { add(o.a, c); break; } return false;
          default: return false;
        }
      }
      return false;
    };
    // Bit-view proof: every wire of the row is a bit field of some word (or a constant): the row is expanded over
    // the individual bits of those words, bit(W, t) in {0, 1} unknown, and holds for every input when all
    // coefficients cancel - `div * 2 + bit * bit === in` of GetLastBitUnsecure, `check[N-1] + div * 2^N === in`,
    // `in === sum[LEN-1]` of Num2Bits when the interval of `in` fits LEN bits
    // (/root/reference/circuits/lib/circuits/int/arithmetic.circom:161-204, bitify/bitify.circom:10-32).  When the
    // only bits left are the high bits W[T..] of one word, each with weight c * 2^t, the row IS the range check
    // (W >> T) == 0 and is lowered to one CHECK_RANGE record (returns 2).
    struct BitTerm { uint32_t base; uint32_t bit; U256 c; };
    std::vector<BitTerm> bt;
    auto prove_views = [&](const std::vector<MT>* parts, uint32_t& rbase, int& rbits) -> int {
      if (!opt.views) return 0;
      bt.clear();
      U256 konst;
      auto expand = [&](uint32_t val, const U256& c) -> bool {
        if (val == 0xFFFFFFFFu) { konst = fr_add(konst, c); return true; }
        U256 k;
        if (const_of(val, k)) { konst = fr_add(konst, fr_mul(c, k)); return true; }
        ViewD d;
        if (!src_view(val, d)) {
          if ((v_cls[val] != CLS_F && v_cls[val] != CLS_Z) || !v_convN[val]) return false;
          uint32_t nf = v_convN[val];
          if (v_def[nf] >= nops || !keep[v_def[nf]] || ops[v_def[nf]].opc != PZK_N_FROM_F || ops[v_def[nf]].dst != nf) return false;
          d.base = nf; d.s = 0; d.n = 254; d.k = 0;
        }
        U256 w = c;
        for (int t = 0; t < d.k; t++) w = fr_add(w, w);
        for (int t = 0; t < d.n; t++) { bt.push_back({d.base, (uint32_t)(d.s + t), w}); w = fr_add(w, w); }
        return true;
      };
      if (!parts[0].empty() && !parts[1].empty()) {
        if (parts[0].size() != 1 || parts[1].size() != 1) return 0;
        uint32_t x = parts[0][0].val, y = parts[1][0].val;
        if (x == 0xFFFFFFFFu || y == 0xFFFFFFFFu) return 0;
        ViewD dx, dy;
        if (!src_view(x, dx) || !src_view(y, dy)) return 0;
        if (dx.n != 1 || dx.k != 0 || dy.n != 1 || dy.k != 0 || dx.base != dy.base || dx.s != dy.s) return 0;
        bt.push_back({dx.base, dx.s, fr_mul(parts[0][0].c, parts[1][0].c)});  // bit * bit = bit
      }
      for (const MT& m : parts[2]) if (!expand(m.val, fr_neg(m.c))) return 0;
      if (!konst.is_zero()) return 0;
      std::sort(bt.begin(), bt.end(), [](const BitTerm& a, const BitTerm& b) { return a.base != b.base ? a.base < b.base : a.bit < b.bit; });
      size_t w = 0;
      for (size_t i = 0; i < bt.size();) {
        BitTerm acc = bt[i]; size_t j = i + 1;
        while (j < bt.size() && bt[j].base == acc.base && bt[j].bit == acc.bit) { acc.c = fr_add(acc.c, bt[j].c); j++; }
        if (!acc.c.is_zero()) bt[w++] = acc;
        i = j;
      }
      bt.resize(w);
      if (bt.empty()) return 1;
      // residual = c * sum_{t >= T} 2^t bit(W, t) ?
      uint32_t B = bt[0].base;
      ViewD id;
      if (!src_view(B, id) || id.base != B) return 0;  // B must be a word
      uint32_t T = bt[0].bit;
      if (T == 0 || bt.size() != (size_t)id.n - T) return 0;
      U256 c = bt[0].c;
      for (size_t i = 0; i < bt.size(); i++) {
        if (bt[i].base != B || bt[i].bit != T + i || !(bt[i].c == c)) return 0;
        c = fr_add(c, c);
      }
      // c0 * 2^(t - T) with c0 != 0: c0 * 2^-T * (W >> T) * 2^T ... the sum is c0 / 2^T * (W - (W mod 2^T)), zero iff W >> T == 0
      rbase = B; rbits = (int)T;
      return 2;
    };
    // plain views of the scalar program (packed truth tables come later): what the bit-view prover reads
    vw.assign(nv, ViewD()); v_tabview.assign(nv, 0); alias_of.assign(nv, 0);
    if (opt.views)
      for (size_t i = 0; i < nops; i++) {
        if (!keep[i]) continue;
        const OpRec& o = ops[i];
        if (o.opc == PZK_NOP || o.opc == PZK_ASSERT_NZ || is_macro(o.opc) || !o.dst) continue;
        ViewD d;
        if (try_view(o, d)) vw[o.dst] = d;
      }
    for (size_t r = 0; r < nrows; r++) {
      uint32_t rbase = 0; int rbits = 0, pv = 0;
      if (merge_row((uint32_t)r, parts)) { row_static[r] = 1; n_static_rows++; }
      else if (prove_by_tables(parts)) { row_static[r] = 2; n_table_rows++; }
      else if (prove_symbolic(parts)) { row_static[r] = 3; n_symbolic_rows++; }
      else if ((pv = prove_views(parts, rbase, rbits)) == 1) { row_static[r] = 5; n_view_rows++; }
      else if (pv == 2) { range_rows[(uint32_t)r] = RangeRow{rbase, (uint16_t)rbits}; n_range_rows++; }
      else if (getenv("PZK_DUMP_ROWS") && !(rows[r].by_def && opt.def_rows_static)) {
        static int dumped = 0;
        if (dumped++ < atoi(getenv("PZK_DUMP_ROWS"))) {
          fprintf(stderr, "row %zu:", r);
          for (int part = 0; part < 3; part++) {
            fprintf(stderr, " [");
            for (auto& m : parts[part]) {
              if (m.val == 0xFFFFFFFFu) { fprintf(stderr, " 1*%llx", (unsigned long long)m.c.w[0]); continue; }
              ViewD d; bool hv = src_view(m.val, d);
              fprintf(stderr, " v%u(cls%d opc%d hi=%llx view=%d:%u,%d,%d,%d t=%s)*%llx", m.val, (int)v_cls[m.val], v_def[m.val] < nops ? (int)ops[v_def[m.val]].opc : -1, (unsigned long long)v_hi[m.val], (int)hv, d.base, d.s, d.n, d.k,
                      (op_stats && v_def[m.val] < op_tag.size() && op_tag[v_def[m.val]] >= 0) ? nm(op_tag[v_def[m.val]]).c_str() : "?", (unsigned long long)m.c.w[0]);
            }
            fprintf(stderr, " ]");
          }
          fprintf(stderr, "\n");
        }
      }
      // `x <== e` stores value(e) into x and adds the row e - x = 0: the wire holds the very value the
      // row compares it with, so the row holds for every input (the same argument as for aliases, one
      // multiplication deeper).  Only `===` rows and rows over `<--` hints can fail at run time.
      else if (rows[r].by_def && opt.def_rows_static) { row_static[r] = 4; n_def_rows++; }
    }
  }
  // ---- views and packed truth tables: rewrite the op list (the proofs above read the scalar program)
  if (opt.views) {
    std::vector<uint8_t> row_needed(nv, 0);
    for (size_t r = 0; r < nrows; r++) {
      if (row_static[r]) continue;
      auto rr = range_rows.find((uint32_t)r);
      if (rr != range_rows.end()) { row_needed[rr->second.base] = 1; continue; }
      uint32_t nt = rows[r].na + rows[r].nb + rows[r].nc;
      for (uint32_t t = 0; t < nt; t++) { uint32_t sig = terms[rows[r].off + t].first; if (sig != 0xFFFFFFFFu && sig_val[sig]) row_needed[sig_val[sig]] = 1; }
    }
    vectorize_sweep(keep, row_needed, for_operands);
    nv = v_cls.size(); nops = ops.size();
    v_def.assign(nv, 0xFFFFFFFFu);
    for (size_t i = 0; i < nops; i++) for_defs(ops[i], [&](uint32_t d) { v_def[d] = (uint32_t)i; });
    // what must exist after the rewrite: words that signals are exported from (own slot or as the base / a root
    // of a view), values of run-time rows; everything else that nothing reads goes
    std::vector<uint8_t> need(nv, 0);
    auto need_word = [&](uint32_t v) { v = canon(v); if (v < nv) need[v] = 1; };
    for (uint32_t v : sig_val) {
      if (!v) continue;
      if (vw[v].base) need_word(vw[v].base);
      else if (v_tabview[v]) { const Table& tb = tables[v_tbl[v]]; for (int j = 0; j < tb.n; j++) { uint32_t root = canon(tb.sup[j]); if (vw[root].base && v_def[root] == 0xFFFFFFFFu) need_word(vw[root].base); else need_word(root); } }
      else need_word(v);
    }
    for (size_t r = 0; r < nrows; r++) {
      if (row_static[r]) continue;
      auto rr = range_rows.find((uint32_t)r);
      if (rr != range_rows.end()) { need_word(rr->second.base); continue; }
      uint32_t nt = rows[r].na + rows[r].nb + rows[r].nc;
      for (uint32_t t = 0; t < nt; t++) { uint32_t sig = terms[rows[r].off + t].first; if (sig != 0xFFFFFFFFu && sig_val[sig]) need_word(sig_val[sig]); }
    }
    for (size_t i = nops; i-- > 0;) {
      const OpRec& o = ops[i];
      bool live = (o.opc == PZK_ASSERT_NZ) || (o.opc == PZK_IN_U) || (o.opc == PZK_IN_F) || (o.opc == PZK_NOP && i == 0);
      if (!live) for_defs(o, [&](uint32_t d) { if (need[d]) live = true; });
      keep[i] = live;
      if (live) for_operands(o, [&](uint32_t v) { if (v != PZK_OPERAND_NONE) need[v] = 1; });
    }
    used.assign(nv, 1);
    print_op_stats("V");
  }
  // ---- rows: each one is checked right after the op that defines the last of its wires
  size_t first_kept = 0;
  while (first_kept < nops && !keep[first_kept]) first_kept++;
  std::vector<uint32_t> row_trigger(nrows, (uint32_t)first_kept), row_recs(nrows, 1);
  // values a run-time row reads (a range row reads its word only)
  auto row_values = [&](uint32_t r, const std::function<void(uint32_t)>& f) {
    auto rr = range_rows.find(r);
    if (rr != range_rows.end()) { f(canon(rr->second.base)); return; }
    uint32_t nt = rows[r].na + rows[r].nb + rows[r].nc;
    for (uint32_t t = 0; t < nt; t++) {
      uint32_t sig = terms[rows[r].off + t].first;
      if (sig == 0xFFFFFFFFu) continue;
      uint32_t v = sig_val[sig];
      if (v) f(canon(v));
    }
  };
  for (size_t r = 0; r < nrows; r++) {
    if (row_kind[r]) continue;
    uint32_t nt = rows[r].na + rows[r].nb + rows[r].nc, live_terms = 0;
    for (uint32_t t = 0; t < nt; t++) {
      uint32_t sig = terms[rows[r].off + t].first;
      if (sig == 0xFFFFFFFFu || sig_val[sig]) live_terms++;
    }
    row_values((uint32_t)r, [&](uint32_t v) { if (v_def[v] != 0xFFFFFFFFu && v_def[v] > row_trigger[r]) row_trigger[r] = v_def[v]; });
    row_recs[r] = range_rows.count((uint32_t)r) ? 1 : 1 + (live_terms + 1) / 2;
  }

  std::vector<uint32_t> row_order(nrows);
  for (size_t r = 0; r < nrows; r++) row_order[r] = (uint32_t)r;
  std::stable_sort(row_order.begin(), row_order.end(), [&](uint32_t a, uint32_t b) { return row_trigger[a] < row_trigger[b]; });
  // ---- peephole: x + (z * 2^k) with a single-use product that is not a wire -> one U_SHLADD record
  // (weighted bit sums: `lc += bit * 2^k`).  Same value in wrapping 64-bit arithmetic, one record less.
  if (opt.fuse_shladd) {
    std::vector<uint32_t> uses(nv, 0), def_op(nv, 0xFFFFFFFFu);
    std::vector<uint8_t> is_sig(nv, 0);
    for (uint32_t v : sig_val) if (v && !(v < vw.size() && (vw[v].base || v_tabview[v]))) is_sig[canon(v)] = 1;
    for (size_t i = 0; i < nops; i++) if (keep[i]) {
      for_operands(ops[i], [&](uint32_t v) { if (v != PZK_OPERAND_NONE) uses[v]++; });
      for_defs(ops[i], [&](uint32_t d) { def_op[d] = (uint32_t)i; });
    }
    auto shift_of = [&](uint32_t v, uint32_t& src, int& k) -> bool {
      if (v >= nv || def_op[v] == 0xFFFFFFFFu || is_sig[v] || uses[v] != 1) return false;
      const OpRec& d = ops[def_op[v]];
      if (d.opc != PZK_U_MUL || !(d.flags & PZK_FLAG_B_IMM) || d.dst != v) return false;
      if (d.b == 0 || (d.b & (d.b - 1)) != 0) return false;
      src = d.a; k = __builtin_ctz(d.b);
      return true;
    };
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      OpRec& o = ops[i];
      if (o.opc != PZK_U_ADD || (o.flags & PZK_FLAG_B_IMM)) continue;
      uint32_t src; int k;
      if (shift_of(o.b, src, k)) { keep[def_op[o.b]] = 0; o.opc = PZK_U_SHLADD; o.b = src; o.imm16 = (uint16_t)k; n_fused++; }
      else if (shift_of(o.a, src, k)) { keep[def_op[o.a]] = 0; o.opc = PZK_U_SHLADD; o.a = o.b; o.b = src; o.imm16 = (uint16_t)k; n_fused++; }
    }
  }
  // words a signal's export entry reads: itself, the base of its view, or the roots of its table view
  auto sig_words = [&](uint32_t v, const std::function<void(uint32_t)>& f) {
    if (v < vw.size() && vw[v].base) { f(canon(vw[v].base)); return; }
    if (v < v_tabview.size() && v_tabview[v]) {
      const Table& tb = tables[v_tbl[v]];
      for (int j = 0; j < tb.n; j++) {
        uint32_t root = canon(tb.sup[j]);
        if (vw[root].base && v_def[root] == 0xFFFFFFFFu) f(canon(vw[root].base)); else f(root);
      }
      return;
    }
    f(canon(v));
  };
  // ---- peephole: x +- (p * q) with a single-use product that is not a wire -> one F_MULADD / Z_MULADD record (the
  // linear combinations of Poseidon's mix layers, the column sums of the big-integer products).  The product's
  // dispatch, its store and the sum's re-load go away; the value of the sum is the same field / integer element.
  if (opt.fuse_muladd) {
    std::vector<uint32_t> uses(nv, 0), def_op(nv, 0xFFFFFFFFu);
    std::vector<uint8_t> is_sig(nv, 0);
    for (uint32_t v : sig_val) if (v) { is_sig[canon(v)] = 1; if (v < nv) is_sig[v] = 1; sig_words(v, [&](uint32_t w) { is_sig[w] = 1; }); }
    for (size_t i = 0; i < nops; i++) if (keep[i]) {
      for_operands(ops[i], [&](uint32_t v) { if (v != PZK_OPERAND_NONE) uses[v]++; });
      for_defs(ops[i], [&](uint32_t d) { def_op[d] = (uint32_t)i; });
    }
    for (size_t r = 0; r < nrows; r++) if (!row_static[r]) row_values((uint32_t)r, [&](uint32_t v) { uses[v]++; });
    // words other signals are views of keep their own record (their descriptor carries the view table)
    std::vector<uint8_t> is_view_base(nv, 0);
    for (uint32_t v : sig_val) if (v) sig_words(v, [&](uint32_t w) { if (w != canon(v)) is_view_base[w] = 1; });
    // a product that IS a wire (every limb product of the RSA multiplier, the S-box products of Poseidon) can move
    // into the sum as well when the sum is its only reader - no other op, no run-time row: it becomes a second result
    // of the record (PZK_FLAG_DST2), digested there, stored only when witnesses are exported
    auto product_of = [&](uint32_t v, int mul_opc) -> int64_t {
      if (v >= nv || def_op[v] == 0xFFFFFFFFu || uses[v] != 1) return -1;
      if (is_sig[v] && (!opt.fuse_muladd_wires || (v < vw.size() && (vw[v].base || v_tabview[v])) || is_view_base[v])) return -1;
      const OpRec& d = ops[def_op[v]];
      if (d.opc != mul_opc || d.dst != v || (d.flags & (PZK_FLAG_EXT | PZK_FLAG_B_IMM))) return -1;
      return (int64_t)def_op[v];
    };
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      OpRec& o = ops[i];
      const bool fop = (o.opc == PZK_F_ADD || o.opc == PZK_F_SUB), zop = (o.opc == PZK_Z_ADD || o.opc == PZK_Z_SUB);
      if ((!fop && !zop) || (o.flags & (PZK_FLAG_B_POOL | PZK_FLAG_B_IMM | PZK_FLAG_EXT))) continue;
      const int mul_opc = fop ? PZK_F_MUL : PZK_Z_MUL;
      const bool sub = (o.opc == PZK_F_SUB || o.opc == PZK_Z_SUB);
      int64_t di; uint32_t c; bool neg_prod, neg_c;
      if ((di = product_of(o.b, mul_opc)) >= 0) { c = o.a; neg_prod = sub; neg_c = false; }
      else if ((di = product_of(o.a, mul_opc)) >= 0) { c = o.b; neg_prod = false; neg_c = sub; }
      else continue;
      const OpRec d = ops[(size_t)di];
      keep[(size_t)di] = 0;
      const bool wire = is_sig[d.dst] != 0;
      o.opc = fop ? PZK_F_MULADD : PZK_Z_MULADD;
      o.flags = (uint8_t)(PZK_FLAG_EXT | (d.flags & PZK_FLAG_B_POOL) | (wire ? PZK_FLAG_DST2 : 0));
      o.imm16 = (uint16_t)((zop ? (d.imm16 & 0xff) : 0) | (neg_prod ? 0x100 : 0) | (neg_c ? 0x200 : 0));
      o.a = d.a; o.b = d.b; o.c = c; o.d = wire ? d.dst : PZK_OPERAND_NONE; o.e = 0; o.f = 0;
      n_fused_mac++;
      if (wire) n_fused_mac_wire++;
    }
    // Z products read 64-bit factors where they are - a U word in the narrow plane or its cache cell - instead of
    // through the 256-bit copy Z_FROM_U makes (the limbs of the RSA operands: 32 bytes per read became 8)
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      OpRec& o = ops[i];
      if (o.opc != PZK_Z_MUL && o.opc != PZK_Z_MULADD) continue;
      auto narrow_src = [&](uint32_t v, uint32_t& src) -> bool {
        if (v >= nv || def_op[v] == 0xFFFFFFFFu || !keep[def_op[v]] || is_sig[v]) return false;
        const OpRec& d = ops[def_op[v]];
        if (d.opc != PZK_Z_FROM_U || d.dst != v) return false;
        src = d.a;
        return true;
      };
      uint32_t src;
      if (narrow_src(o.a, src)) {
        const uint32_t old = o.a;
        o.a = src; o.flags |= PZK_FLAG_A_U; uses[src]++; n_z_u_operands++;
        if (--uses[old] == 0) keep[def_op[old]] = 0;
      }
      if (!(o.flags & PZK_FLAG_B_POOL) && narrow_src(o.b, src)) {
        const uint32_t old = o.b;
        o.b = src; o.flags |= PZK_FLAG_B_U; uses[src]++; n_z_u_operands++;
        if (--uses[old] == 0) keep[def_op[old]] = 0;
      }
    }
  }
  // Fused witness digest (pzk_program.h, PZK_FLAG_DIG): an op whose result is a wire, or the word behind bit-field
  // views that are wires, is followed by a digest descriptor so that the evaluator can fold the value when it is
  // defined instead of storing and re-reading it.  Export entries the descriptor cannot carry (truth-table views
  // over several words, views wider than 64 bits) are folded after the segment from the global slots of their
  // words: those words must always be stored.
  std::vector<uint8_t> dig_work(nv, 0), dig_global(nv, 0);
  if (opt.fused_digest) {
    for (uint32_t sgi = 0; sgi < sig_val.size(); sgi++) {
      uint32_t v = sig_val[sgi];
      if (!v) continue;
      if (v < vw.size() && vw[v].base) {
        const ViewD& d = vw[v];
        uint32_t b = canon(d.base);
        const uint32_t width = (v_cls[b] == CLS_U) ? 64 : 256;
        if ((uint32_t)d.n + d.k <= 64 && d.s < width) dig_work[b] = 1; else { dig_global[b] = 1; n_dig_wide++; }
      } else if (v < v_tabview.size() && v_tabview[v]) {
        sig_words(v, [&](uint32_t w) { dig_global[w] = 1; });
      } else {
        dig_work[canon(v)] = 1;
        uint32_t dv = v_def[canon(v)];
        if (dv < ops.size() && is_macro(ops[dv].opc)) n_dig_macro++;
      }
    }
  }
  // ---- segments over kept ops (+ their rows)
  std::vector<uint32_t> def_seg(nv, 0), last_seg(nv, 0), op_seg(nops, 0);
  {
    uint64_t rec = 0; uint32_t seg = 0;
    size_t rp = 0;
    bool after_solo = false;
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      uint64_t need = ((ops[i].flags & PZK_FLAG_EXT) ? 2 : 1) + ((ops[i].opc == PZK_V_LUT && (ops[i].flags & PZK_FLAG_W64)) ? 1 : 0) + ((opt.fused_digest && !is_macro(ops[i].opc) && ops[i].dst && ops[i].dst < dig_work.size() && dig_work[ops[i].dst]) ? 1 : 0) + ((opt.fused_digest && (ops[i].flags & PZK_FLAG_DST2) && (ops[i].opc == PZK_F_MULADD || ops[i].opc == PZK_Z_MULADD) && dig_work[ops[i].d]) ? 1 : 0);
      size_t q = rp;
      while (q < nrows && row_trigger[row_order[q]] == i) { if (!row_static[row_order[q]]) need += row_recs[row_order[q]]; q++; }
      // the BabyJubjub ladder gets a segment of its own: the runtime runs it as a dedicated kernel
      const bool solo = ops[i].opc == PZK_BJJ_MUL8;   // first record of its segment (the rows it completes follow it)
      if (rec && (rec + need > opt.seg_ops || solo)) { seg++; rec = 0; }
      (void)after_solo;
      rec += need;
      op_seg[i] = seg;
      rp = q;
      for_defs(ops[i], [&](uint32_t d) { def_seg[d] = seg; last_seg[d] = seg; });
    }
    segs.assign(seg + 1, PzkSegment());
  }
  for (size_t i = 0; i < nops; i++) {
    if (!keep[i]) continue;
    uint32_t sg = op_seg[i];
    for_operands(ops[i], [&](uint32_t v) { if (last_seg[v] < sg) last_seg[v] = sg; });
  }
  for (size_t r = 0; r < nrows; r++) {
    if (row_static[r]) continue;
    uint32_t sg = op_seg[row_trigger[r]];
    row_values((uint32_t)r, [&](uint32_t v) { if (last_seg[v] < sg) last_seg[v] = sg; });
  }
  // an export entry is read after the segment that defines the last of its words: all of them live until then
  std::vector<uint32_t> sig_seg(sig_val.size(), 0);
  for (uint32_t sg_i = 0; sg_i < sig_val.size(); sg_i++) {
    uint32_t v = sig_val[sg_i];
    if (!v) continue;
    uint32_t sg = 0;
    sig_words(v, [&](uint32_t w) { if (def_seg[w] > sg) sg = def_seg[w]; });
    sig_words(v, [&](uint32_t w) { if (last_seg[w] < sg) last_seg[w] = sg; });
    sig_seg[sg_i] = sg;
  }
  // ---- slot allocation (free at segment boundaries)
  v_slot.assign(nv, 0xFFFFFFFFu);
  std::vector<std::vector<uint32_t>> dying(segs.size());
  std::vector<uint32_t> free_u, free_f;
  uint32_t next_u = 0, next_f = 0;
  {
    uint32_t cur = 0;
    auto release = [&](uint32_t seg) {
      for (uint32_t v : dying[seg]) {
        if (v_cls[v] == CLS_U || v_cls[v] == CLS_I) free_u.push_back(v_slot[v]); else free_f.push_back(v_slot[v]);
      }
    };
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      while (cur < op_seg[i]) { release(cur); cur++; }
      for_defs(ops[i], [&](uint32_t d) {
        bool narrow = (v_cls[d] == CLS_U || v_cls[d] == CLS_I);
        uint32_t slot;
        if (narrow) { if (!free_u.empty()) { slot = free_u.back(); free_u.pop_back(); } else slot = next_u++; }
        else { if (!free_f.empty()) { slot = free_f.back(); free_f.pop_back(); } else slot = next_f++; }
        v_slot[d] = slot;
        dying[last_seg[d]].push_back(d);
      });
    }
  }
  n_u_slots = next_u; n_f_slots = next_f;
  if (getenv("PZK_LIVE_STATS")) {
    // where do the F slots go: values per (defining opcode, lifetime in segments)
    std::map<std::pair<int, int>, uint64_t> hist;
    std::vector<uint32_t> live(segs.size() + 1, 0);
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      for_defs(ops[i], [&](uint32_t d) {
        if (v_cls[d] == CLS_U || v_cls[d] == CLS_I) return;
        int life = (int)(last_seg[d] - def_seg[d]);
        int b = life == 0 ? 0 : life < 4 ? 1 : life < 32 ? 2 : life < 256 ? 3 : 4;
        hist[{ops[i].opc, b}]++;
        for (uint32_t sgi = def_seg[d]; sgi <= last_seg[d]; sgi++) live[sgi]++;
      });
    }
    for (auto& kv : hist) fprintf(stderr, "F def opc %d life-bucket %d: %llu\n", kv.first.first, kv.first.second, (unsigned long long)kv.second);
    uint32_t mx = 0, at_ = 0;
    for (size_t i = 0; i < segs.size(); i++) if (live[i] > mx) { mx = live[i]; at_ = (uint32_t)i; }
    fprintf(stderr, "max live F %u at segment %u of %zu\n", mx, at_, segs.size());
    for (size_t i = 0; i < segs.size(); i += segs.size() / 40 + 1) fprintf(stderr, " seg %zu live F %u\n", i, live[i]);
  }
  // ---- emit op records, each followed by the constraint rows it completes
  auto slot_of = [&](uint32_t v) -> uint32_t {
    if (v == PZK_OPERAND_NONE) return v;
    if (v_slot[v] == 0xFFFFFFFFu)
      throw CompileError(fmt("internal: operand without a slot (value %u class %d, defined by op %u opc %d kept %d, used=%d)", v,
                             (int)v_cls[v], v_def[v], v_def[v] < ops.size() ? (int)ops[v_def[v]].opc : -1,
                             v_def[v] < ops.size() ? (int)keep[v_def[v]] : -1, (int)used[v]));
    return v_slot[v];
  };
  auto ref_of_sig = [&](uint32_t sig) -> uint32_t {
    if (sig == 0xFFFFFFFFu) return PZK_REF_ONE;
    uint32_t v = sig_val[sig];
    if (!v) return PZK_REF_ZERO;
    uint32_t cls = v_cls[v] == CLS_U ? 0u : (v_cls[v] == CLS_I ? 1u : (v_cls[v] == CLS_Z ? 3u : 2u));
    return (cls << 30) | v_slot[v];
  };
  std::unordered_map<uint64_t, uint32_t> icoef_off;  // int64 coefficient -> list offset
  std::function<uint32_t(uint32_t)> row_ref, row_opnd;
  auto emit_row = [&](uint32_t r) {
    std::vector<MT> parts[3];
    {
      auto rr = range_rows.find(r);
      if (rr != range_rows.end()) {  // the row is (word >> bits) == 0
        uint32_t b = canon(rr->second.base);
        PzkOp h; h.opc = PZK_CHECK_RANGE; h.flags = (v_cls[b] == CLS_U) ? 0 : PZK_FLAG_NBASE; h.imm16 = rr->second.bits;
        h.dst = r; h.a = row_opnd(b); h.b = 0;
        out_ops.push_back(h);
        n_range_emitted++;
        check_bytes += (v_cls[b] == CLS_U) ? 8 : 32;
        return;
      }
    }
    if (merge_row(r, parts)) return;
    // classification by compile-time bounds
    bool is_int = true, all32 = true;
    unsigned __int128 bound[3] = {0, 0, 0};
    const unsigned __int128 SAT = (unsigned __int128)1 << 127;
    for (int part = 0; part < 3 && is_int; part++)
      for (auto& m : parts[part]) {
        int64_t cv;
        if (!small_signed(m.c, cv)) { is_int = false; break; }
        if (cv < INT32_MIN || cv > INT32_MAX) all32 = false;
        unsigned __int128 mag = (unsigned __int128)(cv < 0 ? -cv : cv), vmax = 1;
        if (m.val != 0xFFFFFFFFu) {
          uint32_t v = m.val;
          if (v_cls[v] != CLS_U && v_cls[v] != CLS_I) { is_int = false; break; }
          i128 a = v_lo[v] < 0 ? -v_lo[v] : v_lo[v], b = v_hi[v] < 0 ? -v_hi[v] : v_hi[v];
          vmax = (unsigned __int128)(a > b ? a : b);
        }
        bound[part] += mag * vmax;
        if (bound[part] >= SAT) { is_int = false; break; }
      }
    const unsigned __int128 LIM63 = (unsigned __int128)1 << 63, LIM126 = (unsigned __int128)1 << 126;
    if (is_int && (bound[0] >= LIM63 || bound[1] >= LIM63 || bound[2] >= LIM126)) is_int = false;
    bool is_i64 = is_int && all32 && bound[2] < LIM63 && (bound[0] == 0 || bound[1] == 0 || bound[0] * bound[1] < LIM63);
    std::vector<PzkTerm> ts;
    uint16_t cnt[3] = {0, 0, 0};
    for (int part = 0; part < 3; part++)
      for (auto& m : parts[part]) {
        PzkTerm pt;
        pt.ref = (m.val == 0xFFFFFFFFu) ? PZK_REF_ONE : row_ref(m.val);
        if (is_int) {
          int64_t cv; small_signed(m.c, cv);
          if (cv >= INT32_MIN && cv <= INT32_MAX) pt.coef = (uint32_t)(int32_t)cv;
          else {
            auto it = icoef_off.find((uint64_t)cv);
            uint32_t lo;
            if (it == icoef_off.end()) {
              lo = (uint32_t)out_list.size();
              out_list.push_back((uint32_t)(uint64_t)cv); out_list.push_back((uint32_t)((uint64_t)cv >> 32));
              icoef_off[(uint64_t)cv] = lo;
            } else lo = it->second;
            pt.coef = lo;
            if (pt.ref != PZK_REF_ONE) pt.ref |= PZK_TERM_COEF_LIST; else pt.ref = PZK_REF_ONE_LIST;
          }
        } else {
          pt.coef = coef_id(m.c);
          if (m.val != 0xFFFFFFFFu && v_cls[m.val] == CLS_U && v_lo[m.val] >= 0 && v_hi[m.val] <= 1) pt.ref |= PZK_TERM_BIT;
        }
        ts.push_back(pt); cnt[part]++;
      }
    PzkOp h; h.opc = is_i64 ? PZK_CHECK_I64 : (is_int ? PZK_CHECK_INT : PZK_CHECK_F); h.flags = 0; h.imm16 = cnt[0];
    h.dst = r; h.a = (uint32_t)cnt[1] | ((uint32_t)cnt[2] << 16); h.b = (uint32_t)((ts.size() + 1) / 2);
    out_ops.push_back(h);
    for (size_t k = 0; k < ts.size(); k += 2) {
      uint32_t wds[4] = {ts[k].ref, ts[k].coef, PZK_REF_ZERO, 0};
      if (k + 1 < ts.size()) { wds[2] = ts[k + 1].ref; wds[3] = ts[k + 1].coef; }
      PzkOp raw; memcpy(&raw, wds, 16);
      out_ops.push_back(raw);
    }
    if (is_i64) n_i64_rows++; else if (is_int) n_int_rows++; else n_field_rows++;
    for (auto& t : ts) if (t.ref < PZK_REF_ONE_LIST) check_bytes += (PZK_REF_CLS(t.ref) >= 2) ? 32 : 8;
  };
  out_list = list_pool;
  out_ops.clear();
  // ---- operand cache: use lists (positions in event order: op, then the rows it completes)
  const uint32_t NC = opt.cells;
  seg_quads.assign(segs.size(), 0);
  {
    std::vector<uint64_t> nu(segs.size(), 0), nf(segs.size(), 0);
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i] || is_macro(ops[i].opc)) continue;
      for_defs(ops[i], [&](uint32_t d) { ((v_cls[d] == CLS_U || v_cls[d] == CLS_I) ? nu : nf)[op_seg[i]]++; });
    }
    for (size_t sg = 0; sg < segs.size(); sg++) {
      if (!nf[sg] || NC < 8) continue;
      double share = 4.0 * nf[sg] / (4.0 * nf[sg] + nu[sg]);
      uint32_t q = (uint32_t)(NC / 4 * share + 0.5);
      if (q < 1) q = 1;
      if (nu[sg] && q > NC / 4 - 2) q = NC / 4 - 2;
      if (q > NC / 4) q = NC / 4;
      seg_quads[sg] = q;
    }
  }
  std::vector<uint32_t> use_cnt(nv + 1, 0);
  std::vector<uint32_t> op_pos(nops, 0), seg_last_pos(segs.size(), 0);
  {
    uint32_t pos = 0; size_t rp = 0;
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      op_pos[i] = pos;
      if (!is_macro(ops[i].opc)) for_operands(ops[i], [&](uint32_t v) { use_cnt[v]++; });
      pos++;
      while (rp < nrows && row_trigger[row_order[rp]] == i) {
        uint32_t r = row_order[rp], nt = rows[r].na + rows[r].nb + rows[r].nc;
        if (row_static[r]) { rp++; continue; }
        (void)nt; row_values(r, [&](uint32_t v) { use_cnt[v]++; });
        pos++; rp++;
      }
      seg_last_pos[op_seg[i]] = pos;
    }
  }
  std::vector<uint64_t> use_off(nv + 1, 0);
  for (size_t v = 0; v < nv; v++) use_off[v + 1] = use_off[v] + use_cnt[v];
  std::vector<uint32_t> use_pos(use_off[nv]), use_fill(nv, 0);
  {
    uint32_t pos = 0; size_t rp = 0;
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      if (!is_macro(ops[i].opc)) for_operands(ops[i], [&](uint32_t v) { use_pos[use_off[v] + use_fill[v]++] = pos; });
      pos++;
      while (rp < nrows && row_trigger[row_order[rp]] == i) {
        uint32_t r = row_order[rp], nt = rows[r].na + rows[r].nb + rows[r].nc;
        if (row_static[r]) { rp++; continue; }
        (void)nt; row_values(r, [&](uint32_t v) { use_pos[use_off[v] + use_fill[v]++] = pos; });
        pos++; rp++;
      }
    }
  }
  std::vector<uint32_t> use_idx(nv, 0);
  std::vector<int32_t> cell_of(nv, -1);
  std::vector<uint32_t> cur_next(nv, 0);
  std::vector<uint8_t> needs_global(nv, 0);
  int pass = 0;
  const uint32_t INF = 0xFFFFFFFFu;
  typedef std::pair<uint32_t, uint32_t> HeapE;  // (next use, value)
  std::priority_queue<HeapE> heap_u, heap_f;
  std::vector<uint32_t> free_cells, free_quads, cached_list;
  uint64_t cache_hits = 0, cache_miss = 0;
  auto next_use_after = [&](uint32_t v, uint32_t pos, uint32_t seg_end) -> uint32_t {
    uint32_t& k = use_idx[v];
    while (k < use_cnt[v] && use_pos[use_off[v] + k] <= pos) k++;
    if (k >= use_cnt[v]) return INF;
    uint32_t nx = use_pos[use_off[v] + k];
    return nx < seg_end ? nx : INF;
  };
  auto release = [&](uint32_t v) {
    if (cell_of[v] < 0) return;
    if (v_cls[v] == CLS_U || v_cls[v] == CLS_I) free_cells.push_back((uint32_t)cell_of[v]); else free_quads.push_back((uint32_t)cell_of[v]);
    cell_of[v] = -1;
  };
  auto reset_cache = [&](uint32_t seg) {
    for (uint32_t v : cached_list) cell_of[v] = -1;
    cached_list.clear();
    while (!heap_u.empty()) heap_u.pop();
    while (!heap_f.empty()) heap_f.pop();
    // partition the cells between singles and quads by the segment's definition mix
    uint64_t ub = 0, fb = 0;
    (void)seg;
    free_cells.clear(); free_quads.clear();
    uint32_t quads = seg_quads[seg];
    uint32_t singles = NC - 4 * quads;
    for (uint32_t c = singles; c-- > 0;) free_cells.push_back(c);
    for (uint32_t q = quads; q-- > 0;) free_quads.push_back(singles + 4 * q);
    (void)ub; (void)fb;
  };
  auto touch_operand = [&](uint32_t v, uint32_t pos, uint32_t seg_end) {  // after the event used v
    if (v == PZK_OPERAND_NONE || cell_of[v] < 0) return;
    uint32_t nx = next_use_after(v, pos, seg_end);
    if (nx == INF) { release(v); return; }
    if (nx != cur_next[v]) {
      cur_next[v] = nx;
      ((v_cls[v] == CLS_U || v_cls[v] == CLS_I) ? heap_u : heap_f).push({nx, v});
    }
  };
  auto try_cache_def = [&](uint32_t d, uint32_t pos, uint32_t seg_end) -> int32_t {
    if (NC == 0) return -1;
    uint32_t nx = next_use_after(d, pos, seg_end);
    if (nx == INF) return -1;
    bool narrow = (v_cls[d] == CLS_U || v_cls[d] == CLS_I);
    std::vector<uint32_t>& fr = narrow ? free_cells : free_quads;
    std::priority_queue<HeapE>& hp = narrow ? heap_u : heap_f;
    if (fr.empty()) {
      // evict the cached value whose next use is furthest away, if it is further than ours
      while (!hp.empty()) {
        HeapE top = hp.top();
        if (cell_of[top.second] < 0 || cur_next[top.second] != top.first) { hp.pop(); continue; }
        if (top.first <= nx) return -1;
        hp.pop();
        release(top.second);
        break;
      }
      if (fr.empty()) return -1;
    }
    uint32_t c = fr.back(); fr.pop_back();
    cell_of[d] = (int32_t)c; cur_next[d] = nx;
    hp.push({nx, d});
    cached_list.push_back(d);
    return (int32_t)c;
  };
  auto opnd = [&](uint32_t v) -> uint32_t {
    if (v == PZK_OPERAND_NONE) return v;
    if (cell_of[v] >= 0) { cache_hits++; return PZK_OPERAND_CELL | (uint32_t)cell_of[v]; }
    cache_miss++;
    needs_global[v] = 1;
    return slot_of(v);
  };
  auto tref = [&](uint32_t v) -> uint32_t {
    uint32_t cls = v_cls[v] == CLS_U ? 0u : (v_cls[v] == CLS_I ? 1u : (v_cls[v] == CLS_Z ? 3u : 2u));
    if (cell_of[v] >= 0) { cache_hits++; return (cls << 30) | PZK_TERM_CELL | (uint32_t)cell_of[v]; }
    cache_miss++;
    needs_global[v] = 1;
    return (cls << 30) | v_slot[v];
  };
  row_ref = tref;
  row_opnd = opnd;
  // values that must always reach their global slot: public wires (read by the export kernel)
  for (uint32_t sg = 0; sg < sig_val.size(); sg++) if (sig_val[sg] && sig2wire[sg] <= n_pub_out + n_pub_in) sig_words(sig_val[sg], [&](uint32_t w) { needs_global[w] = 1; });
  for (size_t v = 0; v < nv; v++) if (dig_global[v]) needs_global[v] = 1;
  for (pass = 0; pass < 2; pass++) {
    // pass 0 learns which values are ever read from their global slot; pass 1 emits
    out_list = list_pool; out_ops.clear(); icoef_off.clear();
    std::fill(use_idx.begin(), use_idx.end(), 0u);
    cache_hits = cache_miss = 0; n_i64_rows = n_int_rows = n_field_rows = 0; eval_bytes = check_bytes = 0; n_range_emitted = 0;
    uint32_t cur = 0xFFFFFFFFu;
    size_t rp = 0;
    uint32_t pos = 0;
    for (size_t i = 0; i < nops; i++) {
      if (!keep[i]) continue;
      const OpRec& o = ops[i];
      uint32_t s = op_seg[i];
      if (s != cur) { cur = s; segs[s].op_off = out_ops.size(); reset_cache(s); }
      const uint32_t seg_end = seg_last_pos[s];
      PzkOp r; r.opc = o.opc; r.flags = o.flags; r.imm16 = o.imm16; r.dst = 0; r.a = o.a; r.b = o.b;
      if (o.opc == PZK_U_ADD || o.opc == PZK_U_MUL || o.opc == PZK_U_AND || o.opc == PZK_U_SHR || o.opc == PZK_U_SHLADD) r.flags |= PZK_FLAG_FAST;
      auto cls_bytes = [&](uint32_t v) -> uint64_t { return (v == PZK_OPERAND_NONE) ? 0 : ((v_cls[v] == CLS_U || v_cls[v] == CLS_I) ? 8 : 32); };
      for_operands(o, [&](uint32_t v) { eval_bytes += cls_bytes(v); });
      for_defs(o, [&](uint32_t v) { eval_bytes += cls_bytes(v); });
      uint32_t ext_c = PZK_OPERAND_NONE, ext_d = PZK_OPERAND_NONE;
      bool has_dst = false;
      switch (o.opc) {
        case PZK_NOP: break;
        case PZK_U_CONST: case PZK_F_CONST: case PZK_Z_CONST: case PZK_IN_U: case PZK_IN_F: has_dst = true; break;
        case PZK_ASSERT_NZ: r.a = opnd(o.a); break;
        case PZK_BIGDIV: case PZK_MODINV: case PZK_BJJ_MUL8: {
          uint32_t op0, nop, def0, ndef; macro_layout(o, op0, nop, def0, ndef);
          for (uint32_t j = op0; j < def0 + ndef; j++) { out_list[j] = slot_of(list_pool[j]); needs_global[list_pool[j]] = 1; }
          break;
        }
        case PZK_N_BIT: case PZK_F_CSEL: case PZK_U_EXTRACT: case PZK_N_EXTRACT: has_dst = true; r.a = opnd(o.a); break;
        case PZK_U_LUT: case PZK_U_LUTV: case PZK_V_LUT: has_dst = true; r.a = opnd(o.a); r.b = opnd(o.b); ext_c = opnd(o.c); ext_d = opnd(o.d); break;
        case PZK_U_SEL: case PZK_F_SEL: has_dst = true; r.a = opnd(o.a); r.b = opnd(o.b); ext_c = opnd(o.c); break;
        case PZK_Z_MUL: has_dst = true; r.a = opnd(o.a); if (!(o.flags & PZK_FLAG_B_POOL)) r.b = opnd(o.b); break;
        case PZK_F_MULADD: case PZK_Z_MULADD:
          has_dst = true; r.a = opnd(o.a); if (!(o.flags & PZK_FLAG_B_POOL)) r.b = opnd(o.b); ext_c = opnd(o.c);
          if (o.flags & PZK_FLAG_DST2) {
            // the product: nobody reads it but the export / digest; stored only on request
            ext_d = slot_of(o.d);
            if (ext_d > 0x3fffffu) throw CompileError("too many live slots for the dst encoding");
            if (pass == 1 && !needs_global[o.d]) ext_d |= PZK_DST_OPTIONAL;
          }
          break;
        default:
          has_dst = true; r.a = opnd(o.a);
          if (!(o.flags & (PZK_FLAG_B_IMM | PZK_FLAG_B_POOL))) {
            switch (o.opc) {
              case PZK_F_NEG: case PZK_F_INV: case PZK_F_FROM_U: case PZK_F_FROM_I: case PZK_N_FROM_F: case PZK_Z_FROM_U: case PZK_Z_FROM_I:
              case PZK_F_FROM_N: case PZK_N_FROM_U: case PZK_N_LOW: case PZK_N_FITS: r.b = 0; break;
              default: r.b = opnd(o.b);
            }
          }
      }
      // operands die / get re-prioritised, then the result may take a cell
      if (!is_macro(o.opc)) for_operands(o, [&](uint32_t v) { touch_operand(v, pos, seg_end); });
      if (has_dst) {
        uint32_t slot = slot_of(o.dst);
        if (slot > 0x3fffffu) throw CompileError("too many live slots for the dst encoding");
        int32_t cell = try_cache_def(o.dst, pos, seg_end);
        if (cell >= 0 && cell > 509) throw CompileError("too many cache cells for the dst encoding");
        r.dst = slot | (cell >= 0 ? ((uint32_t)cell + 1) << 22 : 0);
        if (pass == 1 && cell >= 0 && !needs_global[o.dst]) r.dst |= PZK_DST_OPTIONAL;
      }
      const bool with_dig = has_dst && !is_macro(o.opc) && o.dst < dig_work.size() && dig_work[o.dst];
      if (with_dig) r.flags |= PZK_FLAG_DIG;
      const bool with_dig2 = (o.opc == PZK_F_MULADD || o.opc == PZK_Z_MULADD) && (o.flags & PZK_FLAG_DST2) && opt.fused_digest && dig_work[o.d];
      if (with_dig2) r.flags |= PZK_FLAG_DIG2;
      out_ops.push_back(r);
      if (o.flags & PZK_FLAG_EXT) {
        PzkOpExt x; x.c = ext_c; x.d = ext_d; x.e = o.e; x.f = o.f;
        if (o.opc == PZK_U_LUTV && !opt.views) x.e = lutv_off[(uint32_t)i];
        PzkOp raw; memcpy(&raw, &x, sizeof raw);
        out_ops.push_back(raw);
        if (o.opc == PZK_V_LUT && (o.flags & PZK_FLAG_W64)) {
          PzkOpExt y; y.c = o.g; y.d = 0; y.e = 0; y.f = 0;
          memcpy(&raw, &y, sizeof raw);
          out_ops.push_back(raw);
        }
      }
      if (with_dig2) {  // the product's descriptor comes first: the evaluator folds it before it forms the sum
        uint32_t wds[4] = {0u, 1u, slot_of(o.d), 0u};
        PzkOp raw; memcpy(&raw, wds, 16);
        out_ops.push_back(raw);
      }
      if (with_dig) {
        // descriptor: {0, plane (0 = U, 1 = F), slot, 0}; the runtime fills in weights / table when it loads the program
        const bool narrow_d = (v_cls[o.dst] == CLS_U || v_cls[o.dst] == CLS_I);
        uint32_t wds[4] = {0u, narrow_d ? 0u : 1u, slot_of(o.dst), 0u};
        PzkOp raw; memcpy(&raw, wds, 16);
        out_ops.push_back(raw);
      }
      pos++;
      while (rp < nrows && row_trigger[row_order[rp]] == i) {
        uint32_t rr = row_order[rp];
        if (row_static[rr]) { rp++; continue; }
        emit_row(rr);
        uint32_t nt = rows[rr].na + rows[rr].nb + rows[rr].nc;
        (void)nt; row_values(rr, [&](uint32_t v) { touch_operand(v, pos, seg_end); });
        pos++; rp++;
      }
      segs[s].n_ops = out_ops.size() - segs[s].op_off;
    }
    if (rp != nrows) throw CompileError("internal: constraint rows left unplaced");
  }
  cache_hit_refs = cache_hits; cache_miss_refs = cache_miss;
  for (size_t s2 = 0; s2 < segs.size(); s2++) { segs[s2].row_off = 0; segs[s2].n_rows = 0; }
  // ---- exports per segment
  {
    std::vector<std::vector<PzkExport>> by_seg(segs.size());
    for (uint32_t sig = 0; sig < sig_val.size(); sig++) {
      uint32_t v = sig_val[sig];
      PzkExport e; e.wire = sig2wire[sig]; e.aux = 0; e.pad = 0;
      if (!v) { e.ref = PZK_REF_ZERO; by_seg[0].push_back(e); continue; }
      if (v < vw.size() && vw[v].base) {
        const ViewD& d = vw[v];
        uint32_t b = canon(d.base);
        e.ref = (3u << 30) | ((v_cls[b] == CLS_U) ? 0u : PZK_REF_VIEW_N) | slot_of(b);
        e.aux = (uint32_t)d.s | ((uint32_t)d.n << 8) | ((uint32_t)d.k << 16);
        n_view_sigs++;
      } else if (v < v_tabview.size() && v_tabview[v]) {
        const Table& tb = tables[v_tbl[v]];
        e.ref = PZK_REF_TABVIEW; e.aux = (uint32_t)out_list.size();
        out_list.push_back(tb.n);
        for (int j = 0; j < tb.n; j++) {
          uint32_t root = canon(tb.sup[j]);
          if (vw[root].base && v_def[root] == 0xFFFFFFFFu) { out_list.push_back(slot_of(canon(vw[root].base))); out_list.push_back(vw[root].s); }
          else { out_list.push_back(slot_of(root)); out_list.push_back(0); }
        }
        for (int k = 0; k < (1 << tb.n); k++) { uint64_t x = (uint64_t)tb.e[k]; out_list.push_back((uint32_t)x); out_list.push_back((uint32_t)(x >> 32)); }
        n_tabview_sigs++;
      } else {
        uint32_t cv = canon(v);
        uint32_t cls = v_cls[cv] == CLS_U ? 0u : (v_cls[cv] == CLS_I ? 1u : 2u);
        e.ref = (cls << 30) | (v_cls[cv] == CLS_Z ? PZK_REF_Z : 0u) | slot_of(cv);
      }
      by_seg[sig_seg[sig]].push_back(e);
    }
    for (size_t s = 0; s < segs.size(); s++) {
      segs[s].exp_off = out_exports.size();
      out_exports.insert(out_exports.end(), by_seg[s].begin(), by_seg[s].end());
      segs[s].n_exp = out_exports.size() - segs[s].exp_off;
    }
  }
  stats->n_ops = out_ops.size();
  stats->n_u_slots = n_u_slots; stats->n_f_slots = n_f_slots; stats->n_segments = (uint32_t)segs.size();
}

static void json_dims(std::string& s, const std::vector<int>& d) {
  s += "[";
  for (size_t i = 0; i < d.size(); i++) { if (i) s += ","; s += std::to_string(d[i]); }
  s += "]";
}

void Compiler::Impl::build_meta() {
  std::string s = "{";
  s += "\"main\":\"" + nm(unit.main.call->name) + "\",";
  s += "\"n_wires\":" + std::to_string(sig_val.size() + 1) + ",";
  s += "\"n_constraints\":" + std::to_string(rows.size()) + ",";
  s += "\"inputs\":[";
  uint32_t off = 0;
  for (size_t i = 0; i < main_inputs_desc.size(); i++) {
    if (i) s += ",";
    uint32_t cnt = prod(main_inputs_desc[i].second);
    auto it = opt.input_bits.find(main_inputs_desc[i].first);
    s += "{\"name\":\"" + main_inputs_desc[i].first + "\",\"dims\":";
    json_dims(s, main_inputs_desc[i].second);
    s += ",\"offset\":" + std::to_string(off) + ",\"size\":" + std::to_string(cnt) +
         ",\"bits\":" + std::to_string(it == opt.input_bits.end() ? 0 : it->second) + "}";
    off += cnt;
  }
  s += "],\"outputs\":[";
  off = 0;
  for (size_t i = 0; i < main_outputs_desc.size(); i++) {
    if (i) s += ",";
    uint32_t cnt = prod(main_outputs_desc[i].second);
    s += "{\"name\":\"" + main_outputs_desc[i].first + "\",\"dims\":";
    json_dims(s, main_outputs_desc[i].second);
    s += ",\"offset\":" + std::to_string(off) + ",\"size\":" + std::to_string(cnt) + "}";
    off += cnt;
  }
  s += "],\"n_pub_out\":" + std::to_string(n_pub_out) + ",\"n_pub_in\":" + std::to_string(n_pub_in) +
       ",\"n_prv_in\":" + std::to_string(n_prv_in);
  s += ",\"stats\":{\"u_ops\":" + std::to_string(stats->u_ops) + ",\"f_mul\":" + std::to_string(stats->f_mul) +
       ",\"f_inv\":" + std::to_string(stats->f_inv) + ",\"f_inv_real\":" + std::to_string(stats->f_inv_real) + ",\"f_other\":" + std::to_string(stats->f_other) +
       ",\"bigdiv\":" + std::to_string(stats->bigdiv) + ",\"modinv\":" + std::to_string(stats->modinv) + ",\"z_ops\":" + std::to_string(stats->z_ops) + ",\"z_mul\":" + std::to_string(stats->z_mul) + ",\"lut\":" + std::to_string(stats->lut) +
       ",\"op_records\":" + std::to_string(out_ops.size()) + ",\"segments\":" + std::to_string(segs.size()) +
       ",\"static_rows\":" + std::to_string(n_static_rows) + ",\"def_rows\":" + std::to_string(n_def_rows) + ",\"table_rows\":" + std::to_string(n_table_rows) + ",\"symbolic_rows\":" + std::to_string(n_symbolic_rows) + ",\"view_rows\":" + std::to_string(n_view_rows) + ",\"range_rows\":" + std::to_string(n_range_rows) + ",\"vlut\":" + std::to_string(n_vlut) + ",\"vlut_lanes\":" + std::to_string(n_vlut_lanes) + ",\"view_signals\":" + std::to_string(n_view_sigs) + ",\"tabview_signals\":" + std::to_string(n_tabview_sigs) + ",\"extracts\":" + std::to_string(n_extracts) + ",\"fused_shladd\":" + std::to_string(n_fused) + ",\"fused_muladd\":" + std::to_string(n_fused_mac) + ",\"fused_muladd_wire_products\":" + std::to_string(n_fused_mac_wire) + ",\"z_u_operands\":" + std::to_string(n_z_u_operands) + ",\"digest_wide_views\":" + std::to_string(n_dig_wide) + ",\"digest_macro_defs\":" + std::to_string(n_dig_macro) + ",\"i64_rows\":" + std::to_string(n_i64_rows) +
       ",\"int_rows\":" + std::to_string(n_int_rows) + ",\"field_rows\":" + std::to_string(n_field_rows) +
       ",\"eval_bytes\":" + std::to_string(eval_bytes) + ",\"check_bytes\":" + std::to_string(check_bytes) +
       ",\"cells\":" + std::to_string(opt.cells) + ",\"cache_hit_refs\":" + std::to_string(cache_hit_refs) +
       ",\"cache_miss_refs\":" + std::to_string(cache_miss_refs) +
       ",\"u_slots\":" + std::to_string(n_u_slots) + ",\"f_slots\":" + std::to_string(n_f_slots) + "}";
  s += "}";
  meta_json = s;
}

// ---------------------------------------------------------------------------------------
Compiler::Compiler(const std::string& main_path, const CompileOptions& opt) : im(new Impl()) {
  im->opt = opt; im->main_path = main_path; im->stats = &stats;
}
Compiler::~Compiler() {
  for (Layout* l : im->layout_list) delete l;
  for (Comp* c : im->comps) delete c;
  delete im;
}
void Compiler::run() {
  auto t0 = std::chrono::steady_clock::now();
  load_file(im->unit, im->main_path, true, {});
  im->run_main();
  im->backend();
  im->build_meta();
  stats.seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
std::string Compiler::main_io_json() const { return im->meta_json; }

static void wr(FILE* f, const void* p, size_t n) { if (n && fwrite(p, 1, n, f) != n) throw CompileError("write failed"); }
static void pad16(FILE* f, uint64_t& pos) { static const char z[16] = {0}; uint64_t r = (16 - pos % 16) % 16; wr(f, z, r); pos += r; }

void Compiler::write_program(const std::string& path) {
  Impl& m = *im;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw CompileError("cannot write " + path);
  PzkHeader h; memset(&h, 0, sizeof h);
  h.magic = PZK_MAGIC; h.version = PZK_VERSION;
  h.n_wires = (uint32_t)m.sig_val.size() + 1;
  h.n_pub_out = m.n_pub_out; h.n_pub_in = m.n_pub_in; h.n_prv_in = m.n_prv_in;
  h.n_constraints = (uint32_t)m.rows.size();
  h.n_u_slots = m.n_u_slots; h.n_f_slots = m.n_f_slots;
  h.n_segments = (uint32_t)m.segs.size();
  h.n_fpool = (uint32_t)m.fpool.size();
  h.n_coef = (uint32_t)m.coefs.size();
  h.n_inputs = (uint32_t)m.out_inputs.size();
  h.n_list = (uint32_t)m.out_list.size();
  h.n_op_records = m.out_ops.size();
  h.n_rows = m.out_rows.size(); h.n_terms = m.out_terms.size(); h.n_exports = m.out_exports.size();
  h.stat_u_ops = stats.u_ops; h.stat_f_mul = stats.f_mul; h.stat_f_inv = stats.f_inv;
  h.stat_f_other = stats.f_other; h.stat_bigdiv = stats.bigdiv;
  h.reserved[0] = m.meta_json.size();
  h.reserved[1] = m.opt.cells;
  uint64_t pos = 0;
  auto section = [&](const void* p, size_t n) { wr(f, p, n); pos += n; pad16(f, pos); };
  section(&h, sizeof h);
  section(m.segs.data(), m.segs.size() * sizeof(PzkSegment));
  section(m.out_ops.data(), m.out_ops.size() * sizeof(PzkOp));
  section(m.fpool.data(), m.fpool.size() * 32);
  {
    std::vector<PzkCoef> cs(m.coefs.size());
    for (size_t i = 0; i < cs.size(); i++) {
      U256 mm = fr_to_mont(m.coefs[i]), m2 = fr_to_mont(mm);
      memcpy(cs[i].plain, m.coefs[i].w, 32); memcpy(cs[i].mont, mm.w, 32); memcpy(cs[i].mont2, m2.w, 32);
    }
    section(cs.data(), cs.size() * sizeof(PzkCoef));
  }
  section(m.out_list.data(), m.out_list.size() * 4);
  section(m.out_inputs.data(), m.out_inputs.size() * sizeof(PzkInput));
  section(m.out_rows.data(), m.out_rows.size() * sizeof(PzkRow));
  section(m.out_terms.data(), m.out_terms.size() * sizeof(PzkTerm));
  section(m.out_exports.data(), m.out_exports.size() * sizeof(PzkExport));
  section(m.meta_json.data(), m.meta_json.size());
  fclose(f);
}

// one byte per constraint (.r1cs order): how the row is discharged - 0 checked at run time, 1 alias,
// 2 table proof, 3 symbolic proof, 4 definitional (only with def_rows_static).  Lets a test evaluate
// every row independently and confirm that no statically discharged row ever fails.
void Compiler::write_rowkinds(const std::string& path) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw CompileError("cannot write " + path);
  wr(f, im->row_kind.data(), im->row_kind.size());
  fclose(f);
}

// <prefix>.rowsrc: where the constraints that are checked at run time come from (template, file, line) - what the
// circom runtime prints behind "Assert Failed." ("Error in template X line: N").  Binary, little endian:
//   "PZKS" u32 n_files {u32 len, bytes}* u32 n_templates {u32 len, bytes}* u32 n_entries {u32 row, u32 line, u16 file, u16 template}*
// Rows discharged at compile time cannot fail and are left out.
void Compiler::write_rowsrc(const std::string& path) {
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw CompileError("cannot write " + path);
  auto u32w = [&](uint32_t v) { wr(f, &v, 4); };
  auto strw = [&](const std::string& t) { u32w((uint32_t)t.size()); wr(f, t.data(), t.size()); };
  wr(f, "PZKS", 4);
  u32w((uint32_t)im->unit.files.size());
  for (auto& t : im->unit.files) strw(t);
  std::vector<int32_t> tids; std::unordered_map<int32_t, uint16_t> tix;
  for (size_t r = 0; r < im->rows.size(); r++) {
    if (r < im->row_kind.size() && im->row_kind[r]) continue;
    int32_t t = im->rows[r].tname;
    if (!tix.count(t)) { tix[t] = (uint16_t)tids.size(); tids.push_back(t); }
  }
  if (tids.size() > 0xffff) throw CompileError("too many templates for the .rowsrc encoding");
  u32w((uint32_t)tids.size());
  for (int32_t t : tids) strw(t >= 0 ? im->unit.names.str(t) : std::string("?"));
  uint32_t n = 0;
  for (size_t r = 0; r < im->rows.size(); r++) if (!(r < im->row_kind.size() && im->row_kind[r])) n++;
  u32w(n);
  for (size_t r = 0; r < im->rows.size(); r++) {
    if (r < im->row_kind.size() && im->row_kind[r]) continue;
    const RowRec& rr = im->rows[r];
    u32w((uint32_t)r); u32w((uint32_t)rr.line);
    uint16_t fi = (uint16_t)(rr.file >= 0 ? rr.file : 0xffff), ti = tix[rr.tname];
    wr(f, &fi, 2); wr(f, &ti, 2);
  }
  fclose(f);
}

// iden3 r1cs v1
void Compiler::write_r1cs(const std::string& path) {
  Impl& m = *im;
  FILE* f = fopen(path.c_str(), "wb");
  if (!f) throw CompileError("cannot write " + path);
  auto u32 = [&](uint32_t v) { wr(f, &v, 4); };
  auto u64 = [&](uint64_t v) { wr(f, &v, 8); };
  uint32_t n_wires = (uint32_t)m.sig_val.size() + 1;
  wr(f, "r1cs", 4); u32(1); u32(3);
  // header section
  u32(1); u64(4 + 32 + 4 * 4 + 8 + 4);
  u32(32); wr(f, FR_P.w, 32);
  u32(n_wires); u32(m.n_pub_out); u32(m.n_pub_in); u32(m.n_prv_in);
  u64(n_wires); u32((uint32_t)m.rows.size());
  // constraints section
  uint64_t size = 0;
  for (auto& r : m.rows) size += 12 + (uint64_t)(r.na + r.nb + r.nc) * 36;
  u32(2); u64(size);
  std::vector<std::pair<uint32_t, uint32_t>> tmp;
  for (auto& r : m.rows) {
    uint32_t lens[3] = {r.na, r.nb, r.nc};
    uint64_t off = r.off;
    for (int part = 0; part < 3; part++) {
      tmp.clear();
      for (uint32_t t = 0; t < lens[part]; t++, off++) {
        uint32_t sig = m.terms[off].first;
        tmp.emplace_back(sig == 0xFFFFFFFFu ? 0u : m.sig2wire[sig], m.terms[off].second);
      }
      std::sort(tmp.begin(), tmp.end());
      u32((uint32_t)tmp.size());
      for (auto& t : tmp) { u32(t.first); wr(f, m.coefs[t.second].w, 32); }
    }
  }
  // wire -> label map (identity: O0 layout)
  u32(3); u64((uint64_t)n_wires * 8);
  for (uint32_t i = 0; i < n_wires; i++) u64(i);
  fclose(f);
}

// ---------------------------------------------------------------------------------------
// O1-style simplification of the emitted constraint system (the reference's library circuits are compiled with
// `circom --O1`, /root/reference/circuits/lib/circuits/scripts/compile-circuit.sh:34): linear constraints that say
// "signal = signal" or "signal = constant" are removed by substitution, repeatedly, and the signals they eliminate
// leave the witness.  Main inputs and outputs are never eliminated.  Writes <r1cs_path> with the surviving wires
// renumbered in their original order and <sym_path> in which eliminated signals carry witness index -1 - the shape of
// the .sym circom writes - so that pzk_circuit_open_ex(program, program.sym, this .sym) produces the matching witness.
// circom itself is not available here: the result is self-consistent (every surviving constraint holds on the
// renumbered witness, tests/) but its numbering is not calibrated against a circom-built file.
// ---------------------------------------------------------------------------------------
void Compiler::write_o1(const std::string& r1cs_path, const std::string& sym_path) {
  Impl& m = *im;
  const uint32_t n_wires = (uint32_t)m.sig_val.size() + 1;
  const uint32_t n_keep = 1 + m.n_pub_out + m.n_pub_in + m.n_prv_in;  // wire 0 and the main IO: never eliminated
  typedef std::vector<std::pair<uint32_t, U256>> LC;  // (wire, coefficient), sorted by wire, no zero coefficients
  struct Con { LC a, b, c; bool dead = false; };
  std::vector<Con> cons(m.rows.size());
  for (size_t r = 0; r < m.rows.size(); r++) {
    const RowRec& rr = m.rows[r];
    uint64_t off = rr.off;
    LC* parts[3] = {&cons[r].a, &cons[r].b, &cons[r].c};
    uint32_t lens[3] = {rr.na, rr.nb, rr.nc};
    for (int part = 0; part < 3; part++)
      for (uint32_t t = 0; t < lens[part]; t++, off++) {
        uint32_t sig = m.terms[off].first;
        parts[part]->emplace_back(sig == 0xFFFFFFFFu ? 0u : m.sig2wire[sig], m.coefs[m.terms[off].second]);
      }
  }
  std::vector<uint32_t> parent(n_wires);
  for (uint32_t i = 0; i < n_wires; i++) parent[i] = i;
  std::vector<uint8_t> is_const(n_wires, 0);
  std::vector<U256> const_val(n_wires);
  std::function<uint32_t(uint32_t)> find = [&](uint32_t x) { while (parent[x] != x) { parent[x] = parent[parent[x]]; x = parent[x]; } return x; };
  auto normalise = [&](LC& lc) {
    std::map<uint32_t, U256> acc;
    for (auto& t : lc) {
      uint32_t w = find(t.first);
      if (w != 0 && is_const[w]) { acc[0] = fr_add(acc[0], fr_mul(t.second, const_val[w])); continue; }
      acc[w] = fr_add(acc[w], t.second);
    }
    lc.clear();
    for (auto& kv : acc) if (!kv.second.is_zero()) lc.emplace_back(kv.first, kv.second);
  };
  auto const_only = [](const LC& lc, U256& k) { if (lc.empty()) { k = U256(); return true; } if (lc.size() == 1 && lc[0].first == 0) { k = lc[0].second; return true; } return false; };
  uint64_t n_alias = 0, n_constant = 0, n_trivial = 0;
  for (int pass = 0; pass < 64; pass++) {
    bool changed = false;
    for (Con& cn : cons) {
      if (cn.dead) continue;
      normalise(cn.a); normalise(cn.b); normalise(cn.c);
      // the linear form L = C - k B (A = k constant) or C - k A (B = k constant)
      U256 k;
      LC lin;
      bool linear = false;
      if (const_only(cn.a, k)) { linear = true; lin = cn.c; for (auto& t : cn.b) lin.emplace_back(t.first, fr_neg(fr_mul(k, t.second))); }
      else if (const_only(cn.b, k)) { linear = true; lin = cn.c; for (auto& t : cn.a) lin.emplace_back(t.first, fr_neg(fr_mul(k, t.second))); }
      if (!linear) continue;
      normalise(lin);
      if (lin.empty()) { cn.dead = true; n_trivial++; changed = true; continue; }
      U256 c0;
      size_t first = 0;
      if (lin[0].first == 0) { c0 = lin[0].second; first = 1; }
      const size_t nsig = lin.size() - first;
      if (nsig == 1) {
        // cx * x + c0 = 0  ->  x = -c0 / cx
        uint32_t x = lin[first].first;
        if (x < n_keep) continue;
        is_const[x] = 1; const_val[x] = fr_mul(fr_neg(c0), fr_inv(lin[first].second));
        cn.dead = true; n_constant++; changed = true;
      } else if (nsig == 2 && c0.is_zero() && fr_add(lin[first].second, lin[first + 1].second).is_zero()) {
        // cx * (x - y) = 0  ->  x = y: the signal that is not main IO (else the later one) is eliminated
        uint32_t x = lin[first].first, y = lin[first + 1].first;   // x < y
        if (y < n_keep) continue;                                  // both are main IO: the constraint stays
        parent[y] = x;
        cn.dead = true; n_alias++; changed = true;
      }
    }
    if (!changed) break;
  }
  for (Con& cn : cons) if (!cn.dead) { normalise(cn.a); normalise(cn.b); normalise(cn.c); }
  // surviving wires, renumbered in their original order
  std::vector<int64_t> renum(n_wires, -1);
  uint32_t n_new = 0;
  for (uint32_t wv = 0; wv < n_wires; wv++)
    if (wv < n_keep || (find(wv) == wv && !is_const[wv])) renum[wv] = n_new++;
  uint32_t n_cons = 0;
  for (Con& cn : cons) if (!cn.dead) n_cons++;
  {
    FILE* f = fopen(r1cs_path.c_str(), "wb");
    if (!f) throw CompileError("cannot write " + r1cs_path);
    auto u32 = [&](uint32_t v) { wr(f, &v, 4); };
    auto u64 = [&](uint64_t v) { wr(f, &v, 8); };
    wr(f, "r1cs", 4); u32(1); u32(3);
    u32(1); u64(4 + 32 + 4 * 4 + 8 + 4);
    u32(32); wr(f, FR_P.w, 32);
    u32(n_new); u32(m.n_pub_out); u32(m.n_pub_in); u32(m.n_prv_in);
    u64(n_wires); u32(n_cons);
    uint64_t size = 0;
    for (Con& cn : cons) if (!cn.dead) size += 12 + (uint64_t)(cn.a.size() + cn.b.size() + cn.c.size()) * 36;
    u32(2); u64(size);
    for (Con& cn : cons) {
      if (cn.dead) continue;
      const LC* parts[3] = {&cn.a, &cn.b, &cn.c};
      for (int part = 0; part < 3; part++) {
        u32((uint32_t)parts[part]->size());
        for (auto& t : *parts[part]) { u32((uint32_t)renum[t.first]); wr(f, t.second.w, 32); }
      }
    }
    // wire -> label map: the original (O0) wire of every surviving one
    u32(3); u64((uint64_t)n_new * 8);
    for (uint32_t wv = 0; wv < n_wires; wv++) if (renum[wv] >= 0) u64(wv);
    fclose(f);
  }
  {
    FILE* f = fopen(sym_path.c_str(), "w");
    if (!f) throw CompileError("cannot write " + sym_path);
    uint32_t comp_counter = 0;
    std::function<void(Layout*, uint32_t, const std::string&)> walk = [&](Layout* lay, uint32_t base, const std::string& prefix) {
      uint32_t cid = comp_counter++;
      std::vector<std::pair<uint32_t, std::string>> lines;
      for (int n : lay->order) {
        SigInfo& s = lay->sigs[n];
        uint32_t cnt = Impl::prod(s.dims);
        std::vector<int> idx(s.dims.size(), 0);
        for (uint32_t k = 0; k < cnt; k++) {
          std::string name = prefix + "." + m.nm(n);
          for (size_t d = 0; d < idx.size(); d++) name += "[" + std::to_string(idx[d]) + "]";
          lines.emplace_back(base + s.off + k, name);
          for (int d = (int)idx.size() - 1; d >= 0; d--) { if (++idx[d] < s.dims[d]) break; idx[d] = 0; }
        }
      }
      std::sort(lines.begin(), lines.end());
      for (auto& l : lines) {
        uint32_t wv = m.sig2wire[l.first];
        fprintf(f, "%u,%lld,%u,%s\n", wv, (long long)renum[wv], cid, l.second.c_str());
      }
      for (auto& ck : lay->child_order) {
        Child& ch = lay->children[ck];
        std::string nm2 = prefix + "." + m.nm(ck.first);
        auto cd = lay->comp_dims.find(ck.first);
        if (cd != lay->comp_dims.end() && !cd->second.empty()) {
          std::vector<int> idx(cd->second.size());
          int rem = ck.second;
          for (int d = (int)cd->second.size() - 1; d >= 0; d--) { idx[d] = rem % cd->second[d]; rem /= cd->second[d]; }
          for (int v : idx) nm2 += "[" + std::to_string(v) + "]";
        }
        walk(ch.lay, base + ch.rel_base, nm2);
      }
    };
    walk(m.main_lay, 0, "main");
    fclose(f);
  }
  m.o1_stats = fmt("{\"wires\":%u,\"constraints\":%u,\"wires_o0\":%u,\"constraints_o0\":%zu,\"aliases\":%llu,\"constants\":%llu,\"trivial\":%llu}",
                   n_new, n_cons, n_wires, m.rows.size(), (unsigned long long)n_alias, (unsigned long long)n_constant, (unsigned long long)n_trivial);
}
std::string Compiler::o1_stats() const { return im->o1_stats; }

void Compiler::write_sym(const std::string& path) {
  Impl& m = *im;
  FILE* f = fopen(path.c_str(), "w");
  if (!f) throw CompileError("cannot write " + path);
  uint32_t comp_counter = 0;
  std::function<void(Layout*, uint32_t, const std::string&)> walk = [&](Layout* lay, uint32_t base, const std::string& prefix) {
    uint32_t cid = comp_counter++;
    std::vector<std::pair<uint32_t, std::string>> lines;
    for (int n : lay->order) {
      SigInfo& s = lay->sigs[n];
      uint32_t cnt = Impl::prod(s.dims);
      std::vector<int> idx(s.dims.size(), 0);
      for (uint32_t k = 0; k < cnt; k++) {
        std::string name = prefix + "." + m.nm(n);
        for (size_t d = 0; d < idx.size(); d++) name += "[" + std::to_string(idx[d]) + "]";
        lines.emplace_back(base + s.off + k, name);
        for (int d = (int)idx.size() - 1; d >= 0; d--) { if (++idx[d] < s.dims[d]) break; idx[d] = 0; }
      }
    }
    std::sort(lines.begin(), lines.end());
    for (auto& l : lines) {
      uint32_t w = m.sig2wire[l.first];
      fprintf(f, "%u,%u,%u,%s\n", w, w, cid, l.second.c_str());
    }
    for (auto& ck : lay->child_order) {
      Child& c = lay->children[ck];
      std::string name = prefix + "." + m.nm(ck.first);
      const std::vector<int>& dims = lay->comp_dims[ck.first];
      if (!dims.empty()) {
        std::vector<int> idx(dims.size());
        int rem = ck.second;
        for (int d = (int)dims.size() - 1; d >= 0; d--) { idx[d] = rem % dims[d]; rem /= dims[d]; }
        for (int v : idx) name += "[" + std::to_string(v) + "]";
      }
      walk(c.lay, base + c.rel_base, name);
    }
  };
  walk(m.main_lay, 0, "main");
  fclose(f);
}

}  // namespace pzk
