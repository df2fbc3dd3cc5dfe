// compiler.hpp - circom -> typed linear SSA ("program") + R1CS + SYM.
//
// From-scratch lowering of a circom circuit's witness computation (what circom's
// generated witness_calculator.wasm executes for the reference, call sites
// /root/reference/test/automatisationTest.js:37-51) into a straight-line program of
// typed ops (include/pzk_program.h), plus the constraint system in iden3 .r1cs form
// and the .sym name table.  Symbolic execution in two phases per template instance:
//   phase A (shape): run the body with signal values unknown to learn the signal
//            and sub-component layout; cached per (template, arguments);
//   phase B (emit):  run the body when the instance's last input has been assigned,
//            emitting one SSA value per computed signal / temporary and one
//            (A, B, C) row per `<==` / `===`.
#pragma once
#include <cstdint>
#include <functional>
#include <string>
#include <vector>

#include "circom_front.hpp"
#include "pzk_program.h"

namespace pzk {

typedef __int128 i128;

struct Lin {
  std::vector<std::pair<uint32_t, U256>> t;  // (signal index, coefficient), sorted by signal
  U256 k;                                    // constant term
};
struct Alg {
  int deg = 0;  // 0 const, 1 linear, 2 quadratic (a*b + c), 3 not representable
  Lin a, b, c;
};
typedef std::shared_ptr<Alg> AlgP;

struct SVal {
  uint8_t kind = 0;  // 0 constant, 1 SSA value, 2 unknown (phase A)
  uint32_t id = 0;
  U256 c;
  AlgP alg;
  static SVal konst(const U256& v) { SVal s; s.kind = 0; s.c = v; return s; }
  static SVal unk() { SVal s; s.kind = 2; return s; }
  static SVal ssa(uint32_t id) { SVal s; s.kind = 1; s.id = id; return s; }
};

struct AVal;
struct Value {
  bool arr = false;
  SVal s;
  std::shared_ptr<AVal> a;
};
struct AVal {
  std::vector<int> dims;
  std::vector<SVal> v;
};

struct SigInfo { uint32_t off; std::vector<int> dims; int kind; };  // kind 0 mid, 1 in, 2 out

struct Layout;
struct Child {
  int tname;
  std::string key;
  uint32_t rel_base;
  Layout* lay;
};
struct Layout {
  int tname;
  std::vector<Value> args;
  std::unordered_map<int, SigInfo> sigs;
  std::vector<int> order;
  std::unordered_map<int, std::vector<int>> comp_dims;
  std::map<std::pair<int, int>, Child> children;
  std::vector<std::pair<int, int>> child_order;
  uint32_t own = 0, total = 0, n_inputs = 0;
};

struct Comp {
  Layout* lay;
  uint32_t base;
  int64_t pending;
  bool ran = false;
  std::map<std::pair<int, int>, Comp*> kids;
  Comp* parent;
};

struct Table {
  uint8_t n;
  uint32_t sup[4];
  int64_t e[16];
};

struct OpRec {
  uint8_t opc, flags;
  uint16_t imm16;
  uint32_t dst, a, b, c, d;
  uint32_t e = 0, f = 0;  // extension words: LUTV list offset / LUT bit positions, V_LUT rotations / lane mask
  uint32_t g = 0;         // V_LUT: lane mask high (64-lane groups)
};

// A value that is a bit field of a word the program computes anyway:
//   value = ((base >> s) & (2^n - 1)) << k          (base: a U word or a plain 256-bit N value)
struct ViewD {
  uint32_t base = 0;  // value id of the word (0 = not a view)
  uint8_t s = 0, n = 0, k = 0;
};

enum { CLS_U = 0, CLS_I = 1, CLS_F = 2, CLS_N = 3, CLS_Z = 4 };

struct CompileOptions {
  std::map<std::string, int> input_bits;  // main input name -> declared width (bits)
  uint32_t seg_ops = 8192;
  uint32_t cells = 16;  // operand-cache cells (8 bytes) per lane in shared memory
  bool verbose = false;
  bool intrinsics = true;
  bool table_rows_static = true;  // prove rows over table-valued wires by exhaustive evaluation
  bool symbolic_rows_static = true;  // prove rows by expanding their wires through the defining ops
  bool fuse_shladd = true;  // x + z * 2^k with a single-use product -> one U_SHLADD record
  bool fuse_muladd = true;  // x +- p * q with a single-use product -> one F_MULADD / Z_MULADD record
  bool fuse_muladd_wires = true;  // ... also when the product is a wire (second result of the record, PZK_FLAG_DST2)
  bool def_rows_static = false;  // discharge the rows of `x <== e` (they hold by construction) at compile time
  bool views = true;     // bit-field views + bit-view row proofs (Num2Bits / GetLastNBits / running sums cost no ops)
  bool vectorize = true; // pack one-bit truth-table ops over rotated words into V_LUT records (needs views)
  bool fused_digest = true;  // digest descriptors behind the ops that define wires (PZK_FLAG_DIG)
  bool zclass = true;    // exact wide integers (limb products, Karatsuba sums) as Z-class integer ops instead of Fr ops
};

struct CompileStats {
  uint64_t n_signals = 0, n_constraints = 0, n_ops = 0, n_values = 0;
  uint64_t u_ops = 0, f_mul = 0, f_inv = 0, f_inv_real = 0, f_other = 0, bigdiv = 0, lut = 0, modinv = 0, bjj = 0, z_ops = 0, z_mul = 0;
  uint32_t n_u_slots = 0, n_f_slots = 0, n_segments = 0;
  double seconds = 0;
};

class Compiler {
 public:
  Compiler(const std::string& main_path, const CompileOptions& opt);
  ~Compiler();
  void run();
  void write_program(const std::string& path);
  void write_r1cs(const std::string& path);
  void write_sym(const std::string& path);
  void write_rowkinds(const std::string& path);
  void write_rowsrc(const std::string& path);  // source location of every run-time constraint
  void write_o1(const std::string& r1cs_path, const std::string& sym_path);  // O1-simplified constraint system + .sym
  std::string o1_stats() const;
  CompileStats stats;
  std::string main_io_json() const;

 private:
  struct Impl;
  Impl* im;
};

}  // namespace pzk
