// pzkc - command line front door of the circuit compiler:
//   pzkc <main.circom> <out_prefix> [--bits name:width ...] [--seg N] [--no-intrinsics] [--static-def-rows]
// writes <out_prefix>.pzkp / .r1cs / .sym  (the roles of circom's .wasm / .r1cs / .sym,
// /root/reference/circuits/scripts/compile-circuit.sh:34).
#include <cstdio>
#include <cstring>
#include "compiler.hpp"

int main(int argc, char** argv) {
  if (argc < 3) { fprintf(stderr, "usage: pzkc main.circom out_prefix [--bits name:width]... [--seg N]\n"); return 2; }
  pzk::CompileOptions opt;
  for (int i = 3; i < argc; i++) {
    if (!strcmp(argv[i], "--bits") && i + 1 < argc) {
      std::string s = argv[++i]; size_t k = s.find(':');
      if (k == std::string::npos) { fprintf(stderr, "bad --bits\n"); return 2; }
      opt.input_bits[s.substr(0, k)] = atoi(s.c_str() + k + 1);
    } else if (!strcmp(argv[i], "--seg") && i + 1 < argc) opt.seg_ops = (uint32_t)atoi(argv[++i]);
    else if (!strcmp(argv[i], "--cells") && i + 1 < argc) opt.cells = (uint32_t)atoi(argv[++i]);
    else if (!strcmp(argv[i], "--no-intrinsics")) opt.intrinsics = false;
    else if (!strcmp(argv[i], "--static-def-rows")) opt.def_rows_static = true;
    else if (!strcmp(argv[i], "--no-table-proofs")) { opt.table_rows_static = false; opt.views = false; opt.vectorize = false; }
    else if (!strcmp(argv[i], "--no-symbolic-proofs")) opt.symbolic_rows_static = false;
    else if (!strcmp(argv[i], "--no-views")) { opt.views = false; opt.vectorize = false; }
    else if (!strcmp(argv[i], "--no-vectorize")) opt.vectorize = false;
    else { fprintf(stderr, "unknown option %s\n", argv[i]); return 2; }
  }
  try {
    pzk::Compiler c(argv[1], opt);
    c.run();
    std::string p = argv[2];
    c.write_program(p + ".pzkp"); c.write_r1cs(p + ".r1cs"); c.write_sym(p + ".sym"); c.write_rowkinds(p + ".rowkind");
    auto& s = c.stats;
    printf("signals %llu constraints %llu values %llu op_records %llu segments %u u_slots %u f_slots %u\n"
           "u_ops %llu f_mul %llu f_inv %llu f_other %llu bigdiv %llu modinv %llu lut %llu  (%.2fs)\n",
           (unsigned long long)s.n_signals, (unsigned long long)s.n_constraints, (unsigned long long)s.n_values,
           (unsigned long long)s.n_ops, s.n_segments, s.n_u_slots, s.n_f_slots, (unsigned long long)s.u_ops,
           (unsigned long long)s.f_mul, (unsigned long long)s.f_inv, (unsigned long long)s.f_other,
           (unsigned long long)s.bigdiv, (unsigned long long)s.modinv, (unsigned long long)s.lut, s.seconds);
  } catch (std::exception& e) { fprintf(stderr, "pzkc: %s\n", e.what()); return 1; }
  return 0;
}
