// pzk_api.cu - host runtime behind include/pzk.h: program loading, tile scheduling, CUDA
// streams/events, .wtns / .r1cs codecs.  No CPU fallback: every compute entry point needs a
// CUDA device and fails with PZK_ENODEVICE otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <unordered_map>
#include <string>
#include <vector>

#include "compiler.hpp"
#include "pzk.h"
#include "pzk_kernels.cuh"
#include "pzk_r1cs.cuh"

using namespace pzkd;

#define CK(call)                                                                                  \
  do {                                                                                            \
    cudaError_t e_ = (call);                                                                      \
    if (e_ != cudaSuccess) {                                                                      \
      set_err(c, std::string(#call) + ": " + cudaGetErrorString(e_));                             \
      return PZK_ECUDA;                                                                           \
    }                                                                                             \
  } while (0)

struct SegDev {
  PzkSegment s;
  std::vector<PzkExport> pub;  // export entries of public wires defined in this segment
  uint64_t pub_off = 0;        // offset into d_pub_entries
};

struct pzk_circuit {
  std::string err;
  int device = 0;
  std::vector<uint8_t> blob;
  PzkHeader h;
  const PzkSegment* segs = nullptr;
  const PzkOp* ops = nullptr;
  const uint64_t* fpool = nullptr;
  const PzkCoef* coefs = nullptr;
  const uint32_t* list = nullptr;
  const PzkInput* inputs = nullptr;
  const PzkRow* rows = nullptr;
  const PzkTerm* terms = nullptr;
  const PzkExport* exports = nullptr;
  std::string meta;
  std::vector<SegDev> seg;
  uint64_t bytes_per_lane = 0;
  size_t smem_bytes = 0;
  uint64_t wave_lanes = 0;  // lanes of one full wave of resident CTAs
  // device copies
  uint4* d_ops = nullptr;
  u64* d_fpool = nullptr;
  PzkCoef* d_coefs = nullptr;
  unsigned char* d_coef_kind = nullptr;
  u64* d_coef_mag = nullptr;
  u32* d_list = nullptr;
  PzkRow* d_rows = nullptr;
  PzkTerm* d_terms = nullptr;
  PzkExport* d_exports = nullptr;
  PzkExport* d_pub_entries = nullptr;
  // tile state
  uint64_t tile_lanes_cfg = 0, L = 0;
  bool tile_auto_capped = false;
  u64* d_U = nullptr;
  u64* d_F = nullptr;
  // batch state
  uint64_t batch = 0, batch_cap = 0;
  u64* d_inputs = nullptr;
  uint64_t d_inputs_bytes = 0;
  bool packed = false;        // layout of the resident batch
  uint2* d_in_table = nullptr;
  uint32_t packed_stride = 0;
  std::vector<uint2> in_table;
  u32* d_status = nullptr;
  unsigned long long* d_first_bad = nullptr;
  u64* d_public = nullptr;
  uint64_t pub_cap = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;  // H2D of the next tile / D2H of the previous one under the evaluator
  std::vector<cudaEvent_t> run_events; // events of the call in flight (destroyed by pzk_sync)
  u64* d_lane_list = nullptr;          // export lanes + output rows of the call in flight
  bool in_flight = false;
  // witness digest (pzk_batch_set_digest)
  bool digest_on = false;
  u64* d_digest = nullptr;             // [batch_cap][4]
  u64* d_dig_state = nullptr;          // [DIG_STATE_PIECES][batch_cap] carry-save accumulators
  DigRec* d_dig_recs = nullptr;        // digest program: records per segment
  ulonglong2* d_dig_tab = nullptr;     // per-bit coefficient tables
  std::vector<std::pair<uint64_t, uint64_t>> dig_seg;  // (offset, count) of each segment's digest records
  bool dig_fused = false;              // the program carries digest descriptors (PZK_FLAG_DIG): fold in the evaluator
  uint64_t dig_fused_entries = 0, dig_kernel_entries = 0;
  // profiling
  bool prof = false;
  double prof_ms[5] = {0, 0, 0, 0, 0};
  uint64_t prof_launches[5] = {0, 0, 0, 0, 0};
  std::vector<std::pair<int, std::pair<cudaEvent_t, cudaEvent_t>>> pending;
  std::vector<double> seg_ms;  // accumulated eval time per segment (profiling)
  std::vector<int> pending_seg;
};

static void set_err(pzk_circuit* c, const std::string& m) { if (c) c->err = m; }
static void set_err(char* err, size_t n, const std::string& m) { if (err && n) { snprintf(err, n, "%s", m.c_str()); } }

static uint64_t al16(uint64_t x) { return (x + 15) & ~15ull; }

const char* pzk_version(void) { return "pzk 0.1.0 (sm_100a)"; }

int pzk_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int pzk_compile(const char* main_circom_path, const char* out_prefix, const char* const* bits_names,
                const int* bits_widths, int n_bits, uint32_t segment_ops, char* err, size_t err_len) {
  return pzk_compile_ex(main_circom_path, out_prefix, bits_names, bits_widths, n_bits, segment_ops, 0, err, err_len);
}

int pzk_compile_ex(const char* main_circom_path, const char* out_prefix, const char* const* bits_names,
                   const int* bits_widths, int n_bits, uint32_t segment_ops, uint32_t flags, char* err,
                   size_t err_len) {
  try {
    pzk::CompileOptions opt;
    for (int i = 0; i < n_bits; i++) opt.input_bits[bits_names[i]] = bits_widths[i];
    if (segment_ops) opt.seg_ops = segment_ops;
    if (const char* e = getenv("PZK_CELLS")) { int v = atoi(e); if (v >= 0 && v <= 128 && v % 4 == 0) opt.cells = (uint32_t)v; }  // tuning knob
    opt.def_rows_static = (flags & PZK_COMPILE_STATIC_DEF_ROWS) != 0;
    opt.intrinsics = (flags & PZK_COMPILE_NO_INTRINSICS) == 0;
    opt.table_rows_static = (flags & PZK_COMPILE_NO_TABLE_PROOFS) == 0;
    opt.symbolic_rows_static = (flags & PZK_COMPILE_NO_TABLE_PROOFS) == 0;
    opt.views = (flags & (PZK_COMPILE_NO_VIEWS | PZK_COMPILE_NO_TABLE_PROOFS)) == 0;
    opt.vectorize = opt.views && (flags & PZK_COMPILE_NO_VECTORIZE) == 0;
    pzk::Compiler cc(main_circom_path, opt);
    cc.run();
    std::string p = out_prefix;
    cc.write_program(p + ".pzkp");
    cc.write_r1cs(p + ".r1cs");
    cc.write_sym(p + ".sym");
    cc.write_rowkinds(p + ".rowkind");
    cc.write_rowsrc(p + ".rowsrc");
    if (flags & PZK_COMPILE_EMIT_O1) cc.write_o1(p + ".O1.r1cs", p + ".O1.sym");
  } catch (std::exception& e) {
    set_err(err, err_len, e.what());
    return PZK_ECOMPILE;
  }
  return PZK_OK;
}

const char* pzk_last_error(const pzk_circuit* c) { return c ? c->err.c_str() : "null handle"; }
uint32_t pzk_witness_size(const pzk_circuit* c) { return c->h.n_wires; }
uint32_t pzk_input_size(const pzk_circuit* c) { return c->h.n_inputs; }
uint32_t pzk_public_size(const pzk_circuit* c) { return c->h.n_pub_out + c->h.n_pub_in; }
uint32_t pzk_constraint_count(const pzk_circuit* c) { return c->h.n_constraints; }
const char* pzk_circuit_meta_json(const pzk_circuit* c) { return c->meta.c_str(); }
uint64_t pzk_wtns_size(const pzk_circuit* c) { return 12 + 12 + 40 + 12 + 32ull * c->h.n_wires; }

int pzk_circuit_stats(const pzk_circuit* c, uint64_t* op_records, uint64_t* f_mul, uint64_t* f_inv,
                      uint64_t* rows, uint64_t* terms, uint64_t* bytes_per_lane) {
  if (op_records) *op_records = c->h.n_op_records;
  if (f_mul) *f_mul = c->h.stat_f_mul;
  if (f_inv) *f_inv = c->h.stat_f_inv;
  if (rows) *rows = c->h.n_rows;
  if (terms) *terms = c->h.n_terms;
  if (bytes_per_lane) *bytes_per_lane = c->bytes_per_lane;
  return PZK_OK;
}

static void free_tile(pzk_circuit* c) {
  if (c->d_U) cudaFree(c->d_U);
  if (c->d_F) cudaFree(c->d_F);
  c->d_U = c->d_F = nullptr;
  c->L = 0;
}
static void free_batch(pzk_circuit* c) {
  if (c->d_inputs) cudaFree(c->d_inputs);
  if (c->d_status) cudaFree(c->d_status);
  if (c->d_first_bad) cudaFree(c->d_first_bad);
  if (c->d_public) cudaFree(c->d_public);
  if (c->d_digest) cudaFree(c->d_digest);
  if (c->d_dig_state) cudaFree(c->d_dig_state);
  c->d_dig_state = nullptr;
  c->d_inputs = nullptr; c->d_status = nullptr; c->d_first_bad = nullptr; c->d_public = nullptr; c->d_digest = nullptr;
  c->batch_cap = 0; c->pub_cap = 0; c->d_inputs_bytes = 0;
}

void pzk_circuit_close(pzk_circuit* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  free_tile(c); free_batch(c);
  cudaFree(c->d_ops); cudaFree(c->d_fpool); cudaFree(c->d_coefs); cudaFree(c->d_coef_kind); cudaFree(c->d_coef_mag);
  cudaFree(c->d_dig_recs); cudaFree(c->d_dig_tab);
  cudaFree(c->d_list); cudaFree(c->d_in_table); cudaFree(c->d_rows); cudaFree(c->d_terms); cudaFree(c->d_exports); cudaFree(c->d_pub_entries);
  if (c->in_flight) pzk_sync(c);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
  delete c;
}

template <typename T>
static cudaError_t upload(T** dst, const void* src, size_t bytes) {
  cudaError_t e = cudaMalloc((void**)dst, bytes ? bytes : 16);
  if (e != cudaSuccess) return e;
  if (bytes) e = cudaMemcpy(*dst, src, bytes, cudaMemcpyHostToDevice);
  return e;
}

static void classify_coefs(const PzkCoef* coefs, uint32_t n, std::vector<unsigned char>& kind, std::vector<u64>& mag) {
  kind.assign(n ? n : 1, 0); mag.assign(n ? n : 1, 0);
  for (uint32_t i = 0; i < n; i++) {
    pzk::U256 v = pzk::U256::from_limbs(coefs[i].plain[0], coefs[i].plain[1], coefs[i].plain[2], coefs[i].plain[3]);
    if (v.fits64() && v.w[0] < (1ull << 63)) { kind[i] = 1; mag[i] = v.w[0]; continue; }
    pzk::U256 nv = pzk::sub(pzk::FR_P, v);
    if (nv.fits64() && nv.w[0] < (1ull << 63)) { kind[i] = 2; mag[i] = nv.w[0]; }
  }
}

// The digest program: the export entries of every segment compiled into digest records (pzk_kernels.cuh).
// All bit-field views of one word collapse into one table of per-bit 128-bit coefficients.  When the program carries
// digest descriptors (PZK_FLAG_DIG, written by the compiler behind the op that defines a wire or a word with views),
// the records of that value are moved INTO the descriptor - the evaluator folds the value when it is defined - and
// only the rest (outputs of the hint intrinsics, truth-table views over several words, wide views) stays in the
// per-segment list of digest_kernel.
static void build_digest_program(pzk_circuit* c, std::vector<DigRec>& recs, std::vector<ulonglong2>& tab) {
  typedef unsigned __int128 u128;
  c->dig_seg.assign(c->h.n_segments, {0, 0});
  c->dig_fused = false; c->dig_fused_entries = c->dig_kernel_entries = 0;
  PzkOp* ops = const_cast<PzkOp*>(c->ops);
  for (uint32_t s = 0; s < c->h.n_segments; s++) {
    const PzkSegment& sg = c->segs[s];
    c->dig_seg[s].first = recs.size();
    struct Work { u128 plain = 0; uint32_t plain_rep = 0; std::vector<u128> T; uint64_t n_entries = 0; };
    std::map<uint64_t, Work> work;  // (plane << 32 | slot) -> what the wires behind this value contribute
    std::vector<DigRec> generic;
    u128 konst = 0;
    for (uint64_t e = sg.exp_off; e < sg.exp_off + sg.n_exp; e++) {
      const PzkExport& x = c->exports[e];
      const uint32_t cw = pzk_digest_weight(x.wire);
      if (x.ref == PZK_REF_ZERO) continue;
      if (x.ref == PZK_REF_ONE) { konst += cw; continue; }
      DigRec g; g.type_nbits = DIG_GENERIC; g.slot = 0; g.a = (uint32_t)e; g.b = cw;
      if (x.ref == PZK_REF_TABVIEW) { generic.push_back(g); continue; }
      const uint32_t cls = PZK_REF_CLS(x.ref), slot = PZK_REF_SLOT(x.ref);
      if (cls == 3) {
        const uint32_t s_ = x.aux & 255u, n_ = (x.aux >> 8) & 255u, k_ = (x.aux >> 16) & 255u;
        const bool isn = (x.ref & PZK_REF_VIEW_N) != 0;
        const uint32_t width = isn ? 256 : 64;
        if (n_ + k_ <= 64 && s_ < width) {
          Work& w = work[((uint64_t)isn << 32) | slot];
          if (w.T.empty()) w.T.assign(width, 0);
          for (uint32_t b = s_; b < s_ + n_ && b < width; b++) w.T[b] += (u128)cw << (b - s_ + k_);
          w.n_entries++;
        } else if (n_ != 0 && s_ < width) generic.push_back(g);
        continue;
      }
      Work& w = work[((uint64_t)(cls == 2) << 32) | slot];
      w.plain += cw; w.n_entries++;
      w.plain_rep = cls == 0 ? 1u : cls == 1 ? 2u : ((x.ref & PZK_REF_Z) ? 4u : 3u);
    }
    // descriptors of this segment's op stream claim the work of the value they follow
    auto claim = [&](uint64_t at) {
      c->dig_fused = true;
      uint32_t wds[4];
      memcpy(wds, &ops[at], 16);
      const uint64_t key = ((uint64_t)(wds[1] != 0) << 32) | wds[2];
      uint32_t out[4] = {0, 0, 0, 0};
      auto it = work.find(key);
      if (it != work.end()) {
        Work& w = it->second;
        uint32_t nbits = 0;
        for (uint32_t b = 0; b < w.T.size(); b++) if (w.T[b]) nbits = b + 1;
        uint32_t repk = w.plain_rep;
        if (!repk) repk = wds[1] ? 5u : 1u;  // a word that is only the base of views: U word or plain N value
        if ((uint64_t)(w.plain >> 64)) { set_err(c, "digest: weight overflow"); }
        out[0] = repk | (nbits ? 16u : 0u) | (nbits << 8);
        out[1] = (uint32_t)(uint64_t)w.plain; out[2] = (uint32_t)((uint64_t)w.plain >> 32);
        out[3] = (uint32_t)tab.size();
        // nibble tables for the in-stream fold (dig_fold_table): 16 subset sums per four bits
        for (uint32_t g = 0; 4 * g < nbits; g++)
          for (uint32_t v = 0; v < 16; v++) {
            u128 sum = 0;
            for (uint32_t j = 0; j < 4; j++) if (((v >> j) & 1u) && 4 * g + j < nbits) sum += w.T[4 * g + j];
            tab.push_back(make_ulonglong2((u64)sum, (u64)(sum >> 64)));
          }
        if (tab.size() > 0xffffffffull) set_err(c, "digest: table pool overflow");
        c->dig_fused_entries += w.n_entries;
        work.erase(it);
      }
      memcpy(&ops[at], out, 16);
    };
    for (uint64_t pc = sg.op_off; pc < sg.op_off + sg.n_ops; pc++) {
      const PzkOp& o = ops[pc];
      const bool dig = (o.flags & PZK_FLAG_DIG) != 0;
      const bool dig2 = (o.opc == PZK_F_MULADD || o.opc == PZK_Z_MULADD) && (o.flags & PZK_FLAG_DIG2);
      if (o.opc == PZK_CHECK_INT || o.opc == PZK_CHECK_F || o.opc == PZK_CHECK_I64) { pc += o.b; continue; }
      if (o.flags & PZK_FLAG_EXT) pc++;
      if (o.opc == PZK_V_LUT && (o.flags & PZK_FLAG_W64)) pc++;
      if (dig2) claim(++pc);  // the product of a fused multiply-add that is a wire: its descriptor comes first
      if (dig) claim(++pc);
    }
    // what no descriptor claimed: records for digest_kernel
    for (auto& kv : work) {
      Work& w = kv.second;
      const bool fplane = (kv.first >> 32) != 0;
      c->dig_kernel_entries += w.n_entries;
      if (w.plain) {
        u128 pl = w.plain;
        while (pl) {  // 32-bit weights per record
          DigRec r; r.slot = (uint32_t)kv.first; r.b = 0;
          r.type_nbits = w.plain_rep == 1 ? DIG_PLAIN_U : w.plain_rep == 2 ? DIG_PLAIN_I : w.plain_rep == 4 ? DIG_PLAIN_Z : DIG_PLAIN_F;
          r.a = (uint32_t)std::min<u128>(pl, (u128)0xffffffffu);
          pl -= r.a;
          recs.push_back(r);
        }
      }
      uint32_t nbits = 0;
      for (uint32_t b = 0; b < w.T.size(); b++) if (w.T[b]) nbits = b + 1;
      if (nbits) {
        DigRec r; r.type_nbits = (fplane ? DIG_WORD_N : DIG_WORD_U) | (nbits << 8);
        r.slot = (uint32_t)kv.first; r.a = (uint32_t)tab.size(); r.b = 0;
        for (uint32_t b = 0; b < nbits; b++) tab.push_back(make_ulonglong2((u64)w.T[b], (u64)(w.T[b] >> 64)));
        recs.push_back(r);
      }
    }
    c->dig_kernel_entries += generic.size();
    recs.insert(recs.end(), generic.begin(), generic.end());
    while (konst) {  // constant wires (value 1): 64 bits at a time
      DigRec r; r.type_nbits = DIG_CONST; r.slot = 0; r.a = (uint32_t)konst; r.b = (uint32_t)(konst >> 32);
      recs.push_back(r);
      konst >>= 64;
      if (konst) { set_err(c, "digest: constant overflow"); break; }
    }
    c->dig_seg[s].second = recs.size() - c->dig_seg[s].first;
  }
}

// name -> witness index of an iden3 .sym file (lines "labelIdx,witnessIdx,componentIdx,name"; -1 = optimised away)
static int read_sym(const char* path, std::unordered_map<std::string, int64_t>& m, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return PZK_EIO; }
  std::vector<char> line(1 << 16);
  while (fgets(line.data(), (int)line.size(), f)) {
    char* p1 = strchr(line.data(), ',');
    char* p2 = p1 ? strchr(p1 + 1, ',') : nullptr;
    char* p3 = p2 ? strchr(p2 + 1, ',') : nullptr;
    if (!p3) continue;
    long long w = strtoll(p1 + 1, nullptr, 10);
    size_t n = strlen(p3 + 1);
    while (n && (p3[n] == '\n' || p3[n] == '\r')) n--;
    m.emplace(std::string(p3 + 1, n), (int64_t)w);
  }
  fclose(f);
  return PZK_OK;
}

// Re-number the program's wires after an EXTERNAL .sym (one written by the circom compiler for the same circuit at
// any optimisation level): wire <- the witness index the external file gives the signal of the same qualified
// name; signals it lists with -1, or not at all, are dropped; signals merged into one wire are written once.
// After this calculateWitness / calculateWTNSBin / the digest all use the external numbering.
static int remap_wires(pzk_circuit* c, const char* own_sym, const char* ext_sym) {
  std::unordered_map<std::string, int64_t> own, ext;
  std::string e;
  int rc = read_sym(own_sym, own, e);
  if (rc == PZK_OK) rc = read_sym(ext_sym, ext, e);
  if (rc) { set_err(c, e); return rc; }
  std::vector<int64_t> to(c->h.n_wires, -1);
  int64_t max_w = 0;
  for (auto& kv : ext) {
    if (kv.second < 0) continue;
    auto it = own.find(kv.first);
    if (it == own.end()) { set_err(c, "external .sym names `" + kv.first + "`, which is not a signal of this program"); return PZK_EFORMAT; }
    if (it->second <= 0 || it->second >= (int64_t)c->h.n_wires) { set_err(c, "program .sym does not belong to this program"); return PZK_EFORMAT; }
    to[it->second] = kv.second;
    if (kv.second > max_w) max_w = kv.second;
  }
  if (max_w >= (int64_t)c->h.n_wires) { set_err(c, "external .sym: witness index " + std::to_string(max_w) + " exceeds the program's " + std::to_string(c->h.n_wires) + " wires"); return PZK_EFORMAT; }
  const uint32_t n_new = (uint32_t)max_w + 1;
  std::vector<uint8_t> taken(n_new, 0);
  taken[0] = 1;
  PzkSegment* segs = const_cast<PzkSegment*>(c->segs);
  PzkExport* ex = const_cast<PzkExport*>(c->exports);
  uint64_t w = 0;
  for (uint32_t s = 0; s < c->h.n_segments; s++) {
    const uint64_t lo = segs[s].exp_off, hi = lo + segs[s].n_exp;
    segs[s].exp_off = w;
    for (uint64_t k = lo; k < hi; k++) {
      const int64_t nw = to[ex[k].wire];
      if (nw <= 0 || taken[nw]) continue;
      taken[nw] = 1;
      ex[w] = ex[k];
      ex[w].wire = (uint32_t)nw;
      w++;
    }
    segs[s].n_exp = w - segs[s].exp_off;
  }
  for (uint32_t i = 1; i < n_new; i++)
    if (!taken[i]) { set_err(c, "external .sym: witness index " + std::to_string(i) + " has no signal of this program behind it"); return PZK_EFORMAT; }
  c->h.n_exports = w;
  c->h.n_wires = n_new;
  return PZK_OK;
}

static int open_impl(const char* program_path, const char* own_sym, const char* ext_sym, int cuda_device, pzk_circuit** out);

int pzk_circuit_open(const char* program_path, int cuda_device, pzk_circuit** out) {
  return open_impl(program_path, nullptr, nullptr, cuda_device, out);
}
int pzk_circuit_open_ex(const char* program_path, const char* program_sym_path, const char* external_sym_path,
                        int cuda_device, pzk_circuit** out) {
  if (!program_sym_path || !external_sym_path) return PZK_EINVAL;
  return open_impl(program_path, program_sym_path, external_sym_path, cuda_device, out);
}

static int open_impl(const char* program_path, const char* own_sym, const char* ext_sym, int cuda_device, pzk_circuit** out) {
  if (!program_path || !out) return PZK_EINVAL;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return PZK_ENODEVICE;
  if (cuda_device < 0 || cuda_device >= ndev) return PZK_EINVAL;
  pzk_circuit* c = new pzk_circuit();
  *out = c;  // returned even on failure so that pzk_last_error works; caller closes it
  c->device = cuda_device;
  FILE* f = fopen(program_path, "rb");
  if (!f) { set_err(c, std::string("cannot open ") + program_path); return PZK_EIO; }
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  c->blob.resize((size_t)sz + 16);
  if (fread(c->blob.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); set_err(c, "short read"); return PZK_EIO; }
  fclose(f);
  if ((size_t)sz < sizeof(PzkHeader)) { set_err(c, "not a program file"); return PZK_EFORMAT; }
  memcpy(&c->h, c->blob.data(), sizeof c->h);
  if (c->h.magic != PZK_MAGIC || c->h.version != PZK_VERSION) { set_err(c, "bad program magic/version"); return PZK_EFORMAT; }
  const uint8_t* b = c->blob.data();
  uint64_t pos = al16(sizeof(PzkHeader));
  c->segs = (const PzkSegment*)(b + pos); pos = al16(pos + c->h.n_segments * sizeof(PzkSegment));
  c->ops = (const PzkOp*)(b + pos); pos = al16(pos + c->h.n_op_records * sizeof(PzkOp));
  c->fpool = (const uint64_t*)(b + pos); pos = al16(pos + (uint64_t)c->h.n_fpool * 32);
  c->coefs = (const PzkCoef*)(b + pos); pos = al16(pos + (uint64_t)c->h.n_coef * sizeof(PzkCoef));
  c->list = (const uint32_t*)(b + pos); pos = al16(pos + (uint64_t)c->h.n_list * 4);
  c->inputs = (const PzkInput*)(b + pos); pos = al16(pos + (uint64_t)c->h.n_inputs * sizeof(PzkInput));
  c->rows = (const PzkRow*)(b + pos); pos = al16(pos + c->h.n_rows * sizeof(PzkRow));
  c->terms = (const PzkTerm*)(b + pos); pos = al16(pos + c->h.n_terms * sizeof(PzkTerm));
  c->exports = (const PzkExport*)(b + pos); pos = al16(pos + c->h.n_exports * sizeof(PzkExport));
  if (pos + c->h.reserved[0] > (uint64_t)sz) { set_err(c, "truncated program file"); return PZK_EFORMAT; }
  c->meta.assign((const char*)(b + pos), c->h.reserved[0]);
  if (ext_sym) { int rrc = remap_wires(c, own_sym, ext_sym); if (rrc) return rrc; }
  {
    // packed input record: [1-byte inputs (declared <= 8 bits)] pad8 [8-byte inputs (<= 64 bits)] [32-byte field inputs]
    uint32_t n8 = 0, n64 = 0;
    for (uint32_t k = 0; k < c->h.n_inputs; k++) { uint32_t b_ = c->inputs[k].bits; if (b_ && b_ <= 8) n8++; else if (b_) n64++; }
    uint32_t off8 = 0, off64 = (n8 + 7) / 8 * 8, offf = (off64 + n64 * 8 + 15) / 16 * 16;  // field section 16-byte aligned
    c->in_table.resize(c->h.n_inputs ? c->h.n_inputs : 1);
    for (uint32_t k = 0; k < c->h.n_inputs; k++) {
      uint32_t b_ = c->inputs[k].bits;
      if (b_ && b_ <= 8) { c->in_table[k] = make_uint2(0, off8); off8 += 1; }
      else if (b_) { c->in_table[k] = make_uint2(1, off64); off64 += 8; }
      else { c->in_table[k] = make_uint2(2, offf); offf += 32; }
    }
    c->packed_stride = (offf + 15) / 16 * 16;
  }
  c->bytes_per_lane = (uint64_t)c->h.n_u_slots * 8 + (uint64_t)c->h.n_f_slots * 32;
  c->smem_bytes = ((size_t)c->h.reserved[1] + DIG_ACC_WORDS) * 8 * 128;  // operand cache + digest accumulators

  CK(cudaSetDevice(cuda_device));
  CK(cudaStreamCreate(&c->stream));
  CK(cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking));
  if (c->smem_bytes > 48 * 1024) CK(cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  {
    int per_sm = 0, n_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eval_kernel, 128, c->smem_bytes));
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cuda_device));
    c->wave_lanes = (uint64_t)per_sm * n_sm * 128;
  }
  {
    // the digest program first: it fills in the digest descriptors of the op stream before that is uploaded
    std::vector<DigRec> recs; std::vector<ulonglong2> tab;
    build_digest_program(c, recs, tab);
    CK(upload(&c->d_dig_recs, recs.data(), recs.size() * sizeof(DigRec)));
    CK(upload(&c->d_dig_tab, tab.data(), tab.size() * sizeof(ulonglong2)));
  }
  CK(upload(&c->d_ops, c->ops, c->h.n_op_records * sizeof(PzkOp)));
  CK(upload(&c->d_fpool, c->fpool, (size_t)c->h.n_fpool * 32));
  CK(upload(&c->d_coefs, c->coefs, (size_t)c->h.n_coef * sizeof(PzkCoef)));
  {
    std::vector<unsigned char> kind; std::vector<u64> mag;
    classify_coefs(c->coefs, c->h.n_coef, kind, mag);
    CK(upload(&c->d_coef_kind, kind.data(), kind.size()));
    CK(upload(&c->d_coef_mag, mag.data(), mag.size() * 8));
  }
  CK(upload(&c->d_list, c->list, (size_t)c->h.n_list * 4));
  CK(upload(&c->d_in_table, c->in_table.data(), c->in_table.size() * sizeof(uint2)));
  CK(upload(&c->d_rows, c->rows, c->h.n_rows * sizeof(PzkRow)));
  CK(upload(&c->d_terms, c->terms, c->h.n_terms * sizeof(PzkTerm)));
  CK(upload(&c->d_exports, c->exports, c->h.n_exports * sizeof(PzkExport)));
  // public wires per segment
  uint32_t n_pub = c->h.n_pub_out + c->h.n_pub_in;
  std::vector<PzkExport> all_pub;
  c->seg.resize(c->h.n_segments);
  for (uint32_t s = 0; s < c->h.n_segments; s++) {
    c->seg[s].s = c->segs[s];
    c->seg[s].pub_off = all_pub.size();
    for (uint64_t e = c->segs[s].exp_off; e < c->segs[s].exp_off + c->segs[s].n_exp; e++)
      if (c->exports[e].wire >= 1 && c->exports[e].wire <= n_pub) { c->seg[s].pub.push_back(c->exports[e]); all_pub.push_back(c->exports[e]); }
  }
  CK(upload(&c->d_pub_entries, all_pub.data(), all_pub.size() * sizeof(PzkExport)));
  return PZK_OK;
}

int pzk_set_tile_lanes(pzk_circuit* c, uint64_t lanes) {
  if (!c) return PZK_EINVAL;
  c->tile_lanes_cfg = lanes;
  cudaSetDevice(c->device);
  free_tile(c);
  return PZK_OK;
}
uint64_t pzk_get_tile_lanes(const pzk_circuit* c) { return c->L; }
uint64_t pzk_wave_lanes(const pzk_circuit* c) { return c->wave_lanes; }

static int ensure_tile(pzk_circuit* c, uint64_t want) {
  uint64_t L = c->tile_lanes_cfg;
  if (L == 0 && c->L && (c->L >= want || c->tile_auto_capped)) return PZK_OK;  // keep the resident tile
  if (L == 0) {
    size_t free_b = 0, total_b = 0;
    CK(cudaMemGetInfo(&free_b, &total_b));
    uint64_t budget = (uint64_t)(free_b * 0.80);
    L = budget / (c->bytes_per_lane ? c->bytes_per_lane : 1);
    // whole waves only: every segment is one launch, a partially filled last wave is pure tail
    uint64_t wave = c->wave_lanes ? c->wave_lanes : 148ull * 2048;
    uint64_t cap = 3 * wave;
    if (L > cap) L = cap;
    if (L > want) L = want;
    if (L >= wave) L = L / wave * wave;
    c->tile_auto_capped = (L < want);
  }
  if (L > want) L = want;
  L = (L + 127) / 128 * 128;
  if (L == 0) L = 128;
  if (c->L == L) return PZK_OK;
  free_tile(c);
  CK(cudaMalloc((void**)&c->d_U, std::max<uint64_t>((uint64_t)c->h.n_u_slots * 8 * L, 16)));
  CK(cudaMalloc((void**)&c->d_F, std::max<uint64_t>((uint64_t)c->h.n_f_slots * 32 * L, 16)));
  c->L = L;
  return PZK_OK;
}

static int ensure_batch(pzk_circuit* c, uint64_t batch, uint64_t input_bytes) {
  if (batch <= c->batch_cap && input_bytes <= c->d_inputs_bytes) return PZK_OK;
  free_batch(c);
  uint64_t n_pub = c->h.n_pub_out + c->h.n_pub_in;
  CK(cudaMalloc((void**)&c->d_inputs, std::max<uint64_t>(input_bytes, 16)));
  c->d_inputs_bytes = input_bytes;
  CK(cudaMalloc((void**)&c->d_status, batch * 4));
  CK(cudaMalloc((void**)&c->d_first_bad, batch * 8));
  CK(cudaMalloc((void**)&c->d_public, std::max<uint64_t>(batch * n_pub * 32, 16)));
  CK(cudaMalloc((void**)&c->d_digest, batch * 32));
  CK(cudaMalloc((void**)&c->d_dig_state, batch * 8 * DIG_STATE_PIECES));
  c->batch_cap = batch;
  return PZK_OK;
}

static void prof_begin(pzk_circuit* c, int which, cudaEvent_t& a, cudaEvent_t& b) {
  if (!c->prof) return;
  cudaEventCreate(&a); cudaEventCreate(&b);
  cudaEventRecord(a, c->stream);
  (void)which;
}
static void prof_end(pzk_circuit* c, int which, cudaEvent_t a, cudaEvent_t b) {
  if (!c->prof) return;
  cudaEventRecord(b, c->stream);
  c->pending.push_back({which, {a, b}});
}
static void prof_collect(pzk_circuit* c) {
  size_t evk = 0;
  if (c->seg_ms.size() != c->h.n_segments) c->seg_ms.assign(c->h.n_segments, 0.0);
  for (auto& p : c->pending) {
    float ms = 0;
    cudaEventSynchronize(p.second.second);
    cudaEventElapsedTime(&ms, p.second.first, p.second.second);
    c->prof_ms[p.first] += ms;
    c->prof_launches[p.first] += 1;
    if (p.first == 0 && evk < c->pending_seg.size()) c->seg_ms[c->pending_seg[evk++]] += ms;
    cudaEventDestroy(p.second.first); cudaEventDestroy(p.second.second);
  }
  c->pending.clear();
  c->pending_seg.clear();
}

// ---------------------------------------------------------------------------------------
// One pass of the hot path over the resident (or streamed) batch.
// ---------------------------------------------------------------------------------------
struct RunOpts {
  int check_rows = 1;
  // full witnesses of selected lanes (device buffer, canonical): AoS rows or the R1CS checker's blocked layout
  const uint64_t* export_lanes = nullptr;
  uint64_t n_export = 0;
  u64* d_witnesses = nullptr;
  bool blocked = false;
  // streamed call: packed host records in, results out, copies of tile k+1 / k-1 under the kernels of tile k
  const uint8_t* h_packed = nullptr;
  uint32_t* h_status = nullptr;
  int64_t* h_first_bad = nullptr;
  uint8_t* h_public = nullptr;
  uint64_t* h_digest = nullptr;
  bool async = false;  // return after enqueueing; pzk_sync() completes the call
};

static void release_run(pzk_circuit* c) {
  for (cudaEvent_t e : c->run_events) cudaEventDestroy(e);
  c->run_events.clear();
  if (c->d_lane_list) { cudaFree(c->d_lane_list); c->d_lane_list = nullptr; }
  c->in_flight = false;
}

int pzk_sync(pzk_circuit* c) {
  if (!c) return PZK_EINVAL;
  CK(cudaSetDevice(c->device));
  cudaError_t e1 = cudaStreamSynchronize(c->stream), e2 = cudaStreamSynchronize(c->copy_stream);
  cudaError_t e3 = cudaGetLastError();
  release_run(c);
  prof_collect(c);
  if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) {
    set_err(c, std::string("pzk_sync: ") + cudaGetErrorString(e1 != cudaSuccess ? e1 : e2 != cudaSuccess ? e2 : e3));
    return PZK_ECUDA;
  }
  return PZK_OK;
}

static int run_batch(pzk_circuit* c, const RunOpts& o) {
  CK(cudaSetDevice(c->device));
  if (c->in_flight) { int rc0 = pzk_sync(c); if (rc0) return rc0; }
  const bool streamed = o.h_packed != nullptr;
  // a streamed call works in tiles of one wave so that the copies of its neighbours hide under a tile's kernels
  uint64_t want = c->batch;
  if (streamed && c->wave_lanes && want > c->wave_lanes && c->tile_lanes_cfg == 0) want = c->wave_lanes;
  int rc = ensure_tile(c, want);
  if (rc) return rc;
  const uint64_t L = std::min<uint64_t>(c->L, (want + 127) / 128 * 128);
  const uint32_t n_pub = c->h.n_pub_out + c->h.n_pub_in;
  const bool digest = c->digest_on;
  const bool fused = digest && c->dig_fused;   // values are folded when they are defined: nothing extra to store
  const int store_all = (o.n_export > 0 || (digest && !fused)) ? 1 : 0;
  const uint64_t n_tiles = (c->batch + L - 1) / L;
  c->in_flight = true;
  auto new_event = [&]() { cudaEvent_t e; cudaEventCreateWithFlags(&e, cudaEventDisableTiming); c->run_events.push_back(e); return e; };
  std::vector<cudaEvent_t> ev_in(streamed ? n_tiles : 0);
  if (streamed) {
    for (uint64_t t = 0; t < n_tiles; t++) {
      const uint64_t base = t * L, n = std::min<uint64_t>(L, c->batch - base);
      CK(cudaMemcpyAsync(reinterpret_cast<unsigned char*>(c->d_inputs) + base * c->packed_stride, o.h_packed + base * c->packed_stride,
                         n * (uint64_t)c->packed_stride, cudaMemcpyHostToDevice, c->copy_stream));
      ev_in[t] = new_event();
      CK(cudaEventRecord(ev_in[t], c->copy_stream));
    }
  }
  CK(cudaMemsetAsync(c->d_status, 0, c->batch * 4, c->stream));
  CK(cudaMemsetAsync(c->d_first_bad, 0xff, c->batch * 8, c->stream));
  if (digest) CK(cudaMemsetAsync(c->d_dig_state, 0, c->batch_cap * 8 * DIG_STATE_PIECES, c->stream));
  // export lanes sorted into tiles: [lane in tile..., output row...] per tile, one upload
  std::vector<std::vector<u64>> tl(o.n_export ? n_tiles : 0), tr(o.n_export ? n_tiles : 0);
  std::vector<uint64_t> tl_off(n_tiles + 1, 0);
  if (o.n_export) {
    for (uint64_t j = 0; j < o.n_export; j++) { tl[o.export_lanes[j] / L].push_back(o.export_lanes[j] % L); tr[o.export_lanes[j] / L].push_back(j); }
    std::vector<u64> flat;
    for (uint64_t t = 0; t < n_tiles; t++) {
      tl_off[t] = flat.size();
      flat.insert(flat.end(), tl[t].begin(), tl[t].end());
      flat.insert(flat.end(), tr[t].begin(), tr[t].end());
    }
    tl_off[n_tiles] = flat.size();
    CK(cudaMalloc((void**)&c->d_lane_list, flat.size() * 8));
    CK(cudaMemcpyAsync(c->d_lane_list, flat.data(), flat.size() * 8, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));  // `flat` is a local
  }
  cudaEvent_t ra, rb;
  prof_begin(c, 3, ra, rb);
  for (uint64_t t = 0; t < n_tiles; t++) {
    const uint64_t base = t * L, n = std::min<uint64_t>(L, c->batch - base);
    const unsigned grid = (unsigned)((n + 127) / 128);
    if (streamed) CK(cudaStreamWaitEvent(c->stream, ev_in[t], 0));
    const uint64_t n_tl = o.n_export ? tl[t].size() : 0;
    for (uint32_t s = 0; s < c->h.n_segments; s++) {
      const PzkSegment& sg = c->segs[s];
      cudaEvent_t ea, eb;
      uint64_t skip = 0;
      if (sg.n_ops >= 1 && c->ops[sg.op_off].opc == PZK_BJJ_MUL8) {
        // the BabyJubjub ladder is the first record of its segment: the dedicated kernel (pzk_kernels.cuh), then the
        // rest of the segment (the rows the ladder completes) in the evaluator
        BjjParams bp;
        bp.list = c->d_list + c->ops[sg.op_off].a; bp.fpool = c->d_fpool; bp.F = c->d_F; bp.n_lanes = n;
        bp.n_f_slots = c->h.n_f_slots; bp.status = c->d_status + base;
        prof_begin(c, 0, ea, eb);
        bjj_kernel<<<grid, 128, 0, c->stream>>>(bp);
        prof_end(c, 0, ea, eb);
        if (c->prof) c->pending_seg.push_back((int)s);
        skip = 1;
      }
      if (sg.n_ops > skip) {
        EvalParams p;
        p.ops = c->d_ops + sg.op_off + skip; p.n_rec = sg.n_ops - skip; p.U = c->d_U; p.F = c->d_F; p.L = L; p.n_lanes = n;
        p.fpool = c->d_fpool; p.list = c->d_list;
        if (c->packed) {
          p.inputs = reinterpret_cast<const u64*>(reinterpret_cast<const unsigned char*>(c->d_inputs) + base * c->packed_stride);
          p.in_table = c->d_in_table; p.in_stride = c->packed_stride;
        } else { p.inputs = c->d_inputs + base * c->h.n_inputs * 4; p.in_table = nullptr; p.in_stride = 0; }
        p.n_inputs = c->h.n_inputs; p.status = c->d_status + base; p.n_u_slots = c->h.n_u_slots; p.n_f_slots = c->h.n_f_slots;
        p.check_rows = o.check_rows; p.store_all = store_all; p.sc.coefs = c->d_coefs; p.sc.coef_kind = c->d_coef_kind; p.sc.coef_mag = c->d_coef_mag;
        p.first_bad = c->d_first_bad + base;
        p.digest = fused ? 1 : 0; p.dig_tab = c->d_dig_tab; p.dig_state = c->d_dig_state; p.dig_stride = c->batch_cap;
        p.dig_lane_base = base; p.dig_smem_off = (u32)((size_t)c->h.reserved[1] * 8 * 128);
        // the digest accumulators only take shared memory when they are used: the rest stays L1
        const size_t smem = fused ? c->smem_bytes : (size_t)c->h.reserved[1] * 8 * 128;
        prof_begin(c, 0, ea, eb);
        eval_kernel<<<grid, 128, smem, c->stream>>>(p);
        prof_end(c, 0, ea, eb);
        if (c->prof) c->pending_seg.push_back((int)s);
      }
      ExportParams p;
      p.list = c->d_list; p.n_u_slots = c->h.n_u_slots; p.n_f_slots = c->h.n_f_slots;
      p.U = c->d_U; p.F = c->d_F; p.L = L; p.lanes = nullptr; p.rows = nullptr; p.blocked = 0;
      if (!c->seg[s].pub.empty()) {
        p.entries = c->d_pub_entries + c->seg[s].pub_off; p.n_entries = c->seg[s].pub.size();
        p.lane_base = base; p.n_rows = n; p.out = c->d_public; p.out_wires = n_pub; p.wire_off = 1;
        prof_begin(c, 2, ea, eb);
        export_kernel<<<dim3(grid, 1), 128, 0, c->stream>>>(p);
        prof_end(c, 2, ea, eb);
      }
      if (digest && c->dig_seg[s].second) {
        // every wire defined in this segment, folded while its slot still holds it
        DigestParams dp;
        dp.recs = c->d_dig_recs + c->dig_seg[s].first; dp.n_recs = c->dig_seg[s].second; dp.tab = c->d_dig_tab;
        dp.exports = c->d_exports; dp.ex = p; dp.n_lanes = n; dp.state = c->d_dig_state; dp.state_stride = c->batch_cap; dp.lane_base = base;
        unsigned gy = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, (1184 + grid - 1) / grid), std::max<uint64_t>(1, dp.n_recs / 64));
        prof_begin(c, 4, ea, eb);
        digest_kernel<<<dim3(grid, gy), 128, 0, c->stream>>>(dp);
        prof_end(c, 4, ea, eb);
      }
      if (n_tl && sg.n_exp) {
        // full witnesses of the selected lanes of this tile: ONE launch per segment
        p.entries = c->d_exports + sg.exp_off; p.n_entries = sg.n_exp;
        p.lanes = c->d_lane_list + tl_off[t]; p.rows = p.lanes + n_tl; p.lane_base = 0; p.n_rows = n_tl;
        p.out = o.d_witnesses; p.out_wires = c->h.n_wires; p.wire_off = 0; p.blocked = o.blocked ? 1 : 0;
        prof_begin(c, 2, ea, eb);
        if (n_tl >= 64) {
          unsigned gx = (unsigned)((n_tl + 127) / 128);
          unsigned gy = (unsigned)std::min<uint64_t>(std::max<uint64_t>(1, 2368 / gx), std::max<uint64_t>(1, sg.n_exp / 64));
          export_kernel<<<dim3(gx, gy), 128, 0, c->stream>>>(p);
        } else {
          unsigned gx = (unsigned)std::min<uint64_t>(std::max<uint64_t>(sg.n_exp / 128, 1), 1024);
          export_rows_kernel<<<dim3(gx, (unsigned)n_tl), 128, 0, c->stream>>>(p);
        }
        prof_end(c, 2, ea, eb);
      }
    }
    if (digest) digest_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(c->d_dig_state, c->batch_cap, base, n, c->d_digest);
    if (streamed) {
      cudaEvent_t done = new_event();
      CK(cudaEventRecord(done, c->stream));
      CK(cudaStreamWaitEvent(c->copy_stream, done, 0));
      if (o.h_status) CK(cudaMemcpyAsync(o.h_status + base, c->d_status + base, n * 4, cudaMemcpyDeviceToHost, c->copy_stream));
      if (o.h_first_bad) CK(cudaMemcpyAsync(o.h_first_bad + base, c->d_first_bad + base, n * 8, cudaMemcpyDeviceToHost, c->copy_stream));
      if (o.h_public) CK(cudaMemcpyAsync(o.h_public + base * n_pub * 32, c->d_public + base * n_pub * 4, n * n_pub * 32, cudaMemcpyDeviceToHost, c->copy_stream));
      if (o.h_digest && digest) CK(cudaMemcpyAsync(o.h_digest + base * 4, c->d_digest + base * 4, n * 32, cudaMemcpyDeviceToHost, c->copy_stream));
    }
  }
  prof_end(c, 3, ra, rb);
  if (o.async) return PZK_OK;
  return pzk_sync(c);
}

int pzk_batch_upload(pzk_circuit* c, const uint8_t* inputs_le32, uint64_t batch) {
  if (!c || !inputs_le32 || batch == 0) return PZK_EINVAL;
  CK(cudaSetDevice(c->device));
  if (c->in_flight) { int rc0 = pzk_sync(c); if (rc0) return rc0; }
  int rc = ensure_batch(c, batch, batch * c->h.n_inputs * 32);
  if (rc) return rc;
  c->batch = batch; c->packed = false;
  CK(cudaMemcpyAsync(c->d_inputs, inputs_le32, batch * c->h.n_inputs * 32, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PZK_OK;
}

uint32_t pzk_packed_stride(const pzk_circuit* c) { return c->packed_stride; }
int pzk_packed_layout(const pzk_circuit* c, uint32_t* kind, uint32_t* offset) {
  if (!c || !kind || !offset) return PZK_EINVAL;
  for (uint32_t k = 0; k < c->h.n_inputs; k++) { kind[k] = c->in_table[k].x; offset[k] = c->in_table[k].y; }
  return PZK_OK;
}
int pzk_batch_upload_packed(pzk_circuit* c, const uint8_t* packed, uint64_t batch) {
  if (!c || !packed || batch == 0) return PZK_EINVAL;
  CK(cudaSetDevice(c->device));
  if (c->in_flight) { int rc0 = pzk_sync(c); if (rc0) return rc0; }
  int rc = ensure_batch(c, batch, batch * (uint64_t)c->packed_stride);
  if (rc) return rc;
  c->batch = batch; c->packed = true;
  CK(cudaMemcpyAsync(c->d_inputs, packed, batch * (uint64_t)c->packed_stride, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PZK_OK;
}

static int packed_call(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status, int64_t* first_bad,
                       uint8_t* public_le32, uint64_t* digest, bool async) {
  if (!c || !packed || batch == 0) return PZK_EINVAL;
  if (digest && !c->digest_on) { set_err(c, "digest requested but pzk_batch_set_digest(c, 1) was not called"); return PZK_EINVAL; }
  CK(cudaSetDevice(c->device));
  if (c->in_flight) { int rc0 = pzk_sync(c); if (rc0) return rc0; }
  int rc = ensure_batch(c, batch, batch * (uint64_t)c->packed_stride);
  if (rc) return rc;
  c->batch = batch; c->packed = true;
  RunOpts o;
  o.h_packed = packed; o.h_status = status; o.h_first_bad = first_bad; o.h_public = public_le32; o.h_digest = digest; o.async = async;
  return run_batch(c, o);
}

int pzk_witness_batch_packed(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status,
                             int64_t* first_bad, uint8_t* public_le32) {
  return packed_call(c, packed, batch, status, first_bad, public_le32, nullptr, false);
}

int pzk_witness_batch_packed_async(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status,
                                   int64_t* first_bad, uint8_t* public_le32, uint64_t* digest) {
  return packed_call(c, packed, batch, status, first_bad, public_le32, digest, true);
}

int pzk_witness_batch_packed_digest(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status,
                                    int64_t* first_bad, uint8_t* public_le32, uint64_t* digest) {
  return packed_call(c, packed, batch, status, first_bad, public_le32, digest, false);
}

// One host thread, several devices: the batch is cut into contiguous shares (share i = lanes
// [i * batch / n, (i + 1) * batch / n)), every device runs its share asynchronously, then all are
// joined.  No collective: the shares are independent.
int pzk_witness_batch_packed_multi(pzk_circuit* const* handles, int n_handles, const uint8_t* packed, uint64_t batch,
                                   uint32_t* status, int64_t* first_bad, uint8_t* public_le32, uint64_t* digest) {
  if (!handles || n_handles <= 0 || !packed || batch == 0) return PZK_EINVAL;
  int rc = PZK_OK;
  int launched = 0;
  for (int i = 0; i < n_handles; i++) {
    pzk_circuit* c = handles[i];
    if (!c || c->packed_stride != handles[0]->packed_stride || c->h.n_wires != handles[0]->h.n_wires) { rc = PZK_EINVAL; break; }
    const uint64_t lo = batch * (uint64_t)i / n_handles, hi = batch * (uint64_t)(i + 1) / n_handles;
    if (hi == lo) { launched++; continue; }
    const uint64_t n_pub = c->h.n_pub_out + c->h.n_pub_in;
    rc = packed_call(c, packed + lo * c->packed_stride, hi - lo, status ? status + lo : nullptr,
                     first_bad ? first_bad + lo : nullptr, public_le32 ? public_le32 + lo * n_pub * 32 : nullptr,
                     digest ? digest + lo * 4 : nullptr, true);
    if (rc) break;
    launched++;
  }
  for (int i = 0; i < launched && i < n_handles; i++) {
    int r2 = pzk_sync(handles[i]);
    if (rc == PZK_OK) rc = r2;
  }
  return rc;
}

int pzk_batch_set_digest(pzk_circuit* c, int on) {
  if (!c) return PZK_EINVAL;
  c->digest_on = on != 0;
  return PZK_OK;
}
uint32_t pzk_digest_weight_of(uint32_t wire) { return pzk_digest_weight(wire); }
int pzk_batch_download_digest(pzk_circuit* c, uint64_t* digest) {
  if (!c || c->batch == 0 || !digest) return PZK_EINVAL;
  if (!c->digest_on) { set_err(c, "pzk_batch_set_digest(c, 1) was not called before the run"); return PZK_EINVAL; }
  CK(cudaSetDevice(c->device));
  if (c->in_flight) { int rc0 = pzk_sync(c); if (rc0) return rc0; }
  CK(cudaMemcpy(digest, c->d_digest, c->batch * 32, cudaMemcpyDeviceToHost));
  return PZK_OK;
}

int pzk_batch_run(pzk_circuit* c, int check_rows) {
  if (!c || c->batch == 0) return PZK_EINVAL;
  RunOpts o; o.check_rows = check_rows;
  return run_batch(c, o);
}

int pzk_batch_download(pzk_circuit* c, uint32_t* status, int64_t* first_bad, uint8_t* public_le32) {
  if (!c || c->batch == 0) return PZK_EINVAL;
  CK(cudaSetDevice(c->device));
  if (c->in_flight) { int rc0 = pzk_sync(c); if (rc0) return rc0; }
  uint64_t n_pub = c->h.n_pub_out + c->h.n_pub_in;
  if (status) CK(cudaMemcpyAsync(status, c->d_status, c->batch * 4, cudaMemcpyDeviceToHost, c->stream));
  if (first_bad) CK(cudaMemcpyAsync(first_bad, c->d_first_bad, c->batch * 8, cudaMemcpyDeviceToHost, c->stream));
  if (public_le32) CK(cudaMemcpyAsync(public_le32, c->d_public, c->batch * n_pub * 32, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PZK_OK;
}

int pzk_witness_batch(pzk_circuit* c, const uint8_t* inputs_le32, uint64_t batch, uint32_t* status,
                      int64_t* first_bad, uint8_t* public_le32, const uint64_t* export_lanes,
                      uint64_t n_export, uint8_t* witnesses_le32) {
  if (!c || !inputs_le32 || batch == 0) return PZK_EINVAL;
  if (n_export && (!export_lanes || !witnesses_le32)) return PZK_EINVAL;
  for (uint64_t j = 0; j < n_export; j++) if (export_lanes[j] >= batch) return PZK_EINVAL;
  int rc = pzk_batch_upload(c, inputs_le32, batch);
  if (rc) return rc;
  u64* d_wit = nullptr;
  if (n_export) {
    CK(cudaMalloc((void**)&d_wit, n_export * c->h.n_wires * 32));
    CK(cudaMemsetAsync(d_wit, 0, n_export * c->h.n_wires * 32, c->stream));
  }
  RunOpts o; o.export_lanes = export_lanes; o.n_export = n_export; o.d_witnesses = d_wit;
  rc = run_batch(c, o);
  if (rc == PZK_OK && n_export) {
    cudaError_t e = cudaMemcpy(witnesses_le32, d_wit, n_export * c->h.n_wires * 32, cudaMemcpyDeviceToHost);
    if (e != cudaSuccess) { set_err(c, cudaGetErrorString(e)); rc = PZK_ECUDA; }
    for (uint64_t j = 0; j < n_export && rc == PZK_OK; j++) {  // wire 0 = 1
      memset(witnesses_le32 + j * c->h.n_wires * 32, 0, 32);
      witnesses_le32[j * c->h.n_wires * 32] = 1;
    }
  }
  if (d_wit) cudaFree(d_wit);
  if (rc) return rc;
  return pzk_batch_download(c, status, first_bad, public_le32);
}

int pzk_calculate_witness(pzk_circuit* c, const uint8_t* inputs_le32, uint8_t* witness_le32, uint32_t* status,
                          int64_t* first_bad) {
  if (!c || !inputs_le32 || !witness_le32) return PZK_EINVAL;
  uint64_t lane = 0;
  uint32_t st = 0; int64_t fb = -1;
  int rc = pzk_witness_batch(c, inputs_le32, 1, &st, &fb, nullptr, &lane, 1, witness_le32);
  if (status) *status = st;
  if (first_bad) *first_bad = fb;
  return rc;
}

static void put_u32(uint8_t*& p, uint32_t v) { memcpy(p, &v, 4); p += 4; }
static void put_u64(uint8_t*& p, uint64_t v) { memcpy(p, &v, 8); p += 8; }

int pzk_calculate_wtns_bin(pzk_circuit* c, const uint8_t* inputs_le32, uint8_t* out, uint32_t* status,
                           int64_t* first_bad) {
  if (!c || !inputs_le32 || !out) return PZK_EINVAL;
  uint8_t* p = out;
  memcpy(p, "wtns", 4); p += 4;
  put_u32(p, 2); put_u32(p, 2);
  put_u32(p, 1); put_u64(p, 40);
  put_u32(p, 32); memcpy(p, pzk::FR_P.w, 32); p += 32;
  put_u32(p, c->h.n_wires);
  put_u32(p, 2); put_u64(p, 32ull * c->h.n_wires);
  return pzk_calculate_witness(c, inputs_le32, p, status, first_bad);
}

int pzk_profile_get(pzk_circuit* c, int which, double* ms, uint64_t* launches) {
  if (!c || which < 0 || which > 4) return PZK_EINVAL;
  if (ms) *ms = c->prof_ms[which];
  if (launches) *launches = c->prof_launches[which];
  return PZK_OK;
}
int pzk_profile_segments(pzk_circuit* c, double* ms, uint32_t n) {
  if (!c || !ms) return PZK_EINVAL;
  for (uint32_t i = 0; i < n; i++) ms[i] = i < c->seg_ms.size() ? c->seg_ms[i] : 0.0;
  return (int)c->seg_ms.size();
}
void pzk_profile_reset(pzk_circuit* c) {
  c->seg_ms.assign(c->seg_ms.size(), 0.0); for (int i = 0; i < 5; i++) { c->prof_ms[i] = 0; c->prof_launches[i] = 0; } }
void pzk_profile_enable(pzk_circuit* c, int on) { c->prof = on != 0; }

// ---------------------------------------------------------------------------------------
// Generic `wtns check`: any iden3 .r1cs against explicit witnesses.  The matrices are parsed and uploaded ONCE
// per pzk_r1cs handle (the role of snarkjs' readR1cs); witnesses are checked in canonical form, either from host
// memory (pzk_r1cs_check) or handed over on the device by the evaluator (pzk_r1cs_check_circuit).
// ---------------------------------------------------------------------------------------
struct pzk_r1cs {
  int device = 0;
  uint32_t n_wires = 0, n_constraints = 0, n_pub_out = 0, n_pub_in = 0, n_prv_in = 0;
  uint64_t n_terms = 0;
  std::vector<PzkRow> rows;
  std::vector<PzkTerm> terms;
  std::vector<PzkCoef> coefs;
  std::vector<R1csTile> tiles;
  PzkRow* d_rows = nullptr; PzkTerm* d_terms = nullptr; PzkCoef* d_coefs = nullptr;
  unsigned char* d_kind = nullptr; u64* d_mag = nullptr; R1csTile* d_tiles = nullptr;
  // witness planes of the last call are kept and reused while they are large enough
  u64* d_W = nullptr; uint64_t W_lanes = 0;
  u32* d_status = nullptr; unsigned long long* d_bad = nullptr; uint64_t st_cap = 0;
};

struct CoefKey {
  uint64_t w[4];
  bool operator==(const CoefKey& o) const { return w[0] == o.w[0] && w[1] == o.w[1] && w[2] == o.w[2] && w[3] == o.w[3]; }
};
struct CoefKeyHash {
  size_t operator()(const CoefKey& k) const { return (size_t)(k.w[0] * 0x9e3779b97f4a7c15ull ^ k.w[1] * 0xc2b2ae3d27d4eb4full ^ k.w[2] * 0x165667b19e3779f9ull ^ k.w[3]); }
};

// every length in the file is checked against the bytes that are really there (a truncated or malformed
// .r1cs is PZK_EFORMAT, never an out-of-bounds read)
static int parse_r1cs(const char* path, pzk_r1cs& r, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return PZK_EIO; }
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  if (sz < 0) { fclose(f); err = "cannot size file"; return PZK_EIO; }
  std::vector<uint8_t> buf((size_t)sz);
  if (sz && fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); err = "short read"; return PZK_EIO; }
  fclose(f);
  const uint64_t size = (uint64_t)sz;
  if (size < 12 || memcmp(buf.data(), "r1cs", 4) != 0) { err = "not an r1cs file"; return PZK_EFORMAT; }
  uint32_t nsec; memcpy(&nsec, buf.data() + 8, 4);
  uint64_t pos = 12;
  const uint8_t* hdr = nullptr; uint64_t hdr_len = 0; const uint8_t* cons = nullptr; uint64_t cons_len = 0;
  for (uint32_t s = 0; s < nsec; s++) {
    if (pos + 12 > size) { err = "r1cs: truncated section table"; return PZK_EFORMAT; }
    uint32_t type; uint64_t len; memcpy(&type, buf.data() + pos, 4); memcpy(&len, buf.data() + pos + 4, 8);
    pos += 12;
    if (len > size - pos) { err = "r1cs: section " + std::to_string(type) + " runs past the end of the file"; return PZK_EFORMAT; }
    if (type == 1) { hdr = buf.data() + pos; hdr_len = len; }
    if (type == 2) { cons = buf.data() + pos; cons_len = len; }
    pos += len;
  }
  if (!hdr || !cons) { err = "r1cs: missing header or constraint section"; return PZK_EFORMAT; }
  if (hdr_len < 4) { err = "r1cs: header section too short"; return PZK_EFORMAT; }
  uint32_t n8; memcpy(&n8, hdr, 4);
  if (n8 != 32 || hdr_len < 4 + 32 + 28 || memcmp(hdr + 4, pzk::FR_P.w, 32) != 0) { err = "r1cs: prime is not the BN254 scalar field"; return PZK_EFORMAT; }
  memcpy(&r.n_wires, hdr + 36, 4);
  memcpy(&r.n_pub_out, hdr + 40, 4); memcpy(&r.n_pub_in, hdr + 44, 4); memcpy(&r.n_prv_in, hdr + 48, 4);
  memcpy(&r.n_constraints, hdr + 36 + 16 + 8, 4);
  if (r.n_wires == 0) { err = "r1cs: no wires"; return PZK_EFORMAT; }
  if ((uint64_t)r.n_constraints * 12 > cons_len) { err = "r1cs: constraint section shorter than its constraint count"; return PZK_EFORMAT; }
  std::unordered_map<CoefKey, uint32_t, CoefKeyHash> cidx;
  r.rows.reserve(r.n_constraints);
  r.terms.reserve(cons_len / 36);
  const uint8_t* q = cons; const uint8_t* end = cons + cons_len;
  for (uint32_t i = 0; i < r.n_constraints; i++) {
    PzkRow row; row.term_off = (uint32_t)r.terms.size(); row.kind = 0; row.index = i;
    uint16_t cnt[3];
    for (int part = 0; part < 3; part++) {
      if (end - q < 4) { err = "r1cs: truncated"; return PZK_EFORMAT; }
      uint32_t n; memcpy(&n, q, 4); q += 4;
      if (n > 65535) { err = "r1cs: linear combination too long"; return PZK_EFORMAT; }
      if ((uint64_t)(end - q) < (uint64_t)n * 36) { err = "r1cs: truncated"; return PZK_EFORMAT; }
      cnt[part] = (uint16_t)n;
      for (uint32_t k = 0; k < n; k++) {
        uint32_t wire; memcpy(&wire, q, 4);
        CoefKey key; memcpy(key.w, q + 4, 32);
        q += 36;
        if (wire >= r.n_wires) { err = "r1cs: wire index out of range"; return PZK_EFORMAT; }
        auto it = cidx.find(key);
        uint32_t ci;
        if (it == cidx.end()) {
          ci = (uint32_t)r.coefs.size(); cidx.emplace(key, ci);
          PzkCoef pc;
          pzk::U256 v = pzk::fr_reduce(pzk::U256::from_limbs(key.w[0], key.w[1], key.w[2], key.w[3]));
          pzk::U256 m1 = pzk::fr_to_mont(v), m2 = pzk::fr_to_mont(m1);
          memcpy(pc.plain, v.w, 32); memcpy(pc.mont, m1.w, 32); memcpy(pc.mont2, m2.w, 32);
          r.coefs.push_back(pc);
        } else ci = it->second;
        PzkTerm t; t.ref = (2u << 30) | wire; t.coef = ci;
        r.terms.push_back(t);
      }
    }
    row.na = cnt[0]; row.nb = cnt[1]; row.nc = cnt[2];
    r.rows.push_back(row);
  }
  r.n_terms = r.terms.size();
  if (r.n_terms >= (1ull << 32)) { err = "r1cs: too many terms"; return PZK_EFORMAT; }
  // tiles of the A/B/C stream: at most R1CS_TILE_ROWS rows and R1CS_TILE_TERMS terms, even first term
  size_t i = 0, n = r.rows.size();
  while (i < n) {
    R1csTile t; t.row0 = (uint32_t)i; t.term0 = r.rows[i].term_off & ~1u;
    uint32_t end_term = r.rows[i].term_off;
    uint32_t nr = 0;
    while (i < n && nr < R1CS_TILE_ROWS) {
      uint32_t nt = r.rows[i].na + r.rows[i].nb + r.rows[i].nc;
      if (nt + 2 > R1CS_TILE_TERMS) { err = "r1cs: a constraint has too many terms for the streaming tile"; return PZK_EFORMAT; }
      if (r.rows[i].term_off + nt - t.term0 + 1 > R1CS_TILE_TERMS) break;
      end_term = r.rows[i].term_off + nt; nr++; i++;
    }
    t.n_rows = nr; t.n_terms = end_term - t.term0;
    r.tiles.push_back(t);
  }
  return PZK_OK;
}

#define CKE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(err, err_len, std::string(#call) + ": " + cudaGetErrorString(e_)); return PZK_ECUDA; } } while (0)

void pzk_r1cs_close(pzk_r1cs* r) {
  if (!r) return;
  cudaSetDevice(r->device);
  cudaFree(r->d_rows); cudaFree(r->d_terms); cudaFree(r->d_coefs); cudaFree(r->d_kind); cudaFree(r->d_mag); cudaFree(r->d_tiles);
  cudaFree(r->d_W); cudaFree(r->d_status); cudaFree(r->d_bad);
  delete r;
}

int pzk_r1cs_open(const char* r1cs_path, int cuda_device, pzk_r1cs** out, char* err, size_t err_len) {
  if (!r1cs_path || !out) return PZK_EINVAL;
  *out = nullptr;
  pzk_r1cs* r = new pzk_r1cs();
  r->device = cuda_device;
  std::string e;
  int rc = parse_r1cs(r1cs_path, *r, e);   // format errors are reported before any device work
  if (rc) { set_err(err, err_len, e); delete r; return rc; }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err(err, err_len, "no CUDA device"); delete r; return PZK_ENODEVICE; }
  if (cuda_device < 0 || cuda_device >= ndev) { delete r; return PZK_EINVAL; }
  auto fail = [&](cudaError_t ce, const char* what) { set_err(err, err_len, std::string(what) + ": " + cudaGetErrorString(ce)); pzk_r1cs_close(r); return PZK_ECUDA; };
  cudaError_t ce;
  if ((ce = cudaSetDevice(cuda_device)) != cudaSuccess) return fail(ce, "cudaSetDevice");
  std::vector<unsigned char> kind; std::vector<u64> mag;
  classify_coefs(r->coefs.data(), (uint32_t)r->coefs.size(), kind, mag);
  std::vector<PzkTerm> terms_padded = r->terms;
  terms_padded.resize(r->terms.size() + 4, PzkTerm{2u << 30, 0});  // padding terms read wire 0
  if ((ce = upload(&r->d_rows, r->rows.data(), r->rows.size() * sizeof(PzkRow))) != cudaSuccess) return fail(ce, "upload rows");
  if ((ce = upload(&r->d_terms, terms_padded.data(), terms_padded.size() * sizeof(PzkTerm))) != cudaSuccess) return fail(ce, "upload terms");
  if ((ce = upload(&r->d_coefs, r->coefs.data(), r->coefs.size() * sizeof(PzkCoef))) != cudaSuccess) return fail(ce, "upload coefficients");
  if ((ce = upload(&r->d_kind, kind.data(), kind.size())) != cudaSuccess) return fail(ce, "upload coefficient kinds");
  if ((ce = upload(&r->d_mag, mag.data(), mag.size() * 8)) != cudaSuccess) return fail(ce, "upload coefficient magnitudes");
  if ((ce = upload(&r->d_tiles, r->tiles.data(), r->tiles.size() * sizeof(R1csTile))) != cudaSuccess) return fail(ce, "upload tiles");
  if ((ce = cudaFuncSetAttribute(r1cs_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R1CS_SMEM_BYTES)) != cudaSuccess) return fail(ce, "cudaFuncSetAttribute");
  // the host copies are only needed for parsing
  std::vector<PzkTerm>().swap(r->terms); std::vector<PzkCoef>().swap(r->coefs);
  *out = r;
  return PZK_OK;
}

uint32_t pzk_r1cs_wires(const pzk_r1cs* r) { return r ? r->n_wires : 0; }
uint32_t pzk_r1cs_constraints(const pzk_r1cs* r) { return r ? r->n_constraints : 0; }
uint64_t pzk_r1cs_terms(const pzk_r1cs* r) { return r ? r->n_terms : 0; }

// witness planes for `lanes` lanes (rounded up to whole warps), status / first_bad for `batch`
static int r1cs_reserve(pzk_r1cs* r, uint64_t lanes, uint64_t batch, char* err, size_t err_len) {
  lanes = (lanes + 31) / 32 * 32;
  if (lanes > r->W_lanes) {
    if (r->d_W) { cudaFree(r->d_W); r->d_W = nullptr; r->W_lanes = 0; }
    CKE(cudaMalloc((void**)&r->d_W, lanes * r->n_wires * 32));
    r->W_lanes = lanes;
  }
  if (batch + 32 > r->st_cap) {
    if (r->d_status) cudaFree(r->d_status);
    if (r->d_bad) cudaFree(r->d_bad);
    r->d_status = nullptr; r->d_bad = nullptr; r->st_cap = 0;
    CKE(cudaMalloc((void**)&r->d_status, (batch + 32) * 4));
    CKE(cudaMalloc((void**)&r->d_bad, (batch + 32) * 8));
    r->st_cap = batch + 32;
  }
  return PZK_OK;
}

// lanes per pass that fit next to what is already resident on the device
static uint64_t r1cs_lanes_that_fit(pzk_r1cs* r, uint64_t batch, uint64_t extra_per_lane) {
  size_t free_b = 0, total_b = 0;
  cudaMemGetInfo(&free_b, &total_b);
  uint64_t have = (uint64_t)free_b + r->W_lanes * r->n_wires * 32;
  uint64_t per_lane = (uint64_t)r->n_wires * 32 + extra_per_lane;
  uint64_t L = std::max<uint64_t>(32, (uint64_t)(have * 0.85) / per_lane / 32 * 32);
  return std::min<uint64_t>(L, (batch + 31) / 32 * 32);
}

// one launch of the stream kernel over n lanes whose canonical wires are in r->d_W
static int r1cs_launch(pzk_r1cs* r, uint64_t n, uint64_t out_base, cudaStream_t stream, float* ms, char* err, size_t err_len) {
  if (r->n_constraints == 0) { if (ms) *ms = 0; return PZK_OK; }   // nothing to check: every verdict stays 1
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, r->device);
  const unsigned gridl = (unsigned)((n + 127) / 128);
  R1csParams p;
  p.rows = r->d_rows; p.terms = r->d_terms; p.tiles = r->d_tiles; p.n_tiles = (u32)r->tiles.size();
  // row chunks: fill the machine about 2 resident waves deep (6 CTAs per SM), at least one tile per chunk
  uint64_t want_ctas = (uint64_t)n_sm * 12;
  uint64_t chunks = std::max<uint64_t>(1, want_ctas / gridl);
  if (chunks > r->tiles.size()) chunks = r->tiles.size();
  p.tiles_per_chunk = (u32)((r->tiles.size() + chunks - 1) / chunks);
  chunks = (r->tiles.size() + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
  p.coefs = r->d_coefs; p.coef_kind = r->d_kind; p.coef_mag = r->d_mag; p.W = r->d_W; p.n_wires = r->n_wires; p.n_lanes = n;
  p.status = r->d_status + out_base; p.first_bad = r->d_bad + out_base;
  cudaEvent_t ea, eb; cudaEventCreate(&ea); cudaEventCreate(&eb);
  cudaEventRecord(ea, stream);
  r1cs_stream_kernel<<<dim3((unsigned)chunks, gridl), 128, R1CS_SMEM_BYTES, stream>>>(p);
  cudaEventRecord(eb, stream);
  cudaError_t ce = cudaStreamSynchronize(stream);
  if (ce == cudaSuccess) ce = cudaGetLastError();
  float t = 0; cudaEventElapsedTime(&t, ea, eb);
  cudaEventDestroy(ea); cudaEventDestroy(eb);
  if (ce != cudaSuccess) { set_err(err, err_len, std::string("r1cs_stream_kernel: ") + cudaGetErrorString(ce)); return PZK_ECUDA; }
  if (ms) *ms = t;
  return PZK_OK;
}

static int r1cs_collect(pzk_r1cs* r, uint64_t batch, int* verdicts, int64_t* first_bad, char* err, size_t err_len) {
  std::vector<u32> st(batch); std::vector<unsigned long long> bad(batch);
  CKE(cudaMemcpy(st.data(), r->d_status, batch * 4, cudaMemcpyDeviceToHost));
  CKE(cudaMemcpy(bad.data(), r->d_bad, batch * 8, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < batch; i++) {
    verdicts[i] = (st[i] & PZK_LANE_CONSTRAINT) ? 0 : 1;
    if (first_bad) first_bad[i] = (st[i] & PZK_LANE_CONSTRAINT) ? (int64_t)bad[i] : -1;
  }
  return PZK_OK;
}

int pzk_r1cs_check(pzk_r1cs* r, const uint8_t* witnesses_le32, uint64_t batch, int* verdicts, int64_t* first_bad,
                   double* kernel_ms, char* err, size_t err_len) {
  if (!r || !witnesses_le32 || !verdicts || batch == 0) return PZK_EINVAL;
  CKE(cudaSetDevice(r->device));
  const uint64_t L = r1cs_lanes_that_fit(r, batch, (uint64_t)r->n_wires * 32);  // + the AoS staging buffer
  int rc = r1cs_reserve(r, L, batch, err, err_len);
  if (rc) return rc;
  u64* d_wit = nullptr;
  CKE(cudaMalloc((void**)&d_wit, std::min<uint64_t>(L, batch) * r->n_wires * 32));
  CKE(cudaMemset(r->d_status, 0, (batch + 32) * 4));
  CKE(cudaMemset(r->d_bad, 0xff, (batch + 32) * 8));
  double total_ms = 0;
  for (uint64_t base = 0; base < batch && rc == PZK_OK; base += L) {
    const uint64_t n = std::min<uint64_t>(L, batch - base);
    cudaError_t ce = cudaMemcpy(d_wit, witnesses_le32 + base * r->n_wires * 32, n * r->n_wires * 32, cudaMemcpyHostToDevice);
    if (ce != cudaSuccess) { set_err(err, err_len, cudaGetErrorString(ce)); rc = PZK_ECUDA; break; }
    if (n % 32) cudaMemset(r->d_W + (n / 32) * r->n_wires * 128, 0, (uint64_t)r->n_wires * 1024);  // padding lanes of the last warp
    const unsigned gridl = (unsigned)((n + 127) / 128);
    const unsigned gy = (unsigned)std::min<uint64_t>(std::max<uint64_t>(r->n_wires / 64, 1), 4096);
    load_witness_blocked_kernel<<<dim3(gridl, gy), 128>>>(d_wit, r->n_wires, n, r->d_W, r->d_status + base);
    float ms = 0;
    rc = r1cs_launch(r, n, base, 0, &ms, err, err_len);
    total_ms += ms;
  }
  cudaFree(d_wit);
  if (rc) return rc;
  if (kernel_ms) *kernel_ms = total_ms;
  return r1cs_collect(r, batch, verdicts, first_bad, err, err_len);
}

// Device-resident hand-off (the north star's two kernels back to back): the evaluator runs the resident batch
// of `c` and writes every wire of the selected lanes straight into the checker's planes - canonical values,
// blocked by warps, no host round trip of the 72 MB witnesses - then the stream kernel checks EVERY row of the
// .r1cs on them.  Wires are matched by index: the .r1cs must be the one compiled with the program.
int pzk_r1cs_check_circuit(pzk_r1cs* r, pzk_circuit* c, const uint64_t* lanes, uint64_t n_lanes, int* verdicts,
                           int64_t* first_bad, double* eval_ms, double* check_ms, char* err, size_t err_len) {
  if (!r || !c || !lanes || !verdicts || n_lanes == 0 || c->batch == 0) return PZK_EINVAL;
  if (r->device != c->device) { set_err(err, err_len, "the r1cs handle and the circuit live on different devices"); return PZK_EINVAL; }
  if (r->n_wires != c->h.n_wires) {
    set_err(err, err_len, "Invalid witness length. Circuit: " + std::to_string(r->n_wires) + ", witness: " + std::to_string(c->h.n_wires) +
                              " (this .r1cs does not belong to the program: different wire layout)");
    return PZK_EFORMAT;
  }
  for (uint64_t j = 0; j < n_lanes; j++) if (lanes[j] >= c->batch) return PZK_EINVAL;
  CKE(cudaSetDevice(r->device));
  const uint64_t fit = r1cs_lanes_that_fit(r, n_lanes, 0);
  if (fit < (n_lanes + 31) / 32 * 32) { set_err(err, err_len, "not enough device memory for " + std::to_string(n_lanes) + " full witnesses (" + std::to_string(fit) + " fit)"); return PZK_ENOMEM; }
  int rc = r1cs_reserve(r, n_lanes, n_lanes, err, err_len);
  if (rc) return rc;
  CKE(cudaMemsetAsync(r->d_status, 0, (n_lanes + 32) * 4, c->stream));
  CKE(cudaMemsetAsync(r->d_bad, 0xff, (n_lanes + 32) * 8, c->stream));
  // the evaluator's export writes every wire 1.. of every selected lane; padding lanes of the last warp read zeros
  if (n_lanes % 32) CKE(cudaMemsetAsync(r->d_W + (n_lanes / 32) * r->n_wires * 128, 0, (uint64_t)r->n_wires * 1024, c->stream));
  cudaEvent_t ea, eb; cudaEventCreate(&ea); cudaEventCreate(&eb);
  cudaEventRecord(ea, c->stream);
  RunOpts o; o.export_lanes = lanes; o.n_export = n_lanes; o.d_witnesses = r->d_W; o.blocked = true;
  rc = run_batch(c, o);
  if (rc) { set_err(err, err_len, c->err); cudaEventDestroy(ea); cudaEventDestroy(eb); return rc; }
  wire0_blocked_kernel<<<(unsigned)((n_lanes + 127) / 128), 128, 0, c->stream>>>(r->d_W, r->n_wires, n_lanes);
  cudaEventRecord(eb, c->stream);
  CKE(cudaStreamSynchronize(c->stream));
  float t_eval = 0; cudaEventElapsedTime(&t_eval, ea, eb);
  cudaEventDestroy(ea); cudaEventDestroy(eb);
  float t_check = 0;
  rc = r1cs_launch(r, n_lanes, 0, c->stream, &t_check, err, err_len);
  if (rc) return rc;
  if (eval_ms) *eval_ms = t_eval;
  if (check_ms) *check_ms = t_check;
  return r1cs_collect(r, n_lanes, verdicts, first_bad, err, err_len);
}

int pzk_r1cs_check_batch(const char* r1cs_path, const uint8_t* witnesses_le32, uint64_t batch, int cuda_device,
                         int* verdicts, int64_t* first_bad, double* kernel_ms, char* err, size_t err_len) {
  if (!r1cs_path || !witnesses_le32 || !verdicts || batch == 0) return PZK_EINVAL;
  pzk_r1cs* r = nullptr;
  int rc = pzk_r1cs_open(r1cs_path, cuda_device, &r, err, err_len);
  if (rc) return rc;
  rc = pzk_r1cs_check(r, witnesses_le32, batch, verdicts, first_bad, kernel_ms, err, err_len);
  pzk_r1cs_close(r);
  return rc;
}

// section table of a .wtns image, every length checked against wtns_len
static int parse_wtns(const uint8_t* wtns, uint64_t wtns_len, const uint8_t** data, uint32_t* n_wires, char* err, size_t err_len) {
  if (wtns_len < 12 || memcmp(wtns, "wtns", 4) != 0) { set_err(err, err_len, "not a wtns file"); return PZK_EFORMAT; }
  uint32_t nsec; memcpy(&nsec, wtns + 8, 4);
  uint64_t pos = 12;
  const uint8_t* hdr = nullptr; uint64_t hdr_len = 0; const uint8_t* dat = nullptr; uint64_t data_len = 0;
  for (uint32_t s = 0; s < nsec; s++) {
    if (pos + 12 > wtns_len) { set_err(err, err_len, "wtns: truncated section table"); return PZK_EFORMAT; }
    uint32_t type; uint64_t len; memcpy(&type, wtns + pos, 4); memcpy(&len, wtns + pos + 4, 8);
    pos += 12;
    if (len > wtns_len - pos) { set_err(err, err_len, "wtns: section runs past the end of the buffer"); return PZK_EFORMAT; }
    if (type == 1) { hdr = wtns + pos; hdr_len = len; }
    if (type == 2) { dat = wtns + pos; data_len = len; }
    pos += len;
  }
  if (!hdr || !dat) { set_err(err, err_len, "wtns: missing section"); return PZK_EFORMAT; }
  if (hdr_len < 4) { set_err(err, err_len, "wtns: header section too short"); return PZK_EFORMAT; }
  uint32_t n8, nw; memcpy(&n8, hdr, 4);
  if (n8 != 32 || hdr_len < 40 || memcmp(hdr + 4, pzk::FR_P.w, 32) != 0) {
    set_err(err, err_len, "Curve of the witness does not match the curve of the r1cs");
    return PZK_EFORMAT;
  }
  memcpy(&nw, hdr + 36, 4);
  if (data_len != 32ull * nw) { set_err(err, err_len, "wtns: bad data section length"); return PZK_EFORMAT; }
  *data = dat; *n_wires = nw;
  return PZK_OK;
}

int pzk_r1cs_check_wtns(pzk_r1cs* r, const uint8_t* wtns, uint64_t wtns_len, int* verdict, int64_t* first_bad,
                        char* err, size_t err_len) {
  if (!r || !wtns || !verdict) return PZK_EINVAL;
  const uint8_t* data = nullptr; uint32_t nw = 0;
  int rc = parse_wtns(wtns, wtns_len, &data, &nw, err, err_len);
  if (rc) return rc;
  if (r->n_wires != nw) { set_err(err, err_len, "Invalid witness length. Circuit: " + std::to_string(r->n_wires) + ", witness: " + std::to_string(nw)); return PZK_EFORMAT; }
  int64_t fb = -1;
  rc = pzk_r1cs_check(r, data, 1, verdict, &fb, nullptr, err, err_len);
  if (first_bad) *first_bad = fb;
  return rc;
}

int pzk_wtns_check(const char* r1cs_path, const uint8_t* wtns, uint64_t wtns_len, int cuda_device, int* verdict,
                   int64_t* first_bad, char* err, size_t err_len) {
  if (!r1cs_path || !wtns || !verdict) return PZK_EINVAL;
  const uint8_t* data = nullptr; uint32_t nw = 0;
  int rc = parse_wtns(wtns, wtns_len, &data, &nw, err, err_len);   // format errors before any device work
  if (rc) return rc;
  pzk_r1cs* r = nullptr;
  rc = pzk_r1cs_open(r1cs_path, cuda_device, &r, err, err_len);
  if (rc) return rc;
  rc = pzk_r1cs_check_wtns(r, wtns, wtns_len, verdict, first_bad, err, err_len);
  pzk_r1cs_close(r);
  return rc;
}
