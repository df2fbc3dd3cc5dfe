// pzk_api.cu - host runtime behind include/pzk.h: program loading, tile scheduling, CUDA
// streams/events, .wtns / .r1cs codecs.  No CPU fallback: every compute entry point needs a
// CUDA device and fails with PZK_ENODEVICE otherwise.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <map>
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
  // profiling
  bool prof = false;
  double prof_ms[4] = {0, 0, 0, 0};
  uint64_t prof_launches[4] = {0, 0, 0, 0};
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
  c->d_inputs = nullptr; c->d_status = nullptr; c->d_first_bad = nullptr; c->d_public = nullptr;
  c->batch_cap = 0; c->pub_cap = 0; c->d_inputs_bytes = 0;
}

void pzk_circuit_close(pzk_circuit* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  free_tile(c); free_batch(c);
  cudaFree(c->d_ops); cudaFree(c->d_fpool); cudaFree(c->d_coefs); cudaFree(c->d_coef_kind); cudaFree(c->d_coef_mag);
  cudaFree(c->d_list); cudaFree(c->d_in_table); cudaFree(c->d_rows); cudaFree(c->d_terms); cudaFree(c->d_exports); cudaFree(c->d_pub_entries);
  if (c->stream) cudaStreamDestroy(c->stream);
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

int pzk_circuit_open(const char* program_path, int cuda_device, pzk_circuit** out) {
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
  c->smem_bytes = (size_t)c->h.reserved[1] * 8 * 128;

  CK(cudaSetDevice(cuda_device));
  CK(cudaStreamCreate(&c->stream));
  if (c->smem_bytes > 48 * 1024) CK(cudaFuncSetAttribute(eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->smem_bytes));
  {
    int per_sm = 0, n_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, eval_kernel, 128, c->smem_bytes));
    CK(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cuda_device));
    c->wave_lanes = (uint64_t)per_sm * n_sm * 128;
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

// run every tile of the resident batch; optional witness export for selected lanes
static int run_batch(pzk_circuit* c, int check_rows, const uint64_t* export_lanes, uint64_t n_export,
                     u64* d_witnesses /* [n_export][n_wires][4] device */) {
  CK(cudaSetDevice(c->device));
  int rc = ensure_tile(c, c->batch);
  if (rc) return rc;
  const uint64_t L = c->L;
  const uint32_t n_pub = c->h.n_pub_out + c->h.n_pub_in;
  CK(cudaMemsetAsync(c->d_status, 0, c->batch * 4, c->stream));
  CK(cudaMemsetAsync(c->d_first_bad, 0xff, c->batch * 8, c->stream));
  cudaEvent_t ra, rb;
  prof_begin(c, 3, ra, rb);
  u64* d_lane_list = nullptr;
  for (uint64_t base = 0; base < c->batch; base += L) {
    const uint64_t n = std::min<uint64_t>(L, c->batch - base);
    const unsigned grid = (unsigned)((n + 127) / 128);
    // export lanes that fall into this tile
    std::vector<u64> tile_lanes, tile_rows;
    for (uint64_t j = 0; j < n_export; j++)
      if (export_lanes[j] >= base && export_lanes[j] < base + n) { tile_lanes.push_back(export_lanes[j] - base); tile_rows.push_back(j); }
    if (!tile_lanes.empty()) {
      if (d_lane_list) { cudaFree(d_lane_list); d_lane_list = nullptr; }
      CK(cudaMalloc((void**)&d_lane_list, tile_lanes.size() * 8));
      CK(cudaMemcpyAsync(d_lane_list, tile_lanes.data(), tile_lanes.size() * 8, cudaMemcpyHostToDevice, c->stream));
      CK(cudaStreamSynchronize(c->stream));
    }
    for (uint32_t s = 0; s < c->h.n_segments; s++) {
      const PzkSegment& sg = c->segs[s];
      cudaEvent_t ea, eb;
      if (sg.n_ops) {
        EvalParams p;
        p.ops = c->d_ops + sg.op_off; p.n_rec = sg.n_ops; p.U = c->d_U; p.F = c->d_F; p.L = L; p.n_lanes = n;
        p.fpool = c->d_fpool; p.list = c->d_list;
        if (c->packed) {
          p.inputs = reinterpret_cast<const u64*>(reinterpret_cast<const unsigned char*>(c->d_inputs) + base * c->packed_stride);
          p.in_table = c->d_in_table; p.in_stride = c->packed_stride;
        } else { p.inputs = c->d_inputs + base * c->h.n_inputs * 4; p.in_table = nullptr; p.in_stride = 0; }
        p.n_inputs = c->h.n_inputs; p.status = c->d_status + base; p.n_u_slots = c->h.n_u_slots; p.n_f_slots = c->h.n_f_slots;
        p.check_rows = check_rows; p.store_all = (n_export > 0) ? 1 : 0; p.sc.coefs = c->d_coefs; p.sc.coef_kind = c->d_coef_kind; p.sc.coef_mag = c->d_coef_mag;
        p.first_bad = c->d_first_bad + base;
        prof_begin(c, 0, ea, eb);
        eval_kernel<<<grid, 128, c->smem_bytes, c->stream>>>(p);
        prof_end(c, 0, ea, eb);
        if (c->prof) c->pending_seg.push_back((int)s);
      }
      if (!c->seg[s].pub.empty()) {
        ExportParams p;
        p.list = c->d_list;
        p.entries = c->d_pub_entries + c->seg[s].pub_off; p.n_entries = c->seg[s].pub.size();
        p.n_u_slots = c->h.n_u_slots; p.n_f_slots = c->h.n_f_slots;
        p.U = c->d_U; p.F = c->d_F; p.L = L; p.lanes = nullptr; p.lane_base = base; p.n_rows = n;
        p.out = c->d_public; p.out_wires = n_pub; p.wire_off = 1;
        prof_begin(c, 2, ea, eb);
        export_kernel<<<dim3(grid, 1), 128, 0, c->stream>>>(p);
        prof_end(c, 2, ea, eb);
      }
      if (!tile_lanes.empty() && sg.n_exp) {
        // rows of d_witnesses are indexed by position in export_lanes: handle each contiguous run
        for (size_t q = 0; q < tile_lanes.size(); q++) {
          ExportParams p;
          p.list = c->d_list;
          p.entries = c->d_exports + sg.exp_off; p.n_entries = sg.n_exp;
          p.n_u_slots = c->h.n_u_slots; p.n_f_slots = c->h.n_f_slots;
          p.U = c->d_U; p.F = c->d_F; p.L = L; p.lanes = d_lane_list + q; p.lane_base = tile_rows[q]; p.n_rows = 1;
          p.out = d_witnesses; p.out_wires = c->h.n_wires; p.wire_off = 0;
          unsigned gy = (unsigned)std::min<uint64_t>(std::max<uint64_t>(sg.n_exp / 128, 1), 1024);
          prof_begin(c, 2, ea, eb);
          export_rows_kernel<<<gy, 128, 0, c->stream>>>(p);
          prof_end(c, 2, ea, eb);
        }
      }
    }
  }
  prof_end(c, 3, ra, rb);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  if (d_lane_list) cudaFree(d_lane_list);
  prof_collect(c);
  return PZK_OK;
}

int pzk_batch_upload(pzk_circuit* c, const uint8_t* inputs_le32, uint64_t batch) {
  if (!c || !inputs_le32 || batch == 0) return PZK_EINVAL;
  CK(cudaSetDevice(c->device));
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
  int rc = ensure_batch(c, batch, batch * (uint64_t)c->packed_stride);
  if (rc) return rc;
  c->batch = batch; c->packed = true;
  CK(cudaMemcpyAsync(c->d_inputs, packed, batch * (uint64_t)c->packed_stride, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return PZK_OK;
}
int pzk_witness_batch_packed(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status,
                             int64_t* first_bad, uint8_t* public_le32) {
  int rc = pzk_batch_upload_packed(c, packed, batch);
  if (rc) return rc;
  rc = run_batch(c, 1, nullptr, 0, nullptr);
  if (rc) return rc;
  return pzk_batch_download(c, status, first_bad, public_le32);
}

int pzk_batch_run(pzk_circuit* c, int check_rows) {
  if (!c || c->batch == 0) return PZK_EINVAL;
  return run_batch(c, check_rows, nullptr, 0, nullptr);
}

int pzk_batch_download(pzk_circuit* c, uint32_t* status, int64_t* first_bad, uint8_t* public_le32) {
  if (!c || c->batch == 0) return PZK_EINVAL;
  CK(cudaSetDevice(c->device));
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
  rc = run_batch(c, 1, export_lanes, n_export, d_wit);
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
  if (!c || which < 0 || which > 3) return PZK_EINVAL;
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
  c->seg_ms.assign(c->seg_ms.size(), 0.0); for (int i = 0; i < 4; i++) { c->prof_ms[i] = 0; c->prof_launches[i] = 0; } }
void pzk_profile_enable(pzk_circuit* c, int on) { c->prof = on != 0; }

// ---------------------------------------------------------------------------------------
// Generic `wtns check`: any iden3 .r1cs against explicit witnesses (all wires field class).
// ---------------------------------------------------------------------------------------
struct R1csHost {
  uint32_t n_wires = 0, n_constraints = 0;
  std::vector<PzkRow> rows;
  std::vector<PzkTerm> terms;
  std::vector<PzkCoef> coefs;
};

static int parse_r1cs(const char* path, R1csHost& r, std::string& err) {
  FILE* f = fopen(path, "rb");
  if (!f) { err = std::string("cannot open ") + path; return PZK_EIO; }
  fseek(f, 0, SEEK_END); long sz = ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> buf((size_t)sz);
  if (fread(buf.data(), 1, (size_t)sz, f) != (size_t)sz) { fclose(f); err = "short read"; return PZK_EIO; }
  fclose(f);
  if (sz < 12 || memcmp(buf.data(), "r1cs", 4) != 0) { err = "not an r1cs file"; return PZK_EFORMAT; }
  uint32_t nsec; memcpy(&nsec, buf.data() + 8, 4);
  size_t pos = 12;
  const uint8_t* hdr = nullptr; const uint8_t* cons = nullptr; uint64_t cons_len = 0;
  for (uint32_t s = 0; s < nsec && pos + 12 <= (size_t)sz; s++) {
    uint32_t type; uint64_t len; memcpy(&type, buf.data() + pos, 4); memcpy(&len, buf.data() + pos + 4, 8);
    pos += 12;
    if (type == 1) hdr = buf.data() + pos;
    if (type == 2) { cons = buf.data() + pos; cons_len = len; }
    pos += len;
  }
  if (!hdr || !cons) { err = "r1cs: missing header or constraint section"; return PZK_EFORMAT; }
  uint32_t n8; memcpy(&n8, hdr, 4);
  if (n8 != 32 || memcmp(hdr + 4, pzk::FR_P.w, 32) != 0) { err = "r1cs: prime is not the BN254 scalar field"; return PZK_EFORMAT; }
  memcpy(&r.n_wires, hdr + 36, 4);
  memcpy(&r.n_constraints, hdr + 36 + 16 + 8, 4);
  std::map<std::string, uint32_t> cidx;
  const uint8_t* q = cons; const uint8_t* end = cons + cons_len;
  for (uint32_t i = 0; i < r.n_constraints; i++) {
    PzkRow row; row.term_off = (uint32_t)r.terms.size(); row.kind = 0; row.index = i;
    uint16_t cnt[3];
    for (int part = 0; part < 3; part++) {
      if (q + 4 > end) { err = "r1cs: truncated"; return PZK_EFORMAT; }
      uint32_t n; memcpy(&n, q, 4); q += 4;
      if (n > 65535) { err = "r1cs: linear combination too long"; return PZK_EFORMAT; }
      cnt[part] = (uint16_t)n;
      for (uint32_t k = 0; k < n; k++) {
        if (q + 36 > end) { err = "r1cs: truncated"; return PZK_EFORMAT; }
        uint32_t wire; memcpy(&wire, q, 4);
        std::string key((const char*)q + 4, 32);
        q += 36;
        auto it = cidx.find(key);
        uint32_t ci;
        if (it == cidx.end()) {
          ci = (uint32_t)r.coefs.size(); cidx[key] = ci;
          PzkCoef pc; memcpy(pc.plain, key.data(), 32);
          pzk::U256 v = pzk::U256::from_limbs(pc.plain[0], pc.plain[1], pc.plain[2], pc.plain[3]);
          pzk::U256 m1 = pzk::fr_to_mont(pzk::fr_reduce(v)), m2 = pzk::fr_to_mont(m1);
          memcpy(pc.mont, m1.w, 32); memcpy(pc.mont2, m2.w, 32);
          r.coefs.push_back(pc);
        } else ci = it->second;
        PzkTerm t; t.ref = (2u << 30) | wire; t.coef = ci;
        if (wire >= r.n_wires) { err = "r1cs: wire index out of range"; return PZK_EFORMAT; }
        r.terms.push_back(t);
      }
    }
    row.na = cnt[0]; row.nb = cnt[1]; row.nc = cnt[2];
    r.rows.push_back(row);
  }
  return PZK_OK;
}

static int check_batch_impl(R1csHost& r, const uint8_t* witnesses_le32, uint64_t batch, int cuda_device,
                            int* verdicts, int64_t* first_bad, double* kernel_ms, char* err, size_t err_len) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { set_err(err, err_len, "no CUDA device"); return PZK_ENODEVICE; }
#define CKE(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { set_err(err, err_len, std::string(#call) + ": " + cudaGetErrorString(e_)); return PZK_ECUDA; } } while (0)
  CKE(cudaSetDevice(cuda_device));
  // tiles of the A/B/C stream: at most R1CS_TILE_ROWS rows and R1CS_TILE_TERMS terms, even first term
  std::vector<R1csTile> tiles;
  {
    size_t i = 0, n = r.rows.size();
    while (i < n) {
      R1csTile t; t.row0 = (uint32_t)i; t.term0 = r.rows[i].term_off & ~1u;
      uint32_t end_term = r.rows[i].term_off;
      uint32_t nr = 0;
      while (i < n && nr < R1CS_TILE_ROWS) {
        uint32_t nt = r.rows[i].na + r.rows[i].nb + r.rows[i].nc;
        if (nt + 2 > R1CS_TILE_TERMS) { set_err(err, err_len, "r1cs: a constraint has too many terms for the streaming tile"); return PZK_EFORMAT; }
        if (r.rows[i].term_off + nt - t.term0 + 1 > R1CS_TILE_TERMS) break;
        end_term = r.rows[i].term_off + nt; nr++; i++;
      }
      t.n_rows = nr; t.n_terms = end_term - t.term0;
      tiles.push_back(t);
    }
  }
  PzkRow* d_rows; PzkTerm* d_terms; PzkCoef* d_coefs; unsigned char* d_kind; u64* d_mag; R1csTile* d_tiles;
  std::vector<unsigned char> kind; std::vector<u64> mag;
  classify_coefs(r.coefs.data(), (uint32_t)r.coefs.size(), kind, mag);
  std::vector<PzkTerm> terms_padded = r.terms;
  terms_padded.resize(r.terms.size() + 4, PzkTerm{0, 0});
  CKE(upload(&d_rows, r.rows.data(), r.rows.size() * sizeof(PzkRow)));
  CKE(upload(&d_terms, terms_padded.data(), terms_padded.size() * sizeof(PzkTerm)));
  CKE(upload(&d_coefs, r.coefs.data(), r.coefs.size() * sizeof(PzkCoef)));
  CKE(upload(&d_kind, kind.data(), kind.size()));
  CKE(upload(&d_mag, mag.data(), mag.size() * 8));
  CKE(upload(&d_tiles, tiles.data(), tiles.size() * sizeof(R1csTile)));
  CKE(cudaFuncSetAttribute(r1cs_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)R1CS_SMEM_BYTES));
  size_t free_b = 0, total_b = 0;
  CKE(cudaMemGetInfo(&free_b, &total_b));
  uint64_t per_lane = (uint64_t)r.n_wires * 64;  // AoS staging + blocked Montgomery plane
  uint64_t L = std::max<uint64_t>(32, (uint64_t)(free_b * 0.8) / per_lane / 32 * 32);
  if (L > (batch + 31) / 32 * 32) L = (batch + 31) / 32 * 32;
  u64 *d_wit, *d_F; u32* d_status; unsigned long long* d_bad;
  CKE(cudaMalloc((void**)&d_wit, L * r.n_wires * 32));
  CKE(cudaMalloc((void**)&d_F, L * r.n_wires * 32));
  CKE(cudaMemset(d_F, 0, L * r.n_wires * 32));  // padding lanes of the last warp read zeros
  CKE(cudaMalloc((void**)&d_status, (batch + 32) * 4));
  CKE(cudaMalloc((void**)&d_bad, (batch + 32) * 8));
  CKE(cudaMemset(d_status, 0, (batch + 32) * 4));
  CKE(cudaMemset(d_bad, 0xff, (batch + 32) * 8));
  cudaEvent_t ea, eb; cudaEventCreate(&ea); cudaEventCreate(&eb);
  double total_ms = 0;
  int n_sm = 148;
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, cuda_device);
  for (uint64_t base = 0; base < batch; base += L) {
    uint64_t n = std::min<uint64_t>(L, batch - base);
    CKE(cudaMemcpy(d_wit, witnesses_le32 + base * r.n_wires * 32, n * r.n_wires * 32, cudaMemcpyHostToDevice));
    unsigned gridl = (unsigned)((n + 127) / 128);
    unsigned gy = (unsigned)std::min<uint64_t>(std::max<uint64_t>(r.n_wires / 64, 1), 4096);
    load_witness_blocked_kernel<<<dim3(gridl, gy), 128>>>(d_wit, r.n_wires, n, d_F, d_status + base);
    R1csParams p;
    p.rows = d_rows; p.terms = d_terms; p.tiles = d_tiles; p.n_tiles = (u32)tiles.size();
    // enough chunks to fill the machine ~4 CTAs deep, at least 4 tiles per chunk when possible
    uint64_t want_ctas = (uint64_t)n_sm * 4;
    uint64_t chunks = std::max<uint64_t>(1, want_ctas / gridl);
    if (chunks > tiles.size()) chunks = tiles.size();
    p.tiles_per_chunk = (u32)((tiles.size() + chunks - 1) / chunks);
    chunks = (tiles.size() + p.tiles_per_chunk - 1) / p.tiles_per_chunk;
    p.coefs = d_coefs; p.coef_kind = d_kind; p.coef_mag = d_mag; p.F = d_F; p.n_wires = r.n_wires; p.n_lanes = n;
    p.status = d_status + base; p.first_bad = d_bad + base;
    cudaEventRecord(ea);
    r1cs_stream_kernel<<<dim3((unsigned)chunks, gridl), 128, R1CS_SMEM_BYTES>>>(p);
    cudaEventRecord(eb);
    CKE(cudaDeviceSynchronize());
    float ms = 0; cudaEventElapsedTime(&ms, ea, eb); total_ms += ms;
  }
  std::vector<u32> st(batch); std::vector<unsigned long long> bad(batch);
  CKE(cudaMemcpy(st.data(), d_status, batch * 4, cudaMemcpyDeviceToHost));
  CKE(cudaMemcpy(bad.data(), d_bad, batch * 8, cudaMemcpyDeviceToHost));
  for (uint64_t i = 0; i < batch; i++) {
    verdicts[i] = (st[i] & PZK_LANE_CONSTRAINT) ? 0 : 1;
    if (first_bad) first_bad[i] = (st[i] & PZK_LANE_CONSTRAINT) ? (int64_t)bad[i] : -1;
  }
  if (kernel_ms) *kernel_ms = total_ms;
  cudaEventDestroy(ea); cudaEventDestroy(eb);
  cudaFree(d_rows); cudaFree(d_terms); cudaFree(d_coefs); cudaFree(d_kind); cudaFree(d_mag); cudaFree(d_tiles);
  cudaFree(d_wit); cudaFree(d_F); cudaFree(d_status); cudaFree(d_bad);
  return PZK_OK;
}

int pzk_r1cs_check_batch(const char* r1cs_path, const uint8_t* witnesses_le32, uint64_t batch, int cuda_device,
                         int* verdicts, int64_t* first_bad, double* kernel_ms, char* err, size_t err_len) {
  if (!r1cs_path || !witnesses_le32 || !verdicts || batch == 0) return PZK_EINVAL;
  R1csHost r; std::string e;
  int rc = parse_r1cs(r1cs_path, r, e);
  if (rc) { set_err(err, err_len, e); return rc; }
  return check_batch_impl(r, witnesses_le32, batch, cuda_device, verdicts, first_bad, kernel_ms, err, err_len);
}

int pzk_wtns_check(const char* r1cs_path, const uint8_t* wtns, uint64_t wtns_len, int cuda_device, int* verdict,
                   int64_t* first_bad, char* err, size_t err_len) {
  if (!r1cs_path || !wtns || !verdict) return PZK_EINVAL;
  if (wtns_len < 12 || memcmp(wtns, "wtns", 4) != 0) { set_err(err, err_len, "not a wtns file"); return PZK_EFORMAT; }
  uint32_t nsec; memcpy(&nsec, wtns + 8, 4);
  uint64_t pos = 12;
  const uint8_t* hdr = nullptr; const uint8_t* data = nullptr; uint64_t data_len = 0;
  for (uint32_t s = 0; s < nsec && pos + 12 <= wtns_len; s++) {
    uint32_t type; uint64_t len; memcpy(&type, wtns + pos, 4); memcpy(&len, wtns + pos + 4, 8);
    pos += 12;
    if (type == 1) hdr = wtns + pos;
    if (type == 2) { data = wtns + pos; data_len = len; }
    pos += len;
  }
  if (!hdr || !data) { set_err(err, err_len, "wtns: missing section"); return PZK_EFORMAT; }
  uint32_t n8, nw; memcpy(&n8, hdr, 4);
  if (n8 != 32 || memcmp(hdr + 4, pzk::FR_P.w, 32) != 0) {
    set_err(err, err_len, "Curve of the witness does not match the curve of the r1cs");
    return PZK_EFORMAT;
  }
  memcpy(&nw, hdr + 36, 4);
  if (data_len != 32ull * nw) { set_err(err, err_len, "wtns: bad data section length"); return PZK_EFORMAT; }
  R1csHost r; std::string e;
  int rc = parse_r1cs(r1cs_path, r, e);
  if (rc) { set_err(err, err_len, e); return rc; }
  if (r.n_wires != nw) { set_err(err, err_len, "Invalid witness length. Circuit: " + std::to_string(r.n_wires) + ", witness: " + std::to_string(nw)); return PZK_EFORMAT; }
  int64_t fb = -1;
  rc = check_batch_impl(r, data, 1, cuda_device, verdict, &fb, nullptr, err, err_len);
  if (first_bad) *first_bad = fb;
  return rc;
}

