// pzk_napi.cc - Node.js N-API shim over the C ABI of include/pzk.h.
//
// This is the "thin N-API-to-C-ABI shim" of the north star: what a maintainer of the reference adds so that
// test/automatisationTest.js:37-51 (wasm_tester(...).calculateWitness / checkConstraints) and
// circuits/scripts/gen-witness.sh:25 (generate_witness.js -> calculateWTNSBin) run on libpzk.so.
// Node is not part of this image, so the file is NOT compiled or tested here (native/binding.gyp builds it with
// node-gyp against passport-zk-circuits_b200/lib/libpzk.so); it only uses the plain C N-API (node_api.h) and the
// symbols pzk.h declares - tests/test_cpu_host.py checks that every pzk_* symbol it names is exported by the library.
//
//   const pzk = require("./build/Release/pzk.node");
//   const c = pzk.open("artifacts/registerIdentity.pzkp", 0 [, programSym, externalSym]);
//   pzk.meta(c)                                   -> JSON string (main inputs / outputs, offsets, declared widths)
//   pzk.calculateWTNSBin(c, inputsLE32)           -> Uint8Array  (witness.wtns bytes; throws "Assert Failed.")
//   pzk.calculateWitness(c, inputsLE32)           -> Uint8Array  (nWires x 32 bytes, canonical little endian)
//   pzk.witnessBatchPacked(c, packed, batch, wantDigest)
//                                                 -> { status: Uint32Array, firstBad: BigInt64Array,
//                                                      pub: Uint8Array, digest: Uint8Array | null }
//   pzk.packedStride(c) / pzk.packedLayout(c)     -> record layout of `packed`
//   pzk.wtnsCheck(r1csPath, wtnsBytes, device)    -> { ok: boolean, firstBad: number }   (snarkjs wtns check)
//   pzk.close(c)
#include <node_api.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <string>

#include "pzk.h"

#define NAPI_OK(call)                                          \
  do {                                                         \
    if ((call) != napi_ok) {                                   \
      napi_throw_error(env, nullptr, "pzk: N-API call failed"); \
      return nullptr;                                          \
    }                                                          \
  } while (0)

static napi_value fail(napi_env env, const char* msg) {
  napi_throw_error(env, nullptr, msg && *msg ? msg : "pzk error");
  return nullptr;
}

static bool get_args(napi_env env, napi_callback_info info, size_t want, napi_value* argv, size_t* got) {
  size_t argc = 8;
  if (napi_get_cb_info(env, info, &argc, argv, nullptr, nullptr) != napi_ok || argc < want) {
    napi_throw_type_error(env, nullptr, "pzk: wrong number of arguments");
    return false;
  }
  if (got) *got = argc;
  return true;
}

static bool get_string(napi_env env, napi_value v, std::string* out) {
  size_t n = 0;
  if (napi_get_value_string_utf8(env, v, nullptr, 0, &n) != napi_ok) return false;
  out->resize(n);
  return napi_get_value_string_utf8(env, v, &(*out)[0], n + 1, &n) == napi_ok;
}

static pzk_circuit* get_handle(napi_env env, napi_value v) {
  void* p = nullptr;
  if (napi_get_value_external(env, v, &p) != napi_ok || !p) {
    napi_throw_type_error(env, nullptr, "pzk: not a circuit handle");
    return nullptr;
  }
  return *static_cast<pzk_circuit**>(p);
}

static bool get_bytes(napi_env env, napi_value v, uint8_t** data, size_t* len) {
  bool is_ta = false;
  if (napi_is_typedarray(env, v, &is_ta) == napi_ok && is_ta) {
    napi_typedarray_type t; napi_value ab; size_t off = 0, n = 0; void* p = nullptr;
    if (napi_get_typedarray_info(env, v, &t, &n, &p, &ab, &off) != napi_ok) return false;
    size_t width = (t == napi_uint8_array || t == napi_int8_array || t == napi_uint8_clamped_array) ? 1
                 : (t == napi_uint16_array || t == napi_int16_array) ? 2
                 : (t == napi_bigint64_array || t == napi_biguint64_array || t == napi_float64_array) ? 8 : 4;
    *data = static_cast<uint8_t*>(p); *len = n * width;
    return true;
  }
  bool is_buf = false;
  if (napi_is_buffer(env, v, &is_buf) == napi_ok && is_buf) {
    void* p = nullptr;
    if (napi_get_buffer_info(env, v, &p, len) != napi_ok) return false;
    *data = static_cast<uint8_t*>(p);
    return true;
  }
  napi_throw_type_error(env, nullptr, "pzk: expected a TypedArray or Buffer");
  return false;
}

static void finalize_handle(napi_env, void* data, void*) {
  pzk_circuit** slot = static_cast<pzk_circuit**>(data);
  if (*slot) pzk_circuit_close(*slot);
  free(slot);
}

// open(programPath, device = 0, programSym?, externalSym?)
static napi_value Open(napi_env env, napi_callback_info info) {
  napi_value argv[8]; size_t argc = 0;
  if (!get_args(env, info, 1, argv, &argc)) return nullptr;
  std::string path, own_sym, ext_sym;
  if (!get_string(env, argv[0], &path)) return fail(env, "pzk.open: program path must be a string");
  int32_t dev = 0;
  if (argc > 1) napi_get_value_int32(env, argv[1], &dev);
  pzk_circuit* c = nullptr;
  int rc;
  if (argc > 3 && get_string(env, argv[2], &own_sym) && get_string(env, argv[3], &ext_sym))
    rc = pzk_circuit_open_ex(path.c_str(), own_sym.c_str(), ext_sym.c_str(), dev, &c);
  else
    rc = pzk_circuit_open(path.c_str(), dev, &c);
  if (rc != PZK_OK) {
    std::string msg = rc == PZK_ENODEVICE ? "no CUDA device: the witness generator has no CPU fallback"
                                          : (c ? pzk_last_error(c) : "pzk_circuit_open failed");
    if (c) pzk_circuit_close(c);
    return fail(env, msg.c_str());
  }
  pzk_circuit** slot = static_cast<pzk_circuit**>(malloc(sizeof(pzk_circuit*)));
  *slot = c;
  napi_value ext;
  NAPI_OK(napi_create_external(env, slot, finalize_handle, nullptr, &ext));
  return ext;
}

static napi_value Close(napi_env env, napi_callback_info info) {
  napi_value argv[8];
  if (!get_args(env, info, 1, argv, nullptr)) return nullptr;
  void* p = nullptr;
  if (napi_get_value_external(env, argv[0], &p) == napi_ok && p) {
    pzk_circuit** slot = static_cast<pzk_circuit**>(p);
    if (*slot) { pzk_circuit_close(*slot); *slot = nullptr; }
  }
  return nullptr;
}

static napi_value Meta(napi_env env, napi_callback_info info) {
  napi_value argv[8];
  if (!get_args(env, info, 1, argv, nullptr)) return nullptr;
  pzk_circuit* c = get_handle(env, argv[0]);
  if (!c) return nullptr;
  napi_value s;
  NAPI_OK(napi_create_string_utf8(env, pzk_circuit_meta_json(c), NAPI_AUTO_LENGTH, &s));
  return s;
}

static napi_value new_u8(napi_env env, size_t n, uint8_t** data) {
  napi_value ab, ta;
  void* p = nullptr;
  if (napi_create_arraybuffer(env, n, &p, &ab) != napi_ok) return nullptr;
  if (napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &ta) != napi_ok) return nullptr;
  *data = static_cast<uint8_t*>(p);
  return ta;
}

// the wasm throws at the first failing `===` / assert: same message here
static bool throw_on_status(napi_env env, uint32_t status, int64_t first_bad) {
  if (status & PZK_STATUS_INPUT_RANGE) { napi_throw_error(env, nullptr, "Input out of its declared range (status 4)"); return true; }
  if (status & PZK_STATUS_BIGDIV) { napi_throw_error(env, nullptr, "Assert Failed. long division precondition violated (status 8)"); return true; }
  if (status & (PZK_STATUS_ASSERT | PZK_STATUS_CONSTRAINT)) {
    std::string m = "Assert Failed. (status " + std::to_string(status) + ", first failing constraint " + std::to_string((long long)first_bad) + ")";
    napi_throw_error(env, nullptr, m.c_str());
    return true;
  }
  return false;
}

static napi_value WitnessImpl(napi_env env, napi_callback_info info, bool wtns) {
  napi_value argv[8];
  if (!get_args(env, info, 2, argv, nullptr)) return nullptr;
  pzk_circuit* c = get_handle(env, argv[0]);
  if (!c) return nullptr;
  uint8_t* in = nullptr; size_t in_len = 0;
  if (!get_bytes(env, argv[1], &in, &in_len)) return nullptr;
  if (in_len != 32ull * pzk_input_size(c)) return fail(env, "Not all inputs have been set (inputs must be nInputs x 32 bytes)");
  const size_t n = wtns ? (size_t)pzk_wtns_size(c) : 32ull * pzk_witness_size(c);
  uint8_t* out = nullptr;
  napi_value ta = new_u8(env, n, &out);
  if (!ta) return fail(env, "pzk: cannot allocate the witness");
  uint32_t status = 0; int64_t first_bad = -1;
  int rc = wtns ? pzk_calculate_wtns_bin(c, in, out, &status, &first_bad) : pzk_calculate_witness(c, in, out, &status, &first_bad);
  if (rc != PZK_OK) return fail(env, pzk_last_error(c));
  if (throw_on_status(env, status, first_bad)) return nullptr;
  return ta;
}
static napi_value CalculateWTNSBin(napi_env env, napi_callback_info info) { return WitnessImpl(env, info, true); }
static napi_value CalculateWitness(napi_env env, napi_callback_info info) { return WitnessImpl(env, info, false); }

static napi_value PackedStride(napi_env env, napi_callback_info info) {
  napi_value argv[8];
  if (!get_args(env, info, 1, argv, nullptr)) return nullptr;
  pzk_circuit* c = get_handle(env, argv[0]);
  if (!c) return nullptr;
  napi_value v;
  NAPI_OK(napi_create_uint32(env, pzk_packed_stride(c), &v));
  return v;
}

// packedLayout(c) -> { kind: Uint32Array, offset: Uint32Array }  (per flattened input)
static napi_value PackedLayout(napi_env env, napi_callback_info info) {
  napi_value argv[8];
  if (!get_args(env, info, 1, argv, nullptr)) return nullptr;
  pzk_circuit* c = get_handle(env, argv[0]);
  if (!c) return nullptr;
  const size_t n = pzk_input_size(c);
  napi_value ab1, ab2, k, o, r;
  void *p1 = nullptr, *p2 = nullptr;
  NAPI_OK(napi_create_arraybuffer(env, 4 * n, &p1, &ab1));
  NAPI_OK(napi_create_arraybuffer(env, 4 * n, &p2, &ab2));
  if (pzk_packed_layout(c, static_cast<uint32_t*>(p1), static_cast<uint32_t*>(p2)) != PZK_OK) return fail(env, pzk_last_error(c));
  NAPI_OK(napi_create_typedarray(env, napi_uint32_array, n, ab1, 0, &k));
  NAPI_OK(napi_create_typedarray(env, napi_uint32_array, n, ab2, 0, &o));
  NAPI_OK(napi_create_object(env, &r));
  NAPI_OK(napi_set_named_property(env, r, "kind", k));
  NAPI_OK(napi_set_named_property(env, r, "offset", o));
  return r;
}

// witnessBatchPacked(c, packed, batch, wantDigest = false)
static napi_value WitnessBatchPacked(napi_env env, napi_callback_info info) {
  napi_value argv[8]; size_t argc = 0;
  if (!get_args(env, info, 3, argv, &argc)) return nullptr;
  pzk_circuit* c = get_handle(env, argv[0]);
  if (!c) return nullptr;
  uint8_t* in = nullptr; size_t in_len = 0;
  if (!get_bytes(env, argv[1], &in, &in_len)) return nullptr;
  int64_t batch = 0;
  napi_get_value_int64(env, argv[2], &batch);
  bool want_digest = false;
  if (argc > 3) napi_get_value_bool(env, argv[3], &want_digest);
  if (batch <= 0 || in_len != (size_t)batch * pzk_packed_stride(c)) return fail(env, "pzk: packed must hold batch x packedStride bytes");
  const size_t n_pub = pzk_public_size(c);
  napi_value ab_s, ab_b, st, fb, pub, dg, r;
  void *ps = nullptr, *pb = nullptr;
  uint8_t *ppub = nullptr, *pdg = nullptr;
  NAPI_OK(napi_create_arraybuffer(env, 4 * (size_t)batch, &ps, &ab_s));
  NAPI_OK(napi_create_arraybuffer(env, 8 * (size_t)batch, &pb, &ab_b));
  pub = new_u8(env, (size_t)batch * n_pub * 32, &ppub);
  if (!pub) return fail(env, "pzk: cannot allocate the public signals");
  if (want_digest) {
    dg = new_u8(env, (size_t)batch * 32, &pdg);
    if (!dg) return fail(env, "pzk: cannot allocate the digests");
    if (pzk_batch_set_digest(c, 1) != PZK_OK) return fail(env, pzk_last_error(c));
  } else {
    pzk_batch_set_digest(c, 0);
    NAPI_OK(napi_get_null(env, &dg));
  }
  int rc = pzk_witness_batch_packed_digest(c, in, (uint64_t)batch, static_cast<uint32_t*>(ps), static_cast<int64_t*>(pb), ppub,
                                           reinterpret_cast<uint64_t*>(pdg));
  if (rc != PZK_OK) return fail(env, pzk_last_error(c));
  NAPI_OK(napi_create_typedarray(env, napi_uint32_array, (size_t)batch, ab_s, 0, &st));
  NAPI_OK(napi_create_typedarray(env, napi_bigint64_array, (size_t)batch, ab_b, 0, &fb));
  NAPI_OK(napi_create_object(env, &r));
  NAPI_OK(napi_set_named_property(env, r, "status", st));
  NAPI_OK(napi_set_named_property(env, r, "firstBad", fb));
  NAPI_OK(napi_set_named_property(env, r, "pub", pub));
  NAPI_OK(napi_set_named_property(env, r, "digest", dg));
  return r;
}

// wtnsCheck(r1csPath, wtnsBytes, device = 0) -> { ok, firstBad }
static napi_value WtnsCheck(napi_env env, napi_callback_info info) {
  napi_value argv[8]; size_t argc = 0;
  if (!get_args(env, info, 2, argv, &argc)) return nullptr;
  std::string path;
  if (!get_string(env, argv[0], &path)) return fail(env, "pzk.wtnsCheck: r1cs path must be a string");
  uint8_t* w = nullptr; size_t n = 0;
  if (!get_bytes(env, argv[1], &w, &n)) return nullptr;
  int32_t dev = 0;
  if (argc > 2) napi_get_value_int32(env, argv[2], &dev);
  int verdict = 0; int64_t first_bad = -1;
  char err[512] = {0};
  if (pzk_wtns_check(path.c_str(), w, n, dev, &verdict, &first_bad, err, sizeof err) != PZK_OK) return fail(env, err);
  napi_value r, ok, fb;
  NAPI_OK(napi_create_object(env, &r));
  NAPI_OK(napi_get_boolean(env, verdict != 0, &ok));
  NAPI_OK(napi_create_int64(env, first_bad, &fb));
  NAPI_OK(napi_set_named_property(env, r, "ok", ok));
  NAPI_OK(napi_set_named_property(env, r, "firstBad", fb));
  return r;
}

static napi_value Init(napi_env env, napi_value exports) {
  const napi_property_descriptor props[] = {
      {"open", nullptr, Open, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"close", nullptr, Close, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"meta", nullptr, Meta, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"calculateWTNSBin", nullptr, CalculateWTNSBin, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"calculateWitness", nullptr, CalculateWitness, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"packedStride", nullptr, PackedStride, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"packedLayout", nullptr, PackedLayout, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"witnessBatchPacked", nullptr, WitnessBatchPacked, nullptr, nullptr, nullptr, napi_default, nullptr},
      {"wtnsCheck", nullptr, WtnsCheck, nullptr, nullptr, nullptr, napi_default, nullptr},
  };
  napi_define_properties(env, exports, sizeof props / sizeof props[0], props);
  return exports;
}

NAPI_MODULE(NODE_GYP_MODULE_NAME, Init)
