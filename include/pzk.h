/*
 * pzk.h - C ABI of the B200-native batched witness generator / R1CS checker.
 *
 * Drop-in boundary for ONE hot path of fizzy74/passport-zk-circuits: the pair
 *     wasm_tester(circuit).calculateWitness(input)  ->  checkConstraints(w)
 * (/root/reference/test/automatisationTest.js:37-51), the documented
 *     generate_witness.js <wasm> <input.json> <out.wtns>   (calculateWTNSBin)
 * (/root/reference/circuits/scripts/gen-witness.sh:25, prove.sh:25) and
 * `snarkjs wtns check` semantics.
 * A Node N-API shim, a cgo stub or Python ctypes bind exactly these symbols
 * (INTEGRATION.md); nothing torch-typed crosses this boundary.
 *
 * Conventions
 *   - field elements cross the boundary as 32 bytes, little endian, canonical
 *     (non-Montgomery) - the layout of a .wtns section-2 entry;
 *   - inputs are the flattened main-component input signals in witness order
 *     (public inputs first, then private, each in declaration order);
 *     pzk_circuit_meta_json() names them;
 *   - every function returns 0 on success, a negative PZK_E* code otherwise;
 *     pzk_last_error() gives the message.  Handles are single-writer.
 *   - there is NO CPU fallback: without a CUDA device every compute entry point
 *     returns PZK_ENODEVICE.
 */
#ifndef PZK_H
#define PZK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PZK_OK 0
#define PZK_EINVAL -1
#define PZK_EIO -2
#define PZK_ENODEVICE -3
#define PZK_ECUDA -4
#define PZK_ECOMPILE -5
#define PZK_ENOMEM -6
#define PZK_EFORMAT -7

/* per-lane status bits (mirror of the circom runtime's exception codes):
 *   0 = witness valid and every constraint satisfied                                  */
#define PZK_STATUS_ASSERT 1u      /* runtime assert() failed            ("Assert Failed.") */
#define PZK_STATUS_CONSTRAINT 2u  /* a `===` / R1CS row does not hold   ("Assert Failed.") */
#define PZK_STATUS_INPUT_RANGE 4u /* input outside its declared width or >= p              */
#define PZK_STATUS_BIGDIV 8u      /* long_div precondition violated (modulus top limb 0)   */

typedef struct pzk_circuit pzk_circuit;

/* ---- compiler: the role of `circom --r1cs --wasm --sym`
 *      (/root/reference/circuits/scripts/compile-circuit.sh:34).
 * Writes <out_prefix>.pzkp (program), .r1cs and .sym.  bits_names/bits_widths declare
 * main inputs that are narrower than a field element (e.g. "dg1":1, "pubkey":64). */
int pzk_compile(const char* main_circom_path, const char* out_prefix, const char* const* bits_names,
                const int* bits_widths, int n_bits, uint32_t segment_ops, char* err, size_t err_len);

/* Same with options.  PZK_COMPILE_STATIC_DEF_ROWS: the row added by `x <== e` compares the wire x with
 * the expression its value was computed from, so it holds for every input; with this flag such rows
 * are discharged at compile time (like the alias rows `a <== b` always are) and only `===` rows and
 * rows over `<--` hints are evaluated at run time.  status / first_bad are identical either way; the
 * default keeps every non-alias row as a run-time check.  PZK_COMPILE_NO_INTRINSICS: unroll long_div
 * instead of using the device intrinsic (validation of the intrinsic).                             */
#define PZK_COMPILE_STATIC_DEF_ROWS 1u
#define PZK_COMPILE_NO_INTRINSICS 2u
/* Beyond alias rows the compiler discharges two more kinds of rows at compile time, each by a proof it
 * checks itself: rows whose wires are all truth-table functions of <= 8 proven bits (evaluated for every
 * assignment of those bits), and rows that cancel when their wires are expanded through the ops that
 * define them (`z <== x + 2^k * y`, `z <== x * y`).  <prefix>.rowkind records the kind per constraint.
 * This flag keeps both kinds as run-time checks instead.                                              */
#define PZK_COMPILE_NO_TABLE_PROOFS 4u
/* Bit-field views and packed truth tables (pzk_program.h): signals that are bit fields of a word the program
 * computes anyway (Num2Bits / GetLastNBits outputs, running sums, `bit * 2^i`) get no op of their own, one-bit
 * truth tables over rotated words (the XOR3 / Ch / Maj arrays of SHA) are packed 32 or 64 to a record, and rows
 * that are identities over the bits of a word are discharged at compile time (kind 5 in <prefix>.rowkind; a row
 * that reduces to "the high bits of this word are zero" stays a run-time range check).  Same wires, same
 * verdicts; these flags switch the rewrite off (validation; PZK_COMPILE_NO_TABLE_PROOFS implies NO_VIEWS). */
#define PZK_COMPILE_NO_VIEWS 8u
#define PZK_COMPILE_NO_VECTORIZE 16u
/* Also write <prefix>.O1.r1cs and <prefix>.O1.sym: the constraint system after an O1-style simplification (the
 * reference's library circuits are compiled with `circom --O1`): linear constraints "signal = signal" and "signal =
 * constant" are removed by substitution, repeatedly; the signals they eliminate leave the witness and carry witness
 * index -1 in the .sym; main inputs and outputs always stay.  pzk_circuit_open_ex(program, <prefix>.sym,
 * <prefix>.O1.sym) then produces the witness of that system.  Not calibrated against circom's own numbering (circom is
 * not available in this environment): self-consistent, every surviving constraint holds on the renumbered witness. */
#define PZK_COMPILE_EMIT_O1 32u
int pzk_compile_ex(const char* main_circom_path, const char* out_prefix, const char* const* bits_names,
                   const int* bits_widths, int n_bits, uint32_t segment_ops, uint32_t flags, char* err,
                   size_t err_len);

/* ---- circuit handle: the role of `new WitnessCalculator(wasm)` ---------------------- */
int pzk_circuit_open(const char* program_path, int cuda_device, pzk_circuit** out);
/* The same with the wires numbered after an EXTERNAL .sym - the one the circom compiler wrote for this circuit at
 * whatever optimisation level the consumer's .r1cs / .zkey were built with (the reference compiles with --O2 / --O1:
 * /root/reference/circuits/scripts/compile-circuit.sh:34, circuits/lib/circuits/scripts/compile-circuit.sh:34).
 * Signals are matched by qualified name between program_sym_path (written by pzk_compile next to the program) and
 * external_sym_path; a signal the external file lists with witness index -1 (optimised away), or not at all, is
 * dropped, signals it merges into one wire are written once.  pzk_witness_size(), calculateWitness,
 * calculateWTNSBin and the witness digest then follow the external numbering, so the .wtns is consumable by the
 * external .r1cs (pzk_r1cs_open on it) and its proving key.                                                  */
int pzk_circuit_open_ex(const char* program_path, const char* program_sym_path, const char* external_sym_path,
                        int cuda_device, pzk_circuit** out);
void pzk_circuit_close(pzk_circuit* c);
const char* pzk_last_error(const pzk_circuit* c);
uint32_t pzk_witness_size(const pzk_circuit* c);    /* nWitness (wire 0 = 1)              */
uint32_t pzk_input_size(const pzk_circuit* c);      /* flattened main inputs              */
uint32_t pzk_public_size(const pzk_circuit* c);     /* nPubOut + nPubIn                   */
uint32_t pzk_constraint_count(const pzk_circuit* c);
const char* pzk_circuit_meta_json(const pzk_circuit* c); /* main IO names/dims/offsets    */
/* algorithmic cost of one witness, for roofline accounting              */
int pzk_circuit_stats(const pzk_circuit* c, uint64_t* op_records, uint64_t* f_mul, uint64_t* f_inv,
                      uint64_t* rows, uint64_t* terms, uint64_t* bytes_per_lane);

/* ---- calculateWitness(input)  (witness_calculator.js, external) ----------
 * witness: n_wires x 32 bytes.  *status receives the lane status, *first_bad the first
 * failing constraint index or -1.  The witness is written even when status != 0.       */
int pzk_calculate_witness(pzk_circuit* c, const uint8_t* inputs_le32, uint8_t* witness_le32,
                          uint32_t* status, int64_t* first_bad);

/* ---- calculateWTNSBin(input): byte-identical iden3 .wtns v2 --------------------------
 * needs pzk_wtns_size() bytes in `out`.                                                */
uint64_t pzk_wtns_size(const pzk_circuit* c);
int pzk_calculate_wtns_bin(pzk_circuit* c, const uint8_t* inputs_le32, uint8_t* out, uint32_t* status,
                           int64_t* first_bad);

/* ---- batched path (the product): B independent passports ----------------------------
 * inputs_le32 : [B][n_inputs][32] host memory
 * status      : [B]            first_bad : [B] (may be NULL)
 * public_le32 : [B][n_public][32] (may be NULL) - outputs then public inputs
 * export_lanes/n_export/witnesses_le32: optional full witnesses for selected lanes,
 *               witnesses_le32 is [n_export][n_wires][32].
 * Computes every signal of every lane on the device, checks every constraint
 * (checkConstraints semantics) and copies the results back.                            */
int pzk_witness_batch(pzk_circuit* c, const uint8_t* inputs_le32, uint64_t batch, uint32_t* status,
                      int64_t* first_bad, uint8_t* public_le32, const uint64_t* export_lanes,
                      uint64_t n_export, uint8_t* witnesses_le32);

/* Packed inputs: most inputs of the passport circuits are bits or 64-bit limbs (what
 * /root/reference/test/process_passport.js:590-672 writes), so a lane's inputs can cross the bus
 * as one record of pzk_packed_stride() bytes: [one byte per input declared <= 8 bits] padded to 8,
 * [8 bytes little endian per input declared <= 64 bits], pad to 16, [32 bytes per field input], record
 * padded to 16; within a
 * section the inputs keep their flattened order.  pzk_packed_layout() gives kind (0/1/2) and
 * byte offset per flattened input.                                                         */
uint32_t pzk_packed_stride(const pzk_circuit* c);
int pzk_packed_layout(const pzk_circuit* c, uint32_t* kind, uint32_t* offset);
int pzk_witness_batch_packed(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status,
                             int64_t* first_bad, uint8_t* public_le32);
int pzk_batch_upload_packed(pzk_circuit* c, const uint8_t* packed, uint64_t batch);

/* Streams ("stream-async + pzk_sync").  Every packed call works in tiles of one wave of resident
 * CTAs; the host-to-device copy of tile k+1 and the device-to-host copy of tile k-1's results run on a second
 * stream under the kernels of tile k (give it pinned host buffers, or the copies serialise).  The _async form
 * returns as soon as the work is enqueued: the output arrays are valid after pzk_sync(), and any other call on
 * the handle syncs first.  digest may be NULL (see "witness digest" below).                                   */
int pzk_witness_batch_packed_async(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status,
                                   int64_t* first_bad, uint8_t* public_le32, uint64_t* digest);
int pzk_sync(pzk_circuit* c);

/* One host thread, several devices: handles[i] is a circuit handle opened on device i with the
 * same program; the batch is cut into n contiguous shares (share i = lanes [i*batch/n, (i+1)*batch/n)), every
 * device runs its share asynchronously and all are joined before the call returns.  No collective: the shares
 * are independent.  Outputs are indexed by the caller's lane numbers.                                        */
int pzk_witness_batch_packed_multi(pzk_circuit* const* handles, int n_handles, const uint8_t* packed, uint64_t batch,
                                   uint32_t* status, int64_t* first_bad, uint8_t* public_le32, uint64_t* digest);

/* ---- witness digest: proof that every signal of every lane was computed, without moving 72 MB per witness.
 * With the digest switched on, every run folds EVERY wire of EVERY lane, on the device, into one field element
 *     digest[lane] = sum over wires i of c(i) * w_i   mod p        (32 bytes little endian, canonical)
 * w_i the canonical value of wire i exactly as calculateWTNSBin would write it, c(i) = pzk_digest_weight_of(i)
 * (32-bit odd weights from splitmix64: position sensitive).  A consumer that holds the .wtns of a lane - or the
 * oracle - can recompute it; tests compare it with the oracle's witness for every lane.  The reference returns
 * the whole vector (/root/reference/test/automatisationTest.js:40-50); this is its checksum.                */
int pzk_batch_set_digest(pzk_circuit* c, int on);
int pzk_batch_download_digest(pzk_circuit* c, uint64_t* digest /* [batch][4] */);
uint32_t pzk_digest_weight_of(uint32_t wire);
int pzk_witness_batch_packed_digest(pzk_circuit* c, const uint8_t* packed, uint64_t batch, uint32_t* status,
                                    int64_t* first_bad, uint8_t* public_le32, uint64_t* digest);

/* device-resident variant for measurement: upload once, run many times, download once */
int pzk_batch_upload(pzk_circuit* c, const uint8_t* inputs_le32, uint64_t batch);
int pzk_batch_run(pzk_circuit* c, int check_rows); /* all tiles of the uploaded batch    */
int pzk_batch_download(pzk_circuit* c, uint32_t* status, int64_t* first_bad, uint8_t* public_le32);
/* accumulated CUDA-event time (ms) and launch counts per kernel family since the last reset:
 * which = 0 eval, 1 row check, 2 export/public, 3 whole run, 4 witness digest.                           */
int pzk_profile_get(pzk_circuit* c, int which, double* ms, uint64_t* launches);
void pzk_profile_reset(pzk_circuit* c);
/* accumulated evaluator time per program segment (ms); returns the number of segments */
int pzk_profile_segments(pzk_circuit* c, double* ms, uint32_t n);
void pzk_profile_enable(pzk_circuit* c, int on);
/* lanes processed per tile (0 = choose from free device memory)                        */
int pzk_set_tile_lanes(pzk_circuit* c, uint64_t lanes);
uint64_t pzk_get_tile_lanes(const pzk_circuit* c);
/* lanes of one full wave of resident CTAs of the evaluator (tiles are sized in whole waves)  */
uint64_t pzk_wave_lanes(const pzk_circuit* c);

/* ---- snarkjs `wtns check` semantics on an explicit witness -------------------------
 * r1cs_path: iden3 .r1cs v1; wtns: iden3 .wtns v2 bytes.  *verdict = 1 when every
 * constraint holds, else 0 and *first_bad = index of the first failing constraint.
 * Errors: PZK_EFORMAT when the primes differ ("Curve of the witness does not match").   */
int pzk_wtns_check(const char* r1cs_path, const uint8_t* wtns, uint64_t wtns_len, int cuda_device,
                   int* verdict, int64_t* first_bad, char* err, size_t err_len);
/* batched: witnesses_le32 is [batch][n_wires][32] canonical                             */
int pzk_r1cs_check_batch(const char* r1cs_path, const uint8_t* witnesses_le32, uint64_t batch,
                         int cuda_device, int* verdicts, int64_t* first_bad, double* kernel_ms,
                         char* err, size_t err_len);

/* The same check with the matrices parsed and uploaded ONCE (the role of snarkjs' readR1cs): open, check any
 * number of batches, close.  A truncated or malformed file is PZK_EFORMAT; an .r1cs without constraints accepts
 * every witness.                                                                                             */
typedef struct pzk_r1cs pzk_r1cs;
int pzk_r1cs_open(const char* r1cs_path, int cuda_device, pzk_r1cs** out, char* err, size_t err_len);
void pzk_r1cs_close(pzk_r1cs* r);
uint32_t pzk_r1cs_wires(const pzk_r1cs* r);
uint32_t pzk_r1cs_constraints(const pzk_r1cs* r);
uint64_t pzk_r1cs_terms(const pzk_r1cs* r);
int pzk_r1cs_check(pzk_r1cs* r, const uint8_t* witnesses_le32, uint64_t batch, int* verdicts, int64_t* first_bad,
                   double* kernel_ms, char* err, size_t err_len);
int pzk_r1cs_check_wtns(pzk_r1cs* r, const uint8_t* wtns, uint64_t wtns_len, int* verdict, int64_t* first_bad,
                        char* err, size_t err_len);
/* calculateWitness -> checkConstraints without leaving the device (automatisationTest.js:40-51 as two kernels):
 * evaluates the batch resident in `c` (pzk_batch_upload*), writes every wire of the selected lanes straight into
 * the checker's planes and checks EVERY row of the .r1cs on them.  The .r1cs must be the one compiled with the
 * program (same wire layout; a mismatch is "Invalid witness length" / PZK_EFORMAT).  n_lanes full witnesses must
 * fit in device memory (32 bytes x wires each).                                                               */
int pzk_r1cs_check_circuit(pzk_r1cs* r, pzk_circuit* c, const uint64_t* lanes, uint64_t n_lanes, int* verdicts,
                           int64_t* first_bad, double* eval_ms, double* check_ms, char* err, size_t err_len);

int pzk_device_count(void);
const char* pzk_version(void);

#ifdef __cplusplus
}
#endif
#endif
