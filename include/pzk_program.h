/*
 * pzk_program.h - on-disk / in-memory format of a compiled circuit ("program").
 *
 * A program is what the from-scratch compiler (csrc/compiler) lowers a circom
 * circuit to and what the CUDA evaluator executes lane-per-witness.  It plays
 * the role of the `<name>.wasm` that circom emits for the reference
 * (/root/reference/circuits/scripts/compile-circuit.sh:34,
 *  /root/reference/circuits/scripts/gen-witness.sh:25).
 *
 * Value classes
 *   U : exact unsigned 64-bit integer (the signal's canonical value is < 2^64)
 *   F : BN254 Fr element in Montgomery form, 4 x 64-bit limbs (R = 2^256)
 *   N : canonical (non-Montgomery) 256-bit integer, lives in an F slot; only a
 *       temporary for circom's integer operators (>> & \ % < ...) on wide values
 *   Z : exact signed integer |v| < 2^250 in 256-bit two's complement, lives in an F slot: values the compiler
 *       proves to be integers too wide for 64 bits but far below p / 2 (the limb products, Karatsuba sums and
 *       carries of the big-integer multipliers).  + - * are plain integer instructions instead of Montgomery
 *       products; the canonical value of a negative v is p + v
 *
 * Storage: two slot planes per lane tile, slot-major / lane-minor so that a warp
 * touches 32 consecutive words:
 *   U plane:  u64 U[slot][lane]
 *   F plane:  u64 F[slot][limb 0..3][lane]
 * Slots are reused by the compiler once a value is dead (after its last op use
 * and after the segment in which its last constraint row is checked).
 */
#ifndef PZK_PROGRAM_H
#define PZK_PROGRAM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PZK_MAGIC 0x314b5a50u /* "PZK1" */
#define PZK_VERSION 15u

/* ---- opcodes ------------------------------------------------------------ */
enum PzkOpcode {
  PZK_NOP = 0,
  /* U class (wrapping 64-bit ring ops, exact integer ops) */
  PZK_U_CONST = 1, /* dst = (b << 32) | a                                  */
  PZK_U_ADD = 2,   /* dst = a + b                                            */
  PZK_U_SUB = 3,
  PZK_U_MUL = 4,
  PZK_U_DIV = 5, /* integer quotient, x / 0 = 0                            */
  PZK_U_MOD = 6,
  PZK_U_SHR = 7, /* shift counts >= 64 give 0                              */
  PZK_U_SHL = 8,
  PZK_U_AND = 9,
  PZK_U_OR = 10,
  PZK_U_XOR = 11,
  PZK_U_LT = 12, /* unsigned compare -> 0/1                                */
  PZK_U_LE = 13,
  PZK_U_EQ = 14,
  PZK_U_NE = 15,
  PZK_U_SEL = 16, /* dst = a ? b : c        (ext word holds c)              */
  PZK_U_LUT = 17, /* dst = (imm16 >> (a | b<<1 | c<<2 | d<<3)) & 1  (ext); operand j contributes bit
                     (ext.f >> 8j) & 255 of its word (0 for plain one-bit values)                     */
  PZK_I_LT = 18,  /* signed 64-bit compare -> 0/1                           */
  PZK_I_LE = 19,
  PZK_U_LUTV = 20, /* dst = list64[e + (a | b<<1 | c<<2 | d<<3)]     (ext)  */
  PZK_U_SHLADD = 21, /* dst = a + (b << imm16), wrapping: x + z * 2^k fused                */
  PZK_U_EXTRACT = 22, /* dst(U) = ((a >> s) & (2^n - 1)) << k with s = imm16 & 255, k = imm16 >> 8, n = b;
                         a is a U word, or (PZK_FLAG_NBASE) a plain 256-bit value in the F plane: materialises a
                         bit-field view of a word (see "views" below)                                          */
  PZK_V_LUT = 23,     /* packed truth table: 32 or 64 one-bit signals per record.  Bit l of dst =
                         imm16 >> (x0_l | x1_l << 1 | x2_l << 2 | x3_l << 3) & 1 with x_j = rotr_w(operand_j, rot_j),
                         w = 32 (operands masked to 32 bits first) or 64 (PZK_FLAG_W64); result masked with the
                         lane mask.  Extension record: {c, d, rot bytes, lane mask low}; with PZK_FLAG_W64 a second
                         extension record holds {lane mask high, 0, 0, 0}.  Unused operands are PZK_OPERAND_NONE.   */
  /* F class */
  PZK_F_CONST = 24, /* dst = fpool[a]                                       */
  PZK_F_ADD = 25,
  PZK_F_SUB = 26,
  PZK_F_MUL = 27,
  PZK_F_NEG = 28,
  PZK_F_INV = 29,    /* dst = a^-1, inv(0) = 0                                */
  PZK_F_FROM_U = 30, /* dst = Montgomery(a), a in U plane                     */
  PZK_F_SEL = 31,    /* dst = a(U) ? b : c  (ext)                             */
  PZK_F_EQ = 32,     /* dst(U) = (a == b)                                     */
  PZK_F_NE = 33,
  PZK_F_CSEL = 34,   /* dst = fpool[b + a(U)]                                 */
  PZK_F_FROM_I = 35, /* dst = Montgomery(a) for a signed 64-bit a             */
  PZK_BJJ_MUL8 = 37,  /* hint intrinsic for BabyjubjubBase8Multiplication
                         (/root/reference/circuits/lib/circuits/babyjubjub/curve.circom:143-171): the outputs of its
                         2n - 1 BabyjubjubAdd instances (n = 254 scalar bits, MSB first: one doubling and one addition
                         per bit, `(0,0)` standing for "no point" exactly as addZeroBabyjub does, :19-58) computed in
                         projective coordinates with ONE field inversion instead of two per addition.  They only
                         replace the `<--` hints of BabyjubjubAdd (:97,101); the rows `(1 + d tau) out === ...` are
                         still checked.  list = {n, pool a, pool d, pool Bx, pool By, scalar (F), out[2(2n-1)],
                         scratch[2(2n-1)]}                                                                        */
  PZK_CHECK_RANGE = 36, /* constraint row reduced to a range check by the bit-view prover: row dst (.r1cs index)
                           holds iff (a >> imm16) == 0; a is a U word or (PZK_FLAG_NBASE) a plain 256-bit value   */
  /* N class (plain 256-bit integers held in F slots) */
  PZK_N_FROM_F = 40, /* dst(N) = canonical(a)                                */
  PZK_F_FROM_N = 41, /* dst(F) = Montgomery(a mod p)                         */
  PZK_N_FROM_U = 42,
  PZK_N_BIT = 43,    /* dst(U) = (a >> b_imm) & 1                             */
  PZK_N_LOW = 44,    /* dst(U) = a mod 2^64                                   */
  PZK_N_SHR = 45,    /* dst(N) = a >> b(U)                                    */
  PZK_N_AND = 46,
  PZK_N_OR = 47,
  PZK_N_XOR = 48,
  PZK_N_DIV = 49, /* dst(N) = a / b (0 if b == 0)                           */
  PZK_N_MOD = 50,
  PZK_N_SLT = 51, /* dst(U) = signed_rep(a) < signed_rep(b)                 */
  PZK_N_SLE = 52,
  PZK_N_SHL = 53, /* dst(N) = ((a << b(U)) & (2^254-1)) mod p               */
  PZK_N_FITS = 54, /* dst(U) = (a < 2^64)                                   */
  PZK_N_EXTRACT = 55, /* dst(N) = ((a >> s) & (2^n - 1)) << k, fields as PZK_U_EXTRACT */
  /* constraint rows, fused into the op stream right after the op that defines their last wire.
   * header: imm16 = na, a = nb | nc << 16, b = number of 16-byte term records that follow,
   * dst = constraint index (.r1cs order).  Terms (A then B then C) are packed two per record:
   * {ref, coef}.  CHECK_INT: every wire is narrow and the compiler proved |A|,|B| < 2^63,
   * |C| < 2^126 - exact integer check; coef is an inline int32, or (PZK_TERM_COEF_LIST set in
   * ref) an offset into the list pool holding an int64.  CHECK_F: Montgomery arithmetic, coef is
   * an index into the coefficient pool. */
  PZK_CHECK_INT = 56,
  PZK_CHECK_F = 57,
  PZK_CHECK_I64 = 58, /* like CHECK_INT with |A|,|B|,|A*B|,|C| < 2^63 and inline int32 coefficients only:
                         plain wrapping 64-bit arithmetic decides the row */
  /* macro / control */
  PZK_MODINV = 59,    /* mod_inv intrinsic: list = {n, k, 0, a[k], p[k], unused, out[k]} (the BIGDIV layout
                         with m = 0); out = (a mod p)^-1 mod p for an odd prime p, 0 when p | a        */
  PZK_BIGDIV = 60,    /* long_div intrinsic, operands in the list pool       */
  PZK_ASSERT_NZ = 61, /* lane status |= ASSERT when a(U) == 0                */
  PZK_IN_U = 62,      /* dst(U) = input[a], range check: value < 2^imm16      */
  PZK_IN_F = 63,      /* dst(F) = Montgomery(input[a]); value must be < p    */
  /* Z class (exact wide integers, two's complement in the F plane) */
  PZK_Z_ADD = 64,    /* dst = a + b                (b may be a pool constant: PZK_FLAG_B_POOL, raw two's complement) */
  PZK_Z_SUB = 65,
  PZK_Z_MUL = 66,    /* dst = a * b mod 2^256; imm16 = la | lb << 4: when non-zero both operands are non-negative and
                        fit la / lb 64-bit limbs with la + lb <= 4 (schoolbook la x lb instead of the truncated 4 x 4) */
  PZK_Z_FROM_U = 67, /* dst(Z) = a(U)                                                                          */
  PZK_Z_FROM_I = 68, /* dst(Z) = a(I), sign extended                                                            */
  PZK_Z_CONST = 69,  /* dst(Z) = fpool[a] (raw two's complement)                                                */
  /* fused multiply-add: a product whose only reader is a sum is computed inside the sum's record (Poseidon's mix rows,
     the column sums of the RSA limb products).  Extension record {c, d, -, -}; a product that is a wire is a second
     result of the record (PZK_FLAG_DST2 / PZK_FLAG_DIG2): neither stored for nor re-read by the sum.               */
  PZK_F_MULADD = 70, /* dst = +-(a * b) +- c in Fr; b may be a pool constant (PZK_FLAG_B_POOL); imm16 bit 8 negates the
                        product, bit 9 negates c                                                                     */
  PZK_Z_MULADD = 71, /* dst = +-(a * b) +- c mod 2^256; imm16 = la | lb << 4 (as Z_MUL) | negate product << 8 |
                        negate c << 9                                                                                */
  PZK_OPCODE_MAX = 80
};

/* flags */
#define PZK_FLAG_B_IMM 1u  /* U ops: operand b is the 32-bit immediate in .b */
#define PZK_FLAG_B_POOL 2u /* F ops: operand b is fpool[.b]                  */
#define PZK_FLAG_EXT 4u    /* the following 16-byte record is an extension   */
#define PZK_FLAG_FAST 8u   /* set on U_ADD / U_MUL / U_AND / U_SHR / U_SHLADD (57 % of the ops of the passport circuits): the
                              evaluator takes them on a two-compare path in front of its opcode dispatch      */
#define PZK_FLAG_NBASE 16u /* U_EXTRACT / N_EXTRACT / CHECK_RANGE: operand a is a plain 256-bit value (F plane) */
#define PZK_FLAG_W64 32u   /* V_LUT: 64 lanes, a second extension record follows                              */
#define PZK_FLAG_DIG 128u  /* the op's result is a wire (or the word behind bit-field views that are wires): one 16-byte
                              digest descriptor follows the op's own records.  On disk {0, plane (0 = U, 1 = F), slot, 0};
                              the runtime rewrites it when the program is loaded to {rep | table << 4 | nbits << 8,
                              weight low, weight high, table offset} (pzk_kernels.cuh "fused digest"); an evaluator
                              that does not compute the digest skips it                                          */
#define PZK_FLAG_ZSRC 64u  /* N_FROM_F / F_FROM_N: operand a is a Z value: dst(N) = canonical(a) = a < 0 ? p + a : a,
                              dst(F) = Montgomery(a)                                                           */

#define PZK_FLAG_A_U 16u   /* Z_MUL / Z_MULADD: operand a is a U word (zero-extended), read where it is instead of    */
#define PZK_FLAG_B_U 1u    /* through a 256-bit Z_FROM_U copy; likewise operand b                                      */
#define PZK_FLAG_DST2 64u  /* F_MULADD / Z_MULADD: the product a * b is itself a wire whose only reader is this sum: it is a
                              second result of the record, extension word .d = its destination (encoded like .dst)    */
#define PZK_FLAG_DIG2 32u  /* F_MULADD / Z_MULADD with DST2: the digest descriptor of the product follows the extension
                              record, in front of the descriptor of the sum (PZK_FLAG_DIG) if there is one             */

typedef struct PzkOp {
  uint8_t opc;
  uint8_t flags;
  uint16_t imm16;
  uint32_t dst;
  uint32_t a;
  uint32_t b;
} PzkOp;

/* extension record that follows an op with PZK_FLAG_EXT */
typedef struct PzkOpExt {
  uint32_t c;
  uint32_t d;
  uint32_t e;
  uint32_t f;
} PzkOpExt;

/* ---- lane status bits ---------------------------------------------------- */
#define PZK_LANE_OK 0u
#define PZK_LANE_ASSERT 1u       /* runtime assert(...) on signals failed        */
#define PZK_LANE_CONSTRAINT 2u   /* some R1CS row failed ("Assert Failed.")      */
#define PZK_LANE_INPUT_RANGE 4u  /* input outside the declared range / >= p     */
#define PZK_LANE_BIGDIV_PRE 8u   /* long_div precondition violated (top limb 0) */
#define PZK_LANE_HINT 16u        /* a hint intrinsic met a zero denominator (BJJ_MUL8; unreachable on curve points) */

/* ---- constraint rows (slot addressed, per segment) ---------------------- */
/* term.ref : bits 30..31 = class (0 = U unsigned, 1 = I signed 64, 2 = F Montgomery, 3 = Z signed 256-bit integer),
 *            bits 0..29 = slot
 * term.coef: index into the coefficient pool                                */
typedef struct PzkTerm {
  uint32_t ref;
  uint32_t coef;
} PzkTerm;

#define PZK_REF_CLS(r) ((r) >> 30)
#define PZK_TERM_COEF_LIST 0x20000000u
#define PZK_TERM_CELL 0x10000000u /* the wire is read from the shared-memory operand cache */
#define PZK_TERM_BIT 0x08000000u  /* CHECK_F: the wire is a proven bit - its term is a conditional add */
#define PZK_REF_SLOT(r) ((r) & 0x07ffffffu)

/* Operand cache.  Every value is always stored to its global slot; in addition the compiler keeps
 * the values with the nearest next uses in a per-lane shared-memory cache of header.n_cells 8-byte
 * cells (an F value takes 4 consecutive cells).  Because the whole future of the straight-line
 * program is known, the allocation is Belady-optimal per segment and eviction is free.
 *   operand word : bit 31 set -> cell index in bits 0..15, else a global slot
 *   dst word     : bits 0..21 global slot, bits 22..30 cell + 1 (0 = not cached), bit 31 = the
 *                  global store is only needed when witnesses are exported (every later read
 *                  of the value hits the cache)                                                 */
#define PZK_OPERAND_CELL 0x80000000u
#define PZK_DST_SLOT(d) ((d) & 0x3fffffu)
#define PZK_DST_CELL(d) (((d) >> 22) & 0x1ffu)
#define PZK_DST_OPTIONAL 0x80000000u
#define PZK_REF_ZERO 0xFFFFFFFFu
#define PZK_REF_ONE 0xFFFFFFFEu
#define PZK_REF_ONE_LIST 0xFFFFFFFDu /* constant term whose int64 coefficient is in the list pool */
#define PZK_OPERAND_NONE 0xFFFFFFFFu

/* row kinds */
#define PZK_ROW_FIELD 0u /* generic: Montgomery arithmetic                  */
#define PZK_ROW_INT 1u   /* all wires U, coefficients small: exact int math */

typedef struct PzkRow {
  uint32_t term_off; /* first term (A terms, then B, then C)               */
  uint16_t na, nb;
  uint16_t nc;
  uint16_t kind;
  uint32_t index; /* constraint index in the .r1cs file                 */
} PzkRow;

typedef struct PzkSegment {
  uint64_t op_off, n_ops;   /* in 16-byte records (ext records included)    */
  uint64_t row_off, n_rows;
  uint64_t exp_off, n_exp;  /* export entries defined in this segment       */
} PzkSegment;

/* witness export entry: wire <- slot, or wire <- a view of one or more words.
 *
 * Views.  Most signals of the bit-sliced circuits (SHA, Num2Bits, running sums) are bit fields of a
 * word the program computes anyway: out[i] = (in >> i) & 1, sum[i] = in & (2^(i+1) - 1), (1 << i) * bit, the
 * 32 results of one packed PZK_V_LUT record ...  Such a signal has no op of its own; its export entry says
 * how to read it:
 *   class 3 (PZK_REF_VIEW)  : value = ((W >> s) & (2^n - 1)) << k, W = the U word in slot PZK_REF_SLOT(ref),
 *                             or with PZK_REF_VIEW_N a plain 256-bit value in the F plane; aux = s | n << 8 | k << 16
 *   ref == PZK_REF_TABVIEW  : value = table[idx], a truth-table function of up to 4 bits of words;
 *                             aux = offset into the list pool: {n, (ref_j, pos_j) x n, 2^n x int64 (lo, hi)};
 *                             negative entries are field negatives.                                           */
typedef struct PzkExport {
  uint32_t wire;
  uint32_t ref; /* class bits + slot; PZK_REF_ZERO / PZK_REF_ONE constants */
  uint32_t aux;
  uint32_t pad;
} PzkExport;
#define PZK_REF_VIEW_N 0x20000000u
#define PZK_REF_Z 0x10000000u /* export entry of class 2: the F-plane slot holds a Z value (canonical = v < 0 ? p + v : v) */
#define PZK_REF_TABVIEW 0xFFFFFFFCu

typedef struct PzkInput {
  uint32_t wire;   /* witness index of this main input element             */
  uint32_t bits;   /* declared width (0 = full field element)              */
} PzkInput;

/* file = header, then the sections in this order, each 16-byte aligned */
typedef struct PzkHeader {
  uint32_t magic, version;
  uint32_t n_wires;       /* witness length (wire 0 = constant 1)           */
  uint32_t n_pub_out, n_pub_in, n_prv_in;
  uint32_t n_constraints;
  uint32_t n_u_slots, n_f_slots;
  uint32_t n_segments;
  uint32_t n_fpool;       /* 32-byte Montgomery constants                   */
  uint32_t n_coef;        /* coefficient pool entries (3 x 32 bytes each)   */
  uint32_t n_inputs;      /* flattened main inputs                          */
  uint32_t n_list;        /* u32 operand-list pool                          */
  uint64_t n_op_records;
  uint64_t n_rows, n_terms, n_exports;
  uint64_t stat_u_ops, stat_f_mul, stat_f_inv, stat_f_other, stat_bigdiv;
  uint64_t reserved[4]; /* [0] = length of the JSON metadata, [1] = operand-cache cells per lane */
} PzkHeader;

/* coefficient pool entry */
typedef struct PzkCoef {
  uint64_t plain[4]; /* canonical value                                     */
  uint64_t mont[4];  /* c * R mod p    (multiplies F-class wires)            */
  uint64_t mont2[4]; /* c * R^2 mod p  (multiplies U-class wires)            */
} PzkCoef;

#ifdef __cplusplus
}
#endif
#endif
