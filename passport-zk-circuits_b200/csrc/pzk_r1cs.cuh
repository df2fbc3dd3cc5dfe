// pzk_r1cs.cuh - the stand-alone R1CS checker: `snarkjs wtns check` / circom_tester's
// checkConstraints (/root/reference/test/automatisationTest.js:51) for explicit witnesses.
//
// Any iden3 .r1cs, any number of witnesses.  A full registerIdentity witness is 72 MB, so only
// ~1500 fit in HBM at once: the parallelism comes from the ROWS, not from the lanes.
//   grid  = (row chunks, groups of 128 lanes);  CTA = 4 warps, warp w owns lanes [32w, 32w+32)
//   witness layout  W[lane / 32][wire][limb][lane % 32]   (CANONICAL values, a wire of one warp = 1 KB)
// Wires stay canonical (the .wtns representation): a term with coefficient +-1 is a modular add, a term whose
// wire is 0 or 1 in every lane of the warp is a conditional add of the coefficient, any other term is one
// Montgomery product mont(c R, w) = c w, so linear combinations are accumulated as plain residues and no wire
// is ever converted; the product of a quadratic row costs two Montgomery products (a R, then a R * b / R), or
// none when A or B is a bit in every lane.  Loading explicit witnesses is a pure transpose, and the evaluator's
// export writes this layout directly (device-resident hand-off, pzk_api.cu).
// The A/B/C matrices are streamed through shared memory in tiles with TMA bulk copies
// (cp.async.bulk + mbarrier, double buffered): one elected thread issues the copy of the next
// tile's row headers and terms while the four warps consume the current one, so the stream is
// read once per CTA instead of once per warp.  The per-lane work is a gather of 32-byte wires with the next
// term's wire requested while the current one is consumed:
// HBM bound (algorithmic bytes = 32 B x terms x lanes + the matrix stream per CTA).
#pragma once
#include "fr_device.cuh"
#include "pzk_program.h"

namespace pzkd {

struct R1csTile {     // host-built: one shared-memory tile of the stream
  u32 row0, n_rows;   // rows [row0, row0 + n_rows)
  u32 term0, n_terms; // terms [term0, term0 + n_terms), term0 even (16-byte aligned source)
};

struct R1csParams {
  const PzkRow* rows;
  const PzkTerm* terms;
  const R1csTile* tiles;
  u32 n_tiles, tiles_per_chunk;
  const PzkCoef* coefs;
  const unsigned char* coef_kind;  // 0 general, 1 = +small, 2 = -small
  const u64* coef_mag;
  const u64* W;  // [lane/32][wire][limb][32], canonical
  u64 n_wires;
  u64 n_lanes;
  u32* status;
  unsigned long long* first_bad;
};

#define R1CS_TILE_ROWS 128
#define R1CS_TILE_TERMS 768
#define R1CS_SMEM_BYTES (2 * (R1CS_TILE_ROWS * 16 + R1CS_TILE_TERMS * 8) + 16)

__device__ __forceinline__ void mbar_init(u32 bar, u32 count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count)); }
__device__ __forceinline__ void mbar_expect_tx(u32 bar, u32 bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(u32 bar, u32 parity) {
  u32 ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(u32 bar, u32 parity) { while (!mbar_try_wait(bar, parity)) {} }
__device__ __forceinline__ void tma_load_1d(u32 dst, const void* src, u32 bytes, u32 bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void ldW(const u64* Wb, u64 wire, u64* v) {  // Wb: lane-block base + lane%32
  const u64* p = Wb + wire * 128;
  v[0] = p[0]; v[1] = p[32]; v[2] = p[64]; v[3] = p[96];
}
__device__ __forceinline__ bool is_bit256(const u64* w) { return (w[0] >> 1 | w[1] | w[2] | w[3]) == 0; }

// acc (plain residue) += coefficient ci * canonical wire w
__device__ __forceinline__ void r1cs_term(const R1csParams& p, u32 ci, const u64* w, u64* acc) {
  const u32 kind = __ldg(p.coef_kind + ci);
  if (kind && __ldg(p.coef_mag + ci) == 1) {
    if (kind == 1) fr_add(acc, acc, w); else fr_sub(acc, acc, w);
    return;
  }
  // most wires of the passport circuits are bits: when every lane of the warp holds 0 or 1 the
  // term is a conditional add of the coefficient, no multiplication
  u64 c[4];
  if (__all_sync(0xffffffffu, is_bit256(w))) {
    if (w[0]) { ldPool(reinterpret_cast<const u64*>(p.coefs), ci * 3, c); fr_add(acc, acc, c); }
  } else {
    u64 r[4];
    ldPool(reinterpret_cast<const u64*>(p.coefs), ci * 3 + 1, c);  // c * R
    fr_mul(r, c, w);
    fr_add(acc, acc, r);
  }
}

__global__ void __launch_bounds__(128, 6) r1cs_stream_kernel(R1csParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const u32 sbase = (u32)__cvta_generic_to_shared(smem);
  const u32 buf_bytes = R1CS_TILE_ROWS * 16 + R1CS_TILE_TERMS * 8;
  const u32 bar0 = sbase + 2 * buf_bytes;  // two mbarriers (8 bytes each)
  const u32 warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  const u64 lane = ((u64)blockIdx.y * 4 + warp) * 32 + lane32;
  const bool active = lane < p.n_lanes;
  // whole warps take the row loop together (it contains full-mask votes); lanes past the batch in
  // the last warp compute on padding and never write a verdict
  const bool warp_active = ((u64)blockIdx.y * 4 + warp) * 32 < p.n_lanes;
  const u64* Wb = p.W + ((u64)blockIdx.y * 4 + warp) * p.n_wires * 128 + lane32;
  const u32 tile_lo = blockIdx.x * p.tiles_per_chunk;
  const u32 tile_hi = min(p.n_tiles, tile_lo + p.tiles_per_chunk);
  if (tile_lo >= tile_hi) return;
  if (threadIdx.x == 0) {
    mbar_init(bar0, 1); mbar_init(bar0 + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  auto issue = [&](u32 tile, u32 slot) {
    const R1csTile tl = p.tiles[tile];
    const u32 rb = tl.n_rows * 16, tb = ((tl.n_terms + 1) & ~1u) * 8;
    const u32 dst = sbase + slot * buf_bytes;
    mbar_expect_tx(bar0 + 8 * slot, rb + tb);
    tma_load_1d(dst, p.rows + tl.row0, rb, bar0 + 8 * slot);
    tma_load_1d(dst + R1CS_TILE_ROWS * 16, p.terms + tl.term0, tb, bar0 + 8 * slot);
  };
  if (threadIdx.x == 0) issue(tile_lo, 0);
  const u64 R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
  unsigned long long bad = ~0ull;
  u32 phase0 = 0, phase1 = 0;
  for (u32 tile = tile_lo; tile < tile_hi; tile++) {
    const u32 slot = (tile - tile_lo) & 1;
    if (threadIdx.x == 0 && tile + 1 < tile_hi) issue(tile + 1, slot ^ 1);  // prefetch the next tile
    if (slot == 0) { mbar_wait(bar0, phase0); phase0 ^= 1; } else { mbar_wait(bar0 + 8, phase1); phase1 ^= 1; }
    const R1csTile tl = p.tiles[tile];
    const PzkRow* rows = reinterpret_cast<const PzkRow*>(smem + slot * buf_bytes);
    const PzkTerm* terms = reinterpret_cast<const PzkTerm*>(smem + slot * buf_bytes + R1CS_TILE_ROWS * 16);
    if (warp_active && tl.n_rows) {
      // the terms of the tile are one contiguous list: the wire of term t + 1 is requested before term t is
      // consumed (two gathers in flight per warp on top of the warps of the other CTAs of the SM)
      const u32 t_first = rows[0].term_off - tl.term0;
      const u32 t_end = tl.n_terms;
      u32 t = t_first;
      u64 wn[4];
      ldW(Wb, PZK_REF_SLOT(terms[t].ref), wn);
      for (u32 r = 0; r < tl.n_rows; r++) {
        const PzkRow row = rows[r];
        u64 acc[3][4];
#pragma unroll
        for (int q = 0; q < 3; q++) acc[q][0] = acc[q][1] = acc[q][2] = acc[q][3] = 0;
        const u32 na = row.na, nab = na + row.nb, nt = nab + row.nc;
        for (u32 k = 0; k < nt; k++, t++) {
          u64 w[4] = {wn[0], wn[1], wn[2], wn[3]};
          const u32 ci = terms[t].coef;
          if (t + 1 < t_end) ldW(Wb, PZK_REF_SLOT(terms[t + 1].ref), wn);
          if (k < na) r1cs_term(p, ci, w, acc[0]);
          else if (k < nab) r1cs_term(p, ci, w, acc[1]);
          else r1cs_term(p, ci, w, acc[2]);
        }
        bool ok;
        if (na == 0 || nab == na) ok = fr_is_zero(acc[2]);
        else {
          u64 ab[4];
          if (__all_sync(0xffffffffu, is_bit256(acc[0]))) {
#pragma unroll
            for (int j = 0; j < 4; j++) ab[j] = acc[0][0] ? acc[1][j] : 0;
          } else if (__all_sync(0xffffffffu, is_bit256(acc[1]))) {
#pragma unroll
            for (int j = 0; j < 4; j++) ab[j] = acc[1][0] ? acc[0][j] : 0;
          } else {
            u64 ar[4];
            fr_mul(ar, acc[0], R2);   // a R
            fr_mul(ab, ar, acc[1]);   // a b
          }
          ok = fr_eq(ab, acc[2]);
        }
        if (!ok && (unsigned long long)row.index < bad) bad = row.index;
      }
    }
    __syncthreads();  // everyone is done with this slot before it is refilled two tiles later
  }
  if (active && bad != ~0ull) {
    atomicOr(p.status + lane, PZK_LANE_CONSTRAINT);
    atomicMin(p.first_bad + lane, bad);
  }
}

// canonical AoS witnesses [lane][n_wires][4] -> canonical blocked planes [lane/32][wire][limb][32]
// (values >= p are flagged and reduced: snarkjs reads them modulo the prime)
__global__ void __launch_bounds__(128) load_witness_blocked_kernel(const u64* wit, u64 n_wires, u64 n_lanes, u64* W,
                                                                   u32* status) {
  const u64 lane = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= n_lanes) return;
  u64* Wb = W + (lane / 32) * n_wires * 128 + (lane % 32);
  for (u64 wv = blockIdx.y; wv < n_wires; wv += gridDim.y) {
    const ulonglong2* ip = reinterpret_cast<const ulonglong2*>(wit + (lane * n_wires + wv) * 4);
    ulonglong2 lo = ip[0], hi = ip[1];
    u64 v[4] = {lo.x, lo.y, hi.x, hi.y};
    if (geq_p(v)) { atomicOr(status + lane, PZK_LANE_INPUT_RANGE); reduce_p(v); }
    u64* q = Wb + wv * 128;
    q[0] = v[0]; q[32] = v[1]; q[64] = v[2]; q[96] = v[3];
  }
}

// hand-off from the evaluator: its export covers wires 1..n-1, wire 0 is the constant 1
__global__ void wire0_blocked_kernel(u64* W, u64 n_wires, u64 n_lanes) {
  const u64 lane = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= n_lanes) return;
  u64* q = W + (lane / 32) * n_wires * 128 + (lane % 32);
  q[0] = 1; q[32] = 0; q[64] = 0; q[96] = 0;
}

}  // namespace pzkd
