// pzk_r1cs.cuh - the stand-alone R1CS checker: `snarkjs wtns check` / circom_tester's
// checkConstraints (/root/reference/test/automatisationTest.js:51) for explicit witnesses.
//
// Any iden3 .r1cs, any number of witnesses.  A full registerIdentity witness is 72 MB, so only
// ~1000 fit in HBM at once: the parallelism comes from the ROWS, not from the lanes.
//   grid  = (row chunks, groups of 128 lanes);  CTA = 4 warps, warp w owns lanes [32w, 32w+32)
//   witness layout  F[lane / 32][wire][limb][lane % 32]   (Montgomery, a wire of one warp = 1 KB)
// The A/B/C matrices are streamed through shared memory in tiles with TMA bulk copies
// (cp.async.bulk + mbarrier, double buffered): one elected thread issues the copy of the next
// tile's row headers and terms while the four warps consume the current one, so the stream is
// read once per CTA instead of once per warp.  The per-lane work is a gather of 32-byte wires:
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
  const u64* F;  // [lane/32][wire][limb][32]
  u64 n_wires;
  u64 n_lanes;
  u32* status;
  unsigned long long* first_bad;
};

#define R1CS_TILE_ROWS 256
#define R1CS_TILE_TERMS 1536
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

__device__ __forceinline__ void ldW(const u64* Fb, u64 wire, u64* v) {  // Fb: lane-block base + lane%32
  const u64* p = Fb + wire * 128;
  v[0] = p[0]; v[1] = p[32]; v[2] = p[64]; v[3] = p[96];
}

__device__ __forceinline__ void r1cs_lin(const R1csParams& p, const PzkTerm* t, u32 n, const u64* Fb, u64* acc) {
  const u64 RONE[4] = {0xac96341c4ffffffbull, 0x36fc76959f60cd29ull, 0x666ea36f7879462eull, 0x0e0a77c19a07df2full};
  acc[0] = acc[1] = acc[2] = acc[3] = 0;
  for (u32 k = 0; k < n; k++) {
    const u32 ref = t[k].ref, ci = t[k].coef;
    u64 w[4];
    ldW(Fb, PZK_REF_SLOT(ref), w);
    const u32 kind = __ldg(p.coef_kind + ci);
    const u64 mag = __ldg(p.coef_mag + ci);
    if (kind && mag == 1) {
      if (kind == 1) fr_add(acc, acc, w); else fr_sub(acc, acc, w);
      continue;
    }
    // most wires of the passport circuits are bits: when every lane of the warp holds 0 or 1 the
    // term is a conditional add of the coefficient, no multiplication
    const bool is0 = fr_is_zero(w), is1 = fr_eq(w, RONE);
    u64 c[4];
    ldPool(reinterpret_cast<const u64*>(p.coefs), ci * 3 + 1, c);
    if (__all_sync(0xffffffffu, is0 || is1)) {
      if (is1) fr_add(acc, acc, c);
    } else {
      u64 r[4];
      fr_mul(r, c, w);
      fr_add(acc, acc, r);
    }
  }
}

__global__ void __launch_bounds__(128) r1cs_stream_kernel(R1csParams p) {
  extern __shared__ __align__(16) unsigned char smem[];
  const u32 sbase = (u32)__cvta_generic_to_shared(smem);
  const u32 buf_bytes = R1CS_TILE_ROWS * 16 + R1CS_TILE_TERMS * 8;
  const u32 bar0 = sbase + 2 * buf_bytes;  // two mbarriers (8 bytes each)
  const u32 warp = threadIdx.x >> 5, lane32 = threadIdx.x & 31;
  const u64 lane = ((u64)blockIdx.y * 4 + warp) * 32 + lane32;
  const bool active = lane < p.n_lanes;
  // whole warps take the row loop together (it contains a full-mask vote); lanes past the batch in
  // the last warp compute on padding and never write a verdict
  const bool warp_active = ((u64)blockIdx.y * 4 + warp) * 32 < p.n_lanes;
  const u64* Fb = p.F + ((u64)blockIdx.y * 4 + warp) * p.n_wires * 128 + lane32;
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
  unsigned long long bad = ~0ull;
  u32 phase0 = 0, phase1 = 0;
  for (u32 tile = tile_lo; tile < tile_hi; tile++) {
    const u32 slot = (tile - tile_lo) & 1;
    if (threadIdx.x == 0 && tile + 1 < tile_hi) issue(tile + 1, slot ^ 1);  // prefetch the next tile
    if (slot == 0) { mbar_wait(bar0, phase0); phase0 ^= 1; } else { mbar_wait(bar0 + 8, phase1); phase1 ^= 1; }
    const R1csTile tl = p.tiles[tile];
    const PzkRow* rows = reinterpret_cast<const PzkRow*>(smem + slot * buf_bytes);
    const PzkTerm* terms = reinterpret_cast<const PzkTerm*>(smem + slot * buf_bytes + R1CS_TILE_ROWS * 16);
    if (warp_active) {
      for (u32 r = 0; r < tl.n_rows; r++) {
        const PzkRow row = rows[r];
        const PzkTerm* t = terms + (row.term_off - tl.term0);
        u64 a[4], b[4], c[4];
        r1cs_lin(p, t + row.na + row.nb, row.nc, Fb, c);
        bool ok;
        if (row.na == 0 || row.nb == 0) ok = fr_is_zero(c);
        else {
          r1cs_lin(p, t, row.na, Fb, a);
          r1cs_lin(p, t + row.na, row.nb, Fb, b);
          u64 ab[4];
          fr_mul(ab, a, b);
          ok = fr_eq(ab, c);
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

// canonical AoS witnesses [lane][n_wires][4] -> Montgomery blocked planes [lane/32][wire][limb][32]
__global__ void __launch_bounds__(128) load_witness_blocked_kernel(const u64* wit, u64 n_wires, u64 n_lanes, u64* F,
                                                                   u32* status) {
  const u64 lane = (u64)blockIdx.x * blockDim.x + threadIdx.x;
  if (lane >= n_lanes) return;
  u64* Fb = F + (lane / 32) * n_wires * 128 + (lane % 32);
  for (u64 wv = blockIdx.y; wv < n_wires; wv += gridDim.y) {
    const ulonglong2* ip = reinterpret_cast<const ulonglong2*>(wit + (lane * n_wires + wv) * 4);
    ulonglong2 lo = ip[0], hi = ip[1];
    u64 v[4] = {lo.x, lo.y, hi.x, hi.y}, r[4];
    if (geq_p(v)) { atomicOr(status + lane, PZK_LANE_INPUT_RANGE); reduce_p(v); }
    fr_to_mont(r, v);
    u64* q = Fb + wv * 128;
    q[0] = r[0]; q[32] = r[1]; q[64] = r[2]; q[96] = r[3];
  }
}

}  // namespace pzkd
