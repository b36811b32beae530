// hsrle_slice_kernels.cuh -- the small kernels between the all-gathers of a multi-GPU encode (hsrle_slice.cuh).
#pragma once
#include "hsrle_enc_kernels.cuh"

namespace hsrle {

// after the scan: this rank's first message
static __global__ void k_enc_slice_msg1(const EncBufs B)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  SliceMsg &m = *B.msg;
  uint32_t *w = reinterpret_cast<uint32_t *>(&m);
  for (int i = 0; i < 64; i++) w[i] = 0;
  const EncScalars &sc = *B.sc;
  m.lo = B.sliceLo; m.hi = B.sliceHi; m.nStarts = sc.nStarts; m.nEnds = sc.nEnds;
  m.firstEnd = sc.nEnds ? B.runB[0] : 0u; m.status = sc.status;
}

// after all-gather #1: pair starts and ends across the cuts, set the assumed incoming state
static __global__ void k_enc_slice_link(const EncBufs B, const Spec sp)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  EncScalars &sc = *B.sc;
  for (uint32_t q = 0; q < B.world; q++) if (B.all[q].status != ST_OK && sc.status == ST_OK) sc.status = B.all[q].status;
  const SliceLink L = slice_link(B.all, (int)B.rank, (int)B.world);
  if (!L.ok && sc.status == ST_OK) sc.status = ST_BADARG;
  if (sc.status != ST_OK) { sc.nRuns = 0; sc.nSC = 0; sc.endShift = 0; }
  else
  {
    sc.endShift = L.endShift; sc.nRuns = L.nRuns; sc.nSC = (L.nRuns + E2_SCR - 1) / E2_SCR;
    if (L.borrow) B.runB[L.endShift + L.nRuns - 1] = L.borrowedEnd;
  }
  SliceState g; slice_guess_state(sp, (int)B.rank, B.sliceLo, g);
  *B.sliceIn = g;
}

// after an all-gather of outgoing states: replace a wrong assumption and arm the repair rounds
static __global__ void k_enc_slice_inject(const EncBufs B, const Spec sp)
{
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  EncScalars &sc = *B.sc;
  SliceState want; slice_incoming_state(sp, B.all, (int)B.rank, want);
  const SliceState cur = *B.sliceIn;
  const bool changed = sc.status == ST_OK && slice_state_differs(sp, want, cur);
  if (changed)
  {
    *B.sliceIn = want;
    if (sc.nSC > 0) { B.scIn[0] = want.st; if (sp.K) B.scLut[0] = want.lut; B.scDirty[0] = 1; }
  }
  sc.nDirty[0] = changed ? 1u : 0u; sc.firstDirty[0] = 0;
  for (int r = 1; r < E2_MAXROUNDS; r++) { sc.done[r] = 0; sc.nDirty[r] = 0; sc.firstDirty[r] = 0; }
  B.msg->changed = changed ? 1u : 0u;
}

// after all-gather #3: closing header + trailing literal, stream header (rank 0), result
static __global__ void __launch_bounds__(256) k_enc_slice_finish(const EncBufs B, const Spec sp)
{
  const EncScalars &sc = *B.sc;
  uint32_t status = sc.status;
  for (uint32_t q = 0; q < B.world; q++) if (B.all[q].status != ST_OK && status == ST_OK) status = B.all[q].status;
  SlicePlan P; slice_plan(sp, B.all, (int)B.rank, (int)B.world, B.n, P);
  const uint64_t pos = (uint64_t)B.outBase + B.all[B.rank].tokBytes;
  if (status == ST_OK && pos + P.closeLen + P.trailLen > B.cap) status = ST_OVERFLOW;
  // everybody's offsets (the same arithmetic on every rank)
  uint64_t before = 0, total = 0;
  for (uint32_t q = 0; q < B.world; q++)
  {
    SlicePlan Q; slice_plan(sp, B.all, (int)q, (int)B.world, B.n, Q);
    if (q < B.rank) before += Q.partLen;
    total += Q.partLen;
  }
  if (status == ST_OK && total >= 0xFFFFFFF0ull) status = ST_OVERFLOW;
  if (blockIdx.x == 0 && threadIdx.x == 0)
  {
    if (status == ST_OK)
    {
      for (uint32_t k = 0; k < P.closeLen; k++) B.out[pos + k] = P.closeHdr[k];
      if (B.rank == 0)
      { // stream header: uncompressed length, total stream bytes, (rle8: mode 0) -- src/rle8_extreme_cpu.h:91-97,341
        uint8_t *o = B.out + P.partStart;
        const uint32_t nn = B.n, tt = (uint32_t)total;
        for (int k = 0; k < 4; k++) { o[k] = (uint8_t)(nn >> (8 * k)); o[4 + k] = (uint8_t)(tt >> (8 * k)); }
        if (sp.hdr == 9) o[8] = 0;
      }
    }
    uint32_t *r = B.dResult;
    r[0] = status == ST_OK ? (uint32_t)P.partLen : 0u; r[1] = status; r[2] = P.partStart; r[3] = (uint32_t)before;
    r[4] = (uint32_t)total; r[5] = sc.nRuns; r[6] = sc.serialSC; r[7] = P.trailLen;
  }
  if (status != ST_OK || P.trailLen == 0) return;
  // trailing literal: one 4-KiB piece per warp and step
  constexpr uint32_t PIECE = 4096;
  const int lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), nWarps = gridDim.x * (blockDim.x >> 5);
  const uint8_t *src = B.in + P.trailSrc;
  uint8_t *dst = B.out + pos + P.closeLen;
  for (uint64_t off = (uint64_t)warp * PIECE; off < P.trailLen; off += (uint64_t)nWarps * PIECE)
    copy_bytes_warp(dst + off, src + off, (uint32_t)min((uint64_t)PIECE, (uint64_t)P.trailLen - off), lane);
}

} // namespace hsrle
