// hsrle_slice.cuh -- one stream encoded by several GPUs: slice bookkeeping shared by the kernels and the host-side
// stage simulator (tests/sim, test tool only).
//
// The input of ONE reference-identical stream is cut into contiguous slices [lo_r, hi_r), one per rank (lo_r a multiple
// of the scan macro-tile, 128 KiB).  Everything is addressed in absolute input positions; a rank's device buffer holds
// its slice plus a 32-byte halo on both sides.  What crosses a slice boundary, and how it is repaired:
//
//   (1) a match-mask run that spans the cut: its start record lies in one rank, its end record in a later one.  Every
//       rank publishes {#starts, #ends, first end}; `slice_link` pairs them up again (the "boundary-run fix-up").
//   (2) the emit automaton's state (lastRLE, cursor, lastSymbol / LUT): every rank runs its automaton from an assumed
//       incoming state, publishes the outgoing one, `slice_inject` compares with the predecessor's and re-runs the
//       (few) super-chunks that depended on a wrong assumption -- repeated until no rank changed (at most `world` times).
//   (3) the literal that is still open at the end of a slice: its bytes follow the header of the NEXT emitted token
//       (SURVEY App. B.6), which a later rank produces.  A token header is therefore placed by the rank that holds the
//       input position where the token's literal begins; every rank's share of the stream is then one contiguous byte
//       range, and the per-rank sizes (exchanged with the last all-gather) give each rank its offset in the single stream.
//
// The three exchanges are all-gathers of one SliceMsg per rank (NCCL over NVLink in the product; gloo in the CPU tests).
#pragma once
#include "hsrle_enc.cuh"

namespace hsrle {

constexpr uint32_t SLICE_FRONT = 32;      // halo bytes before the slice in the rank's input buffer
constexpr uint32_t SLICE_TAIL = 32;       // halo bytes after it (zero / anything for the last rank)
constexpr uint32_t SLICE_PORCH = 32;      // local output offset of the rank's first token
constexpr uint32_t SLICE_ALIGN = 16u * 8192u;   // = 16 * E1_TILE_VECS: slice starts are multiples of the scan macro-tile

struct SliceState { AutoState st; Lut lut; };

struct SliceMsg
{
  // ---- written after the scan
  uint32_t lo, hi;                 // this rank's input range
  uint32_t nStarts, nEnds;         // run-start / run-end records found in the slice
  uint32_t firstEnd, status;       // first end record (if nEnds > 0)
  // ---- written by the automaton (every time it finishes)
  uint32_t changed, pad0;          // the last inject replaced this rank's incoming state
  AutoState out;                   // outgoing state
  Lut outLut;
  uint64_t tokBytes;               // header + own-slice literal bytes of this rank's tokens
  // ---- written by the emit pass
  uint32_t hasEmit;                // the slice emitted at least one token
  uint32_t firstHdrLen;            // header of its first token ...
  uint32_t firstS;                 // ... the input position where that token's run starts ...
  uint32_t firstLast;              // ... and where its literal starts (lastRLE before it)
  uint8_t firstHdr[24];
  uint32_t pad1[64 - 2 * 4 - 4 - 14 - 2 - 4 - 6];
};
static_assert(sizeof(SliceMsg) == 256, "SliceMsg is 64 words");

// -------------------------------------------------------------------------------------------- (1) boundary-run fix-up
struct SliceLink
{
  uint32_t endShift;     // my start record i pairs with my end record i + endShift
  uint32_t nRuns;        // records this rank evaluates (= its starts)
  uint32_t borrow;       // 1: the last start's end lies in a later slice ...
  uint32_t borrowedEnd;  // ... at this position
  uint32_t ok;
};
HSRLE_HD SliceLink slice_link(const SliceMsg *all, int rank, int world)
{
  SliceLink L; L.endShift = 0; L.nRuns = all[rank].nStarts; L.borrow = 0; L.borrowedEnd = 0; L.ok = 1;
  uint64_t s = 0, e = 0;
  for (int q = 0; q < rank; q++) { s += all[q].nStarts; e += all[q].nEnds; }
  if (s < e || s > e + 1) { L.ok = 0; return L; }
  L.endShift = (uint32_t)(s - e);                       // 1: a run is open where my slice begins; its end is my first end record
  const uint64_t have = all[rank].nEnds, need = (uint64_t)all[rank].nStarts + L.endShift;
  if (need == have + 1)
  {
    L.borrow = 1; L.ok = 0;
    for (int q = rank + 1; q < world; q++) if (all[q].nEnds > 0) { L.borrowedEnd = all[q].firstEnd; L.ok = 1; break; }
  }
  else if (need != have) L.ok = 0;
  return L;
}

// state a rank assumes before it knows its predecessor's: "a run was emitted right before the slice", initial LUT
HSRLE_HD void slice_guess_state(const Spec &sp, int rank, uint32_t lo, SliceState &g)
{
  g.st = enc_initial_state(); lut_init(g.lut, sp.W);
  // (W == 1 never reads or writes `cursor`: it must keep its initial value, or the difference to the true incoming
  //  state would be carried through every super-chunk of the slice and each of them re-evaluated)
  if (rank > 0) { g.st.cursor = sp.W == 1 ? 0u : lo; g.st.last = lo; }
}
// true incoming state of `rank` given everybody's current outgoing state
HSRLE_HD void slice_incoming_state(const Spec &sp, const SliceMsg *all, int rank, SliceState &g)
{
  if (rank == 0) { g.st = enc_initial_state(); lut_init(g.lut, sp.W); }
  else { g.st = all[rank - 1].out; g.lut = all[rank - 1].outLut; }
}
HSRLE_HD bool slice_state_differs(const Spec &sp, const SliceState &a, const SliceState &b)
{
  if (a.st != b.st) return true;
  return sp.K ? !lut_equal(a.lut, b.lut, sp.K) : false;
}

// own-slice part of the literal in[lastBefore, s): bytes below `floor` belong to earlier ranks
HSRLE_HD uint32_t slice_lit_src(uint32_t lastBefore, uint32_t floor) { return lastBefore > floor ? lastBefore : floor; }
HSRLE_HD uint32_t slice_lit_len(uint32_t lastBefore, uint32_t s, uint32_t floor)
{
  const uint32_t src = slice_lit_src(lastBefore, floor);
  return s > src ? s - src : 0u;
}

// -------------------------------------------------------------------------------------------- (3) placement
struct SlicePlan
{
  uint32_t closeLen;        // header this rank writes after its last token (0: none)
  uint8_t closeHdr[24];
  uint32_t trailSrc, trailLen;   // literal bytes of this rank that follow it
  uint32_t partStart;       // local offset of the rank's share of the stream
  uint64_t partLen;
};
// what rank q contributes to the stream, from the all-gathered messages
HSRLE_HD void slice_plan(const Spec &sp, const SliceMsg *all, int q, int world, uint32_t n, SlicePlan &P)
{
  const SliceMsg &m = all[q];
  const uint32_t lo = m.lo, hi = m.hi;
  const bool lastRank = q == world - 1;
  const uint32_t pend = m.out.last;                       // where the literal that is open after my tokens begins
  const bool holds = pend >= lo && (pend < hi || lastRank);
  // the token that closes it: the first token of the next rank that emitted, else the terminator
  uint32_t closeS = n; int closer = -1;
  for (int r = q + 1; r < world; r++) if (all[r].hasEmit) { closer = r; closeS = all[r].firstS; break; }
  P.closeLen = 0;
  if (holds)
  {
    if (closer >= 0) { P.closeLen = all[closer].firstHdrLen; for (uint32_t k = 0; k < 24; k++) P.closeHdr[k] = all[closer].firstHdr[k]; }
    else
    {
      TokenHdr th; enc_terminator(sp, n - pend, th);
      P.closeLen = th.len; for (uint32_t k = 0; k < 24; k++) P.closeHdr[k] = k < th.len ? th.b[k] : 0;
    }
  }
  const uint32_t t0 = pend > lo ? pend : lo, t1 = closeS < hi ? closeS : hi;
  P.trailSrc = t0; P.trailLen = t1 > t0 ? t1 - t0 : 0u;
  const bool foreign = m.hasEmit && m.firstLast < lo;     // my first token's header is placed by the rank that holds firstLast
  P.partStart = SLICE_PORCH + (foreign ? m.firstHdrLen : 0u);
  P.partLen = m.tokBytes - (foreign ? m.firstHdrLen : 0u) + P.closeLen + P.trailLen;
  if (q == 0) { P.partStart -= (uint32_t)sp.hdr; P.partLen += (uint32_t)sp.hdr; }
}

} // namespace hsrle
