// hsrle_enc.cuh -- encoder pipeline of the B200 extreme-RLE codec (three phases; five launches).
//
//   E1  k_enc_scan<W,MINM>   one pass over the input: match mask M[p] = (in[p] == in[p-W]) from coalesced
//                            16-byte-per-lane loads, run boundaries by bit tricks on 32-bit windows of M,
//                            compaction of (start, end, first-period symbol) records with a decoupled
//                            look-back over tiles (single pass, the input is read once).
//   E2  k_enc_auto<codec>    the reference's per-run emit rules (hsrle_core.cuh: enc_eval_t) as a speculative
//                            automaton, round 0.  A super-chunk of 512 records is one WARP: every lane evaluates 16
//                            records from a warmed-up guess of the incoming state, a warp scan of the lane
//                            summaries yields the exact incoming states under the current decisions, mismatching
//                            lanes re-run -- a fixed point of (run -> scan -> compare) is exactly the sequential
//                            result inside the super-chunk.
//       k_enc_fix<codec>     all verify / repair rounds in ONE launch: CTA 0 scans the super-chunk summaries
//                            (state, token bytes), marks super-chunks whose assumed incoming state was wrong AND
//                            whose decisions depended on it (LUT codecs: ScQueries), re-runs them (with helper
//                            CTAs through tickets when there are many), verifies again; when nothing is dirty:
//                            token offsets, header, terminator.  After the round budget an exact sequential
//                            repair (still parallel inside a super-chunk) is the fallback.
//       k_enc_lut_stretch / k_enc_lut_walk (8-bit LUT codecs, hsrle_enc_lutwalk.cuh): the exact table at every
//                            super-chunk start before round 0, from a walk over the symbol changes of the records.
//   E3  k_enc_emit<codec>    token headers + literal scatter at the scanned stream offsets.
//       k_enc_copy_big       grid-wide copy of the few very long literals.
//
// Reference behaviour restated (never copied): scanners src/rle8_extreme_cpu.h:936-1099,
// src/rleX_extreme_cpu_encode.h:46-381; emit rules see hsrle_core.cuh.
#pragma once
#include "hsrle_core.cuh"

namespace hsrle {

enum : uint32_t { ST_OK = 0, ST_OVERFLOW = 1, ST_BADSTREAM = 2, ST_BADARG = 3 };

// ---------------------------------------------------------------- tunables
constexpr int E1_T = 256;                 // threads per scan macro-tile
constexpr int E1_WARPS = E1_T / 32;
constexpr int E1_STEPS = 32;              // at most this many 512-byte steps per warp (16 KiB contiguous per warp); the host picks
                                          // the count per call so that the tiles fill whole waves of resident CTAs (enc_scan_steps)
constexpr int E1_TILE_VECS = E1_T * E1_STEPS;   // 8192 vectors = 128 KiB of input per macro-tile at most
constexpr int E2_T = 128;                 // threads per super-chunk CTA
constexpr int E2_CH = 4;                  // records per thread
constexpr int E2_WARM = 12;               // records a thread warms its state guess up on
constexpr int E2_SCR = E2_T * E2_CH;      // records per super-chunk
constexpr int E2L_CH = 16;                // LUT codecs: records per lane -- one warp owns a whole super-chunk
static_assert(E2L_CH * 32 == E2_SCR && E2L_CH % E2_CH == 0, "a warp covers one super-chunk");
constexpr uint32_t SCF_SENS = 1;          // scFlags: a decision of the super-chunk depended on the incoming LUT
constexpr int FIX_T = 512;                // threads of the single-CTA verify / repair kernel (k_enc_fix)
constexpr int FIX_W = FIX_T / 32;
constexpr int FIX_LIST = 2048;            // dirty super-chunks listed per round (more: the warps scan the dirty flags)
constexpr int E2_MAXIT = 16;              // in-CTA fixed-point rounds before the in-CTA sequential pass
constexpr int E2_ROUNDS = 64;             // verify / repair rounds inside k_enc_fix before the exact sequential repair takes over.  A
                                          // wrong state guess travels one super-chunk per round: short-run streams read with a wider
                                          // symbol (rle24_byte_packed, rle32_sym: few tokens, the state passes through many super-chunks)
                                          // need a dozen rounds ...
constexpr int E2L1_ROUNDS = 512;          // ... and the 8-bit LUT codecs, whose table has long-range memory, 40-90
constexpr int E2_MAXROUNDS = 24;          // (size of the diagnostic arrays in EncScalars)
HSRLE_HDC int enc_rounds(int W, int K) { return (K != 0 && W == 1) ? E2L1_ROUNDS : E2_ROUNDS; }
constexpr int E2_NQ = 6;                  // LUT codecs: recorded table queries per super-chunk (more: always re-run)
constexpr uint32_t MED_COPY = 256;        // literals at least this long go to the grid-wide copy kernel (one warp each)
constexpr uint32_t BIG_COPY = 65536;      // ... and these are split over the whole grid

// ---------------------------------------------------------------- device-resident bookkeeping
struct EncScalars
{
  uint32_t tileTicket;                    // E1 dynamic tile ids
  uint32_t nRuns, nSC;
  uint32_t done[E2_MAXROUNDS];            // E2 per-round "CTAs finished" counters
  uint32_t nDirty[E2_MAXROUNDS];
  uint32_t firstDirty[E2_MAXROUNDS];
  uint32_t status;
  uint32_t total;                         // final stream size
  uint32_t nTok;
  uint32_t nBig, nMed;
  uint32_t serialSC;                      // super-chunks repaired sequentially (diagnostics)
  uint32_t innerSerial;                   // super-chunks that needed the in-CTA sequential pass (diagnostics)
  uint64_t tokBytes;                      // sum over tokens of header + literal bytes
  uint32_t fixCmd, fixRound, fixTicket, fixDone;   // k_enc_fix: helpers leave / grid round number / ticket counter / tickets completed
  uint32_t endShift;                      // slices: start record i pairs with end record i + endShift (hsrle_slice.cuh)
  uint32_t nStarts, nEnds;                // slices: records found by the scan
  uint32_t lwCount, lwOverflow, lwOk;     // 8-bit LUT codecs (hsrle_enc_lutwalk.cuh): stretch boundaries found / too many or too long / scGuess is valid
};

struct CopyDesc { uint32_t dst, src, len; };
struct LwDesc;

// LUT codecs: the decisions of a super-chunk that consulted the part of the table it inherited (marginal candidates whose
// symbol had not been emitted inside the super-chunk yet): symbol, how many first emissions (scFo) preceded it, and the
// answer it got.  The super-chunk's evaluation stays valid under another incoming table iff every answer stays the same.
struct ScQueries
{
  uint32_t n;                             // queries met (only the first E2_NQ are recorded)
  uint8_t known[8], hit[8];
  uint32_t pad;
  uint64_t sym[E2_NQ];
};

struct ChunkSum
{
  uint32_t flags;       // EV_EMIT: `last` written | EV_VALID: `cursor` written | EV_SYMSET: `lastSym` written
  uint32_t last, cursor;
  uint64_t lastSym;
};
HSRLE_HD ChunkSum chunksum_identity() { ChunkSum c; c.flags = 0; c.last = 0; c.cursor = 0; c.lastSym = 0; return c; }
HSRLE_HD void chunksum_apply(AutoState &st, const ChunkSum &c)
{
  if (c.flags & EV_EMIT) st.last = c.last;
  if (c.flags & EV_VALID) st.cursor = c.cursor;
  if (c.flags & EV_SYMSET) st.lastSym = c.lastSym;
}
HSRLE_HD ChunkSum chunksum_combine(const ChunkSum &older, const ChunkSum &newer)
{
  ChunkSum r = older;
  if (newer.flags & EV_EMIT) r.last = newer.last;
  if (newer.flags & EV_VALID) r.cursor = newer.cursor;
  if (newer.flags & EV_SYMSET) r.lastSym = newer.lastSym;
  r.flags |= newer.flags;
  return r;
}
HSRLE_HD AutoState enc_initial_state() { AutoState s; s.cursor = 0; s.last = 0; s.lastSym = 0; return s; }

// scan element: what a segment of records does to the automaton state, plus its token totals
template <int K, class AggT = LutAgg> struct SegSum
{
  ChunkSum cs;
  AggT agg;             // only meaningful for K > 0 (LutAgg, or the packed LutAggB for 8-bit symbols)
  uint64_t bytes;
  uint32_t ntok;
};

struct EncBufs
{
  const uint8_t *in; uint32_t n;
  uint8_t *out; uint32_t cap;
  uint32_t nVec, nTiles, lastVec;
  uint32_t scanSteps;                    // E1: 512-byte steps per warp (a tile is E1_T * scanSteps vectors)
  uint32_t maxRuns, maxSC;
  unsigned long long *tileStatus;        // E1 look-back: [2t] aggregate, [2t+1] inclusive prefix: flag(1) | ends(31) | starts(31)
  uint32_t *runA, *runB; void *runSym;   // records: mask run [a,b), first-period symbol (u32 if W <= 4 else u64)
  AutoState *cIn; Lut *cLut;             // per 4-record chunk: incoming state (written by E2, read by E3); for LUT codecs the
  uint8_t *cKnown;                       //   first cKnown entries of cLut are exact, the rest follow from scLut (enc_chunk_lut)
  AutoState *scIn; Lut *scLut;           // per super-chunk: assumed incoming state
  ChunkSum *scSum; LutAgg *scAgg;        // per super-chunk: summary under the current decisions
  uint64_t *scBytes; uint32_t *scTok;    // per super-chunk: token bytes (LUT codecs: without the symbol bytes of scFo) / tokens
  Lut *scFo; uint8_t *scFlags;           // LUT codecs: symbols in order of first emission (scAgg.m of them), SCF_* flags
  ScQueries *scQ;                        // LUT codecs: the table queries behind SCF_SENS
  uint64_t *scBase;                      // per super-chunk: exclusive token-byte offset
  uint8_t *scDirty;
  LwDesc *lwPool, *lwOrd;                // 8-bit LUT codecs: stretch-boundary descriptors as found / in record order,
  uint32_t *lwBlkOff, *lwBlkCnt;         //   where every block of LW_BLK records put its own,
  uint64_t *scGuess;                     //   and the walk's result: the (packed) table at every super-chunk start
  CopyDesc *bigList, *medList;
  EncScalars *sc;
  uint32_t *dResult;
  // one stream encoded by several GPUs (hsrle_slice.cuh); all zero / null for a whole-stream call
  uint32_t sliceMode, rank, world;
  uint32_t sliceLo, sliceHi;             // this rank's input range; literal bytes below sliceLo belong to earlier ranks
  uint32_t vecBase;                      // first 16-byte vector of the slice
  uint32_t outBase;                      // stream offset (whole stream) / local offset (slice) of the first token
  struct SliceState *sliceIn;            // incoming automaton state of the slice
  struct SliceMsg *msg;                  // this rank's message
  const struct SliceMsg *all;            // everybody's messages after the all-gather
};

// ---------------------------------------------------------------- E1 helpers (host+device)
// nibble of byte-equality bits of two words
HSRLE_HD uint32_t eq_nibble(uint32_t x, uint32_t y)
{
  const uint32_t t = x ^ y;
  const uint32_t nz = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t | 0x7F7F7F7Fu);   // 0x80 in every equal byte
  return (nz * 0x00204081u) >> 28;
}
HSRLE_HD uint32_t funnel_r(uint32_t lo, uint32_t hi, uint32_t s)
{
#ifdef __CUDA_ARCH__
  return __funnelshift_r(lo, hi, s);
#else
  return s == 0 ? lo : ((lo >> s) | (hi << (32 - s)));
#endif
}
// c = { bytes -8..-5, -4..-1, 0..3, 4..7, 8..11, 12..15 } relative to the vector start; returns the 16 raw
// equality bits (in[p] == in[p-W]) of the vector
template <int W> HSRLE_HD uint32_t m16_raw(const uint32_t *c)
{
  uint32_t m = 0;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 0; j < 4; j++)
  {
    const int off = 8 + 4 * j - W, wi = off >> 2, bs = (off & 3) * 8;
    const uint32_t y = bs ? funnel_r(c[wi], c[wi + 1], bs) : c[wi];
    m |= eq_nibble(c[2 + j], y) << (4 * j);
  }
  return m;
}
// validity: M[p] is defined for W <= p < n only
template <int W> HSRLE_HD uint32_t m16_valid(uint32_t v, uint32_t n)
{
  const uint64_t p0 = (uint64_t)v * 16;
  if (p0 >= n) return 0;
  const uint64_t rem = (uint64_t)n - p0;
  uint32_t m = rem >= 16 ? 0xFFFFu : ((1u << rem) - 1u);
  if (v == 0) m &= ~((1u << W) - 1u);
  return m;
}
// run boundaries of the 16 positions of a vector from the 32-bit window A (bit i <-> position p0-8+i)
template <int MINM> HSRLE_HD void m16_boundaries(uint32_t A, uint32_t &starts, uint32_t &ends)
{
  uint32_t ones = A;
#ifdef __CUDA_ARCH__
#pragma unroll
#endif
  for (int j = 1; j < MINM; j++) ones &= (A >> j);
  const uint32_t s = A & ~(A << 1) & ones;
  const uint32_t e = ~A & (A << 1) & (ones << MINM);
  starts = (s >> 8) & 0xFFFFu;
  ends = (e >> 8) & 0xFFFFu;
}

// ---------------------------------------------------------------- E2 helpers (host+device)
HSRLE_HD void lutagg_clear(LutAgg &a) { a.m = 0; }
HSRLE_HD void lutagg_clear(LutAggB &a) { a.m = 0; a.pad = 0; a.v = 0; }
template <int K, class AggT = LutAgg> HSRLE_HD SegSum<K, AggT> segsum_identity()
{
  SegSum<K, AggT> s; s.cs = chunksum_identity(); lutagg_clear(s.agg); s.bytes = 0; s.ntok = 0;
  return s;
}
template <int K, class AggT> HSRLE_HD SegSum<K, AggT> segsum_combine(const SegSum<K, AggT> &older, const SegSum<K, AggT> &newer)
{
  SegSum<K, AggT> r;
  r.cs = chunksum_combine(older.cs, newer.cs);
  if (K) r.agg = lutagg_combine(older.agg, newer.agg, K); else lutagg_clear(r.agg);
  r.bytes = older.bytes + newer.bytes; r.ntok = older.ntok + newer.ntok;
  return r;
}
template <int K, class AggT, class LutT> HSRLE_HD void segsum_apply(AutoState &st, LutT &lut, const SegSum<K, AggT> &s)
{
  chunksum_apply(st, s.cs);
  if (K) lut_apply(lut, K, s.agg);
}

// LUT codecs.  While fewer than K distinct symbols have been emitted inside a super-chunk, the table is
// (those symbols, most recent first) followed by what is left of the super-chunk's incoming table.  A super-chunk
// evaluated from a GUESSED incoming table is still exact -- decisions, state summary, aggregate -- unless one of its
// decisions was marginal (EV_MARG) for a symbol outside the known front part: only such "sensitive" super-chunks
// have to be re-run when the guess turns out wrong.  What does change with the incoming table is whether the first
// emission of each of the <= K symbols of scFo found its symbol in the table (no symbol bytes) or not (W bytes).
template <class LutT> HSRLE_HD uint32_t enc_fo_misses(const LutT &fo, uint32_t m, int K, const LutT &incoming)
{
  LutT l = incoming;
  uint32_t misses = 0;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++)
  {
    if (i < K && i < (int)m)
    {
      const uint64_t f = lut_entry(fo, i);
      const int idx = lut_find(l, K, f);
      if (idx == K) misses++;
      lut_touch(l, K, idx, f);
    }
  }
  return misses;
}
// would any recorded query of a super-chunk get another answer if its incoming table were `incoming`?
template <class LutT> HSRLE_HD bool enc_queries_differ(const ScQueries &q, const LutT &fo, uint32_t m, int K, const LutT &incoming)
{
  if (q.n > (uint32_t)E2_NQ) return true;
  LutT l = incoming;
  bool diff = false;
  HSRLE_UNROLL
  for (int k = 0; k < 7; k++)
  {
    if (k < K)
    { // l = incoming table after the first k first-emissions
      HSRLE_UNROLL
      for (int i = 0; i < E2_NQ; i++)
        if ((uint32_t)i < q.n && q.known[i] == (uint8_t)k) diff = diff || ((lut_find(l, K, q.sym[i]) != K) != (q.hit[i] != 0));
      if (k < (int)m) { const uint64_t f = lut_entry(fo, k); const int idx = lut_find(l, K, f); lut_touch(l, K, idx, f); }
    }
  }
  return diff;
}
// exact incoming table of a chunk: its `known` leading entries, then the super-chunk's exact incoming table
HSRLE_HD void enc_chunk_lut(Lut &chunkLut, uint32_t known, int K, const Lut &scIncoming)
{
  if ((int)known >= K) return;
  LutAgg a; a.m = known;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++) a.s[i] = chunkLut.s[i];
  Lut l = scIncoming;
  lut_apply(l, K, a);
  chunkLut = l;
}

// state guess for a chunk whose predecessor records are unknown: "a run was just emitted right before
// the first candidate", initial LUT
template <class LutT> HSRLE_HD void enc_neutral_state(const Spec &sp, uint32_t firstA, AutoState &st, LutT &lut)
{
  st = enc_initial_state();
  st.last = firstA - sp.W;
  lut_init(lut, sp.W);
}

} // namespace hsrle
