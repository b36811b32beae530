// hsrle_stages.cuh -- the pipeline stages of the B200 extreme-RLE codec as per-item functions.
//
// Each kernel in hsrle_kernels.cu is a thin grid-stride wrapper around one of these; tests/sim drives
// the very same functions from plain host loops so the staged algorithm (candidate scan -> speculative
// emit automaton with verification -> size scan -> scatter; token-boundary maps -> hierarchical
// resolution -> token walk -> expansion) can be checked against the oracle without a GPU.
#pragma once
#include "hsrle_core.cuh"

namespace hsrle {

// ------------------------------------------------------------------------------------------------
// tunables
constexpr uint32_t ENC_VEC = 16;          // bytes per scan lane (one 16-byte load)
constexpr uint32_t ENC_TILE_VECS = 256;   // vectors per scan tile
constexpr uint32_t ENC_CH = 32;           // match-mask runs per automaton chunk
constexpr uint32_t BIG_COPY = 16384;      // literals at least this long go to the grid-wide copy kernel

constexpr uint32_t DEC_B1 = 4096;         // compressed bytes per boundary-map chunk
constexpr uint32_t DEC_G = 16;            // fan-out of the resolution hierarchy
constexpr uint32_t DEC_WIN = 512;         // entry window kept by the upper-level maps
constexpr int DEC_MAX_LEVELS = 6;
constexpr uint32_t DEC_TILE = 4096;       // output bytes per expansion tile

constexpr uint32_t POS_END = 0xFFFFFFFFu; // chain reached the terminator
constexpr uint32_t POS_BAD = 0xFFFFFFFEu; // chain ran into an unparsable position
constexpr uint32_t MAP_END = 0xE000u, MAP_BAD = 0xE001u, MAP_FAR = 0xF000u;

enum : uint32_t { ST_OK = 0, ST_OVERFLOW = 1, ST_BADSTREAM = 2, ST_BADARG = 3 };

// counters / scalars living in device memory
struct EncScalars
{
  uint32_t nRuns;
  uint32_t nChunks;
  uint32_t nDirty;
  uint32_t firstDirty;
  uint32_t nBig;
  uint32_t status;
  uint32_t total;        // final stream size
  uint32_t nTok;         // emitted tokens
  uint64_t tokBytes;     // sum over tokens of header+literal bytes
  uint32_t rounds;       // verification rounds that found work (diagnostics)
  uint32_t serialChunks; // chunks repaired by the serial fallback (diagnostics)
};

struct CopyDesc { uint32_t dst, src, len; };

struct EncBufs
{
  Spec sp;
  const uint8_t *in; uint32_t n;
  uint8_t *out; uint32_t cap;
  uint32_t nVec, nTiles;
  uint32_t maxRuns;
  uint32_t *tileS, *tileE;         // per scan tile: counts, then exclusive bases
  uint32_t *runA, *runB;           // match-mask runs [a,b)
  AutoState *sIn;                  // per chunk: incoming automaton state
  struct ChunkSum *cSum;           // per chunk: state written by the chunk
  Lut *lutIn; LutAgg *lutAgg;      // per chunk (LUT variants)
  uint64_t *cBytes; uint32_t *cTok; // per chunk: token bytes / tokens, then exclusive bases
  uint8_t *dirty;
  CopyDesc *copies;                // one per token (+1 trailing literal)
  uint32_t *bigList;
  EncScalars *sc;
};

// ------------------------------------------------------------------------------------------------
// E1: candidate scan.  For the 16 positions p0..p0+15 return bit masks of qualifying run starts
// (M[p]=1, M[p-1]=0, run length >= minM) and run ends (M[p]=0, M[p-1]=1, run length >= minM), where
// M[p] = (W <= p < n) && in[p]==in[p-W].  Only a 32-position window is needed (minM <= 8).
HSRLE_HD uint32_t cmpeq4_mask(uint32_t x, uint32_t y)
{
#ifdef __CUDA_ARCH__
  const uint32_t r = __vcmpeq4(x, y);
#else
  const uint32_t t = x ^ y;
  const uint32_t z = ~(((t & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | t) & 0x80808080u;
  const uint32_t r = (z >> 7) * 0xFFu;
#endif
  return ((r & 0x08040201u) * 0x01010101u) >> 24;
}
HSRLE_HD uint32_t funnel_l(uint32_t lo, uint32_t hi, uint32_t s)
{
#ifdef __CUDA_ARCH__
  return __funnelshift_l(lo, hi, s);
#else
  return s == 0 ? hi : ((hi << s) | (lo >> (32 - s)));
#endif
}

// w[0..11] = the 48 bytes in[p0-16 .. p0+32) as little-endian words (anything outside [0,n) may be garbage)
HSRLE_HD void mark_from_words(const Spec &sp, const uint32_t *w, uint32_t n, uint64_t p0, uint32_t &starts, uint32_t &ends)
{
  const int W = sp.W;
  const int ws = W >> 2, bs = (W & 3) * 8;
  uint32_t A = 0;
#pragma unroll
  for (int j = 2; j < 10; j++)
  {
    const uint32_t y = bs ? funnel_l(w[j - ws - 1], w[j - ws], bs) : w[j - ws];
    A |= cmpeq4_mask(w[j], y) << (4 * (j - 2));
  }
  // bit i <-> position p0-8+i ; valid iff W <= p < n
  const int64_t base = (int64_t)p0 - 8;
  int64_t lo = (int64_t)W - base; if (lo < 0) lo = 0;
  int64_t hi = (int64_t)n - base; if (hi > 32) hi = 32; if (hi < 0) hi = 0;
  uint32_t valid = 0;
  if (hi > lo) valid = (hi == 32 ? 0xFFFFFFFFu : ((1u << hi) - 1u)) & ~((lo == 0) ? 0u : ((1u << lo) - 1u));
  A &= valid;
  uint32_t ones = A;   // ones[i] = A[i..i+k-1] all set
  for (int j = 1; j < sp.minM; j++) ones &= (A >> j);
  const uint32_t s = A & ~(A << 1) & ones;
  const uint32_t e = ~A & (A << 1) & (ones << sp.minM);
  starts = (s >> 8) & 0xFFFFu;
  ends = (e >> 8) & 0xFFFFu;
}

// generic (byte-wise) loader of the 12 words, used by the simulator and by edge vectors
HSRLE_HD void mark_load_bytes(const uint8_t *in, uint32_t n, uint64_t p0, uint32_t *w)
{
  for (int j = 0; j < 12; j++)
  {
    uint32_t v = 0;
    for (int k = 0; k < 4; k++)
    {
      const int64_t p = (int64_t)p0 - 16 + j * 4 + k;
      if (p >= 0 && p < (int64_t)n) v |= (uint32_t)in[p] << (8 * k);
    }
    w[j] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// E2: emit automaton over one chunk of runs.
//
// The automaton state entering chunk c is "the most recent writer wins" over all earlier chunks:
//   last    <- end of the most recent emitted run        cursor  <- end of the most recent valid candidate
//   lastSym <- symbol of the most recent packed emission  LUT     <- K most recent distinct emitted symbols
// so, GIVEN every chunk's decisions, the incoming states are an exclusive scan of per-chunk summaries.
// Decisions depend on the incoming state, so the pipeline speculates (warm-up over the previous chunk),
// scans, verifies, re-runs the chunks whose input changed, and repeats; a fixed point of
// (run -> scan -> compare) is exactly the sequential result.  A one-thread serial pass is the
// always-correct fallback if the fixed point is not reached within the round budget.
struct ChunkSum
{
  uint32_t flags;       // EV_EMIT: `last` written | EV_VALID: `cursor` written | EV_SYMSET: `lastSym` written
  uint32_t last, cursor;
  uint64_t lastSym;
};
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

struct EncEmit          // where to put tokens (null => counting only)
{
  uint8_t *out;
  CopyDesc *copies;
  uint32_t *bigList; uint32_t *nBig;
  uint32_t outPos;      // stream position of the next token
  uint32_t tokIdx;      // descriptor slot of the next token
};

HSRLE_HD void enc_chunk_run(const EncBufs &B, uint32_t c, AutoState st, Lut lut, ChunkSum &sum, LutAgg *aggOut,
                            uint64_t &bytes, uint32_t &ntok, EncEmit *em)
{
  const uint32_t nRuns = B.sc->nRuns;
  const uint32_t lo = c * ENC_CH;
  uint32_t hi = lo + ENC_CH; if (hi > nRuns) hi = nRuns;
  LutAgg agg; agg.m = 0;
  bytes = 0; ntok = 0;
  uint32_t fl = 0;
  for (uint32_t j = lo; j < hi; j++)
  {
    uint32_t s, e; TokenHdr h;
    const uint32_t lastBefore = st.last;
    const uint32_t ev = enc_eval(B.sp, B.in, B.n, B.runA[j], B.runB[j], st, lut, B.sp.K ? &agg : nullptr, s, e, h);
    fl |= ev;
    if (!(ev & EV_EMIT)) continue;
    const uint32_t lit = s - lastBefore;
    bytes += h.len + lit; ntok++;
    if (em)
    {
      for (uint32_t k = 0; k < h.len; k++) em->out[em->outPos + k] = h.b[k];
      CopyDesc d; d.dst = em->outPos + h.len; d.src = lastBefore; d.len = lit;
      em->copies[em->tokIdx] = d;
      if (lit >= BIG_COPY)
      {
#ifdef __CUDA_ARCH__
        const uint32_t slot = atomicAdd(em->nBig, 1u);
#else
        const uint32_t slot = (*em->nBig)++;
#endif
        em->bigList[slot] = em->tokIdx;
      }
      em->outPos += h.len + lit; em->tokIdx++;
    }
  }
  sum.flags = fl; sum.last = st.last; sum.cursor = st.cursor; sum.lastSym = st.lastSym;
  if (aggOut) *aggOut = agg;
}

HSRLE_HD AutoState enc_initial_state() { AutoState s; s.cursor = 0; s.last = 0; s.lastSym = 0; return s; }

HSRLE_HD void enc_chunk_store(const EncBufs &B, uint32_t c, const ChunkSum &sum, const LutAgg &agg, uint64_t bytes, uint32_t ntok)
{
  B.cSum[c] = sum; B.cBytes[c] = bytes; B.cTok[c] = ntok; B.dirty[c] = 0;
  if (B.sp.K) B.lutAgg[c] = agg;
}

// iteration 0: speculate the incoming state of chunk c by warming up over chunk c-1 from a neutral
// state ("a run was just emitted right before the first candidate"), then run the chunk.
HSRLE_HD void enc_stage_auto_init(const EncBufs &B, uint32_t c)
{
  AutoState st = enc_initial_state();
  Lut lut; lut_init(lut, B.sp.W);
  if (c > 0)
  {
    const uint32_t lo = (c - 1) * ENC_CH, hi = c * ENC_CH;
    st.last = B.runA[lo] - B.sp.W; st.cursor = 0;
    for (uint32_t j = lo; j < hi; j++) { uint32_t s, e; TokenHdr h; enc_eval(B.sp, B.in, B.n, B.runA[j], B.runB[j], st, lut, nullptr, s, e, h); }
  }
  B.sIn[c] = st;
  if (B.sp.K) B.lutIn[c] = lut;
  ChunkSum sum; LutAgg agg; agg.m = 0; uint64_t bytes; uint32_t ntok;
  enc_chunk_run(B, c, st, lut, sum, B.sp.K ? &agg : nullptr, bytes, ntok, nullptr);
  enc_chunk_store(B, c, sum, agg, bytes, ntok);
}

// verification step for chunk c given the exact scan values of the current decisions
HSRLE_HD bool enc_stage_check(const EncBufs &B, uint32_t c, const AutoState &want, const Lut &wantLut)
{
  bool bad = false;
  if (B.sIn[c] != want) { B.sIn[c] = want; bad = true; }
  if (B.sp.K && !lut_equal(B.lutIn[c], wantLut, B.sp.K)) { B.lutIn[c] = wantLut; bad = true; }
  if (bad) B.dirty[c] = 1;
  return bad;
}

// sequential scan + verification over chunks [lo,hi) starting from the given running state (the GPU
// runs this per thread over a slice of chunks inside a block-wide scan)
HSRLE_HD uint32_t enc_scan_check_range(const EncBufs &B, uint32_t lo, uint32_t hi, AutoState st, Lut lut, uint32_t &firstDirty)
{
  uint32_t nd = 0;
  for (uint32_t c = lo; c < hi; c++)
  {
    if (enc_stage_check(B, c, st, lut)) { if (nd == 0 && c < firstDirty) firstDirty = c; nd++; }
    chunksum_apply(st, B.cSum[c]);
    if (B.sp.K) lut_apply(lut, B.sp.K, B.lutAgg[c]);
  }
  return nd;
}

HSRLE_HD void enc_stage_rerun(const EncBufs &B, uint32_t c)
{
  if (!B.dirty[c]) return;
  Lut lut; if (B.sp.K) lut = B.lutIn[c]; else lut_init(lut, B.sp.W);
  ChunkSum sum; LutAgg agg; agg.m = 0; uint64_t bytes; uint32_t ntok;
  enc_chunk_run(B, c, B.sIn[c], lut, sum, B.sp.K ? &agg : nullptr, bytes, ntok, nullptr);
  enc_chunk_store(B, c, sum, agg, bytes, ntok);
}

// exact sequential repair from chunk c0 to the end (one thread): the always-correct fallback.
// sIn[c0] / lutIn[c0] are exact (the scan over chunks < c0 is consistent by definition of c0).
HSRLE_HD void enc_stage_serial(const EncBufs &B, uint32_t c0)
{
  const uint32_t nChunks = B.sc->nChunks;
  if (c0 >= nChunks) return;
  AutoState st = B.sIn[c0];
  Lut lut; if (B.sp.K) lut = B.lutIn[c0]; else lut_init(lut, B.sp.W);
  for (uint32_t c = c0; c < nChunks; c++)
  {
    B.sIn[c] = st; if (B.sp.K) B.lutIn[c] = lut;
    ChunkSum sum; LutAgg agg; agg.m = 0; uint64_t bytes; uint32_t ntok;
    enc_chunk_run(B, c, st, lut, sum, B.sp.K ? &agg : nullptr, bytes, ntok, nullptr);
    enc_chunk_store(B, c, sum, agg, bytes, ntok);
    chunksum_apply(st, sum);
    if (B.sp.K) lut_apply(lut, B.sp.K, agg);
    B.sc->serialChunks++;
  }
}

// E4: final pass of a chunk with verified inputs: write token headers + literal copy descriptors.
// cBytes/cTok hold exclusive bases by now.
HSRLE_HD void enc_stage_emit(const EncBufs &B, uint32_t c)
{
  if (B.sc->status != ST_OK) return;
  EncEmit em;
  em.out = B.out; em.copies = B.copies; em.bigList = B.bigList; em.nBig = &B.sc->nBig;
  em.outPos = (uint32_t)(B.sp.hdr + B.cBytes[c]); em.tokIdx = B.cTok[c];
  Lut lut; if (B.sp.K) lut = B.lutIn[c]; else lut_init(lut, B.sp.W);
  ChunkSum sum; uint64_t bytes; uint32_t ntok;
  enc_chunk_run(B, c, B.sIn[c], lut, sum, nullptr, bytes, ntok, &em);
}

// E3 tail (one thread, after the scans): total size, capacity check, stream header, terminator and
// the trailing-literal descriptor.
HSRLE_HD void enc_stage_finish(const EncBufs &B)
{
  EncScalars &sc = *B.sc;
  const uint32_t nChunks = sc.nChunks;
  AutoState fin = enc_initial_state();
  if (nChunks) { fin = B.sIn[nChunks - 1]; chunksum_apply(fin, B.cSum[nChunks - 1]); }
  const uint32_t L = B.n - fin.last;
  TokenHdr h; enc_terminator(B.sp, L, h);
  const uint64_t total = (uint64_t)B.sp.hdr + sc.tokBytes + h.len + L;
  if (total > B.cap || total >= 0xFFFFFFF0ull) { sc.status = ST_OVERFLOW; sc.total = 0; return; }
  sc.total = (uint32_t)total;
  uint8_t *o = B.out;
  const uint32_t nn = B.n, tt = (uint32_t)total;
  for (int k = 0; k < 4; k++) { o[k] = (uint8_t)(nn >> (8 * k)); o[4 + k] = (uint8_t)(tt >> (8 * k)); }
  if (B.sp.hdr == 9) o[8] = 0;
  const uint32_t pos = (uint32_t)(B.sp.hdr + sc.tokBytes);
  for (uint32_t k = 0; k < h.len; k++) o[pos + k] = h.b[k];
  CopyDesc d; d.dst = pos + h.len; d.src = fin.last; d.len = L;
  B.copies[sc.nTok] = d;
  if (L >= BIG_COPY) { B.bigList[sc.nBig] = sc.nTok; sc.nBig++; }
}

// ================================================================================================
// DECODER
struct DecScalars
{
  uint32_t n, clen, first, single, status;
  uint32_t nChunks;
  uint32_t nTok;
  uint32_t endSeen;
  uint64_t outTotal;
  uint64_t singleSym;
};

struct LutXf            // net effect of a token sequence on the K-entry list
{
  uint64_t sym[7];
  int8_t ref[8];        // >=0: incoming entry ref[i]; -1: explicit sym[i]
};
HSRLE_HD void lutxf_identity(LutXf &x) { for (int i = 0; i < 7; i++) { x.ref[i] = (int8_t)i; x.sym[i] = 0; } x.ref[7] = 0; }
HSRLE_HD void lutxf_touch(LutXf &x, int K, int idx, uint64_t sym)
{ // idx<K: move entry idx to front; idx==K: push explicit symbol
  if (idx == 0) return;
  uint64_t s0; int8_t r0;
  if (idx == K) { s0 = sym; r0 = -1; idx = K - 1; } else { s0 = x.sym[idx]; r0 = x.ref[idx]; }
  for (int j = idx; j > 0; j--) { x.sym[j] = x.sym[j - 1]; x.ref[j] = x.ref[j - 1]; }
  x.sym[0] = s0; x.ref[0] = r0;
}
HSRLE_HD LutXf lutxf_compose(const LutXf &older, const LutXf &newer, int K)
{
  LutXf r; r.ref[7] = 0;
  for (int i = 0; i < 7; i++) { r.sym[i] = 0; r.ref[i] = (int8_t)i; }
  for (int i = 0; i < K; i++)
  {
    if (newer.ref[i] < 0) { r.sym[i] = newer.sym[i]; r.ref[i] = -1; }
    else { r.sym[i] = older.sym[newer.ref[i]]; r.ref[i] = older.ref[newer.ref[i]]; }
  }
  return r;
}
HSRLE_HD void lutxf_apply(const LutXf &x, int K, const Lut &in, Lut &out)
{
  for (int i = 0; i < K; i++) out.s[i] = x.ref[i] < 0 ? x.sym[i] : in.s[x.ref[i]];
}

struct DecBufs
{
  Spec sp;
  const uint8_t *in; uint32_t inSize;
  uint8_t *out; uint32_t outSize;
  uint16_t *map16;                       // per compressed byte: boundary-map code
  uint32_t *lmap[DEC_MAX_LEVELS + 1];    // level l>=1: [group*DEC_WIN + w] -> absolute exit
  uint32_t *lentry[DEC_MAX_LEVELS + 1];  // level l>=0 (0 = chunks): first token start >= group start
  int topLevel;
  uint32_t *cTok; uint64_t *cOut;        // per chunk counts, then exclusive bases
  uint64_t *cSym; uint8_t *cHasSym;      // packed: last explicit symbol of the chunk / carry-in after scan
  LutXf *cXf; Lut *cLutIn;               // LUT variants
  uint32_t *tOut, *tLitSrc, *tLitLen; uint64_t *tSym;   // token records (+1 sentinel)
  uint32_t maxTok;
  uint32_t *tileFirst;
  DecScalars *sc;
};

HSRLE_HD uint64_t dec_level_bytes(int lvl)
{
  uint64_t s = DEC_B1;
  for (int i = 0; i < lvl; i++) s *= DEC_G;
  return s;
}

// header check (one thread) -- src/rle8_extreme_cpu.h:704-761, src/rleX_extreme_cpu.h:84-91
HSRLE_HD void dec_stage_init(const DecBufs &D)
{
  DecScalars &sc = *D.sc;
  sc.status = ST_OK; sc.nTok = 0; sc.endSeen = 0; sc.outTotal = 0; sc.single = 0; sc.singleSym = 0; sc.nChunks = 0;
  if (D.inSize < (uint32_t)D.sp.hdr) { sc.status = ST_BADARG; return; }
  sc.n = load32(D.in); sc.clen = load32(D.in + 4); sc.first = D.sp.hdr;
  if (sc.n > D.outSize || sc.clen > D.inSize || sc.clen < (uint32_t)D.sp.hdr || sc.clen >= 0xFFFFFFF0u) { sc.status = ST_BADARG; return; }
  if (D.sp.hdr == 9)
  {
    const uint8_t mode = D.in[8];
    if (mode == 1) { if (sc.clen < 10) { sc.status = ST_BADARG; return; } sc.single = 1; sc.singleSym = D.in[9]; sc.first = 10; }
    else if (mode != 0) { sc.status = ST_BADARG; return; }
  }
  sc.nChunks = (sc.clen + DEC_B1 - 1) / DEC_B1;
}

// D1: boundary-map code of stream position p (chunk [c0,c1)), given the hop of every position.
// hop semantics: position q -> q + size(q) for a parsable non-final token.
struct HopInfo { uint32_t nxt; uint32_t kind; };   // kind 0: normal, 1: END (final token), 2: BAD
HSRLE_HD HopInfo dec_hop(const DecBufs &D, uint32_t p)
{
  const DecScalars &sc = *D.sc;
  HopInfo h; h.nxt = 0; h.kind = 2;
  if (p >= sc.clen) return h;
  Tok t; dec_parse(D.sp, sc.single != 0, D.in + p, (uint64_t)sc.clen - p, t);
  if (!t.valid) return h;
  if (t.last) { h.kind = 1; return h; }
  h.kind = 0; h.nxt = p + t.hdrLen + t.litLen;
  return h;
}
HSRLE_HD uint16_t dec_map_code(uint32_t c0, uint32_t c1, uint32_t lastTok, const HopInfo &h)
{
  if (h.kind == 1) return (uint16_t)MAP_END;
  if (h.kind == 2) return (uint16_t)MAP_BAD;
  const uint32_t rel = h.nxt - c1;
  if (rel < MAP_END) return (uint16_t)rel;
  return (uint16_t)(MAP_FAR | (lastTok - c0));
}
// one chunk hop through the stored map
HSRLE_HD uint32_t dec_advance_l1(const DecBufs &D, uint32_t x)
{
  const DecScalars &sc = *D.sc;
  if (x >= sc.clen) return POS_BAD;
  const uint32_t c0 = (x / DEC_B1) * DEC_B1;
  uint32_t c1 = c0 + DEC_B1; if (c1 > sc.clen || c1 < c0) c1 = sc.clen;
  const uint32_t code = D.map16[x];
  if (code < MAP_END) return c1 + code;
  if (code == MAP_END) return POS_END;
  if (code < MAP_FAR) return POS_BAD;
  const HopInfo h = dec_hop(D, c0 + (code & 0xFFFu));
  return h.kind == 0 ? h.nxt : (h.kind == 1 ? POS_END : POS_BAD);
}
// one step using the coarsest map (level <= maxLvl) whose entry window contains x
HSRLE_HD uint32_t dec_step(const DecBufs &D, uint32_t x, int maxLvl)
{
  for (int lvl = maxLvl; lvl >= 1; lvl--)
  {
    const uint64_t S = dec_level_bytes(lvl);
    const uint64_t off = (uint64_t)x % S;
    if (off < DEC_WIN) return D.lmap[lvl][(uint64_t)x / S * DEC_WIN + off];
  }
  return dec_advance_l1(D, x);
}
// D1b: up-sweep, level lvl >= 1: item = group*DEC_WIN + w
HSRLE_HD void dec_stage_up(const DecBufs &D, int lvl, uint64_t item)
{
  const DecScalars &sc = *D.sc;
  const uint64_t S = dec_level_bytes(lvl);
  const uint64_t g = item / DEC_WIN, w = item % DEC_WIN;
  const uint64_t start = g * S;
  uint64_t end = start + S; if (end > sc.clen) end = sc.clen;
  uint64_t x = start + w;
  if (x >= sc.clen) { D.lmap[lvl][item] = POS_BAD; return; }
  uint32_t guard = 0;
  while (x < end && ++guard < (1u << 22)) x = dec_step(D, (uint32_t)x, lvl - 1);
  D.lmap[lvl][item] = (uint32_t)x;
}
// D1c: top (one thread): entries of the groups at the top level
HSRLE_HD void dec_stage_top(const DecBufs &D)
{
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const int T = D.topLevel;
  const uint64_t S = dec_level_bytes(T);
  const uint32_t nG = (uint32_t)(((uint64_t)sc.clen + S - 1) / S);
  uint64_t x = sc.first;
  for (uint32_t g = 0; g < nG; g++)
  {
    D.lentry[T][g] = (uint32_t)x;
    uint64_t end = (uint64_t)(g + 1) * S; if (end > sc.clen) end = sc.clen;
    uint32_t guard = 0;
    while (x < end && ++guard < (1u << 22)) x = dec_step(D, (uint32_t)x, T);
  }
}
// D1d: down-sweep from level lvl (>=1) to lvl-1: item = group at level lvl
HSRLE_HD void dec_stage_down(const DecBufs &D, int lvl, uint32_t g)
{
  const DecScalars &sc = *D.sc;
  const uint64_t S = dec_level_bytes(lvl), Sc = dec_level_bytes(lvl - 1);
  uint64_t x = D.lentry[lvl][g];
  for (uint32_t k = 0; k < DEC_G; k++)
  {
    const uint64_t cs = (uint64_t)g * S + (uint64_t)k * Sc;
    if (cs >= sc.clen) break;
    uint64_t ce = cs + Sc; if (ce > sc.clen) ce = sc.clen;
    D.lentry[lvl - 1][(uint64_t)g * DEC_G + k] = (uint32_t)x;
    uint32_t guard = 0;
    while (x < ce && ++guard < (1u << 22)) x = dec_step(D, (uint32_t)x, lvl - 1);
  }
}

// D2: token walk of chunk c from its true entry.
struct DecSymState { uint64_t sym; Lut lut; };

template <bool EMIT>
HSRLE_HD void dec_chunk_walk(const DecBufs &D, uint32_t c)
{
  DecScalars &sc = *D.sc;
  const Spec &sp = D.sp;
  const uint32_t c0 = c * DEC_B1;
  uint32_t c1 = c0 + DEC_B1; if (c1 > sc.clen || c1 < c0) c1 = sc.clen;
  uint32_t x = D.lentry[0][c];
  uint32_t ntok = 0; uint64_t outBytes = 0;
  uint64_t sym = 0; bool hasSym = false;
  Lut lut; LutXf xf;
  uint32_t tokIdx = 0; uint64_t outPos = 0;
  if (EMIT)
  {
    if (sc.status != ST_OK) return;
    tokIdx = D.cTok[c]; outPos = D.cOut[c];
    if (sp.K) lut = D.cLutIn[c];
    else if (sc.single) sym = sc.singleSym;
    else sym = D.cSym[c];
  }
  else if (sp.K) lutxf_identity(xf);

  while (x < c1)
  {
    Tok t; dec_parse(sp, sc.single != 0, D.in + x, (uint64_t)sc.clen - x, t);
    if (!t.valid) { sc.status = ST_BADSTREAM; break; }
    uint64_t runSym = 0;
    if (sp.K)
    {
      const int idx = t.symKind == 0 ? sp.K : t.symKind - 2;
      const uint64_t ex = t.symKind == 0 ? load_sym(D.in + x + t.symOff, sp.W) : 0;
      if (EMIT) { if (idx == sp.K) lut_touch(lut, sp.K, sp.K, ex); else if (idx > 0) { const uint64_t v = lut.s[idx]; lut_touch(lut, sp.K, idx, v); } runSym = lut.s[0]; }
      else lutxf_touch(xf, sp.K, idx, ex);
    }
    else if (t.symKind == 0) { sym = load_sym(D.in + x + t.symOff, sp.W); hasSym = true; runSym = sym; }
    else runSym = sym;
    if (EMIT)
    {
      D.tOut[tokIdx] = (uint32_t)outPos; D.tLitSrc[tokIdx] = x + t.hdrLen; D.tLitLen[tokIdx] = t.litLen; D.tSym[tokIdx] = runSym;
      // expansion tiles whose first byte lies inside this token
      const uint64_t tend = outPos + t.litLen + t.runLen;
      if (tend > outPos)
      {
        uint64_t k = (outPos + DEC_TILE - 1) / DEC_TILE;
        for (; k * DEC_TILE < tend; k++) D.tileFirst[k] = tokIdx;
      }
      tokIdx++;
    }
    ntok++; outBytes += (uint64_t)t.litLen + t.runLen; outPos += (uint64_t)t.litLen + t.runLen;
    if (t.last) { if (!EMIT) sc.endSeen = 1; break; }
    x = x + t.hdrLen + t.litLen;
  }
  if (!EMIT)
  {
    D.cTok[c] = ntok; D.cOut[c] = outBytes;
    if (sp.K) D.cXf[c] = xf; else { D.cSym[c] = sym; D.cHasSym[c] = hasSym ? 1 : 0; }
  }
}

// D3: expansion of the 16 output bytes at v (v multiple of 16, v < n)
HSRLE_HD void dec_expand_vec(const DecBufs &D, uint64_t v, uint8_t *dst16)
{
  const DecScalars &sc = *D.sc;
  const int W = D.sp.W;
  const uint32_t n = sc.n;
  // token covering v: largest j with tOut[j] <= v among [tileFirst[k], tileFirst[k+1]]
  const uint64_t k = v / DEC_TILE;
  uint32_t lo = D.tileFirst[k];
  uint32_t hi = ((k + 1) * DEC_TILE < n) ? D.tileFirst[k + 1] : sc.nTok - 1;
  while (lo < hi)
  {
    const uint32_t mid = lo + (hi - lo + 1) / 2;
    if (D.tOut[mid] <= v) lo = mid; else hi = mid - 1;
  }
  uint32_t j = lo;
  uint64_t tStart = D.tOut[j], tNext = D.tOut[j + 1];
  uint32_t litLen = D.tLitLen[j];
  uint64_t vend = v + 16; if (vend > n) vend = n;
  for (uint64_t pos = v; pos < vend;)
  {
    while (pos >= tNext) { j++; tStart = tNext; tNext = D.tOut[j + 1]; litLen = D.tLitLen[j]; }
    const uint64_t litEnd = tStart + litLen;
    if (pos < litEnd)
    {
      uint64_t e = litEnd < vend ? litEnd : vend;
      const uint8_t *src = D.in + D.tLitSrc[j] + (pos - tStart);
      for (; pos < e; pos++) dst16[pos - v] = *src++;
    }
    else
    {
      uint64_t e = tNext < vend ? tNext : vend;
      const uint64_t sym = D.tSym[j];
      uint32_t ph = (uint32_t)((pos - litEnd) % W);
      for (; pos < e; pos++) { dst16[pos - v] = (uint8_t)(sym >> (8 * ph)); ph++; if (ph == (uint32_t)W) ph = 0; }
    }
  }
}

} // namespace hsrle
