// hsrle_dec.cuh -- decoder pipeline of the B200 extreme-RLE codec.
//
// The stream has no sync markers: token k starts where token k-1 ends.  The decoder finds
// the token chain speculatively per super-chunk (SC, 16 KiB of stream) and resolves it with a merge:
//
//   D1  k_dec_map<codec>     per SC: parse a token at EVERY byte offset; a reverse sweep per 128-byte
//                            mini-block gives "where does a chain that starts here leave the mini-block"
//                            (exTab, kept for D3); a two-level in-place finalisation (warp-blocks, then the
//                            SC) turns that into "where does it leave the SC" for every offset (finTab,
//                            absolute positions).  The first DEC_WIN entries of an SC's finTab row are its
//                            windowed exit map: token chains re-enter the next SC within a few hundred
//                            bytes of its start unless a long literal spans the boundary.
//   D2a k_dec_compose        per segment of DEC_SEG SCs, in reverse SC order: sufExit[c][w] = where the chain
//                            that enters SC c at window offset w leaves the SEGMENT.  Chains that enter an SC
//                            beyond its window (after a long literal) take one finTab look-up instead.
//   D2b k_dec_resolve        one CTA: thread 0 chains the segments from the stream start (one look-up per
//                            segment, one more per long-literal entry), then one thread per segment walks
//                            its SCs forward through finTab and records every SC's true entry.
//   D3a k_dec_walk<codec>    per SC: mark the true chain (entry into every mini-block, hopping through D1's
//                            exit table), walk the tokens of every mini-block in parallel: output bytes, token
//                            count, symbol / LUT state transform of the SC.
//   D3s k_dec_scan           one CTA: exclusive scan of the SC aggregates (output offset, incoming symbol
//                            state); final validation (terminator seen, total == uncompressedLength).
//   D3b k_dec_expand<codec>  per SC: token records (output offset, literal source, run symbol), then one
//                            16-byte aligned output vector per thread and step (literal gather / period-W
//                            run fill).
//
// Reference behaviour restated (never copied): token parse src/rleX_extreme_cpu_decode.h:43-163,
// src/rleX_Xsl.h:580-784, src/rle8_extreme_cpu.h:1558-1632,2020-2087; header checks
// src/rle8_extreme_cpu.h:704-761, src/rleX_extreme_cpu.h:84-91.
#pragma once
#include "hsrle_core.cuh"
#include "hsrle_enc.cuh"   // ST_* status codes

namespace hsrle {

constexpr uint32_t DEC_SCB = 16384;       // stream bytes per super-chunk
constexpr uint32_t DEC_MB = 128;          // mini-block bytes (one thread sweeps one mini-block)
constexpr int DEC_T = DEC_SCB / DEC_MB;   // 128 threads per SC
constexpr uint32_t DEC_PAD = 32;          // readable bytes after the SC in shared memory (longest token header)
constexpr uint32_t DEC_WIN = 512;         // entry window of an SC
constexpr uint32_t DEC_SEG = 32;          // SCs per segment
constexpr int DEC_GROUP = DEC_T;          // look-back group (thread 0: inclusive prefix, threads 1..: aggregates)
constexpr uint32_t DEC_TOKCAP = 1024;     // token records kept in shared memory per expansion pass

constexpr uint32_t POS_END = 0xFFFFFFFFu; // chain reached the terminator
constexpr uint32_t POS_BAD = 0xFFFFFFFEu; // chain ran into an unparsable position
constexpr uint32_t POS_MISS = 0xFFFFFFFDu; // chain entered an SC outside its window (resolved by the slow path)
constexpr uint32_t POS_NONE = 0xFFFFFFFCu; // no token starts in this SC / segment
constexpr uint32_t POS_SPECIAL = 0xFFFFFFF0u;

// exit codes of the per-position table (u16, relative to the SC start)
constexpr uint32_t EX_FAR = 0x8000u;      // first code that is not a position inside [c0, c0 + 0x8000)
constexpr uint32_t EX_END = 0x8000u, EX_BAD = 0x8001u;
constexpr uint32_t EX_FARP = 0xC000u;     // | offset of the far-jumping token (its absolute exit: farTab)

struct DecScalars
{
  uint32_t n, clen, first, single, status;
  uint32_t singleSym;
  uint32_t endSeen;
  uint32_t nTok;
  uint32_t segTicket;                     // k_dec_chain: dynamic segment ids
  uint32_t ticket, done;                  // k_dec_emit: dynamic SC ids, CTAs finished
  uint32_t emitBad;                       // k_dec_emit met an unparsable token on the true chain
  uint32_t nHuge, nMed;                   // long operations handed to k_dec_big
  unsigned long long outTotal;            // output bytes of all tokens
};

// a long literal copy (kind 0: src = stream position of the first byte) or run fill (kind 1: src = output position
// where the run starts, sym = its first period) of nv whole 16-byte output vectors starting at vector v0
struct DecBigOp
{
  uint32_t v0, nv, src, kind;
  uint64_t sym;
};

// net effect of a token sequence on the K-entry LUT (decoder side): entry i of the table afterwards is either the
// incoming entry e[i] (e[i] < 8) or the symbol stored in the stream at position e[i] (>= 8: symbols follow a token
// head, which follows the stream header).  Symbols are fetched only when a transform is applied to a table.
struct LutXf
{
  uint32_t e[7];
  uint32_t pad;
};
HSRLE_HD void lutxf_identity(LutXf &x)
{
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++) x.e[i] = (uint32_t)i;
  x.pad = 0;
}
// (static indices only: see the LUT helpers in hsrle_core.cuh)
HSRLE_HD void lutxf_touch(LutXf &x, int K, int idx, uint32_t symPos)
{ // idx<K: move entry idx to front; idx==K: push the explicit symbol stored at symPos
  if (idx == 0) return;
  uint32_t e0 = symPos;
  int from = K - 1;
  if (idx != K)
  {
    from = idx;
    HSRLE_UNROLL
    for (int i = 1; i < 7; i++) if (i < K && i == idx) e0 = x.e[i];
  }
  HSRLE_UNROLL
  for (int j = 6; j > 0; j--) if (j < K && j <= from) x.e[j] = x.e[j - 1];
  x.e[0] = e0;
}
// key == j ? a : b.  On the device this is an opaque setp/selp pair: left to itself the compiler turns the unrolled
// select chains below into dynamically indexed local-memory arrays.
HSRLE_HD uint32_t sel_eq_u32(uint32_t key, uint32_t j, uint32_t a, uint32_t b)
{
#ifdef __CUDA_ARCH__
  uint32_t r;
  asm("{ .reg .pred p; setp.eq.u32 p, %1, %2; selp.u32 %0, %3, %4, p; }" : "=r"(r) : "r"(key), "r"(j), "r"(a), "r"(b));
  return r;
#else
  return key == j ? a : b;
#endif
}
HSRLE_HD uint64_t sel_eq_u64(uint32_t key, uint32_t j, uint64_t a, uint64_t b)
{
  return (uint64_t)sel_eq_u32(key, j, (uint32_t)a, (uint32_t)b) | ((uint64_t)sel_eq_u32(key, j, (uint32_t)(a >> 32), (uint32_t)(b >> 32)) << 32);
}
HSRLE_HD LutXf lutxf_compose(const LutXf &older, const LutXf &newer, int K)
{
  LutXf r; r.pad = 0;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++)
  {
    uint32_t v = newer.e[i];
    if (i < K)
    {
      HSRLE_UNROLL
      for (int j = 0; j < 7; j++) if (j < K) v = sel_eq_u32(newer.e[i], (uint32_t)j, older.e[j], v);
    }
    r.e[i] = v;
  }
  return r;
}
HSRLE_HD void lutxf_apply(const LutXf &x, int K, int W, const uint8_t *stream, const Lut &in, Lut &out)
{
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++)
  {
    if (i < K)
    {
      uint64_t v = 0;
      if (x.e[i] >= 8u) v = load_sym(stream + x.e[i], W);
      HSRLE_UNROLL
      for (int j = 0; j < 7; j++) if (j < K) v = sel_eq_u64(x.e[i], (uint32_t)j, in.s[j], v);
      out.s[i] = v;
    }
  }
}

// what a token sequence contributes to the decoder state: output bytes, token count, symbol register
template <int K> struct DecAgg
{
  uint64_t out;
  uint32_t ntok;
  uint32_t has;         // K == 0: the sequence set the symbol register
  uint64_t sym;
  LutXf xf;             // K > 0
};
template <int K> HSRLE_HD DecAgg<K> decagg_identity()
{
  DecAgg<K> a; a.out = 0; a.ntok = 0; a.has = 0; a.sym = 0;
  if (K) lutxf_identity(a.xf);
  return a;
}
template <int K> HSRLE_HD DecAgg<K> decagg_combine(const DecAgg<K> &older, const DecAgg<K> &newer)
{
  DecAgg<K> r;
  r.out = older.out + newer.out; r.ntok = older.ntok + newer.ntok;
  if (newer.has) { r.has = 1; r.sym = newer.sym; } else { r.has = older.has; r.sym = older.sym; }
  if (K) r.xf = lutxf_compose(older.xf, newer.xf, K);
  return r;
}

struct DecBufs
{
  const uint8_t *in; uint32_t inSize;
  uint8_t *out; uint32_t outSize;
  uint32_t nSC, nSeg;
  uint16_t *exTab;          // [nSC][DEC_SCB]   per-position mini-block exit tables (D1 -> D3)
  uint16_t *scTab;          // [nSC][DEC_SCB]   per-position SC exit codes (D1 -> D2)
  uint32_t *farTab;         // [nSC][DEC_SCB]   absolute exit of the far-jumping token at that position (sparse)
  uint32_t *winTab;         // [nSC][DEC_WIN]   absolute SC exits of the window positions
  uint32_t *scSkip;         // [nSC]  1: D1's scout found the SC jumped over by a true token (constant tables); cleared per call
  uint32_t *sufExit;        // [nSC][DEC_WIN]   exit of the segment when SC c is entered at window offset w
  uint32_t *flagSeg, *chainFlag;   // [nSeg] "rows published" / "chain position published" (zeroed per call)
  uint32_t *chainPos;       // [nSeg] first position of the true chain at or after the start of the segment (or its end code)
  uint32_t *scEntry;        // [nSC]  true entry (absolute stream position) or POS_NONE
  void *aggBuf, *incBuf;    // [nSC] DecAgg<K>: per-SC totals; inclusive prefixes (last SC of every look-back group)
  uint32_t *flagAgg, *flagInc;   // [nSC] "published" flags of the two (zeroed per call)
  DecBigOp *medList, *hugeList;  // long operations for k_dec_big
  DecScalars *sc;
  uint32_t *dResult;
};

// header check -- src/rle8_extreme_cpu.h:704-761, src/rleX_extreme_cpu.h:84-91 (every CTA evaluates it itself)
HSRLE_HD void dec_header(const Spec &sp, const uint8_t *in, uint32_t inSize, uint32_t outSize, DecScalars &sc)
{
  sc.status = ST_OK; sc.single = 0; sc.singleSym = 0; sc.endSeen = 0; sc.nTok = 0; sc.n = 0; sc.clen = 0; sc.first = sp.hdr;
  sc.segTicket = 0; sc.ticket = 0; sc.done = 0; sc.emitBad = 0; sc.nHuge = 0; sc.nMed = 0; sc.outTotal = 0;
  if (inSize < (uint32_t)sp.hdr) { sc.status = ST_BADARG; return; }
  sc.n = load32(in); sc.clen = load32(in + 4);
  if (sc.n > outSize || sc.clen > inSize || sc.clen < (uint32_t)sp.hdr || sc.clen >= POS_SPECIAL) { sc.status = ST_BADARG; return; }
  if (sp.hdr == 9)
  {
    const uint8_t mode = in[8];
    if (mode == 1) { if (sc.clen < 10) { sc.status = ST_BADARG; return; } sc.single = 1; sc.singleSym = in[9]; sc.first = 10; }
    else if (mode != 0) { sc.status = ST_BADARG; return; }
  }
}

// low 32 bits of the period-W pattern `sym` read at pattern offset d (0 <= d < W)
HSRLE_HD uint32_t pattern_word(uint64_t sym, int W, uint32_t d)
{
  if (W == 1) return (uint32_t)(sym & 0xFF) * 0x01010101u;
  uint64_t r = sym_rot(sym, W, d);
  if (W == 2) return (uint32_t)r | ((uint32_t)r << 16);
  if (W == 3) return (uint32_t)r | ((uint32_t)r << 24);
  return (uint32_t)r;
}

} // namespace hsrle
