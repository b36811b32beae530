// hsrle_dec.cuh -- decoder pipeline of the B200 extreme-RLE codec (two launches).
//
// The stream has no sync markers: token k starts where token k-1 ends.  The decoder finds the token chain
// speculatively per chunk (16 KiB of stream) and resolves it with a merge of windowed exit maps:
//
//   K1  k_dec_map<codec>   CTAs loop over chunks (grid = min(chunks, 8 per SM)).  Once per CTA a SCOUT follows the true chain
//                          from the stream start while its tokens jump over whole chunks: chunks jumped over get a flag and
//                          nothing else.  Per chunk: the image arrives in shared memory by a 1-D bulk copy (TMA); a token is
//                          parsed at EVERY byte offset; a blocked reverse sweep turns "where does the token at p end" into
//                          "where does the chain that starts at p leave its 64 / 128 / 256 / 512-byte block", then -- for the
//                          first DEC_WIN offsets of every 1-KiB sub-chunk -- "... its sub-chunk" (subMap: the rows K2 walks the
//                          sub-chunks from) and, for every offset, "... the chunk" (chunkTab, u16 codes).
//                          The last CTA of every 16-chunk segment composes the segment's exits in reverse chunk order: window
//                          rows (sufMap: chain entering chunk c at offset w < DEC_WINC leaves the SEGMENT at ...) for streams
//                          of short tokens, or the same for EVERY position (segTab) for streams of long literals with many
//                          segments -- decided per call from the first true tokens, identically by every composer.
//                          The last CTA of all is the RESOLVER: one thread follows the true chain from the stream start
//                          (window rows / chunk table or a direct parse after a long literal / one segTab look-up per
//                          segment) and leaves an ANCHOR in the chunks it lands in; one thread per segment derives every
//                          chunk's first true token start from the anchors (chunkEntry); the chunks that have one are listed
//                          in stream order (liveList).
//   K2  k_dec_emit<codec>  persistent CTAs take the LIVE chunks in order.  Thread 0 hops from the chunk's entry to every
//                          sub-chunk's entry (sub-chunk rows); 16 lanes walk the 16 sub-chunks, giving token starts, output
//                          bytes and symbol / LUT state of the chunk; a decoupled look-back over the live chunks yields the
//                          output offset and incoming symbol state; the tokens are expanded into a 16-KiB shared-memory image
//                          of the output (byte-exact, any alignment) that is flushed with aligned 16-byte stores.  Token parts
//                          that cover 64 KiB of whole output tiles and more become grid-wide operations: registered with one
//                          atomic (operation index and first piece number together), taken in 64-KiB pieces in one global
//                          order by every CTA that is out of chunks.
//
// Reference behaviour restated (never copied): token parse src/rleX_extreme_cpu_decode.h:43-163, src/rleX_Xsl.h:580-784,
// src/rle8_extreme_cpu.h:1558-1632,2020-2087; header checks src/rle8_extreme_cpu.h:704-761, src/rleX_extreme_cpu.h:84-91; the
// copy / fill kit src/rleX_extreme_common.h:32-312 is replaced by the shared-memory image and wide stores.
#pragma once
#include "hsrle_core.cuh"
#include "hsrle_enc.cuh"   // ST_* status codes

namespace hsrle {

constexpr uint32_t DEC_CB = 16384;        // chunk: stream bytes per CTA step
constexpr uint32_t DEC_SB = 1024;         // sub-chunk: the unit one lane walks in K2
constexpr int DEC_NSUB = (int)(DEC_CB / DEC_SB);
constexpr uint32_t DEC_WIN = 288;         // entry window of a sub-chunk: a token with a 1-byte range field ends < 8 + 11 + 254 bytes after its start
constexpr uint32_t DEC_WINC = 512;        // entry window of a chunk in the composed segment rows (entries beyond it take the chunk's own table)
constexpr uint32_t DEC_SEG = 16;          // chunks per segment
constexpr uint32_t DEC_IMG_PAD = 512;     // stream bytes loaded after the chunk (token heads and short literals that straddle its end)
constexpr uint32_t DEC_TILE = 16384;      // output image bytes per expansion step
constexpr uint32_t DEC_NSLOT = 1024;      // token records per expansion pass
constexpr uint32_t DEC_HUGE_TILES = 4;    // a token part covering at least this many whole tiles is a grid-wide operation
constexpr uint32_t DEC_BIG_PIECE = 65536; // bytes of a grid-wide operation one CTA takes at a time

constexpr uint32_t POS_END = 0xFFFFFFFFu; // chain reached the terminator
constexpr uint32_t POS_BAD = 0xFFFFFFFEu; // chain ran into an unparsable position
constexpr uint32_t POS_NONE = 0xFFFFFFFCu; // no token starts here
constexpr uint32_t POS_SPECIAL = 0xFFFFFFF0u;

// exit codes of the per-position tables (u16, relative to the chunk start)
constexpr uint32_t EX_FAR = 0x8000u;      // first code that is not a position inside [c0, c0 + 0x8000)
constexpr uint32_t EX_END = 0x8000u, EX_BAD = 0x8001u;
constexpr uint32_t EX_FARP = 0xC000u;     // | offset of the token that jumps beyond c0 + 0x7FFF (its exit: parse it again)

struct DecScalars                         // written once per call by CTA 0 of K1 (header check)
{
  uint32_t n, clen, first, single, status;
  uint32_t singleSym;
};
struct DecCounters                        // zeroed per call
{
  uint32_t segsDone;                      // K1: segments composed
  uint32_t chainBad;                      // K1 resolver: the true chain does not reach the terminator
  uint32_t ticket, chunksDone;            // K2: dynamic chunk ids, chunks finished
  uint32_t emitBad, endSeen;              // K2: unparsable token on the true chain / terminator seen
  uint32_t nBig;                          // grid-wide operations registered (for the result words)
  uint32_t bigTicket;                     // pieces of grid-wide operations handed out
  uint32_t nTok;
  uint32_t nLive;                         // K1 resolver: chunks in which a true token starts (entries of liveList)
  unsigned long long bigReg;              // operations registered << 40 | their pieces: one atomic numbers both consistently
  unsigned long long outTotal;            // output bytes of all tokens
};

// a grid-wide literal copy (kind 0: src = stream position of the byte that lands at output byte dst) or run fill (kind 1:
// src = output position where the run starts, sym = its first period) of `len` output bytes starting at the 16-byte aligned dst
struct DecBigOp
{
  uint64_t sym;
  uint32_t dst, len, src, kind;
  uint32_t pieceBase;                     // number of the operation's first piece in the global piece order
  uint32_t ready;                         // published
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

// what a token sequence contributes to the decoder state: output bytes, token count, symbol register / table
template <int K> struct DecAgg
{
  uint64_t out;
  uint32_t ntok;
  uint32_t symPos;      // K == 0: stream position of the last explicit symbol of the sequence (0: none)
  LutXf xf;             // K > 0
};
template <int K> HSRLE_HD DecAgg<K> decagg_identity()
{
  DecAgg<K> a; a.out = 0; a.ntok = 0; a.symPos = 0;
  if (K) lutxf_identity(a.xf);
  return a;
}
template <int K> HSRLE_HD DecAgg<K> decagg_combine(const DecAgg<K> &older, const DecAgg<K> &newer)
{
  DecAgg<K> r;
  r.out = older.out + newer.out; r.ntok = older.ntok + newer.ntok;
  r.symPos = newer.symPos ? newer.symPos : older.symPos;
  if (K) r.xf = lutxf_compose(older.xf, newer.xf, K);
  return r;
}

struct DecBufs
{
  const uint8_t *in; uint32_t inSize;
  uint8_t *out; uint32_t outSize;
  uint32_t nChunks, nSeg;   // capacities from inSize (the header's compressedLength may be smaller)
  uint32_t emitGrid;        // CTAs of K2
  DecScalars *sc;
  DecCounters *cnt;         // --- zeroed per call from here ...
  uint32_t *segCount;       // [nSeg]    chunks of the segment that published their rows
  uint32_t *anchorAt;       // [nChunks] position of the first true token start the resolver saw in the chunk (0: none)
  uint8_t *skipFlag;        // [nChunks] 1: the scout saw a true token jump over the whole chunk -- no table, no rows were written for it
  uint32_t *flagAgg;        // [nChunks] look-back: 1 = aggregate published, 2 = inclusive prefix published   ... to here
  uint16_t *chunkTab;       // [nChunks][DEC_CB]  exit code of the chunk for EVERY entry offset (read at the few offsets the chain visits)
  uint32_t *sufMap;         // dense streams: [nChunks][DEC_WINC] absolute exit of the SEGMENT (or the first out-of-window landing inside it) per window offset
  uint32_t *segTab;         // sparse streams: [nChunks * DEC_CB] per stream position: where the chain through it leaves its SEGMENT (absolute / POS_END / POS_BAD)
  uint16_t *subMap;         // [nChunks][DEC_NSUB][DEC_WIN] exit codes of the sub-chunks
  uint32_t *chunkEntry;     // [nChunks] first true token start of the chunk (POS_NONE: none) -- written by the resolver, read by K2
  uint32_t *liveList;       // [cnt.nLive] the chunks with an entry, in stream order: K2's ticket / look-back order
  void *aggBuf, *incBuf;    // [nChunks] DecAgg<K>: per-chunk totals / inclusive prefixes
  DecBigOp *bigList;
  uint32_t bigCap;
  uint32_t modeOverride;    // 0: K1's composers decide (rows / segment tables); 1: rows; 2: segment tables (tests: HSRLE_DEC_MODE)
  uint32_t *pieceOp;        // [pieceCap] (zeroed) piece number -> operation index + 1, written when the operation is registered
  uint32_t pieceCap;
  uint32_t *dResult;
  uint32_t *dbg;            // debugging aid (HSRLE_DEBUG): host-mapped stage markers, one word per CTA of K2; null otherwise
};

// header check -- src/rle8_extreme_cpu.h:704-761, src/rleX_extreme_cpu.h:84-91 (every CTA evaluates it itself)
HSRLE_HD void dec_header(const Spec &sp, const uint8_t *in, uint32_t inSize, uint32_t outSize, DecScalars &sc)
{
  sc.status = ST_OK; sc.single = 0; sc.singleSym = 0; sc.n = 0; sc.clen = 0; sc.first = sp.hdr;
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
