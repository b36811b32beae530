// hsrle_dec_v1.cuh -- first-generation decoder stages (boundary maps -> hierarchical resolution -> token
// walk -> per-vector expansion).  Kept while the single-pass decoder replaces it stage by stage.
#pragma once
#include "../../hypersonic-rle-kit_b200/csrc/hsrle_core.cuh"
#include "../../hypersonic-rle-kit_b200/csrc/hsrle_enc.cuh"   // ST_* status codes

namespace hsrle {

constexpr uint32_t DEC_B1 = 4096;         // compressed bytes per boundary-map chunk
constexpr uint32_t DEC_G = 16;            // fan-out of the resolution hierarchy
constexpr uint32_t DEC_WIN = 512;         // entry window kept by the upper-level maps
constexpr int DEC_MAX_LEVELS = 6;
constexpr uint32_t DEC_TILE = 4096;       // output bytes per expansion tile

constexpr uint32_t POS_END = 0xFFFFFFFFu; // chain reached the terminator
constexpr uint32_t POS_BAD = 0xFFFFFFFEu; // chain ran into an unparsable position
constexpr uint32_t MAP_END = 0xE000u, MAP_BAD = 0xE001u, MAP_FAR = 0xF000u;

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
