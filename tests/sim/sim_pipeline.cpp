// TEST TOOL ONLY -- host-side stage simulator.  Drives the *same* per-item stage functions the CUDA
// kernels wrap (hypersonic-rle-kit_b200/csrc/hsrle_stages.cuh) from plain loops, so the staged
// algorithm can be compared with the oracle in a container without a GPU.  Never part of the
// product library: the product has no CPU path.
#include "../../hypersonic-rle-kit_b200/csrc/hsrle_stages.cuh"
#include <cstdio>
#include <cstring>
#include <vector>

using namespace hsrle;

static uint32_t g_stats[8];

extern "C" void sim_get_stats(uint32_t *o) { memcpy(o, g_stats, sizeof(g_stats)); }

extern "C" uint32_t sim_compress(int W, int align, int variant, const uint8_t *in, uint32_t n, uint8_t *out, uint32_t cap, int rounds)
{
  if (!in || !out || n == 0) return 0;
  EncBufs B; memset(&B, 0, sizeof(B));
  B.sp = make_spec(W, align, variant);
  B.in = in; B.n = n; B.out = out; B.cap = cap;
  B.nVec = (uint32_t)(((uint64_t)n + 1 + ENC_VEC - 1) / ENC_VEC);
  B.nTiles = (B.nVec + ENC_TILE_VECS - 1) / ENC_TILE_VECS;
  B.maxRuns = n / (B.sp.minM + 1) + 2;
  std::vector<uint32_t> tileS(B.nTiles + 1), tileE(B.nTiles + 1), runA(B.maxRuns), runB(B.maxRuns);
  const uint32_t maxChunks = B.maxRuns / ENC_CH + 2;
  std::vector<AutoState> sIn(maxChunks); std::vector<ChunkSum> cSum(maxChunks);
  std::vector<Lut> lutIn(maxChunks);
  std::vector<LutAgg> lutAgg(maxChunks);
  std::vector<uint64_t> cBytes(maxChunks); std::vector<uint32_t> cTok(maxChunks); std::vector<uint8_t> dirty(maxChunks);
  std::vector<CopyDesc> copies(B.maxRuns + 2); std::vector<uint32_t> bigList(B.maxRuns + 2);
  EncScalars sc; memset(&sc, 0, sizeof(sc));
  B.tileS = tileS.data(); B.tileE = tileE.data(); B.runA = runA.data(); B.runB = runB.data();
  B.sIn = sIn.data(); B.cSum = cSum.data(); B.lutIn = lutIn.data(); B.lutAgg = lutAgg.data();
  B.cBytes = cBytes.data(); B.cTok = cTok.data(); B.dirty = dirty.data(); B.copies = copies.data(); B.bigList = bigList.data(); B.sc = &sc;

  // E1 count + scan + write
  for (int pass = 0; pass < 2; pass++)
  {
    for (uint32_t t = 0; t < B.nTiles; t++)
    {
      uint32_t ns = 0, ne = 0;
      uint32_t bs = pass ? tileS[t] : 0, be = pass ? tileE[t] : 0;
      for (uint32_t v = t * ENC_TILE_VECS; v < (t + 1) * ENC_TILE_VECS && v < B.nVec; v++)
      {
        uint32_t w[12], s, e;
        mark_load_bytes(in, n, (uint64_t)v * ENC_VEC, w);
        mark_from_words(B.sp, w, n, (uint64_t)v * ENC_VEC, s, e);
        if (pass)
        {
          for (int i = 0; i < 16; i++) { if (s >> i & 1) runA[bs++] = v * ENC_VEC + i; if (e >> i & 1) runB[be++] = v * ENC_VEC + i; }
        }
        ns += __builtin_popcount(s); ne += __builtin_popcount(e);
      }
      if (!pass) { tileS[t] = ns; tileE[t] = ne; }
    }
    if (!pass)
    {
      uint32_t as = 0, ae = 0;
      for (uint32_t t = 0; t < B.nTiles; t++) { uint32_t s = tileS[t], e = tileE[t]; tileS[t] = as; tileE[t] = ae; as += s; ae += e; }
      if (as != ae) { fprintf(stderr, "sim: starts %u != ends %u\n", as, ae); return 0; }
      if (as > B.maxRuns) { fprintf(stderr, "sim: runs %u > max %u\n", as, B.maxRuns); return 0; }
      sc.nRuns = as; sc.nChunks = (as + ENC_CH - 1) / ENC_CH;
    }
  }
  const uint32_t nC = sc.nChunks;
  // E2
  for (uint32_t c = 0; c < nC; c++) enc_stage_auto_init(B, c);
  Lut lut0; lut_init(lut0, B.sp.W);
  uint32_t usedRounds = 0, first = 0xFFFFFFFFu;
  for (int r = 0;; r++)
  {
    first = 0xFFFFFFFFu;
    const uint32_t nd = enc_scan_check_range(B, 0, nC, enc_initial_state(), lut0, first);
    if (r < 5) g_stats[2 + r] = nd;
    if (!nd) break;
    if (r >= rounds) { enc_stage_serial(B, first); break; }
    usedRounds++;
    for (uint32_t c = 0; c < nC; c++) enc_stage_rerun(B, c);
  }
  g_stats[0] = usedRounds; g_stats[1] = sc.serialChunks; g_stats[7] = nC;
  // E3
  uint64_t ab = 0; uint32_t at = 0;
  for (uint32_t c = 0; c < nC; c++) { uint64_t b = cBytes[c]; uint32_t t = cTok[c]; cBytes[c] = ab; cTok[c] = at; ab += b; at += t; }
  sc.tokBytes = ab; sc.nTok = at; sc.status = ST_OK;
  enc_stage_finish(B);
  if (sc.status != ST_OK) return 0;
  // E4
  for (uint32_t c = 0; c < nC; c++) enc_stage_emit(B, c);
  // E5
  for (uint32_t i = 0; i <= sc.nTok; i++) memcpy(out + copies[i].dst, in + copies[i].src, copies[i].len);
  return sc.total;
}

extern "C" uint32_t sim_decompress(int W, int align, int variant, const uint8_t *in, uint32_t inSize, uint8_t *out, uint32_t outSize)
{
  if (!in || !out || inSize == 0 || outSize == 0) return 0;
  DecBufs D; memset(&D, 0, sizeof(D));
  D.sp = make_spec(W, align, variant);
  D.in = in; D.inSize = inSize; D.out = out; D.outSize = outSize;
  DecScalars sc; memset(&sc, 0, sizeof(sc)); D.sc = &sc;
  dec_stage_init(D);
  if (sc.status != ST_OK) return 0;
  const uint32_t nC = (inSize + DEC_B1 - 1) / DEC_B1;
  std::vector<uint16_t> map16((size_t)nC * DEC_B1);
  D.map16 = map16.data();
  // levels
  int T = 0; { uint64_t g = nC; while (g > DEC_G) { g = (g + DEC_G - 1) / DEC_G; T++; } }
  D.topLevel = T;
  std::vector<std::vector<uint32_t>> lmap(T + 1), lentry(T + 1);
  for (int l = 0; l <= T; l++)
  {
    const uint64_t S = dec_level_bytes(l);
    const uint64_t nG = ((uint64_t)inSize + S - 1) / S;
    if (l >= 1) { lmap[l].assign(nG * DEC_WIN, 0); D.lmap[l] = lmap[l].data(); }
    lentry[l].assign(nG + DEC_G, 0); D.lentry[l] = lentry[l].data();
  }
  // D1: per chunk boundary map (reverse sweep == fixpoint of the GPU's pointer doubling)
  for (uint32_t c = 0; c < sc.nChunks; c++)
  {
    const uint32_t c0 = c * DEC_B1; uint32_t c1 = c0 + DEC_B1; if (c1 > sc.clen) c1 = sc.clen;
    for (uint32_t p = c1; p-- > c0;)
    {
      const HopInfo h = dec_hop(D, p);
      if (h.kind == 0 && h.nxt < c1) map16[p] = map16[h.nxt];
      else map16[p] = dec_map_code(c0, c1, p, h);
    }
  }
  for (int l = 1; l <= T; l++)
  {
    const uint64_t S = dec_level_bytes(l);
    const uint64_t nG = ((uint64_t)sc.clen + S - 1) / S;
    for (uint64_t it = 0; it < nG * DEC_WIN; it++) dec_stage_up(D, l, it);
  }
  dec_stage_top(D);
  for (int l = T; l >= 1; l--)
  {
    const uint64_t S = dec_level_bytes(l);
    const uint64_t nG = ((uint64_t)sc.clen + S - 1) / S;
    for (uint64_t g = 0; g < nG; g++) dec_stage_down(D, l, (uint32_t)g);
  }
  // D2a
  std::vector<uint32_t> cTok(nC + 1); std::vector<uint64_t> cOut(nC + 1), cSym(nC + 1); std::vector<uint8_t> cHas(nC + 1);
  std::vector<LutXf> cXf(nC + 1); std::vector<Lut> cLutIn(nC + 1);
  D.cTok = cTok.data(); D.cOut = cOut.data(); D.cSym = cSym.data(); D.cHasSym = cHas.data(); D.cXf = cXf.data(); D.cLutIn = cLutIn.data();
  for (uint32_t c = 0; c < sc.nChunks; c++) dec_chunk_walk<false>(D, c);
  if (sc.status != ST_OK || !sc.endSeen) return 0;
  // scan
  {
    uint32_t at = 0; uint64_t ao = 0; uint64_t sym = 0;
    LutXf acc; lutxf_identity(acc); Lut init; lut_init(init, D.sp.W);
    for (uint32_t c = 0; c < sc.nChunks; c++)
    {
      const uint32_t t = cTok[c]; const uint64_t o = cOut[c];
      cTok[c] = at; cOut[c] = ao; at += t; ao += o;
      if (D.sp.K) { lutxf_apply(acc, D.sp.K, init, cLutIn[c]); acc = lutxf_compose(acc, cXf[c], D.sp.K); }
      else { const uint64_t s = cSym[c]; const bool h = cHas[c]; cSym[c] = sym; if (h) sym = s; }
    }
    sc.nTok = at; sc.outTotal = ao;
    if (ao != sc.n) return 0;
  }
  D.maxTok = sc.nTok + 1;
  std::vector<uint32_t> tOut(sc.nTok + 2), tLitSrc(sc.nTok + 2), tLitLen(sc.nTok + 2); std::vector<uint64_t> tSym(sc.nTok + 2);
  std::vector<uint32_t> tileFirst((size_t)sc.n / DEC_TILE + 2);
  D.tOut = tOut.data(); D.tLitSrc = tLitSrc.data(); D.tLitLen = tLitLen.data(); D.tSym = tSym.data(); D.tileFirst = tileFirst.data();
  for (uint32_t c = 0; c < sc.nChunks; c++) dec_chunk_walk<true>(D, c);
  tOut[sc.nTok] = sc.n; tLitLen[sc.nTok] = 0;
  if (sc.status != ST_OK) return 0;
  for (uint64_t v = 0; v < sc.n; v += 16)
  {
    uint8_t tmp[16];
    dec_expand_vec(D, v, tmp);
    memcpy(out + v, tmp, (size_t)((sc.n - v) < 16 ? (sc.n - v) : 16));
  }
  return sc.n;
}
