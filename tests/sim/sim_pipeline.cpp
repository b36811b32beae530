// TEST TOOL ONLY -- host-side stage simulator.  Drives the *same* host/device rule functions the CUDA
// kernels use (hypersonic-rle-kit_b200/csrc/hsrle_core.cuh, hsrle_enc.cuh, hsrle_dec_v1.cuh) from plain loops,
// reproducing the kernels' staged algorithms (candidate scan -> speculative automaton with scan/verify rounds
// and sequential repair -> emit; boundary maps -> resolution -> walk -> expansion) so they can be compared
// with the oracle in a container without a GPU.  Never part of the product library: the product has no CPU path.
#include "../../hypersonic-rle-kit_b200/csrc/hsrle_enc.cuh"
#include "hsrle_dec_v1.cuh"
#include <cstdio>
#include <cstring>
#include <vector>

using namespace hsrle;

static uint32_t g_stats[8];

extern "C" void sim_get_stats(uint32_t *o) { memcpy(o, g_stats, sizeof(g_stats)); }

// ---------------------------------------------------------------- E1: candidate scan (k_enc_scan)
template <int W, int MINM> static void sim_scan(const uint8_t *in, uint32_t n, std::vector<uint32_t> &runA, std::vector<uint32_t> &runB, std::vector<uint64_t> &runSym)
{
  const uint32_t nVec = (uint32_t)(((uint64_t)n + 1 + 15) / 16);
  std::vector<uint32_t> m16(nVec + 2, 0);      // m16[v+1] = mask of vector v
  for (uint32_t v = 0; v < nVec; v++)
  {
    uint32_t c[6];
    for (int j = 0; j < 6; j++)
    {
      uint32_t w = 0;
      for (int k = 0; k < 4; k++) { const int64_t p = (int64_t)v * 16 - 8 + j * 4 + k; if (p >= 0 && p < (int64_t)n) w |= (uint32_t)in[p] << (8 * k); }
      c[j] = w;
    }
    m16[v + 1] = m16_raw<W>(c) & m16_valid<W>(v, n);
  }
  for (uint32_t v = 0; v < nVec; v++)
  {
    const uint32_t A = (m16[v] >> 8) | (m16[v + 1] << 8) | (m16[v + 2] << 24);
    uint32_t s, e;
    m16_boundaries<MINM>(A, s, e);
    for (int i = 0; i < 16; i++)
    {
      if (s >> i & 1) { const uint32_t a = v * 16 + i; runA.push_back(a); runSym.push_back(load_sym(in + a - W, W)); }
      if (e >> i & 1) runB.push_back(v * 16 + i);
    }
  }
}

// ---------------------------------------------------------------- E2/E3: automaton + emit (k_enc_auto, k_enc_emit)
struct SimEnc
{
  Spec sp; int K;
  const uint8_t *in; uint32_t n;
  std::vector<uint32_t> runA, runB; std::vector<uint64_t> runSym;
  std::vector<AutoState> cIn; std::vector<Lut> cLut;

  template <class Seg> Seg eval_range(uint32_t j0, uint32_t j1, AutoState &st, Lut &lut, uint8_t *out, uint64_t *pos) const
  {
    Seg r; r.cs = chunksum_identity(); r.agg.m = 0; r.bytes = 0; r.ntok = 0;
    uint32_t fl = 0;
    for (uint32_t j = j0; j < j1; j++)
    {
      uint32_t s, e; TokenHdr h;
      const uint32_t lastBefore = st.last;
      const uint32_t ev = enc_eval(sp, runSym[j], n, runA[j], runB[j], st, lut, K ? &r.agg : nullptr, s, e, h);
      fl |= ev & EV_STATE_MASK;
      if (ev & EV_EMIT)
      {
        const uint32_t lit = s - lastBefore;
        if (out) { memcpy(out + *pos, h.b, h.len); memcpy(out + *pos + h.len, in + lastBefore, lit); *pos += h.len + lit; }
        r.bytes += h.len + lit; r.ntok++;
      }
    }
    r.cs.flags = fl; r.cs.last = st.last; r.cs.cursor = st.cursor; r.cs.lastSym = st.lastSym;
    return r;
  }
};

template <int KK> static uint32_t sim_encode_k(SimEnc &E, uint8_t *out, uint32_t cap, int rounds, int maxit)
{
  typedef SegSum<KK> Seg;
  const Spec &sp = E.sp;
  const uint32_t nRuns = (uint32_t)E.runB.size();
  const uint32_t nSC = (nRuns + E2_SCR - 1) / E2_SCR;
  const uint32_t nChunks = nSC * E2_T;
  E.cIn.assign(nChunks + 1, enc_initial_state()); E.cLut.resize(nChunks + 1);
  std::vector<AutoState> scIn(nSC + 1); std::vector<Lut> scLut(nSC + 1);
  std::vector<Seg> scSum(nSC + 1); std::vector<uint64_t> scBase(nSC + 1); std::vector<uint8_t> scDirty(nSC + 1, 0);
  uint32_t innerSerial = 0, serialSC = 0;

  auto process = [&](uint32_t s, bool given, const AutoState &gSt, const Lut &gLut) -> Seg
  {
    const uint32_t lo = s * E2_SCR, cnt = std::min<uint32_t>(E2_SCR, nRuns - lo);
    const uint32_t nT = (cnt + E2_CH - 1) / E2_CH;
    std::vector<AutoState> stIn(nT); std::vector<Lut> lutIn(nT); std::vector<Seg> mine(nT);
    for (uint32_t t = 0; t < nT; t++)
    {
      const uint32_t j0 = lo + t * E2_CH, j1 = std::min(j0 + E2_CH, lo + cnt);
      if (t == 0 && given) { stIn[t] = gSt; lutIn[t] = gLut; }
      else if (t == 0 && s == 0) { stIn[t] = enc_initial_state(); lut_init(lutIn[t], sp.W); }
      else
      {
        const uint32_t w0 = (s == 0) ? (uint32_t)std::max<int64_t>((int64_t)j0 - E2_WARM, 0) : j0 - E2_WARM;
        enc_neutral_state(sp, E.runA[w0], stIn[t], lutIn[t]);
        AutoState ws = stIn[t]; Lut wl = lutIn[t];
        (void)E.eval_range<Seg>(w0, j0, ws, wl, nullptr, nullptr);
        stIn[t] = ws; lutIn[t] = wl;
      }
      AutoState st = stIn[t]; Lut lut = lutIn[t];
      mine[t] = E.eval_range<Seg>(j0, j1, st, lut, nullptr, nullptr);
    }
    bool converged = false;
    for (int it = 0; it < maxit; it++)
    {
      std::vector<AutoState> want(nT); std::vector<Lut> wantLut(nT);
      AutoState run = stIn[0]; Lut runLut = lutIn[0];
      for (uint32_t t = 0; t < nT; t++) { want[t] = run; wantLut[t] = runLut; segsum_apply<KK>(run, runLut, mine[t]); }
      bool changed = false;
      for (uint32_t t = 1; t < nT; t++)
        if (want[t] != stIn[t] || (KK && !lut_equal(wantLut[t], lutIn[t], KK)))
        {
          stIn[t] = want[t]; lutIn[t] = wantLut[t]; changed = true;
          const uint32_t j0 = lo + t * E2_CH, j1 = std::min(j0 + E2_CH, lo + cnt);
          AutoState st = stIn[t]; Lut lut = lutIn[t];
          mine[t] = E.eval_range<Seg>(j0, j1, st, lut, nullptr, nullptr);
        }
      if (!changed) { converged = true; break; }
    }
    if (!converged)
    {
      innerSerial++;
      AutoState st = stIn[0]; Lut lut = lutIn[0];
      for (uint32_t t = 0; t < nT; t++)
      {
        stIn[t] = st; lutIn[t] = lut;
        const uint32_t j0 = lo + t * E2_CH, j1 = std::min(j0 + E2_CH, lo + cnt);
        mine[t] = E.eval_range<Seg>(j0, j1, st, lut, nullptr, nullptr);
      }
    }
    Seg total = segsum_identity<KK>();
    for (uint32_t t = 0; t < nT; t++) { E.cIn[s * E2_T + t] = stIn[t]; E.cLut[s * E2_T + t] = lutIn[t]; total = segsum_combine<KK>(total, mine[t]); }
    scSum[s] = total;
    if (!given) { scIn[s] = stIn[0]; scLut[s] = lutIn[0]; }
    return total;
  };

  AutoState d0 = enc_initial_state(); Lut dl; lut_init(dl, sp.W);
  AutoState fin = d0; Lut finLut = dl; uint64_t tokBytes = 0;
  bool clean = false;
  for (int r = 0; r < E2_ROUNDS && !clean; r++)
  {
    for (uint32_t s = 0; s < nSC; s++)
    {
      if (r == 0) process(s, false, d0, dl);
      else if (scDirty[s]) { const AutoState g = scIn[s]; const Lut gl = scLut[s]; process(s, true, g, gl); }
    }
    // scan_verify
    AutoState st = d0; Lut lut = dl; uint64_t bytes = 0; uint32_t nd = 0, first = 0xFFFFFFFFu;
    for (uint32_t s = 0; s < nSC; s++)
    {
      bool bad = false;
      if (scIn[s] != st) { scIn[s] = st; bad = true; }
      if (KK && !lut_equal(scLut[s], lut, KK)) { scLut[s] = lut; bad = true; }
      scDirty[s] = bad; if (bad) { if (!nd) first = s; nd++; }
      scBase[s] = bytes;
      segsum_apply<KK>(st, lut, scSum[s]); bytes += scSum[s].bytes;
    }
    if (r < 3) g_stats[2 + r] = nd;
    if (nd == 0) { clean = true; fin = st; finLut = lut; tokBytes = bytes; break; }
    if (r == E2_ROUNDS - 1 || r >= rounds)
    { // sequential repair
      AutoState run = scIn[first]; Lut runLut = scLut[first]; uint64_t runBytes = scBase[first];
      for (uint32_t s = first; s < nSC; s++)
      {
        const Seg tot = process(s, true, run, runLut);
        scIn[s] = run; scLut[s] = runLut; scBase[s] = runBytes; serialSC++;
        segsum_apply<KK>(run, runLut, tot); runBytes += tot.bytes;
      }
      fin = run; finLut = runLut; tokBytes = runBytes; clean = true;
    }
  }
  if (nSC == 0) { fin = d0; tokBytes = 0; }
  g_stats[0] = innerSerial; g_stats[1] = serialSC; g_stats[7] = nSC;
  // finish + emit
  const uint32_t L = E.n - fin.last;
  TokenHdr th; enc_terminator(sp, L, th);
  const uint64_t total = (uint64_t)sp.hdr + tokBytes + th.len + L;
  if (total > cap) return 0;
  const uint32_t nn = E.n, tt = (uint32_t)total;
  for (int k = 0; k < 4; k++) { out[k] = (uint8_t)(nn >> (8 * k)); out[4 + k] = (uint8_t)(tt >> (8 * k)); }
  if (sp.hdr == 9) out[8] = 0;
  for (uint32_t s = 0; s < nSC; s++)
  {
    const uint32_t lo = s * E2_SCR, cnt = std::min<uint32_t>(E2_SCR, nRuns - lo);
    uint64_t pos = (uint64_t)sp.hdr + scBase[s];
    for (uint32_t t = 0; t * E2_CH < cnt; t++)
    {
      const uint32_t j0 = lo + t * E2_CH, j1 = std::min(j0 + E2_CH, lo + cnt);
      AutoState st = E.cIn[s * E2_T + t]; Lut lut = E.cLut[s * E2_T + t];
      (void)E.eval_range<Seg>(j0, j1, st, lut, out, &pos);
    }
  }
  const uint64_t pos = (uint64_t)sp.hdr + tokBytes;
  memcpy(out + pos, th.b, th.len);
  memcpy(out + pos + th.len, E.in + fin.last, L);
  return tt;
}

extern "C" uint32_t sim_compress(int W, int align, int variant, const uint8_t *in, uint32_t n, uint8_t *out, uint32_t cap, int rounds, int maxit)
{
  if (!in || !out || n == 0) return 0;
  SimEnc E; E.sp = make_spec(W, align, variant); E.K = E.sp.K; E.in = in; E.n = n;
  memset(g_stats, 0, sizeof(g_stats));
#define SCAN(w, m) if (W == w && E.sp.minM == m) sim_scan<w, m>(in, n, E.runA, E.runB, E.runSym);
  SCAN(1, 5) SCAN(1, 2) SCAN(2, 2) SCAN(3, 3) SCAN(4, 4) SCAN(6, 6) SCAN(8, 8)
#undef SCAN
  if (E.runA.size() != E.runB.size()) { fprintf(stderr, "sim: starts %zu != ends %zu\n", E.runA.size(), E.runB.size()); return 0; }
  if (E.K == 3) return sim_encode_k<3>(E, out, cap, rounds, maxit);
  if (E.K == 7) return sim_encode_k<7>(E, out, cap, rounds, maxit);
  return sim_encode_k<0>(E, out, cap, rounds, maxit);
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

// ================================================================================================
// One stream encoded by several ranks (hsrle_slice.cuh): host-side twin of hsrle_slice_compress_phase, same job
// struct, host pointers.  The scan and the placement run the product's own HD functions (m16_*, slice_link,
// slice_lit_*, slice_plan, enc_eval); the automaton is threaded sequentially from the slice's incoming state (its
// speculative CTA form is covered by sim_compress above and by the GPU parity tests).
#include "../../hypersonic-rle-kit_b200/csrc/hsrle_slice.cuh"
#include "../../include/hsrle_b200.h"
#include <map>

struct SimSlice
{
  std::vector<uint32_t> runA, runB; std::vector<uint64_t> runSym;
  uint32_t endShift = 0, nRuns = 0, status = 0;
  SliceState in;
  uint64_t tokBytes = 0;
};
static std::map<const void *, SimSlice> g_slices;

template <int W, int MINM> static void sim_scan_slice(const uint8_t *in /* absolute */, uint32_t n, uint32_t lo, uint32_t hi, bool last, SimSlice &S)
{
  const uint32_t v0 = lo / 16;
  const uint32_t v1 = last ? (uint32_t)(((uint64_t)n + 1 + 15) / 16) : hi / 16;
  auto mask = [&](int64_t v) -> uint32_t
  { // the halo (32 bytes either side) is real input where it lies inside [0, n)
    if (v < 0) return 0;
    uint32_t c[6];
    for (int j = 0; j < 6; j++)
    {
      uint32_t w = 0;
      for (int k = 0; k < 4; k++) { const int64_t p = v * 16 - 8 + j * 4 + k; if (p >= 0 && p < (int64_t)n && p >= (int64_t)lo - 32 && p < (int64_t)hi + 32) w |= (uint32_t)in[p] << (8 * k); }
      c[j] = w;
    }
    return m16_raw<W>(c) & m16_valid<W>((uint32_t)v, n);
  };
  for (uint32_t v = v0; v < v1; v++)
  {
    const uint32_t A = (mask((int64_t)v - 1) >> 8) | (mask(v) << 8) | (mask((int64_t)v + 1) << 24);
    uint32_t s, e;
    m16_boundaries<MINM>(A, s, e);
    for (int i = 0; i < 16; i++)
    {
      if (s >> i & 1) { const uint32_t a = v * 16 + i; S.runA.push_back(a); S.runSym.push_back(load_sym(in + a - W, W)); }
      if (e >> i & 1) S.runB.push_back(v * 16 + i);
    }
  }
}

// sequential automaton over the slice's records; emits when out != nullptr
static void sim_slice_auto(const Spec &sp, const hsrle_slice_job *J, SimSlice &S, const uint8_t *in, SliceMsg &m, uint8_t *out)
{
  AutoState st = S.in.st; Lut lut = S.in.lut; LutAgg agg; agg.m = 0;
  uint64_t pos = SLICE_PORCH, bytes = 0;
  if (out) m.hasEmit = 0;
  for (uint32_t j = 0; j < S.nRuns; j++)
  {
    uint32_t s, e; TokenHdr h;
    const uint32_t lastBefore = st.last;
    const uint32_t ev = enc_eval(sp, S.runSym[j], J->n, S.runA[j], S.runB[j + S.endShift], st, lut, sp.K ? &agg : nullptr, s, e, h);
    if (!(ev & EV_EMIT)) continue;
    const uint32_t lit = slice_lit_len(lastBefore, s, J->lo), src = slice_lit_src(lastBefore, J->lo);
    if (out)
    {
      if (pos == SLICE_PORCH)
      {
        m.hasEmit = 1; m.firstHdrLen = h.len; m.firstS = s; m.firstLast = lastBefore;
        for (uint32_t k = 0; k < 24; k++) m.firstHdr[k] = k < h.len ? h.b[k] : 0;
      }
      memcpy(out + pos, h.b, h.len); memcpy(out + pos + h.len, in + src, lit);
    }
    pos += h.len + lit; bytes += h.len + lit;
  }
  S.tokBytes = bytes;
  m.out = st; if (sp.K) m.outLut = lut; else lut_init(m.outLut, sp.W);
  m.tokBytes = bytes; m.status = S.status;
}

extern "C" int sim_slice_compress_phase(const hsrle_slice_job *J, int phase, void *)
{
  const int wi = J->codec >> 3, ba = (J->codec >> 2) & 1, var = J->codec & 3;
  const Spec sp = make_spec(width_from_index(wi), ba, var);
  const int W = sp.W;
  SimSlice &S = g_slices[J->dWorkspace];
  const uint8_t *in = J->dIn + SLICE_FRONT - (ptrdiff_t)J->lo;
  SliceMsg &m = *reinterpret_cast<SliceMsg *>(J->dMsg);
  const SliceMsg *all = reinterpret_cast<const SliceMsg *>(J->dAll);
  const bool last = J->rank == J->world - 1;
  switch (phase)
  {
    case 0:
      S = SimSlice(); memset(&m, 0, sizeof(m));
#define SCAN(w, mm) if (W == w && sp.minM == mm) sim_scan_slice<w, mm>(in, J->n, J->lo, J->hi, last, S);
      SCAN(1, 5) SCAN(1, 2) SCAN(2, 2) SCAN(3, 3) SCAN(4, 4) SCAN(6, 6) SCAN(8, 8)
#undef SCAN
      m.lo = J->lo; m.hi = J->hi; m.nStarts = (uint32_t)S.runA.size(); m.nEnds = (uint32_t)S.runB.size();
      m.firstEnd = S.runB.empty() ? 0u : S.runB[0];
      return 0;
    case 1:
    {
      const SliceLink L = slice_link(all, J->rank, J->world);
      if (!L.ok) { S.status = ST_BADARG; S.nRuns = 0; }
      else
      {
        S.endShift = L.endShift; S.nRuns = L.nRuns;
        if (L.borrow) { S.runB.resize(std::max<size_t>(S.runB.size(), (size_t)L.endShift + L.nRuns)); S.runB[L.endShift + L.nRuns - 1] = L.borrowedEnd; }
      }
      slice_guess_state(sp, J->rank, J->lo, S.in);
      sim_slice_auto(sp, J, S, in, m, nullptr);
      return 0;
    }
    case 2:
    {
      SliceState want; slice_incoming_state(sp, all, J->rank, want);
      const bool changed = slice_state_differs(sp, want, S.in);
      if (changed) { S.in = want; sim_slice_auto(sp, J, S, in, m, nullptr); }
      m.changed = changed ? 1u : 0u;
      return 0;
    }
    case 3:
      sim_slice_auto(sp, J, S, in, m, J->dOut);
      return 0;
    case 4:
    {
      uint32_t status = S.status;
      for (int q = 0; q < J->world; q++) if (all[q].status != ST_OK && status == ST_OK) status = all[q].status;
      SlicePlan P; slice_plan(sp, all, J->rank, J->world, J->n, P);
      const uint64_t pos = (uint64_t)SLICE_PORCH + all[J->rank].tokBytes;
      if (status == ST_OK && pos + P.closeLen + P.trailLen > J->outCap) status = ST_OVERFLOW;
      uint64_t before = 0, total = 0;
      for (int q = 0; q < J->world; q++) { SlicePlan Q; slice_plan(sp, all, q, J->world, J->n, Q); if (q < J->rank) before += Q.partLen; total += Q.partLen; }
      if (status == ST_OK)
      {
        memcpy(J->dOut + pos, P.closeHdr, P.closeLen);
        memcpy(J->dOut + pos + P.closeLen, in + P.trailSrc, P.trailLen);
        if (J->rank == 0)
        {
          uint8_t *o = J->dOut + P.partStart;
          const uint32_t nn = J->n, tt = (uint32_t)total;
          for (int k = 0; k < 4; k++) { o[k] = (uint8_t)(nn >> (8 * k)); o[4 + k] = (uint8_t)(tt >> (8 * k)); }
          if (sp.hdr == 9) o[8] = 0;
        }
      }
      uint32_t *r = J->dResult;
      r[0] = status == ST_OK ? (uint32_t)P.partLen : 0u; r[1] = status; r[2] = P.partStart; r[3] = (uint32_t)before;
      r[4] = (uint32_t)total; r[5] = S.nRuns; r[6] = 0; r[7] = P.trailLen;
      g_slices.erase(J->dWorkspace);
      return 0;
    }
  }
  return 1;
}
