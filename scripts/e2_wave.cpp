// e2_wave.cpp (wave variant of e2_convergence.cpp) -- analysis tool (not product code): how many grid-level rounds does the encoder automaton
// (E2) need on a given input, under (a) the exact-state dirty criterion and (b) the sensitivity criterion?
// build: g++ -O2 -std=c++17 -I hypersonic-rle-kit_b200/csrc -o /tmp/e2conv scripts/e2_convergence.cpp
// usage: /tmp/e2conv file W byteAlign variant [scr]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "hsrle_enc.cuh"
using namespace hsrle;

struct Rec { uint32_t a, b; uint64_t sym; };

struct ScOut
{
  ChunkSum cs; LutAgg agg; uint64_t bytes; uint32_t ntok;
  bool sens;          // some decision depended on the incoming part of the LUT
};

static Spec sp;
static uint32_t n;
static std::vector<Rec> recs;

// evaluate one SC sequentially from (st, lut)
static ScOut run_sc(size_t lo, size_t hi, AutoState st, Lut lut)
{
  ScOut r; r.cs = chunksum_identity(); r.agg.m = 0; r.bytes = 0; r.ntok = 0; r.sens = false;
  uint32_t fl = 0;
  int known = 0;
  for (size_t j = lo; j < hi; j++)
  {
    uint32_t s, e; CountSink h;
    const uint32_t lastBefore = st.last;
    // sensitivity probe: evaluate with the symbol forced to hit/miss?  cheaper: replicate the marginal test
    AutoState st2 = st; Lut lutMiss; for (int i = 0; i < 7; i++) lutMiss.s[i] = ~0ull - i;   // nothing matches
    AutoState st3 = st; Lut lutHit = lutMiss;
    uint32_t s2, e2; CountSink h2;
    uint32_t evMiss = enc_eval(sp, recs[j].sym, n, recs[j].a, recs[j].b, st2, lutMiss, nullptr, s2, e2, h2);
    // hit: put the (rotated) symbol in front -- need the rotated symbol: recompute like enc_eval
    {
      uint32_t ss = recs[j].a - sp.W; if (sp.W > 1 && st.cursor > ss) ss = st.cursor;
      lutHit.s[0] = sym_rot(recs[j].sym, sp.W, ss - (recs[j].a - sp.W));
    }
    uint32_t evHit = enc_eval(sp, recs[j].sym, n, recs[j].a, recs[j].b, st3, lutHit, nullptr, s2, e2, h2);
    const bool marginal = ((evMiss ^ evHit) & EV_EMIT) != 0;
    // real evaluation
    Lut before = lut;
    const uint32_t ev = enc_eval(sp, recs[j].sym, n, recs[j].a, recs[j].b, st, lut, &r.agg, s, e, h);
    fl |= ev & EV_STATE_MASK;
    if (sp.K && (ev & EV_VALID))
    {
      uint32_t ss = s;
      const uint64_t sym = sym_rot(recs[j].sym, sp.W, ss - (recs[j].a - sp.W));
      const int idx = lut_find(before, sp.K, sym);
      if (marginal && idx >= known && known < sp.K) r.sens = true;
      if ((ev & EV_EMIT) && idx >= known && known < sp.K) known++;
    }
    if (ev & EV_EMIT) { r.bytes += h.len + (s - lastBefore); r.ntok++; }
  }
  r.cs.flags = fl; r.cs.last = st.last; r.cs.cursor = st.cursor; r.cs.lastSym = st.lastSym;
  return r;
}

int main(int argc, char **argv)
{
  if (argc < 5) return 1;
  FILE *f = fopen(argv[1], "rb"); fseek(f, 0, SEEK_END); n = (uint32_t)ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> in(n + 64, 0); if (fread(in.data(), 1, n, f) != n) return 2; fclose(f);
  const int W = atoi(argv[2]), BA = atoi(argv[3]), V = atoi(argv[4]);
  const size_t SCR = argc > 5 ? atoi(argv[5]) : 512;
  const int WARM = 12; const int J = argc > 6 ? atoi(argv[6]) : 16; const int RULE = argc > 7 ? atoi(argv[7]) : 0;
  sp = make_spec(W, BA, V);
  // records
  {
    uint32_t p = W;
    while (p < n)
    {
      if (in[p] != in[p - W]) { p++; continue; }
      uint32_t a = p; while (p < n && in[p] == in[p - W]) p++;
      if ((int)(p - a) >= sp.minM) { Rec r; r.a = a; r.b = p; r.sym = load_sym(&in[a - W], W); recs.push_back(r); }
    }
  }
  const size_t nSC = (recs.size() + SCR - 1) / SCR;
  printf("W=%d BA=%d V=%d records=%zu SCs=%zu\n", W, BA, V, recs.size(), nSC);
  // exact sequential reference
  std::vector<AutoState> exactIn(nSC); std::vector<Lut> exactLut(nSC);
  {
    AutoState st = enc_initial_state(); Lut lut; lut_init(lut, W);
    for (size_t s = 0; s < nSC; s++)
    {
      exactIn[s] = st; exactLut[s] = lut;
      ScOut o = run_sc(s * SCR, std::min(recs.size(), (s + 1) * SCR), st, lut);
      chunksum_apply(st, o.cs); if (sp.K) lut_apply(lut, sp.K, o.agg);
    }
  }


  // ---- sparse LUT walker prototype (RULE 8): exact table at every symbol-stretch boundary, guesses at super-chunk starts
  std::vector<Lut> walkLut(nSC);
  size_t walkSteps = 0, walkSlow = 0;
  {
    const size_t M = recs.size();
    auto emit_if = [&](size_t j, uint32_t last, bool hit) {
      AutoState st = enc_initial_state(); st.last = last; Lut l; for (int i = 0; i < 7; i++) l.s[i] = ~0ull - i; if (hit) l.s[0] = recs[j].sym;
      uint32_t s, e; CountSink h; return (enc_eval(sp, recs[j].sym, n, recs[j].a, recs[j].b, st, l, nullptr, s, e, h) & EV_EMIT) != 0; };
    auto cert0 = [&](size_t j) { // emitted whatever the state: miss, and the literal before it as long as it can get (from position 0)
      return emit_if(j, 0, false) && emit_if(j, recs[j].a - 1, false) && (recs[j].a < 0x100000u || emit_if(j, recs[j].a - 1 - 0xFFFFFu, false)); };
    auto cert = [&](size_t j) { return cert0(j) || (j > 0 && cert0(j - 1) && emit_if(j, recs[j - 1].b, false)); };
    AutoState st = enc_initial_state(); Lut L; lut_init(L, W);
    size_t jPrev = 0; size_t nextSc = 0;
    auto touch = [&](Lut &l, uint64_t z) { int idx = lut_find(l, sp.K, z); lut_touch(l, sp.K, idx, z); };
    for (size_t j = 1; j <= M; j++)
    {
      if (j < M && recs[j].sym == recs[j - 1].sym) continue;
      // boundary at j: previous stretch [jPrev, j)
      const uint64_t z = recs[jPrev].sym;
      long c = -1;
      for (long i = (long)j - 1; i >= (long)jPrev; i--) if (cert((size_t)i)) { c = i; break; }
      walkSteps++;
      if (c >= 0)
      {
        uint32_t last = recs[c].b;
        for (size_t r = (size_t)c + 1; r < j; r++) if (emit_if(r, last, true)) last = recs[r].b;
        Lut T = L; touch(T, z);
        while (nextSc < nSC && nextSc * SCR < j) { walkLut[nextSc] = (nextSc * SCR == jPrev) ? L : T; nextSc++; }
        L = T; st.last = last;
      }
      else
      {
        for (size_t r = jPrev; r < j; r++)
        {
          if (nextSc < nSC && nextSc * SCR == r) walkLut[nextSc++] = L;
          uint32_t s, e; CountSink h; enc_eval(sp, recs[r].sym, n, recs[r].a, recs[r].b, st, L, nullptr, s, e, h); walkSlow++;
        }
      }
      jPrev = j;
    }
    size_t wrong = 0; for (size_t s = 0; s < nSC; s++) if (!lut_equal(walkLut[s], exactLut[s], sp.K)) wrong++;
    printf("walker: %zu stretches, %zu records simulated one by one, %zu of %zu super-chunk guesses differ from the exact table\n", walkSteps, walkSlow, wrong, nSC);
  }
  auto same_out = [&](const ScOut &a, const ScOut &b) {
    bool bad = a.cs.last != b.cs.last || a.cs.flags != b.cs.flags || a.ntok != b.ntok || a.agg.m != b.agg.m || a.cs.cursor != b.cs.cursor || a.cs.lastSym != b.cs.lastSym;
    for (int i = 0; i < sp.K && i < (int)a.agg.m; i++) bad = bad || a.agg.s[i] != b.agg.s[i];
    return !bad; };
  auto lut_has = [&](const Lut &l, uint64_t x) { for (int i = 0; i < sp.K; i++) if (l.s[i] == x) return true; return false; };
  auto lut_push_front = [&](Lut &l, uint64_t x) { int p = sp.K - 1; for (int i = 0; i < sp.K; i++) if (l.s[i] == x) { p = i; break; } for (int i = p; i > 0; i--) l.s[i] = l.s[i - 1]; l.s[0] = x; };
  {
    std::vector<AutoState> usedIn(nSC); std::vector<Lut> usedLut(nSC), compLut(nSC), prevComp(nSC); std::vector<ScOut> out(nSC);
    for (size_t s = 0; s < nSC; s++)
    {
      AutoState st; Lut lut;
      if (s == 0) { st = enc_initial_state(); lut_init(lut, W); }
      else
      {
        size_t w0 = s * SCR - WARM;
        enc_neutral_state(sp, recs[w0].a, st, lut);
        ScOut o = run_sc(w0, s * SCR, st, lut);
        st.last = o.cs.flags & EV_EMIT ? o.cs.last : st.last; st.cursor = o.cs.flags & EV_VALID ? o.cs.cursor : st.cursor;
        if (o.cs.flags & EV_SYMSET) st.lastSym = o.cs.lastSym;
        if (sp.K) lut_apply(lut, sp.K, o.agg);
      }
      if (RULE == 9) lut = exactLut[s];
      if (RULE == 8) lut = walkLut[s];
      usedIn[s] = st; usedLut[s] = lut; prevComp[s] = lut;
      out[s] = run_sc(s * SCR, std::min(recs.size(), (s + 1) * SCR), st, lut);
    }
    size_t totalReruns = 0;
    for (int round = 0; round < 200; round++)
    {
      AutoState st = enc_initial_state(); Lut lut; lut_init(lut, W);
      size_t nDirty = 0, nWave = 0;
      std::vector<uint8_t> dirty(nSC, 0);
      std::vector<AutoState> compIn(nSC);
      for (size_t s = 0; s < nSC; s++)
      {
        bool bad = usedIn[s] != st;
        const bool lutDiff = sp.K && !lut_equal(usedLut[s], lut, sp.K);
        if (!bad && lutDiff && out[s].sens)
        {
          ScOut o2 = run_sc(s * SCR, std::min(recs.size(), (s + 1) * SCR), st, lut);
          bad = !same_out(o2, out[s]);
          if (!bad) usedLut[s] = lut;
        }
        compIn[s] = st; compLut[s] = lut;
        if (bad) { dirty[s] = 1; nDirty++; }
        chunksum_apply(st, out[s].cs); if (sp.K) lut_apply(lut, sp.K, out[s].agg);
      }
      if (!nDirty) { printf("wave J=%d rule=%d: converged after %d rounds, %zu re-runs\n", J, RULE, round, totalReruns); break; }
      // wave marking
      std::vector<uint8_t> wave(nSC, 0); std::vector<Lut> guess(nSC);
      for (size_t h = 0; h < nSC; h++) if (dirty[h])
      {
        const Lut T = compLut[h], O = prevComp[h];
        for (int j = 1; j < J; j++)
        {
          size_t m = h + j; if (m >= nSC || dirty[m]) break;
          Lut G;
          if (RULE == 0) G = T;
          else if (RULE == 1)
          { G = compLut[m]; for (int i = sp.K - 1; i >= 0; i--) if (!lut_has(O, T.s[i]) && !lut_has(G, T.s[i])) lut_push_front(G, T.s[i]); }
          else
          { // rule 2: D- removed (filled from T's tail), D+ inserted at front
            G = compLut[m];
            for (int i = 0; i < sp.K; i++) if (!lut_has(T, O.s[i]) && lut_has(G, O.s[i]))
            { int p = 0; while (G.s[p] != O.s[i]) p++; for (int q = p; q + 1 < sp.K; q++) G.s[q] = G.s[q + 1];
              uint64_t fill = O.s[i]; for (int q = sp.K - 1; q >= 0; q--) if (!lut_has(G, T.s[q]) || false) { bool in = false; for (int z = 0; z < sp.K - 1; z++) in = in || G.s[z] == T.s[q]; if (!in) { fill = T.s[q]; break; } }
              G.s[sp.K - 1] = fill; }
            for (int i = sp.K - 1; i >= 0; i--) if (!lut_has(O, T.s[i]) && !lut_has(G, T.s[i])) lut_push_front(G, T.s[i]);
          }
          if (lut_equal(G, usedLut[m], sp.K) || !out[m].sens) continue;
          ScOut o2 = run_sc(m * SCR, std::min(recs.size(), (m + 1) * SCR), compIn[m], G);
          if (same_out(o2, out[m])) continue;
          wave[m] = 1; guess[m] = G; nWave++;
        }
      }
      printf("round %d: dirty %zu wave %zu\n", round, nDirty, nWave);
      for (size_t s = 0; s < nSC; s++)
      {
        if (dirty[s]) { usedIn[s] = compIn[s]; usedLut[s] = compLut[s]; }
        else if (wave[s]) { usedIn[s] = compIn[s]; usedLut[s] = guess[s]; }
        else continue;
        out[s] = run_sc(s * SCR, std::min(recs.size(), (s + 1) * SCR), usedIn[s], usedLut[s]); totalReruns++;
      }
      prevComp = compLut;
    }
    // final check
    AutoState st2 = enc_initial_state(); Lut lut2; lut_init(lut2, W); size_t wrong = 0;
    for (size_t s = 0; s < nSC; s++)
    {
      if (st2 != exactIn[s] || (sp.K && !lut_equal(lut2, exactLut[s], sp.K))) wrong++;
      chunksum_apply(st2, out[s].cs); if (sp.K) lut_apply(lut2, sp.K, out[s].agg);
    }
    printf("  SCs whose scanned incoming state differs from the exact one: %zu\n", wrong);
  }
  return 0;
}
