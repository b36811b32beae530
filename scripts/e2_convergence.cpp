// e2_convergence.cpp -- analysis tool (not product code): how many grid-level rounds does the encoder automaton
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
  const int WARM = argc > 6 ? atoi(argv[6]) : 12;
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
  for (int scheme = 0; scheme < 3; scheme++)
  {
    std::vector<AutoState> usedIn(nSC); std::vector<Lut> usedLut(nSC); std::vector<ScOut> out(nSC);
    // round 0: neutral guess warmed up over WARM records
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
      usedIn[s] = st; usedLut[s] = lut;
      out[s] = run_sc(s * SCR, std::min(recs.size(), (s + 1) * SCR), st, lut);
    }
    for (int round = 0; round < 40; round++)
    {
      // verify scan
      AutoState st = enc_initial_state(); Lut lut; lut_init(lut, W);
      size_t nDirty = 0, nSens = 0, nWrongFinal = 0;
      std::vector<uint8_t> dirty(nSC, 0);
      for (size_t s = 0; s < nSC; s++)
      {
        bool bad = usedIn[s] != st;
        const bool lutDiff = sp.K && !lut_equal(usedLut[s], lut, sp.K);
        if (scheme == 0) bad = bad || lutDiff;
        else if (scheme == 1) bad = bad || (lutDiff && out[s].sens);
        else if (!bad && lutDiff && out[s].sens)
        { // ideal criterion: would a re-run change the summary?
          ScOut o2 = run_sc(s * SCR, std::min(recs.size(), (s + 1) * SCR), st, lut);
          bad = o2.cs.last != out[s].cs.last || o2.cs.flags != out[s].cs.flags || o2.ntok != out[s].ntok || o2.agg.m != out[s].agg.m;
          for (int i = 0; i < sp.K && i < (int)o2.agg.m; i++) bad = bad || o2.agg.s[i] != out[s].agg.s[i];
          if (!bad) usedLut[s] = lut;
        }
        nSens += out[s].sens;
        if (bad) { dirty[s] = 1; nDirty++; usedIn[s] = st; usedLut[s] = lut; }
        chunksum_apply(st, out[s].cs); if (sp.K) lut_apply(lut, sp.K, out[s].agg);
      }
      printf("scheme %d round %d: dirty %zu (sens %zu)\n", scheme, round, nDirty, nSens);
      if (!nDirty)
      { // check against the exact states (AutoState always; LUT in scheme 0)
        AutoState st2 = enc_initial_state(); Lut lut2; lut_init(lut2, W);
        for (size_t s = 0; s < nSC; s++)
        {
          if (st2 != exactIn[s] || (sp.K && !lut_equal(lut2, exactLut[s], sp.K))) nWrongFinal++;
          chunksum_apply(st2, out[s].cs); if (sp.K) lut_apply(lut2, sp.K, out[s].agg);
        }
        printf("  converged; SCs whose scanned incoming state differs from the exact one: %zu\n", nWrongFinal);
        break;
      }
      for (size_t s = 0; s < nSC; s++) if (dirty[s]) out[s] = run_sc(s * SCR, std::min(recs.size(), (s + 1) * SCR), usedIn[s], usedLut[s]);
    }
  }
  return 0;
}
