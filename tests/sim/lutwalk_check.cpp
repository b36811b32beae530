// lutwalk_check.cpp -- host model of the stretch walk of the 8-bit LUT encoders (hsrle_enc_lutwalk.cuh), a test tool.
// Uses the header's own decision functions (lw_emit, lw_cert0) and descriptor layout; restates what k_enc_lut_stretch /
// k_enc_lut_walk do with host loops; compares (1) lw_emit with the emit decision of enc_eval (hsrle_core.cuh) on every
// record of the exact sequential pass, (2) the table the walk has at every stretch boundary and every super-chunk start
// with the exact sequential one.  usage: lutwalk_check file variant(2|3)   -> prints counts, exit code 0 iff all zero
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "hsrle_enc_lutwalk.cuh"
using namespace hsrle;

struct Rec { uint32_t a, b; uint32_t sym; };

template <int V> static int run(const std::vector<uint8_t> &in, uint32_t n)
{
  constexpr Spec sp = make_spec(1, 1, V);
  std::vector<Rec> recs;
  for (uint32_t p = 1; p < n;)
  {
    if (in[p] != in[p - 1]) { p++; continue; }
    const uint32_t a = p; while (p < n && in[p] == in[p - 1]) p++;
    if ((int)(p - a) >= sp.minM) recs.push_back({ a, p, in[a - 1] });
  }
  const size_t M = recs.size();
  // exact sequential pass: table before every record; lw_emit against enc_eval
  std::vector<uint64_t> lutBefore(M + 1);
  size_t emitMismatch = 0;
  {
    AutoState st = enc_initial_state(); LutB L; lut_init(L, 1); L.v &= lutb_mask(sp.K);
    for (size_t j = 0; j < M; j++)
    {
      lutBefore[j] = L.v;
      const bool miss = lut_find(L, sp.K, recs[j].sym) == sp.K;
      const bool want = lw_emit<V>(recs[j].b - recs[j].a + 1, recs[j].a - st.last + 1, miss);
      uint32_t s, e; CountSink h;
      const uint32_t ev = enc_eval_t<CountSink, LutB, LutAggB>(sp, recs[j].sym, n, recs[j].a, recs[j].b, st, L, (LutAggB *)nullptr, s, e, h);
      if (((ev & EV_EMIT) != 0) != want) emitMismatch++;
    }
    lutBefore[M] = L.v;
  }
  // k_enc_lut_stretch: a descriptor per boundary
  auto cert = [&](size_t k) {
    if (lw_cert0<V>(recs[k].a, recs[k].b)) return true;
    return k > 0 && lw_cert0<V>(recs[k - 1].a, recs[k - 1].b) && lw_emit<V>(recs[k].b - recs[k].a + 1, recs[k].a - recs[k - 1].b + 1, true); };
  std::vector<LwDesc> desc;
  size_t overflow = 0;
  for (size_t j = 0; j < M; j++)
  {
    if (j > 0 && recs[j].sym == recs[j - 1].sym) continue;
    LwDesc d; d.j = (uint32_t)j; d.lastOut = 0; d.a = 0; d.b = 0; d.info = recs[j].sym & 0xFFu;
    if (j == 0) d.info |= LWD_CERT | LWD_FIRST;
    else
    {
      const uint32_t zPrev = recs[j - 1].sym;
      size_t k = j - 1; d.a = recs[k].a; d.b = recs[k].b;
      int found = 0, steps = 0;
      for (;;)
      {
        if (cert(k)) { found = 1; break; }
        if (k == 0 || recs[k - 1].sym != zPrev) break;
        if (++steps >= LW_BACK) { overflow++; break; }
        k--;
      }
      if (found)
      {
        uint32_t last = recs[k].b;
        for (size_t r = k + 1; r < j; r++) if (lw_emit<V>(recs[r].b - recs[r].a + 1, recs[r].a - last + 1, false)) last = recs[r].b;
        d.lastOut = last; d.info |= LWD_CERT;
      }
      else d.info |= (uint32_t)(j - k) << 8;
    }
    desc.push_back(d);
  }
  // k_enc_lut_walk
  size_t boundaryMismatch = 0, scMismatch = 0, scChecked = 0;
  if (!overflow)
  {
    uint32_t last = 0, zPrev = 0; LutB L; lut_init(L, 1); L.v &= lutb_mask(sp.K);
    std::vector<uint64_t> lb(desc.size());
    for (size_t k = 0; k < desc.size(); k++)
    {
      const LwDesc &d = desc[k];
      if (!(d.info & LWD_FIRST))
      {
        if (d.info & LWD_CERT) { lut_touch(L, sp.K, lut_find(L, sp.K, zPrev), zPrev); last = d.lastOut; }
        else
        {
          const uint32_t cnt = (d.info >> 8) & 0xFFu;
          for (uint32_t r = d.j - cnt; r < d.j; r++)
          {
            const int idx = lut_find(L, sp.K, zPrev);
            if (lw_emit<V>(recs[r].b - recs[r].a + 1, recs[r].a - last + 1, idx == sp.K)) { lut_touch(L, sp.K, idx, zPrev); last = recs[r].b; }
          }
        }
      }
      lb[k] = L.v; zPrev = d.info & 0xFFu;
      if (lb[k] != lutBefore[d.j]) boundaryMismatch++;
    }
    // super-chunk starts: exact at a boundary, else "as if the stretch had emitted" -- a guess that must be right except inside the
    // uncertain head of a stretch
    for (size_t s = 0; s * E2_SCR < M; s++)
    {
      const uint32_t r = (uint32_t)(s * E2_SCR);
      size_t lo = 0, hi = desc.size();
      while (hi - lo > 1) { const size_t mid = (lo + hi) / 2; if (desc[mid].j <= r) lo = mid; else hi = mid; }
      LutB g; g.v = lb[lo];
      if (desc[lo].j != r) { const uint32_t z = desc[lo].info & 0xFFu; lut_touch(g, sp.K, lut_find(g, sp.K, z), z); }
      scChecked++;
      if (g.v != lutBefore[r]) scMismatch++;
    }
  }
  printf("records=%zu boundaries=%zu overflow=%zu emit_mismatch=%zu boundary_mismatch=%zu sc_checked=%zu sc_guess_mismatch=%zu\n",
         M, desc.size(), overflow, emitMismatch, boundaryMismatch, scChecked, scMismatch);
  return (emitMismatch || boundaryMismatch) ? 1 : 0;
}

int main(int argc, char **argv)
{
  if (argc < 3) return 2;
  FILE *f = fopen(argv[1], "rb"); if (!f) return 2;
  fseek(f, 0, SEEK_END); const uint32_t n = (uint32_t)ftell(f); fseek(f, 0, SEEK_SET);
  std::vector<uint8_t> in(n + 64, 0); if (fread(in.data(), 1, n, f) != n) return 2; fclose(f);
  return atoi(argv[2]) == 2 ? run<V_LUT3>(in, n) : run<V_LUT7>(in, n);
}
