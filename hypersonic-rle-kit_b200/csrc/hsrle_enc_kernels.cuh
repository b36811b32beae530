// hsrle_enc_kernels.cuh -- sm_100a kernels of the encoder (see hsrle_enc.cuh for the pipeline).
#pragma once
#include <cuda_runtime.h>
#include "hsrle_enc.cuh"

namespace hsrle {

// ================================================================================================
// small device utilities
template <class T> __device__ __forceinline__ T shfl_up_t(const T &v, int d)
{
  static_assert(sizeof(T) % 4 == 0, "word multiple");
  T r;
  const uint32_t *s = reinterpret_cast<const uint32_t *>(&v);
  uint32_t *o = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 4); i++) o[i] = __shfl_up_sync(0xFFFFFFFFu, s[i], d);
  return r;
}
template <class T> __device__ __forceinline__ T shfl_idx_t(const T &v, int l)
{
  T r;
  const uint32_t *s = reinterpret_cast<const uint32_t *>(&v);
  uint32_t *o = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(T) / 4); i++) o[i] = __shfl_sync(0xFFFFFFFFu, s[i], l);
  return r;
}
__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long *p)
{
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u64(unsigned long long *p, unsigned long long v)
{
  asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ================================================================================================
// E1: candidate scan, single pass
constexpr unsigned long long TS_AGG = 1ull << 62, TS_INC = 2ull << 62, TS_MASK = (1ull << 62) - 1;

template <int W> __device__ __forceinline__ uint32_t m16_generic(const uint8_t *in, uint32_t n, uint32_t lastVec, int64_t v)
{ // slow path for the two halo vectors of a tile
  if (v < 0 || v > (int64_t)lastVec + 1) return 0;
  uint32_t c[6] = { 0, 0, 0, 0, 0, 0 };
  if (v >= 1) { const uint2 p = __ldg(reinterpret_cast<const uint2 *>(in + (size_t)v * 16 - 8)); c[0] = p.x; c[1] = p.y; }
  if (v <= (int64_t)lastVec) { const uint4 x = __ldg(reinterpret_cast<const uint4 *>(in) + v); c[2] = x.x; c[3] = x.y; c[4] = x.z; c[5] = x.w; }
  return m16_raw<W>(c) & m16_valid<W>((uint32_t)v, n);
}

template <int W, int MINM, class SymT>
__global__ void __launch_bounds__(E1_T) k_enc_scan(const EncBufs B)
{
  __shared__ uint16_t mb[E1_TILE_VECS + 2];
  __shared__ uint32_t warpTot[E1_T / 32];
  __shared__ uint32_t sTile;
  __shared__ unsigned long long sBase;
  EncScalars &sc = *B.sc;
  if (threadIdx.x == 0) sTile = atomicAdd(&sc.tileTicket, 1u);
  __syncthreads();
  const uint32_t tile = sTile;
  if (tile >= B.nTiles) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = B.n, lastVec = B.lastVec;
  const uint4 *in16 = reinterpret_cast<const uint4 *>(B.in);
  const uint32_t v0 = tile * E1_TILE_VECS + warp * (32 * E1_VPT) + lane;

  // phase 1: raw equality masks, one 16-byte vector per lane and step (each warp load is 512 contiguous bytes)
  uint32_t c2 = 0, c3 = 0;   // the 8 bytes before the current warp-step
  {
    const uint32_t vb = v0 - lane;
    if (lane == 0 && vb >= 1 && vb - 1 <= lastVec) { const uint2 p = __ldg(reinterpret_cast<const uint2 *>(B.in + (size_t)vb * 16 - 8)); c2 = p.x; c3 = p.y; }
    c2 = __shfl_sync(0xFFFFFFFFu, c2, 0); c3 = __shfl_sync(0xFFFFFFFFu, c3, 0);
  }
#pragma unroll
  for (int i = 0; i < E1_VPT; i++)
  {
    const uint32_t v = v0 + i * 32;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (v <= lastVec) x = __ldg(in16 + v);
    uint32_t p2 = __shfl_up_sync(0xFFFFFFFFu, x.z, 1), p3 = __shfl_up_sync(0xFFFFFFFFu, x.w, 1);
    if (lane == 0) { p2 = c2; p3 = c3; }
    c2 = __shfl_sync(0xFFFFFFFFu, x.z, 31); c3 = __shfl_sync(0xFFFFFFFFu, x.w, 31);
    const uint32_t c[6] = { p2, p3, x.x, x.y, x.z, x.w };
    mb[1 + warp * (32 * E1_VPT) + i * 32 + lane] = (uint16_t)(m16_raw<W>(c) & m16_valid<W>(v, n));
  }
  if (threadIdx.x == 0) mb[0] = (uint16_t)m16_generic<W>(B.in, n, lastVec, (int64_t)tile * E1_TILE_VECS - 1);
  if (threadIdx.x == 32) mb[E1_TILE_VECS + 1] = (uint16_t)m16_generic<W>(B.in, n, lastVec, (int64_t)(tile + 1) * E1_TILE_VECS);
  __syncthreads();

  // phase 2: qualifying run starts / ends of every vector, counts, scan
  uint32_t sMask[E1_VPT], eMask[E1_VPT], excl[E1_VPT];
  uint32_t run = 0;   // starts | ends << 16, running over the steps of this warp
#pragma unroll
  for (int i = 0; i < E1_VPT; i++)
  {
    const int idx = 1 + warp * (32 * E1_VPT) + i * 32 + lane;
    const uint32_t A = ((uint32_t)mb[idx - 1] >> 8) | ((uint32_t)mb[idx] << 8) | ((uint32_t)mb[idx + 1] << 24);
    m16_boundaries<MINM>(A, sMask[i], eMask[i]);
    const uint32_t cnt = __popc(sMask[i]) | (__popc(eMask[i]) << 16);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    excl[i] = run + inc - cnt;
    run += __shfl_sync(0xFFFFFFFFu, inc, 31);
  }
  if (lane == 0) warpTot[warp] = run;
  __syncthreads();
  uint32_t warpBase = 0, tileTot = 0;
#pragma unroll
  for (int w = 0; w < E1_T / 32; w++) { const uint32_t t = warpTot[w]; if (w < warp) warpBase += t; tileTot += t; }

  // decoupled look-back over tiles (warp 0)
  if (warp == 0)
  {
    const unsigned long long agg = (unsigned long long)(tileTot & 0xFFFFu) | ((unsigned long long)(tileTot >> 16) << 31);
    unsigned long long exclusive = 0;
    if (tile == 0) { if (lane == 0) st_volatile_u64(B.tileStatus + 0, TS_INC | agg); }
    else
    {
      if (lane == 0) st_volatile_u64(B.tileStatus + tile, TS_AGG | agg);
      int64_t base = (int64_t)tile - 1;
      for (;;)
      {
        const int64_t idx = base - lane;
        unsigned long long st = TS_INC;   // virtual tiles before the first one: inclusive prefix 0
        if (idx >= 0) { do { st = ld_volatile_u64(B.tileStatus + idx); } while ((st >> 62) == 0); }
        const uint32_t incMask = __ballot_sync(0xFFFFFFFFu, (st >> 62) == 2);
        const int firstInc = incMask ? (__ffs(incMask) - 1) : 32;
        unsigned long long val = (lane <= firstInc) ? (st & TS_MASK) : 0ull;
#pragma unroll
        for (int d = 16; d >= 1; d >>= 1) val += __shfl_xor_sync(0xFFFFFFFFu, val, d);
        exclusive += val;
        if (incMask) break;
        base -= 32;
      }
      if (lane == 0) st_volatile_u64(B.tileStatus + tile, TS_INC | (exclusive + agg));
    }
    if (lane == 0)
    {
      sBase = exclusive;
      if (tile == B.nTiles - 1)
      {
        const unsigned long long tot = exclusive + agg;
        const uint32_t nS = (uint32_t)(tot & 0x7FFFFFFFull), nE = (uint32_t)(tot >> 31);
        if (nS != nE || nS > B.maxRuns) { sc.status = ST_BADARG; sc.nRuns = 0; sc.nSC = 0; }
        else { sc.nRuns = nS; sc.nSC = (nS + E2_SCR - 1) / E2_SCR; }
      }
    }
  }
  __syncthreads();
  const unsigned long long base = sBase;
  const uint32_t baseS = (uint32_t)(base & 0x7FFFFFFFull) + (warpBase & 0xFFFFu);
  const uint32_t baseE = (uint32_t)(base >> 31) + (warpBase >> 16);
  SymT *runSym = reinterpret_cast<SymT *>(B.runSym);
#pragma unroll
  for (int i = 0; i < E1_VPT; i++)
  {
    const uint32_t p0 = (v0 + i * 32) * 16u;
    uint32_t ps = baseS + (excl[i] & 0xFFFFu), pe = baseE + (excl[i] >> 16);
    uint32_t s = sMask[i], e = eMask[i];
    while (s)
    {
      const uint32_t a = p0 + (__ffs(s) - 1); s &= s - 1;
      B.runA[ps] = a;
      runSym[ps] = (SymT)load_sym(B.in + a - W, W);
      ps++;
    }
    while (e) { B.runB[pe++] = p0 + (__ffs(e) - 1); e &= e - 1; }
  }
}

// ================================================================================================
// E2: automaton
// padded record index: one pad slot per 16 records keeps the 16-record-per-thread accesses conflict free
__device__ __forceinline__ int rec_slot(int j) { return j + (j >> 4); }

template <int W, int BA, int V, class SymT> struct EncCta
{
  static constexpr int K = (V == V_LUT3) ? 3 : (V == V_LUT7 ? 7 : 0);
  using Seg = SegSum<K>;
  static constexpr int NREC = E2_SCR + E2_CH;                 // with the halo chunk
  static constexpr int NSLOT = NREC + (NREC >> 4) + 1;

  struct Smem
  {
    uint32_t a[NSLOT], b[NSLOT];
    SymT sym[NSLOT];
    Seg warpTot[E2_T / 32];
    AutoState bcSt; Lut bcLut;                                  // broadcast slots
    AutoState serSt[E2_T]; Lut serLut[K ? E2_T : 1];            // states produced by the in-CTA sequential pass
    uint32_t flag;
  };

  // evaluate records [j0,j1) (local indices) from (st,lut); returns the segment summary
  template <bool COUNT_ONLY>
  static __device__ __forceinline__ Seg eval_range(const Smem &S, uint32_t n, int j0, int j1, AutoState &st, Lut &lut)
  {
    constexpr Spec sp = make_spec(W, BA, V);
    Seg r = segsum_identity<K>();
    uint32_t fl = 0;
    for (int j = j0; j < j1; j++)
    {
      const int q = rec_slot(j);
      uint32_t s, e; CountSink h;
      const uint32_t lastBefore = st.last;
      const uint32_t ev = enc_eval(sp, (uint64_t)S.sym[q], n, S.a[q], S.b[q], st, lut, K ? &r.agg : nullptr, s, e, h);
      fl |= ev;
      if (ev & EV_EMIT) { r.bytes += h.len + (s - lastBefore); r.ntok++; }
    }
    r.cs.flags = fl; r.cs.last = st.last; r.cs.cursor = st.cursor; r.cs.lastSym = st.lastSym;
    return r;
  }

  // exclusive scan of the per-thread segment summaries over the CTA; `total` = combination of all
  static __device__ __forceinline__ Seg block_excl_scan(Smem &S, const Seg &mine, Seg &total)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Seg inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const Seg o = shfl_up_t(inc, d);
      if (lane >= d) inc = segsum_combine<K>(o, inc);
    }
    if (lane == 31) S.warpTot[warp] = inc;
    Seg ex = shfl_up_t(inc, 1);
    if (lane == 0) ex = segsum_identity<K>();
    __syncthreads();
    Seg pre = segsum_identity<K>();
    total = segsum_identity<K>();
#pragma unroll
    for (int w = 0; w < E2_T / 32; w++)
    {
      const Seg t = S.warpTot[w];
      if (w < warp) pre = segsum_combine<K>(pre, t);
      total = segsum_combine<K>(total, t);
    }
    __syncthreads();
    return segsum_combine<K>(pre, ex);
  }

  static __device__ __forceinline__ bool state_differs(const AutoState &a, const Lut &la, const AutoState &b, const Lut &lb)
  {
    bool d = (a != b);
    if (K) d = d || !lut_equal(la, lb, K);
    return d;
  }

  // One super-chunk.  given == true: (gSt,gLut) is the incoming state to use for the first chunk;
  // otherwise it is guessed by warming up over the halo chunk.  Writes the per-chunk incoming states,
  // the super-chunk summary and (when guessed) the assumed incoming state.
  static __device__ void process(const EncBufs &B, Smem &S, uint32_t s, bool given, const AutoState &gSt, const Lut &gLut, Seg &totalOut)
  {
    constexpr Spec sp = make_spec(W, BA, V);
    const uint32_t nRuns = B.sc->nRuns, n = B.n;
    const uint32_t lo = s * E2_SCR;
    const uint32_t cnt = min((uint32_t)E2_SCR, nRuns - lo);
    const int halo = (s > 0) ? E2_CH : 0;
    const SymT *runSym = reinterpret_cast<const SymT *>(B.runSym);
    __syncthreads();   // previous users of the shared arrays are done
    for (int j = threadIdx.x + (E2_CH - halo); j < E2_CH + (int)cnt; j += E2_T)
    {
      const uint32_t g = lo + j - E2_CH;
      const int q = rec_slot(j);
      S.a[q] = B.runA[g]; S.b[q] = B.runB[g]; S.sym[q] = runSym[g];
    }
    __syncthreads();
    const int t = threadIdx.x;
    const int j0 = E2_CH + t * E2_CH;
    const int j1 = min(j0 + E2_CH, E2_CH + (int)cnt);
    const bool active = j0 < j1;

    AutoState stIn; Lut lutIn;
    if (t == 0 && given) { stIn = gSt; lutIn = gLut; }
    else if (t == 0 && s == 0) { stIn = enc_initial_state(); lut_init(lutIn, W); }
    else
    { // warm up over the previous chunk from the neutral guess
      const int w0 = j0 - E2_CH;
      enc_neutral_state(sp, active ? S.a[rec_slot(w0)] : 0u, stIn, lutIn);
      if (active) { AutoState ws = stIn; Lut wl = lutIn; (void)eval_range<true>(S, n, w0, j0, ws, wl); stIn = ws; lutIn = wl; }
    }
    Seg mine = segsum_identity<K>();
    if (active) { AutoState st = stIn; Lut lut = lutIn; mine = eval_range<true>(S, n, j0, j1, st, lut); }

    // fixed point of (scan -> compare -> re-run)
    AutoState st0; Lut lut0;   // incoming state of the super-chunk = what thread 0 used
    if (t == 0) { S.bcSt = stIn; S.bcLut = lutIn; }
    __syncthreads();
    st0 = S.bcSt; lut0 = S.bcLut;
    Seg total;
    bool converged = false;
    for (int it = 0; it < E2_MAXIT; it++)
    {
      const Seg pre = block_excl_scan(S, mine, total);
      AutoState want = st0; Lut wantLut = lut0;
      segsum_apply<K>(want, wantLut, pre);
      int changed = 0;
      if (active && t > 0 && state_differs(want, wantLut, stIn, lutIn))
      {
        stIn = want; lutIn = wantLut; changed = 1;
        AutoState st = stIn; Lut lut = lutIn; mine = eval_range<true>(S, n, j0, j1, st, lut);
      }
      if (!__syncthreads_or(changed)) { converged = true; break; }
    }
    if (!converged)
    { // exact in-CTA sequential pass: thread 0 threads the state through every chunk
      if (t == 0)
      {
        AutoState st = st0; Lut lut = lut0;
        for (int c = 0; c * E2_CH < (int)cnt; c++)
        {
          S.serSt[c] = st; if (K) S.serLut[K ? c : 0] = lut;
          const int a0 = E2_CH + c * E2_CH, a1 = min(a0 + E2_CH, E2_CH + (int)cnt);
          (void)eval_range<true>(S, n, a0, a1, st, lut);
        }
        atomicAdd(&B.sc->innerSerial, 1u);
      }
      __syncthreads();
      if (active)
      {
        stIn = S.serSt[t]; if (K) lutIn = S.serLut[K ? t : 0];
        AutoState st = stIn; Lut lut = lutIn; mine = eval_range<true>(S, n, j0, j1, st, lut);
      }
      (void)block_excl_scan(S, mine, total);
    }
    // publish
    if (active)
    {
      const uint32_t chunk = s * E2_T + t;
      B.cIn[chunk] = stIn;
      if (K) B.cLut[chunk] = lutIn;
    }
    if (t == 0)
    {
      B.scSum[s] = total.cs; if (K) B.scAgg[s] = total.agg;
      B.scBytes[s] = total.bytes; B.scTok[s] = total.ntok;
      if (!given) { B.scIn[s] = st0; if (K) B.scLut[s] = lut0; }
    }
    totalOut = total;
  }

  // ---- run by the last CTA of a round: scan of the super-chunk summaries, verification, finishing
  static __device__ void finish(const EncBufs &B, const AutoState &fin, uint64_t tokBytes, uint32_t nTok)
  { // all threads call; thread 0 writes header/terminator, everybody copies a short trailing literal
    constexpr Spec sp = make_spec(W, BA, V);
    EncScalars &sc = *B.sc;
    __shared__ uint32_t fPos, fLen, fOk;
    if (threadIdx.x == 0)
    {
      fOk = 0;
      sc.tokBytes = tokBytes; sc.nTok = nTok;
      const uint32_t L = B.n - fin.last;
      TokenHdr h; enc_terminator(sp, L, h);
      const uint64_t total = (uint64_t)sp.hdr + tokBytes + h.len + L;
      if (sc.status == ST_OK && (total > B.cap || total >= 0xFFFFFFF0ull)) sc.status = ST_OVERFLOW;
      if (sc.status == ST_OK)
      {
        sc.total = (uint32_t)total;
        uint8_t *o = B.out;
        const uint32_t nn = B.n, tt = (uint32_t)total;
        for (int k = 0; k < 4; k++) { o[k] = (uint8_t)(nn >> (8 * k)); o[4 + k] = (uint8_t)(tt >> (8 * k)); }
        if (sp.hdr == 9) o[8] = 0;
        const uint32_t pos = (uint32_t)(sp.hdr + tokBytes);
        for (uint32_t k = 0; k < h.len; k++) o[pos + k] = h.b[k];
        if (L >= BIG_COPY) { CopyDesc d; d.dst = pos + h.len; d.src = fin.last; d.len = L; B.bigList[atomicAdd(&sc.nBig, 1u)] = d; }
        else { fPos = pos + h.len; fLen = L; fOk = 1; }
      }
      else sc.total = 0;
      uint32_t *r = B.dResult;
      r[0] = sc.status == ST_OK ? sc.total : 0; r[1] = sc.status; r[2] = sc.nRuns; r[3] = sc.nSC;
      r[4] = sc.serialSC; r[5] = sc.innerSerial; r[6] = sc.nTok; r[7] = sc.nDirty[0];
    }
    __syncthreads();
    if (fOk) { const uint32_t src = fin.last; for (uint32_t k = threadIdx.x; k < fLen; k += blockDim.x) B.out[fPos + k] = B.in[src + k]; }
  }

  static __device__ void scan_verify(const EncBufs &B, Smem &S, int round)
  {
    EncScalars &sc = *B.sc;
    const uint32_t nSC = sc.nSC;
    const int t = threadIdx.x;
    __shared__ uint32_t sDirty, sFirst;
    if (t == 0) { sDirty = 0; sFirst = 0xFFFFFFFFu; }
    const uint32_t per = (nSC + E2_T - 1) / E2_T;
    const uint32_t lo = min(nSC, (uint32_t)t * per), hi = min(nSC, lo + per);
    Seg mine = segsum_identity<K>();
    for (uint32_t s = lo; s < hi; s++)
    {
      Seg e; e.cs = B.scSum[s]; if (K) e.agg = B.scAgg[s]; else e.agg.m = 0; e.bytes = B.scBytes[s]; e.ntok = B.scTok[s];
      mine = segsum_combine<K>(mine, e);
    }
    Seg total;
    const Seg pre = block_excl_scan(S, mine, total);
    AutoState st = enc_initial_state(); Lut lut; lut_init(lut, W);
    segsum_apply<K>(st, lut, pre);
    uint64_t bytes = pre.bytes;
    uint32_t nd = 0, first = 0xFFFFFFFFu;
    for (uint32_t s = lo; s < hi; s++)
    {
      bool bad = false;
      if (B.scIn[s] != st) { B.scIn[s] = st; bad = true; }
      if (K && !lut_equal(B.scLut[s], lut, K)) { B.scLut[s] = lut; bad = true; }
      B.scDirty[s] = bad ? 1 : 0;
      if (bad) { if (!nd) first = s; nd++; }
      B.scBase[s] = bytes;
      chunksum_apply(st, B.scSum[s]);
      if (K) lut_apply(lut, K, B.scAgg[s]);
      bytes += B.scBytes[s];
    }
    if (nd) { atomicAdd(&sDirty, nd); atomicMin(&sFirst, first); }
    __syncthreads();
    const uint32_t nDirty = sDirty, firstDirty = sFirst;
    if (t == 0) { sc.nDirty[round] = nDirty; sc.firstDirty[round] = firstDirty; }
    __syncthreads();
    AutoState fin = enc_initial_state(); Lut finLut; lut_init(finLut, W);
    segsum_apply<K>(fin, finLut, total);
    if (nDirty == 0) { finish(B, fin, total.bytes, total.ntok); return; }
    if (round < E2_ROUNDS - 1) return;

    // sequential repair from the first inconsistent super-chunk (its scanned incoming state is exact)
    if (t == 0) { S.bcSt = B.scIn[firstDirty]; if (K) S.bcLut = B.scLut[firstDirty]; }
    __syncthreads();
    AutoState run = S.bcSt; Lut runLut; if (K) runLut = S.bcLut; else lut_init(runLut, W);
    uint64_t runBytes = B.scBase[firstDirty];
    uint32_t runTok = 0;
    for (uint32_t s = 0; s < firstDirty; s++) runTok += B.scTok[s];   // small, uniform across threads
    for (uint32_t s = firstDirty; s < nSC; s++)
    {
      Seg tot;
      process(B, S, s, true, run, runLut, tot);
      if (t == 0) { B.scIn[s] = run; if (K) B.scLut[s] = runLut; B.scBase[s] = runBytes; B.scDirty[s] = 0; atomicAdd(&sc.serialSC, 1u); }
      segsum_apply<K>(run, runLut, tot);
      runBytes += tot.bytes; runTok += tot.ntok;
    }
    __syncthreads();
    finish(B, run, runBytes, runTok);
  }
};

template <int W, int BA, int V, class SymT>
__global__ void __launch_bounds__(E2_T) k_enc_auto(const EncBufs B, int round)
{
  using C = EncCta<W, BA, V, SymT>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  typename C::Smem &S = *reinterpret_cast<typename C::Smem *>(smemRaw);
  EncScalars &sc = *B.sc;
  if (round > 0 && sc.nDirty[round - 1] == 0) return;
  const uint32_t nSC = sc.nSC;
  for (uint32_t s = blockIdx.x; s < nSC; s += gridDim.x)
  {
    typename C::Seg tot;
    if (round == 0) { AutoState d = enc_initial_state(); Lut dl; lut_init(dl, W); C::process(B, S, s, false, d, dl, tot); }
    else if (B.scDirty[s])
    {
      AutoState g = B.scIn[s]; Lut gl; if (C::K) gl = B.scLut[s]; else lut_init(gl, W);
      C::process(B, S, s, true, g, gl, tot);
    }
  }
  // the last CTA to get here scans and verifies
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    S.flag = (atomicAdd(&sc.done[round], 1u) == gridDim.x - 1) ? 1u : 0u;
  }
  __syncthreads();
  if (!S.flag) return;
  __threadfence();
  C::scan_verify(B, S, round);
}

// ================================================================================================
// E3: emit
struct EmitDesc { uint32_t dst, src, len; };

__device__ __forceinline__ void copy_bytes_warp(uint8_t *dst, const uint8_t *src, uint32_t len, int lane)
{
  // dst-aligned 16-byte stores fed by unaligned 4-byte source words (funnel shift); byte head/tail
  if (len < 64) { for (uint32_t i = lane; i < len; i += 32) dst[i] = src[i]; return; }
  const uint32_t head = (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15);
  if ((uint32_t)lane < head) dst[lane] = src[lane];
  dst += head; src += head; len -= head;
  // keep the 4-byte source reads inside [src, src+len): stop the vector body 4 bytes early
  const uint32_t nv = (len - 4) >> 4;
  const uint32_t sh = ((uintptr_t)src & 3) * 8;
  const uint32_t *sw = reinterpret_cast<const uint32_t *>((uintptr_t)src & ~(uintptr_t)3);
  for (uint32_t v = lane; v < nv; v += 32)
  {
    const uint32_t *p = sw + v * 4;
    uint4 o;
    if (sh == 0) { o.x = p[0]; o.y = p[1]; o.z = p[2]; o.w = p[3]; }
    else
    {
      const uint32_t w0 = p[0], w1 = p[1], w2 = p[2], w3 = p[3], w4 = p[4];
      o.x = __funnelshift_r(w0, w1, sh); o.y = __funnelshift_r(w1, w2, sh); o.z = __funnelshift_r(w2, w3, sh); o.w = __funnelshift_r(w3, w4, sh);
    }
    reinterpret_cast<uint4 *>(dst)[v] = o;
  }
  for (uint32_t i = (nv << 4) + lane; i < len; i += 32) dst[i] = src[i];
}

constexpr int E3_NDESC = E2_SCR;

template <int W, int BA, int V, class SymT> struct EncEmitSmem
{
  using C = EncCta<W, BA, V, SymT>;
  uint32_t a[C::NSLOT], b[C::NSLOT];
  SymT sym[C::NSLOT];
  EmitDesc desc[E3_NDESC];
  unsigned long long warpTot[E2_T / 32];
  uint32_t nDesc;
};

template <int W, int BA, int V, class SymT>
__global__ void __launch_bounds__(E2_T) k_enc_emit(const EncBufs B)
{
  using C = EncCta<W, BA, V, SymT>;
  constexpr int K = C::K;
  constexpr Spec sp = make_spec(W, BA, V);
  using Smem = EncEmitSmem<W, BA, V, SymT>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  Smem &S = *reinterpret_cast<Smem *>(smemRaw);
  EncScalars &sc = *B.sc;
  if (sc.status != ST_OK) return;
  const uint32_t nSC = sc.nSC, nRuns = sc.nRuns, n = B.n;
  const SymT *runSym = reinterpret_cast<const SymT *>(B.runSym);
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (uint32_t s = blockIdx.x; s < nSC; s += gridDim.x)
  {
    const uint32_t lo = s * E2_SCR;
    const uint32_t cnt = min((uint32_t)E2_SCR, nRuns - lo);
    __syncthreads();
    if (t == 0) S.nDesc = 0;
    for (int j = t; j < (int)cnt; j += E2_T)
    {
      const int q = rec_slot(j + E2_CH);
      S.a[q] = B.runA[lo + j]; S.b[q] = B.runB[lo + j]; S.sym[q] = runSym[lo + j];
    }
    __syncthreads();
    const int j0 = E2_CH + t * E2_CH, j1 = min(j0 + E2_CH, E2_CH + (int)cnt);
    const bool active = j0 < j1;
    AutoState st0 = enc_initial_state(); Lut lut0; lut_init(lut0, W);
    if (active) { st0 = B.cIn[s * E2_T + t]; if (K) lut0 = B.cLut[s * E2_T + t]; }
    // pass 1: bytes of my chunk
    unsigned long long mine = 0;
    if (active)
    {
      AutoState st = st0; Lut lut = lut0; LutAgg dummy; dummy.m = 0;
      for (int j = j0; j < j1; j++)
      {
        const int q = rec_slot(j);
        uint32_t rs, re; CountSink h;
        const uint32_t lastBefore = st.last;
        const uint32_t ev = enc_eval(sp, (uint64_t)S.sym[q], n, S.a[q], S.b[q], st, lut, K ? &dummy : nullptr, rs, re, h);
        if (ev & EV_EMIT) mine += h.len + (rs - lastBefore);
      }
    }
    unsigned long long inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) S.warpTot[warp] = inc;
    __syncthreads();
    unsigned long long pre = 0;
#pragma unroll
    for (int w = 0; w < E2_T / 32; w++) if (w < warp) pre += S.warpTot[w];
    uint64_t pos = (uint64_t)sp.hdr + B.scBase[s] + pre + (inc - mine);
    // pass 2: headers, short literals inline, longer ones to the warp-cooperative list
    if (active)
    {
      AutoState st = st0; Lut lut = lut0; LutAgg dummy; dummy.m = 0;
      for (int j = j0; j < j1; j++)
      {
        const int q = rec_slot(j);
        uint32_t rs, re; PtrSink h; h.p = B.out + pos;
        const uint32_t lastBefore = st.last;
        const uint32_t ev = enc_eval(sp, (uint64_t)S.sym[q], n, S.a[q], S.b[q], st, lut, K ? &dummy : nullptr, rs, re, h);
        if (!(ev & EV_EMIT)) continue;
        const uint32_t lit = rs - lastBefore;
        const uint32_t dst = (uint32_t)(pos + h.len);
        if (lit <= E3_INLINE) { for (uint32_t k = 0; k < lit; k++) B.out[dst + k] = B.in[lastBefore + k]; }
        else if (lit < BIG_COPY) { const uint32_t slot = atomicAdd(&S.nDesc, 1u); EmitDesc d; d.dst = dst; d.src = lastBefore; d.len = lit; S.desc[slot] = d; }
        else { CopyDesc d; d.dst = dst; d.src = lastBefore; d.len = lit; B.bigList[atomicAdd(&sc.nBig, 1u)] = d; }
        pos += h.len + lit;
      }
    }
    __syncthreads();
    const uint32_t nd = S.nDesc;
    for (uint32_t d = warp; d < nd; d += E2_T / 32)
    {
      const EmitDesc e = S.desc[d];
      copy_bytes_warp(B.out + e.dst, B.in + e.src, e.len, lane);
    }
  }
}

// ================================================================================================
// grid-wide copy of the long literals
constexpr uint32_t BIG_PIECE = 16384;
static __global__ void __launch_bounds__(256) k_enc_copy_big(const EncBufs B)
{
  if (B.sc->status != ST_OK) return;
  const uint32_t nBig = B.sc->nBig;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr uint32_t SUB = BIG_PIECE / 8;    // one sub-piece per warp
  for (uint32_t i = 0; i < nBig; i++)
  {
    const CopyDesc cd = B.bigList[i];
    const uint32_t nPieces = (cd.len + BIG_PIECE - 1) / BIG_PIECE;
    for (uint32_t pc = blockIdx.x; pc < nPieces; pc += gridDim.x)
    {
      const uint32_t off = pc * BIG_PIECE + warp * SUB;
      if (off >= cd.len) continue;
      const uint32_t len = min(SUB, cd.len - off);
      copy_bytes_warp(B.out + cd.dst + off, B.in + cd.src + off, len, lane);
    }
  }
}

} // namespace hsrle
