// hsrle_enc_kernels.cuh -- sm_100a kernels of the encoder (see hsrle_enc.cuh for the pipeline).
#pragma once
#include <cuda_runtime.h>
#include <type_traits>
#include "hsrle_enc.cuh"
#include "hsrle_enc_lutwalk.cuh"
#include "hsrle_slice.cuh"

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
__device__ __forceinline__ uint32_t ld_volatile_u32g(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32g(uint32_t *p, uint32_t v)
{
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ================================================================================================
// E1: candidate scan, single pass.
//
// A CTA owns a macro-tile of E1_WARPS x 16 KiB; each warp walks its own contiguous 16 KiB span in 32 steps of
// 512 bytes (one 16-byte vector per lane, so every warp load is one contiguous 512-byte request).  The
// equality masks of neighbouring vectors travel by shuffles, so phase A needs no block barrier at all.
// Offsets: per-step warp totals -> warp scan -> block scan -> ONE look-back per macro-tile in which all
// 256 threads sum the aggregates of the up-to-255 preceding tiles of the same 256-tile group plus the
// inclusive prefix published by the last tile of the previous group (no serial chain inside a group).
constexpr unsigned long long TS_FLAG = 1ull << 63, TS_MASK = (1ull << 62) - 1;
constexpr int E1_GROUP = 256;

template <int W> __device__ __forceinline__ uint32_t m16_generic(const uint8_t *__restrict__ in, uint32_t n, uint32_t lastVec, int64_t v)
{ // slow path for the vectors just outside a warp span
  if (v < 0 || v > (int64_t)lastVec + 1) return 0;
  uint32_t c[6] = { 0, 0, 0, 0, 0, 0 };
  if (v >= 1) { const uint2 p = __ldg(reinterpret_cast<const uint2 *>(in + (size_t)v * 16 - 8)); c[0] = p.x; c[1] = p.y; }
  if (v <= (int64_t)lastVec) { const uint4 x = __ldg(reinterpret_cast<const uint4 *>(in) + v); c[2] = x.x; c[3] = x.y; c[4] = x.z; c[5] = x.w; }
  return m16_raw<W>(c) & m16_valid<W>((uint32_t)v, n);
}

template <int W, int MINM, class SymT>
__global__ void __launch_bounds__(E1_T) k_enc_scan(const EncBufs B)
{
  __shared__ uint32_t masks[E1_WARPS][E1_STEPS][32];   // starts | ends << 16 of every vector of the macro-tile
  __shared__ uint32_t warpTot[E1_WARPS];
  __shared__ unsigned long long redBuf[E1_WARPS];
  __shared__ uint32_t sTile;
  EncScalars &sc = *B.sc;
  if (threadIdx.x == 0) sTile = atomicAdd(&sc.tileTicket, 1u);
  __syncthreads();
  const uint32_t tile = sTile;
  if (tile >= B.nTiles) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = B.n, lastVec = B.lastVec;
  const uint8_t *__restrict__ in = B.in;
  const uint4 *in16 = reinterpret_cast<const uint4 *>(in);
  const int steps = (int)B.scanSteps;                  // 512-byte steps per warp in this call (multiple of 4, <= E1_STEPS)
  const uint32_t vw = B.vecBase + tile * (uint32_t)(E1_T * steps) + warp * (32 * steps);   // first vector of this warp's span

  // ---- phase A: masks and per-step totals (no block barrier); loads run one group of 4 steps ahead
  uint32_t c2 = 0, c3 = 0;                 // the 8 input bytes before the current step
  if (lane == 0 && vw >= 1 && vw - 1 <= lastVec) { const uint2 p = __ldg(reinterpret_cast<const uint2 *>(in + (size_t)vw * 16 - 8)); c2 = p.x; c3 = p.y; }
  c2 = __shfl_sync(0xFFFFFFFFu, c2, 0); c3 = __shfl_sync(0xFFFFFFFFu, c3, 0);
  const uint32_t mBeforeSpan = __shfl_sync(0xFFFFFFFFu, (lane == 0) ? m16_generic<W>(in, n, lastVec, (int64_t)vw - 1) : 0u, 0);
  const uint32_t mAfterSpan = __shfl_sync(0xFFFFFFFFu, (lane == 0) ? m16_generic<W>(in, n, lastVec, (int64_t)vw + 32 * steps) : 0u, 0);
  uint32_t myStepTot = 0;                  // lane j keeps the total of step j
  auto finalize = [&](int j, uint32_t prevLast, uint32_t cur, uint32_t nextFirst)
  {
    uint32_t left = __shfl_up_sync(0xFFFFFFFFu, cur, 1);
    if (lane == 0) left = prevLast;
    uint32_t right = __shfl_down_sync(0xFFFFFFFFu, cur, 1);
    if (lane == 31) right = nextFirst;
    const uint32_t A = (left >> 8) | (cur << 8) | (right << 24);
    uint32_t sM, eM;
    m16_boundaries<MINM>(A, sM, eM);
    masks[warp][j][lane] = sM | (eM << 16);
    const uint32_t tot = __reduce_add_sync(0xFFFFFFFFu, (uint32_t)__popc(sM) | ((uint32_t)__popc(eM) << 16));
    if (lane == j) myStepTot = tot;
  };
  constexpr int GRP = 4;
  uint4 buf[GRP];
#pragma unroll
  for (int k = 0; k < GRP; k++) { const uint32_t v = vw + k * 32 + lane; buf[k] = (v <= lastVec) ? __ldg(in16 + v) : make_uint4(0, 0, 0, 0); }
  uint32_t pendM = 0, pendPrevLast = mBeforeSpan;
  const int nGroups = steps / GRP;
  for (int g = 0; g < nGroups; g++)
  {
    uint4 nxt[GRP];
    if (g + 1 < nGroups)
    {
#pragma unroll
      for (int k = 0; k < GRP; k++) { const uint32_t v = vw + ((g + 1) * GRP + k) * 32 + lane; nxt[k] = (v <= lastVec) ? __ldg(in16 + v) : make_uint4(0, 0, 0, 0); }
    }
    uint32_t m[GRP];
#pragma unroll
    for (int k = 0; k < GRP; k++)
    {
      const uint32_t vs = vw + (g * GRP + k) * 32;     // first vector of the step
      const uint4 x = buf[k];
      uint32_t p2 = __shfl_up_sync(0xFFFFFFFFu, x.z, 1), p3 = __shfl_up_sync(0xFFFFFFFFu, x.w, 1);
      if (lane == 0) { p2 = c2; p3 = c3; }
      c2 = __shfl_sync(0xFFFFFFFFu, x.z, 31); c3 = __shfl_sync(0xFFFFFFFFu, x.w, 31);
      const uint32_t c[6] = { p2, p3, x.x, x.y, x.z, x.w };
      m[k] = m16_raw<W>(c);
      if (vs == 0 || vs + 32 > lastVec) m[k] &= m16_valid<W>(vs + lane, n);
    }
    if (g > 0) finalize(g * GRP - 1, pendPrevLast, pendM, __shfl_sync(0xFFFFFFFFu, m[0], 0));
    uint32_t pl = (g > 0) ? __shfl_sync(0xFFFFFFFFu, pendM, 31) : mBeforeSpan;
#pragma unroll
    for (int k = 0; k + 1 < GRP; k++)
    {
      finalize(g * GRP + k, pl, m[k], __shfl_sync(0xFFFFFFFFu, m[k + 1], 0));
      pl = __shfl_sync(0xFFFFFFFFu, m[k], 31);
    }
    pendM = m[GRP - 1]; pendPrevLast = pl;
    if (g + 1 < nGroups)
    {
#pragma unroll
      for (int k = 0; k < GRP; k++) buf[k] = nxt[k];
    }
  }
  finalize(steps - 1, pendPrevLast, pendM, mAfterSpan);
  // exclusive scan of the step totals inside the warp
  uint32_t stepInc = myStepTot;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, stepInc, d); if (lane >= d) stepInc += o; }
  const uint32_t stepExcl = stepInc - myStepTot;
  if (lane == 31) warpTot[warp] = stepInc;
  __syncthreads();
  uint32_t warpBase = 0, tileTot = 0;
#pragma unroll
  for (int w = 0; w < E1_WARPS; w++) { const uint32_t t = warpTot[w]; if (w < warp) warpBase += t; tileTot += t; }

  // ---- one look-back per macro-tile, all threads
  const unsigned long long agg = (unsigned long long)(tileTot & 0xFFFFu) | ((unsigned long long)(tileTot >> 16) << 31);
  if (threadIdx.x == 0) st_volatile_u64(B.tileStatus + 2 * (size_t)tile, TS_FLAG | agg);
  const uint32_t g0 = (tile / E1_GROUP) * E1_GROUP;
  unsigned long long part = 0;
  {
    // thread 0: inclusive prefix of the tile before the group; threads 1..: aggregates of tiles g0 .. tile-1
    if (threadIdx.x == 0) { if (g0 > 0) { unsigned long long st; do { st = ld_volatile_u64(B.tileStatus + 2 * (size_t)(g0 - 1) + 1); } while (!(st & TS_FLAG)); part = st & TS_MASK; } }
    else
    {
      const uint32_t p = g0 + threadIdx.x - 1;
      if (p < tile) { unsigned long long st; do { st = ld_volatile_u64(B.tileStatus + 2 * (size_t)p); } while (!(st & TS_FLAG)); part = st & TS_MASK; }
    }
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) part += __shfl_xor_sync(0xFFFFFFFFu, part, d);
    if (lane == 0) redBuf[warp] = part;
  }
  __syncthreads();
  unsigned long long exclusive = 0;
#pragma unroll
  for (int w = 0; w < E1_WARPS; w++) exclusive += redBuf[w];
  if (threadIdx.x == 0)
  {
    st_volatile_u64(B.tileStatus + 2 * (size_t)tile + 1, TS_FLAG | (exclusive + agg));
    if (tile == B.nTiles - 1)
    {
      const unsigned long long tot = exclusive + agg;
      const uint32_t nS = (uint32_t)(tot & 0x7FFFFFFFull), nE = (uint32_t)(tot >> 31);
      sc.nStarts = nS; sc.nEnds = nE;
      if (B.sliceMode) { if (nS > B.maxRuns || nE > B.maxRuns) sc.status = ST_BADARG; }    // paired up by k_enc_slice_link
      else if (nS != nE || nS > B.maxRuns) { sc.status = ST_BADARG; sc.nRuns = 0; sc.nSC = 0; }
      else { sc.nRuns = nS; sc.nSC = (nS + E2_SCR - 1) / E2_SCR; }
    }
  }

  // ---- phase C: write the records
  const uint32_t baseS = (uint32_t)(exclusive & 0x7FFFFFFFull) + (warpBase & 0xFFFFu);
  const uint32_t baseE = (uint32_t)(exclusive >> 31) + (warpBase >> 16);
  SymT *__restrict__ runSym = reinterpret_cast<SymT *>(B.runSym);
  uint32_t *__restrict__ runA = B.runA, *__restrict__ runB = B.runB;
  for (int j = 0; j < steps; j++)
  {
    const uint32_t stepBase = __shfl_sync(0xFFFFFFFFu, stepExcl, j);
    const uint32_t stepTot = __shfl_sync(0xFFFFFFFFu, myStepTot, j);
    if (stepTot == 0) continue;
    const uint32_t m = masks[warp][j][lane];
    uint32_t s = m & 0xFFFFu, e = m >> 16;
    const uint32_t cnt = __popc(s) | (__popc(e) << 16);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    const uint32_t ex = stepBase + inc - cnt;
    uint32_t ps = baseS + (ex & 0xFFFFu), pe = baseE + (ex >> 16);
    const uint32_t p0 = (vw + j * 32 + lane) * 16u;
    while (s)
    {
      const uint32_t a = p0 + (__ffs(s) - 1); s &= s - 1;
      runA[ps] = a;
      runSym[ps] = (SymT)load_sym(in + a - W, W);
      ps++;
    }
    while (e) { runB[pe++] = p0 + (__ffs(e) - 1); e &= e - 1; }
  }
}

// literals of MED_COPY bytes or more are copied by the grid-wide k_enc_copy_big: one warp per "medium"
// literal, the whole grid piece-wise over the few "huge" ones
__device__ __forceinline__ void enc_push_copy(const EncBufs &B, uint32_t dst, uint32_t src, uint32_t len)
{
  CopyDesc d; d.dst = dst; d.src = src; d.len = len;
  if (len >= BIG_COPY) B.bigList[atomicAdd(&B.sc->nBig, 1u)] = d;
  else B.medList[atomicAdd(&B.sc->nMed, 1u)] = d;
}

// ================================================================================================
// E2: automaton
// padded record index: one pad slot per chunk keeps the CH-records-per-thread accesses conflict free
__device__ __forceinline__ int rec_slot(int j) { return j + j / E2_CH; }

// state the automaton starts from: the stream's initial state, or the slice's incoming state (hsrle_slice.cuh)
template <class LutT> __device__ __forceinline__ void enc_stream_incoming(const EncBufs &B, int W, AutoState &st, LutT &lut)
{
  if (B.sliceMode) { st = B.sliceIn->st; lut_from(lut, B.sliceIn->lut); }
  else { st = enc_initial_state(); lut_init(lut, W); }
}

template <int W, int BA, int V, class SymT> struct EncCta
{
  static constexpr int K = (V == V_LUT3) ? 3 : (V == V_LUT7 ? 7 : 0);
  using LutR = typename LutRep<W>::L;     // register representation of the table: one packed word for 8-bit symbols
  using AggR = typename LutRep<W>::A;
  using Seg = SegSum<K, AggR>;
  static constexpr int NREC = E2_SCR + E2_WARM;               // with the warm-up halo
  static constexpr int NSLOT = NREC + NREC / E2_CH + 1;
  // one warp per super-chunk, E2L_CH records per lane: a 32-element warp scan per fixed-point step, no block barrier
  // (LUT codecs: the scan element carries the 7-entry LUT aggregate; plain/packed: 2.6x fewer instructions than four
  // records per thread with block scans)
  static constexpr int NW = E2_T / 32;
  static constexpr int NSLOTL = NREC + NREC / E2L_CH + 2;

  struct WarpRecs
  {
    uint32_t a[NSLOTL], b[NSLOTL];
    SymT sym[NSLOTL];
    AutoState serSt[32]; LutR serLut[K ? 32 : 1];              // states produced by the sequential pass
    AutoState snap[K ? 1 : 32][E2L_CH / E2_CH];                 // plain/packed: per lane, the state at every E3 chunk boundary
    uint64_t fo[8];
    uint32_t foMiss, sens;
    ScQueries q;
  };
  struct Smem { WarpRecs r[NW]; };                            // round 0: one super-chunk per warp
  struct FixSmem                                                // the single-CTA verify / repair kernel
  {
    WarpRecs r[FIX_W];
    Seg warpTot[FIX_W];
    Seg scanTot;
    Seg bcTot;
    AutoState bcSt; LutR bcLut;                                 // broadcast slots
    uint64_t warpBytes[FIX_W];
    uint32_t nDirty, firstDirty, nList;
    uint32_t list[FIX_LIST];                                    // dirty super-chunks of the round (when they fit)
  };

  // evaluate records [j0,j1) (local indices; record j lives in slot j + j / CHS) from (st,lut); returns the segment summary
  // sens0 (LUT codecs): set when a decision was marginal for a symbol that had not been emitted earlier in [j0,j1) --
  // only then can the range's decisions depend on the table it started from (hsrle_enc.cuh)
  template <int CHS>
  static __device__ __forceinline__ Seg eval_range(const uint32_t *ra, const uint32_t *rb, const SymT *rs, uint32_t n, uint32_t floor, int j0, int j1,
                                                   AutoState &st, LutR &lut, uint32_t *sens0 = nullptr, AutoState *snap = nullptr)
  {
    constexpr Spec sp = make_spec(W, BA, V);
    Seg r = segsum_identity<K, AggR>();
    uint32_t fl = 0, known = 0, sens = 0;
    for (int j = j0; j < j1; j++)
    {
      if (snap && ((j - j0) & (E2_CH - 1)) == 0) snap[(j - j0) / E2_CH] = st;      // state at every E3 chunk boundary
      const int q = j + j / CHS;
      uint32_t s, e; CountSink h;
      const uint32_t lastBefore = st.last;
      const uint32_t ev = enc_eval_t<CountSink, LutR, AggR>(sp, (uint64_t)rs[q], n, ra[q], rb[q], st, lut, K ? &r.agg : (AggR *)nullptr, s, e, h);
      fl |= ev & EV_STATE_MASK;
      if (ev & EV_EMIT) { r.bytes += h.len + slice_lit_len(lastBefore, s, floor); r.ntok++; }
      if constexpr (K != 0)
      {
        const uint32_t idx = (ev >> EV_IDX_SHIFT) & 7u;
        if ((ev & EV_VALID) && idx >= known && known < (uint32_t)K) { if (ev & EV_MARG) sens = 1; if (ev & EV_EMIT) known++; }
      }
    }
    if (K && sens0) *sens0 = sens;
    r.cs.flags = fl; r.cs.last = st.last; r.cs.cursor = st.cursor; r.cs.lastSym = st.lastSym;
    return r;
  }

  // exclusive scan of the per-thread segment summaries over the CTA; `total` = combination of all
  static __device__ __forceinline__ Seg block_excl_scan(FixSmem &S, const Seg &mine, Seg &total)
  {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    Seg inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const Seg o = shfl_up_t(inc, d);
      if (lane >= d) inc = segsum_combine(o, inc);
    }
    if (lane == 31) S.warpTot[warp] = inc;
    Seg ex = shfl_up_t(inc, 1);
    if (lane == 0) ex = segsum_identity<K, AggR>();
    __syncthreads();
    if (warp == 0)
    { // the warp totals are scanned by one warp (not combined one after the other by every thread): exclusive prefix per warp
      Seg w = segsum_identity<K, AggR>();
      if (lane < FIX_W) w = S.warpTot[lane];
#pragma unroll
      for (int d = 1; d < FIX_W; d <<= 1)
      {
        const Seg o = shfl_up_t(w, d);
        if (lane >= d) w = segsum_combine(o, w);
      }
      Seg wex = shfl_up_t(w, 1);
      if (lane == 0) wex = segsum_identity<K, AggR>();
      if (lane < FIX_W) S.warpTot[lane] = wex;
      if (lane == FIX_W - 1) S.scanTot = w;
    }
    __syncthreads();
    const Seg pre = S.warpTot[warp];
    total = S.scanTot;
    __syncthreads();
    return segsum_combine(pre, ex);
  }

  static __device__ __forceinline__ bool state_differs(const AutoState &a, const LutR &la, const AutoState &b, const LutR &lb)
  {
    bool d = (a != b);
    if (K) d = d || !lut_equal(la, lb, K);
    return d;
  }

  // ---- One super-chunk per WARP, E2L_CH records per lane.  given == true: (gSt,gLut) is the incoming state to use;
  //      otherwise it is guessed by warming up over the halo records.  Writes the per-chunk incoming states (E3), the
  //      super-chunk summary and (when guessed) the assumed incoming state.  LUT codecs additionally publish what the
  //      sensitivity scheme needs (hsrle_enc.cuh: enc_fo_misses / enc_chunk_lut): scFo, scFlags, cKnown, and scBytes
  //      without the symbol bytes of the first emissions.
  static __device__ __forceinline__ void warp_scan(const Seg &mine, Seg &pre, Seg &total)
  {
    const int lane = threadIdx.x & 31;
    Seg inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const Seg o = shfl_up_t(inc, d);
      if (lane >= d) inc = segsum_combine(o, inc);
    }
    pre = shfl_up_t(inc, 1);
    if (lane == 0) pre = segsum_identity<K, AggR>();
    total = shfl_idx_t(inc, 31);
  }

  //      reuse: the super-chunk was evaluated before -- every lane starts from the state it converged to last time (B.cIn / B.cLut,
  //      the table rebuilt under the new incoming one) instead of a warmed-up guess: a re-run after a changed incoming table then
  //      settles in one or two scans instead of half a dozen.
  static __device__ void process(const EncBufs &B, WarpRecs &R, uint32_t s, bool given, const AutoState &gSt, const LutR &gLut, Seg &totalOut, bool reuse = false)
  {
    constexpr Spec sp = make_spec(W, BA, V);
    const int lane = threadIdx.x & 31;
    const uint32_t nRuns = B.sc->nRuns, n = B.n, floor = B.sliceLo, endShift = B.sc->endShift;
    const uint32_t lo = s * E2_SCR;
    const uint32_t cnt = min((uint32_t)E2_SCR, nRuns - lo);
    const int halo = (s > 0) ? E2_WARM : 0;
    const SymT *runSym = reinterpret_cast<const SymT *>(B.runSym);
    __syncwarp();
    for (int j = lane + (E2_WARM - halo); j < E2_WARM + (int)cnt; j += 32)
    {
      const uint32_t g = lo + j - E2_WARM;
      const int q = j + j / E2L_CH;
      R.a[q] = B.runA[g]; R.b[q] = B.runB[g + endShift]; R.sym[q] = runSym[g];
    }
    if (lane == 0) { R.foMiss = 0; R.sens = 0; R.q.n = 0; }
    __syncwarp();
    const int j0 = E2_WARM + lane * E2L_CH;
    const int j1 = min(j0 + E2L_CH, E2_WARM + (int)cnt);
    const bool active = j0 < j1;

    AutoState stIn; LutR lutIn;
    if (lane == 0 && given) { stIn = gSt; lutIn = gLut; }
    else if (lane == 0 && s == 0) enc_stream_incoming(B, W, stIn, lutIn);
    else if (given && reuse && active)
    {
      const uint32_t chunk = s * E2_T + (uint32_t)lane * (E2L_CH / E2_CH);
      stIn = B.cIn[chunk];
      if constexpr (K != 0)
      {
        Lut g = B.cLut[chunk], gi; lut_to(gi, gLut);
        enc_chunk_lut(g, B.cKnown[chunk], K, gi);
        lut_from(lutIn, g);
      }
      else lut_init(lutIn, W);
    }
    else
    { // warm up over the preceding E2_WARM records from the neutral guess
      const int w0 = max(j0 - E2_WARM, E2_WARM - halo);
      enc_neutral_state(sp, (active && w0 < j0) ? R.a[w0 + w0 / E2L_CH] : 0u, stIn, lutIn);
      if (active && w0 < j0) { AutoState ws = stIn; LutR wl = lutIn; (void)eval_range<E2L_CH>(R.a, R.b, R.sym, n, floor, w0, j0, ws, wl); stIn = ws; lutIn = wl; }
      // 8-bit LUT codecs: the table at the super-chunk start comes from the stretch walk (hsrle_enc_lutwalk.cuh) when it ran
      if constexpr (K != 0 && W == 1) { if (lane == 0 && !given && B.sc->lwOk) lutIn.v = B.scGuess[s]; }
    }
    Seg mine = segsum_identity<K, AggR>();
    uint32_t sens0 = 0;
    AutoState *const snapP = K ? nullptr : R.snap[K ? 0 : lane];   // plain/packed: the chunk states E3 needs fall out of the last evaluation
    if (active) { AutoState st = stIn; LutR lut = lutIn; mine = eval_range<E2L_CH>(R.a, R.b, R.sym, n, floor, j0, j1, st, lut, &sens0, snapP); }
    const AutoState st0 = shfl_idx_t(stIn, 0);
    LutR lut0;
    if constexpr (K != 0) lut0 = shfl_idx_t(lutIn, 0); else lut_init(lut0, W);

    // fixed point of (scan -> compare -> re-run).  A lane whose decisions cannot depend on the table it started from
    // (sens0 == 0) only takes the exact table; its summary stands (its byte count is redone by the final pass).
    Seg pre, total;
    bool converged = false;
    for (int it = 0; it < E2_MAXIT; it++)
    {
      warp_scan(mine, pre, total);
      AutoState want = st0; LutR wantLut = lut0;
      segsum_apply(want, wantLut, pre);
      int changed = 0;
      if (active && lane > 0)
      {
        const bool lutDiff = !lut_equal(wantLut, lutIn, K);
        if (want != stIn || (lutDiff && sens0))
        {
          stIn = want; lutIn = wantLut; changed = 1;
          AutoState st = stIn; LutR lut = lutIn; mine = eval_range<E2L_CH>(R.a, R.b, R.sym, n, floor, j0, j1, st, lut, &sens0, snapP);
        }
        else if (lutDiff) lutIn = wantLut;
      }
      if (!__any_sync(0xFFFFFFFFu, changed)) { converged = true; break; }
    }
    if (!converged)
    { // exact sequential pass: lane 0 threads the state through every lane's records
      if (lane == 0)
      {
        AutoState st = st0; LutR lut = lut0;
        for (int c = 0; c * E2L_CH < (int)cnt; c++)
        {
          R.serSt[c] = st; if (K) R.serLut[K ? c : 0] = lut;
          const int a0 = E2_WARM + c * E2L_CH, a1 = min(a0 + E2L_CH, E2_WARM + (int)cnt);
          (void)eval_range<E2L_CH>(R.a, R.b, R.sym, n, floor, a0, a1, st, lut);
        }
        atomicAdd(&B.sc->innerSerial, 1u);
      }
      __syncwarp();
      if (active)
      {
        stIn = R.serSt[lane]; if (K) lutIn = R.serLut[K ? lane : 0];
        AutoState st = stIn; LutR lut = lutIn; mine = eval_range<E2L_CH>(R.a, R.b, R.sym, n, floor, j0, j1, st, lut, nullptr, snapP);
      }
      warp_scan(mine, pre, total);
    }
    // final pass (LUT codecs): per-chunk incoming states for E3, sensitivity, first emissions
    if constexpr (K == 0)
    {
      if (active)
      {
#pragma unroll
        for (int c = 0; c < E2L_CH / E2_CH; c++)
          if (j0 + c * E2_CH < j1) B.cIn[s * E2_T + (uint32_t)(j0 - E2_WARM) / E2_CH + c] = snapP[c];
      }
    }
    else if (active)
    {
      AutoState st = stIn; LutR lut = lutIn;
      uint32_t known = pre.agg.m, sensL = 0;
      uint64_t bytesL = 0;
      for (int j = j0; j < j1; j++)
      {
        if (((j - j0) & (E2_CH - 1)) == 0)
        {
          const uint32_t chunk = s * E2_T + (uint32_t)(j - E2_WARM) / E2_CH;
          B.cIn[chunk] = st;
          if (K) { lut_to(B.cLut[chunk], lut); B.cKnown[chunk] = (uint8_t)known; }
        }
        const int q = j + j / E2L_CH;
        uint32_t rs, re; CountSink h;
        const uint32_t lastBefore = st.last;
        const uint32_t ev = enc_eval_t<CountSink, LutR, AggR>(sp, (uint64_t)R.sym[q], n, R.a[q], R.b[q], st, lut, (AggR *)nullptr, rs, re, h);
        if (ev & EV_EMIT) bytesL += h.len + slice_lit_len(lastBefore, rs, floor);
        if constexpr (K != 0)
        {
          const uint32_t idx = (ev >> EV_IDX_SHIFT) & 7u;
          if ((ev & EV_VALID) && known < (uint32_t)K && idx >= known)
          {
            if (ev & EV_MARG)
            { // a decision that consulted the inherited part of the table: record the query
              sensL = 1;
              const uint32_t slot = atomicAdd(&R.q.n, 1u);
              if (slot < (uint32_t)E2_NQ)
              {
                R.q.sym[slot] = (ev & EV_EMIT) ? lut_front(lut) : sym_rot((uint64_t)R.sym[q], W, rs - (R.a[q] - W));
                R.q.known[slot] = (uint8_t)known; R.q.hit[slot] = idx < (uint32_t)K ? 1 : 0;
              }
            }
            if (ev & EV_EMIT) { R.fo[known] = lut_front(lut); if (idx == (uint32_t)K) atomicAdd(&R.foMiss, 1u); known++; }
          }
        }
      }
      if (sensL) R.sens = 1;
      mine.bytes = bytesL;
    }
    else mine.bytes = 0;
    if constexpr (K != 0)
    { // exact (under the incoming table this run used) token bytes of the super-chunk
      uint64_t b = mine.bytes;
#pragma unroll
      for (int d = 16; d >= 1; d >>= 1) b += __shfl_xor_sync(0xFFFFFFFFu, b, d);
      total.bytes = b;
    }
    __syncwarp();
    if (lane == 0)
    {
      B.scSum[s] = total.cs;
      B.scBytes[s] = total.bytes - (uint64_t)W * R.foMiss; B.scTok[s] = total.ntok;
      if constexpr (K != 0)
      {
        lutagg_to(B.scAgg[s], total.agg);
        Lut fo;
#pragma unroll
        for (int i = 0; i < 7; i++) fo.s[i] = (i < K && i < (int)total.agg.m) ? R.fo[i] : 0ull;
        B.scFo[s] = fo; B.scFlags[s] = R.sens ? (uint8_t)SCF_SENS : (uint8_t)0;
        B.scQ[s] = R.q;
      }
      if (!given) { B.scIn[s] = st0; if (K) lut_to(B.scLut[s], lut0); }
      __threadfence();                                            // read by the round's last CTA
    }
    totalOut = total;
  }

  // ---- run by the last CTA of a round: scan of the super-chunk summaries, verification, finishing
  static __device__ void finish(const EncBufs &B, const AutoState &fin, const LutR &finLut, uint64_t tokBytes, uint32_t nTok)
  { // all threads call; thread 0 writes header/terminator, everybody copies a short trailing literal
    constexpr Spec sp = make_spec(W, BA, V);
    EncScalars &sc = *B.sc;
    __shared__ uint32_t fPos, fLen, fOk;
    if (B.sliceMode)
    { // a slice only reports: outgoing state and token bytes (placement happens after the size exchange)
      if (threadIdx.x == 0)
      {
        sc.tokBytes = tokBytes; sc.nTok = nTok;
        if (sc.status == ST_OK && (uint64_t)B.outBase + tokBytes + 64 > B.cap) sc.status = ST_OVERFLOW;
        SliceMsg &m = *B.msg;
        m.out = fin; if (K) lut_to(m.outLut, finLut); else lut_init(m.outLut, W);
        m.tokBytes = tokBytes; m.status = sc.status;
        uint32_t *r = B.dResult;
        r[0] = 0; r[1] = sc.status; r[2] = sc.nRuns; r[3] = sc.nSC; r[4] = sc.serialSC; r[5] = sc.innerSerial; r[6] = 0; r[7] = sc.nDirty[0];
      }
      return;
    }
    if (threadIdx.x == 0)
    {
      fOk = 0;
      sc.tokBytes = tokBytes; sc.nTok = nTok;
      const uint32_t L = B.n - fin.last;
      TokenHdr h; enc_terminator(sp, L, h);
      const uint64_t total = (uint64_t)B.outBase + tokBytes + h.len + L;
      if (sc.status == ST_OK && (total > B.cap || total >= 0xFFFFFFF0ull)) sc.status = ST_OVERFLOW;
      if (sc.status == ST_OK)
      {
        sc.total = (uint32_t)total;
        uint8_t *o = B.out;
        const uint32_t nn = B.n, tt = (uint32_t)total;
        for (int k = 0; k < 4; k++) { o[k] = (uint8_t)(nn >> (8 * k)); o[4 + k] = (uint8_t)(tt >> (8 * k)); }
        if (sp.hdr == 9) o[8] = 0;
        const uint32_t pos = (uint32_t)(B.outBase + tokBytes);
        for (uint32_t k = 0; k < h.len; k++) o[pos + k] = h.b[k];
        if (L >= MED_COPY) enc_push_copy(B, pos + h.len, fin.last, L);
        else { fPos = pos + h.len; fLen = L; fOk = 1; }
      }
      else sc.total = 0;
      uint32_t *r = B.dResult;
      r[0] = sc.status == ST_OK ? sc.total : 0; r[1] = sc.status; r[2] = sc.nRuns; r[3] = sc.nSC;
      r[4] = sc.serialSC; r[5] = sc.innerSerial; r[6] = sc.nDirty[1] | (sc.nDirty[2] << 16); r[7] = sc.nDirty[0];
    }
    __syncthreads();
    if (fOk) { const uint32_t src = fin.last; for (uint32_t k = threadIdx.x; k < fLen; k += blockDim.x) B.out[fPos + k] = B.in[src + k]; }
  }

  // would super-chunk s decide anything differently if its incoming table were `lut`?
  static __device__ __forceinline__ bool lut_sensitive(const EncBufs &B, uint32_t s, const LutR &lut)
  {
    if (!K || !(B.scFlags[s] & SCF_SENS)) return false;
    LutR fo; lut_from(fo, B.scFo[s]);
    return enc_queries_differ(B.scQ[s], fo, B.scAgg[s].m, K, lut);
  }

  // exact token bytes of super-chunk s when its incoming table is `lut`
  static __device__ __forceinline__ uint64_t sc_bytes(const EncBufs &B, uint32_t s, const LutR &lut)
  {
    uint64_t b = B.scBytes[s];
    if (K) { const uint32_t m = B.scAgg[s].m; if (m) { LutR fo; lut_from(fo, B.scFo[s]); b += (uint64_t)W * enc_fo_misses(fo, m, K, lut); } }
    return b;
  }
  // a super-chunk's summary as a scan element
  static __device__ __forceinline__ Seg load_seg(const EncBufs &B, uint32_t s, uint64_t bytes)
  {
    Seg e; e.cs = B.scSum[s];
    if (K) lutagg_from(e.agg, B.scAgg[s]); else lutagg_clear(e.agg);
    e.bytes = bytes; e.ntok = B.scTok[s];
    return e;
  }

  // ---- verification: scan of the super-chunk summaries, exact incoming state of every super-chunk, dirty marks.
  //      Returns the number of dirty super-chunks (uniform); the indices of up to FIX_LIST of them are in S.list.
  static __device__ uint32_t verify(const EncBufs &B, FixSmem &S, Seg &total)
  {
    const uint32_t nSC = B.sc->nSC;
    const int t = threadIdx.x;
    if (t == 0) { S.nDirty = 0; S.firstDirty = 0xFFFFFFFFu; S.nList = 0; }
    const uint32_t per = (nSC + FIX_T - 1) / FIX_T;
    const uint32_t lo = min(nSC, (uint32_t)t * per), hi = min(nSC, lo + per);
    // (1) states: scan of the summaries (token bytes follow later: for LUT codecs they depend on the incoming table)
    Seg mine = segsum_identity<K, AggR>();
    for (uint32_t s = lo; s < hi; s++) mine = segsum_combine(mine, load_seg(B, s, 0));
    const Seg pre = block_excl_scan(S, mine, total);              // (barriers inside: the resets above are visible)
    // (2) a LUT super-chunk whose decisions did not depend on the incoming table only takes the exact table
    AutoState st; LutR lut; enc_stream_incoming(B, W, st, lut);
    segsum_apply(st, lut, pre);
    uint32_t nd = 0, first = 0xFFFFFFFFu;
    for (uint32_t s = lo; s < hi; s++)
    {
      bool bad = false;
      if (B.scIn[s] != st) { B.scIn[s] = st; bad = true; }
      if (K)
      {
        LutR cur; lut_from(cur, B.scLut[s]);
        if (!lut_equal(cur, lut, K)) { lut_to(B.scLut[s], lut); if (lut_sensitive(B, s, lut)) bad = true; }
      }
      B.scDirty[s] = bad ? 1 : 0;
      if (bad)
      {
        if (!nd) first = s;
        nd++;
        const uint32_t slot = atomicAdd(&S.nList, 1u);
        if (slot < (uint32_t)FIX_LIST) S.list[slot] = s;
      }
      const Seg e = load_seg(B, s, 0);
      chunksum_apply(st, e.cs);
      if (K) lut_apply(lut, K, e.agg);
    }
    if (nd) { atomicAdd(&S.nDirty, nd); atomicMin(&S.firstDirty, first); }
    __syncthreads();
    return S.nDirty;
  }

  // ---- token byte offsets of the super-chunks (scLut[s] is the exact incoming table of every super-chunk), then either the
  //      finish (nothing dirty) or the exact sequential repair from the first inconsistent super-chunk
  static __device__ void offsets_and_finish(const EncBufs &B, FixSmem &S, const Seg &total, uint32_t nDirty)
  {
    EncScalars &sc = *B.sc;
    const uint32_t nSC = sc.nSC;
    const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
    const uint32_t per = (nSC + FIX_T - 1) / FIX_T;
    const uint32_t lo = min(nSC, (uint32_t)t * per), hi = min(nSC, lo + per);
    uint64_t bytes = 0;
    for (uint32_t s = lo; s < hi; s++)
    {
      B.scBase[s] = bytes;                       // relative to the thread's first super-chunk
      LutR li; if (K) lut_from(li, B.scLut[s]); else lut_init(li, W);
      bytes += sc_bytes(B, s, li);
    }
    uint64_t incB = bytes;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, incB, d); if (lane >= d) incB += o; }
    if (lane == 31) S.warpBytes[warp] = incB;
    __syncthreads();
    uint64_t preB = incB - bytes, totalBytes = 0;
#pragma unroll 1
    for (int w = 0; w < FIX_W; w++) { const uint64_t x = S.warpBytes[w]; if (w < warp) preB += x; totalBytes += x; }
    for (uint32_t s = lo; s < hi; s++) B.scBase[s] += preB;
    __syncthreads();
    AutoState fin; LutR finLut; enc_stream_incoming(B, W, fin, finLut);
    segsum_apply(fin, finLut, total);
    if (nDirty == 0) { finish(B, fin, finLut, totalBytes, total.ntok); return; }

    // exact repair: walk the super-chunks from the first inconsistent one with the exact running state; only those whose
    // assumed incoming state is wrong are re-evaluated (one warp each, one after the other)
    const uint32_t firstDirty = S.firstDirty;
    __threadfence();
    if (t == 0) { S.bcSt = B.scIn[firstDirty]; if (K) lut_from(S.bcLut, B.scLut[firstDirty]); }
    __syncthreads();
    AutoState run = S.bcSt; LutR runLut; if (K) runLut = S.bcLut; else lut_init(runLut, W);
    uint64_t runBytes = B.scBase[firstDirty];
    uint32_t runTok = 0;
    for (uint32_t s = 0; s < firstDirty; s++) runTok += B.scTok[s];   // small, uniform across threads
    for (uint32_t s = firstDirty; s < nSC; s++)
    {
      Seg tot;
      bool need = B.scDirty[s] != 0 || B.scIn[s] != run;               // uniform: every thread reads the same words
      bool lutDiff = false;
      if (K) { LutR cur; lut_from(cur, B.scLut[s]); lutDiff = !lut_equal(cur, runLut, K); }
      if (lutDiff && lut_sensitive(B, s, runLut)) need = true;
      __syncthreads();
      if (need)
      {
        if (warp == 0) { process(B, S.r[0], s, true, run, runLut, tot, true); if (lane == 0) S.bcTot = tot; }
        __syncthreads();
        tot = S.bcTot;
        if (t == 0) { B.scIn[s] = run; if (K) lut_to(B.scLut[s], runLut); B.scDirty[s] = 0; atomicAdd(&sc.serialSC, 1u); }
      }
      else
      {
        tot = load_seg(B, s, sc_bytes(B, s, runLut));
        if (t == 0 && lutDiff) lut_to(B.scLut[s], runLut);
      }
      if (t == 0) B.scBase[s] = runBytes;
      segsum_apply(run, runLut, tot);
      runBytes += tot.bytes; runTok += tot.ntok;
    }
    __syncthreads();
    finish(B, run, runLut, runBytes, runTok);
  }
};

// round 0: every super-chunk evaluated from a warmed-up state guess, one super-chunk per warp
template <int W, int BA, int V, class SymT>
__global__ void __launch_bounds__(E2_T, ((V == V_LUT3 || V == V_LUT7) && W > 1) ? 4 : 1) k_enc_auto(const EncBufs B)
{
  using C = EncCta<W, BA, V, SymT>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  typename C::Smem &S = *reinterpret_cast<typename C::Smem *>(smemRaw);
  const uint32_t nSC = B.sc->nSC;
  constexpr int SCS_PER_CTA = C::NW;
  const uint32_t sFirst = blockIdx.x * SCS_PER_CTA + (threadIdx.x >> 5);
  for (uint32_t s = sFirst; s < nSC; s += gridDim.x * SCS_PER_CTA)
  {
    typename C::Seg tot;
    AutoState d = enc_initial_state(); typename C::LutR dl; lut_init(dl, W);
    C::process(B, S.r[threadIdx.x >> 5], s, false, d, dl, tot);
  }
}

// verify / repair rounds inside ONE launch (no launch or grid-wide barrier per round).  CTA 0 is the master: verify scan over the
// super-chunk summaries -> dirty super-chunks re-run from their exact incoming state -> verify again ... until nothing is dirty
// (then offsets + finish) or the round budget is spent (then the exact sequential repair).  A round with few dirty super-chunks
// is run by the master's 16 warps alone; a round with many is dealt out in tickets of FIX_BLK super-chunks that every resident
// warp of the grid takes (the other CTAs only help: completion is counted per ticket, never per CTA, so nothing waits for a CTA
// that is not resident -- safe with plain launches on concurrent streams).  The helpers leave as soon as the master sees a
// round that it can run alone.  startDirty: the caller marked the dirty super-chunks itself (slices: the true incoming state of
// the slice arrived).
constexpr uint32_t FIX_BLK = 32;            // super-chunks per ticket
constexpr uint32_t FIX_SOLO = 48;           // at most this many dirty super-chunks: the master runs the round alone

template <int W, int BA, int V, class SymT>
__device__ __forceinline__ void enc_fix_tickets(const EncBufs &B, typename EncCta<W, BA, V, SymT>::WarpRecs &R, uint32_t nSC)
{ // warp-level: take tickets of the current round until none is left
  using C = EncCta<W, BA, V, SymT>;
  EncScalars &sc = *B.sc;
  const int lane = threadIdx.x & 31;
  const uint32_t nTickets = (nSC + FIX_BLK - 1) / FIX_BLK;
  for (;;)
  {
    uint32_t k = 0;
    if (lane == 0) k = atomicAdd(&sc.fixTicket, 1u);
    k = __shfl_sync(0xFFFFFFFFu, k, 0);
    if (k >= nTickets) break;
    const uint32_t s0 = k * FIX_BLK, s1 = min(nSC, s0 + FIX_BLK);
    for (uint32_t s = s0; s < s1; s++)
    {
      if (!__ldcg(B.scDirty + s)) continue;
      AutoState g; typename C::LutR gl;
      { // written by the master in this launch: read through L2
        const uint32_t *p = reinterpret_cast<const uint32_t *>(&B.scIn[s]); uint32_t *q = reinterpret_cast<uint32_t *>(&g);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(AutoState) / 4); i++) q[i] = __ldcg(p + i);
        if (C::K)
        {
          Lut tmp; const uint32_t *pl = reinterpret_cast<const uint32_t *>(&B.scLut[s]); uint32_t *ql = reinterpret_cast<uint32_t *>(&tmp);
#pragma unroll
          for (int i = 0; i < (int)(sizeof(Lut) / 4); i++) ql[i] = __ldcg(pl + i);
          lut_from(gl, tmp);
        }
        else lut_init(gl, W);
      }
      typename C::Seg tot;
      C::process(B, R, s, true, g, gl, tot, true);
    }
    __syncwarp();
    if (lane == 0) { __threadfence(); atomicAdd(&sc.fixDone, 1u); }
  }
}

template <int W, int BA, int V, class SymT>
__global__ void __launch_bounds__(FIX_T, 1) k_enc_fix(const EncBufs B, int startDirty, int maxRounds, int followHops)
{
  using C = EncCta<W, BA, V, SymT>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  typename C::FixSmem &S = *reinterpret_cast<typename C::FixSmem *>(smemRaw);
  EncScalars &sc = *B.sc;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (blockIdx.x != 0)
  { // helper CTA: wait for grid rounds, leave when told
    uint32_t seen = 0;
    for (;;)
    {
      uint32_t cmd = 0, rnd = 0;
      if (lane == 0) { cmd = ld_volatile_u32g(&sc.fixCmd); rnd = ld_volatile_u32g(&sc.fixRound); }
      cmd = __shfl_sync(0xFFFFFFFFu, cmd, 0); rnd = __shfl_sync(0xFFFFFFFFu, rnd, 0);
      if (cmd) break;
      if (rnd != seen) { seen = rnd; __threadfence(); enc_fix_tickets<W, BA, V, SymT>(B, S.r[warp], ld_volatile_u32g(&sc.nSC)); }
      else __nanosleep(200);
    }
    return;
  }
  const uint32_t nSC = sc.nSC;
  typename C::Seg total;
  uint32_t nDirty = 0;
  int round = 0;
  bool helpers = gridDim.x > 1;
  if (startDirty)
  { // dirty marks come from the caller: count them (the list is rebuilt by the flags path below)
    if (t == 0) { S.nList = 0; S.nDirty = 0; S.firstDirty = 0xFFFFFFFFu; }
    __syncthreads();
    for (uint32_t s = t; s < nSC; s += FIX_T)
      if (B.scDirty[s]) { atomicAdd(&S.nDirty, 1u); const uint32_t slot = atomicAdd(&S.nList, 1u); if (slot < (uint32_t)FIX_LIST) S.list[slot] = s; }
    __syncthreads();
    nDirty = S.nDirty;
    if (nDirty == 0) nDirty = C::verify(B, S, total);
  }
  else nDirty = C::verify(B, S, total);
  if (t == 0) sc.nDirty[0] = nDirty;
  while (nDirty != 0 && round < maxRounds)
  {
    if (helpers && nDirty <= FIX_SOLO)
    { // from here on the master runs alone: release the helpers
      if (t == 0) { __threadfence(); st_volatile_u32g(&sc.fixCmd, 1u); }
      helpers = false;
    }
    if (helpers)
    { // grid round: states and dirty flags are in global memory; publish the round, take tickets, wait for all tickets
      __syncthreads();
      if (t == 0) { sc.fixTicket = 0; sc.fixDone = 0; __threadfence(); st_volatile_u32g(&sc.fixRound, (uint32_t)round + 1u); }
      __syncthreads();
      enc_fix_tickets<W, BA, V, SymT>(B, S.r[warp], nSC);
      if (t == 0)
      {
        const uint32_t nTickets = (nSC + FIX_BLK - 1) / FIX_BLK;
        while (ld_volatile_u32g(&sc.fixDone) < nTickets) __nanosleep(100);
      }
      __syncthreads();
      __threadfence();                                  // helpers on other SMs rewrote summaries: nothing stale may be read from this SM's L1
    }
    else
    { // solo round: one dirty super-chunk per warp -- and the warp FOLLOWS what its re-run changed for a few super-chunks: while the
      // outgoing state differs from the one the verify scan handed to the next super-chunk, that one takes the new state (re-run if
      // its decisions depend on the difference, passed through otherwise), until the states coalesce, the next listed super-chunk
      // is reached (its owner started from the scan's state: the next verify sorts that out) or the hop budget is spent.
      const uint32_t nl = min(S.nList, (uint32_t)FIX_LIST);
      const bool listed = S.nList <= (uint32_t)FIX_LIST;
      __syncthreads();
      for (uint32_t i = warp; i < (listed ? nl : nSC); i += FIX_W)
      {
        uint32_t s = listed ? S.list[i] : i;
        if (!listed && !B.scDirty[s]) continue;
        AutoState xs = B.scIn[s]; typename C::LutR xl; if (C::K) lut_from(xl, B.scLut[s]); else lut_init(xl, W);
        for (int hop = 0; hop < followHops; hop++)
        {
          typename C::Seg tot;
          C::process(B, S.r[warp], s, true, xs, xl, tot, true);
          if (hop + 1 == followHops) break;
          segsum_apply(xs, xl, tot);                                   // (xs, xl): the new outgoing state
          bool rerun = false;
          for (s++; s < nSC; s++)
          {
            if (B.scDirty[s]) break;                                   // listed this round: not ours
            const AutoState os = B.scIn[s];
            bool lutDiff = false;
            if (C::K) { typename C::LutR ol; lut_from(ol, B.scLut[s]); lutDiff = !lut_equal(ol, xl, C::K); }
            if (os == xs && !lutDiff) break;                           // coalesced with what the scan assumed
            rerun = (os != xs) || (lutDiff && C::lut_sensitive(B, s, xl));
            __syncwarp();
            if (lane == 0) { B.scIn[s] = xs; if (C::K) lut_to(B.scLut[s], xl); }
            __syncwarp();
            if (rerun) break;
            const typename C::Seg e = C::load_seg(B, s, 0);            // decisions unaffected: the state passes through
            segsum_apply(xs, xl, e);
          }
          if (!rerun) break;
        }
      }
      __threadfence_block();
      __syncthreads();
    }
    round++;
    nDirty = C::verify(B, S, total);
  }
  if (helpers && t == 0) { __threadfence(); st_volatile_u32g(&sc.fixCmd, 1u); }
  if (t == 0) { sc.nDirty[1] = (uint32_t)round; sc.nDirty[2] = nDirty; }
  C::offsets_and_finish(B, S, total, nDirty);
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
constexpr uint32_t E3_PIECE = 16;         // bytes per flattened literal piece
// stream bytes of a super-chunk that can be staged in shared memory.  Wider symbols mean fewer but longer tokens: the 512
// tokens of a 32-bit super-chunk carry ~32 KB on the DCT stream (measured: E3 70 -> 35 us with the 64 KB stage); 24-bit
// super-chunks fit 32 KB, and from 48 bits on most literals are long enough for the grid-wide copy kernel anyway
HSRLE_HDC uint32_t e3_stage(int W) { return W == 4 ? 65536u : 32768u; }
HSRLE_HDC uint32_t e3_npiece(int W) { return e3_stage(W) / E3_PIECE + E3_NDESC + 8; }

template <int W, int BA, int V, class SymT> struct EncEmitSmem
{
  using C = EncCta<W, BA, V, SymT>;
  alignas(16) uint8_t stage[e3_stage(W) + 32];
  uint32_t a[C::NSLOT], b[C::NSLOT];
  SymT sym[C::NSLOT];
  EmitDesc desc[E3_NDESC];
  uint32_t pieceOff[E3_NDESC + 1];
  uint16_t pieceDesc[e3_npiece(W)];
  unsigned long long warpTot[E2_T / 32];
  uint32_t warpTot32[E2_T / 32];
  uint32_t nDesc;
};

// E3: per super-chunk, (1) token byte counts per thread -> block scan -> stream position of every chunk,
// (2) token headers and literals are assembled in a shared-memory image of the super-chunk's stream segment
// (literals gathered as 16-byte pieces, one piece per thread and step, wide source loads), (3) the image is
// flushed with 16-byte aligned, fully coalesced stores.  Segments that do not fit the stage (long literals)
// take the direct path: headers straight to the stream, literals >= MED_COPY to the grid-wide copy kernel.
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
  const uint32_t nSC = sc.nSC, nRuns = sc.nRuns, n = B.n, floor = B.sliceLo, endShift = sc.endShift;
  const SymT *__restrict__ runSym = reinterpret_cast<const SymT *>(B.runSym);
  const uint8_t *__restrict__ in = B.in;
  uint8_t *__restrict__ out = B.out;
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
      S.a[q] = B.runA[lo + j]; S.b[q] = B.runB[lo + j + endShift]; S.sym[q] = runSym[lo + j];
    }
    __syncthreads();
    const int j0 = E2_CH + t * E2_CH, j1 = min(j0 + E2_CH, E2_CH + (int)cnt);
    const bool active = j0 < j1;
    using LutR = typename C::LutR;
    using AggR = typename C::AggR;
    AutoState st0 = enc_initial_state(); LutR lut0; lut_init(lut0, W);
    if (active)
    {
      st0 = B.cIn[s * E2_T + t];
      if (K) { Lut g0 = B.cLut[s * E2_T + t]; enc_chunk_lut(g0, B.cKnown[s * E2_T + t], K, B.scLut[s]); lut_from(lut0, g0); }
    }
    // pass 1: bytes of my chunk
    unsigned long long mine = 0;
    if (active)
    {
      AutoState st = st0; LutR lut = lut0;
      for (int j = j0; j < j1; j++)
      {
        const int q = rec_slot(j);
        uint32_t rs, re; CountSink h;
        const uint32_t lastBefore = st.last;
        const uint32_t ev = enc_eval_t<CountSink, LutR, AggR>(sp, (uint64_t)S.sym[q], n, S.a[q], S.b[q], st, lut, (AggR *)nullptr, rs, re, h);
        if (ev & EV_EMIT) mine += h.len + slice_lit_len(lastBefore, rs, floor);
      }
    }
    unsigned long long inc = mine;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) S.warpTot[warp] = inc;
    __syncthreads();
    unsigned long long pre = 0, segLen = 0;
#pragma unroll
    for (int w = 0; w < E2_T / 32; w++) { const unsigned long long x = S.warpTot[w]; if (w < warp) pre += x; segLen += x; }
    const uint64_t seg0 = (uint64_t)B.outBase + B.scBase[s];
    uint64_t pos = seg0 + pre + (inc - mine);
    const bool staged = segLen <= e3_stage(W);         // uniform
    const uint32_t shift = (uint32_t)(seg0 & 15);      // keeps stage image and stream equally aligned
    // pass 2: headers, literal descriptors
    if (active)
    {
      AutoState st = st0; LutR lut = lut0;
      for (int j = j0; j < j1; j++)
      {
        const int q = rec_slot(j);
        uint32_t rs, re; PtrSink h;
        h.p = staged ? (S.stage + shift + (uint32_t)(pos - seg0)) : (out + pos);
        const uint32_t lastBefore = st.last;
        const uint32_t ev = enc_eval_t<PtrSink, LutR, AggR>(sp, (uint64_t)S.sym[q], n, S.a[q], S.b[q], st, lut, (AggR *)nullptr, rs, re, h);
        if (!(ev & EV_EMIT)) continue;
        const uint32_t lit = slice_lit_len(lastBefore, rs, floor), litSrc = slice_lit_src(lastBefore, floor);
        if (B.sliceMode && pos == B.outBase)
        { // the slice's first token: whoever holds the start of its literal places this header (hsrle_slice.cuh)
          SliceMsg &m = *B.msg;
          m.hasEmit = 1; m.firstHdrLen = h.len; m.firstS = rs; m.firstLast = lastBefore;
          for (uint32_t k = 0; k < 24; k++) m.firstHdr[k] = k < h.len ? h.p[k] : 0;
        }
        if (lit)
        {
          if (staged) { EmitDesc d; d.dst = shift + (uint32_t)(pos - seg0) + h.len; d.src = litSrc; d.len = lit; S.desc[atomicAdd(&S.nDesc, 1u)] = d; }
          else if (lit >= MED_COPY) enc_push_copy(B, (uint32_t)(pos + h.len), litSrc, lit);
          else { EmitDesc d; d.dst = (uint32_t)(pos + h.len); d.src = litSrc; d.len = lit; S.desc[atomicAdd(&S.nDesc, 1u)] = d; }
        }
        pos += h.len + lit;
      }
    }
    __syncthreads();
    // flat piece list: exclusive scan of ceil(len/16) over the descriptors (E2_CH per thread)
    const uint32_t nd = S.nDesc;
    uint32_t loc[E2_CH];
    uint32_t sum = 0;
#pragma unroll
    for (int k = 0; k < E2_CH; k++)
    {
      const uint32_t d = t * E2_CH + k;
      loc[k] = sum;
      if (d < nd) sum += (S.desc[d].len + E3_PIECE - 1) / E3_PIECE;
    }
    uint32_t inc32 = sum;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc32, d); if (lane >= d) inc32 += o; }
    if (lane == 31) S.warpTot32[warp] = inc32;
    __syncthreads();
    uint32_t pre32 = 0, totalPieces = 0;
#pragma unroll
    for (int w = 0; w < E2_T / 32; w++) { const uint32_t x = S.warpTot32[w]; if (w < warp) pre32 += x; totalPieces += x; }
    const uint32_t base32 = pre32 + inc32 - sum;
#pragma unroll
    for (int k = 0; k < E2_CH; k++)
    {
      const uint32_t d = t * E2_CH + k;
      if (d < nd)
      {
        const uint32_t o = base32 + loc[k], np = (S.desc[d].len + E3_PIECE - 1) / E3_PIECE;
        S.pieceOff[d] = o;
        if (staged) for (uint32_t q = 0; q < np; q++) S.pieceDesc[o + q] = (uint16_t)d;
      }
    }
    __syncthreads();
    if (staged)
    {
      // four pieces per thread and step: the (up to 20) source words of a step are in flight together
      constexpr int E3_MLP = 4;
      for (uint32_t i0 = t; i0 < totalPieces; i0 += E2_T * E3_MLP)
      {
        uint32_t w[E3_MLP][5], dstOff[E3_MLP], len[E3_MLP], sb[E3_MLP];
#pragma unroll
        for (int u = 0; u < E3_MLP; u++)
        {
          const uint32_t i = i0 + u * E2_T;
          len[u] = 0; dstOff[u] = 0; sb[u] = 0;
          const uint32_t *sw = reinterpret_cast<const uint32_t *>(in);
          uint32_t need = 0;
          if (i < totalPieces)
          {
            const uint32_t a = S.pieceDesc[i];
            const EmitDesc e = S.desc[a];
            const uint32_t off = (i - S.pieceOff[a]) * E3_PIECE;
            len[u] = min(E3_PIECE, e.len - off);
            const uint8_t *src = in + e.src + off;
            sb[u] = (uint32_t)((uintptr_t)src & 3);
            sw = reinterpret_cast<const uint32_t *>((uintptr_t)src & ~(uintptr_t)3);
            need = sb[u] + len[u];                    // bytes needed from the aligned word stream
            dstOff[u] = e.dst + off;
          }
#pragma unroll
          for (int k = 0; k < 5; k++) w[u][k] = ((uint32_t)k * 4 < need) ? __ldg(sw + k) : 0u;
        }
#pragma unroll
        for (int u = 0; u < E3_MLP; u++)
        {
          uint32_t o[4];
#pragma unroll
          for (int k = 0; k < 4; k++) o[k] = __funnelshift_r(w[u][k], w[u][k + 1], sb[u] * 8);
          uint8_t *dst = S.stage + dstOff[u];
#pragma unroll
          for (uint32_t k = 0; k < E3_PIECE; k++) if (k < len[u]) dst[k] = (uint8_t)(o[k >> 2] >> (8 * (k & 3)));
        }
      }
      __syncthreads();
      // flush the image: bytes [shift, shift + segLen) of the stage go to stream bytes [seg0, seg0 + segLen)
      uint8_t *gbase = out + (seg0 - shift);
      const uint32_t end = shift + (uint32_t)segLen;
      const uint32_t headEnd = min(end, 16u);
      if ((uint32_t)t >= shift && (uint32_t)t < headEnd) gbase[t] = S.stage[t];
      const uint32_t vEnd = end >> 4;               // full vectors are 1 .. vEnd-1
      for (uint32_t v = 1 + t; v < vEnd; v += E2_T) reinterpret_cast<uint4 *>(gbase)[v] = reinterpret_cast<const uint4 *>(S.stage)[v];
      const uint32_t tail0 = max(vEnd << 4, 16u);
      if (tail0 + t < end) gbase[tail0 + t] = S.stage[tail0 + t];
    }
    else
    {
      for (uint32_t i = t; i < totalPieces; i += E2_T)
      {
        uint32_t a = 0, b = nd - 1;           // largest d with pieceOff[d] <= i
        while (a < b) { const uint32_t m = (a + b + 1) >> 1; if (S.pieceOff[m] <= i) a = m; else b = m - 1; }
        const EmitDesc e = S.desc[a];
        const uint32_t off = (i - S.pieceOff[a]) * E3_PIECE;
        const uint32_t len = min(E3_PIECE, e.len - off);
        const uint8_t *__restrict__ src = in + e.src + off;
        uint8_t *__restrict__ dst = out + e.dst + off;
        uint8_t tmp[E3_PIECE];
#pragma unroll
        for (uint32_t k = 0; k < E3_PIECE; k++) if (k < len) tmp[k] = src[k];
#pragma unroll
        for (uint32_t k = 0; k < E3_PIECE; k++) if (k < len) dst[k] = tmp[k];
      }
    }
  }
}

// ================================================================================================
// grid-wide copy of the long literals
constexpr uint32_t BIG_PIECE = 16384;
static __global__ void __launch_bounds__(256) k_enc_copy_big(const EncBufs B)
{
  if (B.sc->status != ST_OK) return;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t nMed = B.sc->nMed;
  const uint32_t nWarps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t i = blockIdx.x * (blockDim.x >> 5) + warp; i < nMed; i += nWarps)
  {
    const CopyDesc cd = B.medList[i];
    copy_bytes_warp(B.out + cd.dst, B.in + cd.src, cd.len, lane);
  }
  const uint32_t nBig = B.sc->nBig;
  constexpr uint32_t SUB = BIG_PIECE / 8;    // one sub-piece per warp
  // the pieces of all huge literals are dealt round-robin over the CTAs as ONE sequence (a stream of 64-KiB..1-MiB literals
  // has 4..64 pieces each: starting every literal at CTA 0 left most of the grid idle -- 88 MB in such literals took
  // 1.3 ms); rot = pieces of the literals before this one, modulo the grid
  uint32_t rot = 0;
  for (uint32_t i = 0; i < nBig; i++)
  {
    const CopyDesc cd = B.bigList[i];
    const uint32_t nPieces = (cd.len + BIG_PIECE - 1) / BIG_PIECE;
    const uint32_t first = (blockIdx.x + gridDim.x - rot) % gridDim.x;
    rot = (rot + nPieces) % gridDim.x;
    for (uint32_t pc = first; pc < nPieces; pc += gridDim.x)
    {
      const uint32_t off = pc * BIG_PIECE + warp * SUB;
      if (off >= cd.len) continue;
      const uint32_t len = min(SUB, cd.len - off);
      copy_bytes_warp(B.out + cd.dst + off, B.in + cd.src + off, len, lane);
    }
  }
}

} // namespace hsrle
