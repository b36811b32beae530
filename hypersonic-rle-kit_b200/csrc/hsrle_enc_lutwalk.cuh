// hsrle_enc_lutwalk.cuh -- 8-bit LUT codecs (rle8_3symlut / rle8_7symlut): the table at every super-chunk start, BEFORE the
// speculative automaton (E2) runs.
//
// Why: the table of these codecs (src/rleX_Xsl.h:114-264, restated in hsrle_core.cuh: enc_eval_t, K > 0) keeps the last K
// distinct symbols of EMITTED runs.  In data dominated by one symbol (DCT coefficient planes: 1.26 M candidate runs, 1843 of
// them not of the zero byte) two of the three entries are thousands of records old, and whether a short run of a rare symbol
// is emitted depends on it still being in the table -- a dependency chain that a state guess from the last dozen records can
// never see.  E2's verify / re-run rounds then repair one link of that chain per round (58 and 90 rounds on the benchmark
// input, 3.7 and 10.2 ms).  The table is cheap to get exactly, though, because it only changes where the run symbol changes:
//
//   * a STRETCH is a maximal sequence of consecutive records with the same symbol z.  After the first emitted record of a
//     stretch the table has z in front and nothing else about it changes until the stretch ends.
//   * a record is CERTAIN when it is emitted whatever the automaton state is (symbol absent from the table, literal before it
//     as long as it can be -- or exactly known because the record before it is certain as well).  After a certain record the
//     rest of the stretch evolves independently of anything before it.
//   * so what a stretch with a certain record does to the state is: table <- touch(table, z), `last` <- a value that follows
//     from the records after its last certain one (k_enc_lut_stretch computes it per stretch end, in parallel).  Only stretches
//     without a certain record (the single short runs of rare symbols) have to be stepped through with the real state.
//
// k_enc_lut_stretch   one thread per 16 records: stretch boundaries, and per boundary the descriptor of the stretch that ends there.
// k_enc_lut_walk      one CTA: the descriptors in record order, ONE thread walks them with the exact (last, table) -- a few
//                     thousand steps instead of 1.26 M -- and the table at the start of every super-chunk goes to scGuess[].
//
// scGuess is a GUESS as far as E2 is concerned (a super-chunk that starts inside the uncertain head of a stretch gets the table
// as if the head had emitted): E2's verify scan checks every super-chunk against the composed exact state as before, so the
// result is exact regardless; the walk only removes the long chains.  Inputs with more than LW_CAP stretch boundaries skip the
// walk (sc.lwOk stays 0): many different run symbols flush the table quickly, and the plain guess converges in a few rounds.
#pragma once
#include "hsrle_enc.cuh"

namespace hsrle {

constexpr int LW_T = 256;                 // threads per block of k_enc_lut_stretch / k_enc_lut_walk
constexpr int LW_PER = 16;                // consecutive records per thread
constexpr int LW_BLK = LW_T * LW_PER;     // records per block
constexpr int LW_BACK = 64;               // records a boundary looks back for a certain record of the stretch that ends there
constexpr uint32_t LW_CAP = 16384;        // stretch boundaries the walk takes
constexpr int LW_TILE = 1024;             // descriptors staged in shared memory per walk tile
constexpr uint32_t LWD_CERT = 0x80000000u, LWD_FIRST = 0x40000000u;

// boundary at record j (the stretch [jPrev, j) ends, a stretch of symbol `info & 0xFF` starts)
struct LwDesc
{
  uint32_t j;
  uint32_t lastOut;     // LWD_CERT: `last` after the stretch that ends here
  uint32_t info;        // LWD_CERT | LWD_FIRST (j == 0: nothing ends) | records of the ended stretch << 8 (no certain one) | new symbol
  uint32_t a, b;        // the ended stretch's last record (all of it when it has one record)
};

// emit rule of the 8-bit LUT codecs for a run of `cnt` bytes whose literal distance is rng = s - last + 2 (enc_eval_t, K > 0, W == 1)
template <int V> HSRLE_HD bool lw_emit(uint32_t cnt, uint32_t rng, bool miss)
{
  constexpr Spec sp = make_spec(1, 1, V);
  const uint32_t TR = (1u << sp.RB) - 1;
  const uint32_t stored = cnt - 1;
  const uint32_t pen = (rng <= 0xFFFFFu ? (rng <= TR ? 0u : 2u) : 4u) + (stored <= 0xFFFFFu ? (stored <= 127u ? 0u : 2u) : 4u) + (miss ? 1u : 0u);
  return cnt >= (uint32_t)sp.LONG || cnt >= 3 + pen;
}
// emitted whatever the state: symbol not in the table, `last` as far back as it can be (0)
template <int V> HSRLE_HD bool lw_cert0(uint32_t a, uint32_t b) { return lw_emit<V>(b - a + 1, a + 1, true); }

#ifdef __CUDACC__
template <int V>
__global__ void __launch_bounds__(LW_T) k_enc_lut_stretch(const EncBufs B)
{
  EncScalars &sc = *B.sc;
  const uint32_t nRuns = sc.nRuns;
  const uint32_t nBlk = (nRuns + LW_BLK - 1) / LW_BLK;
  const uint32_t *__restrict__ sym = reinterpret_cast<const uint32_t *>(B.runSym);
  const uint32_t *__restrict__ ra = B.runA, *__restrict__ rb = B.runB;
  __shared__ uint32_t warpCnt[LW_T / 32];
  __shared__ uint32_t sBase;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (uint32_t blk = blockIdx.x; blk < nBlk; blk += gridDim.x)
  {
    const uint32_t j0 = blk * LW_BLK + (uint32_t)t * LW_PER;
    uint32_t s[LW_PER];
    uint32_t mask = 0;
    if (j0 < nRuns)
    {
#pragma unroll
      for (int q = 0; q < LW_PER / 4; q++)
      {
        const uint32_t jq = j0 + 4 * q;
        if (jq + 3 < nRuns)
        {
          const uint4 v = *reinterpret_cast<const uint4 *>(sym + jq);
          s[4 * q] = v.x; s[4 * q + 1] = v.y; s[4 * q + 2] = v.z; s[4 * q + 3] = v.w;
        }
        else
        { // the group that holds the last record: word by word, nothing past it is read
#pragma unroll
          for (int i = 0; i < 4; i++) s[4 * q + i] = (jq + i < nRuns) ? sym[jq + i] : 0u;
        }
      }
      uint32_t prev = j0 ? sym[j0 - 1] : ~s[0];
#pragma unroll
      for (int i = 0; i < LW_PER; i++) { if (j0 + i < nRuns && s[i] != prev) mask |= 1u << i; prev = s[i]; }
    }
    // ordered slot of every boundary of the block
    const uint32_t cnt = __popc(mask);
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warpCnt[warp] = inc;
    __syncthreads();
    uint32_t off = inc - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < LW_T / 32; w++) { const uint32_t x = warpCnt[w]; if (w < warp) off += x; total += x; }
    if (t == 0)
    {
      const uint32_t base = total ? atomicAdd(&sc.lwCount, total) : 0u;
      if (base + total > LW_CAP) sc.lwOverflow = 1;
      B.lwBlkOff[blk] = base; B.lwBlkCnt[blk] = total;
      sBase = base;
    }
    __syncthreads();
    const uint32_t base = sBase;
    __syncthreads();
    if (base + total > LW_CAP) continue;
    while (mask)
    {
      const int i = __ffs(mask) - 1; mask &= mask - 1;
      const uint32_t j = j0 + i;
      LwDesc d; d.j = j; d.lastOut = 0; d.a = 0; d.b = 0; d.info = sym[j] & 0xFFu;
      if (j == 0) d.info |= LWD_CERT | LWD_FIRST;
      else
      { // the stretch of symbol zPrev that ends at j-1: look back for its last certain record
        const uint32_t zPrev = sym[j - 1];
        uint32_t k = j - 1, a = ra[k], b = rb[k];
        d.a = a; d.b = b;
        int found = 0, steps = 0;
        for (;;)
        {
          bool cert = lw_cert0<V>(a, b);
          uint32_t pa = 0, pb = 0;
          if (k > 0) { pa = ra[k - 1]; pb = rb[k - 1]; if (!cert && lw_cert0<V>(pa, pb)) cert = lw_emit<V>(b - a + 1, a - pb + 1, true); }
          if (cert) { found = 1; break; }
          if (k == 0 || sym[k - 1] != zPrev) break;                  // k: first record of the stretch, none certain
          if (++steps >= LW_BACK) { sc.lwOverflow = 1; break; }
          k--; a = pa; b = pb;
        }
        if (found)
        { // after the certain record: z in front of the table, `last` exact -- the rest of the stretch follows
          uint32_t last = b;
          for (uint32_t r = k + 1; r < j; r++) { const uint32_t xa = ra[r], xb = rb[r]; if (lw_emit<V>(xb - xa + 1, xa - last + 1, false)) last = xb; }
          d.lastOut = last; d.info |= LWD_CERT;
        }
        else d.info |= (j - k) << 8;
      }
      B.lwPool[base + off] = d; off++;
    }
  }
}

// the table in registers, one entry each (the walk is one dependent chain: latency per step is all that counts, and a
// move-to-front over K registers is a prefix-OR plus K selects -- a third of the latency of the packed-word version)
template <int K> struct LwRegs
{
  uint32_t r[K];
  __device__ __forceinline__ void from(uint64_t v) {
#pragma unroll
    for (int i = 0; i < K; i++) r[i] = (uint32_t)(v >> (8 * i)) & 0xFFu; }
  __device__ __forceinline__ uint64_t pack() const { uint64_t v = 0;
#pragma unroll
    for (int i = 0; i < K; i++) v |= (uint64_t)r[i] << (8 * i);
    return v; }
  __device__ __forceinline__ void store(unsigned long long *dst) const { uint8_t *p = reinterpret_cast<uint8_t *>(dst);
#pragma unroll
    for (int i = 0; i < K; i++) p[i] = (uint8_t)r[i]; }
  // z to the front when `always`, or when `cond` and z is in the table; returns whether it moved
  __device__ __forceinline__ bool touch(uint32_t z, bool always, bool cond)
  {
    bool hit = false;
#pragma unroll
    for (int i = 0; i < K; i++) hit = hit || (r[i] == z);
    const bool go = always || (cond && hit);
    bool found = false; uint32_t carry = z;
#pragma unroll
    for (int i = 0; i < K; i++) { const uint32_t cur = r[i]; r[i] = (go && !found) ? carry : cur; carry = cur; found = found || (cur == z); }
    return go;
  }
  __device__ __forceinline__ bool has(uint32_t z) const { bool hit = false;
#pragma unroll
    for (int i = 0; i < K; i++) hit = hit || (r[i] == z);
    return hit; }
};

enum : uint32_t { LWS_NONE = 0, LWS_TOUCH = 1, LWS_IFHIT = 2, LWS_SLOW = 3 };

template <int V>
__global__ void __launch_bounds__(LW_T, 1) k_enc_lut_walk(const EncBufs B)
{
  constexpr Spec sp = make_spec(1, 1, V);
  constexpr int K = sp.K;
  EncScalars &sc = *B.sc;
  const uint32_t nRuns = sc.nRuns, D = sc.lwCount, nSC = sc.nSC;
  if (sc.lwOverflow || D > LW_CAP || sc.status != ST_OK) return;                 // lwOk stays 0
  const uint32_t nBlk = (nRuns + LW_BLK - 1) / LW_BLK;
  __shared__ uint2 step[LW_TILE];                 // .x = kind | zPrev << 8, .y = `last` when the step emits
  __shared__ unsigned long long lb[LW_TILE];      // table at the start of the stretch that begins at boundary k
  __shared__ uint32_t jArr[LW_TILE];
  __shared__ uint8_t zNew[LW_TILE];
  __shared__ uint32_t slowBits[LW_TILE / 32];
  __shared__ uint32_t warpCnt[LW_T / 32];
  __shared__ uint32_t carry;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  // (1) the blocks' descriptor lists in record order
  if (t == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nBlk; b0 += LW_T)
  {
    const uint32_t blk = b0 + t;
    const uint32_t cnt = blk < nBlk ? B.lwBlkCnt[blk] : 0u;
    uint32_t inc = cnt;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
    if (lane == 31) warpCnt[warp] = inc;
    __syncthreads();
    uint32_t off = carry + inc - cnt, total = 0;
#pragma unroll
    for (int w = 0; w < LW_T / 32; w++) { const uint32_t x = warpCnt[w]; if (w < warp) off += x; total += x; }
    if (cnt)
    {
      const LwDesc *src = B.lwPool + B.lwBlkOff[blk];
      for (uint32_t i = 0; i < cnt; i++) B.lwOrd[off + i] = src[i];
    }
    __syncthreads();
    if (t == 0) carry += total;
    __syncthreads();
  }
  // (2) walk, a tile of boundaries at a time
  uint32_t last = 0;
  LwRegs<K> L; { LutB l0; lut_init(l0, 1); L.from(l0.v); }
  for (uint32_t tb = 0; tb < D; tb += LW_TILE)
  {
    const uint32_t nt = min((uint32_t)LW_TILE, D - tb);
    // everything about a step that does not depend on the running state, by all threads
    for (uint32_t i0 = (uint32_t)warp * 32; i0 < nt; i0 += LW_T)
    {
      const uint32_t i = i0 + lane;
      uint32_t kind = LWS_NONE;
      if (i < nt)
      {
        const LwDesc d = B.lwOrd[tb + i];
        uint32_t zPrev = 0, newLast = d.lastOut;
        if (!(d.info & LWD_FIRST))
        {
          const LwDesc dp = B.lwOrd[tb + i - 1];
          zPrev = dp.info & 0xFFu;
          if (d.info & LWD_CERT) kind = LWS_TOUCH;
          else if (((d.info >> 8) & 0xFFu) == 1 && (dp.info & LWD_CERT) && !(dp.info & LWD_FIRST))
          { // one uncertain record right after a stretch whose `last` is known: its decision is a function of the table alone
            const uint32_t cnt = d.b - d.a + 1, rng = d.a - dp.lastOut + 1;
            newLast = d.b;
            kind = lw_emit<V>(cnt, rng, true) ? LWS_TOUCH : (lw_emit<V>(cnt, rng, false) ? LWS_IFHIT : LWS_NONE);
          }
          else kind = LWS_SLOW;
        }
        step[i] = make_uint2(kind | (zPrev << 8), newLast);
        jArr[i] = d.j; zNew[i] = (uint8_t)(d.info & 0xFFu);
        lb[i] = 0;
      }
      const uint32_t sb = __ballot_sync(0xFFFFFFFFu, kind == LWS_SLOW);
      if (lane == 0) slowBits[i0 >> 5] = sb;
    }
    __syncthreads();
    if (t == 0)
    {
      uint32_t k = 0;
      while (k < nt)
      { // next step that needs the slow path (bit scan over the tile's flags), then a branch-free run up to it
        uint32_t e = nt;
        for (uint32_t w = k >> 5; w < (nt + 31) / 32; w++)
        {
          uint32_t m = slowBits[w];
          if (w == (k >> 5)) m &= ~0u << (k & 31);
          if (m) { e = min(nt, w * 32 + (uint32_t)__ffs(m) - 1); break; }
        }
#pragma unroll 4
        for (; k < e; k++)
        {
          const uint2 st = step[k];
          const uint32_t zPrev = (st.x >> 8) & 0xFFu;
          const bool go = L.touch(zPrev, (st.x & LWS_TOUCH) != 0, (st.x & LWS_IFHIT) != 0);
          last = go ? st.y : last;
          L.store(&lb[k]);
        }
        if (k < nt)
        { // stretch without a certain record whose `last` depends on what came before: step through its records
          const uint32_t zPrev = (step[k].x >> 8) & 0xFFu;
          const LwDesc d = B.lwOrd[tb + k];
          const uint32_t cnt = (d.info >> 8) & 0xFFu;
          for (uint32_t r = d.j - cnt; r < d.j; r++)
          {
            uint32_t a = d.a, b = d.b;
            if (r + 1 != d.j) { a = B.runA[r]; b = B.runB[r]; }
            const bool hit = L.has(zPrev);
            if (lw_emit<V>(b - a + 1, a - last + 1, !hit)) { L.touch(zPrev, true, false); last = b; }
          }
          L.store(&lb[k]);
          k++;
        }
      }
    }
    __syncthreads();
    // (3) super-chunks that start inside this tile's stretches
    const uint32_t jFirst = jArr[0];
    const uint32_t jNext = (tb + nt < D) ? B.lwOrd[tb + nt].j : nRuns;
    const uint32_t sLo = (jFirst + E2_SCR - 1) / E2_SCR, sHi = min(nSC, (jNext + E2_SCR - 1) / E2_SCR);
    for (uint32_t s = sLo + t; s < sHi; s += LW_T)
    {
      const uint32_t r = s * E2_SCR;
      uint32_t lo = 0, hi = nt;                                                // last boundary with j <= r
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (jArr[mid] <= r) lo = mid; else hi = mid; }
      LutB g; g.v = lb[lo];
      if (jArr[lo] != r) { const uint32_t z = zNew[lo]; lut_touch(g, K, lut_find(g, K, z), z); }
      B.scGuess[s] = g.v;
    }
    __syncthreads();
  }
  if (t == 0) { __threadfence(); sc.lwOk = 1; }
}
#endif

} // namespace hsrle
