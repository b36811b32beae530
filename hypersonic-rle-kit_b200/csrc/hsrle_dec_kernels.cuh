// hsrle_dec_kernels.cuh -- sm_100a kernels of the decoder (see hsrle_dec.cuh for the pipeline).
#pragma once
#include <cuda_runtime.h>
#include "hsrle_dec.cuh"
#include "hsrle_enc_kernels.cuh"   // shfl helpers, volatile access

namespace hsrle {

// ================================================================================================
// shared pieces of D1 and D3: SC image in shared memory, per-position exit table
struct DecScSmem
{
  alignas(16) uint8_t data[DEC_SCB + DEC_PAD];
  uint16_t ex[DEC_SCB];
};

// load stream bytes [c0, c0 + DEC_SCB + DEC_PAD) (zero beyond clen) -- 16-byte coalesced
__device__ __forceinline__ void dec_load_sc(uint8_t *data, const uint8_t *__restrict__ in, uint32_t c0, uint32_t clen)
{
  constexpr int NV = (DEC_SCB + DEC_PAD) / 16;
  const uint4 *src = reinterpret_cast<const uint4 *>(in + c0);
  uint4 *dst = reinterpret_cast<uint4 *>(data);
  const uint32_t avail = clen > c0 ? clen - c0 : 0;
  for (int v = threadIdx.x; v < NV; v += blockDim.x)
  {
    const uint32_t b = (uint32_t)v * 16;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (b < avail) x = __ldg(src + v);     // the 16-byte block holding byte clen-1 lies inside the caller's allocation
    if (b + 16 > avail)
    { // zero the bytes at and beyond clen so that nothing depends on them
      uint32_t w[4] = { x.x, x.y, x.z, x.w };
#pragma unroll
      for (int k = 0; k < 4; k++)
      {
        const uint32_t bb = b + 4 * k;
        if (bb >= avail) w[k] = 0;
        else if (bb + 4 > avail) w[k] &= (1u << (8 * (avail - bb))) - 1u;
      }
      x = make_uint4(w[0], w[1], w[2], w[3]);
    }
    dst[v] = x;
  }
}

// exit code of the token at SC-relative offset p (absolute c0 + p)
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t dec_hop_code(const uint8_t *data, uint32_t p, uint32_t c0, uint32_t clen, bool single, uint32_t &nxtRel)
{
  constexpr Spec sp = make_spec(W, BA, V);
  nxtRel = 0;
  const uint32_t pa = c0 + p;
  if (pa >= clen) return EX_BAD;
  Tok t; dec_parse(sp, single, data + p, (uint64_t)clen - pa, t);
  if (!t.valid) return EX_BAD;
  if (t.last) return EX_END;
  const uint64_t nr = (uint64_t)p + t.hdrLen + t.litLen;
  if (nr < EX_FAR) { nxtRel = (uint32_t)nr; return (uint32_t)nr; }
  return EX_FAR | p;
}

// per-position exit table of the SC: reverse sweep of one mini-block per thread
template <int W, int BA, int V>
__device__ __forceinline__ void dec_sweep(DecScSmem &S, uint32_t c0, uint32_t clen, bool single)
{
  const uint32_t b0 = threadIdx.x * DEC_MB, b1 = b0 + DEC_MB;
  for (uint32_t p = b1; p-- > b0;)
  {
    uint32_t nr;
    const uint32_t code = dec_hop_code<W, BA, V>(S.data, p, c0, clen, single, nr);
    S.ex[p] = (uint16_t)((code < EX_FAR && nr < b1) ? S.ex[nr] : code);
  }
}

// absolute exit position encoded by a table code (re-parses the far-jumping token)
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t dec_code_to_pos(const DecScSmem &S, uint32_t code, uint32_t c0, uint32_t clen, bool single)
{
  constexpr Spec sp = make_spec(W, BA, V);
  if (code < EX_FAR) return c0 + code;
  if (code == EX_END) return POS_END;
  if (code >= EX_END) return POS_BAD;
  const uint32_t p = code & 0x3FFFu;
  Tok t; dec_parse(sp, single, S.data + p, (uint64_t)clen - (c0 + p), t);
  return (uint32_t)((uint64_t)c0 + p + t.hdrLen + t.litLen);   // <= clen < POS_SPECIAL for a valid token
}

// ================================================================================================
// D1: windowed exit maps
template <int W, int BA, int V>
__global__ void __launch_bounds__(DEC_T) k_dec_map(const DecBufs D)
{
  constexpr Spec sp = make_spec(W, BA, V);
  extern __shared__ __align__(16) unsigned char smemRaw[];
  DecScSmem &S = *reinterpret_cast<DecScSmem *>(smemRaw);
  DecScalars hs; dec_header(sp, D.in, D.inSize, D.outSize, hs);
  if (blockIdx.x == 0 && threadIdx.x == 0) *D.sc = hs;
  if (hs.status != ST_OK) return;
  const uint32_t c = blockIdx.x;
  const uint32_t c0 = c * DEC_SCB;
  if (c0 >= hs.clen)
  {
    for (uint32_t w = threadIdx.x; w < DEC_WIN; w += DEC_T) D.map[(size_t)c * DEC_WIN + w] = POS_BAD;
    return;
  }
  const bool single = hs.single != 0;
  dec_load_sc(S.data, D.in, c0, hs.clen);
  __syncthreads();
  dec_sweep<W, BA, V>(S, c0, hs.clen, single);
  __syncthreads();
  // hop mini-block to mini-block from every window entry
  for (uint32_t w = threadIdx.x; w < DEC_WIN; w += DEC_T)
  {
    uint32_t code = S.ex[w];
    while (code < DEC_SCB) code = S.ex[code];
    D.map[(size_t)c * DEC_WIN + w] = dec_code_to_pos<W, BA, V>(S, code, c0, hs.clen, single);
  }
}

// ================================================================================================
// D2a: compose the maps of one segment
static __global__ void __launch_bounds__(DEC_WIN) k_dec_compose(const DecBufs D)
{
  extern __shared__ __align__(16) unsigned char smemRaw[];
  uint32_t *maps = reinterpret_cast<uint32_t *>(smemRaw);   // [DEC_SEG][DEC_WIN]
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t g = blockIdx.x;
  const uint32_t cFirst = g * DEC_SEG;
  const uint32_t nHere = min(DEC_SEG, D.nSC - cFirst);
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(D.map + (size_t)cFirst * DEC_WIN);
    uint4 *dst = reinterpret_cast<uint4 *>(maps);
    for (uint32_t v = threadIdx.x; v < nHere * DEC_WIN / 4; v += blockDim.x) dst[v] = src[v];
  }
  __syncthreads();
  const uint32_t w = threadIdx.x;
  uint32_t pos = cFirst * DEC_SCB + w;
  uint32_t *trail = D.trail + (size_t)g * DEC_SEG * DEC_WIN;
  for (uint32_t i = 0; i < nHere; i++)
  {
    const uint32_t c0 = (cFirst + i) * DEC_SCB;
    uint32_t tr = POS_NONE;
    if (pos < POS_SPECIAL && pos - c0 < DEC_SCB)
    {
      tr = pos;
      const uint32_t off = pos - c0;
      pos = off < DEC_WIN ? maps[i * DEC_WIN + off] : POS_MISS;
    }
    trail[i * DEC_WIN + w] = tr;
  }
  D.segExit[(size_t)g * DEC_WIN + w] = pos;
}

// ================================================================================================
// D2b: chain the segments, pick the true entries
constexpr int D2B_T = 1024;
constexpr uint32_t D2B_BATCH = 48;        // segment maps staged in shared memory at a time (96 KiB)

template <int W, int BA, int V>
__device__ uint32_t dec_walk_global(const DecBufs &D, const DecScalars &sc, uint32_t pos, uint32_t end)
{ // slow path: walk tokens in global memory from pos until the chain leaves [.., end)
  constexpr Spec sp = make_spec(W, BA, V);
  while (pos < end)
  {
    Tok t; dec_parse(sp, sc.single != 0, D.in + pos, (uint64_t)sc.clen - pos, t);
    if (!t.valid) return POS_BAD;
    if (t.last) return POS_END;
    pos = (uint32_t)((uint64_t)pos + t.hdrLen + t.litLen);
  }
  return pos;
}

template <int W, int BA, int V>
__global__ void __launch_bounds__(D2B_T) k_dec_resolve(const DecBufs D)
{
  extern __shared__ __align__(16) unsigned char smemRaw[];
  uint32_t *segMaps = reinterpret_cast<uint32_t *>(smemRaw);     // [D2B_BATCH][DEC_WIN]
  __shared__ uint32_t segW[D2B_BATCH];                           // window entry of the true chain per segment / POS_NONE / POS_MISS
  __shared__ uint32_t sPos;
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t clen = sc.clen;
  if (threadIdx.x == 0) sPos = sc.first;
  __syncthreads();
  for (uint32_t gb = 0; gb < D.nSeg; gb += D2B_BATCH)
  {
    const uint32_t nb = min(D2B_BATCH, D.nSeg - gb);
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(D.segExit + (size_t)gb * DEC_WIN);
      uint4 *dst = reinterpret_cast<uint4 *>(segMaps);
      for (uint32_t v = threadIdx.x; v < nb * DEC_WIN / 4; v += blockDim.x) dst[v] = src[v];
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
      uint32_t pos = sPos;
      for (uint32_t k = 0; k < nb; k++)
      {
        const uint32_t g = gb + k;
        const uint32_t s0 = g * DEC_SEG * DEC_SCB;
        const uint64_t s1 = (uint64_t)s0 + (uint64_t)DEC_SEG * DEC_SCB;
        if (pos >= POS_SPECIAL || pos >= s1) { segW[k] = POS_NONE; continue; }
        const uint32_t w = pos - s0;
        uint32_t e = w < DEC_WIN ? segMaps[k * DEC_WIN + w] : POS_MISS;
        if (e != POS_MISS) { segW[k] = w; pos = e; continue; }
        // slow path: SC by SC through this segment
        segW[k] = POS_MISS;
        for (uint32_t i = 0; i < DEC_SEG && g * DEC_SEG + i < D.nSC; i++)
        {
          const uint32_t c = g * DEC_SEG + i;
          const uint32_t c0 = c * DEC_SCB;
          const uint32_t c1 = (uint32_t)min((uint64_t)c0 + DEC_SCB, (uint64_t)clen);
          if (pos >= POS_SPECIAL || pos - c0 >= DEC_SCB) { D.scEntry[c] = POS_NONE; continue; }
          D.scEntry[c] = pos;
          const uint32_t off = pos - c0;
          pos = off < DEC_WIN ? D.map[(size_t)c * DEC_WIN + off] : dec_walk_global<W, BA, V>(D, sc, pos, c1);
        }
      }
      sPos = pos;
    }
    __syncthreads();
    for (uint32_t c = gb * DEC_SEG + threadIdx.x; c < min(D.nSC, (gb + nb) * DEC_SEG); c += blockDim.x)
    {
      const uint32_t k = c / DEC_SEG - gb, i = c % DEC_SEG;
      const uint32_t w = segW[k];
      if (w == POS_MISS) continue;
      D.scEntry[c] = (w == POS_NONE) ? POS_NONE : D.trail[((size_t)(gb + k) * DEC_SEG + i) * DEC_WIN + w];
    }
    __syncthreads();
  }
}

// ================================================================================================
// D3: token walk, look-back, expansion
template <int K> struct DecExpandSmem
{
  DecScSmem sc;                              // data + exit table; the exit table is reused for the token records
  uint32_t mbEntry[DEC_T];                   // SC-relative entry of the true chain into every mini-block (or ~0)
  DecAgg<K> warpAgg[DEC_T / 32];
  uint32_t ticket;
};
// token records of an expansion pass live where the exit table was:
//   tSym[DEC_TOKCAP] (u64) | tOut[DEC_TOKCAP+1] | tLitLen[DEC_TOKCAP] | tLitSrc[DEC_TOKCAP]
static_assert(DEC_TOKCAP * 8 + (DEC_TOKCAP + 4) * 4 + DEC_TOKCAP * 8 <= DEC_SCB * 2, "token records must fit the exit table");

template <int K> __device__ __forceinline__ DecAgg<K> dec_block_excl_scan(DecAgg<K> *warpBuf, const DecAgg<K> &mine, DecAgg<K> &total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  DecAgg<K> inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const DecAgg<K> o = shfl_up_t(inc, d);
    if (lane >= d) inc = decagg_combine<K>(o, inc);
  }
  if (lane == 31) warpBuf[warp] = inc;
  DecAgg<K> ex = shfl_up_t(inc, 1);
  if (lane == 0) ex = decagg_identity<K>();
  __syncthreads();
  DecAgg<K> pre = decagg_identity<K>();
  total = decagg_identity<K>();
#pragma unroll
  for (int w = 0; w < DEC_T / 32; w++)
  {
    const DecAgg<K> t = warpBuf[w];
    if (w < warp) pre = decagg_combine<K>(pre, t);
    total = decagg_combine<K>(total, t);
  }
  __syncthreads();
  return decagg_combine<K>(pre, ex);
}

// symbol of a token given the running symbol state; updates the state
template <int W, int BA, int V>
__device__ __forceinline__ uint64_t dec_token_symbol(const Tok &t, const uint8_t *tokPtr, uint64_t &symReg, Lut &lut)
{
  constexpr Spec sp = make_spec(W, BA, V);
  constexpr int K = sp.K;
  if (K)
  {
    const int idx = t.symKind == 0 ? K : t.symKind - 2;
    if (idx == K) lut_touch(lut, K, K, load_sym(tokPtr + t.symOff, W));
    else if (idx > 0) { const uint64_t v = lut.s[idx]; lut_touch(lut, K, idx, v); }
    return lut.s[0];
  }
  if (t.symKind == 0) symReg = load_sym(tokPtr + t.symOff, W);
  return symReg;
}

template <int W, int BA, int V>
__global__ void __launch_bounds__(DEC_T) k_dec_expand(const DecBufs D)
{
  constexpr Spec sp = make_spec(W, BA, V);
  constexpr int K = sp.K;
  using Agg = DecAgg<K>;
  using Smem = DecExpandSmem<K>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  Smem &S = *reinterpret_cast<Smem *>(smemRaw);
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  if (threadIdx.x == 0) S.ticket = atomicAdd(D.ticket, 1u);   // SCs in ticket order: predecessors are running or done
  __syncthreads();
  const uint32_t c = S.ticket;
  if (c >= D.nSC) return;
  const uint32_t c0 = c * DEC_SCB;
  const uint32_t clen = sc.clen, n = sc.n;
  const bool single = sc.single != 0;
  const int t = threadIdx.x;
  Agg *aggBuf = reinterpret_cast<Agg *>(D.aggBuf), *incBuf = reinterpret_cast<Agg *>(D.incBuf);
  const uint32_t entry = D.scEntry[c];
  const bool has = entry < POS_SPECIAL;

  // ---- token chain of this SC
  Agg mine = decagg_identity<K>();
  uint32_t myEntry = 0xFFFFFFFFu;
  bool sawEnd = false, sawBad = false;
  if (has)
  {
    dec_load_sc(S.sc.data, D.in, c0, clen);
    __syncthreads();
    dec_sweep<W, BA, V>(S.sc, c0, clen, single);
    S.mbEntry[t] = 0xFFFFFFFFu;
    __syncthreads();
    if (t == 0)
    { // entries of the true chain into the mini-blocks
      uint32_t p = entry - c0;
      while (p < DEC_SCB)
      {
        S.mbEntry[p / DEC_MB] = p;
        const uint32_t code = S.sc.ex[p];
        p = code < EX_FAR ? code : DEC_SCB;          // leaves the SC (or ends / breaks inside this mini-block)
      }
    }
    __syncthreads();
    myEntry = S.mbEntry[t];
    if (myEntry != 0xFFFFFFFFu)
    { // walk my mini-block: sizes and symbol summary
      uint32_t p = myEntry;
      const uint32_t b1 = (t + 1) * DEC_MB;
      while (p < b1)
      {
        Tok tk; dec_parse(sp, single, S.sc.data + p, (uint64_t)clen - (c0 + p), tk);
        if (c0 + p >= clen || !tk.valid) { sawBad = true; break; }
        mine.out += (uint64_t)tk.litLen + tk.runLen; mine.ntok++;
        if (K)
        {
          const int idx = tk.symKind == 0 ? K : tk.symKind - 2;
          lutxf_touch(mine.xf, K, idx, tk.symKind == 0 ? load_sym(S.sc.data + p + tk.symOff, W) : 0);
        }
        else if (tk.symKind == 0) { mine.has = 1; mine.sym = load_sym(S.sc.data + p + tk.symOff, W); }
        if (tk.last) { sawEnd = true; break; }
        const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
        p = nx < DEC_SCB ? (uint32_t)nx : DEC_SCB;
      }
    }
  }
  if (__syncthreads_or(sawBad ? 1 : 0)) { if (t == 0) D.sc->status = ST_BADSTREAM; }
  Agg total;
  const Agg pre = dec_block_excl_scan<K>(S.warpAgg, mine, total);

  // ---- look-back over SCs: thread 0 takes the inclusive prefix of the SC before the group, thread k the
  //      aggregate of SC g0+k-1
  if (t == 0)
  {
    aggBuf[c] = total;
    __threadfence();
    atomicExch(&D.aggFlag[c], 1u);
  }
  const uint32_t g0 = (c / DEC_GROUP) * DEC_GROUP;
  Agg part = decagg_identity<K>();
  if (t == 0)
  {
    if (g0 > 0) { while (atomicAdd(&D.incFlag[g0 - 1], 0u) == 0u) { } __threadfence(); part = incBuf[g0 - 1]; }
  }
  else
  {
    const uint32_t p = g0 + t - 1;
    if (p < c) { while (atomicAdd(&D.aggFlag[p], 0u) == 0u) { } __threadfence(); part = aggBuf[p]; }
  }
  Agg exclusive;
  (void)dec_block_excl_scan<K>(S.warpAgg, part, exclusive);
  if (t == 0)
  {
    incBuf[c] = decagg_combine<K>(exclusive, total);
    __threadfence();
    atomicExch(&D.incFlag[c], 1u);
  }
  // the SC that holds the final token reports the result
  if (__syncthreads_or(sawEnd ? 1 : 0))
  {
    if (t == 0)
    {
      const uint64_t outTotal = exclusive.out + total.out;
      DecScalars &w = *D.sc;
      w.endSeen = 1; w.nTok = exclusive.ntok + total.ntok;
      if (outTotal != n && w.status == ST_OK) w.status = ST_BADSTREAM;
    }
  }
  if (!has || total.ntok == 0) return;
  if (exclusive.out + total.out > (uint64_t)n) { if (t == 0) D.sc->status = ST_BADSTREAM; return; }   // corrupt stream: never expand past n

  // ---- state at the start of my mini-block
  const Agg before = decagg_combine<K>(exclusive, pre);
  uint64_t symReg = single ? (uint64_t)sc.singleSym : before.sym;     // register starts as zero (src/rleX_extreme_cpu_decode.h:33)
  Lut lut; lut_init(lut, W);
  if (K) { Lut l0 = lut; lutxf_apply(before.xf, K, l0, lut); }
  uint64_t outPos = before.out;
  const uint32_t tokFirst = pre.ntok;                                  // index of my first token inside the SC
  const uint64_t scOut0 = exclusive.out, scOut1 = exclusive.out + total.out;
  // the exit table is dead now: token records of a pass
  uint64_t *tSym = reinterpret_cast<uint64_t *>(S.sc.ex);
  uint32_t *tOut = reinterpret_cast<uint32_t *>(tSym + DEC_TOKCAP);
  uint32_t *tLitLen = tOut + DEC_TOKCAP + 4;
  uint32_t *tLitSrc = tLitLen + DEC_TOKCAP;

  // ---- expansion in passes of DEC_TOKCAP tokens
  uint32_t p = myEntry;
  uint32_t k = tokFirst;                                               // my next token index
  const uint32_t b1 = (t + 1) * DEC_MB;
  for (uint32_t pass0 = 0; pass0 < total.ntok; pass0 += DEC_TOKCAP)
  {
    const uint32_t passN = min(DEC_TOKCAP, total.ntok - pass0);
    __syncthreads();
    if (myEntry != 0xFFFFFFFFu)
    {
      while (p < b1 && k < pass0 + passN)
      {
        Tok tk; dec_parse(sp, single, S.sc.data + p, (uint64_t)clen - (c0 + p), tk);
        if (!tk.valid) break;
        const uint64_t sym = dec_token_symbol<W, BA, V>(tk, S.sc.data + p, symReg, lut);
        const uint32_t r = k - pass0;
        tOut[r] = (uint32_t)outPos; tLitLen[r] = tk.litLen; tLitSrc[r] = c0 + p + tk.hdrLen; tSym[r] = sym;
        outPos += (uint64_t)tk.litLen + tk.runLen; k++;
        if (tk.last) { p = DEC_SCB; break; }
        const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
        p = nx < DEC_SCB ? (uint32_t)nx : DEC_SCB;
      }
    }
    __syncthreads();
    // output range of this pass
    const uint64_t o0 = tOut[0];
    // the end of the pass: start of the first token of the next pass, or the end of the SC's output
    // (computed by the thread that owns token pass0+passN, if any)
    if (myEntry != 0xFFFFFFFFu && k == pass0 + passN && p < b1 && pass0 + passN < total.ntok) tOut[passN] = (uint32_t)outPos;
    if (pass0 + passN >= total.ntok && t == 0) tOut[passN] = (uint32_t)min(scOut1, (uint64_t)0xFFFFFFFFu);
    __syncthreads();
    const uint64_t o1raw = tOut[passN];
    const uint64_t oEnd = min(o1raw, (uint64_t)n);                      // never write beyond the declared size
    if (o0 >= oEnd) continue;
    const uint64_t v0 = o0 >> 4, v1 = (oEnd + 15) >> 4;
    for (uint64_t v = v0 + t; v < v1; v += DEC_T)
    {
      const uint64_t vb = v << 4;
      const uint64_t lo = max(vb, o0), hi = min(vb + 16, oEnd);
      // token covering lo: largest r with tOut[r] <= lo
      uint32_t a = 0, b = passN - 1;
      while (a < b) { const uint32_t m = (a + b + 1) >> 1; if (tOut[m] <= lo) a = m; else b = m - 1; }
      uint32_t r = a;
      uint64_t tStart = tOut[r], tNext = tOut[r + 1];
      uint32_t litLen = tLitLen[r];
      uint32_t w4[4] = { 0, 0, 0, 0 };
      uint64_t pos = lo;
      // fast path: the whole vector lies inside one run
      if (lo == vb && hi == vb + 16 && pos >= tStart + litLen && vb + 16 <= tNext)
      {
        const uint64_t sym = tSym[r];
        const uint32_t ph = (uint32_t)(pos - (tStart + litLen)) % (uint32_t)W;
#pragma unroll
        for (int j = 0; j < 4; j++) w4[j] = pattern_word(sym, W, (ph + 4 * j) % W);
      }
      else
      {
#pragma unroll
        for (int i = 0; i < 16; i++)
        {
          const uint64_t q = vb + i;
          if (q >= lo && q < hi)
          {
            while (q >= tNext) { r++; tStart = tNext; tNext = tOut[r + 1]; litLen = tLitLen[r]; }
            const uint64_t litEnd = tStart + litLen;
            uint32_t byte;
            if (q < litEnd)
            {
              const uint32_t sp_ = tLitSrc[r] + (uint32_t)(q - tStart);
              byte = (sp_ - c0 < DEC_SCB + DEC_PAD) ? S.sc.data[sp_ - c0] : __ldg(D.in + sp_);
            }
            else
            {
              const uint32_t ph = (uint32_t)(q - litEnd) % (uint32_t)W;
              byte = (uint32_t)(tSym[r] >> (8 * ph)) & 0xFFu;
            }
            w4[i >> 2] |= byte << (8 * (i & 3));
          }
        }
      }
      if (lo == vb && hi == vb + 16) *reinterpret_cast<uint4 *>(D.out + vb) = make_uint4(w4[0], w4[1], w4[2], w4[3]);
      else for (uint64_t q = lo; q < hi; q++) D.out[q] = (uint8_t)(w4[(q - vb) >> 2] >> (8 * ((q - vb) & 3)));
    }
  }
  (void)scOut0;
}

// final status (one thread): runs after D3
static __global__ void k_dec_finish(const DecBufs D)
{
  const DecScalars &sc = *D.sc;
  uint32_t status = sc.status;
  if (status == ST_OK && !sc.endSeen) status = ST_BADSTREAM;
  D.dResult[0] = status == ST_OK ? sc.n : 0; D.dResult[1] = status; D.dResult[2] = sc.nTok; D.dResult[3] = D.nSC;
  D.dResult[4] = sc.clen; D.dResult[5] = sc.single; D.dResult[6] = 0; D.dResult[7] = 0;
}

} // namespace hsrle
