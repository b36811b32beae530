// hsrle_dec_kernels.cuh -- sm_100a kernels of the decoder (see hsrle_dec.cuh for the pipeline).
#pragma once
#include <cuda_runtime.h>
#include "hsrle_dec.cuh"
#include "hsrle_enc_kernels.cuh"   // shfl helpers, volatile access

namespace hsrle {

// ================================================================================================
// shared-memory image of an SC.  One thread owns one 128-byte mini-block, so byte and table addresses are
// skewed by one word per mini-block: lane t then starts at bank t instead of bank 0 (no 32-way conflicts).
__device__ __forceinline__ uint32_t skew8(uint32_t x) { return x + ((x >> 7) << 2); }     // byte index
__device__ __forceinline__ uint32_t skew16(uint32_t x) { return x + ((x >> 7) << 1); }    // u16 index
constexpr uint32_t DEC_DATA_BYTES = DEC_SCB + DEC_PAD + ((DEC_SCB + DEC_PAD) / 128 + 1) * 4;
constexpr uint32_t DEC_EX_ELEMS = DEC_SCB + (DEC_SCB / 128 + 1) * 2;

struct SkewReader
{
  const uint8_t *data; uint32_t p;
  __device__ __forceinline__ uint32_t u8(uint32_t o) const { return data[skew8(p + o)]; }
};

// load stream bytes [c0, c0 + DEC_SCB + DEC_PAD) (zero beyond clen) into the skewed image -- 16-byte coalesced
__device__ __forceinline__ void dec_load_sc(uint8_t *data, const uint8_t *__restrict__ in, uint32_t c0, uint32_t clen)
{
  constexpr int NV = (DEC_SCB + DEC_PAD) / 16;
  const uint4 *src = reinterpret_cast<const uint4 *>(in + c0);
  const uint32_t avail = clen > c0 ? clen - c0 : 0;
  for (int v = threadIdx.x; v < NV; v += blockDim.x)
  {
    const uint32_t b = (uint32_t)v * 16;
    uint4 x = make_uint4(0, 0, 0, 0);
    if (b < avail) x = __ldg(src + v);     // the 16-byte block holding byte clen-1 lies inside the caller's allocation
    uint32_t w[4] = { x.x, x.y, x.z, x.w };
    if (b + 16 > avail)
    { // zero the bytes at and beyond clen so that nothing depends on them
#pragma unroll
      for (int k = 0; k < 4; k++)
      {
        const uint32_t bb = b + 4 * k;
        if (bb >= avail) w[k] = 0;
        else if (bb + 4 > avail) w[k] &= (1u << (8 * (avail - bb))) - 1u;
      }
    }
    uint32_t *dst = reinterpret_cast<uint32_t *>(data + skew8(b));
    dst[0] = w[0]; dst[1] = w[1]; dst[2] = w[2]; dst[3] = w[3];
  }
}

// ------------------------------------------------------------------------------------------------
// length-only token parse from a 24-byte register window (bytes p .. p+23 of the stream).  Same decisions
// as dec_parse (hsrle_core.cuh), restricted to what the chain needs: kind and the distance to the next token.
struct TokWin { uint32_t w[6]; };
template <int K> __device__ __forceinline__ uint32_t win_u8(const TokWin &x) { return (x.w[K >> 2] >> (8 * (K & 3))) & 0xFFu; }
template <int K> __device__ __forceinline__ uint32_t win_u32(const TokWin &x)
{
  if constexpr ((K & 3) == 0) return x.w[K >> 2];
  else return __funnelshift_r(x.w[K >> 2], x.w[(K >> 2) + 1], 8 * (K & 3));
}
enum : uint32_t { TK_OK = 0, TK_END = 1, TK_BAD = 2 };

// [S symbol bytes][cnt][rng] with 8-bit fields and 0-escapes (plain tokens; S = 0 for single-symbol streams)
template <int S> __device__ __forceinline__ uint64_t toklen_plain(const TokWin &x, uint64_t avail, uint32_t &kind)
{
  const uint32_t c = win_u8<S>(x);
  const bool e1 = c == 0;
  const uint32_t cnt32 = win_u32<S + 1>(x);
  const uint32_t r = e1 ? win_u8<S + 5>(x) : win_u8<S + 1>(x);
  const uint32_t r32 = e1 ? win_u32<S + 6>(x) : win_u32<S + 2>(x);
  const bool e2 = r == 0;
  const uint32_t hdr = S + 2 + (e1 ? 4u : 0u) + (e2 ? 4u : 0u);
  const uint32_t rng = e2 ? r32 : r;
  const uint64_t len = (uint64_t)hdr + rng - 1;
  kind = (hdr > avail) ? TK_BAD : (rng == 0) ? TK_END : (len > avail) ? TK_BAD : (e1 && cnt32 == 0) ? TK_END : TK_OK;
  return len;
}
// packed tokens: b0 = same<<7 | cnt7, optional u32 cnt, optional symbol, rng in the 7-bit or the 8-bit style
template <int W, bool RNG7> __device__ __forceinline__ uint64_t toklen_packed(const TokWin &x, uint64_t avail, uint32_t &kind)
{
  const uint32_t b0 = win_u8<0>(x);
  const bool e1 = (b0 & 0x7F) == 0, same = (b0 & 0x80) != 0;
  const uint32_t cnt32 = win_u32<1>(x);
  const uint32_t o = 1 + (e1 ? 4u : 0u) + (same ? 0u : (uint32_t)W);
  const uint32_t r = same ? (e1 ? win_u8<5>(x) : win_u8<1>(x)) : (e1 ? win_u8<5 + W>(x) : win_u8<1 + W>(x));
  uint32_t hdr, rng; bool endMark;
  if (RNG7)
  {
    const uint32_t r32 = same ? (e1 ? win_u32<5>(x) : win_u32<1>(x)) : (e1 ? win_u32<5 + W>(x) : win_u32<1 + W>(x));
    const bool esc = (r & 1) != 0;
    rng = esc ? (r32 >> 1) : (r >> 1);
    hdr = o + (esc ? 4u : 1u);
    endMark = esc && rng == 0;
  }
  else
  {
    const uint32_t r32 = same ? (e1 ? win_u32<6>(x) : win_u32<2>(x)) : (e1 ? win_u32<6 + W>(x) : win_u32<2 + W>(x));
    const bool esc = r == 0;
    rng = esc ? r32 : r;
    hdr = o + (esc ? 5u : 1u);
    endMark = esc && rng == 0;
  }
  const uint64_t len = (uint64_t)hdr + rng - 1;
  kind = (hdr > avail) ? TK_BAD : endMark ? TK_END : (rng == 0 || len > avail) ? TK_BAD : (e1 && cnt32 == 0) ? TK_END : TK_OK;
  return len;
}
// LUT tokens: u16 head = idx | cnt7 | rng, optional symbol, optional u16/u32 cnt, optional u16/u32 rng
template <int W, int K> __device__ __forceinline__ uint64_t toklen_lut(const TokWin &x, uint64_t avail, uint32_t &kind)
{
  constexpr int RB = (K == 3) ? 7 : 6;
  const uint32_t head = win_u32<0>(x) & 0xFFFFu;
  const bool miss = (head >> (K == 3 ? 14 : 13)) == (uint32_t)K;
  const uint32_t c7 = (head >> RB) & 0x7F, r = head & ((1u << RB) - 1u);
  const uint32_t ce = c7 == 1 ? 2u : (c7 == 0 ? 4u : 0u);
  const uint32_t x1 = miss ? win_u32<2 + W>(x) : win_u32<2>(x);
  const uint32_t cnt = c7 == 1 ? (x1 & 0xFFFFu) : (c7 == 0 ? x1 : c7);
  const uint32_t x2 = miss ? (ce == 0 ? win_u32<2 + W>(x) : ce == 2 ? win_u32<4 + W>(x) : win_u32<6 + W>(x))
                           : (ce == 0 ? win_u32<2>(x) : ce == 2 ? win_u32<4>(x) : win_u32<6>(x));
  const uint32_t re = r == 1 ? 2u : (r == 0 ? 4u : 0u);
  const uint32_t rng = r == 1 ? (x2 & 0xFFFFu) : (r == 0 ? x2 : r);
  const uint32_t hdr = 2 + (miss ? (uint32_t)W : 0u) + ce + re;
  const bool endMark = r == 1 && rng == 0;
  const uint64_t len = (uint64_t)hdr + rng - 2;
  kind = (hdr > avail) ? TK_BAD : endMark ? TK_END : (rng < 2 || len > avail) ? TK_BAD : (cnt == 0) ? TK_END : TK_OK;
  return len;
}
template <int W, int BA, int V>
__device__ __forceinline__ uint64_t toklen(const TokWin &x, bool single, uint64_t avail, uint32_t &kind)
{
  constexpr Spec sp = make_spec(W, BA, V);
  if constexpr (sp.K != 0) return toklen_lut<W, sp.K>(x, avail, kind);
  else
  {
    if constexpr (W == 1) { if (single) return toklen_plain<0>(x, avail, kind); }
    if constexpr (V == V_PLAIN) return toklen_plain<W>(x, avail, kind);
    else return toklen_packed<W, sp.rng7 != 0>(x, avail, kind);
  }
}

// per-position exit table of the SC: reverse sweep of one mini-block per thread with a sliding window.
// ex[q] (u16, SC-relative): < EX_FAR: where the chain that starts at q leaves q's mini-block; EX_FAR | q': the
// token at q' jumps beyond c0 + 0x7FFF; EX_END / EX_BAD.
template <int W, int BA, int V>
__device__ __forceinline__ void dec_sweep(const uint8_t *data, uint16_t *ex, uint32_t c0, uint32_t clen, bool single)
{
  const uint32_t b0 = threadIdx.x * DEC_MB, b1 = b0 + DEC_MB;
  TokWin x;
  { // window at p = b1 - 1
    uint32_t p = b1 - 1;
#pragma unroll
    for (int k = 0; k < 6; k++)
    {
      uint32_t v = 0;
#pragma unroll
      for (int j = 0; j < 4; j++) v |= (uint32_t)data[skew8(p + 4 * k + j)] << (8 * j);
      x.w[k] = v;
    }
  }
  for (uint32_t p = b1; p-- > b0;)
  {
    const uint32_t pa = c0 + p;
    uint32_t kind;
    const uint64_t len = toklen<W, BA, V>(x, single, pa < clen ? (uint64_t)(clen - pa) : 0ull, kind);
    const uint64_t nr = (uint64_t)p + len;
    uint32_t code;
    if (kind != TK_OK) code = kind == TK_END ? EX_END : EX_BAD;
    else if (nr < b1) code = ex[skew16((uint32_t)nr)];
    else code = nr < EX_FAR ? (uint32_t)nr : (EX_FAR | p);
    ex[skew16(p)] = (uint16_t)code;
    // slide the window down by one byte
    if (p > b0)
    {
      const uint32_t nb = data[skew8(p - 1)];
#pragma unroll
      for (int k = 5; k > 0; k--) x.w[k] = __funnelshift_l(x.w[k - 1], x.w[k], 8);
      x.w[0] = (x.w[0] << 8) | nb;
    }
  }
}

// absolute exit position encoded by a table code (re-parses the far-jumping token)
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t dec_code_to_pos(const uint8_t *data, uint32_t code, uint32_t c0, uint32_t clen, bool single)
{
  constexpr Spec sp = make_spec(W, BA, V);
  if (code < EX_FAR) return c0 + code;
  if (code == EX_END) return POS_END;
  if (code >= EX_END) return POS_BAD;
  const uint32_t p = code & 0x3FFFu;
  SkewReader rd; rd.data = data; rd.p = p;
  Tok t; dec_parse_rd(sp, single, rd, (uint64_t)clen - (c0 + p), t);
  return (uint32_t)((uint64_t)c0 + p + t.hdrLen + t.litLen);   // <= clen < POS_SPECIAL for a valid token
}

// ================================================================================================
// D1: per-position exit tables: mini-block level (exTab, for D3) and SC level (finTab, for D2)
struct DecMapSmem
{
  alignas(16) uint8_t data[DEC_DATA_BYTES];
  alignas(16) uint16_t ex[DEC_EX_ELEMS];
};
constexpr uint32_t DEC_WB = 32 * DEC_MB;    // warp-block: the 32 mini-blocks swept by one warp

template <int W, int BA, int V>
__global__ void __launch_bounds__(DEC_T) k_dec_map(const DecBufs D)
{
  constexpr Spec sp = make_spec(W, BA, V);
  extern __shared__ __align__(16) unsigned char smemRaw[];
  DecMapSmem &S = *reinterpret_cast<DecMapSmem *>(smemRaw);
  DecScalars hs; dec_header(sp, D.in, D.inSize, D.outSize, hs);
  if (blockIdx.x == 0 && threadIdx.x == 0) *D.sc = hs;
  if (hs.status != ST_OK) return;
  const uint32_t c = blockIdx.x;
  const uint32_t c0 = c * DEC_SCB;
  if (c0 >= hs.clen) return;
  const bool single = hs.single != 0;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  dec_load_sc(S.data, D.in, c0, hs.clen);
  __syncthreads();
  dec_sweep<W, BA, V>(S.data, S.ex, c0, hs.clen, single);
  __syncthreads();
  // keep the mini-block table for D3 (two entries per 4-byte store)
  {
    uint32_t *dst = reinterpret_cast<uint32_t *>(D.exTab + (size_t)c * DEC_SCB);
    for (uint32_t q = t * 2; q < DEC_SCB; q += DEC_T * 2) dst[q >> 1] = *reinterpret_cast<const uint32_t *>(S.ex + skew16(q));
  }
  __syncthreads();
  // finalise in place, level 1: inside every warp-block, mini-blocks in reverse order (a code below the end of
  // the warp-block points into a later mini-block of the same warp-block, which is final already)
  {
    const uint32_t wb0 = warp * DEC_WB, wb1 = wb0 + DEC_WB;
    for (int mb = 30; mb >= 0; mb--)
    {
      uint32_t code[4];
#pragma unroll
      for (int k = 0; k < 4; k++) code[k] = S.ex[skew16(wb0 + mb * DEC_MB + lane + 32 * k)];
#pragma unroll
      for (int k = 0; k < 4; k++) if (code[k] < wb1) code[k] = S.ex[skew16(code[k])];
#pragma unroll
      for (int k = 0; k < 4; k++) S.ex[skew16(wb0 + mb * DEC_MB + lane + 32 * k)] = (uint16_t)code[k];
      __syncwarp();
    }
  }
  __syncthreads();
  // level 2: warp-blocks in reverse order
  for (int wb = DEC_T / 32 - 2; wb >= 0; wb--)
  {
    for (uint32_t p = wb * DEC_WB + t; p < (wb + 1) * DEC_WB; p += DEC_T)
    {
      uint32_t code = S.ex[skew16(p)];
      if (code < DEC_SCB) { code = S.ex[skew16(code)]; S.ex[skew16(p)] = (uint16_t)code; }
    }
    __syncthreads();
  }
  // absolute SC exits of every position
  uint32_t *fin = D.finTab + (size_t)c * DEC_SCB;
  for (uint32_t p = t; p < DEC_SCB; p += DEC_T) fin[p] = dec_code_to_pos<W, BA, V>(S.data, S.ex[skew16(p)], c0, hs.clen, single);
}

// ================================================================================================
// D2a: per segment, where does the chain that enters SC c at window offset w leave the segment
static __global__ void __launch_bounds__(DEC_WIN) k_dec_compose(const DecBufs D)
{
  extern __shared__ __align__(16) unsigned char smemRaw[];
  uint32_t *suf = reinterpret_cast<uint32_t *>(smemRaw);   // [DEC_SEG][DEC_WIN]
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t nSC = (sc.clen + DEC_SCB - 1) / DEC_SCB;
  const uint32_t g = blockIdx.x;
  const uint32_t cFirst = g * DEC_SEG;
  if (cFirst >= nSC) return;
  const uint32_t nHere = min(DEC_SEG, nSC - cFirst);
  const uint64_t segEnd = (uint64_t)(cFirst + nHere) * DEC_SCB;
  const uint32_t *__restrict__ fin = D.finTab;
  const uint32_t w = threadIdx.x;
  for (int i = (int)nHere - 1; i >= 0; i--)
  {
    uint32_t x = fin[(size_t)(cFirst + i) * DEC_SCB + w];
    while (x < POS_SPECIAL && (uint64_t)x < segEnd)
    {
      const uint32_t c2 = x / DEC_SCB, off = x - c2 * DEC_SCB;
      if (off < DEC_WIN) { x = suf[(c2 - cFirst) * DEC_WIN + off]; break; }   // a later SC of the segment: final already
      x = fin[x];                                                             // entry beyond the window: one SC at a time
    }
    suf[i * DEC_WIN + w] = x;
    __syncthreads();
  }
  uint32_t *dst = D.sufExit + (size_t)cFirst * DEC_WIN;
  for (uint32_t i = 0; i < nHere; i++) dst[i * DEC_WIN + w] = suf[i * DEC_WIN + w];
}

// ================================================================================================
// D2b: chain the segments, record the true entry of every SC
constexpr int D2B_T = 1024;

static __global__ void __launch_bounds__(D2B_T) k_dec_resolve(const DecBufs D)
{
  DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t clen = sc.clen;
  const uint32_t nSC = (clen + DEC_SCB - 1) / DEC_SCB;
  const uint32_t nSeg = (nSC + DEC_SEG - 1) / DEC_SEG;
  const uint32_t *__restrict__ fin = D.finTab;
  for (uint32_t c = threadIdx.x; c < D.nSC; c += blockDim.x) D.scEntry[c] = POS_NONE;
  for (uint32_t g = threadIdx.x; g < nSeg; g += blockDim.x) D.segEntry[g] = POS_NONE;
  __syncthreads();
  if (threadIdx.x == 0)
  {
    uint32_t pos = sc.first, lastSeg = 0xFFFFFFFFu;
    while (pos < POS_SPECIAL)
    {
      if (pos >= clen) { pos = POS_BAD; break; }
      const uint32_t c = pos / DEC_SCB, g = c / DEC_SEG, off = pos - c * DEC_SCB;
      if (g != lastSeg) { D.segEntry[g] = pos; lastSeg = g; }
      pos = off < DEC_WIN ? D.sufExit[(size_t)c * DEC_WIN + off] : fin[pos];
    }
    if (pos != POS_END) sc.status = ST_BADSTREAM;
  }
  __syncthreads();
  for (uint32_t g = threadIdx.x; g < nSeg; g += blockDim.x)
  {
    uint32_t pos = D.segEntry[g];
    const uint64_t segEnd = (uint64_t)(g + 1) * DEC_SEG * DEC_SCB;
    while (pos < POS_SPECIAL && (uint64_t)pos < segEnd && pos < clen)
    {
      D.scEntry[pos / DEC_SCB] = pos;
      pos = fin[pos];
    }
  }
}

// ================================================================================================
// D3: token walk (D3a), scan over SCs (D3s), expansion (D3b)
template <int K> struct DecWalkSmem
{
  alignas(16) uint8_t data[DEC_DATA_BYTES];  // skewed SC image
  alignas(16) uint16_t ex[DEC_SCB];          // exit table from D1 (linear)
  uint32_t mbEntry[DEC_T];                   // SC-relative entry of the true chain into every mini-block (or 0xFFFF)
  DecAgg<K> warpAgg[DEC_T / 32];
};
template <int K> struct DecExpandSmem
{
  alignas(16) uint8_t data[DEC_DATA_BYTES];  // skewed SC image
  alignas(16) uint64_t tSym[DEC_TOKCAP];     // token records of an expansion pass
  uint32_t tOut[DEC_TOKCAP + 4];
  uint32_t tLitLen[DEC_TOKCAP];
  uint32_t tLitSrc[DEC_TOKCAP];
  DecAgg<K> warpAgg[DEC_T / 32];
};

template <int K> __device__ __forceinline__ DecAgg<K> dec_block_excl_scan(DecAgg<K> *warpBuf, const DecAgg<K> &mine, DecAgg<K> &total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
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
  for (int w = 0; w < nw; w++)
  {
    const DecAgg<K> t = warpBuf[w];
    if (w < warp) pre = decagg_combine<K>(pre, t);
    total = decagg_combine<K>(total, t);
  }
  __syncthreads();
  return decagg_combine<K>(pre, ex);
}

// sizes and symbol summary of the tokens of my mini-block (walk #1)
template <int W, int BA, int V>
__device__ __forceinline__ void dec_walk_sizes(const uint8_t *data, uint32_t myEntry, uint32_t c0, uint32_t clen, bool single,
                                               DecAgg<make_spec(W, BA, V).K> &mine, bool &sawEnd, bool &sawBad)
{
  constexpr Spec sp = make_spec(W, BA, V);
  constexpr int K = sp.K;
  uint32_t p = myEntry;
  const uint32_t b1 = (p / DEC_MB + 1) * DEC_MB;
  while (p < b1)
  {
    SkewReader rd; rd.data = data; rd.p = p;
    Tok tk; dec_parse_rd(sp, single, rd, (uint64_t)clen - (c0 + p), tk);
    if (c0 + p >= clen || !tk.valid) { sawBad = true; break; }
    mine.out += (uint64_t)tk.litLen + tk.runLen; mine.ntok++;
    if (K)
    {
      const int idx = tk.symKind == 0 ? K : tk.symKind - 2;
      lutxf_touch(mine.xf, K, idx, tk.symKind == 0 ? rd_sym(rd, tk.symOff, W) : 0);
    }
    else if (tk.symKind == 0) { mine.has = 1; mine.sym = rd_sym(rd, tk.symOff, W); }
    if (tk.last) { sawEnd = true; break; }
    const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
    p = nx < DEC_SCB ? (uint32_t)nx : DEC_SCB;
  }
}

// D3a: per SC, mark the true chain (entries into the mini-blocks) and total its tokens
template <int W, int BA, int V>
__global__ void __launch_bounds__(DEC_T) k_dec_walk(const DecBufs D)
{
  constexpr Spec sp = make_spec(W, BA, V);
  constexpr int K = sp.K;
  using Agg = DecAgg<K>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  DecWalkSmem<K> &S = *reinterpret_cast<DecWalkSmem<K> *>(smemRaw);
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t c = blockIdx.x, c0 = c * DEC_SCB, clen = sc.clen;
  const bool single = sc.single != 0;
  const int t = threadIdx.x;
  Agg *aggBuf = reinterpret_cast<Agg *>(D.aggBuf);
  const uint32_t entry = D.scEntry[c];
  if (entry >= POS_SPECIAL)
  { // no token starts here
    if (t == 0) aggBuf[c] = decagg_identity<K>();
    D.mbEntry[(size_t)c * DEC_T + t] = 0xFFFFu;
    return;
  }
  dec_load_sc(S.data, D.in, c0, clen);
  {
    const uint4 *src = reinterpret_cast<const uint4 *>(D.exTab + (size_t)c * DEC_SCB);
    uint4 *dst = reinterpret_cast<uint4 *>(S.ex);
    for (uint32_t v = t; v < DEC_SCB * 2 / 16; v += DEC_T) dst[v] = __ldg(src + v);
  }
  S.mbEntry[t] = 0xFFFFu;
  __syncthreads();
  if (t == 0)
  {
    uint32_t p = entry - c0;
    while (p < DEC_SCB)
    {
      S.mbEntry[p / DEC_MB] = p;
      const uint32_t code = S.ex[p];
      p = code < EX_FAR ? code : DEC_SCB;          // leaves the SC (or ends / breaks inside this mini-block)
    }
  }
  __syncthreads();
  const uint32_t myEntry = S.mbEntry[t];
  D.mbEntry[(size_t)c * DEC_T + t] = (uint16_t)myEntry;
  Agg mine = decagg_identity<K>();
  bool sawEnd = false, sawBad = false;
  if (myEntry != 0xFFFFu) dec_walk_sizes<W, BA, V>(S.data, myEntry, c0, clen, single, mine, sawEnd, sawBad);
  if (__syncthreads_or(sawBad ? 1 : 0)) { if (t == 0) D.sc->status = ST_BADSTREAM; }
  if (__syncthreads_or(sawEnd ? 1 : 0)) { if (t == 0) D.sc->endSeen = 1; }
  Agg total;
  (void)dec_block_excl_scan<K>(S.warpAgg, mine, total);
  if (t == 0) aggBuf[c] = total;
}

// D3s: exclusive scan of the SC aggregates (one CTA); final status and result
constexpr int D3S_T = 1024;
template <int K>
__global__ void __launch_bounds__(D3S_T) k_dec_scan(const DecBufs D)
{
  using Agg = DecAgg<K>;
  __shared__ Agg warpAgg[D3S_T / 32];
  DecScalars &sc = *D.sc;
  const int t = threadIdx.x;
  Agg *aggBuf = reinterpret_cast<Agg *>(D.aggBuf), *incBuf = reinterpret_cast<Agg *>(D.incBuf);
  if (sc.status == ST_OK)
  {
    const uint32_t nSC = D.nSC;
    const uint32_t per = (nSC + D3S_T - 1) / D3S_T;
    const uint32_t lo = min(nSC, (uint32_t)t * per), hi = min(nSC, lo + per);
    Agg mine = decagg_identity<K>();
    for (uint32_t c = lo; c < hi; c++) mine = decagg_combine<K>(mine, aggBuf[c]);
    Agg total;
    Agg run = dec_block_excl_scan<K>(warpAgg, mine, total);
    for (uint32_t c = lo; c < hi; c++) { const Agg a = aggBuf[c]; incBuf[c] = run; run = decagg_combine<K>(run, a); }   // incBuf = EXCLUSIVE prefix
    if (t == 0)
    {
      sc.nTok = total.ntok;
      if (!sc.endSeen || total.out != (uint64_t)sc.n) sc.status = ST_BADSTREAM;
    }
  }
  __syncthreads();
  if (t == 0)
  {
    const uint32_t status = sc.status;
    D.dResult[0] = status == ST_OK ? sc.n : 0; D.dResult[1] = status; D.dResult[2] = sc.nTok; D.dResult[3] = D.nSC;
    D.dResult[4] = sc.clen; D.dResult[5] = sc.single; D.dResult[6] = 0; D.dResult[7] = 0;
  }
}

// symbol of a token given the running symbol state; updates the state
template <int W, int BA, int V>
__device__ __forceinline__ uint64_t dec_token_symbol(const Tok &t, const SkewReader &rd, uint64_t &symReg, Lut &lut)
{
  constexpr Spec sp = make_spec(W, BA, V);
  constexpr int K = sp.K;
  if (K)
  {
    const int idx = t.symKind == 0 ? K : t.symKind - 2;
    if (idx == K) lut_touch(lut, K, K, rd_sym(rd, t.symOff, W));
    else if (idx > 0) { const uint64_t v = lut_get(lut, K, idx); lut_touch(lut, K, idx, v); }
    return lut.s[0];
  }
  if (t.symKind == 0) symReg = rd_sym(rd, t.symOff, W);
  return symReg;
}

// D3b: expansion.  Per SC: token records (output offset, literal source, symbol) in passes of DEC_TOKCAP
// tokens, then one 16-byte aligned output vector per thread and step.
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
  const uint32_t c = blockIdx.x, c0 = c * DEC_SCB;
  const uint32_t clen = sc.clen, n = sc.n;
  const bool single = sc.single != 0;
  const int t = threadIdx.x;
  const Agg *aggBuf = reinterpret_cast<const Agg *>(D.aggBuf), *exclBuf = reinterpret_cast<const Agg *>(D.incBuf);
  if (D.scEntry[c] >= POS_SPECIAL) return;
  const Agg total = aggBuf[c];
  if (total.ntok == 0) return;
  const Agg exclusive = exclBuf[c];
  dec_load_sc(S.data, D.in, c0, clen);
  const uint32_t myEntry = D.mbEntry[(size_t)c * DEC_T + t];
  __syncthreads();
  Agg mine = decagg_identity<K>();
  { bool e = false, b = false; if (myEntry != 0xFFFFu) dec_walk_sizes<W, BA, V>(S.data, myEntry, c0, clen, single, mine, e, b); }
  Agg totalChk;
  const Agg pre = dec_block_excl_scan<K>(S.warpAgg, mine, totalChk);

  // ---- state at the start of my mini-block
  const Agg before = decagg_combine<K>(exclusive, pre);
  uint64_t symReg = single ? (uint64_t)sc.singleSym : before.sym;     // register starts as zero (src/rleX_extreme_cpu_decode.h:33)
  Lut lut; lut_init(lut, W);
  if (K) { Lut l0 = lut; lutxf_apply(before.xf, K, l0, lut); }
  uint64_t outPos = before.out;
  const uint64_t scOut1 = exclusive.out + total.out;
  uint64_t *tSym = S.tSym;
  uint32_t *tOut = S.tOut, *tLitLen = S.tLitLen, *tLitSrc = S.tLitSrc;

  // ---- expansion in passes of DEC_TOKCAP tokens
  uint32_t p = myEntry == 0xFFFFu ? DEC_SCB : myEntry;
  uint32_t k = pre.ntok;                                               // my next token index inside the SC
  const uint32_t b1 = (t + 1) * DEC_MB;
  for (uint32_t pass0 = 0; pass0 < total.ntok; pass0 += DEC_TOKCAP)
  {
    const uint32_t passN = min(DEC_TOKCAP, total.ntok - pass0);
    __syncthreads();
    while (p < b1 && k < pass0 + passN)
    {
      SkewReader rd; rd.data = S.data; rd.p = p;
      Tok tk; dec_parse_rd(sp, single, rd, (uint64_t)clen - (c0 + p), tk);
      if (!tk.valid) { p = DEC_SCB; break; }
      const uint64_t sym = dec_token_symbol<W, BA, V>(tk, rd, symReg, lut);
      const uint32_t r = k - pass0;
      tOut[r] = (uint32_t)outPos; tLitLen[r] = tk.litLen; tLitSrc[r] = c0 + p + tk.hdrLen; tSym[r] = sym;
      outPos += (uint64_t)tk.litLen + tk.runLen; k++;
      if (tk.last) { p = DEC_SCB; break; }
      const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
      p = nx < DEC_SCB ? (uint32_t)nx : DEC_SCB;
    }
    // sentinel: start of the first token of the next pass (written by its owner), or the end of the SC's output
    if (k == pass0 + passN && p < b1 && pass0 + passN < total.ntok) tOut[passN] = (uint32_t)outPos;
    if (pass0 + passN >= total.ntok && t == 0) tOut[passN] = (uint32_t)min(scOut1, (uint64_t)0xFFFFFFFFu);
    __syncthreads();
    const uint64_t o0 = tOut[0];
    const uint64_t oEnd = min((uint64_t)tOut[passN], (uint64_t)n);      // never write beyond the declared size
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
      // fast path: the whole vector lies inside one run
      if (lo == vb && hi == vb + 16 && lo >= tStart + litLen && vb + 16 <= tNext)
      {
        const uint64_t sym = tSym[r];
        const uint32_t ph = (uint32_t)(lo - (tStart + litLen)) % (uint32_t)W;
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
              byte = (sp_ - c0 < DEC_SCB + DEC_PAD) ? S.data[skew8(sp_ - c0)] : __ldg(D.in + sp_);
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
}

} // namespace hsrle
