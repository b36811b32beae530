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
template <int S> __device__ __forceinline__ uint32_t toklen_plain(const TokWin &x, uint32_t avail, uint32_t &kind)
{
  const uint32_t c = win_u8<S>(x);
  const bool e1 = c == 0;
  const uint32_t cnt32 = win_u32<S + 1>(x);
  const uint32_t r = e1 ? win_u8<S + 5>(x) : win_u8<S + 1>(x);
  const uint32_t r32 = e1 ? win_u32<S + 6>(x) : win_u32<S + 2>(x);
  const bool e2 = r == 0;
  const uint32_t hdr = S + 2 + (e1 ? 4u : 0u) + (e2 ? 4u : 0u);
  const uint32_t rng = e2 ? r32 : r;
  // (a token fits iff hdr <= avail and rng - 1 <= avail - hdr: no 64-bit arithmetic; len is only used when it fits)
  kind = (hdr > avail) ? TK_BAD : (rng == 0) ? TK_END : (rng - 1 > avail - hdr) ? TK_BAD : (e1 && cnt32 == 0) ? TK_END : TK_OK;
  return hdr + rng - 1;
}
// packed tokens: b0 = same<<7 | cnt7, optional u32 cnt, optional symbol, rng in the 7-bit or the 8-bit style
template <int W, bool RNG7> __device__ __forceinline__ uint32_t toklen_packed(const TokWin &x, uint32_t avail, uint32_t &kind)
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
  kind = (hdr > avail) ? TK_BAD : endMark ? TK_END : (rng == 0 || rng - 1 > avail - hdr) ? TK_BAD : (e1 && cnt32 == 0) ? TK_END : TK_OK;
  return hdr + rng - 1;
}
// LUT tokens: u16 head = idx | cnt7 | rng, optional symbol, optional u16/u32 cnt, optional u16/u32 rng
template <int W, int K> __device__ __forceinline__ uint32_t toklen_lut(const TokWin &x, uint32_t avail, uint32_t &kind)
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
  kind = (hdr > avail) ? TK_BAD : endMark ? TK_END : (rng < 2 || rng - 2 > avail - hdr) ? TK_BAD : (cnt == 0) ? TK_END : TK_OK;
  return hdr + rng - 2;
}
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t toklen(const TokWin &x, bool single, uint32_t avail, uint32_t &kind)
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

// ================================================================================================
// D1: per-position exit tables, one CTA per SC.
//
// ex[q] (u16, SC-relative code): < EX_FAR: where the chain that starts at q leaves q's mini-block (exTab, kept for D3) /
// the SC (scTab, after the two-level in-place finalisation), a position of [c0, c0 + 0x8000); EX_END / EX_BAD;
// EX_FARP | q': the chain reaches the token at q', which jumps beyond c0 + 0x7FFF and whose absolute exit is
// farTab[c0 + q'].  D2 reads the table as codes and translates the few entries it needs; only the first DEC_WIN entries
// of an SC (its windowed exit map, read by every call of D2) are also kept as absolute positions (winTab).
constexpr int DM_T = 256;
constexpr int DM_SCOUT = 8;                // true-chain tokens the scout of D1 follows at most
constexpr uint32_t DEC_HB = 64;             // half mini-block: the unit one thread sweeps
constexpr uint32_t DEC_LIN_BYTES = DEC_SCB + 48;
__device__ __forceinline__ uint32_t skew16h(uint32_t x) { return x + ((x >> 6) << 1); }   // u16 index, one pad word per 64 entries
constexpr uint32_t DEC_EXH_ELEMS = DEC_SCB + (DEC_SCB / 64 + 1) * 2;

// The SC image is only read by phase A, which turns it into exit codes front to back, 1 KiB of stream (2112 bytes of
// table) per step of the CTA: image and table share one buffer, the image at its far end, and a block barrier per step
// keeps the table's write front below the image's read front (step it writes below 2112 (it + 1) and reads from
// DM_DATA0 + 1024 it on: DM_DATA0 >= 1088 * 15 + 2112).  35 KB instead of 50 KB per CTA: 6 CTAs per SM instead of 4.
constexpr uint32_t DM_DATA0 = 18560;
static_assert(DM_DATA0 % 16 == 0 && DM_DATA0 >= 1088 * (DEC_SCB / 1024 - 1) + 2112, "image must stay ahead of the table");
struct DecMapSmem
{
  alignas(16) uint8_t buf[DM_DATA0 + DEC_LIN_BYTES];
  __device__ __forceinline__ uint8_t *data() { return buf + DM_DATA0; }                    // linear SC image (+ the longest token head after it)
  __device__ __forceinline__ uint16_t *ex() { return reinterpret_cast<uint16_t *>(buf); }   // DEC_EXH_ELEMS entries
};
static_assert(DEC_EXH_ELEMS * 2 <= DM_DATA0 + DEC_LIN_BYTES, "table fits the buffer");
constexpr uint32_t DEC_WB = 2048;            // warp-block: the 16 mini-blocks finalised by one warp

// load stream bytes [c0, c0 + DEC_LIN_BYTES) (zero beyond clen), linear -- 16-byte coalesced
__device__ __forceinline__ void dec_load_sc_linear(uint8_t *data, const uint8_t *__restrict__ in, uint32_t c0, uint32_t clen)
{
  constexpr int NV = DEC_LIN_BYTES / 16;
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
    reinterpret_cast<uint4 *>(data)[v] = make_uint4(w[0], w[1], w[2], w[3]);
  }
}

// absolute position (or POS_END / POS_BAD) a final exit code of SC c stands for
__device__ __forceinline__ uint32_t dec_code_pos(uint32_t code, uint32_t c0, const uint32_t *farRow)
{
  if (code < EX_FAR) return c0 + code;
  if (code < EX_FARP) return code == EX_END ? POS_END : POS_BAD;
  return __ldcg(farRow + (code & 0x3FFFu));
}
// where the chain that starts at stream position x (< clen) leaves x's SC
__device__ __forceinline__ uint32_t dec_sc_exit(const DecBufs &D, uint32_t x)
{
  const uint32_t c0 = (x / DEC_SCB) * DEC_SCB;
  return dec_code_pos(__ldg(D.scTab + x), c0, D.farTab + (size_t)c0);
}

template <int W, int BA, int V>
__global__ void __launch_bounds__(DM_T) k_dec_map(const DecBufs D)
{
  constexpr Spec sp = make_spec(W, BA, V);
  extern __shared__ __align__(16) unsigned char smemRaw[];
  DecMapSmem &S = *reinterpret_cast<DecMapSmem *>(smemRaw);
  DecScalars hs; dec_header(sp, D.in, D.inSize, D.outSize, hs);
  if (blockIdx.x == 0 && threadIdx.x == 0) *D.sc = hs;
  if (hs.status != ST_OK) return;
  const uint32_t c = blockIdx.x;
  const uint32_t c0 = c * DEC_SCB;
  const uint32_t clen = hs.clen;
  if (c0 >= clen) return;
  const bool single = hs.single != 0;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  uint32_t *farRow = D.farTab + (size_t)c0;
  // Scout: the first tokens of the TRUE chain can be followed from the stream start without any table as long as each
  // of them jumps over whole SCs (incompressible input is one token with a literal of the whole input).  An SC such a
  // token jumps over holds no token start: its tables only have to be safe for the speculative chains that land in it
  // ("unparsable"), not computed.  The scout stops at the first ordinary token, i.e. after one parse on ordinary data.
  {
    __shared__ uint32_t sSkip;
    if (t == 0)
    {
      uint32_t skip = 0, p = hs.first;
      for (int i = 0; i < DM_SCOUT && p < c0; i++)
      {
        Tok tk; dec_parse(sp, single, D.in + p, (uint64_t)clen - p, tk);
        if (!tk.valid) break;
        if (tk.last) { skip = 1; break; }                             // the final token starts before this SC: nothing starts in it
        const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
        if (nx >= (uint64_t)c0 + DEC_SCB) { skip = 1; break; }        // starts before this SC, next token after it
        if (tk.litLen < 4 * DEC_SCB) break;
        p = (uint32_t)nx;
      }
      sSkip = skip;
    }
    __syncthreads();
    if (sSkip)
    {
      const uint32_t bad2 = EX_BAD | (EX_BAD << 16);
      uint4 *dst = reinterpret_cast<uint4 *>(D.scTab + (size_t)c * DEC_SCB);
      for (uint32_t q = t; q < DEC_SCB / 8; q += DM_T) dst[q] = make_uint4(bad2, bad2, bad2, bad2);
      for (uint32_t w = t; w < DEC_WIN; w += DM_T) D.winTab[(size_t)c * DEC_WIN + w] = POS_BAD;
      if (t == 0) D.scSkip[c] = 1u;
      return;
    }
  }
  uint16_t *const ex = S.ex();
  dec_load_sc_linear(S.data(), D.in, c0, clen);
  __syncthreads();

  // ---- phase A: a token parse at EVERY byte offset, four consecutive offsets per thread and step (seven aligned
  //      words give the four 24-byte windows); raw code = where that token ends
  {
    const uint32_t *d32 = reinterpret_cast<const uint32_t *>(S.data());
    const uint32_t availSC = clen - c0;              // stream bytes from the start of the SC
    for (int it = 0; it < (int)(DEC_SCB / (4 * DM_T)); it++)
    {
      const uint32_t p4 = (uint32_t)(it * DM_T + t) * 4;
      uint32_t w[7];
#pragma unroll
      for (int k = 0; k < 7; k++) w[k] = d32[(p4 >> 2) + k];
      uint32_t codes[4];
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        TokWin x;
#pragma unroll
        for (int k = 0; k < 6; k++) x.w[k] = j ? __funnelshift_r(w[k], w[k + 1], 8 * j) : w[k];
        const uint32_t p = p4 + j;
        uint32_t kind;
        const uint32_t len = toklen<W, BA, V>(x, single, p < availSC ? availSC - p : 0u, kind);
        const uint32_t nr = p + len;                  // <= availSC when the token fits
        uint32_t code = kind == TK_END ? EX_END : EX_BAD;
        if (kind == TK_OK)
        {
          code = nr;
          if (nr >= EX_FAR) { farRow[p] = c0 + nr; code = EX_FARP | p; }   // far jump: its absolute exit is parked in its own farTab entry
        }
        codes[j] = code;
      }
      uint32_t *dst = reinterpret_cast<uint32_t *>(ex + skew16h(p4));
      dst[0] = codes[0] | (codes[1] << 16); dst[1] = codes[2] | (codes[3] << 16);
      __syncthreads();                               // table write front vs image read front (DecMapSmem)
    }
  }
  __syncthreads();
  // ---- phase B: chain exits.  B1: every thread sweeps one half mini-block in reverse (a code below the end of the
  //      half points to a later entry of the same half, final already)
  {
    const uint32_t h0 = (uint32_t)t * DEC_HB, h1 = h0 + DEC_HB;
    for (uint32_t p = h1; p-- > h0;)
    {
      uint32_t code = ex[skew16h(p)];
      if (code < h1) { code = ex[skew16h(code)]; ex[skew16h(p)] = (uint16_t)code; }
    }
  }
  __syncthreads();
  //      B2: entries of the lower halves that exit into the upper half of their mini-block take its (final) entry
  for (uint32_t i = t; i < DEC_SCB / 2; i += DM_T)
  {
    const uint32_t mb = i / DEC_HB, p = mb * DEC_MB + (i % DEC_HB);
    const uint32_t code = ex[skew16h(p)];
    if (code < (mb + 1) * DEC_MB) ex[skew16h(p)] = ex[skew16h(code)];
  }
  __syncthreads();
  // keep the mini-block table for D3 (eight entries per 16-byte store)
  {
    uint4 *dst = reinterpret_cast<uint4 *>(D.exTab + (size_t)c * DEC_SCB);
    for (uint32_t q = t * 8; q < DEC_SCB; q += DM_T * 8)
    {
      const uint32_t *src = reinterpret_cast<const uint32_t *>(ex + skew16h(q));
      dst[q >> 3] = make_uint4(src[0], src[1], src[2], src[3]);
    }
  }
  __syncthreads();
  // warp-block level, in place: inside every warp-block, mini-blocks in reverse order (a code below the end of the
  // warp-block points into a later mini-block of the same warp-block, which is final already)
  {
    const uint32_t wb0 = warp * DEC_WB, wb1 = wb0 + DEC_WB;
    for (int mb = (int)(DEC_WB / DEC_MB) - 2; mb >= 0; mb--)
    {
      uint32_t code[4];
#pragma unroll
      for (int k = 0; k < 4; k++) code[k] = ex[skew16h(wb0 + mb * DEC_MB + lane + 32 * k)];
#pragma unroll
      for (int k = 0; k < 4; k++) if (code[k] < wb1) code[k] = ex[skew16h(code[k])];
#pragma unroll
      for (int k = 0; k < 4; k++) ex[skew16h(wb0 + mb * DEC_MB + lane + 32 * k)] = (uint16_t)code[k];
      __syncwarp();
    }
  }
  __syncthreads();
  // SC level: warp-blocks in reverse order
  for (int wb = (int)(DEC_SCB / DEC_WB) - 2; wb >= 0; wb--)
  {
    for (uint32_t p = wb * DEC_WB + t; p < (wb + 1) * DEC_WB; p += DM_T)
    {
      uint32_t code = ex[skew16h(p)];
      if (code < DEC_SCB) { code = ex[skew16h(code)]; ex[skew16h(p)] = (uint16_t)code; }
    }
    __syncthreads();
  }
  // keep the SC table for D2 (as codes: D2 translates the few entries it reads)
  {
    uint4 *dst = reinterpret_cast<uint4 *>(D.scTab + (size_t)c * DEC_SCB);
    for (uint32_t q = t * 8; q < DEC_SCB; q += DM_T * 8)
    {
      const uint32_t *src = reinterpret_cast<const uint32_t *>(ex + skew16h(q));
      dst[q >> 3] = make_uint4(src[0], src[1], src[2], src[3]);
    }
  }
  // absolute SC exits of the window positions
  for (uint32_t w = t; w < DEC_WIN; w += DM_T) D.winTab[(size_t)c * DEC_WIN + w] = dec_code_pos(ex[skew16h(w)], c0, farRow);
}

// ================================================================================================
// D2: k_dec_chain -- one CTA per segment of DEC_SEG SCs, one thread per window offset:
//   (1) the windowed SC-exit rows of the segment are fetched into shared memory in one round trip;
//   (2) composed in reverse SC order: suf[i][w] = where the chain that enters SC i at window offset w leaves the
//       SEGMENT (entries beyond the window -- after a long literal -- cost one finTab look-up per SC instead);
//   (3) the segment's rows are published; (4) thread 0 follows the true chain from the stream start (or from the
//       nearest chain position a preceding segment has published) through the published rows up to its own segment;
//   (5) and walks its SCs forward through the raw rows, recording every SC's true entry.
constexpr int DC_T = (int)DEC_WIN;
struct DecChainSmem
{
  uint32_t raw[DEC_SEG][DEC_WIN];
  uint32_t suf[DEC_SEG][DEC_WIN];
  uint32_t ticket, entry;
};
constexpr int DC_LOOKBACK = 8;

__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v);
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p);

static __global__ void __launch_bounds__(DC_T) k_dec_chain(const DecBufs D)
{
  extern __shared__ __align__(16) unsigned char smemRaw[];
  DecChainSmem &S = *reinterpret_cast<DecChainSmem *>(smemRaw);
  DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;                 // header check of D1: uniform over the grid
  const uint32_t w = threadIdx.x;
  if (w == 0) S.ticket = atomicAdd(&sc.segTicket, 1u);
  __syncthreads();
  const uint32_t g = S.ticket;
  const uint32_t clen = sc.clen;
  const uint32_t nSC = (clen + DEC_SCB - 1) / DEC_SCB;
  const uint32_t nSegEff = (nSC + DEC_SEG - 1) / DEC_SEG;
  const uint32_t cFirst = g * DEC_SEG;
  for (uint32_t c = cFirst + w; c < min(cFirst + DEC_SEG, D.nSC); c += DC_T) D.scEntry[c] = POS_NONE;
  if (cFirst >= nSC) return;
  const uint32_t nHere = min(DEC_SEG, nSC - cFirst);
  // a segment whose SCs were all jumped over by a true token (D1's scout) is never entered by the true chain: nobody will
  // read its rows.  (The last segment still settles how the chain ended.)
  if (__syncthreads_and(w >= nHere || D.scSkip[cFirst + w] != 0u) && g != nSegEff - 1) return;
  const uint64_t segEnd = (uint64_t)(cFirst + nHere) * DEC_SCB;
  const uint64_t segBytes = (uint64_t)DEC_SEG * DEC_SCB;
  // (1)
#pragma unroll 8
  for (uint32_t i = 0; i < nHere; i++) S.raw[i][w] = __ldg(D.winTab + (size_t)(cFirst + i) * DEC_WIN + w);
  // (1b) chains that enter a later SC of the segment beyond its window need a table look-up in global memory per
  //      SC.  The first hops of all rows are independent of each other: take them now, many loads in flight, instead
  //      of one dependent round trip per row inside the sequential composition below.
  constexpr int DC_PREHOPS = 2, DC_BATCH = 16;
#pragma unroll 1
  for (int hop = 0; hop < DC_PREHOPS; hop++)
  {
#pragma unroll 1
    for (uint32_t i0 = 0; i0 < nHere; i0 += DC_BATCH)
    {
      uint32_t x[DC_BATCH], code[DC_BATCH], farv[DC_BATCH]; bool need[DC_BATCH];
#pragma unroll
      for (int k = 0; k < DC_BATCH; k++)
      {
        const uint32_t i = i0 + k;
        x[k] = i < nHere ? (hop == 0 ? S.raw[i][w] : S.suf[i][w]) : POS_NONE;
        need[k] = x[k] < POS_SPECIAL && (uint64_t)x[k] < segEnd && (x[k] % DEC_SCB) >= DEC_WIN;
      }
#pragma unroll
      for (int k = 0; k < DC_BATCH; k++) code[k] = __ldg(D.scTab + (need[k] ? x[k] : 0u));          // unconditional: all in flight together
#pragma unroll
      for (int k = 0; k < DC_BATCH; k++)
      {
        const bool far = need[k] && code[k] >= EX_FARP;
        // (the dummy address of the not-far case is a word that is always initialised: the scalars D1 wrote)
        farv[k] = __ldcg(far ? D.farTab + ((size_t)(x[k] / DEC_SCB) * DEC_SCB + (code[k] & 0x3FFFu)) : reinterpret_cast<const uint32_t *>(D.sc));
      }
#pragma unroll
      for (int k = 0; k < DC_BATCH; k++)
      {
        const uint32_t i = i0 + k;
        if (i >= nHere) continue;
        uint32_t y = x[k];
        if (need[k])
        {
          const uint32_t c0 = (x[k] / DEC_SCB) * DEC_SCB;
          y = code[k] < EX_FAR ? c0 + code[k] : code[k] < EX_FARP ? (code[k] == EX_END ? POS_END : POS_BAD) : farv[k];
        }
        S.suf[i][w] = y;
      }
    }
  }
  __syncthreads();
  // (2)
  for (int i = (int)nHere - 1; i >= 0; i--)
  {
    uint32_t x = S.suf[i][w];
    while (x < POS_SPECIAL && (uint64_t)x < segEnd)
    {
      const uint32_t c2 = x / DEC_SCB, off = x - c2 * DEC_SCB;
      if (off < DEC_WIN) { x = S.suf[c2 - cFirst][off]; break; }    // a later SC of the segment: final already
      x = dec_sc_exit(D, x);                                        // entry beyond the window: one SC at a time
    }
    S.suf[i][w] = x;
    __syncthreads();
  }
  // (3)
  {
    uint32_t *dst = D.sufExit + (size_t)cFirst * DEC_WIN;
    for (uint32_t i = 0; i < nHere; i++) dst[i * DEC_WIN + w] = S.suf[i][w];
  }
  __threadfence();
  __syncthreads();
  if (w == 0)
  {
    st_volatile_u32(D.flagSeg + g, 1u);
    // (4)
    uint32_t pos = sc.first;
    for (int k = (int)g - 1; k >= 0 && k >= (int)g - DC_LOOKBACK; k--)
      if (ld_volatile_u32(D.chainFlag + k)) { __threadfence(); pos = __ldcg(D.chainPos + k); break; }
    while (pos < POS_SPECIAL && pos < clen && (uint64_t)pos / segBytes < g)
    {
      const uint32_t c = pos / DEC_SCB, off = pos - c * DEC_SCB;
      if (off < DEC_WIN)
      { // rows are zeroed per call and an exit is never 0: the entry itself tells whether its segment has published
        uint32_t v;
        do { v = ld_volatile_u32(D.sufExit + (size_t)c * DEC_WIN + off); } while (v == 0u);
        pos = v;
      }
      else pos = dec_sc_exit(D, pos);
    }
    // pos: the first position of the true chain at or after the start of this segment (or how the chain ended)
    D.chainPos[g] = pos; __threadfence(); st_volatile_u32(D.chainFlag + g, 1u);
    S.entry = (pos < POS_SPECIAL && pos < clen && (uint64_t)pos < segEnd) ? pos : POS_NONE;
    if (g == nSegEff - 1)
    { // the last segment settles whether the chain reaches the terminator
      uint32_t x = pos;
      while (x < POS_SPECIAL)
      {
        if (x >= clen) { x = POS_BAD; break; }
        const uint32_t c = x / DEC_SCB, off = x - c * DEC_SCB;
        x = off < DEC_WIN ? S.suf[c - cFirst][off] : dec_sc_exit(D, x);
      }
      if (x != POS_END) sc.status = ST_BADSTREAM;
    }
    // (5)
    uint32_t p = S.entry;
    while (p < POS_SPECIAL && (uint64_t)p < segEnd && p < clen)
    {
      const uint32_t c = p / DEC_SCB, off = p - c * DEC_SCB;
      D.scEntry[c] = p;
      p = off < DEC_WIN ? S.raw[c - cFirst][off] : dec_sc_exit(D, p);
    }
  }
}

// ================================================================================================
// D3: k_dec_emit -- token walk, scan over the SCs (decoupled look-back), expansion; k_dec_big -- grid-wide long ops
constexpr int DX_T = 256;                   // threads per SC (the first DEC_T of them walk one mini-block each)
constexpr int DX_GROUP = DX_T;              // look-back group: aggregates of the group + inclusive prefix of the group before
constexpr uint32_t DX_LONG_VECS = 256;      // literal copies / run fills of more 16-byte vectors than this are done by the whole CTA
constexpr uint32_t DX_HUGE_VECS = 16384;    // ... and from here on (256 KiB) by the whole grid (k_dec_big)
constexpr int DX_BIGCAP = 96;

template <int K> struct DecEmitSmem
{
  alignas(16) uint8_t data[DEC_DATA_BYTES];    // skewed SC image
  union
  {
    alignas(16) uint16_t ex[DEC_SCB];          // exit table from D1 (linear), only needed to mark the chain
    struct
    {
      uint64_t tSym[DEC_TOKCAP];               // token records of an expansion pass
      uint32_t tOut[DEC_TOKCAP + 4];
      uint32_t tLitLen[DEC_TOKCAP];
      uint32_t tLitSrc[DEC_TOKCAP];
    } rec;
  } u;
  uint32_t mbEntry[DEC_T];                     // SC-relative entry of the true chain into every mini-block (or 0xFFFF)
  DecAgg<K> warpAgg[DX_T / 32 + 1];
  DecAgg<K> bc;
  DecBigOp big[DX_BIGCAP];
  uint4 lowMask[17];                           // lowMask[i]: the bytes below i of a 16-byte vector
  uint32_t nBig, ticket, flag;
};

__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v)
{
  asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// field-wise helpers (K == 0 aggregates carry no LUT transform: nothing of it is moved)
template <int K> __device__ __forceinline__ DecAgg<K> decagg_shfl_up(const DecAgg<K> &v, int d)
{
  DecAgg<K> r;
  r.out = __shfl_up_sync(0xFFFFFFFFu, v.out, d);
  r.ntok = __shfl_up_sync(0xFFFFFFFFu, v.ntok, d);
  r.has = __shfl_up_sync(0xFFFFFFFFu, v.has, d);
  r.sym = __shfl_up_sync(0xFFFFFFFFu, v.sym, d);
  if (K)
  {
#pragma unroll
    for (int i = 0; i < 7; i++) if (i < K) r.xf.e[i] = __shfl_up_sync(0xFFFFFFFFu, v.xf.e[i], d);
  }
  return r;
}
template <int K> __device__ __forceinline__ void decagg_store(DecAgg<K> *dst, const DecAgg<K> &v)
{
  dst->out = v.out; dst->ntok = v.ntok; dst->has = v.has; dst->sym = v.sym;
  if (K)
  {
#pragma unroll
    for (int i = 0; i < 7; i++) if (i < K) dst->xf.e[i] = v.xf.e[i];
  }
}
// L2 loads (the aggregates are written by other CTAs while this kernel runs: never through the non-coherent L1)
template <int K> __device__ __forceinline__ DecAgg<K> decagg_load_cg(const DecAgg<K> *src)
{
  DecAgg<K> r = decagg_identity<K>();
  r.out = __ldcg(&src->out); r.ntok = __ldcg(&src->ntok); r.has = __ldcg(&src->has); r.sym = __ldcg(&src->sym);
  if (K)
  {
#pragma unroll
    for (int i = 0; i < 7; i++) if (i < K) r.xf.e[i] = __ldcg(&src->xf.e[i]);
  }
  return r;
}

// exclusive scan over the CTA in thread order; total = combination of everything.  warpBuf holds nWarps + 1 entries.
template <int K> __device__ __forceinline__ DecAgg<K> dec_block_excl_scan(DecAgg<K> *warpBuf, const DecAgg<K> &mine, DecAgg<K> &total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nw = blockDim.x >> 5;
  DecAgg<K> inc = mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const DecAgg<K> o = decagg_shfl_up<K>(inc, d);
    if (lane >= d) inc = decagg_combine<K>(o, inc);
  }
  if (lane == 31) decagg_store<K>(&warpBuf[warp], inc);
  DecAgg<K> ex = decagg_shfl_up<K>(inc, 1);
  if (lane == 0) ex = decagg_identity<K>();
  __syncthreads();
  if (warp == 0)
  { // the warp totals are scanned by one warp: exclusive prefix per warp, grand total in slot nw
    DecAgg<K> w = decagg_identity<K>();
    if (lane < nw) w = warpBuf[lane];
    for (int d = 1; d < nw; d <<= 1)
    {
      const DecAgg<K> o = decagg_shfl_up<K>(w, d);
      if (lane >= d) w = decagg_combine<K>(o, w);
    }
    DecAgg<K> wex = decagg_shfl_up<K>(w, 1);
    if (lane == 0) wex = decagg_identity<K>();
    if (lane < nw) decagg_store<K>(&warpBuf[lane], wex);
    if (lane == nw - 1) decagg_store<K>(&warpBuf[nw], w);
  }
  __syncthreads();
  const DecAgg<K> pre = warpBuf[warp];
  total = warpBuf[nw];
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
      lutxf_touch(mine.xf, K, idx, c0 + p + tk.symOff);
    }
    else if (tk.symKind == 0) { mine.has = 1; mine.sym = rd_sym(rd, tk.symOff, W); }
    if (tk.last) { sawEnd = true; break; }
    const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
    p = nx < DEC_SCB ? (uint32_t)nx : DEC_SCB;
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

// ---- 16-byte output vectors
// stream bytes [src, src+16) (any alignment; only aligned words holding at least one of them are read)
__device__ __forceinline__ uint4 dec_lit_vec(const uint8_t *__restrict__ in, uint32_t src)
{
  const uint32_t sb = src & 3u;
  const uint32_t *sw = reinterpret_cast<const uint32_t *>(in + (src - sb));
  const uint32_t w0 = __ldg(sw), w1 = __ldg(sw + 1), w2 = __ldg(sw + 2), w3 = __ldg(sw + 3);
  if (sb == 0) return make_uint4(w0, w1, w2, w3);
  const uint32_t w4 = __ldg(sw + 4), sh = sb * 8;
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
// as dec_lit_vec for a vector of which only bytes [a, b) (0 <= a < b <= 16) are needed: words holding none of them are
// not read (they may lie before the stream or after its end), `src` may be negative
__device__ __forceinline__ uint4 dec_lit_vec_part(const uint8_t *__restrict__ in, int64_t src, uint32_t a, uint32_t b)
{
  const uint32_t sb = (uint32_t)src & 3u;
  const uint8_t *base = in + (src - (int64_t)sb);                  // aligned stream position of word 0
  uint32_t w[5];
#pragma unroll
  for (int k = 0; k < 5; k++)
  { // word k holds the vector bytes [4k - sb, 4k - sb + 4)
    const int v0 = 4 * k - (int)sb;
    w[k] = (v0 < (int)b && v0 + 4 > (int)a) ? __ldg(reinterpret_cast<const uint32_t *>(base + 4 * k)) : 0u;
  }
  const uint32_t sh = sb * 8;
  return make_uint4(__funnelshift_r(w[0], w[1], sh), __funnelshift_r(w[1], w[2], sh), __funnelshift_r(w[2], w[3], sh), __funnelshift_r(w[3], w[4], sh));
}
// 16 bytes of the period-W pattern `sym` starting at pattern offset ph (0 <= ph < W)
template <int W> __device__ __forceinline__ uint4 dec_run_vec(uint64_t sym, uint32_t ph)
{
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; j++) w[j] = pattern_word(sym, W, (ph + 4 * j) % W);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <int W> __device__ __forceinline__ void dec_big_vec(const DecBigOp &op, uint32_t k, const uint8_t *__restrict__ in, uint8_t *__restrict__ out)
{ // vector k of a long operation
  const uint32_t v = op.v0 + k;
  uint4 x;
  if (op.kind == 0) x = dec_lit_vec(in, op.src + 16u * k);
  else x = dec_run_vec<W>(op.sym, (uint32_t)(((uint64_t)v * 16 - op.src) % (uint32_t)W));
  *reinterpret_cast<uint4 *>(out + (size_t)v * 16) = x;
}

// the CTA's last act: the last CTA of the grid settles the status and the result
__device__ __forceinline__ void dec_emit_done(const DecBufs &D, uint32_t *flagSlot)
{
  __syncthreads();
  if (threadIdx.x == 0)
  {
    __threadfence();
    *flagSlot = (atomicAdd(&D.sc->done, 1u) == gridDim.x - 1) ? 1u : 0u;
    if (*flagSlot)
    {
      __threadfence();
      DecScalars &sc = *D.sc;
      uint32_t status = ld_volatile_u32(&sc.status);
      if (status == ST_OK)
      {
        const uint32_t bad = ld_volatile_u32(&sc.emitBad), end = ld_volatile_u32(&sc.endSeen);
        const unsigned long long tot = ld_volatile_u64(&sc.outTotal);
        if (bad || !end || tot != (unsigned long long)sc.n) { status = ST_BADSTREAM; sc.status = status; }
      }
      D.dResult[0] = status == ST_OK ? sc.n : 0; D.dResult[1] = status; D.dResult[2] = ld_volatile_u32(&sc.nTok); D.dResult[3] = D.nSC;
      D.dResult[4] = sc.clen; D.dResult[5] = sc.single; D.dResult[6] = ld_volatile_u32(&sc.nHuge); D.dResult[7] = 0;
    }
  }
}

template <int W, int BA, int V>
__global__ void __launch_bounds__(DX_T) k_dec_emit(const DecBufs D)
{
  constexpr Spec sp = make_spec(W, BA, V);
  constexpr int K = sp.K;
  using Agg = DecAgg<K>;
  using Smem = DecEmitSmem<K>;
  extern __shared__ __align__(16) unsigned char smemRaw[];
  Smem &S = *reinterpret_cast<Smem *>(smemRaw);
  DecScalars &sc = *D.sc;
  const int t = threadIdx.x;
  // status as the kernels before left it: uniform over the grid (errors found here go to sc.emitBad)
  if (sc.status != ST_OK) { dec_emit_done(D, &S.flag); return; }
  if (t == 0) { S.ticket = atomicAdd(&sc.ticket, 1u); S.nBig = 0; }
  if (t < 17)
  {
    uint32_t w[4];
#pragma unroll
    for (int k = 0; k < 4; k++) { const int nb = min(max(t - 4 * k, 0), 4); w[k] = nb >= 4 ? 0xFFFFFFFFu : ((1u << (8 * nb)) - 1u); }
    S.lowMask[t] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  __syncthreads();
  const uint32_t c = S.ticket, c0 = c * DEC_SCB;
  const uint32_t clen = sc.clen, n = sc.n;
  const bool single = sc.single != 0;
  Agg *aggBuf = reinterpret_cast<Agg *>(D.aggBuf), *incBuf = reinterpret_cast<Agg *>(D.incBuf);
  const uint32_t entry = D.scEntry[c];
  const bool hasTok = entry < POS_SPECIAL;
  if (!hasTok && c != gridDim.x - 1 && (c % DX_GROUP) != DX_GROUP - 1)
  { // no token starts in this SC (it lies inside a long literal) and nobody needs its prefix: publish the identity and go
    if (t == 0) { decagg_store<K>(&aggBuf[c], decagg_identity<K>()); __threadfence(); st_volatile_u32(D.flagAgg + c, 1u); }
    dec_emit_done(D, &S.flag);
    return;
  }

  // ---- the SC's tokens: chain marks, walk #1
  Agg mine = decagg_identity<K>();
  uint32_t myEntry = 0xFFFFu;
  if (hasTok)
  {
    dec_load_sc(S.data, D.in, c0, clen);
    {
      const uint4 *src = reinterpret_cast<const uint4 *>(D.exTab + (size_t)c * DEC_SCB);
      uint4 *dst = reinterpret_cast<uint4 *>(S.u.ex);
      for (uint32_t v = t; v < DEC_SCB * 2 / 16; v += DX_T) dst[v] = __ldg(src + v);
    }
    if (t < DEC_T) S.mbEntry[t] = 0xFFFFu;
    __syncthreads();
    if (t == 0)
    {
      uint32_t p = entry - c0;
      while (p < DEC_SCB)
      {
        S.mbEntry[p / DEC_MB] = p;
        const uint32_t code = S.u.ex[p];
        p = code < EX_FAR ? code : DEC_SCB;          // leaves the SC (or ends / breaks inside this mini-block)
      }
    }
    __syncthreads();
    bool sawEnd = false, sawBad = false;
    if (t < DEC_T)
    {
      myEntry = S.mbEntry[t];
      if (myEntry != 0xFFFFu) dec_walk_sizes<W, BA, V>(S.data, myEntry, c0, clen, single, mine, sawEnd, sawBad);
    }
    if (sawBad) sc.emitBad = 1;
    if (sawEnd) sc.endSeen = 1;
  }
  Agg total;
  const Agg pre = dec_block_excl_scan<K>(S.warpAgg, mine, total);

  // ---- scan over the SCs: publish my aggregate, gather the group's aggregates and the prefix of the group before
  if (t == 0) { decagg_store<K>(&aggBuf[c], total); __threadfence(); st_volatile_u32(D.flagAgg + c, 1u); }
  const uint32_t g0 = (c / DX_GROUP) * DX_GROUP;
  Agg part = decagg_identity<K>();
  {
    const uint32_t p = g0 + t;
    if (p < c)
    {
      while (ld_volatile_u32(D.flagAgg + p) == 0u) { }
      __threadfence();
      part = decagg_load_cg<K>(&aggBuf[p]);
    }
  }
  Agg exclusive;
  (void)dec_block_excl_scan<K>(S.warpAgg, part, exclusive);
  if (g0 > 0)
  {
    if (t == 0)
    {
      while (ld_volatile_u32(D.flagInc + (g0 - 1)) == 0u) { }
      __threadfence();
      const Agg before = decagg_load_cg<K>(&incBuf[g0 - 1]);
      decagg_store<K>(&S.bc, before);
    }
    __syncthreads();
    const Agg before = S.bc;
    exclusive = decagg_combine<K>(before, exclusive);
  }
  const bool lastOfGrid = c == gridDim.x - 1;
  if (t == 0 && ((c % DX_GROUP) == DX_GROUP - 1 || lastOfGrid))
  {
    const Agg inclusive = decagg_combine<K>(exclusive, total);
    decagg_store<K>(&incBuf[c], inclusive); __threadfence(); st_volatile_u32(D.flagInc + c, 1u);
    if (lastOfGrid) { st_volatile_u64(&sc.outTotal, inclusive.out); st_volatile_u32(&sc.nTok, inclusive.ntok); }
  }
  if (!hasTok || total.ntok == 0) { dec_emit_done(D, &S.flag); return; }

  // ---- state at the start of my mini-block
  const Agg before = decagg_combine<K>(exclusive, pre);
  uint64_t symReg = single ? (uint64_t)sc.singleSym : before.sym;     // register starts as zero (src/rleX_extreme_cpu_decode.h:33)
  Lut lut; lut_init(lut, W);
  if (K && myEntry != 0xFFFFu) { Lut l0 = lut; lutxf_apply(before.xf, K, W, D.in, l0, lut); }
  uint64_t outPos = before.out;
  const uint64_t scOut1 = exclusive.out + total.out;
  uint64_t *tSym = S.u.rec.tSym;
  uint32_t *tOut = S.u.rec.tOut, *tLitLen = S.u.rec.tLitLen, *tLitSrc = S.u.rec.tLitSrc;
  const uint8_t *__restrict__ in = D.in;
  uint8_t *__restrict__ out = D.out;

  // ---- expansion in passes of DEC_TOKCAP tokens
  uint32_t p = myEntry == 0xFFFFu ? DEC_SCB : myEntry;
  uint32_t k = pre.ntok;                                               // my next token index inside the SC
  const uint32_t b1 = (t + 1) * DEC_MB;
  for (uint32_t pass0 = 0; pass0 < total.ntok; pass0 += DEC_TOKCAP)
  {
    const uint32_t passN = min(DEC_TOKCAP, total.ntok - pass0);
    __syncthreads();                                                   // the exit table / the previous pass's records are dead
    if (t < DEC_T)
    {
      while (p < b1 && k < pass0 + passN)
      {
        SkewReader rd; rd.data = S.data; rd.p = p;
        Tok tk; dec_parse_rd(sp, single, rd, (uint64_t)clen - (c0 + p), tk);
        if (!tk.valid) { p = DEC_SCB; break; }
        const uint64_t sym = dec_token_symbol<W, BA, V>(tk, rd, symReg, lut);
        const uint32_t r = k - pass0;
        tOut[r] = (uint32_t)min(outPos, (uint64_t)n); tLitLen[r] = tk.litLen; tLitSrc[r] = c0 + p + tk.hdrLen; tSym[r] = sym;
        outPos += (uint64_t)tk.litLen + tk.runLen; k++;
        if (tk.last) { p = DEC_SCB; break; }
        const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
        p = nx < DEC_SCB ? (uint32_t)nx : DEC_SCB;
      }
      // sentinel: start of the first token of the next pass (written by its owner), or the end of the SC's output
      if (k == pass0 + passN && p < b1 && pass0 + passN < total.ntok) tOut[passN] = (uint32_t)min(outPos, (uint64_t)n);
      if (pass0 + passN >= total.ntok && t == 0) tOut[passN] = (uint32_t)min(scOut1, (uint64_t)n);   // never write beyond the declared size
    }
    __syncthreads();
    const uint32_t pStart = tOut[0], pEnd = tOut[passN];
    if (pStart >= pEnd) continue;

    // -- phase M: vectors that mix segments (literal / run of neighbouring tokens), one owning token per thread;
    //    item passN is the partial vector at the start of the pass (its first bytes belong to the pass / SC before).
    //    A vector is assembled segment by segment: 16 bytes of literal source or run pattern, masked to the segment.
    {
      auto mixed = [&](uint32_t vb, uint32_t r0)
      { // all positions are output offsets <= pEnd <= n: 32-bit arithmetic throughout
        const uint32_t lo = max(vb, pStart), hi = (vb > 0xFFFFFFEFu || pEnd - vb < 16u) ? pEnd : vb + 16u;
        uint32_t rr = r0;
        uint32_t tStart = tOut[rr], tNext = tOut[rr + 1];
        uint32_t litLen = tLitLen[rr];
        uint32_t a0 = 0, a1 = 0, a2 = 0, a3 = 0;
        uint32_t x = lo;
        while (x < hi)
        {
          while (x >= tNext) { rr++; tStart = tNext; tNext = tOut[rr + 1]; litLen = tLitLen[rr]; }
          const uint32_t litEnd = litLen >= tNext - tStart ? tNext : tStart + litLen;
          uint4 cv; uint32_t segEnd;
          if (x < litEnd)
          {
            segEnd = min(litEnd, hi);
            cv = dec_lit_vec_part(in, (int64_t)tLitSrc[rr] + ((int64_t)vb - (int64_t)tStart), x - vb, segEnd - vb);
          }
          else
          {
            segEnd = min(tNext, hi);
            cv = dec_run_vec<W>(tSym[rr], (vb + 48u - litEnd) % (uint32_t)W);     // small and positive even if vb + 48 wraps
          }
          const uint4 mb = S.lowMask[segEnd - vb], ma = S.lowMask[x - vb];   // bytes [x, segEnd) of the vector
          a0 |= cv.x & mb.x & ~ma.x; a1 |= cv.y & mb.y & ~ma.y;
          a2 |= cv.z & mb.z & ~ma.z; a3 |= cv.w & mb.w & ~ma.w;
          x = segEnd;
        }
        if (lo == vb && hi - vb == 16u) *reinterpret_cast<uint4 *>(out + vb) = make_uint4(a0, a1, a2, a3);
        else
        {
          const uint32_t ba = lo - vb, bb = hi - vb;
#pragma unroll
          for (int i = 0; i < 16; i++)
          {
            const uint32_t wv = (i >> 2) == 0 ? a0 : (i >> 2) == 1 ? a1 : (i >> 2) == 2 ? a2 : a3;
            if ((uint32_t)i >= ba && (uint32_t)i < bb) out[(size_t)vb + i] = (uint8_t)(wv >> (8 * (i & 3)));
          }
        }
      };
      for (uint32_t r = t; r <= passN; r += DX_T)
      {
        if (r == passN) { if (pStart & 15u) mixed(pStart & ~15u, 0); }
        else
        {
          const uint32_t o = tOut[r], o1 = tOut[r + 1];
          const uint32_t m = (uint32_t)min((uint64_t)o + tLitLen[r], (uint64_t)o1);
          if (m > o && (m & 15u) && (m & ~15u) >= o) mixed(m & ~15u, r);
          if (o1 > m && (o1 & 15u) && (o1 & ~15u) >= m) mixed(o1 & ~15u, r);
        }
      }
    }

    // -- phase F: vectors inside one literal / one run, eight lanes per token; long ones are deferred
    {
      auto defer = [&](const DecBigOp &op)
      {
        if (op.nv >= DX_HUGE_VECS) { D.hugeList[atomicAdd(&sc.nHuge, 1u)] = op; return; }
        const uint32_t slot = atomicAdd(&S.nBig, 1u);
        if (slot < (uint32_t)DX_BIGCAP) S.big[slot] = op;
        else D.medList[atomicAdd(&sc.nMed, 1u)] = op;        // more long operations than the CTA's list holds: k_dec_big
      };
      const int grp = t >> 3, l8 = t & 7;
      for (uint32_t r = grp; r < passN; r += DX_T / 8)
      {
        const uint32_t o = tOut[r], o1 = tOut[r + 1];
        const uint32_t m = (uint32_t)min((uint64_t)o + tLitLen[r], (uint64_t)o1);
        { // literal: vectors [ceil(o/16), floor(m/16))
          const uint32_t v0 = (o >> 4) + ((o & 15u) ? 1u : 0u), v1 = m >> 4;
          if (v1 > v0)
          {
            const uint32_t src0 = tLitSrc[r] + (v0 * 16u - o);
            if (v1 - v0 > DX_LONG_VECS)
            {
              if (l8 == 0)
              {
                DecBigOp op; op.v0 = v0; op.nv = v1 - v0; op.src = src0; op.kind = 0; op.sym = 0;
                defer(op);
              }
            }
            else for (uint32_t v = v0 + l8; v < v1; v += 8) *reinterpret_cast<uint4 *>(out + (size_t)v * 16) = dec_lit_vec(in, src0 + (v - v0) * 16u);
          }
        }
        { // run: vectors [ceil(m/16), floor(o1/16))
          const uint32_t v0 = (m >> 4) + ((m & 15u) ? 1u : 0u), v1 = o1 >> 4;
          if (v1 > v0)
          {
            const uint64_t sym = tSym[r];
            if (v1 - v0 > DX_LONG_VECS)
            {
              if (l8 == 0)
              {
                DecBigOp op; op.v0 = v0; op.nv = v1 - v0; op.src = m; op.kind = 1; op.sym = sym;
                defer(op);
              }
            }
            else for (uint32_t v = v0 + l8; v < v1; v += 8) *reinterpret_cast<uint4 *>(out + (size_t)v * 16) = dec_run_vec<W>(sym, (v * 16u - m) % (uint32_t)W);
          }
        }
      }
    }
    __syncthreads();
    // -- long operations: the whole CTA, or (huge) the whole grid in k_dec_big
    {
      const uint32_t nb = min(S.nBig, (uint32_t)DX_BIGCAP);
      for (uint32_t i = 0; i < nb; i++)
      {
        const DecBigOp op = S.big[i];
        for (uint32_t kk = t; kk < op.nv; kk += DX_T) dec_big_vec<W>(op, kk, in, out);
      }
      __syncthreads();
      if (t == 0) S.nBig = 0;
    }
  }
  dec_emit_done(D, &S.flag);
}

// grid-wide execution of the huge literal copies / run fills (a 1 GiB single-symbol frame is ONE run)
constexpr uint32_t DBIG_PIECE = 1024;      // vectors per CTA step (16 KiB)
template <int W>
__global__ void __launch_bounds__(256) k_dec_big(const DecBufs D)
{
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t nMed = sc.nMed;
  for (uint32_t i = blockIdx.x; i < nMed; i += gridDim.x)
  {
    const DecBigOp op = D.medList[i];
    for (uint32_t kk = threadIdx.x; kk < op.nv; kk += 256) dec_big_vec<W>(op, kk, D.in, D.out);
  }
  const uint32_t nHuge = sc.nHuge;
  for (uint32_t i = 0; i < nHuge; i++)
  {
    const DecBigOp op = D.hugeList[i];
    const uint32_t nPieces = (op.nv + DBIG_PIECE - 1) / DBIG_PIECE;
    for (uint32_t pc = blockIdx.x; pc < nPieces; pc += gridDim.x)
    {
      const uint32_t k0 = pc * DBIG_PIECE, k1 = min(op.nv, k0 + DBIG_PIECE);
      for (uint32_t kk = k0 + threadIdx.x; kk < k1; kk += 256) dec_big_vec<W>(op, kk, D.in, D.out);
    }
  }
}

} // namespace hsrle
