// hsrle_dec_kernels.cuh -- sm_100a kernels of the decoder (see hsrle_dec.cuh for the pipeline).
#pragma once
#include <cuda_runtime.h>
#include "hsrle_dec.cuh"
#include "hsrle_enc_kernels.cuh"   // shfl helpers, volatile access

namespace hsrle {

// ================================================================================================
// small device utilities
__device__ __forceinline__ uint32_t ld_volatile_u32(const uint32_t *p) { return ld_volatile_u32g(p); }
__device__ __forceinline__ void st_volatile_u32(uint32_t *p, uint32_t v) { st_volatile_u32g(p, v); }

__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t *p)
{
  uint32_t v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_u32(uint32_t *p, uint32_t v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// ---- 1-D bulk copies (TMA) global -> shared, completion on an mbarrier
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, uint32_t count)
{
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, uint32_t bytes)
{
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// dst, src 16-byte aligned, bytes a multiple of 16
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, unsigned long long *bar)
{
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// (bounded: a copy that never lands -- it cannot, with the sizes and alignments used here -- must not hang the device; the
//  caller records the failure and the call returns 0)
__device__ __forceinline__ bool mbar_wait(unsigned long long *bar, uint32_t parity)
{
  uint32_t ok = 0;
  for (uint32_t spin = 0; !ok && spin < (1u << 22); spin++)
  {
    asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
  }
  return ok != 0;
}
// generic-proxy accesses to shared memory before this point are ordered before later async-proxy (bulk copy) accesses
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ------------------------------------------------------------------------------------------------
// token parse from a 24-byte register window (bytes p .. p+23 of the stream).  Same decisions as dec_parse
// (hsrle_core.cuh), branch free.
struct TokWin { uint32_t w[6]; };
template <int K> __device__ __forceinline__ uint32_t win_u8(const TokWin &x) { return (x.w[K >> 2] >> (8 * (K & 3))) & 0xFFu; }
template <int K> __device__ __forceinline__ uint32_t win_u32(const TokWin &x)
{
  if constexpr ((K & 3) == 0) return x.w[K >> 2];
  else return __funnelshift_r(x.w[K >> 2], x.w[(K >> 2) + 1], 8 * (K & 3));
}
enum : uint32_t { TK_OK = 0, TK_END = 1, TK_BAD = 2 };
// what K2 needs beyond the length: header bytes, stored count, where the explicit symbol sits (0xFF: none), LUT index
struct TokF { uint32_t hdr, cnt, symOff, idx; };

// [S symbol bytes][cnt][rng] with 8-bit fields and 0-escapes (plain tokens; S = 0 for single-symbol streams)
template <int S> __device__ __forceinline__ uint32_t toklen_plain(const TokWin &x, uint32_t avail, uint32_t &kind, TokF &f)
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
  f.hdr = hdr; f.cnt = e1 ? cnt32 : c; f.symOff = S ? 0u : 0xFFu; f.idx = 0;
  return hdr + rng - 1;
}
// packed tokens: b0 = same<<7 | cnt7, optional u32 cnt, optional symbol, rng in the 7-bit or the 8-bit style
template <int W, bool RNG7> __device__ __forceinline__ uint32_t toklen_packed(const TokWin &x, uint32_t avail, uint32_t &kind, TokF &f)
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
  f.hdr = hdr; f.cnt = e1 ? cnt32 : (b0 & 0x7F); f.symOff = same ? 0xFFu : (e1 ? 5u : 1u); f.idx = 0;
  return hdr + rng - 1;
}
// LUT tokens: u16 head = idx | cnt7 | rng, optional symbol, optional u16/u32 cnt, optional u16/u32 rng
template <int W, int K> __device__ __forceinline__ uint32_t toklen_lut(const TokWin &x, uint32_t avail, uint32_t &kind, TokF &f)
{
  constexpr int RB = (K == 3) ? 7 : 6;
  const uint32_t head = win_u32<0>(x) & 0xFFFFu;
  const uint32_t idx = head >> (K == 3 ? 14 : 13);
  const bool miss = idx == (uint32_t)K;
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
  f.hdr = hdr; f.cnt = cnt; f.symOff = miss ? 2u : 0xFFu; f.idx = idx;
  return hdr + rng - 2;
}
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t toklen(const TokWin &x, bool single, uint32_t avail, uint32_t &kind, TokF &f)
{
  constexpr Spec sp = make_spec(W, BA, V);
  if constexpr (sp.K != 0) return toklen_lut<W, sp.K>(x, avail, kind, f);
  else
  {
    if constexpr (W == 1) { if (single) return toklen_plain<0>(x, avail, kind, f); }
    if constexpr (V == V_PLAIN) return toklen_plain<W>(x, avail, kind, f);
    else return toklen_packed<W, sp.rng7 != 0>(x, avail, kind, f);
  }
}
// the token at byte p of a shared-memory stream image (p + 28 bytes readable)
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t toklen_at(const uint8_t *img, uint32_t p, bool single, uint32_t avail, uint32_t &kind, TokF &f)
{
  const uint32_t *d32 = reinterpret_cast<const uint32_t *>(img) + (p >> 2);
  const uint32_t sh = (p & 3u) * 8u;
  uint32_t w[7];
#pragma unroll
  for (int k = 0; k < 7; k++) w[k] = d32[k];
  TokWin x;
#pragma unroll
  for (int k = 0; k < 6; k++) x.w[k] = __funnelshift_r(w[k], w[k + 1], sh);
  return toklen<W, BA, V>(x, single, avail, kind, f);
}
// output bytes of the run part of a token with stored count c
template <int W, int BA, int V> __device__ __forceinline__ uint32_t tok_run_bytes(uint32_t c, bool single)
{
  constexpr Spec sp = make_spec(W, BA, V);
  if (c == 0) return 0;
  if constexpr (sp.K != 0) return (W == 1 || sp.byteAlign) ? c + 1 : (c + 3 / W - 2) * W;
  else
  {
    if (single) return c + (V == V_PLAIN ? 3 : 1);
    return (W == 1 || sp.byteAlign) ? c + sp.SHORT - 1 : (c + sp.SHORT / W - 1) * W;
  }
}

// ================================================================================================
// K1: k_dec_map -- windowed exit rows per chunk, segment composition, chain resolution (anchors)
constexpr int DM_T = 256;
constexpr int DM_SCOUT = 8;                // true-chain tokens the scout follows at most
constexpr int DM_FARC = 128;               // entries of the composer's far-token cache
constexpr int DM_SAMPLE = 12;              // true tokens sampled for the dense / sparse decision
constexpr uint32_t DM_SPARSE_LEN = 256;    // mean stream bytes per sampled token from which the stream counts as sparse
constexpr uint32_t DEC_HB = 64;            // the unit one thread sweeps
__device__ __forceinline__ uint32_t skew16h(uint32_t x) { return x + ((x >> 6) << 1); }   // u16 index, one pad word per 64 entries
constexpr uint32_t DEC_EXH_ELEMS = DEC_CB + (DEC_CB / 64 + 1) * 2;
constexpr uint32_t DEC_IMG_BYTES = DEC_CB + DEC_IMG_PAD + 64;

struct DecMapSmem
{
  alignas(16) uint8_t img[DEC_IMG_BYTES];          // chunk image + pad (bulk-copy destination)
  alignas(16) uint16_t ex[DEC_EXH_ELEMS];           // per-position exit codes (skewed)
  alignas(8) unsigned long long mbar;
  uint32_t flag, pos, gBase, done;
  uint32_t nScout, scP[DM_SCOUT], scN[DM_SCOUT];     // the scout's jumps: token start, next token start
  uint8_t segSkip[DEC_SEG];                          // composer: skip flags of the segment's chunks
  uint32_t mode;                                     // 1: sparse stream (segment tables), 0: dense (window rows)
  unsigned long long farCache[DM_FARC];              // composer: far token (chunk in segment << 14 | position) << 32 | where it ends
};
constexpr uint32_t DM_STAGE_ROWS = 20;     // segment rows the (dense-mode) resolver stages at a time
static_assert(DEC_IMG_BYTES % 16 == 0 && DEC_IMG_BYTES + DEC_EXH_ELEMS * 2 >= DM_STAGE_ROWS * DEC_WINC * 4, "rows fit the image + table area");
static_assert(DEC_SEG * DEC_WINC * 4 <= DEC_IMG_BYTES + DEC_EXH_ELEMS * 2, "segment rows fit");

// absolute position (or POS_END / POS_BAD) a final exit code stands for; far-jumping tokens are parsed again from the image
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t dec_code_abs(uint32_t code, uint32_t c0, const uint8_t *img, bool single, uint32_t availSC)
{
  if (code < EX_FAR) return c0 + code;
  if (code < EX_FARP) return code == EX_END ? POS_END : POS_BAD;
  const uint32_t p = code & 0x3FFFu;
  uint32_t kind; TokF f;
  const uint32_t len = toklen_at<W, BA, V>(img, p, single, availSC - p, kind, f);
  return kind == TK_OK ? c0 + p + len : (kind == TK_END ? POS_END : POS_BAD);
}

// a far-jumping token parsed from the stream: ONE round trip for the seven aligned words that hold its 24-byte window (the buffer is
// readable up to the next 16-byte boundary after the stream: words beyond that read as zero), then the register parse of phase A
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t dec_far_abs_inl(const uint8_t *in, uint32_t tp, bool single, uint32_t clen)
{
  const uint32_t endAligned = (clen + 15u) & ~15u, base = tp & ~3u, sh = (tp & 3u) * 8u;
  uint32_t w[7];
#pragma unroll
  for (int k = 0; k < 7; k++) w[k] = (base + 4u * k < endAligned) ? __ldcg(reinterpret_cast<const uint32_t *>(in + base) + k) : 0u;
  TokWin x;
#pragma unroll
  for (int k = 0; k < 6; k++) x.w[k] = __funnelshift_r(w[k], w[k + 1], sh);
  uint32_t kind; TokF f;
  const uint32_t len = toklen<W, BA, V>(x, single, clen - tp, kind, f);
  return kind == TK_OK ? tp + len : (kind == TK_END ? POS_END : POS_BAD);      // tp + len <= clen (the token fits)
}
template <int W, int BA, int V>
__device__ __noinline__ uint32_t dec_far_abs(const uint8_t *in, uint32_t tp, bool single, uint32_t clen)
{
  return dec_far_abs_inl<W, BA, V>(in, tp, single, clen);
}

// the same for a code read back from the chunk table in global memory: far-jumping tokens are parsed from the stream
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t dec_tab_abs(const DecBufs &D, uint32_t code, uint32_t c0, bool single, uint32_t clen)
{
  constexpr Spec sp = make_spec(W, BA, V);
  if (code < EX_FAR) return c0 + code;
  if (code < EX_FARP) return code == EX_END ? POS_END : POS_BAD;
  return dec_far_abs_inl<W, BA, V>(D.in, c0 + (code & 0x3FFFu), single, clen);
}

// the resolver: follows the true chain from the stream start, leaves an anchor in every chunk it visits
template <int W, int BA, int V>
__device__ void dec_resolve(const DecBufs &D, const DecScalars &hs, DecMapSmem &S, uint32_t *rows)
{
  const int t = threadIdx.x;
  const uint32_t clen = hs.clen;
  const uint32_t nChunks = (clen + DEC_CB - 1) / DEC_CB, nSeg = (nChunks + DEC_SEG - 1) / DEC_SEG;
  const bool single = hs.single != 0;
#if defined(HSRLE_PHASE_TIMERS)
  unsigned long long rt0 = 0; if (t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(rt0));
#define HSRLE_RT(i) do { if (D.dbg && t == 0) { unsigned long long now_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_)); D.dbg[3100 + (i)] = (uint32_t)(now_ - rt0); } } while (0)
#else
#define HSRLE_RT(i) do { } while (0)
#endif
  if (S.mode)
  {
    // one look-up per SEGMENT the chain visits: segTab holds, for every stream position, where the chain through it leaves its segment
    if (t == 0)
    {
      uint32_t pos = hs.first;
      for (uint32_t guard = 0;; guard++)
      {
        if (guard > nSeg + 1u) { pos = POS_BAD; D.cnt->chainBad = 0x600; break; }   // (every step leaves a segment: never reached)
        if (pos >= POS_SPECIAL) break;
        if (pos >= clen) { pos = POS_BAD; break; }
        const uint32_t c = pos / DEC_CB;
        if (__ldcg(D.skipFlag + c)) { pos = POS_BAD; break; }          // (the true chain never lands in a chunk one of its own tokens jumps over)
        D.anchorAt[c] = pos;
        pos = __ldcg(D.segTab + pos);
      }
      S.pos = pos;
    }
    __syncthreads();
  }
  else
  {
    if (t == 0) { S.pos = hs.first; S.gBase = 0; S.done = 0; }
    __syncthreads();
    uint32_t lastAnchor = 0xFFFFFFFFu;
    for (;;)
    {
      const uint32_t g0 = S.gBase;
      // stage the rows of the first chunks of segments [g0, g0 + DM_STAGE_ROWS)
      const uint32_t nr = min((uint32_t)DM_STAGE_ROWS, nSeg - g0);
      for (uint32_t i = t; i < nr * DEC_WINC; i += DM_T)
      {
        const uint32_t g = g0 + i / DEC_WINC, w = i % DEC_WINC;
        rows[i] = __ldcg(D.sufMap + (size_t)g * DEC_SEG * DEC_WINC + w);
      }
      __syncthreads();
      if (t == 0)
      {
        uint32_t pos = S.pos;
        bool fin = false;
        for (uint32_t guard = 0;; guard++)
        {
          if (guard > (1u << 24)) { pos = POS_BAD; fin = true; D.cnt->chainBad = 0x600; break; }   // (every step advances by at least one chunk: never reached)
          if (pos >= POS_SPECIAL) { fin = true; break; }
          if (pos >= clen) { pos = POS_BAD; fin = true; break; }
          const uint32_t c = pos / DEC_CB, o = pos - c * DEC_CB, g = c / DEC_SEG;
          if (g >= g0 + nr) break;                                    // beyond the staged rows: stage again from there
          if (c != lastAnchor) { D.anchorAt[c] = pos; lastAnchor = c; }
          if (o < DEC_WINC)
          {
            if (c == g * DEC_SEG) pos = rows[(g - g0) * DEC_WINC + o];
            else pos = __ldcg(D.sufMap + (size_t)c * DEC_WINC + o);
          }
          else if (__ldcg(D.skipFlag + c)) pos = POS_BAD;
          else
          { // after a long literal.  The token at the landing first: when it leaves the chunk by itself (long literals come in
            // sequences) that is the next landing, one round trip; else the chunk table takes over from where it ends
            const uint32_t nx = dec_far_abs_inl<W, BA, V>(D.in, pos, single, clen);
            if (nx < POS_SPECIAL && nx / DEC_CB == c) pos = dec_tab_abs<W, BA, V>(D, __ldcg(D.chunkTab + (size_t)nx), c * DEC_CB, single, clen);
            else pos = nx;
          }
        }
        S.pos = pos;
        if (fin) S.done = 1; else S.gBase = pos / DEC_CB / DEC_SEG;
      }
      __syncthreads();
      if (S.done) break;
    }
  }
  if (t == 0 && S.pos != POS_END) D.cnt->chainBad = 1;
  __threadfence();
  __syncthreads();
  HSRLE_RT(0);
  // every chunk's first true token start: one thread per segment walks its chunks through the chunk tables from the anchors
  for (uint32_t g = t; g < nSeg; g += DM_T)
  {
    const uint32_t cFirst = g * DEC_SEG, cEnd = min(cFirst + DEC_SEG, nChunks);
    uint32_t x = POS_NONE;                                           // chain position (absolute), once an anchor was met
    for (uint32_t c = cFirst; c < cEnd; c++)
    {
      const uint32_t a = __ldcg(D.anchorAt + c);
      if (a) x = a;
      uint32_t e = POS_NONE;
      if (x < POS_SPECIAL && x / DEC_CB == c)
      {
        e = x;
        x = __ldcg(D.skipFlag + c) ? POS_BAD : dec_tab_abs<W, BA, V>(D, __ldcg(D.chunkTab + (size_t)c * DEC_CB + (x - c * DEC_CB)), c * DEC_CB, single, clen);
      }
      D.chunkEntry[c] = e;
    }
  }
  __syncthreads();
  HSRLE_RT(1);
  // the LIVE chunks (a true token starts in them), in stream order: K2 takes its tickets over this list -- the chunks inside
  // long literals (all but one of an incompressible 1-GiB frame) cost it nothing
  {
    const uint32_t per = (nChunks + DM_T - 1) / DM_T;
    const uint32_t lo = min(nChunks, (uint32_t)t * per), hi = min(nChunks, lo + per);
    uint32_t mine = 0;
    for (uint32_t c = lo; c < hi; c++) if (__ldcg(D.chunkEntry + c) < POS_SPECIAL) mine++;
    rows[t] = mine;
    __syncthreads();
    uint32_t off = 0, total = 0;
    for (int k = 0; k < DM_T; k++) { const uint32_t x = rows[k]; if (k < t) off += x; total += x; }
    for (uint32_t c = lo; c < hi; c++) if (__ldcg(D.chunkEntry + c) < POS_SPECIAL) D.liveList[off++] = c;
    if (t == 0)
    {
      if (total == 0) { D.liveList[0] = 0; total = 1; }               // (broken chain: K2 still has to settle the result)
      D.cnt->nLive = total;
    }
  }
  HSRLE_RT(2);
}

template <int W, int BA, int V>
__global__ void __launch_bounds__(DM_T, 4) k_dec_map(const DecBufs D)
{
  constexpr Spec sp = make_spec(W, BA, V);
  extern __shared__ __align__(16) unsigned char smemRaw[];
  DecMapSmem &S = *reinterpret_cast<DecMapSmem *>(smemRaw);
  uint32_t *const rows = reinterpret_cast<uint32_t *>(smemRaw);      // segment / resolver rows alias the image + table area
  DecScalars hs; dec_header(sp, D.in, D.inSize, D.outSize, hs);
  if (blockIdx.x == 0 && threadIdx.x == 0) *D.sc = hs;
  if (hs.status != ST_OK) return;
  const uint32_t clen = hs.clen;
  const uint32_t nChunks = (clen + DEC_CB - 1) / DEC_CB, nSeg = (nChunks + DEC_SEG - 1) / DEC_SEG;
  const bool single = hs.single != 0;
  const int t = threadIdx.x;
  uint16_t *const ex = S.ex;
  // Scout (once per CTA): the first tokens of the TRUE chain can be followed from the stream start without any table as long as
  // each of them jumps over whole chunks (incompressible input is one token with a literal of the whole input).  A chunk such a
  // token jumps over holds no token start: it gets a flag instead of a table.
  if (t == 0)
  {
    mbar_init(&S.mbar, 1);
    uint32_t k = 0, p = hs.first;
    for (int i = 0; i < DM_SCOUT; i++)
    {
      Tok tk; dec_parse(sp, single, D.in + p, (uint64_t)clen - p, tk);
      if (!tk.valid) break;
      if (tk.last) { S.scP[k] = p; S.scN[k] = 0xFFFFFFFFu; k++; break; }   // nothing starts after the final token
      const uint64_t nx = (uint64_t)p + tk.hdrLen + tk.litLen;
      if (tk.litLen < 4 * DEC_CB) break;
      S.scP[k] = p; S.scN[k] = (uint32_t)min(nx, (uint64_t)0xFFFFFFFFu); k++;
      p = (uint32_t)nx;
    }
    S.nScout = k;
  }
  uint32_t phase = 0;
  __syncthreads();                                                    // (also: mbarrier initialised before anybody waits)
  for (uint32_t c = blockIdx.x; c < nChunks; c += gridDim.x)
  {
  const uint32_t c0 = c * DEC_CB;
  const uint32_t availSC = clen - c0;                                 // stream bytes from the start of the chunk
  bool skipped = false;
  for (uint32_t i = 0; i < S.nScout; i++) skipped = skipped || (S.scP[i] < c0 && (uint64_t)S.scN[i] >= (uint64_t)c0 + DEC_CB);
  __syncthreads();                                                    // the previous chunk's rows (they alias the image) are done with
  if (skipped)
  { // no table, no rows, no image: one flag (every reader of the chunk table checks it)
    if (t == 0) D.skipFlag[c] = 1;
  }
  else
  {
    // ---- the chunk image: one bulk copy (rounded up to 16 bytes: the caller's buffer is readable up to the next 16-byte
    //      boundary after the stream, include/hsrle_b200.h)
    if (t == 0)
    {
      fence_async_smem();
      const uint32_t bytes = min(DEC_CB + DEC_IMG_PAD, (availSC + 15u) & ~15u);
      mbar_expect_tx(&S.mbar, bytes);
      bulk_load(S.img, D.in + c0, bytes, &S.mbar);
    }
    if (!mbar_wait(&S.mbar, phase)) D.cnt->chainBad = 0x100;
    phase ^= 1u;
    __syncthreads();
    // ---- phase A: a token parse at EVERY byte offset, four consecutive offsets per thread and step (seven aligned
    //      words give the four 24-byte windows); raw code = where that token ends
    {
      const uint32_t *d32 = reinterpret_cast<const uint32_t *>(S.img);
#pragma unroll 2
      for (int it = 0; it < (int)(DEC_CB / (4 * DM_T)); it++)
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
          uint32_t kind; TokF f;
          const uint32_t len = toklen<W, BA, V>(x, single, p < availSC ? availSC - p : 0u, kind, f);
          const uint32_t nr = p + len;                  // <= availSC when the token fits
          uint32_t code = kind == TK_END ? EX_END : EX_BAD;
          if (kind == TK_OK) code = nr < EX_FAR ? nr : (EX_FARP | p);
          codes[j] = code;
        }
        uint32_t *dst = reinterpret_cast<uint32_t *>(ex + skew16h(p4));
        dst[0] = codes[0] | (codes[1] << 16); dst[1] = codes[2] | (codes[3] << 16);
      }
    }
    __syncthreads();
    // ---- phase B: chain exits.  B1: every thread sweeps one 64-byte block in reverse (a code below the end of the block points
    //      to a later entry of the same block, final already)
    {
      const uint32_t h0 = (uint32_t)t * DEC_HB, h1 = h0 + DEC_HB;
      for (uint32_t p = h1; p-- > h0;)
      {
        uint32_t code = ex[skew16h(p)];
        if (code < h1) { code = ex[skew16h(code)]; ex[skew16h(p)] = (uint16_t)code; }
      }
    }
    __syncthreads();
    //      B2: block sizes 128, 256, 512: the lower half of every block takes the (final) entry of the upper half it exits into
#pragma unroll
    for (int lg = 6; lg <= 8; lg++)
    {
      constexpr uint32_t one = 1u;
      const uint32_t half = one << lg;
#pragma unroll 4
      for (uint32_t i = t; i < DEC_CB / 2; i += DM_T)
      {
        const uint32_t blk = i >> lg, p = (blk << (lg + 1)) + (i & (half - 1));
        const uint32_t code = ex[skew16h(p)];
        if (code < ((blk + 1) << (lg + 1))) ex[skew16h(p)] = ex[skew16h(code)];
      }
      __syncthreads();
    }
    //      B3: sub-chunk level, window offsets only (after B2 every entry of a 512-byte block leaves the block)
#pragma unroll 4
    for (uint32_t s = 0; s < (uint32_t)DEC_NSUB; s++)
    {
      for (uint32_t w = t; w < DEC_WIN; w += DM_T)
      {
        const uint32_t p = s * DEC_SB + w;
        const uint32_t code = ex[skew16h(p)];
        if (code < (s + 1) * DEC_SB) ex[skew16h(p)] = ex[skew16h(code)];
      }
    }
    __syncthreads();
    // ---- the sub-chunk rows (K2 walks the sub-chunks from them)
    {
      uint32_t *dst = reinterpret_cast<uint32_t *>(D.subMap + (size_t)c * DEC_NSUB * DEC_WIN);
#pragma unroll 4
      for (uint32_t s = 0; s < (uint32_t)DEC_NSUB; s++)
      {
        if (t < DEC_WIN / 2)
        {
          const uint32_t p = s * DEC_SB + 2u * t;
          dst[s * (DEC_WIN / 2) + t] = (uint32_t)ex[skew16h(p)] | ((uint32_t)ex[skew16h(p + 1)] << 16);
        }
      }
    }
    __syncthreads();
    // ---- chunk level, in place, for EVERY offset.  First every warp finishes its own 2-KiB region (512-byte blocks in reverse order:
    //      a code below the end of the region points into a later block of it, final already; warp barriers only), then the regions
    //      in reverse order by the whole CTA (a code below the end of the chunk points into a later region) -- one look-up per offset
    {
      const uint32_t r0 = (uint32_t)(t >> 5) * 2048u, r1 = r0 + 2048u;
      for (int b = 2; b >= 0; b--)
      {
#pragma unroll 4
        for (uint32_t p = r0 + (uint32_t)b * 512 + (t & 31); p < r0 + (uint32_t)(b + 1) * 512; p += 32)
        {
          const uint32_t code = ex[skew16h(p)];
          if (code < r1) ex[skew16h(p)] = ex[skew16h(code)];
        }
        __syncwarp();
      }
    }
    __syncthreads();
    for (int rg = (int)(DEC_CB / 2048) - 2; rg >= 0; rg--)
    {
#pragma unroll 4
      for (uint32_t p = (uint32_t)rg * 2048 + t; p < (uint32_t)(rg + 1) * 2048; p += DM_T)
      {
        const uint32_t code = ex[skew16h(p)];
        if (code < DEC_CB) ex[skew16h(p)] = ex[skew16h(code)];
      }
      __syncthreads();
    }
    // ---- the chunk table (eight entries per 16-byte store)
    {
      uint4 *dst = reinterpret_cast<uint4 *>(D.chunkTab + (size_t)c * DEC_CB);
      for (uint32_t q = t * 8; q < DEC_CB; q += DM_T * 8)
      {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(ex + skew16h(q));
        dst[q >> 3] = make_uint4(src[0], src[1], src[2], src[3]);
      }
    }
  }
  // ---- the last chunk of a segment composes the segment's rows; the last segment resolves the chain
  const uint32_t g = c / DEC_SEG, cFirst = g * DEC_SEG, nHere = min(DEC_SEG, nChunks - cFirst);
  __threadfence();
  __syncthreads();
  if (t == 0) S.flag = (atomicAdd(D.segCount + g, 1u) == nHere - 1) ? 1u : 0u;
  __syncthreads();
  if (!S.flag) continue;
  __threadfence();
  if (t == 0)
  { // a segment of jumped-over chunks has no rows
    uint32_t live = 0;
    for (uint32_t i = 0; i < nHere; i++) live += __ldcg(D.skipFlag + cFirst + i) ? 0u : 1u;
    S.pos = live;
    // Window rows or segment tables?  Sampled on the first true tokens (every composer gets the same answer).  Chains of short tokens
    // land inside the entry windows: window rows are all the resolver needs (the narrow codecs on compressible data).  With long
    // literals nearly every landing is deep inside a chunk and the resolver makes one or two dependent look-ups per TOKEN; a table with
    // the exit of the segment for EVERY position makes that one look-up per SEGMENT.  Building it costs a pass over the chunk tables,
    // per chunk and one chunk of a segment after the other: it pays when there is at least a token per chunk to save (mean token
    // below the chunk size) and there are enough segments to keep the SMs busy (measured: 3.0 -> 1.8 ms on a 1-GiB run-mixed frame
    // of the 64-bit codecs; a loss for the 88 MB streams -- 67 segments -- and for 8-bit streams of the same frame: 34 643 chunks for
    // 5 768 tokens).
    uint32_t p = hs.first, k = 0;
    const bool manySegs = gridDim.x >= 256u && nSeg >= gridDim.x / 8u;   // (the grid is 8 CTAs per SM when there are that many chunks)
    for (; manySegs && k < (uint32_t)DM_SAMPLE && p < clen; k++)
    {
      const uint32_t nx = dec_far_abs_inl<W, BA, V>(D.in, p, single, clen);
      if (nx >= clen) break;                                          // (also POS_END / POS_BAD)
      p = nx;
    }
    const uint32_t meanTok = k ? (p - hs.first) / k : 0u;
    S.mode = (k >= 4 && meanTok >= DM_SPARSE_LEN && meanTok <= DEC_CB && manySegs) ? 1u : 0u;
    if (D.modeOverride) S.mode = D.modeOverride - 1u;
  }
  __syncthreads();
#if defined(HSRLE_PHASE_TIMERS)
  unsigned long long ct0 = 0; if (t == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ct0));
#endif
  if (S.pos != 0 && S.mode)
  { // SPARSE streams -- segment table, EVERY offset of every chunk: where the chain through that position leaves the segment (absolute position,
    // POS_END, POS_BAD).  Chunks in reverse order: an exit into a later chunk of the segment takes that position's (final) entry.
    // The look-ups repeat a few hundred distinct positions per chunk (chains merge): they hit in cache.
    const uint64_t segEnd = (uint64_t)(cFirst + nHere) * DEC_CB;
    if (t < (int)DEC_SEG) S.segSkip[t] = ((uint32_t)t < nHere) ? __ldcg(D.skipFlag + cFirst + t) : (uint8_t)1;
    for (uint32_t i = t; i < (uint32_t)DM_FARC; i += DM_T) S.farCache[i] = ~0ull;
    __syncthreads();
    for (int i = (int)nHere - 1; i >= 0; i--)
    {
      const uint32_t ci = cFirst + (uint32_t)i;
      if (!S.segSkip[i])
      {
        const uint16_t *tab = D.chunkTab + (size_t)ci * DEC_CB;
        uint32_t *dst = D.segTab + (size_t)ci * DEC_CB;
        constexpr int NV = (int)(DEC_CB / (DM_T * 8));
        uint4 v[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) v[k] = __ldcg(reinterpret_cast<const uint4 *>(tab + (uint32_t)(k * DM_T + t) * 8));
#pragma unroll 1
        for (int k = 0; k < NV; k++)
        {
          const uint32_t q = (uint32_t)(k * DM_T + t) * 8;
          uint4 vv = v[0];
#pragma unroll
          for (int kk = 1; kk < NV; kk++) if (kk == k) vv = v[kk];
          const uint32_t w[4] = { vv.x, vv.y, vv.z, vv.w };
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; j++)
          {
            const uint32_t code = (w[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu;
            uint32_t x;
            if (code < EX_FAR) x = ci * DEC_CB + code;
            else if (code < EX_FARP) x = code == EX_END ? POS_END : POS_BAD;
            else
            { // far-jumping token: parsed from the stream, then found in a small cache (many positions share a chain's last token)
              const uint32_t key = ((uint32_t)i << 14) | (code & 0x3FFFu), slot = (key * 2654435761u) >> 25;
              const unsigned long long e = S.farCache[slot];
              if ((uint32_t)(e >> 32) == key) x = (uint32_t)e;
              else { x = dec_far_abs<W, BA, V>(D.in, ci * DEC_CB + (code & 0x3FFFu), single, clen); S.farCache[slot] = ((unsigned long long)key << 32) | x; }
            }
            if (x < POS_SPECIAL)
            {
              if (x >= clen) x = POS_BAD;
              else if ((uint64_t)x < segEnd) x = S.segSkip[x / DEC_CB - cFirst] ? POS_BAD : D.segTab[x];   // (written by this CTA: plain, cached load)
            }
            o[j] = x;
          }
          reinterpret_cast<uint4 *>(dst + q)[0] = make_uint4(o[0], o[1], o[2], o[3]);
          reinterpret_cast<uint4 *>(dst + q)[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      }
      __syncthreads();
    }
  }
#if defined(HSRLE_PHASE_TIMERS)
  if (D.dbg && t == 0) { unsigned long long now_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now_)); atomicMax(D.dbg + 3110, (uint32_t)(now_ - ct0)); atomicAdd(D.dbg + 3111, 1u); }
#endif
  if (S.pos != 0 && !S.mode)
  { // DENSE streams -- window rows: exit of the segment per entry offset < DEC_WINC of every chunk (the resolver's other landings take the chunk tables)
    for (uint32_t i = t; i < nHere * DEC_WINC; i += DM_T)
    {
      const uint32_t ci = cFirst + i / DEC_WINC, w = i % DEC_WINC;
      rows[i] = __ldcg(D.skipFlag + ci) ? POS_BAD : dec_tab_abs<W, BA, V>(D, __ldcg(D.chunkTab + (size_t)ci * DEC_CB + w), ci * DEC_CB, single, clen);
    }
    __syncthreads();
    const uint64_t segEnd = (uint64_t)(cFirst + nHere) * DEC_CB;
    for (int i = (int)nHere - 2; i >= 0; i--)
    {
      for (uint32_t w = t; w < DEC_WINC; w += DM_T)
      {
        const uint32_t x = rows[i * DEC_WINC + w];
        if (x < POS_SPECIAL && (uint64_t)x < segEnd)
        {
          const uint32_t c2 = x / DEC_CB, off = x - c2 * DEC_CB;
          if (off < DEC_WINC) rows[i * DEC_WINC + w] = rows[(c2 - cFirst) * DEC_WINC + off];      // a later chunk of the segment: final already
        }
      }
      __syncthreads();
    }
    for (uint32_t i = t; i < nHere * DEC_WINC; i += DM_T) D.sufMap[(size_t)cFirst * DEC_WINC + i] = rows[i];
  }
  __threadfence();
  __syncthreads();
  if (t == 0) S.flag = (atomicAdd(&D.cnt->segsDone, 1u) == nSeg - 1) ? 1u : 0u;
  __syncthreads();
  if (!S.flag) continue;
  __threadfence();
  dec_resolve<W, BA, V>(D, hs, S, rows);
  }
}

// ================================================================================================
// K2: k_dec_emit -- persistent CTAs: chunk entry, token walk, look-back over the chunks, expansion, grid-wide long operations
constexpr int DX_T = 256;
constexpr uint32_t DX_INLINE = 64;          // token parts up to this many bytes (inside a tile) are written by the token's own thread
constexpr int DX_NSKIP = 16;                // whole-tile ranges of a chunk handed to the grid
constexpr uint32_t DX_LONGCAP = DEC_TILE / DX_INLINE * 2 + 8;

constexpr uint32_t DX_POSCAP = DEC_SB / 2;  // tokens that can start in one sub-chunk (the shortest token has two bytes)

template <int K> struct DecEmitSmem
{
  alignas(16) uint8_t img[DEC_IMG_BYTES];           // chunk image + pad (bulk-copy destination)
  alignas(16) uint8_t tile[DEC_TILE];                // output image of one expansion step; before that: the chunk's sub-chunk rows
  uint16_t pos[DEC_NSUB][DX_POSCAP];                  // token starts of every sub-chunk (chunk-relative), written by the walkers
  uint16_t ref[K ? DEC_NSUB : 1][K ? DX_POSCAP : 1];  // K > 0: per token, where its symbol comes from: < 8 table slot at the start of the
                                                      //        sub-chunk, else 8 + chunk-relative position of an explicit symbol
  uint32_t rOut[DEC_NSLOT + 4];                       // token records of an expansion group: output offset relative to the chunk's first
  uint32_t rLit[DEC_NSLOT];                           //   output byte (+ sentinel), literal bytes,
  uint32_t rRun[DEC_NSLOT];                           //   run bytes,
  uint16_t rSrc[DEC_NSLOT];                           //   literal source (chunk-relative),
  uint16_t rSymI[DEC_NSLOT];                          //   symbol: K == 0 chunk-relative position of the governing explicit symbol (0xFFFF: the
  uint8_t rSub[DEC_NSLOT];                            //   chunk's incoming one); K > 0 the token's `ref`; and the token's sub-chunk
  uint32_t subEntry[DEC_NSUB];
  uint32_t subCnt[DEC_NSUB + 1];                      // tokens per sub-chunk -> exclusive prefix
  uint32_t subEnd[DEC_NSUB];                          // 1: the sub-chunk's walker met the last token, 2: an unparsable one
  LutXf subXf[K ? DEC_NSUB : 1];                      // K > 0: table transform of every sub-chunk
  Lut subLut[K ? DEC_NSUB : 1];                       // K > 0: table at the start of the sub-chunk
  DecAgg<K> bc;                                       // look-back result
  DecAgg<K> lbAgg[DX_T / 32];                         // look-back: per-warp combination of the window, down to its nearest inclusive prefix
  uint32_t lbHit[DX_T / 32];
  unsigned long long warpSum[DX_T / 32 + 1];          // block scans: output bytes ...
  uint32_t warpSymI[DX_T / 32 + 1];                   // ... and last explicit symbol
  unsigned long long sumOut;                          // output bytes of the chunk / of the groups before the current one
  uint32_t sumSym;                                    // K == 0: chunk-relative position + 1 of the last explicit symbol so far (0: none)
  uint64_t inSym;                                     // K == 0: symbol register at the start of the chunk
  uint32_t longList[DX_LONGCAP];
  uint32_t skipLo[DX_NSKIP], skipHi[DX_NSKIP];
  alignas(8) unsigned long long mbar[2];
  uint32_t nLong, nSkip, ticket, entry, flag, bcast;
};

// field-wise helpers (K == 0 aggregates carry no LUT transform: nothing of it is moved)
template <int K> __device__ __forceinline__ DecAgg<K> decagg_shfl_down(const DecAgg<K> &v, int d)
{
  DecAgg<K> r;
  r.out = __shfl_down_sync(0xFFFFFFFFu, v.out, d);
  r.ntok = __shfl_down_sync(0xFFFFFFFFu, v.ntok, d);
  r.symPos = __shfl_down_sync(0xFFFFFFFFu, v.symPos, d);
  if (K)
  {
#pragma unroll
    for (int i = 0; i < 7; i++) if (i < K) r.xf.e[i] = __shfl_down_sync(0xFFFFFFFFu, v.xf.e[i], d);
  }
  return r;
}
template <int K> __device__ __forceinline__ DecAgg<K> decagg_shfl_up(const DecAgg<K> &v, int d)
{
  DecAgg<K> r;
  r.out = __shfl_up_sync(0xFFFFFFFFu, v.out, d);
  r.ntok = __shfl_up_sync(0xFFFFFFFFu, v.ntok, d);
  r.symPos = __shfl_up_sync(0xFFFFFFFFu, v.symPos, d);
  if (K)
  {
#pragma unroll
    for (int i = 0; i < 7; i++) if (i < K) r.xf.e[i] = __shfl_up_sync(0xFFFFFFFFu, v.xf.e[i], d);
  }
  return r;
}
template <int K> __device__ __forceinline__ void decagg_store(DecAgg<K> *dst, const DecAgg<K> &v)
{
  dst->out = v.out; dst->ntok = v.ntok; dst->symPos = v.symPos;
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
  r.out = __ldcg(&src->out); r.ntok = __ldcg(&src->ntok); r.symPos = __ldcg(&src->symPos);
  if (K)
  {
#pragma unroll
    for (int i = 0; i < 7; i++) if (i < K) r.xf.e[i] = __ldcg(&src->xf.e[i]);
  }
  return r;
}

// ---- output helpers
// 16 bytes of the period-W pattern `sym` starting at pattern offset ph (0 <= ph < W)
template <int W> __device__ __forceinline__ uint4 dec_run_vec(uint64_t sym, uint32_t ph)
{
  uint32_t w[4];
#pragma unroll
  for (int j = 0; j < 4; j++) w[j] = pattern_word(sym, W, (ph + 4 * j) % W);
  return make_uint4(w[0], w[1], w[2], w[3]);
}
template <int W> __device__ __forceinline__ uint32_t run_byte(uint64_t sym, uint32_t ph) { return (uint32_t)(sym >> (8 * ph)) & 0xFFu; }
// global stream bytes [src, src+16) (any alignment; only aligned words holding at least one of them are read)
__device__ __forceinline__ uint4 dec_lit_vec(const uint8_t *__restrict__ in, uint32_t src)
{
  const uint32_t sb = src & 3u;
  const uint32_t *sw = reinterpret_cast<const uint32_t *>(in + (src - sb));
  const uint32_t w0 = __ldg(sw), w1 = __ldg(sw + 1), w2 = __ldg(sw + 2), w3 = __ldg(sw + 3);
  if (sb == 0) return make_uint4(w0, w1, w2, w3);
  const uint32_t w4 = __ldg(sw + 4), sh = sb * 8;
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}
// shared-memory bytes [p, p+16) of a byte array (any alignment)
__device__ __forceinline__ uint4 smem_vec(const uint8_t *base, uint32_t p)
{
  const uint32_t *sw = reinterpret_cast<const uint32_t *>(base) + (p >> 2);
  const uint32_t sh = (p & 3u) * 8u;
  const uint32_t w0 = sw[0], w1 = sw[1], w2 = sw[2], w3 = sw[3], w4 = sw[4];
  return make_uint4(__funnelshift_r(w0, w1, sh), __funnelshift_r(w1, w2, sh), __funnelshift_r(w2, w3, sh), __funnelshift_r(w3, w4, sh));
}

// one literal byte of a token of the current chunk: from the image when it holds it, else from the stream
__device__ __forceinline__ uint32_t lit_byte(const uint8_t *img, const uint8_t *__restrict__ in, uint32_t c0, uint32_t rel)
{
  return rel < DEC_CB + DEC_IMG_PAD ? (uint32_t)img[rel] : (uint32_t)__ldg(in + (size_t)c0 + rel);
}

// run fill of `len` bytes at tile offset d by one warp: byte head to the next 16-byte boundary, vectors, byte tail
template <int W> __device__ __forceinline__ void warp_fill(uint8_t *tile, uint32_t d, uint32_t len, uint64_t sym, uint32_t ph, int lane)
{
  const uint32_t head = min(len, (16u - (d & 15u)) & 15u);
  if ((uint32_t)lane < head) tile[d + lane] = (uint8_t)run_byte<W>(sym, (ph + lane) % W);
  const uint32_t d1 = d + head, nv = (len - head) >> 4, ph1 = (ph + head) % W;
  for (uint32_t v = lane; v < nv; v += 32) *reinterpret_cast<uint4 *>(tile + d1 + 16 * v) = dec_run_vec<W>(sym, (ph1 + 16 * v) % W);
  const uint32_t done = head + (nv << 4), tail = len - done;
  if ((uint32_t)lane < tail) tile[d + done + lane] = (uint8_t)run_byte<W>(sym, (ph + done + lane) % W);
}
// literal copy of `len` bytes to tile offset d by one warp; source = chunk-relative stream position rel
__device__ __forceinline__ void warp_copy(uint8_t *tile, uint32_t d, uint32_t len, const uint8_t *img, const uint8_t *__restrict__ in, uint32_t c0, uint32_t rel, int lane)
{
  const uint32_t head = min(len, (16u - (d & 15u)) & 15u);
  if ((uint32_t)lane < head) tile[d + lane] = (uint8_t)lit_byte(img, in, c0, rel + lane);
  const uint32_t d1 = d + head, nv = (len - head) >> 4, r1 = rel + head;
  for (uint32_t v = lane; v < nv; v += 32)
  {
    const uint32_t r = r1 + 16 * v;
    const uint4 x = (r + 20 <= DEC_CB + DEC_IMG_PAD) ? smem_vec(img, r) : dec_lit_vec(in, c0 + r);
    *reinterpret_cast<uint4 *>(tile + d1 + 16 * v) = x;
  }
  const uint32_t done = head + (nv << 4), tail = len - done;
  if ((uint32_t)lane < tail) tile[d + done + lane] = (uint8_t)lit_byte(img, in, c0, rel + done + lane);
}

// one piece (<= DEC_BIG_PIECE bytes, 16-byte aligned destination) of a grid-wide operation, by the whole CTA
template <int W, int K>
__device__ void dec_big_piece(const DecBufs &D, DecEmitSmem<K> &S, const DecBigOp &op, uint32_t off, uint32_t len, uint32_t &phase0, uint32_t &phase1)
{
  const int t = threadIdx.x;
  uint8_t *__restrict__ out = D.out + (size_t)op.dst + off;
  if (op.kind == 1)
  { // run fill: pattern phase of the piece's first byte
    const uint32_t ph0 = (uint32_t)(((uint64_t)op.dst + off - op.src) % (uint32_t)W);
    for (uint32_t v = t; v < (len >> 4); v += DX_T) reinterpret_cast<uint4 *>(out)[v] = dec_run_vec<W>(op.sym, (ph0 + 16 * v) % W);
    return;
  }
  // literal copy: unaligned source words straight from the stream (L2), 16-byte stores
  const uint32_t src0 = op.src + off;                       // stream position of the piece's first byte
  (void)S; (void)phase0; (void)phase1;
  for (uint32_t v = t; v < (len >> 4); v += DX_T) reinterpret_cast<uint4 *>(out)[v] = dec_lit_vec(D.in, src0 + 16u * v);
}

// length of the token at x for the walkers: the common shapes (no escaped count / range field) from two or three bytes of the
// image, everything else through the general parse.  Returns the kind; len as toklen.
template <int W, int BA, int V>
__device__ __forceinline__ uint32_t tok_walk_len(const uint8_t *img, uint32_t x, bool single, uint32_t avail, uint32_t &len, uint32_t &idx)
{
  constexpr Spec sp = make_spec(W, BA, V);
  constexpr int K = sp.K;
  idx = 0;
  if constexpr (K != 0)
  {
    constexpr int RB = (K == 3) ? 7 : 6;
    const uint32_t head = (uint32_t)img[x] | ((uint32_t)img[x + 1] << 8);
    idx = head >> (K == 3 ? 14 : 13);
    const uint32_t c7 = (head >> RB) & 0x7F, r = head & ((1u << RB) - 1u);
    const uint32_t hdr = 2u + (idx == (uint32_t)K ? (uint32_t)W : 0u);
    if (c7 >= 2 && r >= 2 && hdr <= avail && r - 2 <= avail - hdr) { len = hdr + r - 2; return TK_OK; }
  }
  else if (V == V_PLAIN || (W == 1 && single))
  {
    const uint32_t S = (W == 1 && single) ? 0u : (uint32_t)W;
    const uint32_t c = img[x + S], r = img[x + S + 1];
    if (c != 0 && r != 0 && S + 2 <= avail && r - 1 <= avail - (S + 2)) { len = S + 1 + r; return TK_OK; }
  }
  else
  {
    const uint32_t b0 = img[x];
    const uint32_t o = 1u + ((b0 & 0x80u) ? 0u : (uint32_t)W);
    const uint32_t r = img[x + o];
    if ((b0 & 0x7Fu) != 0)
    {
      if (sp.rng7) { if ((r & 1u) == 0 && r >= 2 && o + 1 <= avail && (r >> 1) - 1 <= avail - (o + 1)) { len = o + (r >> 1); return TK_OK; } }
      else if (r != 0 && o + 1 <= avail && r - 1 <= avail - (o + 1)) { len = o + r; return TK_OK; }
    }
  }
  uint32_t kind; TokF f;
  len = toklen_at<W, BA, V>(img, x, single, avail, kind, f);
  idx = f.idx;
  return kind;
}

// W bytes of the image at p as a symbol
template <int W> __device__ __forceinline__ uint64_t img_sym(const uint8_t *img, uint32_t p)
{
  uint64_t v = 0;
#pragma unroll
  for (int i = 0; i < W; i++) v |= (uint64_t)img[p + i] << (8 * i);
  return v;
}

// a short literal part by one thread: bytes up to the next word of the image, words, bytes
__device__ __forceinline__ void thread_copy(uint8_t *tile, uint32_t d, uint32_t len, const uint8_t *img, const uint8_t *__restrict__ in, uint32_t c0, uint32_t rel)
{
  if (rel + len + 4 > DEC_CB + DEC_IMG_PAD) { for (uint32_t i = 0; i < len; i++) tile[d + i] = (uint8_t)lit_byte(img, in, c0, rel + i); return; }
  uint32_t i = 0;
  const uint32_t head = min(len, (4u - (d & 3u)) & 3u);
  for (; i < head; i++) tile[d + i] = img[rel + i];
  const uint32_t nw = (len - i) >> 2;
  if (nw)
  {
    const uint32_t r = rel + i, sh = (r & 3u) * 8u;
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(img) + (r >> 2);
    uint32_t *dw = reinterpret_cast<uint32_t *>(tile + d + i);
    uint32_t lo = sw[0];
    for (uint32_t k = 0; k < nw; k++) { const uint32_t hi = sw[k + 1]; dw[k] = __funnelshift_r(lo, hi, sh); lo = hi; }
    i += nw << 2;
  }
  for (; i < len; i++) tile[d + i] = img[rel + i];
}
// a short run part by one thread
template <int W> __device__ __forceinline__ void thread_fill(uint8_t *tile, uint32_t d, uint32_t len, uint64_t sym, uint32_t ph)
{
  if constexpr (W == 1)
  {
    const uint32_t b = (uint32_t)sym & 0xFFu, w = b * 0x01010101u;
    uint32_t i = 0;
    const uint32_t head = min(len, (4u - (d & 3u)) & 3u);
    for (; i < head; i++) tile[d + i] = (uint8_t)b;
    const uint32_t nw = (len - i) >> 2;
    uint32_t *dw = reinterpret_cast<uint32_t *>(tile + d + i);
    for (uint32_t k = 0; k < nw; k++) dw[k] = w;
    i += nw << 2;
    for (; i < len; i++) tile[d + i] = (uint8_t)b;
  }
  else
  {
    for (uint32_t i = 0; i < len; i++) { tile[d + i] = (uint8_t)run_byte<W>(sym, ph); ph = (ph + 1 == (uint32_t)W) ? 0u : ph + 1; }
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
  uint16_t (*const subRows)[DEC_WIN] = reinterpret_cast<uint16_t (*)[DEC_WIN]>(S.tile);     // the sub-chunk rows live in the tile buffer until the expansion
  constexpr uint32_t SUBROW_BYTES = DEC_NSUB * DEC_WIN * 2;
  static_assert(SUBROW_BYTES <= DEC_TILE, "sub-chunk rows fit the tile buffer");
  const DecScalars hs = *D.sc;
  DecCounters &cnt = *D.cnt;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  if (hs.status != ST_OK)
  { // header check of K1 failed: uniform over the grid
    if (blockIdx.x == 0 && t == 0) { D.dResult[0] = 0; D.dResult[1] = hs.status; for (int i = 2; i < 8; i++) D.dResult[i] = 0; }
    return;
  }
  const uint32_t clen = hs.clen, n = hs.n;
  const uint32_t nChunks = (clen + DEC_CB - 1) / DEC_CB;
  const uint32_t nLive = cnt.nLive;                                   // (K1's resolver)
  const bool single = hs.single != 0;
  const uint8_t *__restrict__ in = D.in;
  uint8_t *__restrict__ out = D.out;
  Agg *aggBuf = reinterpret_cast<Agg *>(D.aggBuf), *incBuf = reinterpret_cast<Agg *>(D.incBuf);
  if (t == 0) { mbar_init(&S.mbar[0], 1); mbar_init(&S.mbar[1], 1); }
  uint32_t phase0 = 0, phase1 = 0;
  __syncthreads();
#if defined(HSRLE_STAGE_MARKS)   // (hang diagnosis: build with -DHSRLE_STAGE_MARKS; every mark is a system-scope fence)
#define HSRLE_DBG(code) do { if (D.dbg && t == 0) { *((volatile uint32_t *)D.dbg + blockIdx.x) = (uint32_t)(code); __threadfence_system(); } } while (0)
#else
#define HSRLE_DBG(code) do { } while (0)
#endif
  // phase timers (build with -DHSRLE_PHASE_TIMERS, run with HSRLE_DEBUG): thread 0 sums the clock ticks between phase marks;
  // dbg[3000 + phase] gets the totals (in 64-tick units).  Not in production builds: the counters cost registers.
#if defined(HSRLE_PHASE_TIMERS)
  unsigned long long tPh[12] = { 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0 };
  long long tLast = D.dbg ? clock64() : 0;
#define HSRLE_PH(ph) do { if (D.dbg && t == 0) { const long long now_ = clock64(); tPh[ph] += (unsigned long long)(now_ - tLast); tLast = now_; } } while (0)
#else
#define HSRLE_PH(ph) do { } while (0)
#endif

  // block-wide exclusive scan of (output bytes, last explicit symbol) over the threads; totals returned in (totOut, totSym)
  auto block_scan = [&](unsigned long long myOut, uint32_t mySym, unsigned long long &exOut, uint32_t &exSym, unsigned long long &totOut, uint32_t &totSym)
  {
    unsigned long long io = myOut; uint32_t is = mySym;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1)
    {
      const unsigned long long o = __shfl_up_sync(0xFFFFFFFFu, io, d); const uint32_t q = __shfl_up_sync(0xFFFFFFFFu, is, d);
      if (lane >= d) { io += o; is = max(is, q); }
    }
    if (lane == 31) { S.warpSum[warp] = io; S.warpSymI[warp] = is; }
    unsigned long long eo = __shfl_up_sync(0xFFFFFFFFu, io, 1); uint32_t es = __shfl_up_sync(0xFFFFFFFFu, is, 1);
    if (lane == 0) { eo = 0; es = 0; }
    __syncthreads();
    unsigned long long po = 0, to = 0; uint32_t ps = 0, tsy = 0;
#pragma unroll
    for (int w = 0; w < DX_T / 32; w++) { const unsigned long long x = S.warpSum[w]; const uint32_t y = S.warpSymI[w]; if (w < warp) { po += x; ps = max(ps, y); } to += x; tsy = max(tsy, y); }
    __syncthreads();
    exOut = po + eo; exSym = max(ps, es); totOut = to; totSym = tsy;
  };

  for (;;)
  {
    __syncthreads();
    if (t == 0) { S.ticket = atomicAdd(&cnt.ticket, 1u); }
    __syncthreads();
    const uint32_t slot = S.ticket;                                   // position in the list of live chunks (look-back order)
    if (slot >= nLive) break;
    const uint32_t c = __ldcg(D.liveList + slot);
    const uint32_t c0 = c * DEC_CB;
    const uint32_t availSC = clen - c0;
    HSRLE_DBG(0x1000000u | c);
    HSRLE_PH(0);
    // ---- the chunk's first true token start (K1's resolver)
    if (t == 0) S.entry = __ldcg(D.chunkEntry + c);
    if (t < DEC_NSUB) { S.subEntry[t] = 0xFFFFFFFFu; S.subEnd[t] = 0; S.subCnt[t] = 0; }
    __syncthreads();
    const uint32_t entry = S.entry;
    const bool hasTok = entry < POS_SPECIAL && entry < clen;
    // ---- image + sub-chunk rows by bulk copies (a chunk no token starts in -- it lies inside a long literal -- needs neither)
    if (hasTok)
    {
      if (t == 0)
      {
        fence_async_smem();
        const uint32_t bytes = min(DEC_CB + DEC_IMG_PAD, (availSC + 15u) & ~15u);
        mbar_expect_tx(&S.mbar[0], bytes + SUBROW_BYTES);
        bulk_load(S.img, in + c0, bytes, &S.mbar[0]);
        bulk_load(S.tile, D.subMap + (size_t)c * DEC_NSUB * DEC_WIN, SUBROW_BYTES, &S.mbar[0]);
      }
      if (!mbar_wait(&S.mbar[0], phase0)) cnt.emitBad = 0x200;
      phase0 ^= 1u;
    }
    HSRLE_DBG(0x3000000u | c);
    HSRLE_PH(1);
    // ---- sub-chunk entries: hops through the sub-chunk rows (a landing outside a window is walked token by token)
    if (t == 0 && hasTok)
    {
      uint32_t x = entry - c0;
      while (x < DEC_CB && c0 + x < clen)
      {
        const uint32_t s = x / DEC_SB, os = x - s * DEC_SB;
        if (S.subEntry[s] == 0xFFFFFFFFu) S.subEntry[s] = x;
        uint32_t code;
        if (os < DEC_WIN) code = subRows[s][os];
        else
        {
          uint32_t len, idx;
          const uint32_t kind = tok_walk_len<W, BA, V>(S.img, x, single, availSC - x, len, idx);
          code = kind == TK_OK ? min(x + len, 0x7FFFu) : EX_BAD;   // (ends are handled by the walkers)
        }
        if (code >= EX_FAR) break;                                  // the chain ends or leaves the chunk far
        x = code;
      }
    }
    __syncthreads();
    HSRLE_DBG(0x4000000u | c);
    HSRLE_PH(2);
    // ---- the walkers: one lane per sub-chunk (two per warp), token starts only; LUT codecs also track where every token's
    //      symbol comes from (table slot at the start of the sub-chunk / explicit symbol) and the sub-chunk's table transform
    if ((lane & 15) == 0)
    {
      const int s = warp * 2 + (lane >> 4);
      uint32_t x = S.subEntry[s], k = 0, endKind = 0;
      LutXf xf; if (K) lutxf_identity(xf);
      if (x != 0xFFFFFFFFu)
      {
        const uint32_t subEnd = (uint32_t)(s + 1) * DEC_SB;
        while (x < subEnd)
        {
          if (c0 + x >= clen) { endKind = 2; break; }
          uint32_t len, idx;
          const uint32_t kind = tok_walk_len<W, BA, V>(S.img, x, single, availSC - x, len, idx);
          if (kind == TK_BAD) { endKind = 2; break; }
          S.pos[s][k] = (uint16_t)x;
          if (K)
          {
            lutxf_touch(xf, K, (int)idx, 8u + x + 2u);              // explicit symbols sit right behind the two head bytes (chunk-relative + 8 here)
            S.ref[K ? s : 0][K ? k : 0] = (uint16_t)xf.e[0];
          }
          k++;
          if (kind == TK_END) { endKind = 1; break; }
          x = (len >= 0x10000u) ? 0x10000u : x + len;
        }
      }
      S.subCnt[s] = k; S.subEnd[s] = endKind;
      if (K) S.subXf[K ? s : 0] = xf;
    }
    __syncthreads();
    HSRLE_PH(3);
    // ---- token counts -> exclusive prefix; end / error flags
    if (warp == 0)
    {
      uint32_t v = lane < DEC_NSUB ? S.subCnt[lane] : 0u;
      const uint32_t e = lane < DEC_NSUB ? S.subEnd[lane] : 0u;
      if (__any_sync(0xFFFFFFFFu, e == 2u) && lane == 0) cnt.emitBad = 1;
      if (__any_sync(0xFFFFFFFFu, e == 1u) && lane == 0) cnt.endSeen = 1;
      uint32_t inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) { const uint32_t o = __shfl_up_sync(0xFFFFFFFFu, inc, d); if (lane >= d) inc += o; }
      if (lane < DEC_NSUB) S.subCnt[lane] = inc - v;
      if (lane == DEC_NSUB - 1) S.subCnt[DEC_NSUB] = inc;
    }
    __syncthreads();
    const uint32_t ntokChunk = S.subCnt[DEC_NSUB];
    // groups of whole sub-chunks with at most DEC_NSLOT tokens each (a sub-chunk alone never has more than DX_POSCAP)
    auto group_end = [&](uint32_t s0) -> uint32_t
    {
      uint32_t s1 = s0 + 1;
      while (s1 < (uint32_t)DEC_NSUB && S.subCnt[s1 + 1] - S.subCnt[s0] <= DEC_NSLOT) s1++;
      return s1;
    };
    // parse the tokens of sub-chunks [s0, s1) in parallel (four consecutive tokens per thread), scan output bytes and the
    // last explicit symbol; with `keep` the records are stored
    auto parse_group = [&](uint32_t s0, uint32_t s1, bool keep, unsigned long long outBase, uint32_t symBase, unsigned long long &gOut, uint32_t &gSym)
    {
      const uint32_t tok0 = S.subCnt[s0], nt = S.subCnt[s1] - tok0;
      uint32_t lit[4], run[4], src[4], sy[4], sb[4];
      unsigned long long myOut = 0; uint32_t mySym = 0;
#pragma unroll
      for (int j = 0; j < 4; j++)
      {
        const uint32_t i = (uint32_t)t * 4 + j;
        lit[j] = 0; run[j] = 0; src[j] = 0; sy[j] = 0; sb[j] = 0;
        if (i < nt)
        {
          uint32_t s = s0;
          while (S.subCnt[s + 1] <= tok0 + i) s++;                 // (sub-chunks without tokens are stepped over)
          const uint32_t k = tok0 + i - S.subCnt[s];
          const uint32_t x = S.pos[s][k];
          uint32_t kind; TokF f;
          const uint32_t len = toklen_at<W, BA, V>(S.img, x, single, availSC - x, kind, f);
          uint32_t runB = tok_run_bytes<W, BA, V>(f.cnt, single), litB = len - f.hdr;
          if (kind == TK_END) { runB = 0; if (len < f.hdr) litB = 0; }
          lit[j] = litB; run[j] = runB; src[j] = x + f.hdr; sb[j] = s;
          if (K) sy[j] = S.ref[K ? s : 0][K ? k : 0];
          else sy[j] = f.symOff != 0xFFu ? x + f.symOff + 1u : 0u;   // + 1: 0 means "no explicit symbol"
          myOut += (unsigned long long)litB + runB;
          if (!K) mySym = max(mySym, sy[j]);
        }
      }
      unsigned long long exOut; uint32_t exSym;
      block_scan(myOut, mySym, exOut, exSym, gOut, gSym);
      if (keep)
      {
        unsigned long long o = outBase + exOut; uint32_t ls = max(symBase, exSym);
#pragma unroll
        for (int j = 0; j < 4; j++)
        {
          const uint32_t i = (uint32_t)t * 4 + j;
          if (i < nt)
          {
            if (!K) ls = max(ls, sy[j]);
            S.rOut[i] = (uint32_t)min(o, (unsigned long long)0xFFFFFFFFu); S.rLit[i] = lit[j]; S.rRun[i] = run[j]; S.rSrc[i] = (uint16_t)src[j];
            S.rSymI[i] = K ? (uint16_t)sy[j] : (uint16_t)(ls ? ls - 1u : 0xFFFFu); S.rSub[i] = (uint8_t)sb[j];
            o += (unsigned long long)lit[j] + run[j];
          }
        }
        if (t == 0) S.rOut[nt] = (uint32_t)min(outBase + gOut, (unsigned long long)0xFFFFFFFFu);      // sentinel: where the group ends
      }
    };
    // ---- chunk totals (single group: its records are kept)
    const uint32_t gEnd0 = group_end(0);
    const bool oneGroup = gEnd0 == (uint32_t)DEC_NSUB;
    unsigned long long chunkOutLen = 0; uint32_t chunkSym = 0;
    for (uint32_t s0 = 0; s0 < (uint32_t)DEC_NSUB;)
    {
      const uint32_t s1 = group_end(s0);
      unsigned long long gOut; uint32_t gSym;
      parse_group(s0, s1, oneGroup, 0ull, 0u, gOut, gSym);
      chunkOutLen += gOut; chunkSym = max(chunkSym, gSym);
      s0 = s1;
    }
    HSRLE_DBG(0x5000000u | c);
    HSRLE_PH(4);
    // ---- publish the chunk's aggregate; decoupled look-back for the exclusive prefix, 256 chunks per step (one per thread; the first
    //      wave of CTAs starts together, so the look-back is deep there); symbol state
    Agg tot = decagg_identity<K>();
    Agg subEx = decagg_identity<K>();                                 // K > 0 (warp 0): table transform of the sub-chunks before lane's
    if (warp == 0)
    {
      tot.out = chunkOutLen; tot.ntok = ntokChunk; tot.symPos = chunkSym ? c0 + chunkSym - 1u : 0u;
      if (K)
      {
        Agg mine = decagg_identity<K>();
        if (lane < DEC_NSUB)
        { // the walkers' entries are chunk-relative + 8: make them stream positions
          mine.xf = S.subXf[K ? lane : 0];
#pragma unroll
          for (int i = 0; i < 7; i++) if (i < K && mine.xf.e[i] >= 8u) mine.xf.e[i] += c0 - 8u;
        }
        Agg inc = mine;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1)
        {
          const Agg o = decagg_shfl_up<K>(inc, d);
          if (lane >= d) inc = decagg_combine<K>(o, inc);
        }
        subEx = decagg_shfl_up<K>(inc, 1);
        if (lane == 0) subEx = decagg_identity<K>();
#pragma unroll
        for (int i = 0; i < 7; i++) if (i < K) tot.xf.e[i] = __shfl_sync(0xFFFFFFFFu, inc.xf.e[i], 31);
      }
      if (lane == 0) { decagg_store<K>(&aggBuf[slot], tot); st_release_u32(D.flagAgg + slot, 1u); }
    }
    Agg excl = decagg_identity<K>();                                  // (maintained by warp 0)
    for (int64_t base = (int64_t)slot - 1; base >= 0; base -= DX_T)
    {
      const int64_t p = base - t;
      uint32_t fl = 2u;                                               // threads before chunk 0 behave like an (identity) inclusive prefix
      Agg v = decagg_identity<K>();
      if (p >= 0)
      {
        uint32_t spin = 0;
        do { fl = ld_acquire_u32(D.flagAgg + p); } while (fl == 0u && ++spin < (1u << 24));
        if (fl == 0u) { cnt.emitBad = 0x400; fl = 2u; }              // (cannot happen: chunks are taken in order)
        v = decagg_load_cg<K>(fl == 2u ? &incBuf[p] : &aggBuf[p]);
      }
      const uint32_t incMask = __ballot_sync(0xFFFFFFFFu, fl == 2u);
      const int j = incMask ? (__ffs(incMask) - 1) : 32;             // nearest inclusive prefix in this warp's window
      if (lane > j) v = decagg_identity<K>();
      // ordered reduction: lane 0 ends with the combination of lanes 0 .. 31 (older chunks are the higher lanes)
#pragma unroll
      for (int d = 1; d < 32; d <<= 1)
      {
        const Agg o = decagg_shfl_down<K>(v, d);
        if (lane + d < 32) v = decagg_combine<K>(o, v);
      }
      if (lane == 0) { decagg_store<K>(&S.lbAgg[warp], v); S.lbAgg[warp].ntok = v.ntok; S.lbHit[warp] = incMask ? 1u : 0u; }
      __syncthreads();
      if (warp == 0)
      { // merge the warps' windows, nearest first, until one reached an inclusive prefix
        bool hit = false;
        Agg r = decagg_identity<K>();
        for (int w = 0; w < DX_T / 32 && !hit; w++)
        {
          Agg a = decagg_identity<K>();
          a.out = S.lbAgg[w].out; a.ntok = S.lbAgg[w].ntok; a.symPos = S.lbAgg[w].symPos;
          if (K) a.xf = S.lbAgg[w].xf;
          r = decagg_combine<K>(a, r);
          hit = S.lbHit[w] != 0;
        }
        excl = decagg_combine<K>(r, excl);
        if (lane == 0) S.flag = hit ? 1u : 0u;
      }
      __syncthreads();
      if (S.flag) break;
    }
    if (warp == 0)
    {
      const Agg incl = decagg_combine<K>(excl, tot);
      if (lane == 0)
      {
        decagg_store<K>(&incBuf[slot], incl); st_release_u32(D.flagAgg + slot, 2u);
        if (slot == nLive - 1) { cnt.outTotal = incl.out; cnt.nTok = incl.ntok; }
        decagg_store<K>(&S.bc, excl);
        if (!K) S.inSym = single ? (uint64_t)hs.singleSym : (excl.symPos ? load_sym(in + excl.symPos, W) : 0ull);
      }
      if (K && lane < DEC_NSUB)
      { // the table at the start of every sub-chunk
        const Agg before = decagg_combine<K>(excl, subEx);
        Lut l0; lut_init(l0, W);
        Lut l1; lutxf_apply(before.xf, K, W, in, l0, l1);
        S.subLut[K ? lane : 0] = l1;
      }
    }
    __syncthreads();
    HSRLE_DBG(0x6000000u | c);
    HSRLE_PH(5);
    const uint64_t chunkOut0 = S.bc.out;
    // run symbol of a token record
    auto rec_sym = [&](uint32_t r) -> uint64_t
    {
      const uint32_t v = S.rSymI[r];
      if (K) return v < 8u ? lut_get(S.subLut[K ? S.rSub[r] : 0], K, (int)v) : img_sym<W>(S.img, v - 8u);
      return v == 0xFFFFu ? S.inSym : img_sym<W>(S.img, v);
    };
    // ---- expansion, group by group
    unsigned long long outBase = 0; uint32_t symBase = 0;
    for (uint32_t s0 = 0; s0 < (uint32_t)DEC_NSUB && ntokChunk;)
    {
      const uint32_t s1 = group_end(s0);
      const uint32_t passN = S.subCnt[s1] - S.subCnt[s0];
      __syncthreads();
      if (!oneGroup)
      {
        unsigned long long gOut; uint32_t gSym;
        parse_group(s0, s1, true, outBase, symBase, gOut, gSym);
        outBase += gOut; symBase = max(symBase, gSym);
      }
      s0 = s1;
      if (t == 0) { S.nSkip = 0; }
      __syncthreads();
      if (passN == 0) continue;
      HSRLE_DBG(0x7000000u | c);
      // output range of the group, absolute (never beyond the declared size)
      const uint64_t pLo = min(chunkOut0 + S.rOut[0], (uint64_t)n);
      const uint64_t pHi = min(chunkOut0 + S.rOut[passN], (uint64_t)n);
      if (pHi <= pLo) continue;
      const uint64_t T0 = pLo & ~(uint64_t)15;
      const uint32_t nTiles = (uint32_t)((pHi - T0 + DEC_TILE - 1) / DEC_TILE);
      // -- token parts that cover many whole tiles go to the grid
      if (nTiles > DEC_HUGE_TILES)
      {
        for (uint32_t r = t; r < passN; r += DX_T)
        {
          const uint64_t a = chunkOut0 + S.rOut[r];
          const uint64_t m = a + S.rLit[r], e = m + S.rRun[r];
#pragma unroll 1
          for (int part = 0; part < 2; part++)
          {
            const uint64_t pa = max(part ? m : a, pLo), pb = min(part ? e : m, pHi);
            if (pb <= pa) continue;
            const uint64_t kA = (pa - T0 + DEC_TILE - 1) / DEC_TILE, kB = (pb - T0) / DEC_TILE;
            if (kB < kA + DEC_HUGE_TILES) continue;
            const uint32_t slot = atomicAdd(&S.nSkip, 1u);
            if (slot >= (uint32_t)DX_NSKIP) continue;
            const uint32_t opLen = (uint32_t)((kB - kA) * DEC_TILE), np = (opLen + DEC_BIG_PIECE - 1) / DEC_BIG_PIECE;
            const unsigned long long reg = atomicAdd(&cnt.bigReg, (1ull << 40) | (unsigned long long)np);
            const uint32_t idx = (uint32_t)(reg >> 40), base = (uint32_t)(reg & 0xFFFFFFFFFFull);
            if (idx >= D.bigCap || base + np > D.pieceCap) { cnt.emitBad = 0x800; S.skipLo[slot] = 0; S.skipHi[slot] = 0; continue; }   // (cannot happen: the caps are upper bounds)
            S.skipLo[slot] = (uint32_t)kA; S.skipHi[slot] = (uint32_t)kB;
            DecBigOp &op = D.bigList[idx];
            op.dst = (uint32_t)(T0 + kA * DEC_TILE); op.len = opLen; op.kind = (uint32_t)part; op.sym = part ? rec_sym(r) : 0ull;
            op.src = part ? (uint32_t)m : (uint32_t)(c0 + S.rSrc[r] + (T0 + kA * DEC_TILE - a));
            op.pieceBase = base;
            __threadfence();
            st_volatile_u32(&op.ready, 1u);
            for (uint32_t k = 0; k < np; k++) st_volatile_u32(D.pieceOp + base + k, idx + 1u);
          }
        }
        __syncthreads();
      }
      const uint32_t nSkip = min(S.nSkip, (uint32_t)DX_NSKIP);
      for (uint32_t tk = 0; tk < nTiles; tk++)
      {
        { // inside a range handed to the grid?
          uint32_t jump = 0;
          for (uint32_t i = 0; i < nSkip; i++) if (tk >= S.skipLo[i] && tk < S.skipHi[i]) jump = S.skipHi[i];
          if (jump) { tk = jump - 1; continue; }
        }
        const uint64_t tb = T0 + (uint64_t)tk * DEC_TILE;
        const uint64_t lo = max(tb, pLo), hi = min(tb + DEC_TILE, pHi);
        __syncthreads();                                            // the previous tile is flushed
        if (t == 0) S.nLong = 0;
        __syncthreads();
        HSRLE_PH(6);
        // -- every token of the group that overlaps the tile: short parts by the token's thread, long parts listed
        for (uint32_t r = t; r < passN; r += DX_T)
        {
          const uint64_t a = chunkOut0 + S.rOut[r];
          if (a >= hi) continue;
          const uint64_t m = a + S.rLit[r], e = m + S.rRun[r];
          if (e <= lo) continue;
          { // literal part
            const uint64_t pa = max(a, lo), pb = min(m, hi);
            if (pb > pa)
            {
              const uint32_t len = (uint32_t)(pb - pa), d = (uint32_t)(pa - tb), rel = S.rSrc[r] + (uint32_t)(pa - a);
              if (len <= DX_INLINE) thread_copy(S.tile, d, len, S.img, in, c0, rel);
              else { const uint32_t q = atomicAdd(&S.nLong, 1u); if (q < DX_LONGCAP) S.longList[q] = r * 2u; }
            }
          }
          { // run part
            const uint64_t pa = max(m, lo), pb = min(e, hi);
            if (pb > pa)
            {
              const uint32_t len = (uint32_t)(pb - pa), d = (uint32_t)(pa - tb);
              if (len <= DX_INLINE) thread_fill<W>(S.tile, d, len, rec_sym(r), (uint32_t)((pa - m) % (uint32_t)W));
              else { const uint32_t q = atomicAdd(&S.nLong, 1u); if (q < DX_LONGCAP) S.longList[q] = r * 2u + 1u; }
            }
          }
        }
        __syncthreads();
        HSRLE_PH(7);
        // -- long parts: one warp each
        {
          const uint32_t nl = min(S.nLong, (uint32_t)DX_LONGCAP);
          for (uint32_t q = warp; q < nl; q += DX_T / 32)
          {
            const uint32_t r = S.longList[q] >> 1, part = S.longList[q] & 1u;
            const uint64_t a = chunkOut0 + S.rOut[r];
            const uint64_t m = a + S.rLit[r], e = m + S.rRun[r];
            if (!part)
            {
              const uint64_t pa = max(a, lo), pb = min(m, hi);
              warp_copy(S.tile, (uint32_t)(pa - tb), (uint32_t)(pb - pa), S.img, in, c0, S.rSrc[r] + (uint32_t)(pa - a), lane);
            }
            else
            {
              const uint64_t pa = max(m, lo), pb = min(e, hi);
              warp_fill<W>(S.tile, (uint32_t)(pa - tb), (uint32_t)(pb - pa), rec_sym(r), (uint32_t)((pa - m) % (uint32_t)W), lane);
            }
          }
        }
        __syncthreads();
        HSRLE_PH(8);
        // -- flush: whole vectors with 16-byte stores, the ragged ends (shared with the neighbouring chunks) byte-wise
        {
          const uint32_t b0 = (uint32_t)(lo - tb), b1 = (uint32_t)(hi - tb);
          uint8_t *g = out + tb;
          const uint32_t v0 = (b0 + 15u) >> 4, v1 = b1 >> 4;
          for (uint32_t v = v0 + t; v < v1; v += DX_T) reinterpret_cast<uint4 *>(g)[v] = reinterpret_cast<const uint4 *>(S.tile)[v];
          if (v1 >= v0)
          {
            if ((uint32_t)t < (v0 << 4) - b0) g[b0 + t] = S.tile[b0 + t];                 // head bytes [b0, 16 v0)
            if ((uint32_t)t < b1 - (v1 << 4)) g[(v1 << 4) + t] = S.tile[(v1 << 4) + t];    // tail bytes [16 v1, b1)
          }
          else if (b0 + (uint32_t)t < b1) g[b0 + t] = S.tile[b0 + t];                      // both ends inside one vector
        }
      }
    }
    HSRLE_DBG(0x9000000u | c);
    HSRLE_PH(9);
    // ---- chunk done; the last one settles the status and the result
    __syncthreads();
    if (t == 0)
    {
      __threadfence();
      if (atomicAdd(&cnt.chunksDone, 1u) == nLive - 1)
      {
        __threadfence();
        uint32_t status = ST_OK;
        const uint32_t bad = ld_volatile_u32(&cnt.emitBad) | ld_volatile_u32(&cnt.chainBad), end = ld_volatile_u32(&cnt.endSeen);
        const unsigned long long tot = ld_volatile_u64(&cnt.outTotal);
        if (bad || !end || tot != (unsigned long long)n) status = ST_BADSTREAM;
        D.dResult[0] = status == ST_OK ? n : 0; D.dResult[1] = status; D.dResult[2] = ld_volatile_u32(&cnt.nTok); D.dResult[3] = nChunks;
        D.dResult[4] = clen; D.dResult[5] = hs.single; D.dResult[6] = (uint32_t)(ld_volatile_u64(&cnt.bigReg) >> 40); D.dResult[7] = bad;
      }
    }
  }

  HSRLE_PH(10);
#if defined(HSRLE_PHASE_TIMERS)
  if (D.dbg && t == 0) { for (int i = 0; i < 12; i++) atomicAdd(D.dbg + 3000 + i, (uint32_t)(tPh[i] >> 6)); }
#endif
  // ---- grid-wide operations: every CTA that is out of chunks takes pieces (DEC_BIG_PIECE bytes each) in the global piece order until
  //      all chunks are done and every registered piece is taken.  One atomic per piece; the operation a piece belongs to is looked up
  //      in pieceOp (written at registration).  One thread polls (the others wait at the barrier and cost no issue slots).
  for (uint32_t guard = 0; guard < (1u << 24); guard++)
  {
    __syncthreads();
    if (t == 0)
    {
      const uint32_t T = atomicAdd(&cnt.bigTicket, 1u);
      uint32_t found = 0, opi = 0;
      for (uint32_t spin = 0; spin < (1u << 22); spin++)
      { // "all done" is sampled before the registration count: whatever was registered before the last chunk finished is seen
        const uint32_t done = (ld_volatile_u32(&cnt.chunksDone) >= nLive) ? 1u : 0u;
        __threadfence();
        const unsigned long long reg = ld_volatile_u64(&cnt.bigReg);
        if (T < (uint32_t)(reg & 0xFFFFFFFFFFull)) { found = 1; break; }
        if (done) break;
        __nanosleep(500);
      }
      if (found)
      {
        uint32_t spin = 0;
        while ((opi = ld_volatile_u32(D.pieceOp + T)) == 0u && ++spin < (1u << 24)) { }
        if (opi) { while (ld_volatile_u32(&D.bigList[opi - 1].ready) == 0u && ++spin < (1u << 24)) { } }
        if (!opi || spin >= (1u << 24)) { cnt.emitBad = 0x900; found = 0; }
      }
      S.flag = found; S.entry = opi; S.bcast = T;
    }
    __syncthreads();
    if (!S.flag) break;
    __threadfence();
    const DecBigOp &op = D.bigList[S.entry - 1];
    DecBigOp o;
    o.sym = __ldcg(&op.sym); o.dst = __ldcg(&op.dst); o.len = __ldcg(&op.len); o.src = __ldcg(&op.src); o.kind = __ldcg(&op.kind);
    const uint32_t k = S.bcast - __ldcg(&op.pieceBase);
    dec_big_piece<W, K>(D, S, o, k * DEC_BIG_PIECE, min(DEC_BIG_PIECE, o.len - k * DEC_BIG_PIECE), phase0, phase1);
  }
}

} // namespace hsrle
