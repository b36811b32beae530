// hsrle_kernels.cu -- sm_100a kernels + C ABI of the B200 extreme-RLE codec.
//
// Encoder pipeline (device resident, stream ordered, no host sync):
//   enc_count/enc_write : coalesced 16-byte-per-lane scan of the match mask M[p]=(in[p]==in[p-W]),
//                         run boundaries by bit tricks, compaction into (start,end) lists
//   enc_auto_*          : the reference's per-run emit rules as a speculative automaton over chunks of
//                         runs; incoming states by scan, verified, re-run until a fixed point
//   enc_sizes           : scan of token byte sizes -> stream offsets, header, terminator
//   enc_emit/enc_copy_* : token headers and literal scatter
// Decoder pipeline:
//   dec_map             : per 4 KiB chunk, "where does a token chain starting at byte p leave the chunk"
//                         for every p (pointer doubling in shared memory)
//   dec_up/top/down     : hierarchical composition of those maps -> true entry offset of every chunk
//   dec_walk/scan       : token walk per chunk, scans for output offsets and symbol / LUT state
//   dec_expand          : one 16-byte output vector per lane (literal gather / period-W run fill)
//
// Reference entry points this file replaces: src/rle.h:100-394 (see include/hsrle_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>
#include <map>

#include "../../include/hsrle_b200.h"
#include "hsrle_stages.cuh"

namespace hsrle {

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

// optional per-kernel CUDA-event timing (bench.py's roofline leg); off by default
struct TimedLaunch { const char *name; cudaEvent_t a, b; };
static bool g_timing = false;
static std::vector<TimedLaunch> g_timed;

#define HSRLE_LAUNCH(kern, grid, block, smem, stream, ...)                 \
  do {                                                                     \
    TimedLaunch tl_{ #kern, nullptr, nullptr };                            \
    if (g_timing) { cudaEventCreate(&tl_.a); cudaEventCreate(&tl_.b); cudaEventRecord(tl_.a, (stream)); } \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
    if (g_timing) { cudaEventRecord(tl_.b, (stream)); g_timed.push_back(tl_); } \
    g_launches.fetch_add(1, std::memory_order_relaxed);                    \
  } while (0)

static bool cuda_ok(cudaError_t e, const char *what)
{
  if (e == cudaSuccess) return true;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

constexpr int GS_BLOCK = 128;            // threads of grid-stride "one item per thread" kernels
constexpr int GS_GRID = 148 * 8;         // 148 SMs x 8 resident CTAs
constexpr int SCAN_T = 512;              // threads of the single-CTA scan kernels

// ================================================================================================
// block helpers
__device__ __forceinline__ uint64_t warp_incl_scan_u64(uint64_t v)
{
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1)
  {
    const uint64_t o = __shfl_up_sync(0xFFFFFFFFu, v, d);
    if (lane >= d) v += o;
  }
  return v;
}
// exclusive scan of one u64 per thread across the block; total returned in `total`
__device__ __forceinline__ uint64_t block_excl_scan_u64(uint64_t v, uint64_t *smemWarp /* >= 33 */, uint64_t &total)
{
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  const uint64_t inc = warp_incl_scan_u64(v);
  if (lane == 31) smemWarp[warp] = inc;
  __syncthreads();
  if (warp == 0)
  {
    uint64_t w = lane < nw ? smemWarp[lane] : 0;
    const uint64_t wi = warp_incl_scan_u64(w);
    smemWarp[lane] = wi - w;
    if (lane == 31) smemWarp[32] = wi;
  }
  __syncthreads();
  const uint64_t r = smemWarp[warp] + inc - v;
  total = smemWarp[32];
  __syncthreads();
  return r;
}

// ================================================================================================
// E1: candidate scan
__device__ __forceinline__ void enc_vec_masks(const EncBufs &B, uint32_t v, uint32_t &starts, uint32_t &ends)
{
  const uint4 *in16 = reinterpret_cast<const uint4 *>(B.in);
  const uint32_t lastVec = (B.n - 1) >> 4;   // last vector holding input bytes (n > 0)
  uint4 a = make_uint4(0, 0, 0, 0), b = a, c = a;
  if (v >= 1 && v - 1 <= lastVec) a = __ldg(in16 + (v - 1));
  if (v <= lastVec) b = __ldg(in16 + v);
  if (v + 1 <= lastVec) c = __ldg(in16 + (v + 1));
  const uint32_t w[12] = { a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w };
  mark_from_words(B.sp, w, B.n, (uint64_t)v * ENC_VEC, starts, ends);
}

__global__ void __launch_bounds__(ENC_TILE_VECS) k_enc_count(const EncBufs B)
{
  __shared__ uint64_t sw[34];
  const uint32_t v = blockIdx.x * ENC_TILE_VECS + threadIdx.x;
  uint32_t s = 0, e = 0;
  if (v < B.nVec) enc_vec_masks(B, v, s, e);
  uint64_t total;
  block_excl_scan_u64((uint64_t)__popc(s) | ((uint64_t)__popc(e) << 32), sw, total);
  if (threadIdx.x == 0) { B.tileS[blockIdx.x] = (uint32_t)total; B.tileE[blockIdx.x] = (uint32_t)(total >> 32); }
}

__global__ void __launch_bounds__(ENC_TILE_VECS) k_enc_write(const EncBufs B)
{
  __shared__ uint64_t sw[34];
  const uint32_t v = blockIdx.x * ENC_TILE_VECS + threadIdx.x;
  uint32_t s = 0, e = 0;
  if (v < B.nVec) enc_vec_masks(B, v, s, e);
  uint64_t total;
  const uint64_t ex = block_excl_scan_u64((uint64_t)__popc(s) | ((uint64_t)__popc(e) << 32), sw, total);
  uint32_t ps = B.tileS[blockIdx.x] + (uint32_t)ex, pe = B.tileE[blockIdx.x] + (uint32_t)(ex >> 32);
  const uint32_t p0 = v * ENC_VEC;
  while (s) { const int i = __ffs(s) - 1; s &= s - 1; B.runA[ps++] = p0 + i; }
  while (e) { const int i = __ffs(e) - 1; e &= e - 1; B.runB[pe++] = p0 + i; }
}

// single CTA: exclusive scan of the per-tile (starts, ends) counts; publishes nRuns / nChunks
__global__ void __launch_bounds__(1024) k_enc_scan_tiles(const EncBufs B)
{
  __shared__ uint64_t sw[34];
  const uint32_t nT = B.nTiles;
  const uint32_t per = (nT + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = min(nT, threadIdx.x * per), hi = min(nT, lo + per);
  uint64_t sum = 0;
  for (uint32_t t = lo; t < hi; t++) sum += (uint64_t)B.tileS[t] | ((uint64_t)B.tileE[t] << 32);
  uint64_t total;
  uint64_t run = block_excl_scan_u64(sum, sw, total);
  for (uint32_t t = lo; t < hi; t++)
  {
    const uint64_t c = (uint64_t)B.tileS[t] | ((uint64_t)B.tileE[t] << 32);
    B.tileS[t] = (uint32_t)run; B.tileE[t] = (uint32_t)(run >> 32);
    run += c;
  }
  if (threadIdx.x == 0)
  {
    const uint32_t nS = (uint32_t)total, nE = (uint32_t)(total >> 32);
    EncScalars &sc = *B.sc;
    if (nS != nE || nS > B.maxRuns) { sc.status = ST_BADARG; sc.nRuns = 0; sc.nChunks = 0; }
    else { sc.nRuns = nS; sc.nChunks = (nS + ENC_CH - 1) / ENC_CH; }
  }
}

// ================================================================================================
// E2: automaton
__global__ void __launch_bounds__(GS_BLOCK) k_enc_auto_init(const EncBufs B)
{
  const uint32_t nC = B.sc->nChunks;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nC; c += gridDim.x * blockDim.x) enc_stage_auto_init(B, c);
}

__global__ void __launch_bounds__(GS_BLOCK) k_enc_rerun(const EncBufs B)
{
  if (B.sc->nDirty == 0) return;
  const uint32_t nC = B.sc->nChunks;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nC; c += gridDim.x * blockDim.x) enc_stage_rerun(B, c);
}

struct ScanElem { ChunkSum sum; LutAgg agg; };

// single CTA: exclusive scan of the chunk summaries -> exact incoming state of every chunk under the
// current decisions; compare with what each chunk assumed; mark dirty.
__global__ void __launch_bounds__(SCAN_T) k_enc_scan_check(const EncBufs B, int round)
{
  extern __shared__ __align__(16) unsigned char smemRaw[];
  EncScalars &sc = *B.sc;
  if (round > 0 && sc.nDirty == 0) return;           // already at the fixed point
  ScanElem *el = reinterpret_cast<ScanElem *>(smemRaw);
  ScanElem *blk = el + SCAN_T;
  __shared__ uint32_t sDirty, sFirst;
  const int K = B.sp.K;
  const uint32_t nC = sc.nChunks;
  const uint32_t per = (nC + SCAN_T - 1) / SCAN_T;
  const uint32_t lo = min(nC, threadIdx.x * per), hi = min(nC, lo + per);
  if (threadIdx.x == 0) { sDirty = 0; sFirst = 0xFFFFFFFFu; }
  // phase A: composite of my slice
  ScanElem me; me.sum.flags = 0; me.sum.last = 0; me.sum.cursor = 0; me.sum.lastSym = 0; me.agg.m = 0;
  for (uint32_t c = lo; c < hi; c++)
  {
    me.sum = chunksum_combine(me.sum, B.cSum[c]);
    if (K) me.agg = lutagg_combine(me.agg, B.lutAgg[c], K);
  }
  el[threadIdx.x] = me;
  __syncthreads();
  // phase B: two-level exclusive scan (32 x 16)
  constexpr int NB = SCAN_T / 32;
  if (threadIdx.x < NB)
  {
    ScanElem acc; acc.sum.flags = 0; acc.sum.last = 0; acc.sum.cursor = 0; acc.sum.lastSym = 0; acc.agg.m = 0;
    for (int i = 0; i < 32; i++)
    {
      const ScanElem cur = el[threadIdx.x * 32 + i];
      el[threadIdx.x * 32 + i] = acc;
      acc.sum = chunksum_combine(acc.sum, cur.sum);
      if (K) acc.agg = lutagg_combine(acc.agg, cur.agg, K);
    }
    blk[threadIdx.x] = acc;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    ScanElem acc; acc.sum.flags = 0; acc.sum.last = 0; acc.sum.cursor = 0; acc.sum.lastSym = 0; acc.agg.m = 0;
    for (int i = 0; i < NB; i++)
    {
      const ScanElem cur = blk[i];
      blk[i] = acc;
      acc.sum = chunksum_combine(acc.sum, cur.sum);
      if (K) acc.agg = lutagg_combine(acc.agg, cur.agg, K);
    }
  }
  __syncthreads();
  // phase C: walk my slice with the exact running state
  {
    const ScanElem b = blk[threadIdx.x / 32], w = el[threadIdx.x];
    AutoState st = enc_initial_state();
    chunksum_apply(st, b.sum); chunksum_apply(st, w.sum);
    Lut lut; lut_init(lut, B.sp.W);
    if (K) { lut_apply(lut, K, b.agg); lut_apply(lut, K, w.agg); }
    uint32_t first = 0xFFFFFFFFu;
    const uint32_t nd = enc_scan_check_range(B, lo, hi, st, lut, first);
    if (nd) { atomicAdd(&sDirty, nd); atomicMin(&sFirst, first); }
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    sc.nDirty = sDirty; sc.firstDirty = sFirst;
    if (sDirty) sc.rounds++;
  }
}

__global__ void k_enc_serial(const EncBufs B)
{
  if (B.sc->nDirty == 0) return;
  enc_stage_serial(B, B.sc->firstDirty);
  B.sc->nDirty = 0;
}

// single CTA: exclusive scan of per-chunk token bytes / token counts, then header + terminator
__global__ void __launch_bounds__(1024) k_enc_sizes(const EncBufs B, uint32_t *dResult)
{
  __shared__ uint64_t sw[34];
  EncScalars &sc = *B.sc;
  const uint32_t nC = sc.nChunks;
  const uint32_t per = (nC + blockDim.x - 1) / blockDim.x;
  const uint32_t lo = min(nC, threadIdx.x * per), hi = min(nC, lo + per);
  uint64_t sb = 0, st = 0;
  for (uint32_t c = lo; c < hi; c++) { sb += B.cBytes[c]; st += B.cTok[c]; }
  uint64_t totB, totT;
  uint64_t rb = block_excl_scan_u64(sb, sw, totB);
  uint64_t rt = block_excl_scan_u64(st, sw, totT);
  for (uint32_t c = lo; c < hi; c++)
  {
    const uint64_t b = B.cBytes[c]; const uint32_t t = B.cTok[c];
    B.cBytes[c] = rb; B.cTok[c] = (uint32_t)rt;
    rb += b; rt += t;
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    sc.tokBytes = totB; sc.nTok = (uint32_t)totT;
    if (sc.status == ST_OK) enc_stage_finish(B); else sc.total = 0;
    dResult[0] = sc.status == ST_OK ? sc.total : 0; dResult[1] = sc.status; dResult[2] = sc.nRuns; dResult[3] = sc.nChunks;
    dResult[4] = sc.rounds; dResult[5] = sc.serialChunks; dResult[6] = sc.nTok; dResult[7] = sc.nBig;
  }
}

__global__ void __launch_bounds__(GS_BLOCK) k_enc_emit(const EncBufs B)
{
  if (B.sc->status != ST_OK) return;
  const uint32_t nC = B.sc->nChunks;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nC; c += gridDim.x * blockDim.x) enc_stage_emit(B, c);
}

// literal scatter: one warp per token literal (short ones), grid-wide pieces for the long ones
__global__ void __launch_bounds__(256) k_enc_copy_small(const EncBufs B)
{
  if (B.sc->status != ST_OK) return;
  const uint32_t nD = B.sc->nTok + 1;
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t nWarps = gridDim.x * (blockDim.x >> 5);
  for (uint32_t d = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); d < nD; d += nWarps)
  {
    const CopyDesc cd = B.copies[d];
    if (cd.len >= BIG_COPY) continue;
    const uint8_t *src = B.in + cd.src; uint8_t *dst = B.out + cd.dst;
    for (uint32_t i = lane; i < cd.len; i += 32) dst[i] = src[i];
  }
}

__device__ __forceinline__ void copy_bytes_block(uint8_t *dst, const uint8_t *src, uint32_t len)
{
  // dst-aligned 16-byte stores fed by unaligned 4-byte source words (funnel shift); byte head/tail
  const uint32_t head = min(len, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
  for (uint32_t i = threadIdx.x; i < head; i += blockDim.x) dst[i] = src[i];
  dst += head; src += head; len -= head;
  const uint32_t nv = len >> 4;
  const uint32_t sh = ((uintptr_t)src & 3) * 8;
  const uint32_t *sw = reinterpret_cast<const uint32_t *>((uintptr_t)src & ~(uintptr_t)3);
  for (uint32_t v = threadIdx.x; v < nv; v += blockDim.x)
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
  const uint32_t done = nv << 4;
  for (uint32_t i = done + threadIdx.x; i < len; i += blockDim.x) dst[i] = src[i];
}

constexpr uint32_t BIG_PIECE = 16384;
__global__ void __launch_bounds__(256) k_enc_copy_big(const EncBufs B)
{
  if (B.sc->status != ST_OK) return;
  const uint32_t nBig = B.sc->nBig;
  for (uint32_t i = 0; i < nBig; i++)
  {
    const CopyDesc cd = B.copies[B.bigList[i]];
    // keep the funnel-shift reads inside the literal: the last piece ends 16 bytes early, tail by bytes
    const uint32_t nPieces = (cd.len + BIG_PIECE - 1) / BIG_PIECE;
    for (uint32_t pc = blockIdx.x; pc < nPieces; pc += gridDim.x)
    {
      const uint32_t off = pc * BIG_PIECE;
      const uint32_t len = min(BIG_PIECE, cd.len - off);
      const bool lastPiece = pc + 1 == nPieces;
      if (!lastPiece) copy_bytes_block(B.out + cd.dst + off, B.in + cd.src + off, len);
      else
      {
        const uint32_t safe = len > 32 ? len - 32 : 0;
        copy_bytes_block(B.out + cd.dst + off, B.in + cd.src + off, safe);
        for (uint32_t k = safe + threadIdx.x; k < len; k += blockDim.x) B.out[cd.dst + off + k] = B.in[cd.src + off + k];
      }
    }
  }
}

// ================================================================================================
// DECODER kernels
__global__ void k_dec_init(const DecBufs D) { dec_stage_init(D); }

__global__ void __launch_bounds__(256) k_dec_map(const DecBufs D)
{
  __shared__ uint16_t nxt[DEC_B1];
  __shared__ uint16_t code[DEC_B1];
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t c = blockIdx.x;
  if (c >= sc.nChunks) return;
  const uint32_t c0 = c * DEC_B1;
  uint32_t c1 = c0 + DEC_B1; if (c1 > sc.clen || c1 < c0) c1 = sc.clen;
  const uint32_t len = c1 - c0;
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x)
  {
    const uint32_t p = c0 + i;
    const HopInfo h = dec_hop(D, p);
    if (h.kind == 0 && h.nxt < c1) { nxt[i] = (uint16_t)(h.nxt - c0); code[i] = 0; }
    else { nxt[i] = (uint16_t)i; code[i] = dec_map_code(c0, c1, p, h); }
  }
  __syncthreads();
  // pointer doubling to the last token of every chain (self loops are fixed points)
  volatile uint16_t *vn = nxt;
  for (;;)
  {
    int changed = 0;
    for (uint32_t i = threadIdx.x; i < len; i += blockDim.x)
    {
      const uint16_t q = vn[i];
      const uint16_t r = vn[q];
      if (r != q) { vn[i] = r; changed = 1; }
    }
    if (!__syncthreads_or(changed)) break;
  }
  for (uint32_t i = threadIdx.x; i < len; i += blockDim.x) D.map16[c0 + i] = code[nxt[i]];
}

__global__ void __launch_bounds__(GS_BLOCK) k_dec_up(const DecBufs D, int lvl)
{
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint64_t S = dec_level_bytes(lvl);
  const uint64_t nItems = (((uint64_t)sc.clen + S - 1) / S) * DEC_WIN;
  for (uint64_t it = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; it < nItems; it += (uint64_t)gridDim.x * blockDim.x) dec_stage_up(D, lvl, it);
}

__global__ void k_dec_top(const DecBufs D) { dec_stage_top(D); }

__global__ void __launch_bounds__(GS_BLOCK) k_dec_down(const DecBufs D, int lvl)
{
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint64_t S = dec_level_bytes(lvl);
  const uint64_t nG = ((uint64_t)sc.clen + S - 1) / S;
  for (uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; g < nG; g += (uint64_t)gridDim.x * blockDim.x) dec_stage_down(D, lvl, (uint32_t)g);
}

template <bool EMIT>
__global__ void __launch_bounds__(GS_BLOCK) k_dec_walk(const DecBufs D)
{
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint32_t nC = sc.nChunks;
  for (uint32_t c = blockIdx.x * blockDim.x + threadIdx.x; c < nC; c += gridDim.x * blockDim.x) dec_chunk_walk<EMIT>(D, c);
}

struct DecElem { uint64_t out; uint32_t tok; uint32_t has; uint64_t sym; };

// single CTA: exclusive scans of tokens / output bytes / symbol carry / LUT transform over chunks
constexpr int DSCAN_T = 256;
__global__ void __launch_bounds__(DSCAN_T) k_dec_scan(const DecBufs D, uint32_t *dResult)
{
  __shared__ DecElem el[DSCAN_T];
  __shared__ LutXf xf[DSCAN_T];
  __shared__ DecElem accE;
  __shared__ LutXf accX;
  DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) { if (threadIdx.x == 0) { dResult[0] = 0; dResult[1] = sc.status; } return; }
  const int K = D.sp.K;
  const uint32_t nC = sc.nChunks;
  const uint32_t per = (nC + DSCAN_T - 1) / DSCAN_T;
  const uint32_t lo = min(nC, threadIdx.x * per), hi = min(nC, lo + per);
  DecElem me; me.out = 0; me.tok = 0; me.has = 0; me.sym = 0;
  LutXf mx; lutxf_identity(mx);
  for (uint32_t c = lo; c < hi; c++)
  {
    me.out += D.cOut[c]; me.tok += D.cTok[c];
    if (K) mx = lutxf_compose(mx, D.cXf[c], K);
    else if (D.cHasSym[c]) { me.has = 1; me.sym = D.cSym[c]; }
  }
  el[threadIdx.x] = me; if (K) xf[threadIdx.x] = mx;
  __syncthreads();
  if (threadIdx.x == 0)
  { // sequential exclusive scan over the 256 slice composites
    DecElem a; a.out = 0; a.tok = 0; a.has = 0; a.sym = 0;
    LutXf ax; lutxf_identity(ax);
    for (int i = 0; i < DSCAN_T; i++)
    {
      const DecElem cur = el[i]; el[i] = a;
      a.out += cur.out; a.tok += cur.tok; if (cur.has) { a.has = 1; a.sym = cur.sym; }
      if (K) { const LutXf cx = xf[i]; xf[i] = ax; ax = lutxf_compose(ax, cx, K); }
    }
    accE = a; accX = ax;
  }
  __syncthreads();
  {
    DecElem a = el[threadIdx.x];
    LutXf ax; if (K) ax = xf[threadIdx.x]; else lutxf_identity(ax);
    Lut init; lut_init(init, D.sp.W);
    for (uint32_t c = lo; c < hi; c++)
    {
      const uint64_t o = D.cOut[c]; const uint32_t t = D.cTok[c];
      D.cOut[c] = a.out; D.cTok[c] = a.tok; a.out += o; a.tok += t;
      if (K) { Lut l; lutxf_apply(ax, K, init, l); D.cLutIn[c] = l; ax = lutxf_compose(ax, D.cXf[c], K); }
      else { const uint64_t s = D.cSym[c]; const bool h = D.cHasSym[c] != 0; D.cSym[c] = a.sym; if (h) a.sym = s; }
    }
  }
  __syncthreads();
  if (threadIdx.x == 0)
  {
    sc.nTok = accE.tok; sc.outTotal = accE.out;
    if (!sc.endSeen || accE.out != sc.n || accE.tok > D.maxTok) sc.status = ST_BADSTREAM;
    else { D.tOut[sc.nTok] = sc.n; D.tLitLen[sc.nTok] = 0; }
    dResult[0] = sc.status == ST_OK ? sc.n : 0; dResult[1] = sc.status; dResult[2] = sc.nTok; dResult[3] = sc.nChunks;
    dResult[4] = sc.clen; dResult[5] = sc.single; dResult[6] = 0; dResult[7] = 0;
  }
}

__global__ void __launch_bounds__(256) k_dec_expand(const DecBufs D)
{
  const DecScalars &sc = *D.sc;
  if (sc.status != ST_OK) return;
  const uint64_t n = sc.n;
  const uint64_t nv = (n + 15) >> 4;
  for (uint64_t vi = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; vi < nv; vi += (uint64_t)gridDim.x * blockDim.x)
  {
    const uint64_t v = vi << 4;
    __align__(16) uint8_t tmp[16];
    dec_expand_vec(D, v, tmp);
    if (v + 16 <= n) *reinterpret_cast<uint4 *>(D.out + v) = *reinterpret_cast<const uint4 *>(tmp);
    else for (uint64_t i = 0; v + i < n; i++) D.out[v + i] = tmp[i];
  }
}

// ================================================================================================
// host side: workspace carving + launch sequences
struct Carver
{
  uint8_t *base; size_t off;
  template <typename T> T *take(size_t count)
  {
    off = (off + 255) & ~(size_t)255;
    T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
    off += count * sizeof(T);
    return p;
  }
};

static size_t enc_carve(EncBufs &B, const Spec &sp, uint32_t n, void *ws)
{
  Carver cv{ (uint8_t *)ws, 0 };
  B.sp = sp; B.n = n;
  B.nVec = (uint32_t)(((uint64_t)n + 1 + ENC_VEC - 1) / ENC_VEC);
  B.nTiles = (B.nVec + ENC_TILE_VECS - 1) / ENC_TILE_VECS;
  B.maxRuns = n / (sp.minM + 1) + 2;
  const size_t maxChunks = B.maxRuns / ENC_CH + 2;
  B.sc = cv.take<EncScalars>(1);
  B.tileS = cv.take<uint32_t>(B.nTiles + 1); B.tileE = cv.take<uint32_t>(B.nTiles + 1);
  B.runA = cv.take<uint32_t>(B.maxRuns); B.runB = cv.take<uint32_t>(B.maxRuns);
  B.sIn = cv.take<AutoState>(maxChunks); B.cSum = cv.take<ChunkSum>(maxChunks);
  B.lutIn = cv.take<Lut>(sp.K ? maxChunks : 1); B.lutAgg = cv.take<LutAgg>(sp.K ? maxChunks : 1);
  B.cBytes = cv.take<uint64_t>(maxChunks); B.cTok = cv.take<uint32_t>(maxChunks); B.dirty = cv.take<uint8_t>(maxChunks);
  B.copies = cv.take<CopyDesc>((size_t)B.maxRuns + 2);
  B.bigList = cv.take<uint32_t>((size_t)n / BIG_COPY + 4);
  return cv.off + 256;
}

static int dec_top_level(uint32_t inSize)
{
  uint64_t g = ((uint64_t)inSize + DEC_B1 - 1) / DEC_B1;
  int T = 0;
  while (g > DEC_G && T < DEC_MAX_LEVELS) { g = (g + DEC_G - 1) / DEC_G; T++; }
  return T;
}

static size_t dec_carve(DecBufs &D, const Spec &sp, uint32_t inSize, uint32_t outSize, void *ws)
{
  Carver cv{ (uint8_t *)ws, 0 };
  D.sp = sp; D.inSize = inSize; D.outSize = outSize;
  const size_t nC = ((size_t)inSize + DEC_B1 - 1) / DEC_B1;
  D.sc = cv.take<DecScalars>(1);
  D.map16 = cv.take<uint16_t>(nC * DEC_B1);
  D.topLevel = dec_top_level(inSize);
  for (int l = 0; l <= DEC_MAX_LEVELS; l++) { D.lmap[l] = nullptr; D.lentry[l] = nullptr; }
  for (int l = 0; l <= D.topLevel; l++)
  {
    const uint64_t S = dec_level_bytes(l);
    const size_t nG = (size_t)(((uint64_t)inSize + S - 1) / S);
    if (l >= 1) D.lmap[l] = cv.take<uint32_t>(nG * DEC_WIN);
    D.lentry[l] = cv.take<uint32_t>(nG + DEC_G);
  }
  D.cTok = cv.take<uint32_t>(nC + 1); D.cOut = cv.take<uint64_t>(nC + 1);
  D.cSym = cv.take<uint64_t>(nC + 1); D.cHasSym = cv.take<uint8_t>(nC + 1);
  D.cXf = cv.take<LutXf>(sp.K ? nC + 1 : 1); D.cLutIn = cv.take<Lut>(sp.K ? nC + 1 : 1);
  D.maxTok = inSize / 2 + 2;
  D.tOut = cv.take<uint32_t>((size_t)D.maxTok + 2); D.tLitSrc = cv.take<uint32_t>((size_t)D.maxTok + 2);
  D.tLitLen = cv.take<uint32_t>((size_t)D.maxTok + 2); D.tSym = cv.take<uint64_t>((size_t)D.maxTok + 2);
  D.tileFirst = cv.take<uint32_t>((size_t)outSize / DEC_TILE + 4);
  return cv.off + 256;
}

static bool spec_from_codec(int codec, Spec &sp)
{
  if (codec < 0 || codec >= 48) return false;
  const int wi = codec >> 3, ba = (codec >> 2) & 1, var = codec & 3;
  const int W = width_from_index(wi);
  if (W == 1 && !ba) return false;
  sp = make_spec(W, ba, var);
  return true;
}

constexpr int ENC_ROUNDS = 4;

static int enc_enqueue(int codec, const uint8_t *dIn, uint32_t n, uint8_t *dOut, uint32_t cap, void *ws, size_t wsSize, uint32_t *dResult, cudaStream_t st)
{
  Spec sp;
  if (!spec_from_codec(codec, sp) || !dIn || !dOut || !ws || !dResult || n == 0) { g_err = "bad argument"; return 1; }
  if (((uintptr_t)dIn & 15) || ((uintptr_t)dOut & 15) || ((uintptr_t)ws & 255)) { g_err = "device pointers must be 16-byte aligned (workspace 256)"; return 1; }
  EncBufs B; memset(&B, 0, sizeof(B));
  const size_t need = enc_carve(B, sp, n, ws);
  if (need > wsSize) { g_err = "workspace too small"; return 1; }
  B.in = dIn; B.out = dOut; B.cap = cap;
  if (!cuda_ok(cudaMemsetAsync(B.sc, 0, sizeof(EncScalars), st), "memset")) return 2;
  HSRLE_LAUNCH(k_enc_count, B.nTiles, ENC_TILE_VECS, 0, st, B);
  HSRLE_LAUNCH(k_enc_scan_tiles, 1, 1024, 0, st, B);
  HSRLE_LAUNCH(k_enc_write, B.nTiles, ENC_TILE_VECS, 0, st, B);
  HSRLE_LAUNCH(k_enc_auto_init, GS_GRID, GS_BLOCK, 0, st, B);
  const size_t scanSmem = (SCAN_T + SCAN_T / 32) * sizeof(ScanElem);
  for (int r = 0; r < ENC_ROUNDS; r++)
  {
    HSRLE_LAUNCH(k_enc_scan_check, 1, SCAN_T, scanSmem, st, B, r);
    HSRLE_LAUNCH(k_enc_rerun, GS_GRID, GS_BLOCK, 0, st, B);
  }
  HSRLE_LAUNCH(k_enc_scan_check, 1, SCAN_T, scanSmem, st, B, ENC_ROUNDS);
  HSRLE_LAUNCH(k_enc_serial, 1, 1, 0, st, B);
  HSRLE_LAUNCH(k_enc_sizes, 1, 1024, 0, st, B, dResult);
  HSRLE_LAUNCH(k_enc_emit, GS_GRID, GS_BLOCK, 0, st, B);
  HSRLE_LAUNCH(k_enc_copy_small, GS_GRID, 256, 0, st, B);
  HSRLE_LAUNCH(k_enc_copy_big, 148 * 4, 256, 0, st, B);
  return cuda_ok(cudaGetLastError(), "encode launch") ? 0 : 2;
}

static int dec_enqueue(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize, void *ws, size_t wsSize, uint32_t *dResult, cudaStream_t st)
{
  Spec sp;
  if (!spec_from_codec(codec, sp) || !dIn || !dOut || !ws || !dResult || inSize == 0 || outSize == 0) { g_err = "bad argument"; return 1; }
  if (((uintptr_t)dIn & 15) || ((uintptr_t)dOut & 15) || ((uintptr_t)ws & 255)) { g_err = "device pointers must be 16-byte aligned (workspace 256)"; return 1; }
  DecBufs D; memset(&D, 0, sizeof(D));
  const size_t need = dec_carve(D, sp, inSize, outSize, ws);
  if (need > wsSize) { g_err = "workspace too small"; return 1; }
  D.in = dIn; D.out = dOut;
  const uint32_t nC = (uint32_t)(((uint64_t)inSize + DEC_B1 - 1) / DEC_B1);
  HSRLE_LAUNCH(k_dec_init, 1, 1, 0, st, D);
  HSRLE_LAUNCH(k_dec_map, nC, 256, 0, st, D);
  for (int l = 1; l <= D.topLevel; l++) HSRLE_LAUNCH(k_dec_up, GS_GRID, GS_BLOCK, 0, st, D, l);
  HSRLE_LAUNCH(k_dec_top, 1, 1, 0, st, D);
  for (int l = D.topLevel; l >= 1; l--) HSRLE_LAUNCH(k_dec_down, GS_GRID, GS_BLOCK, 0, st, D, l);
  HSRLE_LAUNCH(k_dec_walk<false>, GS_GRID, GS_BLOCK, 0, st, D);
  HSRLE_LAUNCH(k_dec_scan, 1, DSCAN_T, 0, st, D, dResult);
  HSRLE_LAUNCH(k_dec_walk<true>, GS_GRID, GS_BLOCK, 0, st, D);
  HSRLE_LAUNCH(k_dec_expand, GS_GRID * 2, 256, 0, st, D);
  return cuda_ok(cudaGetLastError(), "decode launch") ? 0 : 2;
}

// ------------------------------------------------------------------------------------------------
// library-owned context for the synchronous entry points
struct Context
{
  std::mutex mu;
  int dev = -1;
  bool tried = false;
  cudaStream_t stream = nullptr;
  void *ws = nullptr; size_t wsSize = 0;
  uint8_t *dIn = nullptr; size_t dInSize = 0;
  uint8_t *dOut = nullptr; size_t dOutSize = 0;
  uint32_t *dResult = nullptr; uint32_t *hResult = nullptr;

  bool init()
  {
    if (tried) return dev >= 0;
    tried = true;
    int count = 0;
    if (!cuda_ok(cudaGetDeviceCount(&count), "cudaGetDeviceCount") || count == 0) { if (g_err.empty()) g_err = "no CUDA device"; return false; }
    int d = 0;
    if (!cuda_ok(cudaGetDevice(&d), "cudaGetDevice")) return false;
    if (!cuda_ok(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking), "stream")) return false;
    if (!cuda_ok(cudaMalloc(&dResult, 64), "malloc result")) return false;
    if (!cuda_ok(cudaMallocHost(&hResult, 64), "malloc host result")) return false;
    cudaFuncSetAttribute(k_enc_scan_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((SCAN_T + SCAN_T / 32) * sizeof(ScanElem)));
    dev = d;
    return true;
  }
  bool grow(void **p, size_t *cur, size_t need)
  {
    if (*cur >= need) return true;
    if (*p) cudaFree(*p);
    *p = nullptr; *cur = 0;
    need += need / 8 + 4096;
    if (!cuda_ok(cudaMalloc(p, need), "cudaMalloc workspace")) return false;
    *cur = need;
    return true;
  }
};
static Context g_ctx;
static std::once_flag g_attrOnce;

static void set_func_attrs()
{
  std::call_once(g_attrOnce, [] {
    cudaFuncSetAttribute(k_enc_scan_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((SCAN_T + SCAN_T / 32) * sizeof(ScanElem)));
  });
}

static uint32_t run_sync(bool compress, int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize)
{
  Context &C = g_ctx;
  Spec sp;
  if (!spec_from_codec(codec, sp)) return 0;
  EncBufs B; DecBufs D;
  const size_t need = compress ? enc_carve(B, sp, inSize, nullptr) : dec_carve(D, sp, inSize, outSize, nullptr);
  if (!C.grow(&C.ws, &C.wsSize, need)) return 0;
  const int rc = compress ? enc_enqueue(codec, dIn, inSize, dOut, outSize, C.ws, C.wsSize, C.dResult, C.stream)
                          : dec_enqueue(codec, dIn, inSize, dOut, outSize, C.ws, C.wsSize, C.dResult, C.stream);
  if (rc) return 0;
  if (!cuda_ok(cudaMemcpyAsync(C.hResult, C.dResult, 32, cudaMemcpyDeviceToHost, C.stream), "result copy")) return 0;
  if (!cuda_ok(cudaStreamSynchronize(C.stream), "synchronize")) return 0;
  return C.hResult[1] == ST_OK ? C.hResult[0] : 0;
}

} // namespace hsrle

// ================================================================================================
// C ABI
using namespace hsrle;

extern "C" {

uint32_t rle_compress_bounds(const uint32_t inSize)
{
  if (inSize > (1u << 30)) return 0;
  return inSize + (16 + 4 + 1 + 4 + 1 + 64) * 2 + 12 + 1;
}
uint32_t rle_decompress_additional_size(void) { return 128; }

int hsrle_codec_id(int symbolBits, int byteAligned, int variant)
{
  int wi;
  switch (symbolBits) { case 8: wi = 0; break; case 16: wi = 1; break; case 24: wi = 2; break; case 32: wi = 3; break; case 48: wi = 4; break; case 64: wi = 5; break; default: return -1; }
  if (variant < 0 || variant > 3) return -1;
  if (wi == 0) byteAligned = 1;
  return wi * 8 + (byteAligned ? 4 : 0) + variant;
}

int hsrle_codec_id_from_name(const char *name)
{
  if (!name) return -1;
  int bits = 0; const char *p = name;
  if (strncmp(p, "rle", 3) != 0) return -1;
  p += 3;
  while (*p >= '0' && *p <= '9') { bits = bits * 10 + (*p - '0'); p++; }
  if (*p != '_') return -1;
  p++;
  const std::string rest(p);
  if (bits == 8)
  {
    if (rest == "multi" || rest == "") return hsrle_codec_id(8, 1, 0);
    if (rest == "packed_multi" || rest == "packed") return hsrle_codec_id(8, 1, 1);
    if (rest == "3symlut") return hsrle_codec_id(8, 1, 2);
    if (rest == "7symlut") return hsrle_codec_id(8, 1, 3);
    return -1;
  }
  if (rest == "sym") return hsrle_codec_id(bits, 0, 0);
  if (rest == "byte") return hsrle_codec_id(bits, 1, 0);
  if (rest == "sym_packed") return hsrle_codec_id(bits, 0, 1);
  if (rest == "byte_packed") return hsrle_codec_id(bits, 1, 1);
  if (rest == "3symlut_sym") return hsrle_codec_id(bits, 0, 2);
  if (rest == "3symlut_byte") return hsrle_codec_id(bits, 1, 2);
  if (rest == "7symlut_sym") return hsrle_codec_id(bits, 0, 3);
  if (rest == "7symlut_byte") return hsrle_codec_id(bits, 1, 3);
  return -1;
}

size_t hsrle_compress_workspace_size(int codec, uint32_t inSize)
{
  Spec sp; if (!spec_from_codec(codec, sp) || inSize == 0) return 0;
  EncBufs B; return enc_carve(B, sp, inSize, nullptr);
}
size_t hsrle_decompress_workspace_size(int codec, uint32_t inSize, uint32_t outSize)
{
  Spec sp; if (!spec_from_codec(codec, sp) || inSize == 0) return 0;
  DecBufs D; return dec_carve(D, sp, inSize, outSize, nullptr);
}

int hsrle_compress_device_async(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize,
                                void *dWorkspace, size_t workspaceSize, uint32_t *dResult, void *cudaStream)
{
  set_func_attrs();
  return enc_enqueue(codec, dIn, inSize, dOut, outSize, dWorkspace, workspaceSize, dResult, (cudaStream_t)cudaStream);
}
int hsrle_decompress_device_async(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize,
                                  void *dWorkspace, size_t workspaceSize, uint32_t *dResult, void *cudaStream)
{
  set_func_attrs();
  return dec_enqueue(codec, dIn, inSize, dOut, outSize, dWorkspace, workspaceSize, dResult, (cudaStream_t)cudaStream);
}

uint32_t hsrle_compress_device(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize)
{
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (!g_ctx.init()) return 0;
  return run_sync(true, codec, dIn, inSize, dOut, outSize);
}
uint32_t hsrle_decompress_device(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize)
{
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  if (!g_ctx.init()) return 0;
  return run_sync(false, codec, dIn, inSize, dOut, outSize);
}

uint32_t hsrle_compress_host(int codec, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize)
{
  // preconditions of the reference: src/rle8_extreme_cpu.h:88, src/rleX_extreme_cpu.h:49, src/rleX_Xsl.h:271
  if (pIn == NULL || inSize == 0 || pOut == NULL || outSize < rle_compress_bounds(inSize)) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  Context &C = g_ctx;
  if (!C.init()) return 0;
  if (!C.grow((void **)&C.dIn, &C.dInSize, (size_t)inSize + 64)) return 0;
  if (!C.grow((void **)&C.dOut, &C.dOutSize, (size_t)outSize + 64)) return 0;
  if (!cuda_ok(cudaMemcpyAsync(C.dIn, pIn, inSize, cudaMemcpyHostToDevice, C.stream), "H2D")) return 0;
  const uint32_t r = run_sync(true, codec, C.dIn, inSize, C.dOut, outSize);
  if (r == 0) return 0;
  if (!cuda_ok(cudaMemcpyAsync(pOut, C.dOut, r, cudaMemcpyDeviceToHost, C.stream), "D2H")) return 0;
  if (!cuda_ok(cudaStreamSynchronize(C.stream), "synchronize")) return 0;
  return r;
}

uint32_t hsrle_decompress_host(int codec, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize)
{
  if (pIn == NULL || pOut == NULL || inSize == 0 || outSize == 0) return 0;
  // header check on the host first (src/rle8_extreme_cpu.h:707-712): only the stream itself is uploaded
  if (inSize < 8) return 0;
  uint32_t n, clen; memcpy(&n, pIn, 4); memcpy(&clen, pIn + 4, 4);
  if (n > outSize || clen > inSize || clen < 8) return 0;
  if (n == 0) return 0;
  std::lock_guard<std::mutex> lk(g_ctx.mu);
  Context &C = g_ctx;
  if (!C.init()) return 0;
  if (!C.grow((void **)&C.dIn, &C.dInSize, (size_t)clen + 64)) return 0;
  if (!C.grow((void **)&C.dOut, &C.dOutSize, (size_t)n + 64)) return 0;
  if (!cuda_ok(cudaMemcpyAsync(C.dIn, pIn, clen, cudaMemcpyHostToDevice, C.stream), "H2D")) return 0;
  const uint32_t r = run_sync(false, codec, C.dIn, clen, C.dOut, n);
  if (r == 0) return 0;
  if (!cuda_ok(cudaMemcpyAsync(pOut, C.dOut, r, cudaMemcpyDeviceToHost, C.stream), "D2H")) return 0;
  if (!cuda_ok(cudaStreamSynchronize(C.stream), "synchronize")) return 0;
  return r;
}

void hsrle_timing_begin(void)
{
  for (auto &t : g_timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  g_timed.clear(); g_timing = true;
}
// Synchronises the device and writes "kernel:launches:total_ms;..." for every kernel launched since
// hsrle_timing_begin().  Returns the number of characters written.
int hsrle_timing_end(char *buf, int bufSize)
{
  g_timing = false;
  cudaDeviceSynchronize();
  std::map<std::string, std::pair<int, double>> acc;
  for (auto &t : g_timed)
  {
    float ms = 0; cudaEventElapsedTime(&ms, t.a, t.b);
    auto &e = acc[t.name]; e.first++; e.second += ms;
    cudaEventDestroy(t.a); cudaEventDestroy(t.b);
  }
  g_timed.clear();
  std::string out;
  for (auto &kv : acc) { char tmp[256]; snprintf(tmp, sizeof(tmp), "%s:%d:%.6f;", kv.first.c_str(), kv.second.first, kv.second.second); out += tmp; }
  if (!buf || bufSize <= 0) return 0;
  const int nw = (int)std::min<size_t>(out.size(), (size_t)bufSize - 1);
  memcpy(buf, out.data(), nw); buf[nw] = 0;
  return nw;
}

const char *hsrle_last_error(void) { return g_err.c_str(); }
int hsrle_device(void) { std::lock_guard<std::mutex> lk(g_ctx.mu); g_ctx.init(); return g_ctx.dev; }
uint64_t hsrle_kernel_launches(void) { return g_launches.load(); }

#define HSRLE_PAIR(cname, dname, bits, ba, var)                                                                                   \
  uint32_t cname(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize)                                 \
  { return hsrle_compress_host(hsrle_codec_id(bits, ba, var), pIn, inSize, pOut, outSize); }                                       \
  uint32_t dname(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize)                                 \
  { return hsrle_decompress_host(hsrle_codec_id(bits, ba, var), pIn, inSize, pOut, outSize); }

HSRLE_PAIR(rle8_multi_compress, rle8_decompress, 8, 1, 0)
HSRLE_PAIR(rle8_packed_multi_compress, rle8_packed_decompress, 8, 1, 1)
HSRLE_PAIR(rle8_3symlut_compress, rle8_3symlut_decompress, 8, 1, 2)
HSRLE_PAIR(rle8_7symlut_compress, rle8_7symlut_decompress, 8, 1, 3)
#define HSRLE_WIDTH(bits)                                                                        \
  HSRLE_PAIR(rle##bits##_sym_compress, rle##bits##_sym_decompress, bits, 0, 0)                   \
  HSRLE_PAIR(rle##bits##_byte_compress, rle##bits##_byte_decompress, bits, 1, 0)                 \
  HSRLE_PAIR(rle##bits##_sym_packed_compress, rle##bits##_sym_packed_decompress, bits, 0, 1)     \
  HSRLE_PAIR(rle##bits##_byte_packed_compress, rle##bits##_byte_packed_decompress, bits, 1, 1)   \
  HSRLE_PAIR(rle##bits##_3symlut_sym_compress, rle##bits##_3symlut_sym_decompress, bits, 0, 2)   \
  HSRLE_PAIR(rle##bits##_3symlut_byte_compress, rle##bits##_3symlut_byte_decompress, bits, 1, 2) \
  HSRLE_PAIR(rle##bits##_7symlut_sym_compress, rle##bits##_7symlut_sym_decompress, bits, 0, 3)   \
  HSRLE_PAIR(rle##bits##_7symlut_byte_compress, rle##bits##_7symlut_byte_decompress, bits, 1, 3)
HSRLE_WIDTH(16)
HSRLE_WIDTH(24)
HSRLE_WIDTH(32)
HSRLE_WIDTH(48)
HSRLE_WIDTH(64)

} // extern "C"
