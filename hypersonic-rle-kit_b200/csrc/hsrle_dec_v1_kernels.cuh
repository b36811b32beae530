// hsrle_dec_v1_kernels.cuh -- first-generation decoder kernels (thin grid-stride wrappers of hsrle_dec_v1.cuh).
#pragma once
#include <cuda_runtime.h>
#include "hsrle_dec_v1.cuh"

namespace hsrle {

constexpr int GS_BLOCK = 128;            // threads of grid-stride "one item per thread" kernels
constexpr int GS_GRID = 148 * 8;         // 148 SMs x 8 resident CTAs

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

} // namespace hsrle
