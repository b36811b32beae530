// hsrle_api.cu -- host side of the B200 extreme-RLE codec: C ABI, workspace carving, launch sequences.
//
// Reference entry points this file replaces: src/rle.h:100-394 (see include/hsrle_b200.h).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <time.h>
#include <unistd.h>
#include <mutex>
#include <string>
#include <vector>
#include <map>

#include "../../include/hsrle_b200.h"
#include "hsrle_dispatch.h"
#include "hsrle_enc_kernels.cuh"
#include "hsrle_dec_kernels.cuh"
#include "hsrle_slice_kernels.cuh"

namespace hsrle {

static thread_local std::string g_err;
static std::atomic<uint64_t> g_launches{0};

// optional per-kernel CUDA-event timing (bench.py's roofline leg); off by default
struct TimedLaunch { const char *name; cudaEvent_t a, b; };
static std::atomic<bool> g_timing{false};
static std::mutex g_timedMu;
static std::vector<TimedLaunch> g_timed;

#define HSRLE_LAUNCH_NAMED(name, kern, grid, block, smem, stream, ...)     \
  do {                                                                     \
    TimedLaunch tl_{ name, nullptr, nullptr };                             \
    if (g_timing && cudaEventCreate(&tl_.a) == cudaSuccess)                \
    {                                                                      \
      if (cudaEventCreate(&tl_.b) != cudaSuccess) { cudaEventDestroy(tl_.a); tl_.a = nullptr; } \
      else cudaEventRecord(tl_.a, (stream));                               \
    }                                                                      \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__);              \
    if (tl_.a) { cudaEventRecord(tl_.b, (stream)); std::lock_guard<std::mutex> lk_(g_timedMu); g_timed.push_back(tl_); } \
    g_launches.fetch_add(1, std::memory_order_relaxed);                    \
  } while (0)
#define HSRLE_LAUNCH(kern, grid, block, smem, stream, ...) HSRLE_LAUNCH_NAMED(#kern, kern, grid, block, smem, stream, __VA_ARGS__)

static bool cuda_ok(cudaError_t e, const char *what)
{
  if (e == cudaSuccess) return true;
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return false;
}

// per-device caches: cudaFuncSetAttribute and the SM count belong to the CURRENT device, and one process may use several
constexpr int MAX_DEV = 64;
static int current_dev()
{
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= MAX_DEV) dev = 0;
  return dev;
}
static std::atomic<int> g_numSM[MAX_DEV];
static int num_sms()
{
  const int dev = current_dev();
  int n = g_numSM[dev].load(std::memory_order_relaxed);
  if (n == 0)
  {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    g_numSM[dev].store(n, std::memory_order_relaxed);
  }
  return n;
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

static const EncKernels *enc_kernels_for(int codec)
{
  const int wi = codec >> 3;
  const EncKernels *tab = nullptr;
  switch (wi)
  {
    case 0: tab = enc_kernels_w1(); break; case 1: tab = enc_kernels_w2(); break; case 2: tab = enc_kernels_w3(); break;
    case 3: tab = enc_kernels_w4(); break; case 4: tab = enc_kernels_w6(); break; case 5: tab = enc_kernels_w8(); break;
    default: return nullptr;
  }
  const EncKernels *k = tab + (codec & 7);
  return k->scan ? k : nullptr;
}

// E1 tile size: one CTA per tile, about 4 CTAs resident per SM.  A grid that is a little larger than a whole number of
// waves leaves the last wave almost empty (88 MB in 128-KiB tiles: 675 tiles on 592 slots), so pick the steps per warp
// (multiple of 4) that minimises waves x tile time, with a fixed per-tile cost (ticket, look-back, two barriers) of
// about two steps.  Depends only on the vector count and the SM count: deterministic per call, as the workspace layout
// needs it.
static uint32_t enc_scan_steps(uint32_t nVec, bool pow2Only = false)
{
  const uint64_t slots = (uint64_t)num_sms() * 4;
  uint32_t best = E1_STEPS; uint64_t bestCost = ~0ull;
  for (uint32_t steps = 8; steps <= (uint32_t)E1_STEPS; steps += 4)
  {
    if (pow2Only && (steps & (steps - 1))) continue;
    const uint64_t tiles = ((uint64_t)nVec + (uint64_t)E1_T * steps - 1) / ((uint64_t)E1_T * steps);
    const uint64_t waves = (tiles + slots - 1) / slots;
    const uint64_t cost = waves * (steps + 2);
    if (cost < bestCost || (cost == bestCost && steps > best)) { best = steps; bestCost = cost; }
  }
  return best;
}

// grid of the verify / repair kernel: the master CTA plus helpers for rounds with many dirty super-chunks.  Few super-chunks
// (small inputs) never have such rounds: the master alone.
// super-chunks a repair warp follows a changed state through per round (1: none); HSRLE_FOLLOW overrides (experiments)
static int enc_follow_hops()
{
  static const int v = getenv("HSRLE_FOLLOW") ? std::max(1, atoi(getenv("HSRLE_FOLLOW"))) : 1;
  return v;
}
// HSRLE_LUTWALK=0 switches the stretch walk of the 8-bit LUT codecs off (experiments; the result is the same either way)
static bool lut_walk_enabled()
{
  static const bool v = !(getenv("HSRLE_LUTWALK") && atoi(getenv("HSRLE_LUTWALK")) == 0);
  return v;
}
static int enc_fix_grid(const EncBufs &B, int sms)
{
  return B.maxSC <= 4 * FIX_SOLO ? 1 : std::min<int>(sms, 1 + (int)(B.maxSC / 64));
}

static size_t enc_carve(EncBufs &B, const Spec &sp, uint32_t n, void *ws, size_t *zeroBytes, bool sliceTiles = false)
{
  Carver cv{ (uint8_t *)ws, 0 };
  B.n = n;
  B.nVec = (uint32_t)(((uint64_t)n + 1 + 15) / 16);
  B.lastVec = (n - 1) >> 4;
  B.scanSteps = enc_scan_steps(B.nVec, sliceTiles);
  B.nTiles = (B.nVec + E1_T * B.scanSteps - 1) / (E1_T * B.scanSteps) + (sliceTiles ? 1u : 0u);   // (slices: the last rank's vector count is a little larger)
  B.maxRuns = n / (sp.minM + 1) + 2;
  B.maxSC = B.maxRuns / E2_SCR + 2;
  const size_t maxChunks = (size_t)B.maxSC * E2_T;
  // zero-initialised region first: scalars + look-back status words
  B.sc = cv.take<EncScalars>(1);
  B.tileStatus = cv.take<unsigned long long>(2 * (size_t)B.nTiles + 2);
  if (zeroBytes) *zeroBytes = cv.off;
  B.runA = cv.take<uint32_t>(B.maxRuns); B.runB = cv.take<uint32_t>(B.maxRuns);
  B.runSym = cv.take<uint64_t>((sp.W <= 4 ? ((size_t)B.maxRuns + 1) / 2 : (size_t)B.maxRuns) + 8);   // (+8: k_enc_lut_stretch reads whole 16-record groups)
  B.cIn = cv.take<AutoState>(maxChunks); B.cLut = cv.take<Lut>(sp.K ? maxChunks : 1); B.cKnown = cv.take<uint8_t>(sp.K ? maxChunks : 1);
  B.scFo = cv.take<Lut>(sp.K ? B.maxSC : 1); B.scFlags = cv.take<uint8_t>(sp.K ? B.maxSC : 1); B.scQ = cv.take<ScQueries>(sp.K ? B.maxSC : 1);
  B.scIn = cv.take<AutoState>(B.maxSC); B.scLut = cv.take<Lut>(sp.K ? B.maxSC : 1);
  B.scSum = cv.take<ChunkSum>(B.maxSC); B.scAgg = cv.take<LutAgg>(sp.K ? B.maxSC : 1);
  B.scBytes = cv.take<uint64_t>(B.maxSC); B.scTok = cv.take<uint32_t>(B.maxSC); B.scBase = cv.take<uint64_t>(B.maxSC);
  B.scDirty = cv.take<uint8_t>(B.maxSC);
  if (sp.K && sp.W == 1)
  { // stretch walk (hsrle_enc_lutwalk.cuh)
    const size_t nBlk = (size_t)B.maxRuns / LW_BLK + 2;
    B.lwPool = cv.take<LwDesc>(LW_CAP); B.lwOrd = cv.take<LwDesc>(LW_CAP);
    B.lwBlkOff = cv.take<uint32_t>(nBlk); B.lwBlkCnt = cv.take<uint32_t>(nBlk);
    B.scGuess = cv.take<uint64_t>(B.maxSC);
  }
  B.bigList = cv.take<CopyDesc>((size_t)n / BIG_COPY + 4);
  B.medList = cv.take<CopyDesc>((size_t)n / MED_COPY + 4);
  return cv.off + 256;
}

static const DecKernels *dec_kernels_for(int codec)
{
  const int wi = codec >> 3;
  const DecKernels *tab = nullptr;
  switch (wi)
  {
    case 0: tab = dec_kernels_w1(); break; case 1: tab = dec_kernels_w2(); break; case 2: tab = dec_kernels_w3(); break;
    case 3: tab = dec_kernels_w4(); break; case 4: tab = dec_kernels_w6(); break; case 5: tab = dec_kernels_w8(); break;
    default: return nullptr;
  }
  const DecKernels *k = tab + (codec & 7);
  return k->map ? k : nullptr;
}

static size_t dec_carve(DecBufs &D, const Spec &sp, uint32_t inSize, uint32_t outSize, void *ws, size_t *zeroBytes)
{
  Carver cv{ (uint8_t *)ws, 0 };
  D.inSize = inSize; D.outSize = outSize;
  D.nChunks = (uint32_t)(((uint64_t)inSize + DEC_CB - 1) / DEC_CB);
  D.nSeg = (D.nChunks + DEC_SEG - 1) / DEC_SEG;
  const size_t aggBytes = sp.K ? sizeof(DecAgg<7>) : sizeof(DecAgg<0>);
  // zero-initialised region first
  D.cnt = cv.take<DecCounters>(1);
  D.segCount = cv.take<uint32_t>((size_t)D.nSeg + 1);
  D.anchorAt = cv.take<uint32_t>((size_t)D.nChunks + 1);
  D.skipFlag = cv.take<uint8_t>((size_t)D.nChunks + 1);
  D.flagAgg = cv.take<uint32_t>((size_t)D.nChunks + 1);
  D.bigCap = (uint32_t)(((size_t)outSize / ((size_t)DEC_TILE * DEC_HUGE_TILES)) + 64);      // every such operation covers at least 64 KiB of output
  D.bigList = cv.take<DecBigOp>(D.bigCap);                                                  // (their `ready` words must start as 0)
  D.pieceCap = (uint32_t)((size_t)outSize / DEC_BIG_PIECE) + D.bigCap + 64;                 // whole pieces + one partial piece per operation
  D.pieceOp = cv.take<uint32_t>(D.pieceCap);
  if (zeroBytes) *zeroBytes = cv.off;
  D.sc = cv.take<DecScalars>(1);
  D.chunkTab = cv.take<uint16_t>((size_t)D.nChunks * DEC_CB);
  D.sufMap = cv.take<uint32_t>((size_t)D.nChunks * DEC_WINC);
  D.segTab = cv.take<uint32_t>((size_t)D.nChunks * DEC_CB);
  D.chunkEntry = cv.take<uint32_t>((size_t)D.nChunks + 1);
  D.liveList = cv.take<uint32_t>((size_t)D.nChunks + 1);
  D.subMap = cv.take<uint16_t>((size_t)D.nChunks * DEC_NSUB * DEC_WIN);
  D.aggBuf = cv.take<uint8_t>(((size_t)D.nChunks + 1) * aggBytes); D.incBuf = cv.take<uint8_t>(((size_t)D.nChunks + 1) * aggBytes);
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

static std::mutex g_attrMu;
static bool g_attrDone[MAX_DEV][48];
static bool enc_prepare(int codec, const EncKernels *k)
{
  const int dev = current_dev();
  std::lock_guard<std::mutex> lk(g_attrMu);
  if (g_attrDone[dev][codec]) return true;
  // the bandwidth kernels ask for the same (maximum) shared-memory carve-out: CTAs of kernels with different carve-outs
  // cannot share an SM, and the calls of several streams are meant to overlap.  Not the automaton: its register-capped
  // LUT variants keep spilled table entries in L1 (measured: +10 % time with the small L1)
  cudaFuncSetAttribute((const void *)k->scan, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute((const void *)k->emit, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute((const void *)k_enc_copy_big, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (!cuda_ok(cudaFuncSetAttribute((const void *)k->autom, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->autoSmem), "attr auto")) return false;
  if (!cuda_ok(cudaFuncSetAttribute((const void *)k->fix, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->fixSmem), "attr fix")) return false;
  if (!cuda_ok(cudaFuncSetAttribute((const void *)k->emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->emitSmem), "attr emit")) return false;
  g_attrDone[dev][codec] = true;
  return true;
}

static int enc_enqueue(int codec, const uint8_t *dIn, uint32_t n, uint8_t *dOut, uint32_t cap, void *ws, size_t wsSize, uint32_t *dResult, cudaStream_t st)
{
  Spec sp;
  if (!spec_from_codec(codec, sp) || !dIn || !dOut || !ws || !dResult || n == 0) { g_err = "bad argument"; return 1; }
  if (((uintptr_t)dIn & 15) || ((uintptr_t)dOut & 15) || ((uintptr_t)ws & 255)) { g_err = "device pointers must be 16-byte aligned (workspace 256)"; return 1; }
  const EncKernels *k = enc_kernels_for(codec);
  if (!k) { g_err = "codec not built"; return 1; }
  if (!enc_prepare(codec, k)) return 2;
  EncBufs B; memset(&B, 0, sizeof(B));
  size_t zeroBytes = 0;
  const size_t need = enc_carve(B, sp, n, ws, &zeroBytes);
  if (need > wsSize) { g_err = "workspace too small"; return 1; }
  B.in = dIn; B.out = dOut; B.cap = cap; B.dResult = dResult; B.outBase = (uint32_t)sp.hdr;
  if (!cuda_ok(cudaMemsetAsync(ws, 0, zeroBytes, st), "memset")) return 2;
  const int sms = num_sms();
  HSRLE_LAUNCH_NAMED("k_enc_scan", k->scan, B.nTiles, E1_T, 0, st, B);
  if (k->lutStretch && lut_walk_enabled())
  { // 8-bit LUT codecs: the table at every super-chunk start, from the stretch walk
    const int g = (int)std::min<uint64_t>((uint64_t)B.maxRuns / LW_BLK + 1, (uint64_t)sms * 8);
    HSRLE_LAUNCH_NAMED("k_enc_lut_stretch", k->lutStretch, g, LW_T, 0, st, B);
    HSRLE_LAUNCH_NAMED("k_enc_lut_walk", k->lutWalk, 1, LW_T, 0, st, B);
  }
  const int autoGrid = (int)std::min<uint64_t>((uint64_t)B.maxSC, (uint64_t)sms * 6);
  HSRLE_LAUNCH_NAMED("k_enc_auto", k->autom, autoGrid, E2_T, k->autoSmem, st, B);
  HSRLE_LAUNCH_NAMED("k_enc_fix", k->fix, enc_fix_grid(B, sms), FIX_T, k->fixSmem, st, B, 0, enc_rounds(sp.W, sp.K), enc_follow_hops());
  HSRLE_LAUNCH_NAMED("k_enc_emit", k->emit, autoGrid, E2_T, k->emitSmem, st, B);
  HSRLE_LAUNCH(k_enc_copy_big, sms * 4, 256, 0, st, B);
  return cuda_ok(cudaGetLastError(), "encode launch") ? 0 : 2;
}

// ------------------------------------------------------------------------------------------------
// one stream encoded by several GPUs: per-rank phases between the all-gathers (hsrle_slice.cuh)
static size_t slice_carve(EncBufs &B, const Spec &sp, uint32_t n, uint32_t lo, uint32_t hi, int rank, int world, void *ws, size_t *zeroBytes)
{
  const uint32_t len = hi - lo;
  const size_t base = enc_carve(B, sp, len ? len : 1, ws, zeroBytes, true);     // record / chunk capacities from the slice length
  Carver cv{ (uint8_t *)ws, base };
  B.sliceIn = cv.take<SliceState>(1);
  B.n = n;
  B.sliceMode = 1; B.rank = (uint32_t)rank; B.world = (uint32_t)world; B.sliceLo = lo; B.sliceHi = hi;
  B.vecBase = lo / 16;
  const bool last = rank == world - 1;
  const uint32_t nVecGlobal = (uint32_t)(((uint64_t)n + 1 + 15) / 16);
  const uint32_t vecs = last ? nVecGlobal - B.vecBase : len / 16;
  B.nVec = vecs;
  // a slice's tiles must end exactly where the slice ends (vectors past the cut belong to the next rank: a tile that reaches over it would
  // count their run boundaries): tile sizes that divide the 128-KiB slice alignment only (32 / 64 / 128 KiB)
  // (enc_carve picked among those -- sliceTiles -- and sized the look-back words for at least this many tiles)
  B.nTiles = (vecs + E1_T * B.scanSteps - 1) / (E1_T * B.scanSteps);
  B.lastVec = last ? (n - 1) >> 4 : hi / 16;        // non-last ranks may load the first vector of the tail halo
  B.maxRuns += 2;
  B.outBase = SLICE_PORCH;
  return cv.off + 256;
}

static int slice_phase(const hsrle_slice_job *J, int phase, cudaStream_t st)
{
  Spec sp;
  if (!J || !spec_from_codec(J->codec, sp) || !J->dIn || !J->dOut || !J->dWorkspace || !J->dMsg || !J->dAll || !J->dResult) { g_err = "bad argument"; return 1; }
  if (J->world < 1 || J->world > 64 || J->rank < 0 || J->rank >= J->world || J->n == 0 || J->lo > J->hi || J->hi > J->n) { g_err = "bad slice"; return 1; }
  if ((J->lo % SLICE_ALIGN) || (J->rank != J->world - 1 && (J->hi % SLICE_ALIGN)) || (J->rank == J->world - 1 && J->hi != J->n) || (J->rank == 0 && J->lo != 0))
  { g_err = "slice bounds must be multiples of 128 KiB and cover [0, n)"; return 1; }
  if ((uint64_t)J->n + 64 >= 0xFFFFFFF0ull) { g_err = "stream too long for one frame"; return 1; }
  if (((uintptr_t)J->dIn & 15) || ((uintptr_t)J->dOut & 15) || ((uintptr_t)J->dWorkspace & 255)) { g_err = "device pointers must be 16-byte aligned (workspace 256)"; return 1; }
  const EncKernels *k = enc_kernels_for(J->codec);
  if (!k) { g_err = "codec not built"; return 1; }
  if (!enc_prepare(J->codec, k)) return 2;
  EncBufs B; memset(&B, 0, sizeof(B));
  size_t zeroBytes = 0;
  const size_t need = slice_carve(B, sp, J->n, J->lo, J->hi, J->rank, J->world, J->dWorkspace, &zeroBytes);
  if (need > J->workspaceSize) { g_err = "workspace too small"; return 1; }
  // absolute addressing: byte p of the input is B.in[p]; the rank's buffer starts SLICE_FRONT bytes before its slice
  B.in = J->dIn + SLICE_FRONT - (ptrdiff_t)J->lo;
  B.out = J->dOut; B.cap = J->outCap; B.dResult = J->dResult;
  B.msg = reinterpret_cast<SliceMsg *>(J->dMsg); B.all = reinterpret_cast<const SliceMsg *>(J->dAll);
  const int sms = num_sms();
  const int autoGrid = (int)std::min<uint64_t>((uint64_t)B.maxSC, (uint64_t)sms * 6);
  switch (phase)
  {
    case 0:   // scan
      if (!cuda_ok(cudaMemsetAsync(J->dWorkspace, 0, zeroBytes, st), "memset")) return 2;
      if (B.nTiles) HSRLE_LAUNCH_NAMED("k_enc_scan", k->scan, B.nTiles, E1_T, 0, st, B);
      HSRLE_LAUNCH(k_enc_slice_msg1, 1, 32, 0, st, B);
      break;
    case 1:   // boundary-run fix-up, automaton from the assumed incoming state
      HSRLE_LAUNCH(k_enc_slice_link, 1, 32, 0, st, B, sp);
      HSRLE_LAUNCH_NAMED("k_enc_auto", k->autom, autoGrid, E2_T, k->autoSmem, st, B);
      HSRLE_LAUNCH_NAMED("k_enc_fix", k->fix, enc_fix_grid(B, sms), FIX_T, k->fixSmem, st, B, 0, enc_rounds(sp.W, sp.K), enc_follow_hops());
      break;
    case 2:   // true incoming state, repair rounds
      HSRLE_LAUNCH(k_enc_slice_inject, 1, 32, 0, st, B, sp);
      HSRLE_LAUNCH_NAMED("k_enc_fix", k->fix, 1, FIX_T, k->fixSmem, st, B, 1, enc_rounds(sp.W, sp.K), enc_follow_hops());
      break;
    case 3:   // tokens
      HSRLE_LAUNCH_NAMED("k_enc_emit", k->emit, autoGrid, E2_T, k->emitSmem, st, B);
      HSRLE_LAUNCH(k_enc_copy_big, sms * 4, 256, 0, st, B);
      break;
    case 4:   // closing header, trailing literal, stream header, result
      HSRLE_LAUNCH(k_enc_slice_finish, sms * 2, 256, 0, st, B, sp);
      break;
    default: g_err = "bad phase"; return 1;
  }
  return cuda_ok(cudaGetLastError(), "slice launch") ? 0 : 2;
}

static std::mutex g_dattrMu;
static bool g_dattrDone[MAX_DEV][48];
static int g_emitGrid[MAX_DEV][48];
static bool dec_prepare(int codec, const DecKernels *k)
{
  const int dev = current_dev();
  std::lock_guard<std::mutex> lk(g_dattrMu);
  if (g_dattrDone[dev][codec]) return true;
  cudaFuncSetAttribute((const void *)k->map, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  cudaFuncSetAttribute((const void *)k->emit, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  if (!cuda_ok(cudaFuncSetAttribute((const void *)k->map, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->mapSmem), "attr map")) return false;
  if (!cuda_ok(cudaFuncSetAttribute((const void *)k->emit, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)k->emitSmem), "attr emit")) return false;
  // the emit kernel is persistent: as many CTAs as fit the device at once
  int perSM = 0;
  if (!cuda_ok(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, (const void *)k->emit, DX_T, k->emitSmem), "occupancy emit")) return false;
  g_emitGrid[dev][codec] = std::max(1, perSM) * num_sms();
  g_dattrDone[dev][codec] = true;
  return true;
}

static int dec_enqueue(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize, void *ws, size_t wsSize, uint32_t *dResult, cudaStream_t st)
{
  Spec sp;
  if (!spec_from_codec(codec, sp) || !dIn || !dOut || !ws || !dResult || inSize == 0 || outSize == 0) { g_err = "bad argument"; return 1; }
  if (((uintptr_t)dIn & 15) || ((uintptr_t)dOut & 15) || ((uintptr_t)ws & 255)) { g_err = "device pointers must be 16-byte aligned (workspace 256)"; return 1; }
  const DecKernels *k = dec_kernels_for(codec);
  if (!k) { g_err = "codec not built"; return 1; }
  if (!dec_prepare(codec, k)) return 2;
  DecBufs D; memset(&D, 0, sizeof(D));
  size_t zeroBytes = 0;
  const size_t need = dec_carve(D, sp, inSize, outSize, ws, &zeroBytes);
  if (need > wsSize) { g_err = "workspace too small"; return 1; }
  D.in = dIn; D.out = dOut; D.dResult = dResult;
  D.emitGrid = (uint32_t)g_emitGrid[current_dev()][codec];
  { // tests force either way of composing the segment exits (the result must not depend on it)
    const char *m = getenv("HSRLE_DEC_MODE");
    D.modeOverride = !m ? 0u : (!strcmp(m, "rows") ? 1u : (!strcmp(m, "segtab") ? 2u : 0u));
  }
  if (!cuda_ok(cudaMemsetAsync(ws, 0, zeroBytes, st), "memset")) return 2;
  static const bool dbg = getenv("HSRLE_DEBUG") != nullptr;
  static uint32_t *hDbg = nullptr;
  if (dbg)
  {
    if (!hDbg) cudaHostAlloc((void **)&hDbg, 4096 * 4, cudaHostAllocMapped);
    memset(hDbg, 0, 4096 * 4);
    cudaHostGetDevicePointer((void **)&D.dbg, hDbg, 0);
    for (int i = 3000; i < 3200; i++) hDbg[i] = 0;
  }
  // (CTAs loop over chunks: the scout runs once per CTA, and the chunks it rules out cost a flag each)
  HSRLE_LAUNCH_NAMED("k_dec_map", k->map, std::min<uint32_t>(D.nChunks, (uint32_t)num_sms() * 8u), DM_T, k->mapSmem, st, D);
  if (dbg) { cudaError_t e = cudaStreamSynchronize(st); fprintf(stderr, "[hsrle] k_dec_map done: %s (chunks %u)\n", cudaGetErrorString(e), D.nChunks); fflush(stderr); }
  HSRLE_LAUNCH_NAMED("k_dec_emit", k->emit, D.emitGrid, DX_T, k->emitSmem, st, D);
  if (dbg)
  {
    for (int ms = 0; ms < 8000 && cudaStreamQuery(st) == cudaErrorNotReady; ms += 50) { struct timespec ts = { 0, 50000000 }; nanosleep(&ts, nullptr); }
    if (cudaStreamQuery(st) == cudaErrorNotReady)
    {
      fprintf(stderr, "[hsrle] k_dec_emit STUCK (grid %u); stage markers of the CTAs (stage<<24 | arg), zeros omitted:\n", D.emitGrid);
      std::map<uint32_t, int> hist;
      for (uint32_t b = 0; b < D.emitGrid && b < 4096; b++) { hist[hDbg[b] >> 24]++; if ((hDbg[b] >> 24) != 0xA && hDbg[b]) fprintf(stderr, "  cta %u: %08x\n", b, hDbg[b]); }
      for (auto &kv : hist) fprintf(stderr, "  stage %x: %d CTAs\n", kv.first, kv.second);
      fflush(stderr);
      _exit(3);
    }
    fprintf(stderr, "[hsrle] k_dec_emit done: %s (grid %u); phase ticks/64:", cudaGetErrorString(cudaStreamSynchronize(st)), D.emitGrid);
    for (int i = 0; i < 12; i++) fprintf(stderr, " %u", hDbg[3000 + i]);
    fprintf(stderr, " | K1 resolver ns: chain %u, entries %u, live list %u | sparse composition: max ns %u, segments %u, far parses %u\n", hDbg[3100], hDbg[3101], hDbg[3102], hDbg[3110], hDbg[3111], hDbg[3112]); fflush(stderr);
  }
  return cuda_ok(cudaGetLastError(), "decode launch") ? 0 : 2;
}

// ------------------------------------------------------------------------------------------------
// library-owned context for the synchronous entry points
// One context per call in flight: the reference's entry points are re-entrant pure functions
// (src/simd_platform.c:100-103 is their only shared state), so concurrent callers must not serialise here either;
// each call owns a stream, staging buffers and a workspace and overlaps with the other threads' transfers and kernels.
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
// Contexts live in a pool: a call takes a free one (or creates one) and gives it back when it returns, so threads
// that come and go keep re-using the same streams and device buffers.
static std::mutex g_poolMu;
static std::vector<Context *> g_pool;
struct ContextLease
{
  Context *c;
  ContextLease()
  {
    int cur = -1;
    if (cudaGetDevice(&cur) != cudaSuccess) cur = -1;
    std::lock_guard<std::mutex> lk(g_poolMu);
    c = nullptr;
    for (size_t i = g_pool.size(); i-- > 0;)     // streams and buffers belong to the device they were created on
      if (g_pool[i]->dev == cur || !g_pool[i]->tried) { c = g_pool[i]; g_pool.erase(g_pool.begin() + i); break; }
    if (!c) c = new Context();
  }
  ~ContextLease() { std::lock_guard<std::mutex> lk(g_poolMu); g_pool.push_back(c); }
};

static uint32_t run_sync(Context &C, bool compress, int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize)
{
  Spec sp;
  if (!spec_from_codec(codec, sp)) return 0;
  // size the context's workspace for the hungriest codec at this input size, so that a pooled context never has to
  // re-allocate (cudaFree synchronises the device) when it meets another codec
  size_t need = 0;
  for (int id = 0; id < 48; id++)
  {
    Spec s2; if (!spec_from_codec(id, s2)) continue;
    EncBufs B; DecBufs D;
    need = std::max(need, compress ? enc_carve(B, s2, inSize, nullptr, nullptr) : dec_carve(D, s2, inSize, outSize, nullptr, nullptr));
  }
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
  EncBufs B; return enc_carve(B, sp, inSize, nullptr, nullptr);
}
size_t hsrle_decompress_workspace_size(int codec, uint32_t inSize, uint32_t outSize)
{
  Spec sp; if (!spec_from_codec(codec, sp) || inSize == 0) return 0;
  DecBufs D; return dec_carve(D, sp, inSize, outSize, nullptr, nullptr);
}

int hsrle_compress_device_async(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize,
                                void *dWorkspace, size_t workspaceSize, uint32_t *dResult, void *cudaStream)
{
  return enc_enqueue(codec, dIn, inSize, dOut, outSize, dWorkspace, workspaceSize, dResult, (cudaStream_t)cudaStream);
}
int hsrle_decompress_device_async(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize,
                                  void *dWorkspace, size_t workspaceSize, uint32_t *dResult, void *cudaStream)
{
  return dec_enqueue(codec, dIn, inSize, dOut, outSize, dWorkspace, workspaceSize, dResult, (cudaStream_t)cudaStream);
}

uint32_t hsrle_compress_device(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize)
{
  ContextLease L; Context &C = *L.c;
  if (!C.init()) return 0;
  return run_sync(C, true, codec, dIn, inSize, dOut, outSize);
}
uint32_t hsrle_decompress_device(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize)
{
  ContextLease L; Context &C = *L.c;
  if (!C.init()) return 0;
  return run_sync(C, false, codec, dIn, inSize, dOut, outSize);
}

uint32_t hsrle_compress_host(int codec, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize)
{
  // preconditions of the reference: src/rle8_extreme_cpu.h:88, src/rleX_extreme_cpu.h:49, src/rleX_Xsl.h:271
  if (pIn == NULL || inSize == 0 || pOut == NULL || outSize < rle_compress_bounds(inSize)) return 0;
  ContextLease L; Context &C = *L.c;
  if (!C.init()) return 0;
  if (!C.grow((void **)&C.dIn, &C.dInSize, (size_t)inSize + 64)) return 0;
  if (!C.grow((void **)&C.dOut, &C.dOutSize, (size_t)outSize + 64)) return 0;
  if (!cuda_ok(cudaMemcpyAsync(C.dIn, pIn, inSize, cudaMemcpyHostToDevice, C.stream), "H2D")) return 0;
  const uint32_t r = run_sync(C, true, codec, C.dIn, inSize, C.dOut, outSize);
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
  ContextLease L; Context &C = *L.c;
  if (!C.init()) return 0;
  if (!C.grow((void **)&C.dIn, &C.dInSize, (size_t)clen + 64)) return 0;
  if (!C.grow((void **)&C.dOut, &C.dOutSize, (size_t)n + 64)) return 0;
  if (!cuda_ok(cudaMemcpyAsync(C.dIn, pIn, clen, cudaMemcpyHostToDevice, C.stream), "H2D")) return 0;
  const uint32_t r = run_sync(C, false, codec, C.dIn, clen, C.dOut, n);
  if (r == 0) return 0;
  if (!cuda_ok(cudaMemcpyAsync(pOut, C.dOut, r, cudaMemcpyDeviceToHost, C.stream), "D2H")) return 0;
  if (!cuda_ok(cudaStreamSynchronize(C.stream), "synchronize")) return 0;
  return r;
}

size_t hsrle_slice_workspace_size(int codec, uint32_t sliceBytes)
{
  Spec sp; if (!spec_from_codec(codec, sp)) return 0;
  EncBufs B; return slice_carve(B, sp, sliceBytes ? sliceBytes : 1, 0, sliceBytes, 0, 1, nullptr, nullptr);
}
int hsrle_slice_compress_phase(const hsrle_slice_job *job, int phase, void *cudaStream)
{
  return slice_phase(job, phase, (cudaStream_t)cudaStream);
}

void hsrle_timing_begin(void)
{
  std::lock_guard<std::mutex> lk(g_timedMu);
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
  std::lock_guard<std::mutex> lk(g_timedMu);
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
int hsrle_device(void) { ContextLease L; L.c->init(); return L.c->dev; }
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
