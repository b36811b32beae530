/* hsrle_b200.h -- C ABI of the B200-native extreme-RLE codec (drop-in for the hot path of
 * hypersonic-rle-kit's src/rle.h).
 *
 * Part 1 keeps the reference's entry points VERBATIM (names, argument meaning, return convention:
 * bytes written / 0 on error, src/rle.h:100-394, typedef src/codec_funcs.h:262-266).  They take HOST
 * pointers; the library copies to the GPU, runs the sm_100a kernels and copies back.  There is no CPU
 * path: without a usable CUDA device every call returns 0.
 *
 * Part 2 adds what a GPU-resident caller needs: device-pointer twins (sync and stream-async), a
 * workspace query, and the frame helpers for inputs above the format's u32 ceiling.
 *
 * Stream format and encoder decisions are byte-identical to the reference's AVX2 path.
 */
#ifndef HSRLE_B200_H
#define HSRLE_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---------------------------------------------------------------- Part 1: reference entry points */

/* replaces src/rle8_extreme_cpu.c:22-28 (declared src/rle.h:100): inSize + 193, or 0 above 1 GiB */
uint32_t rle_compress_bounds(const uint32_t inSize);
/* replaces src/rle8_extreme_cpu.c:17-20 (declared src/rle.h:105): 128.  The GPU decoder never writes
 * past uncompressedLength, but callers written against the reference may keep the slack. */
uint32_t rle_decompress_additional_size(void);

/* replaces src/rle.h:101 and src/rle.h:103 */
uint32_t rle8_multi_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle8_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:173 and src/rle.h:175 */
uint32_t rle8_packed_multi_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle8_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:199 and src/rle.h:200 */
uint32_t rle8_3symlut_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle8_3symlut_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:207 and src/rle.h:208 */
uint32_t rle8_7symlut_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle8_7symlut_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:107 and src/rle.h:108 */
uint32_t rle16_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:129 and src/rle.h:130 */
uint32_t rle16_sym_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_sym_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:250 and src/rle.h:251 */
uint32_t rle16_3symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_3symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:255 and src/rle.h:256 */
uint32_t rle16_7symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_7symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:151 and src/rle.h:152 */
uint32_t rle16_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:177 and src/rle.h:178 */
uint32_t rle16_byte_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_byte_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:252 and src/rle.h:253 */
uint32_t rle16_3symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_3symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:257 and src/rle.h:258 */
uint32_t rle16_7symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle16_7symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:118 and src/rle.h:119 */
uint32_t rle24_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:140 and src/rle.h:141 */
uint32_t rle24_sym_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_sym_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:352 and src/rle.h:353 */
uint32_t rle24_3symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_3symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:357 and src/rle.h:358 */
uint32_t rle24_7symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_7symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:162 and src/rle.h:163 */
uint32_t rle24_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:188 and src/rle.h:189 */
uint32_t rle24_byte_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_byte_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:354 and src/rle.h:355 */
uint32_t rle24_3symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_3symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:359 and src/rle.h:360 */
uint32_t rle24_7symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle24_7symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:110 and src/rle.h:111 */
uint32_t rle32_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:132 and src/rle.h:133 */
uint32_t rle32_sym_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_sym_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:284 and src/rle.h:285 */
uint32_t rle32_3symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_3symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:289 and src/rle.h:290 */
uint32_t rle32_7symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_7symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:154 and src/rle.h:155 */
uint32_t rle32_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:180 and src/rle.h:181 */
uint32_t rle32_byte_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_byte_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:286 and src/rle.h:287 */
uint32_t rle32_3symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_3symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:291 and src/rle.h:292 */
uint32_t rle32_7symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle32_7symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:121 and src/rle.h:122 */
uint32_t rle48_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:143 and src/rle.h:144 */
uint32_t rle48_sym_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_sym_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:386 and src/rle.h:387 */
uint32_t rle48_3symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_3symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:391 and src/rle.h:392 */
uint32_t rle48_7symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_7symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:165 and src/rle.h:166 */
uint32_t rle48_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:191 and src/rle.h:192 */
uint32_t rle48_byte_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_byte_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:388 and src/rle.h:389 */
uint32_t rle48_3symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_3symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:393 and src/rle.h:394 */
uint32_t rle48_7symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle48_7symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:113 and src/rle.h:114 */
uint32_t rle64_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:135 and src/rle.h:136 */
uint32_t rle64_sym_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_sym_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:318 and src/rle.h:319 */
uint32_t rle64_3symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_3symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:323 and src/rle.h:324 */
uint32_t rle64_7symlut_sym_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_7symlut_sym_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:157 and src/rle.h:158 */
uint32_t rle64_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:183 and src/rle.h:184 */
uint32_t rle64_byte_packed_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_byte_packed_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:320 and src/rle.h:321 */
uint32_t rle64_3symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_3symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
/* replaces src/rle.h:325 and src/rle.h:326 */
uint32_t rle64_7symlut_byte_compress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);
uint32_t rle64_7symlut_byte_decompress(const uint8_t *pIn, const uint32_t inSize, uint8_t *pOut, const uint32_t outSize);

/* ---------------------------------------------------------------- Part 2: GPU-resident interface */

/* codec id = widthIndex*8 + byteAligned*4 + variant
 *   widthIndex: 0..5 for 8/16/24/32/48/64-bit symbols; byteAligned: 0 = "sym", 1 = "byte" (8-bit: 1)
 *   variant: 0 plain, 1 packed, 2 3symlut, 3 7symlut */
int hsrle_codec_id(int symbolBits, int byteAligned, int variant);
/* id of a reference function name without the _compress/_decompress suffix, e.g. "rle24_3symlut_byte",
 * "rle8_multi", "rle8_packed_multi"; -1 if unknown */
int hsrle_codec_id_from_name(const char *name);

/* Scratch bytes needed by the async calls below for an input of inSize bytes (compress) or a stream of
 * inSize bytes expanding to at most outSize bytes (decompress). */
size_t hsrle_compress_workspace_size(int codec, uint32_t inSize);
size_t hsrle_decompress_workspace_size(int codec, uint32_t inSize, uint32_t outSize);

/* Device-resident, stream-ordered, no host synchronisation.  All pointers are device pointers, 16-byte
 * aligned.  dResult[0] receives the byte count (0 on error), dResult[1] a status code (0 = ok,
 * 1 = output too small, 2 = corrupt stream, 3 = bad argument/header); dResult[2..7] diagnostics.
 * Returns 0 when the work was enqueued, non-zero on a launch/argument error.
 * Readable extent: the kernels read whole 16-byte vectors, so dIn must be readable up to the next 16-byte boundary after
 * dIn + inSize (any cudaMalloc'ed buffer is: allocations are 256-byte granular); the bytes past inSize are never used. */
int hsrle_compress_device_async(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize,
                                void *dWorkspace, size_t workspaceSize, uint32_t *dResult, void *cudaStream);
int hsrle_decompress_device_async(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize,
                                  void *dWorkspace, size_t workspaceSize, uint32_t *dResult, void *cudaStream);

/* Device-resident, synchronous convenience forms (library-owned workspace, returns the byte count).  They run on a
 * library-owned non-blocking stream: work the caller still has in flight on dIn / dOut (on any of its streams) must
 * have completed before the call; on return the result is complete. */
uint32_t hsrle_compress_device(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize);
uint32_t hsrle_decompress_device(int codec, const uint8_t *dIn, uint32_t inSize, uint8_t *dOut, uint32_t outSize);

/* Host-pointer forms by codec id (what the Part 1 functions call). */
uint32_t hsrle_compress_host(int codec, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize);
uint32_t hsrle_decompress_host(int codec, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize);

/* ---- One stream encoded by several GPUs (one process per GPU; configs[3] of the benchmark) ------------------
 * The input of ONE reference-identical stream of n bytes is cut into contiguous slices [lo, hi), one per rank:
 * lo a multiple of 128 KiB, rank 0 starts at 0, the last rank ends at n.  A rank's input buffer dIn holds the
 * bytes [lo - 32, hi + 32) of the input (byte 32 of the buffer is input byte lo; halo bytes outside [0, n) may be
 * anything).  The encode runs in phases with an all-gather of every rank's 256-byte message (dMsg -> dAll, rank
 * order) after phases 0, 1, 2 and 3 -- NCCL in hsrle_b200.sliced, any transport works:
 *   phase 0  candidate scan                                   -> all-gather
 *   phase 1  boundary-run fix-up + emit automaton             -> all-gather
 *   phase 2  incoming-state check + repair rounds             -> all-gather; repeat phase 2 while any rank's
 *            message has word 6 ("changed") set (at most `world` times)
 *   phase 3  tokens                                           -> all-gather
 *   phase 4  closing header + trailing literal + stream header; dResult = { partLen, status, partStart,
 *            partOffset, totalStreamBytes, ... }: this rank's share of the stream is dOut[partStart, partStart +
 *            partLen) and belongs at byte partOffset of the single stream.
 * dOut needs (hi - lo) + (hi - lo) / 256 + 1024 bytes.  All calls are stream-ordered and never synchronise. */
typedef struct hsrle_slice_job
{
  int codec, rank, world;
  uint32_t n, lo, hi;
  const uint8_t *dIn;
  uint8_t *dOut; uint32_t outCap;
  void *dWorkspace; size_t workspaceSize;
  uint32_t *dMsg;            /* 64 words */
  const uint32_t *dAll;      /* world x 64 words */
  uint32_t *dResult;         /* 8 words */
} hsrle_slice_job;
size_t hsrle_slice_workspace_size(int codec, uint32_t sliceBytes);
int hsrle_slice_compress_phase(const hsrle_slice_job *job, int phase, void *cudaStream);

/* Last CUDA error text seen by the library on this thread ("" if none) and the device in use (-1 = none). */
const char *hsrle_last_error(void);
int hsrle_device(void);
/* Number of kernels the library has launched so far in this process (monotonic). */
uint64_t hsrle_kernel_launches(void);
/* Optional per-kernel CUDA-event timing (measurement aid): begin() arms it; end() synchronises the
 * device and writes "kernel:launches:total_ms;..." into buf, returning the characters written. */
void hsrle_timing_begin(void);
int hsrle_timing_end(char *buf, int bufSize);

#ifdef __cplusplus
}
#endif
#endif
