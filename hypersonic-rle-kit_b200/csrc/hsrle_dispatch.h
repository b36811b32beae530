// hsrle_dispatch.h -- per-codec kernel tables.  Every symbol width is instantiated in its own translation
// unit (hsrle_inst_w*.cu) so the 44 codec specialisations compile in parallel.
#pragma once
#include <cuda_runtime.h>
#include "hsrle_enc.cuh"
#include "hsrle_dec.cuh"

namespace hsrle {

struct EncKernels
{
  void (*scan)(const EncBufs);
  void (*autom)(const EncBufs);
  void (*fix)(const EncBufs, int, int, int);
  void (*emit)(const EncBufs);
  void (*lutStretch)(const EncBufs);   // 8-bit LUT codecs only (hsrle_enc_lutwalk.cuh), else null
  void (*lutWalk)(const EncBufs);
  size_t autoSmem, fixSmem, emitSmem;
  int symBytes;           // 4 or 8: element size of EncBufs::runSym
  int minM;
};

struct DecKernels
{
  void (*map)(const DecBufs);
  void (*emit)(const DecBufs);
  size_t mapSmem, emitSmem;
  size_t aggBytes;        // sizeof(DecAgg<K>)
};

// table index = byteAlign*4 + variant (see hsrle_codec_id); entries with scan == nullptr do not exist
const EncKernels *enc_kernels_w1();
const EncKernels *enc_kernels_w2();
const EncKernels *enc_kernels_w3();
const EncKernels *enc_kernels_w4();
const EncKernels *enc_kernels_w6();
const EncKernels *enc_kernels_w8();
const DecKernels *dec_kernels_w1();
const DecKernels *dec_kernels_w2();
const DecKernels *dec_kernels_w3();
const DecKernels *dec_kernels_w4();
const DecKernels *dec_kernels_w6();
const DecKernels *dec_kernels_w8();

} // namespace hsrle
