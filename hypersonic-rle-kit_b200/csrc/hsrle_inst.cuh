// hsrle_inst.cuh -- instantiates every codec of one symbol width (HSRLE_INST_W) and exports its kernel table.
#include "hsrle_dispatch.h"
#include <type_traits>
#include "hsrle_enc_kernels.cuh"
#include "hsrle_dec_kernels.cuh"

namespace hsrle {

template <int W, int BA, int V> static EncKernels make_enc_kernels()
{
  using SymT = typename std::conditional<(W <= 4), uint32_t, uint64_t>::type;
  constexpr Spec sp = make_spec(W, BA, V);
  using C = EncCta<W, BA, V, SymT>;
  EncKernels k;
  k.scan = &k_enc_scan<W, sp.minM, SymT>;
  k.autom = &k_enc_auto<W, BA, V, SymT>;
  k.fix = &k_enc_fix<W, BA, V, SymT>;
  k.emit = &k_enc_emit<W, BA, V, SymT>;
  k.lutStretch = nullptr; k.lutWalk = nullptr;
  if constexpr (W == 1 && sp.K != 0) { k.lutStretch = &k_enc_lut_stretch<V>; k.lutWalk = &k_enc_lut_walk<V>; }
  k.autoSmem = sizeof(typename C::Smem);
  k.fixSmem = sizeof(typename C::FixSmem);
  k.emitSmem = sizeof(EncEmitSmem<W, BA, V, SymT>);
  k.symBytes = (int)sizeof(SymT);
  k.minM = sp.minM;
  return k;
}

template <int W, int BA, int V> static DecKernels make_dec_kernels()
{
  constexpr Spec sp = make_spec(W, BA, V);
  DecKernels k;
  k.map = &k_dec_map<W, BA, V>;
  k.emit = &k_dec_emit<W, BA, V>;
  k.mapSmem = sizeof(DecMapSmem);
  k.emitSmem = sizeof(DecEmitSmem<sp.K>);
  k.aggBytes = sizeof(DecAgg<sp.K>);
  return k;
}

#define HSRLE_CAT2(a, b) a##b
#define HSRLE_CAT(a, b) HSRLE_CAT2(a, b)

const EncKernels *HSRLE_CAT(enc_kernels_w, HSRLE_INST_W)()
{
  static EncKernels tab[8];
  static bool init = false;
  if (!init)
  {
    for (int i = 0; i < 8; i++) tab[i] = EncKernels{ nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, 0, 0, 0, 0, 0 };
    constexpr int W = HSRLE_INST_W;
    if constexpr (W > 1)
    {
      tab[0] = make_enc_kernels<W, 0, V_PLAIN>(); tab[1] = make_enc_kernels<W, 0, V_PACKED>();
      tab[2] = make_enc_kernels<W, 0, V_LUT3>(); tab[3] = make_enc_kernels<W, 0, V_LUT7>();
    }
    tab[4] = make_enc_kernels<W, 1, V_PLAIN>(); tab[5] = make_enc_kernels<W, 1, V_PACKED>();
    tab[6] = make_enc_kernels<W, 1, V_LUT3>(); tab[7] = make_enc_kernels<W, 1, V_LUT7>();
    init = true;
  }
  return tab;
}

const DecKernels *HSRLE_CAT(dec_kernels_w, HSRLE_INST_W)()
{
  static DecKernels tab[8];
  static bool init = false;
  if (!init)
  {
    for (int i = 0; i < 8; i++) tab[i] = DecKernels{ nullptr, nullptr, 0, 0, 0 };
    constexpr int W = HSRLE_INST_W;
    if constexpr (W > 1)
    {
      tab[0] = make_dec_kernels<W, 0, V_PLAIN>(); tab[1] = make_dec_kernels<W, 0, V_PACKED>();
      tab[2] = make_dec_kernels<W, 0, V_LUT3>(); tab[3] = make_dec_kernels<W, 0, V_LUT7>();
    }
    tab[4] = make_dec_kernels<W, 1, V_PLAIN>(); tab[5] = make_dec_kernels<W, 1, V_PACKED>();
    tab[6] = make_dec_kernels<W, 1, V_LUT3>(); tab[7] = make_dec_kernels<W, 1, V_LUT7>();
    init = true;
  }
  return tab;
}

} // namespace hsrle
