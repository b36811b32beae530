// hsrle_core.cuh -- codec rules shared by every kernel of the B200 extreme-RLE pipeline.
//
// Everything here is `__host__ __device__` so that the exact stage logic the kernels run can also be
// driven by a host-side stage simulator in tests/sim (test tool only, never shipped in the product
// library -- the product has no CPU path).
//
// Reference behaviour restated here (never copied; see SURVEY.md App. A/B for the derivation):
//   emit rules / token layout, plain+packed : src/rle8_extreme_cpu.h:950-1060, src/rleX_extreme_cpu_encode.h:172-312
//   emit rule / token layout, 3LUT/7LUT     : src/rleX_Xsl.h:114-264
//   token parse                             : src/rleX_extreme_cpu_decode.h:43-163, src/rleX_Xsl.h:580-784,
//                                             src/rle8_extreme_cpu.h:1558-1632,2020-2087
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define HSRLE_HD __host__ __device__ __forceinline__
#define HSRLE_HDC constexpr __host__ __device__ __forceinline__
#else
#define HSRLE_HD inline
#define HSRLE_HDC constexpr inline
#endif

#ifdef __CUDA_ARCH__
#define HSRLE_UNROLL _Pragma("unroll")
#else
#define HSRLE_UNROLL
#endif

namespace hsrle {

enum : int { V_PLAIN = 0, V_PACKED = 1, V_LUT3 = 2, V_LUT7 = 3 };

// One codec family member: symbol width W (bytes), alignment, token variant.
struct Spec
{
  int W;          // 1,2,3,4,6,8
  int byteAlign;  // 1 = "byte" variants (run length any byte count), 0 = "sym" (multiple of W)
  int variant;    // V_*
  int K;          // LUT entries (0,3,7)
  int hdr;        // stream header bytes (9 for rle8 plain/packed, else 8)
  int rng7;       // 7-bit copy-range field style
  int R;          // max short-form copy range
  int SHORT, MEDIUM, LONG;
  int minM;       // shortest run of the match mask M[p] = (in[p]==in[p-W]) that can yield a candidate
  int RB;         // LUT: range bits in the u16 head (7 / 6)
};

HSRLE_HDC Spec make_spec(int W, int byteAlign, int variant)
{
  Spec sp{};
  sp.W = W; sp.byteAlign = (W == 1) ? 1 : byteAlign; sp.variant = variant;
  sp.K = variant == V_LUT3 ? 3 : (variant == V_LUT7 ? 7 : 0);
  sp.hdr = (W == 1 && (variant == V_PLAIN || variant == V_PACKED)) ? 9 : 8;
  sp.rng7 = 0; sp.R = 255; sp.SHORT = 0; sp.MEDIUM = 0; sp.LONG = 0; sp.RB = variant == V_LUT3 ? 7 : 6;
  if (sp.K)
  {
    sp.SHORT = 3; sp.LONG = 2 + 4 + 4 + W; sp.R = (1 << sp.RB) - 1;
  }
  else if (W == 1)
  {
    if (variant == V_PLAIN) { sp.SHORT = 6; sp.LONG = 6; }
    else { sp.R = 127; sp.rng7 = 1; sp.SHORT = 3; sp.MEDIUM = 4; sp.LONG = 11; }
  }
  else if (variant == V_PLAIN) { sp.SHORT = W + 4; sp.LONG = W + 11; }
  else if (sp.byteAlign) { sp.R = 127; sp.rng7 = 1; sp.SHORT = 3; sp.MEDIUM = W + 3; sp.LONG = W + 11; }
  else { sp.SHORT = 3; sp.MEDIUM = W + 3; sp.LONG = W + 10; }
  // W==1: a byte run of length L is an M-run of length L-1; runs below the smallest emittable length
  // never change any state, so the scanner drops them.  W>1: a candidate needs W consecutive M bits.
  sp.minM = (W == 1) ? ((variant == V_PLAIN) ? 5 : 2) : W;
  return sp;
}

// codec id used across the C ABI: id = widthIndex*8 + byteAlign*4 + variant, widthIndex over {1,2,3,4,6,8}
HSRLE_HD int width_from_index(int wi) { return wi == 0 ? 1 : wi == 1 ? 2 : wi == 2 ? 3 : wi == 3 ? 4 : wi == 4 ? 6 : 8; }

// ------------------------------------------------------------------------------------------------
// little-endian unaligned field access
HSRLE_HD uint64_t load_sym(const uint8_t *p, int W)
{
  uint64_t v = 0;
  for (int i = 0; i < W; i++) v |= (uint64_t)p[i] << (8 * i);
  return v;
}
HSRLE_HD uint32_t load16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
HSRLE_HD uint32_t load32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

struct TokenHdr
{
  uint32_t len;
  uint8_t b[20];
  HSRLE_HD void put8(uint32_t v) { b[len++] = (uint8_t)v; }
  HSRLE_HD void put16(uint32_t v) { put8(v); put8(v >> 8); }
  HSRLE_HD void put32(uint32_t v) { put8(v); put8(v >> 8); put8(v >> 16); put8(v >> 24); }
  HSRLE_HD void putsym(uint64_t s, int W) { for (int i = 0; i < W; i++) put8((uint32_t)(s >> (8 * i))); }
};

// Header byte sinks for enc_eval: CountSink only measures, TokenHdr buffers, PtrSink writes through.
struct CountSink
{
  uint32_t len;
  HSRLE_HD void put8(uint32_t) { len++; }
  HSRLE_HD void put16(uint32_t) { len += 2; }
  HSRLE_HD void put32(uint32_t) { len += 4; }
  HSRLE_HD void putsym(uint64_t, int W) { len += W; }
};
struct PtrSink
{
  uint32_t len;
  uint8_t *p;
  HSRLE_HD void put8(uint32_t v) { p[len++] = (uint8_t)v; }
  HSRLE_HD void put16(uint32_t v) { put8(v); put8(v >> 8); }
  HSRLE_HD void put32(uint32_t v) { put8(v); put8(v >> 8); put8(v >> 16); put8(v >> 24); }
  HSRLE_HD void putsym(uint64_t s, int W) { for (int i = 0; i < W; i++) put8((uint32_t)(s >> (8 * i))); }
};

// The W bytes at offset d (0 <= d < W) of the period-W pattern whose first period is sym0.
HSRLE_HD uint64_t sym_rot(uint64_t sym0, int W, uint32_t d)
{
  if (W == 1 || d == 0) return sym0;
  const int sh = 8 * (int)d;
  uint64_t r = (sym0 >> sh) | (sym0 << (8 * W - sh));
  if (W < 8) r &= (1ull << (8 * W)) - 1ull;
  return r;
}

// ------------------------------------------------------------------------------------------------
// encoder automaton state
struct AutoState
{
  uint32_t cursor;    // end of the previous valid candidate (W>1 candidate search restarts here)
  uint32_t last;      // lastRLE: input index just after the previous emitted run
  uint64_t lastSym;   // packed: symbol of the previous emitted token
};
HSRLE_HD bool operator==(const AutoState &a, const AutoState &b) { return a.cursor == b.cursor && a.last == b.last && a.lastSym == b.lastSym; }
HSRLE_HD bool operator!=(const AutoState &a, const AutoState &b) { return !(a == b); }

struct Lut
{
  uint64_t s[7];
};
HSRLE_HD bool lut_equal(const Lut &a, const Lut &b, int K)
{
  bool eq = true;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++) if (i < K) eq = eq && (a.s[i] == b.s[i]);
  return eq;
}

HSRLE_HD uint64_t broadcast_byte(uint32_t v, int W)
{
  uint64_t r = 0;
  for (int i = 0; i < W; i++) r |= (uint64_t)v << (8 * i);
  return r;
}

HSRLE_HD void lut_init(Lut &l, int W)
{ // src/rleX_Xsl.h:279-287 / src/rleX_Xsl_multibyte_encoder.h:31-39
  l.s[0] = broadcast_byte(0x00, W); l.s[1] = broadcast_byte(0x7F, W); l.s[2] = broadcast_byte(0xFF, W);
  l.s[3] = broadcast_byte(0x01, W); l.s[4] = broadcast_byte(0x7E, W); l.s[5] = broadcast_byte(0x80, W);
  l.s[6] = broadcast_byte(0xFE, W);
}

// All LUT helpers index their arrays with compile-time constants only (fully unrolled, predicated), so that
// the tables live in registers on the device; K is a compile-time constant at every call site.
HSRLE_HD int lut_find(const Lut &l, int K, uint64_t sym)
{
  int idx = K;
  HSRLE_UNROLL
  for (int i = 6; i >= 0; i--) if (i < K && l.s[i] == sym) idx = i;
  return idx;
}

// move-to-front: idx<K moves entry idx (== sym) to the front; idx==K pushes a new symbol (dropping the last entry)
HSRLE_HD void lut_touch(Lut &l, int K, int idx, uint64_t sym)
{
  const int from = idx == K ? K - 1 : idx;
  HSRLE_UNROLL
  for (int j = 6; j > 0; j--) if (j < K && j <= from) l.s[j] = l.s[j - 1];
  l.s[0] = sym;
}
// entry idx (< K) of the table
HSRLE_HD uint64_t lut_get(const Lut &l, int K, int idx)
{
  uint64_t v = l.s[0];
  HSRLE_UNROLL
  for (int i = 1; i < 7; i++) if (i < K && i == idx) v = l.s[i];
  return v;
}

// "K most recent distinct emitted symbols" aggregate of a segment; composition is associative.
struct LutAgg
{
  uint32_t m;
  uint64_t s[7];
};
HSRLE_HD void lutagg_push(LutAgg &a, int K, uint64_t sym)
{
  int idx = (int)a.m;
  HSRLE_UNROLL
  for (int i = 6; i >= 0; i--) if (i < K && i < (int)a.m && a.s[i] == sym) idx = i;
  if (idx == (int)a.m) { if ((int)a.m < K) a.m++; else idx = K - 1; }
  HSRLE_UNROLL
  for (int j = 6; j > 0; j--) if (j < K && j <= idx) a.s[j] = a.s[j - 1];
  a.s[0] = sym;
}
// apply a segment aggregate (newer) on top of a full LUT (older): newer symbols first, then the
// older entries not among them, truncated to K
HSRLE_HD void lut_apply(Lut &l, int K, const LutAgg &a)
{
  if (a.m == 0) return;
  Lut r;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++) r.s[i] = a.s[i];
  int k = (int)a.m;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++)
  {
    if (i < K)
    {
      bool dup = false;
      HSRLE_UNROLL
      for (int j = 0; j < 7; j++) if (j < K) dup |= (j < (int)a.m && a.s[j] == l.s[i]);
      if (!dup && k < K)
      {
        HSRLE_UNROLL
        for (int q = 0; q < 7; q++) if (q < K && q == k) r.s[q] = l.s[i];
        k++;
      }
    }
  }
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++) if (i < K) l.s[i] = r.s[i];
}
// older then newer -> combined aggregate
HSRLE_HD LutAgg lutagg_combine(const LutAgg &older, const LutAgg &newer, int K)
{
  LutAgg r = newer;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++)
  {
    if (i < K && i < (int)older.m)
    {
      bool dup = false;
      HSRLE_UNROLL
      for (int j = 0; j < 7; j++) if (j < K) dup |= (j < (int)newer.m && newer.s[j] == older.s[i]);
      if (!dup && (int)r.m < K)
      {
        HSRLE_UNROLL
        for (int q = 0; q < 7; q++) if (q < K && q == (int)r.m) r.s[q] = older.s[i];
        r.m++;
      }
    }
  }
  return r;
}

// ------------------------------------------------------------------------------------------------
// W == 1: table entries are single bytes, so the whole K-entry table is ONE 64-bit word (entry i in byte i) and every
// table operation is a handful of integer instructions instead of 7-way unrolled 64-bit select chains.  Same semantics
// as Lut / LutAgg above (the kernels pick the representation per symbol width; global-memory records keep the generic
// layout and are converted at the boundaries).
struct LutB { uint64_t v; };
struct LutAggB { uint32_t m; uint32_t pad; uint64_t v; };
HSRLE_HD uint64_t lutb_mask(int cnt) { return cnt >= 8 ? ~0ull : ((1ull << (8 * cnt)) - 1ull); }
HSRLE_HD int lutb_ctz(uint64_t z)
{
#ifdef __CUDA_ARCH__
  return __ffsll((long long)z) - 1;
#else
  return __builtin_ctzll(z);
#endif
}
// index of byte `sym` among the low `cnt` bytes of v (cnt if absent); the lowest zero byte of the xor is found exactly
HSRLE_HD int lutb_find(uint64_t v, int cnt, uint32_t sym)
{
  const uint64_t x = v ^ (0x0101010101010101ull * (uint64_t)(sym & 0xFFu));
  const uint64_t z = (x - 0x0101010101010101ull) & ~x & 0x8080808080808080ull & lutb_mask(cnt);
  return z ? (lutb_ctz(z) >> 3) : cnt;
}
// bytes below `from` move up by one, bytes above it stay, `sym` becomes byte 0
HSRLE_HD uint64_t lutb_front(uint64_t v, int from, uint32_t sym)
{
  const uint64_t low = lutb_mask(from), high = ~lutb_mask(from + 1);
  return ((v & low) << 8) | (v & high) | (uint64_t)(sym & 0xFFu);
}
HSRLE_HD void lut_init(LutB &l, int) { l.v = 0x00FE807E01FF7F00ull; }      // bytes 00 7F FF 01 7E 80 FE (src/rleX_Xsl.h:279-287)
HSRLE_HD bool lut_equal(const LutB &a, const LutB &b, int K) { return ((a.v ^ b.v) & lutb_mask(K)) == 0; }
HSRLE_HD int lut_find(const LutB &l, int K, uint64_t sym) { return lutb_find(l.v, K, (uint32_t)sym); }
HSRLE_HD void lut_touch(LutB &l, int K, int idx, uint64_t sym) { l.v = lutb_front(l.v, idx == K ? K - 1 : idx, (uint32_t)sym) & lutb_mask(K); }
HSRLE_HD uint64_t lut_front(const LutB &l) { return l.v & 0xFFu; }
HSRLE_HD uint64_t lut_front(const Lut &l) { return l.s[0]; }
HSRLE_HD void lutagg_push(LutAggB &a, int K, uint64_t sym)
{
  int idx = lutb_find(a.v, (int)a.m, (uint32_t)sym);
  if (idx == (int)a.m) { if ((int)a.m < K) a.m++; else idx = K - 1; }
  a.v = lutb_front(a.v, idx, (uint32_t)sym) & lutb_mask(K);
}
HSRLE_HD void lut_apply(LutB &l, int K, const LutAggB &a)
{
  if (a.m == 0) return;
  uint64_t r = a.v & lutb_mask((int)a.m);
  int k = (int)a.m;
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++)
  {
    if (i < K)
    {
      const uint32_t b = (uint32_t)(l.v >> (8 * i)) & 0xFFu;
      if (k < K && lutb_find(a.v, (int)a.m, b) == (int)a.m) { r |= (uint64_t)b << (8 * k); k++; }
    }
  }
  l.v = r;
}
HSRLE_HD LutAggB lutagg_combine(const LutAggB &older, const LutAggB &newer, int K)
{
  LutAggB r; r.pad = 0; r.m = newer.m; r.v = newer.v & lutb_mask((int)newer.m);
  HSRLE_UNROLL
  for (int i = 0; i < 7; i++)
  {
    if (i < K && i < (int)older.m)
    {
      const uint32_t b = (uint32_t)(older.v >> (8 * i)) & 0xFFu;
      if ((int)r.m < K && lutb_find(newer.v, (int)newer.m, b) == (int)newer.m) { r.v |= (uint64_t)b << (8 * r.m); r.m++; }
    }
  }
  return r;
}
// conversions (global-memory records use the generic layout)
HSRLE_HD void lut_from(LutB &d, const Lut &s) { d.v = 0; HSRLE_UNROLL for (int i = 0; i < 7; i++) d.v |= (s.s[i] & 0xFFull) << (8 * i); }
HSRLE_HD void lut_from(Lut &d, const Lut &s) { d = s; }
HSRLE_HD void lut_to(Lut &d, const LutB &s) { HSRLE_UNROLL for (int i = 0; i < 7; i++) d.s[i] = (s.v >> (8 * i)) & 0xFFull; }
HSRLE_HD void lut_to(Lut &d, const Lut &s) { d = s; }
HSRLE_HD void lutagg_from(LutAggB &d, const LutAgg &s) { d.m = s.m; d.pad = 0; d.v = 0; HSRLE_UNROLL for (int i = 0; i < 7; i++) d.v |= (s.s[i] & 0xFFull) << (8 * i); }
HSRLE_HD void lutagg_from(LutAgg &d, const LutAgg &s) { d = s; }
HSRLE_HD void lutagg_to(LutAgg &d, const LutAggB &s) { d.m = s.m; HSRLE_UNROLL for (int i = 0; i < 7; i++) d.s[i] = (s.v >> (8 * i)) & 0xFFull; }
HSRLE_HD void lutagg_to(LutAgg &d, const LutAgg &s) { d = s; }
HSRLE_HD uint64_t lut_entry(const LutB &l, int i) { return (l.v >> (8 * i)) & 0xFFull; }
HSRLE_HD uint64_t lut_entry(const Lut &l, int i)
{
  uint64_t v = l.s[0];
  HSRLE_UNROLL
  for (int k = 1; k < 7; k++) if (k == i) v = l.s[k];
  return v;
}
HSRLE_HD uint64_t lutagg_entry(const LutAggB &a, int i) { return (a.v >> (8 * i)) & 0xFFull; }
HSRLE_HD uint64_t lutagg_entry(const LutAgg &a, int i)
{
  uint64_t v = a.s[0];
  HSRLE_UNROLL
  for (int k = 1; k < 7; k++) if (k == i) v = a.s[k];
  return v;
}
template <int W> struct LutRep { using L = Lut; using A = LutAgg; };
template <> struct LutRep<1> { using L = LutB; using A = LutAggB; };

// ------------------------------------------------------------------------------------------------
// Encoder: evaluate one match-mask run [a,b) (M[p]==1 for a<=p<b, maximal, b-a >= sp.minM).
// Returns EV_* flags.  With EV_EMIT, [s,e) is the run, `h` its header bytes (everything before the
// literal), and the literal is in[lastBefore, s).  State is advanced either way.
enum : uint32_t { EV_VALID = 1, EV_EMIT = 2, EV_SYMSET = 4, EV_STATE_MASK = 7,
                  EV_MARG = 8,           // LUT codecs: the decision would flip between "symbol in the table" and "not in the table"
                  EV_IDX_SHIFT = 8 };    // LUT codecs: table index of the symbol (K = absent) in bits 8..10
// `sym0` = the W input bytes in[a-W, a) (the first period of the run the mask run [a,b) belongs to).
template <class Sink, class LutT, class AggT>
HSRLE_HD uint32_t enc_eval_t(const Spec &sp, uint64_t sym0, uint32_t n, uint32_t a, uint32_t b, AutoState &st, LutT &lut, AggT *agg,
                         uint32_t &s, uint32_t &e, Sink &h)
{
  const int W = sp.W;
  if (W == 1) { s = a - 1; e = b; }
  else
  { // s = max(cursor, a-W); valid iff b >= s+2W; whole symbols, then (byte) partial
    s = a - W; if (st.cursor > s) s = st.cursor;
    if ((uint64_t)b < (uint64_t)s + 2 * W) return 0;
    const uint32_t estar = s + ((b - s) / W) * W;
    e = (sp.byteAlign && (uint64_t)estar + W <= n) ? b : estar;
    st.cursor = e;
  }
  const uint32_t cnt = e - s;
  const uint64_t sym = sym_rot(sym0, W, s - (a - W));
  h.len = 0;

  if (sp.K)
  { // process_symbol, src/rleX_Xsl.h:114-264
    const int K = sp.K;
    const uint32_t TR = (1u << sp.RB) - 1, TC = 127;
    const uint32_t rng = s - st.last + 2;
    const int idx = lut_find(lut, K, sym);
    const uint32_t stored = (W == 1 || sp.byteAlign) ? cnt - 1 : cnt / W - 3 / W + 2;
    const uint32_t penBase = (rng <= 0xFFFFFu ? (rng <= TR ? 0u : 2u) : 4u) + (stored <= 0xFFFFFu ? (stored <= TC ? 0u : 2u) : 4u);
    const uint32_t pen = penBase + (idx == K ? 1u : 0u);
    const uint32_t info = ((uint32_t)idx << EV_IDX_SHIFT) | ((cnt < (uint32_t)sp.LONG && cnt == 3 + penBase) ? (uint32_t)EV_MARG : 0u);
    if (!(cnt >= (uint32_t)sp.LONG || cnt >= 3 + pen)) return EV_VALID | info;
    lut_touch(lut, K, idx, sym);
    if (agg) lutagg_push(*agg, K, sym);
    const uint32_t c7 = stored <= TC ? stored : (stored <= 0xFFFFu ? 1u : 0u);
    const uint32_t r7 = rng <= TR ? rng : (rng <= 0xFFFFu ? 1u : 0u);
    h.put16(((uint32_t)idx << (K == 3 ? 14 : 13)) | (c7 << sp.RB) | r7);
    if (idx == K) h.putsym(sym, W);
    if (stored != c7) { if (stored <= 0xFFFFu) h.put16(stored); else h.put32(stored); }
    if (rng != r7) { if (rng <= 0xFFFFu) h.put16(rng); else h.put32(rng); }
    st.last = e;
    return EV_VALID | EV_EMIT | info;
  }

  const uint32_t rng = s - st.last + 1;
  bool emit, same = false, symSet = false;
  if (sp.variant == V_PLAIN)
  {
    if (W == 1) emit = cnt >= 6;
    else emit = (rng <= (uint32_t)sp.R && cnt >= (uint32_t)sp.SHORT) || cnt >= (uint32_t)sp.LONG;
  }
  else
  {
    bool simdRegion = true;
    if (W == 1)
    { // AVX2 contract: evaluated by the 32-byte loop iff the block that sees the run end starts before n-32
      const uint64_t p = (uint64_t)s + 1 + 32ull * ((e - s - 1) / 32);
      simdRegion = (p + 32 < n) && (e < n);
    }
    same = (sym == st.lastSym);
    if (simdRegion)
    {
      emit = cnt >= (uint32_t)sp.LONG || (rng <= (uint32_t)sp.R && ((same && cnt >= (uint32_t)sp.SHORT) || cnt >= (uint32_t)sp.MEDIUM));
      if (emit) { st.lastSym = sym; symSet = true; }
    }
    else { emit = cnt >= (uint32_t)sp.LONG; same = false; }
  }
  if (!emit) return EV_VALID;

  uint32_t stored;
  if (W == 1 || sp.byteAlign) stored = cnt - sp.SHORT + 1;
  else stored = cnt / W - sp.SHORT / W + 1;
  if (sp.variant == V_PLAIN)
  {
    h.putsym(sym, W);
    if (stored <= 255) h.put8(stored); else { h.put8(0); h.put32(stored); }
  }
  else
  {
    const uint32_t sameMask = same ? 0x80u : 0u;
    if (stored <= 127) h.put8(stored | sameMask); else { h.put8(sameMask); h.put32(stored); }
    if (!same) h.putsym(sym, W);
  }
  if (sp.rng7) { if (rng <= 127) h.put8(rng << 1); else h.put32((rng << 1) | 1u); }
  else { if (rng <= 255) h.put8(rng); else { h.put8(0); h.put32(rng); } }
  st.last = e;
  return EV_VALID | EV_EMIT | (symSet ? EV_SYMSET : 0u);
}
template <class Sink>
HSRLE_HD uint32_t enc_eval(const Spec &sp, uint64_t sym0, uint32_t n, uint32_t a, uint32_t b, AutoState &st, Lut &lut, LutAgg *agg,
                       uint32_t &s, uint32_t &e, Sink &h)
{
  return enc_eval_t<Sink, Lut, LutAgg>(sp, sym0, n, a, b, st, lut, agg, s, e, h);
}

// Terminator written after the last token; L = trailing literal length (n - last).
HSRLE_HD void enc_terminator(const Spec &sp, uint32_t L, TokenHdr &h)
{
  h.len = 0;
  if (sp.K)
  {
    if (L == 0) { h.put16((1u << sp.RB) | 1u); h.put16(0); h.put16(0); }
    else { h.put16(1u << sp.RB); h.put16(0); h.put32(L + 2); }
    return;
  }
  if (sp.variant == V_PLAIN) { h.putsym(0, sp.W); h.put8(0); h.put32(0); }
  else { h.put8(0x80); h.put32(0); }
  if (L == 0) { if (sp.rng7) h.put32(1); else { h.put8(0); h.put32(0); } }
  else if (sp.rng7) h.put32(((L + 1) << 1) | 1u);
  else { h.put8(0); h.put32(L + 1); }
}

// ------------------------------------------------------------------------------------------------
// Decoder: parse the token that starts at p (avail = readable stream bytes from p).
struct Tok
{
  uint32_t hdrLen;    // bytes before the literal
  uint32_t litLen;
  uint32_t runLen;    // output bytes of the run part (0 for the final token)
  int symKind;        // 0: explicit symbol at p+symOff | 1: same as previous | 2+idx: LUT entry idx (idx<K) | -1: none
  uint32_t symOff;
  bool last;          // decoding stops after this token's literal
  bool valid;
};

// byte readers: plain pointer, or anything with u8(offset)
struct PtrReader
{
  const uint8_t *p;
  HSRLE_HD uint32_t u8(uint32_t o) const { return p[o]; }
};
template <class Rd> HSRLE_HD uint32_t rd16(const Rd &r, uint32_t o) { return r.u8(o) | (r.u8(o + 1) << 8); }
template <class Rd> HSRLE_HD uint32_t rd32(const Rd &r, uint32_t o) { return r.u8(o) | (r.u8(o + 1) << 8) | (r.u8(o + 2) << 16) | (r.u8(o + 3) << 24); }
template <class Rd> HSRLE_HD uint64_t rd_sym(const Rd &r, uint32_t o, int W)
{
  uint64_t v = 0;
  for (int i = 0; i < W; i++) v |= (uint64_t)r.u8(o + i) << (8 * i);
  return v;
}

template <class Rd>
HSRLE_HD void dec_parse_rd(const Spec &sp, bool single, const Rd &p, uint64_t avail, Tok &t)
{
  const int W = sp.W;
  uint32_t ip = 0, cnt, rng;
  t.valid = false; t.last = false; t.symKind = -1; t.symOff = 0; t.hdrLen = 0; t.litLen = 0; t.runLen = 0;
#define HSRLE_NEED(k) do { if ((uint64_t)ip + (k) > avail) return; } while (0)
  if (sp.K)
  {
    const int K = sp.K;
    HSRLE_NEED(2);
    const uint32_t head = rd16(p, 0); ip = 2;
    const int idx = (int)(head >> (K == 3 ? 14 : 13));
    cnt = (head >> sp.RB) & 0x7F;
    rng = head & ((1u << sp.RB) - 1);
    if (idx == K) { HSRLE_NEED(W); t.symKind = 0; t.symOff = ip; ip += W; }
    else t.symKind = 2 + idx;
    if (cnt == 1) { HSRLE_NEED(2); cnt = rd16(p, ip); ip += 2; }
    else if (cnt == 0) { HSRLE_NEED(4); cnt = rd32(p, ip); ip += 4; }
    if (rng == 1) { HSRLE_NEED(2); rng = rd16(p, ip); ip += 2; if (rng == 0) { t.last = true; t.hdrLen = ip; t.valid = true; return; } }
    else if (rng == 0) { HSRLE_NEED(4); rng = rd32(p, ip); ip += 4; }
    if (rng < 2) return;
    t.litLen = rng - 2;
    if (cnt == 0) t.last = true;
    else if (W == 1 || sp.byteAlign) t.runLen = cnt + 1;
    else t.runLen = (cnt + 3 / W - 2) * W;
  }
  else
  {
    if (single)
    {
      HSRLE_NEED(1); cnt = p.u8(ip++);
      if (cnt == 0) { HSRLE_NEED(4); cnt = rd32(p, ip); ip += 4; }
      t.symKind = 1;
    }
    else if (sp.variant == V_PLAIN)
    {
      HSRLE_NEED(W + 1);
      t.symKind = 0; t.symOff = 0; ip = W;
      cnt = p.u8(ip++);
      if (cnt == 0) { HSRLE_NEED(4); cnt = rd32(p, ip); ip += 4; }
    }
    else
    {
      HSRLE_NEED(1);
      const uint32_t b0 = p.u8(ip++);
      cnt = b0 & 0x7F;
      if (cnt == 0) { HSRLE_NEED(4); cnt = rd32(p, ip); ip += 4; }
      if (!(b0 & 0x80)) { HSRLE_NEED(W); t.symKind = 0; t.symOff = ip; ip += W; }
      else t.symKind = 1;
    }
    if (sp.rng7 && !single)
    {
      HSRLE_NEED(1);
      if (p.u8(ip) & 1) { HSRLE_NEED(4); rng = rd32(p, ip) >> 1; ip += 4; if (rng == 0) { t.last = true; t.hdrLen = ip; t.valid = true; return; } }
      else { rng = p.u8(ip++) >> 1; }
    }
    else
    {
      HSRLE_NEED(1); rng = p.u8(ip++);
      if (rng == 0) { HSRLE_NEED(4); rng = rd32(p, ip); ip += 4; if (rng == 0) { t.last = true; t.hdrLen = ip; t.valid = true; return; } }
    }
    if (rng < 1) return;
    t.litLen = rng - 1;
    if (cnt == 0) t.last = true;
    else if (single) t.runLen = cnt + (sp.variant == V_PLAIN ? 3 : 1);
    else if (W == 1 || sp.byteAlign) t.runLen = cnt + sp.SHORT - 1;
    else t.runLen = (cnt + sp.SHORT / W - 1) * W;
  }
#undef HSRLE_NEED
  t.hdrLen = ip;
  if ((uint64_t)ip + t.litLen > avail) return;
  t.valid = true;
}
HSRLE_HD void dec_parse(const Spec &sp, bool single, const uint8_t *p, uint64_t avail, Tok &t)
{
  PtrReader r; r.p = p;
  dec_parse_rd(sp, single, r, avail, t);
}

} // namespace hsrle
