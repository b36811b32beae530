/* TEST INFRASTRUCTURE ONLY -- see rle_oracle.h.  Plain C11 restatement of the reference's extreme
 * RLE codecs as a *run-list specification*: (1) enumerate run candidates, (2) apply the per-run emit
 * rule with its sequential state, (3) serialise tokens.  The SIMD loops of the reference are an
 * implementation detail; what is restated here is their observable behaviour on the AVX2 path
 * (the only ISA-dependent encoder is rle8_packed_multi, src/rle8_extreme_cpu.h:119-123,976-1001).
 *
 * Parity: pinned against the compiled reference (oracle/_ref) by tests/test_oracle_vs_ref.py.
 */
#include "rle_oracle.h"
#include <string.h>

typedef struct
{
  int W, align, variant;
  int hdr;        /* stream header bytes: 9 for rle8 plain/packed (src/rle8_extreme_cpu.c:5-15), else 8 */
  int rng7;       /* 7-bit copy-range field style (PREFER_7_BIT_OR_4_BYTE_COPY) */
  int64_t R;      /* max short-form copy range */
  int64_t SHORT, MEDIUM, LONG;
  int K;          /* LUT entries (3/7) or 0 */
} spec_t;

static int make_spec(spec_t *sp, int W, int align, int variant)
{
  if (!(W == 1 || W == 2 || W == 3 || W == 4 || W == 6 || W == 8)) return 0;
  if (variant < ORC_PLAIN || variant > ORC_LUT7) return 0;
  memset(sp, 0, sizeof(*sp));
  sp->W = W; sp->align = (W == 1) ? ORC_BYTE : align; sp->variant = variant;
  sp->hdr = (W == 1 && (variant == ORC_PLAIN || variant == ORC_PACKED)) ? 9 : 8;
  if (variant == ORC_LUT3 || variant == ORC_LUT7)
  {
    sp->K = variant == ORC_LUT3 ? 3 : 7;
    sp->SHORT = 3;              /* RLE8_XSYMLUT_MIN_RANGE_SHORT, src/rleX_Xsl.h:1 */
    sp->LONG = 2 + 4 + 4 + W;   /* RLE8_XSYMLUT_MIN_RANGE_LONG,  src/rleX_Xsl.h:2 */
    return 1;
  }
  if (W == 1)
  {
    if (variant == ORC_PLAIN) { sp->R = 255; sp->SHORT = 6; sp->LONG = 6; }                    /* src/rle8_extreme_cpu.h:4 */
    else { sp->R = 127; sp->rng7 = 1; sp->SHORT = 3; sp->MEDIUM = 4; sp->LONG = 11; }          /* src/rle8_extreme_cpu.h:14-16 */
    return 1;
  }
  if (variant == ORC_PLAIN)
  { /* src/rleX_extreme_cpu.h:9-11 (FULL_COPY_SIZE 4: plain never uses the 7-bit style) */
    sp->R = 255; sp->SHORT = W + 4; sp->LONG = W + 11;
  }
  else if (sp->align == ORC_BYTE)
  { /* byte_packed is compiled with PREFER_7_BIT_OR_4_BYTE_COPY, src/rleX_extreme_cpu.c:28-44 */
    sp->R = 127; sp->rng7 = 1; sp->SHORT = 3; sp->MEDIUM = W + 3; sp->LONG = W + 11;
  }
  else
  { /* sym_packed: the macro is #undef'd again first, src/rleX_extreme_cpu.c:46-63 */
    sp->R = 255; sp->SHORT = 3; sp->MEDIUM = W + 3; sp->LONG = W + 10;
  }
  return 1;
}

/* ---------------------------------------------------------------- little-endian field helpers */
static void put16(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); }
static void put32(uint8_t *p, uint32_t v) { p[0] = (uint8_t)v; p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)(v >> 16); p[3] = (uint8_t)(v >> 24); }
static uint32_t get16(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8); }
static uint32_t get32(const uint8_t *p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

/* ---------------------------------------------------------------- candidates */
/* W==1: maximal byte runs of length >= 2 (scanner: src/rle8_extreme_cpu.h:950-1091).
 * W>1 : first k >= cursor with word(k)==word(k+W) and k+2W <= n, extended by whole symbols, then
 *       (byte-aligned only) by up to W-1 leading pattern bytes if a whole word still fits
 *       (src/rleX_extreme_cpu_encode.h:57-166,315-380). */
static int next_candidate(const uint8_t *in, int64_t n, const spec_t *sp, int64_t cursor, int64_t *ps, int64_t *pe)
{
  const int W = sp->W;
  if (W == 1)
  {
    int64_t k = cursor;
    while (k + 1 < n && in[k] != in[k + 1]) k++;
    if (k + 1 >= n) return 0;
    int64_t e = k + 2;
    while (e < n && in[e] == in[k]) e++;
    *ps = k; *pe = e;
    return 1;
  }
  int64_t k = cursor;
  while (k + 2 * W <= n && memcmp(in + k, in + k + W, (size_t)W) != 0) k++;
  if (k + 2 * W > n) return 0;
  int64_t e = k + 2 * W;
  while (e + W <= n && memcmp(in + e, in + k, (size_t)W) == 0) e += W;
  if (sp->align == ORC_BYTE && e + W <= n)
  {
    int j = 0;
    while (j < W - 1 && in[e + j] == in[k + j]) j++;
    e += j;
  }
  *ps = k; *pe = e;
  return 1;
}

/* ---------------------------------------------------------------- encoder */
typedef struct
{
  const spec_t *sp;
  const uint8_t *in; int64_t n;
  uint8_t *out; int64_t cap; int64_t idx; int overflow;
  int64_t last;                 /* lastRLE */
  uint8_t lastSym[8];           /* packed: symbol of the previous emitted token (starts all-zero) */
  uint8_t lut[7][8];            /* LUT variants: move-to-front list */
} enc_t;

static void emit_bytes(enc_t *E, const void *p, int64_t len)
{
  if (E->idx + len > E->cap) { E->overflow = 1; return; }
  memcpy(E->out + E->idx, p, (size_t)len);
  E->idx += len;
}
static void emit8(enc_t *E, uint32_t v) { uint8_t b = (uint8_t)v; emit_bytes(E, &b, 1); }
static void emit16(enc_t *E, uint32_t v) { uint8_t b[2]; put16(b, v); emit_bytes(E, b, 2); }
static void emit32(enc_t *E, uint32_t v) { uint8_t b[4]; put32(b, v); emit_bytes(E, b, 4); }

static void emit_rng(enc_t *E, int64_t rng, int forceLong)
{
  const spec_t *sp = E->sp;
  if (sp->rng7)
  { /* src/rle8_extreme_cpu.h:1042-1052 */
    if (rng <= 127 && !forceLong) emit8(E, (uint32_t)(rng << 1));
    else emit32(E, ((uint32_t)rng << 1) | 1u);
  }
  else
  { /* src/rle8_extreme_cpu.h:1029-1040 */
    if (rng <= 255 && !forceLong) emit8(E, (uint32_t)rng);
    else { emit8(E, 0); emit32(E, (uint32_t)rng); }
  }
}

/* plain / packed token: returns 1 if emitted.  Rules: SURVEY App. B.1/B.3. */
static int eval_plain_packed(enc_t *E, int64_t s, int64_t e)
{
  const spec_t *sp = E->sp;
  const int W = sp->W;
  const int64_t cnt = e - s;
  const int64_t rng = s - E->last + 1;
  const uint8_t *sym = E->in + s;
  int emit, same = 0;

  if (sp->variant == ORC_PLAIN)
  {
    if (W == 1) emit = cnt >= 6;                                                   /* src/rle8_extreme_cpu.h:974 */
    else emit = (rng <= sp->R && cnt >= sp->SHORT) || cnt >= sp->LONG;             /* src/rleX_extreme_cpu_encode.h:177,239 */
  }
  else
  {
    int simdRegion = 1;
    if (W == 1)
    { /* AVX2 path: the run is evaluated inside the 32-byte loop iff the block that detects its end
         starts before n-32 (src/rle8_extreme_cpu.h:950-978); otherwise the scalar tail rule applies
         (src/rle8_extreme_cpu.h:119-159,206-246). */
      const int64_t p = s + 1 + 32 * ((e - s - 1) / 32);
      simdRegion = (p < E->n - 32) && (e < E->n);
    }
    same = memcmp(sym, E->lastSym, (size_t)W) == 0;
    if (simdRegion)
      emit = cnt >= sp->LONG || (rng <= sp->R && ((same && cnt >= sp->SHORT) || cnt >= sp->MEDIUM));
    else
    { emit = cnt >= sp->LONG; same = 0; }
    if (emit && simdRegion) memcpy(E->lastSym, sym, (size_t)W);
  }
  if (!emit) return 0;

  int64_t stored;
  if (W == 1) stored = cnt - sp->SHORT + 1;
  else if (sp->align == ORC_BYTE) stored = cnt - sp->SHORT + 1;
  else stored = cnt / W - sp->SHORT / W + 1;

  if (sp->variant == ORC_PLAIN)
  {
    emit_bytes(E, sym, W);
    if (stored <= 255) emit8(E, (uint32_t)stored); else { emit8(E, 0); emit32(E, (uint32_t)stored); }
  }
  else
  {
    const uint32_t sameMask = same ? 0x80u : 0u;
    if (stored <= 127) emit8(E, (uint32_t)stored | sameMask); else { emit8(E, sameMask); emit32(E, (uint32_t)stored); }
    if (!same) emit_bytes(E, sym, W);
  }
  emit_rng(E, rng, 0);
  emit_bytes(E, E->in + E->last, s - E->last);
  E->last = e;
  return 1;
}

/* LUT token (`process_symbol`, src/rleX_Xsl.h:114-264): returns 1 if emitted. */
static int eval_lut(enc_t *E, int64_t s, int64_t e)
{
  const spec_t *sp = E->sp;
  const int W = sp->W, K = sp->K;
  const int RB = (K == 3) ? 7 : 6;              /* RLE8_XSYMLUT_RANGE_BITS, src/rleX_Xsl.h:4-15 */
  const int64_t TR = (1 << RB) - 1, TC = 127;
  const int64_t cnt = e - s;
  const int64_t rng = s - E->last + 2;
  const uint8_t *sym = E->in + s;
  int idx = 0;
  for (; idx < K; idx++) if (memcmp(sym, E->lut[idx], (size_t)W) == 0) break;

  int64_t stored;
  if (W == 1 || sp->align == ORC_BYTE) stored = cnt - 3 + 2;
  else stored = cnt / W - 3 / W + 2;

  const int64_t pen = (rng <= 0xFFFFF ? (rng <= TR ? 0 : 2) : 4) + (stored <= 0xFFFFF ? (stored <= TC ? 0 : 2) : 4) + (idx == K ? 1 : 0);
  if (!(cnt >= sp->LONG || cnt >= 3 + pen)) return 0;

  /* move-to-front (src/rleX_Xsl.h:135-188) */
  if (idx != 0)
  {
    const int from = idx == K ? K - 1 : idx;
    for (int j = from; j > 0; j--) memcpy(E->lut[j], E->lut[j - 1], 8);
    memset(E->lut[0], 0, 8);
    memcpy(E->lut[0], sym, (size_t)W);
  }

  const uint32_t c7 = stored <= TC ? (uint32_t)stored : (stored <= 0xFFFF ? 1u : 0u);
  const uint32_t r7 = rng <= TR ? (uint32_t)rng : (rng <= 0xFFFF ? 1u : 0u);
  const uint32_t head = ((uint32_t)idx << (K == 3 ? 14 : 13)) | (c7 << RB) | r7;
  emit16(E, head);
  if (idx == K) emit_bytes(E, sym, W);
  if (stored != (int64_t)c7) { if (stored <= 0xFFFF) emit16(E, (uint32_t)stored); else emit32(E, (uint32_t)stored); }
  if (rng != (int64_t)r7) { if (rng <= 0xFFFF) emit16(E, (uint32_t)rng); else emit32(E, (uint32_t)rng); }
  emit_bytes(E, E->in + E->last, s - E->last);
  E->last = e;
  return 1;
}

uint32_t oracle_compress_bounds(uint32_t inSize)
{
  if (inSize > (1u << 30)) return 0;
  return inSize + (16 + 4 + 1 + 4 + 1 + 64) * 2 + 12 + 1;
}

uint32_t oracle_decompress_additional_size(void) { return 128; }

uint32_t oracle_compress(int W, int align, int variant, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize)
{
  spec_t sp;
  if (!make_spec(&sp, W, align, variant)) return 0;
  if (pIn == NULL || inSize == 0 || pOut == NULL || outSize < oracle_compress_bounds(inSize)) return 0;

  enc_t E;
  memset(&E, 0, sizeof(E));
  E.sp = &sp; E.in = pIn; E.n = inSize; E.out = pOut; E.cap = outSize;
  if (sp.K)
  { /* LUT init list, each byte broadcast to W bytes (src/rleX_Xsl.h:279-287) */
    static const uint8_t init7[7] = { 0x00, 0x7F, 0xFF, 0x01, 0x7E, 0x80, 0xFE };
    for (int k = 0; k < sp.K; k++) memset(E.lut[k], init7[k], (size_t)W);
  }

  uint8_t hdr[9] = { 0 };
  put32(hdr, inSize);
  emit_bytes(&E, hdr, sp.hdr);   /* compressedLength patched below; mode byte 0 = multi */

  int64_t cursor = 0, s, e;
  while (next_candidate(pIn, E.n, &sp, cursor, &s, &e))
  {
    if (sp.K) eval_lut(&E, s, e); else eval_plain_packed(&E, s, e);
    cursor = e;
  }

  const int64_t L = E.n - E.last;
  static const uint8_t zeros[8] = { 0 };
  if (sp.K)
  { /* src/rleX_Xsl.h:316-340 */
    const int RB = (sp.K == 3) ? 7 : 6;
    if (L == 0) { emit16(&E, (1u << RB) | 1u); emit16(&E, 0); emit16(&E, 0); }
    else { emit16(&E, 1u << RB); emit16(&E, 0); emit32(&E, (uint32_t)(L + 2)); emit_bytes(&E, pIn + E.last, L); }
  }
  else
  { /* src/rle8_extreme_cpu.h:281-337, src/rleX_extreme_cpu_encode.h:454-601 */
    if (sp.variant == ORC_PLAIN) { emit_bytes(&E, zeros, W); emit8(&E, 0); emit32(&E, 0); }
    else { emit8(&E, 0x80); emit32(&E, 0); }
    if (L == 0)
    {
      if (sp.rng7) emit32(&E, 1); else { emit8(&E, 0); emit32(&E, 0); }
    }
    else
    {
      emit_rng(&E, L + 1, 1);
      emit_bytes(&E, pIn + E.last, L);
    }
  }
  if (E.overflow) return 0;
  put32(pOut + 4, (uint32_t)E.idx);
  return (uint32_t)E.idx;
}

/* ---------------------------------------------------------------- decoder */
static void fill_run(uint8_t *out, int64_t len, const uint8_t *sym, int W)
{
  for (int64_t i = 0; i < len; i++) out[i] = sym[i % W];
}

uint32_t oracle_decompress(int W, int align, int variant, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize)
{
  spec_t sp;
  if (!make_spec(&sp, W, align, variant)) return 0;
  if (pIn == NULL || pOut == NULL || inSize == 0 || outSize == 0) return 0;
  if (inSize < (uint32_t)sp.hdr) return 0;
  const uint32_t n = get32(pIn), clen = get32(pIn + 4);
  if (n > outSize || clen > inSize) return 0;     /* src/rle8_extreme_cpu.h:707-712 */

  int64_t ip = sp.hdr, op = 0;
  const int64_t iend = clen;
  int single = 0;
  uint8_t sym[8] = { 0 };
  uint8_t lut[7][8];

  if (sp.hdr == 9)
  {
    const uint8_t mode = pIn[8];
    if (mode == 1) { single = 1; sym[0] = pIn[ip++]; }   /* src/rle8_extreme_cpu.h:736-757 */
    else if (mode != 0) return 0;
  }
  if (sp.K)
  {
    static const uint8_t init7[7] = { 0x00, 0x7F, 0xFF, 0x01, 0x7E, 0x80, 0xFE };
    for (int k = 0; k < sp.K; k++) { memset(lut[k], 0, 8); memset(lut[k], init7[k], (size_t)W); }
  }

#define NEED(k) do { if (ip + (k) > iend) return 0; } while (0)
  for (;;)
  {
    int64_t cnt, rng, runBytes;
    if (sp.K)
    { /* src/rleX_Xsl.h:580-784 */
      const int K = sp.K, RB = (K == 3) ? 7 : 6;
      NEED(2);
      const uint32_t head = get16(pIn + ip); ip += 2;
      const int idx = (int)(head >> (K == 3 ? 14 : 13));
      cnt = (head >> RB) & 0x7F;
      rng = head & ((1u << RB) - 1);
      if (idx == K)
      {
        NEED(W);
        for (int j = K - 1; j > 0; j--) memcpy(lut[j], lut[j - 1], 8);
        memset(lut[0], 0, 8); memcpy(lut[0], pIn + ip, (size_t)W); ip += W;
      }
      else if (idx > 0)
      {
        uint8_t t[8]; memcpy(t, lut[idx], 8);
        for (int j = idx; j > 0; j--) memcpy(lut[j], lut[j - 1], 8);
        memcpy(lut[0], t, 8);
      }
      if (cnt == 1) { NEED(2); cnt = get16(pIn + ip); ip += 2; }
      else if (cnt == 0) { NEED(4); cnt = get32(pIn + ip); ip += 4; }
      if (rng == 1) { NEED(2); rng = get16(pIn + ip); ip += 2; if (rng == 0) break; }
      else if (rng == 0) { NEED(4); rng = get32(pIn + ip); ip += 4; }
      rng -= 2;
      memcpy(sym, lut[0], 8);
      if (W == 1 || sp.align == ORC_BYTE) runBytes = cnt + 1; else runBytes = (cnt + 3 / W - 2) * W;
      if (cnt == 0) runBytes = 0;
    }
    else
    {
      if (single)
      { /* src/rle8_extreme_cpu.h:2020-2087 */
        NEED(1); cnt = pIn[ip++];
        if (cnt == 0) { NEED(4); cnt = get32(pIn + ip); ip += 4; }
      }
      else if (sp.variant == ORC_PLAIN)
      {
        NEED(W + 1);
        memcpy(sym, pIn + ip, (size_t)W); ip += W;
        cnt = pIn[ip++];
        if (cnt == 0) { NEED(4); cnt = get32(pIn + ip); ip += 4; }
      }
      else
      {
        NEED(1);
        const uint8_t b0 = pIn[ip++];
        cnt = b0 & 0x7F;
        if (cnt == 0) { NEED(4); cnt = get32(pIn + ip); ip += 4; }
        if (!(b0 & 0x80)) { NEED(W); memcpy(sym, pIn + ip, (size_t)W); ip += W; }
      }
      if (sp.rng7 && !single)
      {
        NEED(1);
        if (pIn[ip] & 1) { NEED(4); rng = get32(pIn + ip) >> 1; ip += 4; if (rng == 0) break; }
        else { rng = pIn[ip++] >> 1; }
      }
      else
      {
        NEED(1); rng = pIn[ip++];
        if (rng == 0) { NEED(4); rng = get32(pIn + ip); ip += 4; if (rng == 0) break; }
      }
      rng -= 1;
      if (single) runBytes = cnt + (sp.variant == ORC_PLAIN ? 3 : 1);
      else if (W == 1) runBytes = cnt + sp.SHORT - 1;
      else if (sp.align == ORC_BYTE) runBytes = cnt + sp.SHORT - 1;
      else runBytes = (cnt + sp.SHORT / W - 1) * W;
      if (cnt == 0) runBytes = 0;
    }
    if (rng < 0) return 0;
    NEED(rng);
    if (op + rng > (int64_t)n) return 0;
    memcpy(pOut + op, pIn + ip, (size_t)rng); ip += rng; op += rng;
    if (cnt == 0) break;
    if (op + runBytes > (int64_t)n) return 0;
    fill_run(pOut + op, runBytes, sym, W); op += runBytes;
  }
#undef NEED
  if (op != (int64_t)n) return 0;
  return n;
}
