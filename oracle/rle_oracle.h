/* TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the hypersonic-rle-kit extreme RLE
 * hot path.  Nothing in the product (hypersonic-rle-kit_b200/, include/) may include, link or call
 * this.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
 * use it, and only as the checker.
 *
 * Parity status: PINNED by differential testing against the compiled, unmodified reference
 * (oracle/_ref/libhsrle_ref.so, built by oracle/Makefile from /root/reference/src) -- see
 * tests/test_oracle_vs_ref.py and tests/golden/.  The reference ships no golden vectors of its own
 * (its tests are round-trip only, src/rle_fuzz.c:609-744), so the committed fixtures under
 * tests/golden/ were generated from the compiled reference by tests/golden/make_golden.py.
 */
#ifndef RLE_ORACLE_H
#define RLE_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

enum { ORC_PLAIN = 0, ORC_PACKED = 1, ORC_LUT3 = 2, ORC_LUT7 = 3 };
enum { ORC_SYM = 0, ORC_BYTE = 1 };

/* W: symbol width in bytes (1,2,3,4,6,8).  align: ORC_SYM / ORC_BYTE (ignored for W==1).
 * Returns bytes written (0 on error), like the reference's *_compress (src/rle.h:100-394).
 * Unlike the reference, out-of-bounds reads never happen: a W-byte word compare that would touch an
 * index >= inSize fails (SURVEY App. C.1 convention; the reference reads past the buffer there,
 * src/rleX_extreme_cpu_encode.h:369-371). */
uint32_t oracle_compress(int W, int align, int variant, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize);

/* Decodes any stream of the given codec family; for W==1 PLAIN/PACKED also mode-1 (single) streams
 * (src/rle8_extreme_cpu.h:702-764).  Writes exactly uncompressedLength bytes (no overshoot). */
uint32_t oracle_decompress(int W, int align, int variant, const uint8_t *pIn, uint32_t inSize, uint8_t *pOut, uint32_t outSize);

uint32_t oracle_compress_bounds(uint32_t inSize);      /* src/rle8_extreme_cpu.c:22-28 */
uint32_t oracle_decompress_additional_size(void);      /* src/rle8_extreme_cpu.c:17-20 */

#ifdef __cplusplus
}
#endif
#endif
