/* TEST INFRASTRUCTURE ONLY.  Tiny shim linked into oracle/_ref/libhsrle_ref.so next to the
 * unmodified reference objects: lets the test harness steer the reference's ISA dispatch the same
 * way `hsrlekit --max-simd` does (src/main.c:172-313 just clears the simd_platform.h globals). */
#include "simd_platform.h"
#include <stdint.h>

/* level: 0 = host default, 1 = cap at AVX2 (no AVX-512), 2 = cap at SSE4.2 (no AVX) */
void hsrle_ref_set_max_simd(int level)
{
  _DetectCPUFeatures();
  if (level >= 1) { avx512FSupported = false; avx512PFSupported = false; avx512ERSupported = false; avx512CDSupported = false;
                    avx512BWSupported = false; avx512DQSupported = false; avx512VLSupported = false; avx512IFMASupported = false;
                    avx512VBMISupported = false; avx512VNNISupported = false; avx512VBMI2Supported = false; avx512POPCNTDQSupported = false;
                    avx512BITALGSupported = false; avx5124VNNIWSupported = false; avx5124FMAPSSupported = false; }
  if (level >= 2) { avx2Supported = false; avxSupported = false; fma3Supported = false; }
}
int hsrle_ref_has_avx2(void) { _DetectCPUFeatures(); return avx2Supported ? 1 : 0; }
int hsrle_ref_has_avx512f(void) { _DetectCPUFeatures(); return avx512FSupported ? 1 : 0; }
