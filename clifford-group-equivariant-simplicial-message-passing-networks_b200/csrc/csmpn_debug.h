/* Bring-up / tuning diagnostics of the tensor-core engine.  NOT part of the product ABI (include/csmpn_b200.h): these
 * entry points exist only in a library built with CSMPN_DEBUG_BUILD=1 (csrc/build.py adds -DCSMPN_DEBUG_TOOLS and
 * compiles tc_probe.cu); tools/tc_probe_*.py, tools/tc_timeline.py and tools/dw_timeline.py are their only callers. */
#ifndef CSMPN_DEBUG_H_
#define CSMPN_DEBUG_H_
#include "../../include/csmpn_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostics: when set to a device buffer of 1024 int64 (NULL to disable), the second forward kernel of engine 1 records
 * (phase code, clock64) pairs of two threads of CTA 0: the per-tile timeline used to tune the pipeline (tools/tc_timeline.py). */
int csmpn_tc_debug_buffer(int64_t* device_buffer_1024);

/* ---- tensor-core diagnostics ------------------------------------------------------------------------------------
 * Single-tile tcgen05 probe (csrc/tc_probe.cu): D = A x B on the TF32 tensor pipe from the shared-memory "plane"
 * operand layout every tensor-core kernel of this library uses; dumps the [128 lanes, N] TMEM accumulator.
 * mode 0: A [128,K], B [N,K] (both K-major); mode 1: A [128,K], B [K,N] (B MN-major); mode 2: A [K,M], B [K,N] (both
 * MN-major, M in {64,128}).  flags: 1 / 2 swap LBO and SBO of A / B (must fail), 4 = hi/lo split with three MMAs. */
int csmpn_tc_probe(int mode, int M, int N, int K, int flags, const float* A, const float* B, float* dump,
                   csmpn_stream_t stream);
/* Raw variant: byte-exact shared-memory images of both operands and explicit descriptor fields (prm16: M, N, a_mn, b_mn,
 * a_lbo, a_sbo, a_layout, b_lbo, b_sbo, b_layout, ksteps, a_kinc, b_kinc, a_off, b_off, 0), for layout exploration. */
int csmpn_tc_probe_raw(const float* a_img, int a_words, const float* b_img, int b_words, const uint32_t* prm16,
                       float* dump, csmpn_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CSMPN_DEBUG_H_ */
