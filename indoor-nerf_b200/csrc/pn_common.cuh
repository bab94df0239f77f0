// Common plumbing for libpocketnerf.so (sm_100a).  No torch types anywhere in csrc/.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/pocketnerf.h"

#if defined(__CUDACC__)
#define PN_HD __host__ __device__ __forceinline__
#else
#define PN_HD inline
#endif

// Individually rounded fp32 operations.  The reference runs eager torch: every *, +, -, / is its
// own kernel and is rounded on its own, so wherever the result must match bit for bit the kernels
// must not let ptxas contract a*b+c into an FMA.  On the host (test-only emulation build, compiled
// with -ffp-contract=off) the plain operators have the same semantics.
#if defined(__CUDA_ARCH__)
PN_HD float pn_mul(float a, float b) { return __fmul_rn(a, b); }
PN_HD float pn_add(float a, float b) { return __fadd_rn(a, b); }
PN_HD float pn_sub(float a, float b) { return __fsub_rn(a, b); }
PN_HD float pn_div(float a, float b) { return __fdiv_rn(a, b); }
#else
PN_HD float pn_mul(float a, float b) { volatile float r = a * b; return r; }
PN_HD float pn_add(float a, float b) { volatile float r = a + b; return r; }
PN_HD float pn_sub(float a, float b) { volatile float r = a - b; return r; }
PN_HD float pn_div(float a, float b) { volatile float r = a / b; return r; }
#endif

namespace pn {

// ---- error reporting (thread-local message, no exceptions across the ABI) ------------------------
void set_error(const char *fmt, ...);
int check_launch(const char *what);   // cudaGetLastError() after a launch -> 0 / PN_ECUDA
void count_launch(int n = 1);

#define PN_REQUIRE(cond, code, ...)          \
  do {                                       \
    if (!(cond)) {                           \
      pn::set_error(__VA_ARGS__);            \
      return (code);                         \
    }                                        \
  } while (0)

static inline cudaStream_t as_stream(pn_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// Number of SMs of the current device (cached per device).
int sm_count();

static inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pn
