/* Stub for <cblas.h>: the reference's Utils/memory.h:779,814 call cblas_dnrm2 / cblas_snrm2 in two dead helpers.
 * No CBLAS is installed in this image; these two inline definitions let the reference header compile unmodified.
 * Used only by oracle/ref_harness.cu (test infrastructure). */
#ifndef XM_B200_STUB_CBLAS_H
#define XM_B200_STUB_CBLAS_H
#include <math.h>
static inline double cblas_dnrm2(const int n, const double* x, const int incx) {
    double s = 0; for (int i = 0; i < n; ++i) s += x[i * incx] * x[i * incx]; return sqrt(s);
}
static inline float cblas_snrm2(const int n, const float* x, const int incx) {
    float s = 0; for (int i = 0; i < n; ++i) s += x[i * incx] * x[i * incx]; return sqrtf(s);
}
#endif
