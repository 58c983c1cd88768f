/* Minimal CBLAS declarations for building the UNMODIFIED reference sources as a CPU
 * oracle (TEST INFRASTRUCTURE, never linked into the product). The reference includes
 * <cblas.h> inside extern "C" (distances.hpp:15-17, quantizers.hpp:20-22) and calls only
 * cblas_sgemm(RowMajor, NoTrans, Trans, ...) and cblas_sgemv. They are mapped either to
 * scipy's bundled OpenBLAS (symbols scipy_cblas_*) or to the naive fallbacks that
 * oracle/ref_harness.cpp defines when QADC_REF_NAIVE_BLAS is set. */
#ifndef QADC_SHIM_CBLAS_H
#define QADC_SHIM_CBLAS_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
#ifdef QADC_REF_NAIVE_BLAS
#define QADC_SGEMM qadc_naive_sgemm
#define QADC_SGEMV qadc_naive_sgemv
#else
#define QADC_SGEMM scipy_cblas_sgemm
#define QADC_SGEMV scipy_cblas_sgemv
#endif
void QADC_SGEMM(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                int m, int n, int k, float alpha, const float* a, int lda,
                const float* b, int ldb, float beta, float* c, int ldc);
void QADC_SGEMV(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, int m, int n, float alpha,
                const float* a, int lda, const float* x, int incx, float beta, float* y,
                int incy);
#define cblas_sgemm QADC_SGEMM
#define cblas_sgemv QADC_SGEMV
#endif
