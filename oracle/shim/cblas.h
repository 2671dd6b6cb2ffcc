/* Minimal CBLAS declaration used only to compile the reference's BLAS back-end
 * (src/blas/blas_conv_layer.c, blas_dense_layer.c call nothing but cblas_sgemm).
 * TEST INFRASTRUCTURE ONLY - part of the oracle build, never of the product. */
#ifndef ORACLE_SHIM_CBLAS_H
#define ORACLE_SHIM_CBLAS_H
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_sgemm(enum CBLAS_ORDER order, enum CBLAS_TRANSPOSE ta, enum CBLAS_TRANSPOSE tb,
                 int M, int N, int K, float alpha, const float *A, int lda,
                 const float *B, int ldb, float beta, float *C, int ldc);
#endif
