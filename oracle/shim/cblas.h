/* Test infrastructure only (see oracle/README.md).
 * Minimal CBLAS declaration so that the reference's src/common/matrix.cpp
 * (which does `#include "cblas.h"`, matrix.hpp:8) compiles here without a
 * system OpenBLAS: cblas_sgemm is mapped onto the LP64 `scipy_cblas_sgemm`
 * exported by the OpenBLAS that ships inside scipy.libs. */
#pragma once
#ifdef __cplusplus
extern "C" {
#endif
typedef enum { CblasRowMajor = 101, CblasColMajor = 102 } CBLAS_ORDER;
typedef enum { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 } CBLAS_TRANSPOSE;
void scipy_cblas_sgemm(CBLAS_ORDER, CBLAS_TRANSPOSE, CBLAS_TRANSPOSE, int, int, int, float,
                       const float *, int, const float *, int, float, float *, int);
int scipy_openblas_get_num_threads(void);
void scipy_openblas_set_num_threads(int);
#ifdef __cplusplus
}
#endif
#define cblas_sgemm scipy_cblas_sgemm
