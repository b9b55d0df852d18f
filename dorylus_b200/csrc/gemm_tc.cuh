// tcgen05 (5th-gen tensor core) path for the dense apply Z = AH . W  (internal).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace dory {

// C[M x ldc] = A[M x lda] . W[Kpad x ldw]  with an error-compensated 3xTF32 split (fp32-level
// accuracy); epilogue EPI_TANH also writes C2 = tanh(C).
// Returns the number of kernels launched, 0 when the shape is not supported by this kernel (the
// caller then uses the fp32 SIMT path), or -1 on a launch error.
int launch_gemm_tc(const float *A, uint32_t lda, uint64_t M, const float *W, uint32_t ldw, uint32_t Kpad,
                   float *C, float *C2, uint32_t ldc, int epilogue, cudaStream_t s);

// dW[Mpad x ldc] = A^T . G with A [K x lda] (first Mpad columns) and G [K x ldg], both row-major: the
// contraction runs over the rows (vertices).  3xTF32 on tcgen05 with MN-major operands, split over
// the vertex range into `ws` partials that are reduced in ascending order (deterministic).
// Same return convention as launch_gemm_tc.
int launch_gemm_tn_tc(const float *A, uint32_t lda, uint32_t Mpad, const float *G, uint32_t ldg, uint64_t K, float *C,
                      uint32_t ldc, float *ws, size_t ws_floats, cudaStream_t s);

}  // namespace dory
