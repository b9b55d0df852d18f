// tcgen05 (5th-gen tensor core) path for the dense apply Z = AH . W  (internal).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

namespace dory {

// C[M x ldc] = A[M x lda] . W[Kpad x ldw]  with an error-compensated 3xTF32 split (fp32-level
// accuracy); epilogue EPI_TANH also writes C2 = tanh(C).
// Returns the number of kernels launched, 0 when the shape is not supported by this kernel (the
// caller then uses the fp32 SIMT path), or -1 on a launch error.
// K <= 128 with N = 32 / 64 takes the small-tile kernel (several CTAs per SM, `small_stages` shared-memory stages
// each, 0 = choose); small_stages < 0 keeps the deep-ring kernel (N = 64 only).
int launch_gemm_tc(const float *A, uint32_t lda, uint64_t M, const float *W, uint32_t ldw, uint32_t Kpad,
                   float *C, float *C2, uint32_t ldc, int epilogue, cudaStream_t s, int small_stages = 0);

// C[M x prows] = G[M x ldw] . W^T (W [prows x ldw] row-major): grad = d . W^T of ApplyVertex backward
// (CPU_comm.cpp:152-157) on the same small-tile kernel.  Same return convention.
int launch_gemm_nt_tc(const float *G, uint32_t ldg, uint64_t M, const float *W, uint32_t ldw, uint32_t prows, float *C,
                      uint32_t ldc, int stages, cudaStream_t s);

// Last-layer apply, fused: logits = A[V x lda] . W[Kpad x ldw] (3xTF32) and, in the epilogue (a TMEM lane is a
// vertex row), everything launch_softmax_ce does except the final sum of the per-row statistics
// (launch_softmax_stats).  `stages` = shared-memory stages per CTA, 0 = choose.  Same return convention.
struct SoftmaxCEArgs;
int launch_gemm_tc_softmax(const float *A, uint32_t lda, const float *W, uint32_t ldw, uint32_t Kpad, const SoftmaxCEArgs &a,
                           int stages, cudaStream_t s);

// dW[Mpad x ldc] = A^T . G with A [K x lda] (first Mpad columns) and G [K x ldg], both row-major: the
// contraction runs over the rows (vertices).  3xTF32 on tcgen05 with MN-major operands, split over
// the vertex range into `ws` partials that are reduced in ascending order (deterministic).
// Same return convention as launch_gemm_tc.
int launch_gemm_tn_tc(const float *A, uint32_t lda, uint32_t Mpad, const float *G, uint32_t ldg, uint64_t K, float *C,
                      uint32_t ldc, float *ws, size_t ws_floats, cudaStream_t s);

}  // namespace dory
