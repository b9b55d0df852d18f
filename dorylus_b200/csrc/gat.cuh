// GAT edge operators (internal).  Reference: CPUComm::edgNNForwardGAT / edgNNBackwardGAT
// (commmanager/CPU_comm.cpp:190-242) and Engine::predictGAT (engine/ops/gat_ops.cpp:247-265).
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <cstdint>

namespace dory {

// az[e] = z[dst(e)] . a ; A[e] = leakyReLU(az[e], 0.01) for every in-edge e (CSC order).
int launch_gat_edge_forward(const float *z, uint32_t ld, uint32_t F, const float *a, const uint64_t *colPtrs,
                            uint32_t V, float *az, float *A, cudaStream_t s);

struct GatEdgeBackwardArgs {
    const float *grad;        // [V x ld]
    const float *z;           // [V x ld]
    uint32_t ld, F, V;
    const float *a;           // [F] (a_i, stored F x 1 with ld 4)
    const float *az;          // [E]
    const uint64_t *colPtrs;  // [V+1]
    float *dA;                // [E]   out
    float *da;                // [F x 1, ld 4] out (gradient handed to sendaUpdate)
    float *scratch;           // >= 2*V + 2*ld floats
    size_t scratch_floats;
    float *ws;                // split workspace for the F x F / column reductions
    size_t ws_floats;
};
int launch_gat_edge_backward(const GatEdgeBackwardArgs &a, cudaStream_t s);

// grad[r,:] = softmax(logits[r,:]) - lab[r,:] for r in [low, up).
int launch_gat_predict(const float *logits, uint32_t ld_logits, const float *lab, float *grad, uint32_t ld,
                       uint32_t C, uint32_t low, uint32_t up, cudaStream_t s);

}  // namespace dory
