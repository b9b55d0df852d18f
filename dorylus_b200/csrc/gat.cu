// GAT edge operators.  The reference materialises per-edge tensors (az, A, dLRelu, and a dense
// E x F' "dAct" -- 58 GB for Reddit at F' = 128, CPU_comm.cpp:220).  Because the reference's score
// is one-sided (az[e] = z[dst(e)] . a_i, quirk Q8) every per-edge quantity factors through a
// per-destination scalar, so nothing E x F' is ever formed here:
//   forward : s[v] = z[v].a           az[e] = s[dst(e)]         A[e]  = lrelu(az[e])
//   backward: t[v] = grad[v].a        dA[e] = t[dst(e)] * lrelu'(az[e])
//             c[v] = sum_{e in in(v)} lrelu'(az[e])
//             dAct_reduce = grad^T . c   (== column sums of the reference's dAct)
//             da = (z^T z) . dAct_reduce
#include "common.cuh"
#include "gat.cuh"

namespace dory {
namespace {

constexpr unsigned kFull = 0xffffffffu;
constexpr float kAlpha = 0.01f;  // CPU_comm.cpp:385,398

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// a_i is stored as an F x 1 weight with row pitch 4 floats (padded_ld(1)).
constexpr uint32_t kALd = 4;

__global__ void __launch_bounds__(256)
gat_edge_forward_kernel(const float *__restrict__ z, uint32_t ld, uint32_t F, const float *__restrict__ a,
                        const uint64_t *__restrict__ colPtrs, uint32_t V, float *__restrict__ az,
                        float *__restrict__ A) {
    const uint32_t v = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (v >= V) return;
    const float *zr = z + (size_t)v * ld;
    float s = 0.f;
    for (uint32_t j = lane; j < F; j += 32) s = fmaf(zr[j], a[(size_t)j * kALd], s);
    s = warp_sum(s);
    const float act = s > 0.f ? s : kAlpha * s;
    for (uint64_t e = colPtrs[v] + lane; e < colPtrs[v + 1]; e += 32) {
        az[e] = s;
        A[e] = act;
    }
}

__global__ void __launch_bounds__(256)
gat_edge_backward_kernel(const float *__restrict__ grad, uint32_t ld, uint32_t F, const float *__restrict__ a,
                         const float *__restrict__ az, const uint64_t *__restrict__ colPtrs, uint32_t V,
                         float *__restrict__ dA, float *__restrict__ cvec) {
    const uint32_t v = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (v >= V) return;
    const float *gr = grad + (size_t)v * ld;
    float t = 0.f;
    for (uint32_t j = lane; j < F; j += 32) t = fmaf(gr[j], a[(size_t)j * kALd], t);
    t = warp_sum(t);
    float c = 0.f;
    for (uint64_t e = colPtrs[v] + lane; e < colPtrs[v + 1]; e += 32) {
        const float d = az[e] > 0.f ? 1.f : kAlpha;
        dA[e] = t * d;
        c += d;
    }
    c = warp_sum(c);
    if (lane == 0) cvec[v] = c;
}

// partial[b][j] = sum over the block's vertex range of grad[v][j] * c[v]   (fixed order per block)
__global__ void __launch_bounds__(256)
col_weighted_sum_kernel(const float *__restrict__ grad, uint32_t ld, const float *__restrict__ cvec,
                        uint32_t V, uint32_t rows_per_block, float *__restrict__ partial) {
    const uint32_t r0 = blockIdx.x * rows_per_block;
    const uint32_t r1 = min(V, r0 + rows_per_block);
    for (uint32_t j = threadIdx.x; j < ld; j += blockDim.x) {
        float s = 0.f;
        for (uint32_t v = r0; v < r1; ++v) s = fmaf(grad[(size_t)v * ld + j], cvec[v], s);
        partial[(size_t)blockIdx.x * ld + j] = s;
    }
}

__global__ void col_final_kernel(const float *__restrict__ partial, uint32_t nblk, uint32_t ld,
                                 float *__restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    float s = 0.f;
    for (uint32_t b = 0; b < nblk; ++b) s += partial[(size_t)b * ld + j];
    out[j] = s;
}

// da[i] = sum_j zz[i][j] * r[j]
__global__ void matvec_kernel(const float *__restrict__ zz, uint32_t ldz, const float *__restrict__ r,
                              uint32_t F, float *__restrict__ da) {
    const uint32_t i = blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= F) return;
    float s = 0.f;
    for (uint32_t j = lane; j < F; j += 32) s = fmaf(zz[(size_t)i * ldz + j], r[j], s);
    s = warp_sum(s);
    if (lane == 0) da[(size_t)i * kALd] = s;
}

constexpr int kMaxPerLane = 8;
__global__ void __launch_bounds__(256)
gat_predict_kernel(const float *__restrict__ logits, uint32_t ldl, const float *__restrict__ lab,
                   float *__restrict__ grad, uint32_t ld, uint32_t C, uint32_t low, uint32_t up) {
    const uint32_t row = low + blockIdx.x * 8 + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (row >= up) return;
    const float *zr = logits + (size_t)row * ldl;
    float v[kMaxPerLane];
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
        const uint32_t c = lane + 32 * j;
        v[j] = c < C ? zr[c] : -INFINITY;
        mx = fmaxf(mx, v[j]);
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
        const uint32_t c = lane + 32 * j;
        v[j] = c < C ? expf(v[j] - mx) : 0.f;
        sum += v[j];
    }
    sum = warp_sum(sum);
    const float denom = 1e-20f + sum;  // tensors.cpp:13-18
#pragma unroll
    for (int j = 0; j < kMaxPerLane; ++j) {
        const uint32_t c = lane + 32 * j;
        if (c < C) grad[(size_t)row * ld + c] = v[j] / denom - lab[(size_t)row * ld + c];
    }
}

}  // namespace

int launch_gat_edge_forward(const float *z, uint32_t ld, uint32_t F, const float *a, const uint64_t *colPtrs,
                            uint32_t V, float *az, float *A, cudaStream_t s) {
    if (V == 0) return 0;
    gat_edge_forward_kernel<<<(V + 7) / 8, 256, 0, s>>>(z, ld, F, a, colPtrs, V, az, A);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

int launch_gat_edge_backward(const GatEdgeBackwardArgs &a, cudaStream_t s) {
    if (a.V == 0) return 0;
    if (a.scratch_floats < (size_t)a.V + 2 * a.ld) return -1;
    float *cvec = a.scratch;
    float *reduced = a.scratch + a.V;  // [ld]
    int launches = 0;
    gat_edge_backward_kernel<<<(a.V + 7) / 8, 256, 0, s>>>(a.grad, a.ld, a.F, a.a, a.az, a.colPtrs, a.V, a.dA, cvec);
    ++launches;
    // dAct_reduce = grad^T . c
    const uint32_t rows_per_block = 512;
    const uint32_t nblk = (a.V + rows_per_block - 1) / rows_per_block;
    const size_t zz_floats = (size_t)a.ld * a.ld;
    if (a.ws_floats < (size_t)nblk * a.ld + zz_floats) return -1;
    float *partial = a.ws;
    col_weighted_sum_kernel<<<nblk, 256, 0, s>>>(a.grad, a.ld, cvec, a.V, rows_per_block, partial);
    col_final_kernel<<<(a.ld + 127) / 128, 128, 0, s>>>(partial, nblk, a.ld, reduced);
    launches += 2;
    // zz = z^T . z  (F' x F'), then da = zz . dAct_reduce
    float *zz = a.ws + (size_t)nblk * a.ld;
    GemmArgs g{};
    g.A = a.z; g.lda = a.ld; g.B = a.z; g.ldb = a.ld; g.C = zz; g.ldc = a.ld;
    g.M = a.ld; g.N = a.ld; g.K = a.V;
    g.transA = true; g.transB = false; g.epilogue = EPI_NONE;
    g.ws = zz + zz_floats;  // split over vertices, partials reduced in a fixed order
    g.ws_floats = a.ws_floats - ((size_t)nblk * a.ld + zz_floats);
    int n = launch_gemm(g, s);
    if (n < 0) return -1;
    launches += n;
    matvec_kernel<<<(a.F + 7) / 8, 256, 0, s>>>(zz, a.ld, reduced, a.F, a.da);
    ++launches;
    return cudaGetLastError() == cudaSuccess ? launches : -1;
}

int launch_gat_predict(const float *logits, uint32_t ld_logits, const float *lab, float *grad, uint32_t ld,
                       uint32_t C, uint32_t low, uint32_t up, cudaStream_t s) {
    if (up <= low) return 0;
    if (C > 32 * kMaxPerLane) return -1;
    gat_predict_kernel<<<(up - low + 7) / 8, 256, 0, s>>>(logits, ld_logits, lab, grad, ld, C, low, up);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

}  // namespace dory
