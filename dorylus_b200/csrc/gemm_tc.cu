// Dense apply Z = AH . W on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), with an
// error-compensated 3xTF32 split so that the result keeps fp32-level accuracy (the parity bar is
// 1e-5 against cblas_sgemm, which plain TF32 -- 10 mantissa bits -- cannot meet).
//
// Stands in for Matrix::dot -> cblas_sgemm in CPUComm::vtxNNForwardGCN/GAT (reference
// commmanager/CPU_comm.cpp:98-107,161-169) and for cublasSgemm + cudnnActivationForward in the
// reference GPU backend (GPU-Computation/comp_server.cu:104-176); the tanh epilogue is fused and
// writes both z and h.
//
// Shape of the problem: M = |V_p| (232,965 on Reddit) is huge, K = F_in (608 padded), N = F_out
// (128 / 64 padded).  One CTA owns a 128-row output tile and the whole N extent:
//   warp 0     TMA producer: per 32-wide K block one box of A (128 x 32 fp32, 128B swizzle) and the
//              matching boxes of W^T_hi and W^T_lo (pre-split once per call, K-major)
//   warps 2-5  split the A box in place into hi = a & 0xffffe000 (exactly representable in TF32)
//              and lo = a - hi (second buffer), then publish it to the async proxy
//   warp 1     one elected lane issues, per K block, 4 x {hi.hi -> D0, hi.lo -> D1, lo.hi -> D1}
//              tcgen05.mma.kind::tf32 (M=128, N, K=8); D0 / D1 are two fp32 accumulators in TMEM so
//              that the small correction terms never round against the large partial sums
//   warps 2-5  epilogue: tcgen05.ld D0 + D1 -> registers -> (tanh) -> 128-bit global stores
// Stages are handed around with mbarriers (TMA -> converters -> MMA -> TMA); tcgen05.commit frees a
// stage and finally signals the epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"
#include "gemm_tc.cuh"

namespace dory {
namespace {

constexpr int BM = 128;      // rows per CTA tile (UMMA M, cta_group::1)
constexpr int BK = 32;       // fp32 per K block = one 128-byte swizzle span
constexpr int UMMA_K = 8;    // K per tcgen05.mma.kind::tf32
constexpr int kThreads = 192;
constexpr int kConvThreads = 128;

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(bar), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap *map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem desc] . B[smem desc], kind::tf32, issued by ONE thread for the CTA
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc,
                                            uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// 32 lanes x 32 consecutive fp32 columns of TMEM -> 32 registers per thread
__device__ __forceinline__ void tc_ld_32x32(uint32_t taddr, float *v) {
    uint32_t *r = reinterpret_cast<uint32_t *>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// K-major operand tile, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO); the
// leading-dimension offset is unused for swizzled K-major layouts (set to 1 like CUTLASS).
// Bit layout: cute/arch/mma_sm100_desc.hpp (SmemDescriptor), version = 1, layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// Instruction descriptor (InstrDescriptor in the same header): c_format F32 (1) at bit 4,
// a/b format TF32 (2) at bits 7 / 10, both operands K-major, N>>3 at bit 17, M>>4 at bit 24.
template <int BN>
__host__ __device__ constexpr uint32_t umma_idesc_tf32() {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
}

template <int BN>
struct SmemLayout {
    static constexpr int kStages = BN == 128 ? 3 : 4;
    static constexpr uint32_t kABytes = BM * BK * 4;  // 16 KB
    static constexpr uint32_t kBBytes = BN * BK * 4;  // 16 KB / 8 KB
    static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;
    static constexpr uint32_t kTxBytes = kABytes + 2 * kBBytes;  // what TMA delivers per stage
    static constexpr uint32_t kBarOffset = kStages * kStageBytes;
    static constexpr uint32_t kTotal = kBarOffset + 256 + 1024;  // barriers + alignment slack
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBhi,
               const __grid_constant__ CUtensorMap mapBlo, float *__restrict__ C, float *__restrict__ C2,
               uint32_t ldc, uint64_t M, uint32_t nkb, int epilogue) {
    using L = SmemLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // 128B swizzle needs 1024 B alignment
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    auto stage_a_hi = [&](int s) { return base + s * L::kStageBytes; };
    auto stage_a_lo = [&](int s) { return base + s * L::kStageBytes + L::kABytes; };
    auto stage_b_hi = [&](int s) { return base + s * L::kStageBytes + 2 * L::kABytes; };
    auto stage_b_lo = [&](int s) { return base + s * L::kStageBytes + 2 * L::kABytes + L::kBBytes; };
    const uint32_t bar0 = base + L::kBarOffset;
    auto bar_full = [&](int s) { return bar0 + 8 * s; };                   // TMA bytes landed
    auto bar_conv = [&](int s) { return bar0 + 8 * (L::kStages + s); };    // A split published
    auto bar_empty = [&](int s) { return bar0 + 8 * (2 * L::kStages + s); };  // MMAs retired
    const uint32_t bar_done = bar0 + 8 * (3 * L::kStages);                 // accumulators final
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen_base + L::kBarOffset + 8 * (3 * L::kStages + 1));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapBhi);
        tma_prefetch_desc(&mapBlo);
        for (int s = 0; s < L::kStages; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_conv(s), kConvThreads);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {  // TMEM: two fp32 accumulators of BN columns each
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(2 * BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int m0 = blockIdx.x * BM;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const int s = kb % L::kStages;
                const uint32_t ph = (kb / L::kStages) & 1;
                mbar_wait(bar_empty(s), ph ^ 1);
                mbar_expect_tx(bar_full(s), L::kTxBytes);
                tma_load_2d(stage_a_hi(s), &mapA, kb * BK, m0, bar_full(s));
                tma_load_2d(stage_b_hi(s), &mapBhi, kb * BK, 0, bar_full(s));
                tma_load_2d(stage_b_lo(s), &mapBlo, kb * BK, 0, bar_full(s));
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        constexpr uint32_t idesc = umma_idesc_tf32<BN>();
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const int s = kb % L::kStages;
            const uint32_t ph = (kb / L::kStages) & 1;
            mbar_wait(bar_full(s), ph);
            mbar_wait(bar_conv(s), ph);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * UMMA_K * 4;  // bytes along K inside the swizzle span
                    const uint64_t a_hi = umma_desc_sw128(stage_a_hi(s) + koff);
                    const uint64_t a_lo = umma_desc_sw128(stage_a_lo(s) + koff);
                    const uint64_t b_hi = umma_desc_sw128(stage_b_hi(s) + koff);
                    const uint64_t b_lo = umma_desc_sw128(stage_b_lo(s) + koff);
                    const uint32_t acc = (kb | (uint32_t)k) ? 1u : 0u;
                    tc_mma_tf32(tmem, a_hi, b_hi, idesc, acc);       // D0 += hi . hi
                    tc_mma_tf32(tmem + BN, a_hi, b_lo, idesc, acc);  // D1 += hi . lo
                    tc_mma_tf32(tmem + BN, a_lo, b_hi, idesc, 1u);   // D1 += lo . hi
                }
                tc_commit(bar_empty(s));                 // stage reusable once these MMAs retire
                if (kb + 1 == nkb) tc_commit(bar_done);  // accumulators complete
            }
            __syncwarp();
        }
    } else {
        // ===================== A splitter, then epilogue =====================
        const int t = threadIdx.x - 64;  // 0..127
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const int s = kb % L::kStages;
            const uint32_t ph = (kb / L::kStages) & 1;
            mbar_wait(bar_full(s), ph);
            float4 *hi = reinterpret_cast<float4 *>(gen_base + s * L::kStageBytes);
            float4 *lo = reinterpret_cast<float4 *>(gen_base + s * L::kStageBytes + L::kABytes);
            // the split is element-wise, so the swizzled placement is preserved by construction
#pragma unroll
            for (int i = 0; i < (BM * BK / 4) / kConvThreads; ++i) {
                const int idx = t + i * kConvThreads;
                const float4 v = hi[idx];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                l.x = v.x - h.x;
                l.y = v.y - h.y;
                l.z = v.z - h.z;
                l.w = v.w - h.w;
                hi[idx] = h;
                lo[idx] = l;
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
            mbar_arrive(bar_conv(s));
        }
        // epilogue: TMEM lane quarter of this warp (warp % 4), one output row per thread
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const int q = warp & 3;
        const uint64_t m = (uint64_t)m0 + q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            float d0[32], d1[32];
            tc_ld_32x32(lane_base + c * 32, d0);
            tc_ld_32x32(lane_base + BN + c * 32, d1);
            tc_wait_ld();
            if (m < M) {
                float4 *out = reinterpret_cast<float4 *>(C + m * ldc + c * 32);
                float4 *out2 = reinterpret_cast<float4 *>(C2 + m * ldc + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    float4 v = make_float4(d0[4 * j] + d1[4 * j], d0[4 * j + 1] + d1[4 * j + 1],
                                           d0[4 * j + 2] + d1[4 * j + 2], d0[4 * j + 3] + d1[4 * j + 3]);
                    out[j] = v;
                    if (epilogue == EPI_TANH) out2[j] = make_float4(tanhf(v.x), tanhf(v.y), tanhf(v.z), tanhf(v.w));
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN) : "memory");
    }
}

// ------------------------------------------------------------------ last layer: logits + soft-max, skinny K
//
// The last ApplyVertex of a GCN is [V x K] . [K x C] with K <= 128 and C <= 64 followed by a row soft-max
// (CPU_comm.cpp:98-121, 276-297, 448-471).  gemm_tc_kernel above holds 3-4 stages of 48-64 KB: one CTA per
// SM, and with one or two K blocks per tile a CTA is all latency (launch, TMA, split, MMA, epilogue, one
// after the other: 4.5 us per 128-row tile on the 8.2 M-row Friendster partition).  Here a CTA takes only
// `nst` <= 2 stages (BN = 64: 48 KB each) so that 2-4 CTAs share an SM and cover each other's latencies,
// and the epilogue is the whole of softmax_ce_kernel: a TMEM lane is an output row, so tcgen05.ld hands every
// thread the C logits of ONE vertex -- the soft-max, both arg-maxes, the validation statistics, the maskout
// (quirk Q6) and d = (P - Y) / scale need no shuffle and the logits never reach HBM.
// Soft-max, validation statistics, maskout and d = (P - Y) / scale of ONE vertex row whose logits are in v
// (softmax_ce_kernel of dense.cu, sequential over the classes): CPU_comm.cpp:108-121, 276-297, 448-471.
// The label row is read from, and d written to, the thread's row of the warp's staging slab in shared memory
// (`wrow`, chunk j at wrow[swz(j)]): the slab is filled and drained with coalesced global accesses.
template <int BN, class Swz>
__device__ __forceinline__ void softmax_ce_row(float (&v)[BN], uint64_t row, const SoftmaxCEArgs &a, float4 *wrow, Swz swz) {
    float mx = -INFINITY;
#pragma unroll
    for (int j = 0; j < BN; ++j)
        if ((uint32_t)j < a.C) mx = fmaxf(mx, v[j]);
    float sum = 0.f;
#pragma unroll
    for (int j = 0; j < BN; ++j) {
        v[j] = (uint32_t)j < a.C ? expf(v[j] - mx) : 0.f;
        sum += v[j];
    }
    // One reciprocal per row instead of a division per class (the row's 2 x C divisions were most of this
    // epilogue's instructions): P and d differ from the divided form by at most one rounding.
    const float rdenom = 1.f / (1e-20f + sum);  // CPU_comm.cpp:285-290
    const float rscale = 1.f / a.denom;
    // maskout (quirk Q6): the classes [clo, chi) of this row lie in the overwritten range of the dense V x C array
    uint32_t clo = 0, chi = 0;
    if (a.strictMask) {
        if (row >= a.trainEnd) chi = a.C;
    } else {
        const uint64_t f0 = row * a.C, mb = (uint64_t)a.trainEnd * a.C, me = mb + a.maskFloats;
        if (f0 + a.C > mb && f0 < me) {
            clo = mb > f0 ? (uint32_t)(mb - f0) : 0u;
            chi = me - f0 < a.C ? (uint32_t)(me - f0) : a.C;
        }
    }
    float pbest = -INFINITY, lbest = -INFINITY, lab_at_pbest = 0.f, p_at_lbest = 1.f;
    float4 *pred4 = a.pred ? reinterpret_cast<float4 *>(a.pred + row * a.ld) : nullptr;
#pragma unroll
    for (int j4 = 0; j4 < BN / 4; ++j4) {
        const float4 l4 = wrow[swz(j4)];
        const float lb[4] = {l4.x, l4.y, l4.z, l4.w};
        float d[4], pr[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const uint32_t c = 4 * j4 + k;
            const float p = v[c] * rdenom;
            pr[k] = p;
            // first maximum wins, like the reference's argmax helper (classes >= C hold p = 0 and label 0: the
            // strict comparisons never pick them unless the row has no positive entry at all, where class 0 won already)
            if (c < a.C && p > pbest) pbest = p, lab_at_pbest = lb[k];
            if (c < a.C && lb[k] > lbest) lbest = lb[k], p_at_lbest = p;
            // maskout, then hadamardSub and the scale (CPU_comm.cpp:118-121, 464-471); padding columns stay zero
            const bool keep = c < a.C && !(c >= clo && c < chi);
            d[k] = keep ? (p - lb[k]) * rscale : 0.f;
        }
        wrow[swz(j4)] = make_float4(d[0], d[1], d[2], d[3]);
        if (pred4) pred4[j4] = make_float4(pr[0], pr[1], pr[2], pr[3]);
    }
    if (row >= a.trainEnd && row < a.valEnd) {  // getTrainStat over the validation slice (CPU_comm.cpp:448-462)
        a.rowstat[row - a.trainEnd] = lab_at_pbest;
        a.rowstat[(size_t)a.V + (row - a.trainEnd)] = -logf(p_at_lbest);
    }
}

// SOFTMAX = false: the same pipeline with the plain epilogue of gemm_tc_kernel (C, and C2 = tanh(C) for EPI_TANH)
// -- the skinny products of the Amazon / Friendster widths (K <= 128, N <= 64) and grad = G . W^T.
struct SmallOut {
    float *C, *C2;
    uint32_t ldc;
    int epilogue;
};

template <int BN, bool SOFTMAX>
__global__ void __launch_bounds__(kThreads, (SOFTMAX && BN == 64) ? 3 : 4)
gemm_tc_small_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapBhi,
                     const __grid_constant__ CUtensorMap mapBlo, uint32_t nkb, uint32_t nst, uint64_t M, const SmallOut o,
                     const SoftmaxCEArgs a) {
    using L = SmemLayout<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // [base, base + 1024): barriers and the TMEM slot; stages follow (1024 B aligned for the 128 B swizzle)
    constexpr int kMaxStages = 4;
    auto stage_off = [&](uint32_t s) { return 1024u + s * L::kStageBytes; };
    auto stage_a_hi = [&](uint32_t s) { return base + stage_off(s); };
    auto stage_a_lo = [&](uint32_t s) { return base + stage_off(s) + L::kABytes; };
    auto stage_b_hi = [&](uint32_t s) { return base + stage_off(s) + 2 * L::kABytes; };
    auto stage_b_lo = [&](uint32_t s) { return base + stage_off(s) + 2 * L::kABytes + L::kBBytes; };
    auto bar_full = [&](uint32_t s) { return base + 8 * s; };
    auto bar_conv = [&](uint32_t s) { return base + 8 * (kMaxStages + s); };
    auto bar_empty = [&](uint32_t s) { return base + 8 * (2 * kMaxStages + s); };
    const uint32_t bar_done = base + 8 * (3 * kMaxStages);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen_base + 8 * (3 * kMaxStages + 1));

    const uint64_t m0 = (uint64_t)blockIdx.x * BM;
    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapBhi);
        tma_prefetch_desc(&mapBlo);
        for (uint32_t s = 0; s < nst; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_conv(s), kConvThreads);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(2 * BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    // Epilogue traffic goes through a staging slab in shared memory (the A buffers of stage 0, dead once the
    // accumulators are final): a thread owns a ROW of the tile (TMEM lane), but a warp instruction that touches 32
    // rows 128-256 B apart costs the L1 one tag cycle per row.  One warp instruction here covers RPI whole rows
    // (lane -> row i*RPI + lane / CPR, 16-byte chunk lane % CPR); in the slab chunk c of row r sits at
    // r*CPR + (c ^ r) & 7 (within its group of 8), conflict-free for both the row-per-thread and the coalesced view.
    constexpr int CPR = BN / 4, RPI = 32 / CPR, NIT = 32 / RPI;
    const int q = warp & 3;  // TMEM lane quarter of this warp = its 32-row slab of the tile
    const int crow = lane / CPR, cchunk = lane % CPR;
    auto swz = [](int r, int c) { return r * CPR + ((c & ~7) | ((c ^ r) & 7)); };
    float4 labr[SOFTMAX ? NIT : 1];
    if constexpr (SOFTMAX) {
        if (warp >= 2) {  // the labels are needed only in the epilogue: fetch them now, behind the TMA / MMA latency
            const float4 *lab4 = reinterpret_cast<const float4 *>(a.lab);
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const uint64_t row = m0 + (uint64_t)q * 32 + i * RPI + crow;
                labr[i] = row < M ? lab4[row * CPR + cchunk] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const uint32_t s = kb % nst, ph = (kb / nst) & 1;
                mbar_wait(bar_empty(s), ph ^ 1);
                mbar_expect_tx(bar_full(s), L::kTxBytes);
                tma_load_2d(stage_a_hi(s), &mapA, kb * BK, (int)m0, bar_full(s));
                tma_load_2d(stage_b_hi(s), &mapBhi, kb * BK, 0, bar_full(s));
                tma_load_2d(stage_b_lo(s), &mapBlo, kb * BK, 0, bar_full(s));
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_tf32<BN>();
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const uint32_t s = kb % nst, ph = (kb / nst) & 1;
            mbar_wait(bar_full(s), ph);
            mbar_wait(bar_conv(s), ph);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * UMMA_K * 4;
                    const uint64_t a_hi = umma_desc_sw128(stage_a_hi(s) + koff);
                    const uint64_t a_lo = umma_desc_sw128(stage_a_lo(s) + koff);
                    const uint64_t b_hi = umma_desc_sw128(stage_b_hi(s) + koff);
                    const uint64_t b_lo = umma_desc_sw128(stage_b_lo(s) + koff);
                    const uint32_t acc = (kb | (uint32_t)k) ? 1u : 0u;
                    tc_mma_tf32(tmem, a_hi, b_hi, idesc, acc);
                    tc_mma_tf32(tmem + BN, a_hi, b_lo, idesc, acc);
                    tc_mma_tf32(tmem + BN, a_lo, b_hi, idesc, 1u);
                }
                tc_commit(bar_empty(s));
                if (kb + 1 == nkb) tc_commit(bar_done);
            }
            __syncwarp();
        }
    } else {
        const int t = threadIdx.x - 64;
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const uint32_t s = kb % nst, ph = (kb / nst) & 1;
            mbar_wait(bar_full(s), ph);
            float4 *hi = reinterpret_cast<float4 *>(gen_base + stage_off(s));
            float4 *lo = reinterpret_cast<float4 *>(gen_base + stage_off(s) + L::kABytes);
#pragma unroll
            for (int i = 0; i < (BM * BK / 4) / kConvThreads; ++i) {
                const int idx = t + i * kConvThreads;
                const float4 v = hi[idx];
                float4 h, l;
                h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
                h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
                h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
                h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
                l.x = v.x - h.x;
                l.y = v.y - h.y;
                l.z = v.z - h.z;
                l.w = v.w - h.w;
                hi[idx] = h;
                lo[idx] = l;
            }
            fence_proxy_async();
            mbar_arrive(bar_conv(s));
        }
        // ---- epilogue: thread = vertex row (TMEM lane quarter of this warp)
        mbar_wait(bar_done, 0);
        tc_fence_after();
        const uint64_t row = m0 + (uint64_t)q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        float4 *wbuf = reinterpret_cast<float4 *>(gen_base + stage_off(0)) + q * 32 * CPR;  // this warp's slab
        float4 *wrow = wbuf + lane * CPR;
        auto swz_mine = [&](int c) { return (c & ~7) | ((c ^ lane) & 7); };
        if constexpr (SOFTMAX) {
#pragma unroll
            for (int i = 0; i < NIT; ++i) wbuf[swz(i * RPI + crow, cchunk)] = labr[i];
            __syncwarp();
            float v[BN];
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                float d0[32], d1[32];
                tc_ld_32x32(lane_base + c * 32, d0);
                tc_ld_32x32(lane_base + BN + c * 32, d1);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) v[c * 32 + j] = d0[j] + d1[j];
            }
            if (row < M) softmax_ce_row<BN>(v, row, a, wrow, swz_mine);
            __syncwarp();
            float4 *d4 = reinterpret_cast<float4 *>(a.d);
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const uint64_t r = m0 + (uint64_t)q * 32 + i * RPI + crow;
                if (r < M) d4[r * CPR + cchunk] = wbuf[swz(i * RPI + crow, cchunk)];
            }
        } else {
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                float d0[32], d1[32];
                tc_ld_32x32(lane_base + c * 32, d0);
                tc_ld_32x32(lane_base + BN + c * 32, d1);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    wrow[swz_mine(c * 8 + j)] = make_float4(d0[4 * j] + d1[4 * j], d0[4 * j + 1] + d1[4 * j + 1],
                                                            d0[4 * j + 2] + d1[4 * j + 2], d0[4 * j + 3] + d1[4 * j + 3]);
            }
            __syncwarp();
            float4 *c4 = reinterpret_cast<float4 *>(o.C), *t4 = reinterpret_cast<float4 *>(o.C2);
#pragma unroll
            for (int i = 0; i < NIT; ++i) {
                const uint64_t r = m0 + (uint64_t)q * 32 + i * RPI + crow;
                if (r >= M) continue;
                const float4 z = wbuf[swz(i * RPI + crow, cchunk)];
                c4[r * CPR + cchunk] = z;  // the output pitch is BN
                if (o.epilogue == EPI_TANH) t4[r * CPR + cchunk] = make_float4(tanhf(z.x), tanhf(z.y), tanhf(z.z), tanhf(z.w));
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN) : "memory");
    }
}

// Wt_hi[n][k], Wt_lo[n][k] (K-major, [BN x Kpad]) from W[k][n] ([Kpad x ldw]); rows n >= ldw are zero.
__global__ void split_transpose_kernel(const float *__restrict__ W, uint32_t ldw, uint32_t Kpad, uint32_t BN,
                                       float *__restrict__ hi, float *__restrict__ lo) {
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t n = blockIdx.y;
    if (k >= Kpad) return;
    const float w = n < ldw ? W[(size_t)k * ldw + n] : 0.f;
    const float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    hi[(size_t)n * Kpad + k] = h;
    lo[(size_t)n * Kpad + k] = w - h;
}

// ------------------------------------------------------------------ dW = A^T . G  (split over the vertices)
//
// Weight gradient of ApplyVertex: dW[Fin x Fout] = AH^T . G with AH [V x Fin] and G [V x Fout] both
// row-major, i.e. the contraction runs over the ROWS of both operands (CPUComm::vtxNNBackwardGCN,
// CPU_comm.cpp:146-151: ah.dot(interGrad, true, false); cublasSgemm with CUBLAS_OP_T in the
// reference GPU backend).  For the tensor core that makes both operands MN-major: a TMA box of
// 32 vertices x 32 floats is 32 rows of 128 B (row = one vertex, the K index).  The only shared-memory
// layout tcgen05 accepts for MN-major 32-bit operands is the 128-byte swizzle with 32-byte atomicity
// (layout type SWIZZLE_128B_BASE32B; the four 32 B chunks of a row are XOR-ed with row & 3, TMA mode
// CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B): its canonical form is 4-row (K) groups 512 B apart and 32-float
// (M / N) blocks one box (4096 B) apart, which is what the boxes are -- so no transpose is ever
// materialised; the instruction descriptor just says "A and B are MN-major".
//   grid = (M tiles of 128, splits): a CTA reduces its vertex range into one 128 x BN fp32 tile
//   (3xTF32: hi.hi -> D0, hi.lo + lo.hi -> D1, as above, both operands split in shared memory by the
//   converter warps) and writes the partial to ws[split]; splitk_reduce adds the partials in
//   ascending order (deterministic dW).
__device__ __forceinline__ uint64_t umma_desc_mn_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
    d |= (uint64_t)(4096 >> 4) << 16;  // leading byte offset: next 32-element block along M / N
    d |= (uint64_t)(512 >> 4) << 32;   // stride byte offset: next 4-row group along K
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)1 << 61;            // SWIZZLE_128B_BASE32B
    return d;
}

template <int BN>
__host__ __device__ constexpr uint32_t umma_idesc_tf32_mn() {
    return umma_idesc_tf32<BN>() | (1u << 15) | (1u << 16);  // a_major = b_major = MN
}

template <int BN>
struct SmemLayoutTN {
    static constexpr int kStages = BN == 128 ? 3 : 4;
    static constexpr uint32_t kBox = 32 * BK * 4;        // one TMA box: 32 floats x 32 vertices = 4 KB
    static constexpr uint32_t kABytes = BM * BK * 4;     // 4 boxes
    static constexpr uint32_t kBBytes = BN * BK * 4;     // BN / 32 boxes
    static constexpr uint32_t kStageBytes = 2 * kABytes + 2 * kBBytes;  // hi and lo of both operands
    static constexpr uint32_t kTxBytes = kABytes + kBBytes;
    static constexpr uint32_t kBarOffset = kStages * kStageBytes;
    static constexpr uint32_t kTotal = kBarOffset + 256 + 1024;
};

template <int BN>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tn_tc_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapG,
                  float *__restrict__ ws, uint32_t ldc, uint32_t Mpad, uint64_t K, uint64_t kchunk,
                  size_t split_stride) {
    using L = SmemLayoutTN<BN>;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t *gen_base = smem_raw + (base - smem_u32(smem_raw));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    auto stage_a_hi = [&](int s) { return base + s * L::kStageBytes; };
    auto stage_a_lo = [&](int s) { return base + s * L::kStageBytes + L::kABytes; };
    auto stage_b_hi = [&](int s) { return base + s * L::kStageBytes + 2 * L::kABytes; };
    auto stage_b_lo = [&](int s) { return base + s * L::kStageBytes + 2 * L::kABytes + L::kBBytes; };
    const uint32_t bar0 = base + L::kBarOffset;
    auto bar_full = [&](int s) { return bar0 + 8 * s; };
    auto bar_conv = [&](int s) { return bar0 + 8 * (L::kStages + s); };
    auto bar_empty = [&](int s) { return bar0 + 8 * (2 * L::kStages + s); };
    const uint32_t bar_done = bar0 + 8 * (3 * L::kStages);
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(gen_base + L::kBarOffset + 8 * (3 * L::kStages + 1));

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&mapA);
        tma_prefetch_desc(&mapG);
        for (int s = 0; s < L::kStages; ++s) {
            mbar_init(bar_full(s), 1);
            mbar_init(bar_conv(s), kConvThreads);
            mbar_init(bar_empty(s), 1);
        }
        mbar_init(bar_done, 1);
        fence_barrier_init();
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                     "n"(2 * BN)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    const int m0 = blockIdx.x * BM;
    const uint64_t kbeg = (uint64_t)blockIdx.y * kchunk;
    const uint64_t kend = kbeg + kchunk < K ? kbeg + kchunk : K;
    const uint32_t nkb = (uint32_t)((kend - kbeg + BK - 1) / BK);  // rows past K are zero-filled by TMA

    if (warp == 0) {
        if (lane == 0) {
            for (uint32_t kb = 0; kb < nkb; ++kb) {
                const int s = kb % L::kStages;
                const uint32_t ph = (kb / L::kStages) & 1;
                mbar_wait(bar_empty(s), ph ^ 1);
                mbar_expect_tx(bar_full(s), L::kTxBytes);
                const int k0 = (int)(kbeg + (uint64_t)kb * BK);
#pragma unroll
                for (int i = 0; i < BM / 32; ++i) tma_load_2d(stage_a_hi(s) + i * L::kBox, &mapA, m0 + 32 * i, k0, bar_full(s));
#pragma unroll
                for (int j = 0; j < BN / 32; ++j) tma_load_2d(stage_b_hi(s) + j * L::kBox, &mapG, 32 * j, k0, bar_full(s));
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = umma_idesc_tf32_mn<BN>();
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const int s = kb % L::kStages;
            const uint32_t ph = (kb / L::kStages) & 1;
            mbar_wait(bar_full(s), ph);
            mbar_wait(bar_conv(s), ph);
            tc_fence_after();
            if (lane == 0) {
#pragma unroll
                for (int k = 0; k < BK / UMMA_K; ++k) {
                    const uint32_t koff = k * 1024;  // one 8-vertex group of every box
                    const uint64_t a_hi = umma_desc_mn_sw128(stage_a_hi(s) + koff);
                    const uint64_t a_lo = umma_desc_mn_sw128(stage_a_lo(s) + koff);
                    const uint64_t b_hi = umma_desc_mn_sw128(stage_b_hi(s) + koff);
                    const uint64_t b_lo = umma_desc_mn_sw128(stage_b_lo(s) + koff);
                    const uint32_t acc = (kb | (uint32_t)k) ? 1u : 0u;
                    tc_mma_tf32(tmem, a_hi, b_hi, idesc, acc);
                    tc_mma_tf32(tmem + BN, a_hi, b_lo, idesc, acc);
                    tc_mma_tf32(tmem + BN, a_lo, b_hi, idesc, 1u);
                }
                tc_commit(bar_empty(s));
                if (kb + 1 == nkb) tc_commit(bar_done);
            }
            __syncwarp();
        }
    } else {
        const int t = threadIdx.x - 64;
        auto split = [](const float4 v, float4 &h, float4 &l) {
            h.x = __uint_as_float(__float_as_uint(v.x) & 0xffffe000u);
            h.y = __uint_as_float(__float_as_uint(v.y) & 0xffffe000u);
            h.z = __uint_as_float(__float_as_uint(v.z) & 0xffffe000u);
            h.w = __uint_as_float(__float_as_uint(v.w) & 0xffffe000u);
            l.x = v.x - h.x;
            l.y = v.y - h.y;
            l.z = v.z - h.z;
            l.w = v.w - h.w;
        };
        for (uint32_t kb = 0; kb < nkb; ++kb) {
            const int s = kb % L::kStages;
            const uint32_t ph = (kb / L::kStages) & 1;
            mbar_wait(bar_full(s), ph);
            float4 *ahi = reinterpret_cast<float4 *>(gen_base + s * L::kStageBytes);
            float4 *alo = reinterpret_cast<float4 *>(gen_base + s * L::kStageBytes + L::kABytes);
            float4 *bhi = reinterpret_cast<float4 *>(gen_base + s * L::kStageBytes + 2 * L::kABytes);
            float4 *blo = reinterpret_cast<float4 *>(gen_base + s * L::kStageBytes + 2 * L::kABytes + L::kBBytes);
#pragma unroll
            for (int i = 0; i < (BM * BK / 4) / kConvThreads; ++i) {
                const int idx = t + i * kConvThreads;
                float4 h, l;
                split(ahi[idx], h, l);
                ahi[idx] = h;
                alo[idx] = l;
            }
#pragma unroll
            for (int i = 0; i < (BN * BK / 4) / kConvThreads; ++i) {
                const int idx = t + i * kConvThreads;
                float4 h, l;
                split(bhi[idx], h, l);
                bhi[idx] = h;
                blo[idx] = l;
            }
            fence_proxy_async();
            mbar_arrive(bar_conv(s));
        }
        float *out = ws + (size_t)blockIdx.y * split_stride;
        const int q = warp & 3;
        const uint32_t m = (uint32_t)m0 + q * 32 + lane;
        const uint32_t lane_base = tmem + ((uint32_t)(q * 32) << 16);
        if (nkb) {
            mbar_wait(bar_done, 0);
            tc_fence_after();
        }
#pragma unroll 1
        for (int c = 0; c < BN / 32; ++c) {
            float d0[32], d1[32];
            if (nkb) {
                tc_ld_32x32(lane_base + c * 32, d0);
                tc_ld_32x32(lane_base + BN + c * 32, d1);
                tc_wait_ld();
            } else {  // empty vertex range (more splits than 32-vertex blocks): contributes zero
#pragma unroll
                for (int j = 0; j < 32; ++j) d0[j] = d1[j] = 0.f;
            }
            if (m < Mpad) {
                float4 *o = reinterpret_cast<float4 *>(out + (size_t)m * ldc + c * 32);
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    o[j] = make_float4(d0[4 * j] + d1[4 * j], d0[4 * j + 1] + d1[4 * j + 1], d0[4 * j + 2] + d1[4 * j + 2],
                                       d0[4 * j + 3] + d1[4 * j + 3]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * BN) : "memory");
    }
}

// C[i] = sum_z ws[z][i], ascending z
__global__ void tn_reduce_kernel(const float4 *__restrict__ ws, float4 *__restrict__ C, size_t n4, size_t stride4,
                                 int nsplit) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n4) return;
    float4 s = ws[i];
    for (int z = 1; z < nsplit; ++z) {
        const float4 v = ws[(size_t)z * stride4 + i];
        s.x += v.x;
        s.y += v.y;
        s.z += v.z;
        s.w += v.w;
    }
    C[i] = s;
}

// ------------------------------------------------------------------ host side
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                              const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                              CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeFn encode_fn() {
    static EncodeFn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeFn>(p);
    }
    return fn;
}

bool make_map(CUtensorMap *map, const float *ptr, uint64_t rows, uint32_t cols, uint32_t ld, uint32_t boxRows,
              CUtensorMapSwizzle swizzle = CU_TENSOR_MAP_SWIZZLE_128B) {
    EncodeFn fn = encode_fn();
    if (!fn) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    const cuuint32_t box[2] = {(cuuint32_t)BK, boxRows};
    const cuuint32_t estr[2] = {1, 1};
    return fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(ptr), dims, strides, box, estr,
              CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct WeightSplit {
    float *hi = nullptr, *lo = nullptr;
    CUtensorMap mapHi, mapLo;
};
struct Cache {
    std::mutex mu;
    std::map<std::tuple<const float *, uint32_t, uint32_t>, WeightSplit> weights;  // (W, ldw, Kpad)
    std::map<std::tuple<const float *, uint32_t, uint64_t>, CUtensorMap> amaps;    // (A, lda, M)
};
Cache &cache() {
    static Cache c;
    return c;
}

template <int BN>
int launch_bn(const float *A, uint32_t lda, uint64_t M, const float *W, uint32_t ldw, uint32_t Kpad, float *C,
              float *C2, uint32_t ldc, int epilogue, cudaStream_t s) {
    using L = SmemLayout<BN>;
    Cache &c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto wkey = std::make_tuple(W, ldw, Kpad);
    auto wit = c.weights.find(wkey);
    if (wit == c.weights.end()) {
        WeightSplit ws;
        const size_t bytes = (size_t)BN * Kpad * 4;
        if (cudaMalloc(&ws.hi, bytes) != cudaSuccess || cudaMalloc(&ws.lo, bytes) != cudaSuccess) return -1;
        if (!make_map(&ws.mapHi, ws.hi, BN, Kpad, Kpad, BN) || !make_map(&ws.mapLo, ws.lo, BN, Kpad, Kpad, BN)) return 0;
        wit = c.weights.emplace(wkey, ws).first;
    }
    auto akey = std::make_tuple(A, lda, M);
    auto ait = c.amaps.find(akey);
    if (ait == c.amaps.end()) {
        CUtensorMap m;
        if (!make_map(&m, A, M, Kpad, lda, BM)) return 0;
        ait = c.amaps.emplace(akey, m).first;
    }
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kTotal) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    // weights change every epoch (Adam): re-split them on every call (0.3 MB)
    dim3 sg((Kpad + 127) / 128, BN);
    split_transpose_kernel<<<sg, 128, 0, s>>>(W, ldw, Kpad, BN, wit->second.hi, wit->second.lo);
    const unsigned grid = (unsigned)((M + BM - 1) / BM);
    gemm_tc_kernel<BN><<<grid, kThreads, L::kTotal, s>>>(ait->second, wit->second.mapHi, wit->second.mapLo, C,
                                                        C2 ? C2 : C, ldc, M, Kpad / BK, C2 ? epilogue : EPI_NONE);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 2;
}

// W_hi / W_lo of a weight matrix used UNtransposed: grad = G . W^T contracts over W's columns, so W [prows x ld]
// row-major already is the K-major [N x K] operand.
__global__ void split_kernel(const float *__restrict__ W, size_t n, float *__restrict__ hi, float *__restrict__ lo) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float w = W[i];
    const float h = __uint_as_float(__float_as_uint(w) & 0xffffe000u);
    hi[i] = h;
    lo[i] = w - h;
}

// transposed = false: B operand from W^T (Z = A . W, W [Kpad x ldw], BN = ldw); true: from W itself
// (C = A . W^T, W [BN x Kpad] with pitch Kpad).
template <int BN, bool SOFTMAX>
int launch_small(const float *A, uint32_t lda, uint64_t M, const float *W, uint32_t ldw, uint32_t Kpad, bool transposed,
                 const SmallOut &o, const SoftmaxCEArgs &a, int stages, cudaStream_t s) {
    using L = SmemLayout<BN>;
    Cache &c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    // the split copies are [BN x Kpad]: BN is W's pitch for the transposed copy, but W's padded ROW count for the
    // untransposed one, which the pitch does not determine -- it is part of the key (engines created one after the
    // other in a process get the same addresses back from the allocator)
    auto wkey = std::make_tuple(W, transposed ? (ldw | ((uint32_t)BN << 8) | 0x80000000u) : ldw, Kpad);
    auto wit = c.weights.find(wkey);
    if (wit == c.weights.end()) {
        WeightSplit ws;
        const size_t bytes = (size_t)BN * Kpad * 4;
        if (cudaMalloc(&ws.hi, bytes) != cudaSuccess || cudaMalloc(&ws.lo, bytes) != cudaSuccess) return -1;
        if (!make_map(&ws.mapHi, ws.hi, BN, Kpad, Kpad, BN) || !make_map(&ws.mapLo, ws.lo, BN, Kpad, Kpad, BN)) return 0;
        wit = c.weights.emplace(wkey, ws).first;
    }
    auto akey = std::make_tuple(A, lda, M | ((uint64_t)Kpad << 40));  // the map's extent along K is part of it
    auto ait = c.amaps.find(akey);
    if (ait == c.amaps.end()) {
        CUtensorMap m;
        if (!make_map(&m, A, M, Kpad, lda, BM)) return 0;
        ait = c.amaps.emplace(akey, m).first;
    }
    const uint32_t nkb = Kpad / BK;
    // one stage per CTA (three to four CTAs per SM) unless the K loop is long enough to want its own ring
    const uint32_t nst = stages > 0 ? std::min<uint32_t>((uint32_t)stages, std::min<uint32_t>(nkb, 4)) : (nkb <= 2 ? 1u : 2u);
    const uint32_t smem = 1024 + 1024 + nst * L::kStageBytes;
    static uint32_t granted = 0;
    if (smem > granted) {
        if (cudaFuncSetAttribute(gemm_tc_small_kernel<BN, SOFTMAX>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
            return -1;
        granted = smem;
    }
    // weights change every epoch (Adam): re-split them on every call
    if (transposed) {
        const size_t n = (size_t)BN * Kpad;
        split_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(W, n, wit->second.hi, wit->second.lo);
    } else {
        dim3 sg((Kpad + 127) / 128, BN);
        split_transpose_kernel<<<sg, 128, 0, s>>>(W, ldw, Kpad, BN, wit->second.hi, wit->second.lo);
    }
    const unsigned grid = (unsigned)((M + BM - 1) / BM);
    gemm_tc_small_kernel<BN, SOFTMAX><<<grid, kThreads, smem, s>>>(ait->second, wit->second.mapHi, wit->second.mapLo, nkb, nst,
                                                                   M, o, a);
    if (cudaGetLastError() != cudaSuccess) return -1;
    return 2;
}

}  // namespace

int launch_gemm_tc_softmax(const float *A, uint32_t lda, const float *W, uint32_t ldw, uint32_t Kpad, const SoftmaxCEArgs &a,
                           int stages, cudaStream_t s) {
    // Supported: K a multiple of the 32-float swizzle span, the label / logit pitch exactly 32 or 64 (one row's
    // classes in one accumulator tile).  Anything else takes the fp32 path (launch_gemm_softmax_ce, dense.cu).
    if (Kpad % BK != 0 || Kpad == 0 || lda < Kpad || a.ld != ldw || a.C > a.ld || a.V == 0) return 0;
    const SmallOut none{nullptr, nullptr, 0, EPI_NONE};
    if (ldw == 64) return launch_small<64, true>(A, lda, a.V, W, ldw, Kpad, false, none, a, stages, s);
    if (ldw == 32) return launch_small<32, true>(A, lda, a.V, W, ldw, Kpad, false, none, a, stages, s);
    return 0;
}

int launch_gemm_nt_tc(const float *G, uint32_t ldg, uint64_t M, const float *W, uint32_t ldw, uint32_t prows, float *C,
                      uint32_t ldc, int stages, cudaStream_t s) {
    // C[M x prows] = G[M x ldw] . W^T with W [prows x ldw]: the contraction runs over W's pitch (<= 128: the K
    // loop of the small kernel), the output pitch is W's padded row count (32 or 64).
    if (ldw % BK != 0 || ldw == 0 || ldw > 128 || ldg != ldw || ldc != prows || M == 0) return 0;
    const SmallOut o{C, C, ldc, EPI_NONE};
    const SoftmaxCEArgs none{};
    if (prows == 64) return launch_small<64, false>(G, ldg, M, W, ldw, ldw, true, o, none, stages, s);
    if (prows == 32) return launch_small<32, false>(G, ldg, M, W, ldw, ldw, true, o, none, stages, s);
    return 0;
}

namespace {

template <int BN>
int launch_tn_bn(const float *A, uint32_t lda, uint32_t Mpad, const float *G, uint64_t K, float *C, float *ws,
                 size_t ws_floats, cudaStream_t s) {
    using L = SmemLayoutTN<BN>;
    const uint32_t mtiles = (Mpad + BM - 1) / BM;
    const size_t cfloats = (size_t)Mpad * BN;
    // enough splits for one CTA per SM (each holds 192 KB of shared memory), chunks of >= 256 vertices
    uint64_t nsplit = std::min<uint64_t>(std::max<uint64_t>(1, 148 / mtiles), (K + 255) / 256);
    nsplit = std::max<uint64_t>(1, std::min<uint64_t>(nsplit, ws_floats / cfloats));
    uint64_t kchunk = (K + nsplit - 1) / nsplit;
    kchunk = (kchunk + BK - 1) / BK * BK;
    nsplit = (K + kchunk - 1) / kchunk;
    if (ws_floats < cfloats * nsplit) return 0;
    Cache &c = cache();
    std::lock_guard<std::mutex> lock(c.mu);
    auto get_map = [&](const float *p, uint32_t cols, uint32_t ld) -> const CUtensorMap * {
        auto key = std::make_tuple(p, ld | 0x80000000u, K);  // high bit: 32 x 32 boxes (transposed use)
        auto it = c.amaps.find(key);
        if (it == c.amaps.end()) {
            CUtensorMap m;
            if (!make_map(&m, p, K, cols, ld, 32, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B)) return nullptr;
            it = c.amaps.emplace(key, m).first;
        }
        return &it->second;
    };
    const CUtensorMap *ma = get_map(A, Mpad, lda), *mg = get_map(G, BN, BN);
    if (!ma || !mg) return 0;
    static bool attr_set = false;
    if (!attr_set) {
        if (cudaFuncSetAttribute(gemm_tn_tc_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::kTotal) != cudaSuccess)
            return -1;
        attr_set = true;
    }
    dim3 grid(mtiles, (unsigned)nsplit);
    gemm_tn_tc_kernel<BN><<<grid, kThreads, L::kTotal, s>>>(*ma, *mg, nsplit > 1 ? ws : C, BN, Mpad, K, kchunk, cfloats);
    int launches = 1;
    if (nsplit > 1) {
        const size_t n4 = cfloats / 4;
        tn_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, s>>>(reinterpret_cast<const float4 *>(ws),
                                                                     reinterpret_cast<float4 *>(C), n4, n4, (int)nsplit);
        ++launches;
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

}  // namespace

int launch_gemm_tn_tc(const float *A, uint32_t lda, uint32_t Mpad, const float *G, uint32_t ldg, uint64_t K, float *C,
                      uint32_t ldc, float *ws, size_t ws_floats, cudaStream_t s) {
    // Supported: output pitch == G's pitch == 64 or 128, M (padded) a multiple of 32 within A's pitch,
    // a vertex range worth splitting.  Anything else takes the fp32 SIMT path (dense.cu).
    if (ldg != ldc || Mpad % 32 != 0 || Mpad == 0 || Mpad > lda || K < 1024 || !ws) return 0;
    if (ldg == 128) return launch_tn_bn<128>(A, lda, Mpad, G, K, C, ws, ws_floats, s);
    if (ldg == 64) return launch_tn_bn<64>(A, lda, Mpad, G, K, C, ws, ws_floats, s);
    return 0;
}

int launch_gemm_tc(const float *A, uint32_t lda, uint64_t M, const float *W, uint32_t ldw, uint32_t Kpad, float *C,
                   float *C2, uint32_t ldc, int epilogue, cudaStream_t s, int small_stages) {
    // Supported: K a multiple of the 32-float swizzle span, N (padded) exactly 64 or 128, and the
    // output pitch equal to N.  Anything else takes the fp32 SIMT path (dense.cu).
    if (Kpad % BK != 0 || Kpad == 0 || lda < Kpad || ldc != ldw || M == 0) return 0;
    if (Kpad / BK <= 4 && (ldw == 64 || ldw == 32) && small_stages >= 0) {  // skinny: several CTAs per SM instead of a deep ring
        const SmallOut o{C, C2 ? C2 : C, ldc, C2 ? epilogue : EPI_NONE};
        const SoftmaxCEArgs none{};
        return ldw == 64 ? launch_small<64, false>(A, lda, M, W, ldw, Kpad, false, o, none, small_stages, s)
                         : launch_small<32, false>(A, lda, M, W, ldw, Kpad, false, o, none, small_stages, s);
    }
    if (ldw == 128) return launch_bn<128>(A, lda, M, W, ldw, Kpad, C, C2, ldc, epilogue, s);
    if (ldw == 64) return launch_bn<64>(A, lda, M, W, ldw, Kpad, C, C2, ldc, epilogue, s);
    return 0;
}

}  // namespace dory
