// Neighbour aggregation: out[v,:] = self(v) + sum_{e in adj(v)} vals[e] * src[idx[e],:]
//
// Stands in for Engine::aggregateGCN / aggregateGAT (reference engine/ops/gcn_ops.cpp:130-191,
// gat_ops.cpp:173-243) and for the reference GPU backend's cusparseSpMM + cublasSgeam +
// cublasSdgmm + thrust::plus chain (GPU-Computation/comp_unit.cu:48-91), as ONE kernel with the
// self term fused.
//
// Mapping to B200:
//   * a destination row is owned by one warp (rows below the heavy threshold) or by one CTA of
//     8 warps (heavy rows; the warps split the edge list and combine through shared memory in a
//     fixed order) -- no atomics, so results are bit-reproducible run to run;
//   * lanes are grouped LG per source row and read VEC float4 each: every gather of a source row
//     is a run of fully coalesced 128-bit loads covering whole 128 B lines (rows are padded to
//     a line multiple, common.cuh), 32/LG edges are in flight per instruction and U such
//     instructions are issued back to back before the first FMA (memory-level parallelism);
//   * edge ids / weights are read once per 32 edges with one coalesced 128 B load each and
//     distributed by warp shuffles; they bypass L1 and are marked evict-first in L2 (streamed
//     once per slab) while feature rows are marked evict-last, so that L2 keeps the slab;
//   * rows are issued heaviest first (longest-processing-time-first), so the power-law tail
//     fills in behind the hubs instead of stretching the last wave (engine.cu: build_row_lists);
//   * wide rows are cut into 128-float column slabs (gridDim.y) and the source rows into windows
//     (ptr_stride / ptr_off / ptr_span, one launch per window group): a pass gathers from
//     rows x 512 B ~ 60 MB, which is what stays resident in L2.  Walking whole 2.4 KB rows of the
//     561 MB layer-0 block instead misses L2 78 % of the time and is bound by 193 GB of DRAM
//     re-reads per aggregation (profiles/round1_spmm_v1_full.md vs round1_final_full.md);
//   * low-degree rows take spmm_group_kernel (a lane group per row) instead of a warp per row.
#include <cooperative_groups.h>

#include <algorithm>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace dory {
namespace {

constexpr int kWarpsPerCta = 8;
constexpr unsigned kFull = 0xffffffffu;

__device__ __forceinline__ uint64_t policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint32_t ld_stream_u32(const uint32_t *p, uint64_t pol) {
    uint32_t v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float ld_stream_f32(const float *p, uint64_t pol) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ float4 ld_row_f4(const float4 *p, uint64_t pol) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void st_stream_f4(float4 *p, const float4 &v, uint64_t pol) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(p), "f"(v.x), "f"(v.y),
                 "f"(v.z), "f"(v.w), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void fma4(float4 &a, const float4 &x, float w) {
    a.x = fmaf(x.x, w, a.x);
    a.y = fmaf(x.y, w, a.y);
    a.z = fmaf(x.z, w, a.z);
    a.w = fmaf(x.w, w, a.w);
}
__device__ __forceinline__ void add4(float4 &a, const float4 &b) {
    a.x += b.x;
    a.y += b.y;
    a.z += b.z;
    a.w += b.w;
}

// LG   : lanes cooperating on one source row (power of two, 4..32)
// VEC  : float4 per lane  -> a CTA column slab is LG*VEC float4 wide
// TEAM : warps per destination row (1: warp-per-row, kWarpsPerCta: CTA-per-row)
// CL   : CL > 1 (TEAM == kWarpsPerCta only) is launched as thread-block clusters of CL CTAs.  The
//        first a.n_vheavy clusters each own ONE hub row: the CL * 8 warps split its edge list, every
//        CTA reduces its warps in shared memory and rank 0 then adds the CTA partials through
//        distributed shared memory in rank order -- still no atomics, still the same bits every run.
//        The remaining clusters are just CL independent CTA-per-row rows.  A hub row's edge list is
//        the longest dependent chain of the launch (a warp retires ~4 edges per L2 round trip); at 8
//        partitions a 27 K-edge row on one CTA outlasted the rest of its launch
//        (profiles/round1_ops_n8.md).  Hubs and ordinary heavy rows share one launch so that the few
//        hub clusters do not run alone on an otherwise idle chip.
// U    : gather instructions issued back to back before their FMAs (each covers 32/LG edges)
// OCC  : CTAs of 8 warps ptxas must fit per SM (register budget 65536 / (256 * OCC) per thread);
//        the kernel is bound by bytes in flight, so resident warps are worth more than registers
template <int LG, int VEC, int TEAM, int U, int OCC, int CL = 1>
__global__ void __launch_bounds__(32 * kWarpsPerCta, OCC)
spmm_kernel(const SpmmArgs a, const uint32_t *__restrict__ rowlist, uint32_t nrows) {
    static_assert(CL == 1 || TEAM == kWarpsPerCta, "clusters split CTA-per-row rows only");
    int warps = TEAM;  // warps walking this row
    bool hub = false;
    constexpr int EPW = 32 / LG;  // edges covered by one warp-wide gather instruction
    static_assert(U >= 1 && (LG % U == 0 || U % LG == 0), "U must divide the number of steps per 32-edge batch");
    constexpr int UU = U < LG ? U : LG;  // steps per inner block (a batch of 32 edges has LG steps)
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LG, l = lane % LG;

    uint32_t rid;
    int team_rank;
    if (TEAM == 1) {
        rid = blockIdx.x * kWarpsPerCta + warp;
        team_rank = 0;
        if (rid >= nrows) return;
    } else if (CL == 1) {
        rid = blockIdx.x;
        team_rank = warp;
    } else {
        const uint32_t cid = blockIdx.x / CL, crank = cg::this_cluster().block_rank();
        hub = cid < a.n_vheavy;
        if (hub) {
            rid = cid;
            team_rank = (int)crank * TEAM + warp;
            warps = TEAM * CL;
        } else {
            rid = a.n_vheavy + (cid - a.n_vheavy) * CL + crank;
            team_rank = warp;
            if (rid >= nrows) return;  // padding CTA of the last non-hub cluster (never cluster-syncs)
        }
    }
    const uint32_t row = rowlist ? rowlist[rid] : a.low + rid;
    const uint32_t col0 = blockIdx.y * (LG * VEC);  // slab start, float4 units
    const uint64_t pbase = (uint64_t)row * a.ptr_stride + a.ptr_off;
    const uint64_t e_begin = a.ptrs[pbase], e_end = a.ptrs[pbase + a.ptr_span];
    const float4 *__restrict__ src4 = reinterpret_cast<const float4 *>(a.src);
    const uint32_t ld4 = a.ld >> 2;
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();

    float4 acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    bool act[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) act[j] = (col0 + l + j * LG) < a.nvec;

    // ids / weights of the NEXT 32-edge batch are requested before the current batch is walked, so
    // their (HBM-streamed) latency is off the dependent chain id -> shuffle -> gather
    uint32_t s_n = 0;
    float w_n = 0.f;
    {
        const uint64_t my = e_begin + (uint64_t)team_rank * 32 + lane;
        if (my < e_end) {
            s_n = ld_stream_u32(a.idx + my, pol_stream);
            w_n = ld_stream_f32(a.vals + my, pol_stream);
        }
    }
    for (uint64_t e0 = e_begin + (uint64_t)team_rank * 32; e0 < e_end; e0 += 32 * (uint64_t)warps) {
        const uint32_t s_l = s_n;
        const float w_l = w_n;
        {
            const uint64_t nx = e0 + 32 * (uint64_t)warps + lane;
            s_n = 0;
            w_n = 0.f;
            if (nx < e_end) {
                s_n = ld_stream_u32(a.idx + nx, pol_stream);
                w_n = ld_stream_f32(a.vals + nx, pol_stream);
            }
        }
        const int n = (int)min((uint64_t)32, e_end - e0);
#pragma unroll 1
        for (int k0 = 0; k0 < LG; k0 += UU) {
            if (k0 * EPW >= n) break;  // warp-uniform
            float4 x[UU][VEC];
            float w[UU];
            // issue every gather of the block first ...
#pragma unroll
            for (int u = 0; u < UU; ++u) {
                const int sl = (k0 + u) * EPW + g;
                const uint32_t s = __shfl_sync(kFull, s_l, sl);
                w[u] = __shfl_sync(kFull, w_l, sl);
                const bool ev = sl < n;  // a padding edge never touches memory (0 * Inf would be NaN)
                const float4 *rp = src4 + (size_t)s * ld4 + col0 + l;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    x[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (act[j] && ev) x[u][j] = ld_row_f4(rp + j * LG, pol_keep);
                }
            }
            // ... then consume them in edge order
#pragma unroll
            for (int u = 0; u < UU; ++u)
#pragma unroll
                for (int j = 0; j < VEC; ++j) fma4(acc[j], x[u][j], w[u]);
        }
    }

    // combine the EPW edge groups of the warp (fixed butterfly order)
#pragma unroll
    for (int off = LG; off < 32; off <<= 1) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            acc[j].x += __shfl_xor_sync(kFull, acc[j].x, off);
            acc[j].y += __shfl_xor_sync(kFull, acc[j].y, off);
            acc[j].z += __shfl_xor_sync(kFull, acc[j].z, off);
            acc[j].w += __shfl_xor_sync(kFull, acc[j].w, off);
        }
    }

    auto self_term = [&](size_t o) -> float4 {
        float4 self = make_float4(0.f, 0.f, 0.f, 0.f);
        if (a.self_mode == SELF_NORM) {
            const float sw = a.selfw[row];
            const float4 x = src4[o];
            self = make_float4(x.x * sw, x.y * sw, x.z * sw, x.w * sw);
        } else if (a.self_mode == SELF_ONE) {
            self = src4[o];
        } else if (a.self_mode == SELF_ACCUM) {
            self = reinterpret_cast<const float4 *>(a.out)[o];
        }
        return self;
    };

    if (TEAM > 1) {
        __shared__ float4 part[TEAM][LG * VEC];
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) part[warp][l + j * LG] = acc[j];
        }
        __syncthreads();
        if (CL == 1 || !hub) {
            // every thread of the CTA finishes a strided share of the slab's columns
            for (int c = threadIdx.x; c < LG * VEC; c += 32 * TEAM) {
                if (col0 + c >= a.nvec) continue;
                float4 t = part[0][c];
#pragma unroll
                for (int wv = 1; wv < TEAM; ++wv) add4(t, part[wv][c]);
                const size_t o = (size_t)row * ld4 + col0 + c;
                float4 self = self_term(o);
                add4(self, t);
                st_stream_f4(reinterpret_cast<float4 *>(a.out) + o, self, pol_stream);
            }
        } else {
            // CTA partial -> part[0]; rank 0 adds the peers' partials over DSMEM in rank order
            cg::cluster_group cl = cg::this_cluster();
            for (int c = threadIdx.x; c < LG * VEC; c += 32 * TEAM) {
                float4 t = part[0][c];
#pragma unroll
                for (int wv = 1; wv < TEAM; ++wv) add4(t, part[wv][c]);
                part[0][c] = t;  // column c is read and written by this thread only
            }
            cl.sync();
            if (cl.block_rank() == 0) {
                for (int c = threadIdx.x; c < LG * VEC; c += 32 * TEAM) {
                    if (col0 + c >= a.nvec) continue;
                    float4 t = part[0][c];
                    for (int r = 1; r < CL; ++r) add4(t, cl.map_shared_rank(&part[0][0], r)[c]);
                    const size_t o = (size_t)row * ld4 + col0 + c;
                    float4 self = self_term(o);
                    add4(self, t);
                    st_stream_f4(reinterpret_cast<float4 *>(a.out) + o, self, pol_stream);
                }
            }
            cl.sync();  // peers keep their shared memory alive until rank 0 has read it
        }
    } else if (g == 0) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            if (!act[j]) continue;
            const size_t o = (size_t)row * ld4 + col0 + l + j * LG;
            float4 self = self_term(o);
            add4(self, acc[j]);
            st_stream_f4(reinterpret_cast<float4 *>(a.out) + o, self, pol_stream);
        }
    }
}

// Low-degree rows (Amazon / Friendster shapes: ~25 edges per vertex).  A warp that owns ONE such row
// spends its life in three dependent latencies (offsets -> ids -> rows) for a single 32-edge batch.
// Here every LANE GROUP of LG lanes owns its own row (32/LG rows per warp) and walks it edge by edge
// with no cross-group reduction at all -- and in the reference's own summation order (self term,
// then edges in file order).  The row must fit one slab: nvec <= LG * VEC.
//   * ids / weights: the LG lanes of a group read LG consecutive edges with one load each (a
//     contiguous 4*LG-byte piece per group) and hand them round by shuffles inside the group; the
//     NEXT batch of LG is requested before the current one is gathered, so the id latency is off the
//     dependent chain id -> address -> row.  These loads allocate in L1 with the default L2 policy:
//     a group consumes its row's 128 B line of ids 16 B at a time over several iterations, and with
//     the streaming (no-allocate, evict-first) loads of the warp-per-row kernel every piece went back
//     to HBM -- 7.1 GB of DRAM reads for 1.5 GB of compulsory bytes (profiles/round1_lowdeg.md);
//   * U gathers are in flight per group before their FMAs.
template <int LG, int VEC, int U, int OCC>
__global__ void __launch_bounds__(32 * kWarpsPerCta, OCC)
spmm_group_kernel(const SpmmArgs a, const uint32_t *__restrict__ rowlist, uint32_t nrows) {
    static_assert(LG % U == 0, "U must divide the id batch");
    constexpr int G = 32 / LG;  // rows per warp
    const int lane = threadIdx.x & 31;
    const int g = lane / LG, l = lane % LG;
    const uint32_t rid = (blockIdx.x * kWarpsPerCta + (threadIdx.x >> 5)) * G + g;
    const bool live = rid < nrows;
    const uint32_t row = live ? (rowlist ? rowlist[rid] : a.low + rid) : 0;
    const float4 *__restrict__ src4 = reinterpret_cast<const float4 *>(a.src);
    const uint32_t ld4 = a.ld >> 2;
    const uint64_t pol_keep = policy_evict_last();
    uint64_t e = 0, e_end = 0;
    if (live) {
        const uint64_t pbase = (uint64_t)row * a.ptr_stride + a.ptr_off;
        e = a.ptrs[pbase];
        e_end = a.ptrs[pbase + a.ptr_span];
    }
    uint32_t s_n = 0;
    float w_n = 0.f;
    if (e + l < e_end) {
        s_n = __ldg(a.idx + e + l);
        w_n = __ldg(a.vals + e + l);
    }
    bool act[VEC];
    float4 acc[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
        act[j] = live && (uint32_t)(l + j * LG) < a.nvec;
        acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    // self term first, like the reference (gcn_ops.cpp:166-171)
    if (a.self_mode != SELF_ZERO) {
        const float sw = a.self_mode == SELF_NORM ? a.selfw[row] : 1.f;
        const float4 *base = a.self_mode == SELF_ACCUM ? reinterpret_cast<const float4 *>(a.out) : src4;
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            if (act[j]) {
                const float4 x = base[(size_t)row * ld4 + l + j * LG];
                acc[j] = make_float4(x.x * sw, x.y * sw, x.z * sw, x.w * sw);
            }
    }
    // groups of one warp have different trip counts: the loop runs to the longest row of the warp
    // (rows are issued in degree classes, so the spread is < 2x)
    while (__any_sync(kFull, e < e_end)) {
        const uint32_t s_c = s_n;
        const float w_c = w_n;
        s_n = 0;
        w_n = 0.f;
        if (e + LG + l < e_end) {
            s_n = __ldg(a.idx + e + LG + l);
            w_n = __ldg(a.vals + e + LG + l);
        }
#pragma unroll
        for (int k0 = 0; k0 < LG; k0 += U) {
            float4 x[U][VEC];
            float w[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int from = g * LG + k0 + u;
                const uint32_t s = __shfl_sync(kFull, s_c, from);
                w[u] = __shfl_sync(kFull, w_c, from);
                const bool ev = e + (k0 + u) < e_end;  // a padding edge never touches memory
                const float4 *rp = src4 + (size_t)s * ld4 + l;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    x[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (act[j] && ev) x[u][j] = ld_row_f4(rp + j * LG, pol_keep);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int j = 0; j < VEC; ++j) fma4(acc[j], x[u][j], w[u]);
        }
        e += LG;
    }
#pragma unroll
    for (int j = 0; j < VEC; ++j)
        if (act[j]) reinterpret_cast<float4 *>(a.out)[(size_t)row * ld4 + l + j * LG] = acc[j];
}

template <int LG, int VEC, int U, int OCC>
int launch_group(const SpmmArgs &a, cudaStream_t s) {
    constexpr int G = 32 / LG;
    const uint32_t rowsPerCta = kWarpsPerCta * G;
    spmm_group_kernel<LG, VEC, U, OCC><<<(a.n_light + rowsPerCta - 1) / rowsPerCta, 32 * kWarpsPerCta, 0, s>>>(
        a, a.light, a.n_light);
    return cudaGetLastError() == cudaSuccess ? 1 : -1;
}

// Light rows through the lane-group kernel; returns 0 when the row is too wide for it.
int launch_light_groups(const SpmmArgs &a, cudaStream_t s) {
    const uint32_t n = a.nvec;
    if (n <= 4) return launch_group<4, 1, 4, 6>(a, s);
    if (n <= 8) return launch_group<4, 2, 4, 4>(a, s);
    if (n <= 12) return launch_group<4, 3, 2, 4>(a, s);
    if (n <= 16) return launch_group<4, 4, 2, 4>(a, s);
    if (n <= 32) return launch_group<8, 4, 2, 4>(a, s);
    return 0;
}

constexpr int kClusterCtas = 8;  // CTAs sharing one very heavy row (portable cluster size limit)

template <int LG, int VEC, int U, int OCC>
int launch_cfg(const SpmmArgs &a, cudaStream_t s) {
    int launches = 0;
    const uint32_t slab = LG * VEC;
    const uint32_t nslab = (a.nvec + slab - 1) / slab;
    const uint32_t n_vheavy = a.heavy ? std::min(a.n_vheavy, a.n_heavy) : 0;
    if (n_vheavy) {  // hub rows present: the whole heavy launch goes out as clusters of 8 CTAs
        SpmmArgs h = a;
        h.n_vheavy = n_vheavy;
        const uint32_t clusters = n_vheavy + (a.n_heavy - n_vheavy + kClusterCtas - 1) / kClusterCtas;
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = dim3(clusters * kClusterCtas, nslab);
        cfg.blockDim = dim3(32 * kWarpsPerCta);
        cfg.dynamicSmemBytes = 0;
        cfg.stream = s;
        cudaLaunchAttribute attr{};
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = kClusterCtas;
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        if (cudaLaunchKernelEx(&cfg, spmm_kernel<LG, VEC, kWarpsPerCta, U, OCC, kClusterCtas>, h, a.heavy, a.n_heavy) !=
            cudaSuccess)
            return -1;
        ++launches;
    } else if (a.n_heavy) {
        dim3 grid(a.n_heavy, nslab);
        spmm_kernel<LG, VEC, kWarpsPerCta, U, OCC><<<grid, 32 * kWarpsPerCta, 0, s>>>(a, a.heavy, a.n_heavy);
        ++launches;
    }
    if (a.n_light) {
        dim3 grid((a.n_light + kWarpsPerCta - 1) / kWarpsPerCta, nslab);
        spmm_kernel<LG, VEC, 1, U, OCC><<<grid, 32 * kWarpsPerCta, 0, s>>>(a, a.light, a.n_light);
        ++launches;
    }
    if (cudaGetLastError() != cudaSuccess) return -1;
    return launches;
}

template <int LG, int VEC, int U>
int launch_occ(const SpmmArgs &a, int occ, cudaStream_t s) {
    if (occ >= 8) return launch_cfg<LG, VEC, U, 8>(a, s);
    if (occ >= 6) return launch_cfg<LG, VEC, U, 6>(a, s);
    if (occ >= 5) return launch_cfg<LG, VEC, U, 5>(a, s);
    return launch_cfg<LG, VEC, U, 4>(a, s);
}

template <int LG, int VEC>
int launch_unroll(const SpmmArgs &a, int unroll, int occ, cudaStream_t s) {
    if (unroll >= 2) return launch_occ<LG, VEC, 2>(a, occ, s);
    return launch_occ<LG, VEC, 1>(a, occ, s);
}

}  // namespace

// (lanes per row, float4 per lane) pairs that have a kernel instantiation below
bool spmm_shape_supported(int lg, int vec) {
    if (lg == 0 || vec == 0) return lg == 0 && vec == 0;  // both derived from the row width, or both given
    switch (lg) {
    case 4: return vec == 1 || vec == 2 || vec == 4;
    case 8: return vec == 1 || vec == 2 || vec == 4;
    case 16: return vec == 1 || vec == 2;
    case 32: return vec == 1 || vec == 2 || vec == 4;
    default: return false;
    }
}

#define DORY_SPMM_CASE(LG_, VEC_) \
    if (lg == LG_ && vec == VEC_) return launch_unroll<LG_, VEC_>(a, unroll, occ, s)

int launch_spmm(const SpmmArgs &a, cudaStream_t s) {
    // low-degree light rows that fit one slab: a lane group per row (cfg_light: 0 auto, 1 warp, 2 group)
    const bool fits = a.nvec <= 32;
    const bool groups = a.n_light && fits && (a.cfg_light == 2 || (a.cfg_light == 0 && a.light_avg_degree < 96));
    if (!groups) return launch_spmm_rows(a, s);
    int launches = 0;
    if (a.n_heavy) {
        SpmmArgs h = a;
        h.n_light = 0;
        launches = launch_spmm_rows(h, s);
        if (launches < 0) return -1;
    }
    const int n = launch_light_groups(a, s);
    return n < 0 ? -1 : launches + n;
}

int launch_spmm_rows(const SpmmArgs &a, cudaStream_t s) {
    int lg = a.cfg_lg, vec = a.cfg_vec, unroll = a.cfg_unroll;
    if (lg == 0 || vec == 0) {
        // Default (tools/spmm_sweep.py on the Reddit shape, profiles/): 8 lanes x 4 float4 per
        // gathered row = 128-float slabs, 4 edges per gather instruction.  Narrow rows use fewer
        // lanes so that no lane idles.
        const uint32_t n = a.nvec;
        if (n <= 4) lg = 4, vec = 1;
        else if (n <= 8) lg = 8, vec = 1;
        else if (n <= 16) lg = 8, vec = 2;
        else lg = 8, vec = 4;
    }
    if (unroll == 0) unroll = 1;
    const int occ = a.cfg_occ ? a.cfg_occ : 4;
    if (!spmm_shape_supported(lg, vec)) return -2;  // dory_set_option refuses these; defensive
    DORY_SPMM_CASE(4, 1);
    DORY_SPMM_CASE(4, 2);
    DORY_SPMM_CASE(4, 4);  // 8 edges per gather instruction: candidate for 33..64-float rows (not a default yet)
    DORY_SPMM_CASE(8, 1);
    DORY_SPMM_CASE(8, 2);
    DORY_SPMM_CASE(8, 4);
    DORY_SPMM_CASE(16, 1);
    DORY_SPMM_CASE(16, 2);
    DORY_SPMM_CASE(32, 1);
    DORY_SPMM_CASE(32, 2);
    DORY_SPMM_CASE(32, 4);
    return -1;
}

}  // namespace dory
