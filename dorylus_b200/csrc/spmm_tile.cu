// Shared-memory-staged neighbour aggregation (same operator as spmm.cu: Engine::aggregateGCN,
// reference engine/ops/gcn_ops.cpp:130-191) for graphs whose vertex numbering has locality.
//
// A CTA owns one TILE of destination rows (tile_plan.h).  It stages the tile's source WINDOW -- the
// run of consecutive source rows that holds most of the tile's edges, one column slab of it -- into
// shared memory with bulk TMA copies (cp.async.bulk, completion on an mbarrier; SASS: UBLKCP), then
// walks every row's edge list in two parts: the edges whose source lies in the window read their rows
// from shared memory (128 B/clk per SM, no L1 tag look-ups, no L2 traffic), the rest gather from L2
// like spmm.cu.  The window's bytes cross L2 -> SM once per tile instead of once per edge.
//
//   * high-degree graphs (Reddit shape): rows above `teamDegree` edges are walked by all 16 warps of the
//     CTA together (edge list split 32 edges per warp, partials combined through shared memory in warp
//     order), the others by one warp each, handed out longest first through a shared counter;
//   * low-degree graphs (Amazon / Friendster shapes): a lane group per row, 32 / LG rows per warp, the
//     reference's own summation order -- the row must fit one column slab;
//   * no atomics anywhere: the result is bit-reproducible run to run.
#include <cuda.h>

#include <algorithm>
#include <cstdio>
#include <map>
#include <mutex>
#include <tuple>

#include "common.cuh"

namespace dory {
namespace {

constexpr int kTileWarps = 16;
constexpr unsigned kFull = 0xffffffffu;

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
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// 1-D bulk copy global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}
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
__device__ __forceinline__ float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
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

// 2-D tiled TMA load (cp.async.bulk.tensor): box {boxCols floats, kBoxRows rows} at (col, row) of the source block
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap *map, int col, int row, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(col), "r"(row)
        : "memory");
}

constexpr uint32_t kBoxRows = 64;  // rows per TMA box of a column-slab window

// Bytes the staged window occupies in shared memory: whole rows are one contiguous run; a column slab is
// fetched in boxes of kBoxRows rows (the last one may overhang the window; the buffer has room for it).
__host__ __device__ __forceinline__ uint32_t window_rows_padded(uint32_t wrows, bool contiguous) {
    return contiguous ? wrows : (wrows + kBoxRows - 1) / kBoxRows * kBoxRows;
}

// Posts the copies of the window (`wrows` rows of `slab4` float4, source rows wlo.. at pitch ld4, column
// col0) on `bar`; returns the bytes they will deliver.  Whole rows: 4 KB bulk copies issued by all threads.
// Column slab: one tiled TMA load per kBoxRows rows through the tensor map of (source block, slab width).
__device__ __forceinline__ uint32_t window_bytes(uint32_t ld4, uint32_t slab4, uint32_t wrows) {
    return window_rows_padded(wrows, slab4 == ld4) * slab4 * 16u;
}
__device__ __forceinline__ void issue_window(float4 *win, const float4 *src4, const CUtensorMap *map, uint32_t ld4, uint32_t col0,
                                             uint32_t slab4, uint32_t wlo, uint32_t wrows, uint32_t bar) {
    if (slab4 == ld4) {
        const uint32_t total4 = wrows * ld4;
        const float4 *base = src4 + (size_t)wlo * ld4;
        for (uint32_t c = threadIdx.x * 256u; c < total4; c += blockDim.x * 256u)
            bulk_g2s(smem_u32(win + c), base + c, min(256u, total4 - c) * 16u, bar);
    } else {
        const uint32_t nbox = (wrows + kBoxRows - 1) / kBoxRows;
        for (uint32_t b = threadIdx.x; b < nbox; b += blockDim.x)
            tma_load_2d(smem_u32(win + (size_t)b * kBoxRows * slab4), map, (int)(col0 * 4), (int)(wlo + b * kBoxRows), bar);
    }
}

// One part of a row's edge list walked by `warps` warps (this one is number `rank`): 32 edges per warp
// per step, ids / weights of the next step requested before the current one is gathered.
// SMEM: source rows come from the staged window (row s at win + (s - sub) * stride4), else from `base`.
template <int LG, int VEC, bool SMEM>
__device__ __forceinline__ void walk_part(float4 (&acc)[VEC], const bool (&act)[VEC], uint64_t e_begin, uint64_t e_end, int rank,
                                          int warps, const uint32_t *__restrict__ idx, const float *__restrict__ vals,
                                          const float4 *base, uint32_t stride4, uint32_t sub, uint64_t pol_stream,
                                          uint64_t pol_keep) {
    constexpr int EPW = 32 / LG;
    const int lane = threadIdx.x & 31;
    const int g = lane / LG, l = lane % LG;
    const uint32_t smem_base = SMEM ? smem_u32(base) : 0u;
    uint32_t s_n = 0;
    float w_n = 0.f;
    {
        const uint64_t my = e_begin + (uint64_t)rank * 32 + lane;
        if (my < e_end) {
            s_n = ld_stream_u32(idx + my, pol_stream);
            w_n = ld_stream_f32(vals + my, pol_stream);
        }
    }
    for (uint64_t e0 = e_begin + (uint64_t)rank * 32; e0 < e_end; e0 += 32 * (uint64_t)warps) {
        const uint32_t s_l = s_n;
        const float w_l = w_n;
        {
            const uint64_t nx = e0 + 32 * (uint64_t)warps + lane;
            s_n = 0;
            w_n = 0.f;
            if (nx < e_end) {
                s_n = ld_stream_u32(idx + nx, pol_stream);
                w_n = ld_stream_f32(vals + nx, pol_stream);
            }
        }
        const int n = (int)min((uint64_t)32, e_end - e0);
        constexpr int UU = SMEM ? (LG < 2 ? LG : 2) : 1;  // two steps in flight from shared memory
#pragma unroll 1
        for (int k0 = 0; k0 < LG; k0 += UU) {
            if (k0 * EPW >= n) break;  // warp-uniform
            float4 x[UU][VEC];
            float w[UU];
#pragma unroll
            for (int u = 0; u < UU; ++u) {
                const int sl = (k0 + u) * EPW + g;
                const uint32_t s = __shfl_sync(kFull, s_l, sl);
                w[u] = __shfl_sync(kFull, w_l, sl);
                const bool ev = sl < n;  // a padding edge never touches memory
                const float4 *rp = base + (size_t)(s - sub) * stride4 + l;  // global rows (SMEM: unused)
                const uint32_t sp = smem_base + ((s - sub) * stride4 + l) * 16u;
#pragma unroll
                for (int j = 0; j < VEC; ++j) {
                    x[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (act[j] && ev) x[u][j] = SMEM ? lds_f4(sp + j * LG * 16u) : ld_row_f4(rp + j * LG, pol_keep);
                }
            }
#pragma unroll
            for (int u = 0; u < UU; ++u)
#pragma unroll
                for (int j = 0; j < VEC; ++j) fma4(acc[j], x[u][j], w[u]);
        }
    }
}

__device__ __forceinline__ float4 self_term(const SpmmArgs &a, uint32_t row, size_t o) {
    const float4 *src4 = reinterpret_cast<const float4 *>(a.src);
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
}

// ---------------------------------------------------------------------------- high-degree tiles
template <int LG, int VEC, int OCC>
__global__ void __launch_bounds__(32 * kTileWarps, OCC)
spmm_tile_kernel(const SpmmArgs a, const TilePlanDev t, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) float4 win[];
    __shared__ __align__(8) uint64_t bar_mem;
    __shared__ uint32_t next_row;
    __shared__ float4 part[kTileWarps][LG * VEC];
    constexpr int SLAB = LG * VEC;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LG, l = lane % LG;
    const uint32_t tile = blockIdx.x;
    const uint32_t col0 = blockIdx.y * SLAB;  // slab start, float4 units
    const uint32_t ld4 = a.ld >> 2;
    // window row pitch: the whole row when it fits one slab, else a full slab for EVERY slab -- the last one
    // overhangs the row pitch and its TMA box is zero-filled there (the box shape is fixed per launch)
    const uint32_t slab4 = min((uint32_t)SLAB, ld4);
    const uint32_t wlo = t.tile_wlo[tile], wrows = t.tile_wrows[tile];
    const uint32_t r_begin = t.tile_ptr[tile], r_end = t.tile_ptr[tile + 1];
    const uint32_t n_team = min(t.tile_team[tile], r_end - r_begin);
    const float4 *__restrict__ src4 = reinterpret_cast<const float4 *>(a.src);
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        next_row = r_begin + n_team;
    }
    if (wrows && threadIdx.x == 0) mbar_expect_tx(bar, window_bytes(ld4, slab4, wrows));
    __syncthreads();
    if (wrows) {
        issue_window(win, src4, &tmap, ld4, col0, slab4, wlo, wrows, bar);
        mbar_wait(bar, 0);
    }
    const uint64_t pol_stream = policy_evict_first();
    const uint64_t pol_keep = policy_evict_last();
    bool act[VEC];
#pragma unroll
    for (int j = 0; j < VEC; ++j) act[j] = (col0 + l + j * LG) < a.nvec;

    auto walk_row = [&](uint32_t row, int rank, int warps, float4(&acc)[VEC]) {
#pragma unroll
        for (int j = 0; j < VEC; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint64_t e0 = t.ptrs[2 * (size_t)row], e1 = t.ptrs[2 * (size_t)row + 1], e2 = t.ptrs[2 * (size_t)row + 2];
        walk_part<LG, VEC, true>(acc, act, e0, e1, rank, warps, t.idx, t.vals, win, slab4, wlo, pol_stream, pol_keep);
        walk_part<LG, VEC, false>(acc, act, e1, e2, rank, warps, t.idx, t.vals, src4 + col0, ld4, 0u, pol_stream, pol_keep);
        // combine the edge groups of the warp (fixed butterfly order)
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
    };

    // rows the whole CTA walks together
    for (uint32_t i = r_begin; i < r_begin + n_team; ++i) {
        const uint32_t row = t.rows[i];
        float4 acc[VEC];
        walk_row(row, warp, kTileWarps, acc);
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) part[warp][l + j * LG] = acc[j];
        }
        __syncthreads();
        for (int c = threadIdx.x; c < SLAB; c += 32 * kTileWarps) {
            if (col0 + c >= a.nvec) continue;
            float4 s = part[0][c];
#pragma unroll
            for (int wv = 1; wv < kTileWarps; ++wv) add4(s, part[wv][c]);
            const size_t o = (size_t)row * ld4 + col0 + c;
            float4 self = self_term(a, row, o);
            add4(self, s);
            reinterpret_cast<float4 *>(a.out)[o] = self;
        }
        __syncthreads();
    }
    // the others: a warp each, handed out in the tile's (degree-descending) order
    for (;;) {
        uint32_t i = 0;
        if (lane == 0) i = atomicAdd(&next_row, 1u);
        i = __shfl_sync(kFull, i, 0);
        if (i >= r_end) break;
        const uint32_t row = t.rows[i];
        float4 acc[VEC];
        walk_row(row, 0, 1, acc);
        if (g == 0) {
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                if (!act[j]) continue;
                const size_t o = (size_t)row * ld4 + col0 + l + j * LG;
                float4 self = self_term(a, row, o);
                add4(self, acc[j]);
                reinterpret_cast<float4 *>(a.out)[o] = self;
            }
        }
    }
}

// ---------------------------------------------------------------------------- low-degree tiles
// A lane group of LG lanes per row, G = 32 / LG rows per warp (spmm_group_kernel's mapping): self term
// first, then the in-window edges, then the rest -- per group in the row's own edge order within each part.
// The row fits one slab (nvec <= LG * VEC).  Besides the window, the tile's offsets, edge ids and edge
// weights are staged too: a tile's rows are consecutive, so each of the three is ONE contiguous run of the
// plan's arrays and one bulk copy.  The walk then reads nothing but shared memory, except the rows of the
// out-of-window edges (L2) -- no dependent chain of global loads per row (offsets -> ids -> rows), which is
// what bounds the gather kernel at this degree.
template <int LG, int VEC, int WARPS, int OCC>
__global__ void __launch_bounds__(32 * WARPS, OCC)
spmm_tile_group_kernel(const SpmmArgs a, const TilePlanDev t, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t bar_mem;
    constexpr int G = 32 / LG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LG, l = lane % LG;
    const uint32_t tile = blockIdx.x;
    const uint32_t ld4 = a.ld >> 2;
    const uint32_t slab4 = min((uint32_t)(LG * VEC), ld4);
    const uint32_t wlo = t.tile_wlo[tile], wrows = t.tile_wrows[tile];
    const uint32_t r_begin = t.tile_ptr[tile], nrows = t.tile_ptr[tile + 1] - r_begin;
    const uint64_t e0 = t.tile_e0[tile], e1 = t.tile_e1[tile];
    const uint32_t row0 = t.rows[r_begin];  // rows of a tile are consecutive: row0 .. row0 + nrows - 1
    const float4 *__restrict__ src4 = reinterpret_cast<const float4 *>(a.src);
    // shared-memory layout: window | offsets | ids | weights
    float4 *win = reinterpret_cast<float4 *>(smem);
    uint64_t *sptr = reinterpret_cast<uint64_t *>(smem + t.smem_ptr_off);
    uint32_t *sidx = reinterpret_cast<uint32_t *>(smem + t.smem_idx_off);
    float *sval = reinterpret_cast<float *>(smem + t.smem_val_off);
    const uint64_t ea = e0 & ~(uint64_t)3;                       // 16-byte aligned start of the edge run
    const uint32_t ebytes = (uint32_t)(((e1 + 3) & ~(uint64_t)3) - ea) * 4u;
    const uint32_t pbytes = (2u * nrows + 2u) * 8u;              // offsets 2*row0 .. 2*(row0+nrows), padded to 16 B
    const uint32_t bar = smem_u32(&bar_mem);
    if (threadIdx.x == 0) {
        mbar_init(bar, 1);
        fence_barrier_init();
        mbar_expect_tx(bar, (wrows ? window_bytes(ld4, slab4, wrows) : 0u) + pbytes + 2u * ebytes);
    }
    __syncthreads();
    if (wrows) issue_window(win, src4, &tmap, ld4, 0u, slab4, wlo, wrows, bar);
    if (warp == WARPS - 1) {  // the last warp posts the three edge-data runs (32 KB pieces)
        if (lane == 0) bulk_g2s(smem_u32(sptr), t.ptrs + 2 * (size_t)row0, pbytes, bar);
        for (uint32_t c = lane * 32768u; c < ebytes; c += 32u * 32768u) {
            const uint32_t n = min(32768u, ebytes - c);
            bulk_g2s(smem_u32(sidx) + c, reinterpret_cast<const uint8_t *>(t.idx + ea) + c, n, bar);
            bulk_g2s(smem_u32(sval) + c, reinterpret_cast<const uint8_t *>(t.vals + ea) + c, n, bar);
        }
    }
    mbar_wait(bar, 0);
    const uint64_t pol_keep = policy_evict_last();
    const uint32_t win_base = smem_u32(win);

    for (uint32_t i0 = (uint32_t)warp * G; i0 < nrows; i0 += WARPS * G) {
        const uint32_t i = i0 + g;
        const bool live = i < nrows;
        const uint32_t row = row0 + (live ? i : 0u);
        uint32_t e = 0, e_mid = 0, e_end = 0;  // positions inside the staged run
        if (live) {
            e = (uint32_t)(sptr[2 * i] - ea);
            e_mid = (uint32_t)(sptr[2 * i + 1] - ea);
            e_end = (uint32_t)(sptr[2 * i + 2] - ea);
        }
        bool act[VEC];
        float4 acc[VEC];
#pragma unroll
        for (int j = 0; j < VEC; ++j) {
            act[j] = live && (uint32_t)(l + j * LG) < a.nvec;
            acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        if (a.self_mode != SELF_ZERO) {  // self term first, like the reference (gcn_ops.cpp:166-171)
            const float sw = a.self_mode == SELF_NORM ? a.selfw[row] : 1.f;
            const float4 *sbase = a.self_mode == SELF_ACCUM ? reinterpret_cast<const float4 *>(a.out) : src4;
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (act[j]) {
                    const float4 x = sbase[(size_t)row * ld4 + l + j * LG];
                    acc[j] = make_float4(x.x * sw, x.y * sw, x.z * sw, x.w * sw);
                }
        }
        // two parts: [e, e_mid) from the window, [e_mid, e_end) from L2
#pragma unroll
        for (int part = 0; part < 2; ++part) {
            uint32_t pe = part == 0 ? e : e_mid;
            const uint32_t pend = part == 0 ? e_mid : e_end;
            while (__any_sync(kFull, pe < pend)) {
                uint32_t s_c = 0;
                float w_c = 0.f;
                if (pe + l < pend) {
                    s_c = sidx[pe + l];
                    w_c = sval[pe + l];
                }
                constexpr int U = VEC >= 3 ? 2 : 4;  // gathers in flight per group before their FMAs
#pragma unroll
                for (int k0 = 0; k0 < LG; k0 += U) {
                    float4 x[U][VEC];
                    float w[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int from = g * LG + k0 + u;
                        const uint32_t s = __shfl_sync(kFull, s_c, from);
                        w[u] = __shfl_sync(kFull, w_c, from);
                        const bool ev = pe + (k0 + u) < pend;  // a padding edge never touches memory
#pragma unroll
                        for (int j = 0; j < VEC; ++j) {
                            x[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (act[j] && ev)
                                x[u][j] = part == 0 ? lds_f4(win_base + ((s - wlo) * slab4 + l + j * LG) * 16u)
                                                    : ld_row_f4(src4 + (size_t)s * ld4 + l + j * LG, pol_keep);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u)
#pragma unroll
                        for (int j = 0; j < VEC; ++j) fma4(acc[j], x[u][j], w[u]);
                }
                pe += LG;
            }
        }
#pragma unroll
        for (int j = 0; j < VEC; ++j)
            if (act[j]) reinterpret_cast<float4 *>(a.out)[(size_t)row * ld4 + l + j * LG] = acc[j];
    }
}

// ---------------------------------------------------------------------------- low-degree tiles, pipelined
// The same walk as spmm_tile_group_kernel, as a PERSISTENT CTA with a two-stage TMA pipeline: one producer
// warp stages tile k+1 (window, offsets, ids, weights; completion on full[stage]) while the CW consumer
// warps walk tile k out of the other stage, and hands a stage back to the producer through empty[stage].
// A CTA that stages, waits and then computes keeps its shared memory idle for the whole load latency, and
// the window leaves room for only two or three such CTAs per SM; here the load of the next tile is always
// in flight behind the current one.  Tiles are dealt round-robin (tile = blockIdx.x + k * gridDim.x) from
// the plan's heaviest-first order.
struct StageHeader {
    uint32_t row0, nrows, wlo, wrows;
    uint64_t ea;  // first edge of the staged (16 B-aligned) run
    uint32_t pad[2];
};

template <int LG, int VEC, int CW, int OCC>
__global__ void __launch_bounds__(32 * (CW + 1), OCC)
spmm_tile_pipe_kernel(const SpmmArgs a, const TilePlanDev t, const __grid_constant__ CUtensorMap tmap) {
    extern __shared__ __align__(128) uint8_t smem[];
    __shared__ __align__(8) uint64_t full_mem[2], empty_mem[2];
    constexpr int G = 32 / LG;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int g = lane / LG, l = lane % LG;
    const uint32_t ld4 = a.ld >> 2;
    const uint32_t slab4 = min((uint32_t)(LG * VEC), ld4);
    const float4 *__restrict__ src4 = reinterpret_cast<const float4 *>(a.src);
    const uint32_t full0 = smem_u32(&full_mem[0]), empty0 = smem_u32(&empty_mem[0]);
    if (threadIdx.x == 0) {
        for (int st = 0; st < 2; ++st) {
            mbar_init(full0 + 8 * st, 1);
            mbar_init(empty0 + 8 * st, CW);
        }
        fence_barrier_init();
    }
    __syncthreads();
    const uint32_t n_my = t.n_tiles > blockIdx.x ? (t.n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;

    if (warp == CW) {
        // ------------------------------------------------------------------ producer warp
        for (uint32_t k = 0; k < n_my; ++k) {
            const uint32_t st = k & 1;
            uint8_t *base = smem + (size_t)st * t.stage_bytes;
            if (k >= 2) mbar_wait(empty0 + 8 * st, ((k >> 1) - 1) & 1);
            const uint32_t tile = blockIdx.x + k * gridDim.x;
            const uint32_t wlo = t.tile_wlo[tile], wrows = t.tile_wrows[tile];
            const uint32_t r_begin = t.tile_ptr[tile], nrows = t.tile_ptr[tile + 1] - r_begin;
            const uint64_t e0 = t.tile_e0[tile], e1 = t.tile_e1[tile];
            const uint32_t row0 = t.rows[r_begin];
            const uint64_t ea = e0 & ~(uint64_t)3;
            const uint32_t ebytes = (uint32_t)(((e1 + 3) & ~(uint64_t)3) - ea) * 4u;
            const uint32_t pbytes = (2u * nrows + 2u) * 8u;
            const uint32_t bar = full0 + 8 * st;
            if (lane == 0) {
                StageHeader *h = reinterpret_cast<StageHeader *>(base);
                h->row0 = row0;
                h->nrows = nrows;
                h->wlo = wlo;
                h->wrows = wrows;
                h->ea = ea;
                mbar_expect_tx(bar, (wrows ? window_bytes(ld4, slab4, wrows) : 0u) + pbytes + 2u * ebytes);
            }
            __syncwarp();
            float4 *win = reinterpret_cast<float4 *>(base + t.smem_win_off);
            if (wrows) {
                if (slab4 == ld4) {
                    const uint32_t total4 = wrows * ld4;
                    const float4 *wb = src4 + (size_t)wlo * ld4;
                    for (uint32_t c = lane * 256u; c < total4; c += 32u * 256u)
                        bulk_g2s(smem_u32(win + c), wb + c, min(256u, total4 - c) * 16u, bar);
                } else {
                    const uint32_t nbox = (wrows + kBoxRows - 1) / kBoxRows;
                    for (uint32_t b = lane; b < nbox; b += 32)
                        tma_load_2d(smem_u32(win + (size_t)b * kBoxRows * slab4), &tmap, 0, (int)(wlo + b * kBoxRows), bar);
                }
            }
            if (lane == 31) bulk_g2s(smem_u32(base + t.smem_ptr_off), t.ptrs + 2 * (size_t)row0, pbytes, bar);
            for (uint32_t c = lane * 8192u; c < ebytes; c += 32u * 8192u) {
                const uint32_t n = min(8192u, ebytes - c);
                bulk_g2s(smem_u32(base + t.smem_idx_off) + c, reinterpret_cast<const uint8_t *>(t.idx + ea) + c, n, bar);
                bulk_g2s(smem_u32(base + t.smem_val_off) + c, reinterpret_cast<const uint8_t *>(t.vals + ea) + c, n, bar);
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const uint64_t pol_keep = policy_evict_last();
    for (uint32_t k = 0; k < n_my; ++k) {
        const uint32_t st = k & 1;
        uint8_t *base = smem + (size_t)st * t.stage_bytes;
        mbar_wait(full0 + 8 * st, (k >> 1) & 1);
        const StageHeader h = *reinterpret_cast<const StageHeader *>(base);
        const uint64_t *sptr = reinterpret_cast<const uint64_t *>(base + t.smem_ptr_off);
        const uint32_t *sidx = reinterpret_cast<const uint32_t *>(base + t.smem_idx_off);
        const float *sval = reinterpret_cast<const float *>(base + t.smem_val_off);
        const uint32_t win_base = smem_u32(base + t.smem_win_off);
        const uint32_t wlo = h.wlo;
        for (uint32_t i0 = (uint32_t)warp * G; i0 < h.nrows; i0 += CW * G) {
            const uint32_t i = i0 + g;
            const bool live = i < h.nrows;
            const uint32_t row = h.row0 + (live ? i : 0u);
            uint32_t e = 0, e_mid = 0, e_end = 0;  // positions inside the staged run
            if (live) {
                e = (uint32_t)(sptr[2 * i] - h.ea);
                e_mid = (uint32_t)(sptr[2 * i + 1] - h.ea);
                e_end = (uint32_t)(sptr[2 * i + 2] - h.ea);
            }
            bool act[VEC];
            float4 acc[VEC];
#pragma unroll
            for (int j = 0; j < VEC; ++j) {
                act[j] = live && (uint32_t)(l + j * LG) < a.nvec;
                acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (a.self_mode != SELF_ZERO) {  // self term first, like the reference (gcn_ops.cpp:166-171)
                const float sw = a.self_mode == SELF_NORM ? a.selfw[row] : 1.f;
                const float4 *sbase = a.self_mode == SELF_ACCUM ? reinterpret_cast<const float4 *>(a.out) : src4;
#pragma unroll
                for (int j = 0; j < VEC; ++j)
                    if (act[j]) {
                        const float4 x = sbase[(size_t)row * ld4 + l + j * LG];
                        acc[j] = make_float4(x.x * sw, x.y * sw, x.z * sw, x.w * sw);
                    }
            }
#pragma unroll
            for (int part = 0; part < 2; ++part) {
                uint32_t pe = part == 0 ? e : e_mid;
                const uint32_t pend = part == 0 ? e_mid : e_end;
                while (__any_sync(kFull, pe < pend)) {
                    uint32_t s_c = 0;
                    float w_c = 0.f;
                    if (pe + l < pend) {
                        s_c = sidx[pe + l];
                        w_c = sval[pe + l];
                    }
                    constexpr int U = VEC >= 3 ? 2 : 4;
#pragma unroll
                    for (int k0 = 0; k0 < LG; k0 += U) {
                        float4 x[U][VEC];
                        float w[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int from = g * LG + k0 + u;
                            const uint32_t s = __shfl_sync(kFull, s_c, from);
                            w[u] = __shfl_sync(kFull, w_c, from);
                            const bool ev = pe + (k0 + u) < pend;  // a padding edge never touches memory
#pragma unroll
                            for (int j = 0; j < VEC; ++j) {
                                x[u][j] = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (act[j] && ev)
                                    x[u][j] = part == 0 ? lds_f4(win_base + ((s - wlo) * slab4 + l + j * LG) * 16u)
                                                        : ld_row_f4(src4 + (size_t)s * ld4 + l + j * LG, pol_keep);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u)
#pragma unroll
                            for (int j = 0; j < VEC; ++j) fma4(acc[j], x[u][j], w[u]);
                    }
                    pe += LG;
                }
            }
#pragma unroll
            for (int j = 0; j < VEC; ++j)
                if (act[j]) reinterpret_cast<float4 *>(a.out)[(size_t)row * ld4 + l + j * LG] = acc[j];
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * st);  // this warp is done with the stage
    }
}

// ---------------------------------------------------------------------------- host side
using EncodeFn = CUresult (*)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                              const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                              CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

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

// Tensor map of the source block [rows x ld floats] with a box of {boxCols floats, kBoxRows rows}, no swizzle:
// the box lands in shared memory row-major with a pitch of boxCols floats, which is the window's layout.
// Cached per (block, pitch, rows, box width).
bool window_map(const float *src, uint32_t ld, uint64_t rows, uint32_t boxCols, CUtensorMap &out) {
    static std::mutex mu;
    static std::map<std::tuple<const float *, uint32_t, uint64_t, uint32_t>, CUtensorMap> cache;
    std::lock_guard<std::mutex> lock(mu);
    auto key = std::make_tuple(src, ld, rows, boxCols);
    auto it = cache.find(key);
    if (it == cache.end()) {
        EncodeFn fn = encode_fn();
        if (!fn) return false;
        CUtensorMap m;
        const cuuint64_t dims[2] = {ld, rows};
        const cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
        const cuuint32_t box[2] = {boxCols, kBoxRows};
        const cuuint32_t estr[2] = {1, 1};
        if (fn(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(src), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
        if (cache.size() > 256) cache.clear();  // tensors of engines long gone
        it = cache.emplace(key, m).first;
    }
    out = it->second;
    return true;
}

// Opt-in dynamic shared memory; remembered per kernel instantiation (and device) so that the attribute
// call is off the launch path after the first use.
template <class K>
bool set_smem(K kernel, size_t bytes, size_t (&granted)[16]) {
    int dev = 0;
    cudaGetDevice(&dev);
    dev &= 15;
    if (bytes <= granted[dev]) return true;
    if (bytes > 227u * 1024u) {
        fprintf(stderr, "[dorylus_b200] tile kernel: window of %zu bytes exceeds shared memory\n", bytes);
        return false;
    }
    const cudaError_t ce = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
    if (ce != cudaSuccess) {
        fprintf(stderr, "[dorylus_b200] tile kernel: cudaFuncSetAttribute(%zu bytes): %s\n", bytes, cudaGetErrorString(ce));
        return false;
    }
    granted[dev] = bytes;
    return true;
}

template <int LG, int VEC, int OCC>
int launch_tile(const SpmmArgs &a, const TilePlanDev &t, cudaStream_t s) {
    const uint32_t slab = LG * VEC, ld4 = a.ld / 4;
    const uint32_t nslab = (a.nvec + slab - 1) / slab;
    const uint32_t slab4 = std::min<uint32_t>(slab, ld4);
    const bool contiguous = slab4 == ld4;
    const size_t smem = (size_t)window_rows_padded(t.max_wrows, contiguous) * slab4 * 16;
    CUtensorMap map{};
    if (!contiguous && !window_map(a.src, a.ld, a.src_rows, slab4 * 4, map)) return -1;
    static thread_local size_t granted[16] = {};  // per kernel instantiation
    if (!set_smem(spmm_tile_kernel<LG, VEC, OCC>, smem, granted)) return -1;
    spmm_tile_kernel<LG, VEC, OCC><<<dim3(t.n_tiles, nslab), 32 * kTileWarps, smem, s>>>(a, t, map);
    const cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) fprintf(stderr, "[dorylus_b200] tile kernel <%d,%d> launch (%u tiles x %u slabs, %zu B smem): %s\n", LG, VEC, t.n_tiles, nslab, smem, cudaGetErrorString(ce));
    return ce == cudaSuccess ? 1 : -1;
}

template <int LG, int VEC, int WARPS, int OCC>
int launch_tile_group(const SpmmArgs &a, TilePlanDev t, cudaStream_t s) {
    const uint32_t ld4 = a.ld / 4;
    const uint32_t slab4 = std::min<uint32_t>(LG * VEC, ld4);
    const bool contiguous = slab4 == ld4;
    auto up = [](size_t x) { return (x + 127) & ~(size_t)127; };
    const size_t winBytes = up((size_t)window_rows_padded(t.max_wrows, contiguous) * slab4 * 16);
    const size_t ptrBytes = up(((size_t)2 * t.max_tile_rows + 2) * 8);
    const size_t edgeBytes = up(((size_t)t.max_tile_edges + 8) * 4);
    t.smem_ptr_off = (uint32_t)winBytes;
    t.smem_idx_off = (uint32_t)(winBytes + ptrBytes);
    t.smem_val_off = (uint32_t)(winBytes + ptrBytes + edgeBytes);
    const size_t smem = winBytes + ptrBytes + 2 * edgeBytes;
    CUtensorMap map{};
    if (!contiguous && !window_map(a.src, a.ld, a.src_rows, slab4 * 4, map)) return -1;
    static thread_local size_t granted[16] = {};  // per kernel instantiation
    if (!set_smem(spmm_tile_group_kernel<LG, VEC, WARPS, OCC>, smem, granted)) return -1;
    spmm_tile_group_kernel<LG, VEC, WARPS, OCC><<<t.n_tiles, 32 * WARPS, smem, s>>>(a, t, map);
    const cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) fprintf(stderr, "[dorylus_b200] tile group kernel <%d,%d> launch (%u tiles, %zu B smem): %s\n", LG, VEC, t.n_tiles, smem, cudaGetErrorString(ce));
    return ce == cudaSuccess ? 1 : -1;
}

template <int LG, int VEC, int CW, int OCC>
int launch_tile_pipe(const SpmmArgs &a, TilePlanDev t, cudaStream_t s) {
    const uint32_t ld4 = a.ld / 4;
    const uint32_t slab4 = std::min<uint32_t>(LG * VEC, ld4);
    const bool contiguous = slab4 == ld4;
    auto up = [](size_t x) { return (x + 127) & ~(size_t)127; };
    const size_t winBytes = up((size_t)window_rows_padded(t.max_wrows, contiguous) * slab4 * 16);
    const size_t ptrBytes = up(((size_t)2 * t.max_tile_rows + 2) * 8);
    const size_t edgeBytes = up(((size_t)t.max_tile_edges + 8) * 4);
    t.smem_win_off = 128;  // after the stage header
    t.smem_ptr_off = (uint32_t)(128 + winBytes);
    t.smem_idx_off = (uint32_t)(128 + winBytes + ptrBytes);
    t.smem_val_off = (uint32_t)(128 + winBytes + ptrBytes + edgeBytes);
    t.stage_bytes = (uint32_t)(128 + winBytes + ptrBytes + 2 * edgeBytes);
    const size_t smem = 2 * (size_t)t.stage_bytes;
    CUtensorMap map{};
    if (!contiguous && !window_map(a.src, a.ld, a.src_rows, slab4 * 4, map)) return -1;
    static thread_local size_t granted[16] = {};  // per kernel instantiation
    if (!set_smem(spmm_tile_pipe_kernel<LG, VEC, CW, OCC>, smem, granted)) return -1;
    static int sms = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        if (sms <= 0) sms = 148;
    }
    const uint32_t perSm = (uint32_t)std::max<size_t>(1, std::min<size_t>(OCC, (227u * 1024u) / (smem + 1024)));
    const uint32_t grid = std::min<uint32_t>(t.n_tiles, (uint32_t)sms * perSm);
    spmm_tile_pipe_kernel<LG, VEC, CW, OCC><<<grid, 32 * (CW + 1), smem, s>>>(a, t, map);
    const cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) fprintf(stderr, "[dorylus_b200] tile pipe kernel <%d,%d> launch (%u CTAs, %zu B smem): %s\n", LG, VEC, grid, smem, cudaGetErrorString(ce));
    return ce == cudaSuccess ? 1 : -1;
}

}  // namespace

size_t tile_smem_bytes(uint32_t ld, uint32_t nvec, uint32_t windowRows, bool lowDegree, int slabFloats) {
    const uint32_t ld4 = ld / 4;
    uint32_t slab4;
    if (lowDegree) slab4 = std::min(ld4, nvec <= 4 ? 4u : nvec <= 8 ? 8u : nvec <= 12 ? 12u : nvec <= 16 ? 16u : 32u);
    else slab4 = std::min<uint32_t>(ld4, (uint32_t)slabFloats / 4);
    return (size_t)window_rows_padded(windowRows, slab4 == ld4) * slab4 * 16;
}

size_t tile_edge_smem_bytes(uint64_t maxTileEdges, uint32_t maxTileRows) {
    auto up = [](size_t x) { return (x + 127) & ~(size_t)127; };
    return up(((size_t)2 * maxTileRows + 2) * 8) + 2 * up(((size_t)maxTileEdges + 8) * 4);
}

// Returns the number of kernels launched, 0 when this shape has no tile kernel, -1 on a launch error.
int launch_spmm_tile(const SpmmArgs &a, const TilePlanDev &t, cudaStream_t s) {
    if (t.n_tiles == 0) return 0;
    if (t.low_degree && t.pipeline) {
        const uint32_t n = a.nvec;
        if (n <= 4) return launch_tile_pipe<4, 1, 8, 3>(a, t, s);
        if (n <= 8) return launch_tile_pipe<4, 2, 8, 3>(a, t, s);
        if (n <= 12) return launch_tile_pipe<4, 3, 8, 2>(a, t, s);
        if (n <= 16) return launch_tile_pipe<4, 4, 8, 2>(a, t, s);
        if (n <= 32) return launch_tile_pipe<8, 4, 8, 2>(a, t, s);
        return 0;
    }
    if (t.low_degree) {
        const uint32_t n = a.nvec;
        if (n <= 4) return launch_tile_group<4, 1, 8, 4>(a, t, s);
        if (n <= 8) return launch_tile_group<4, 2, 8, 4>(a, t, s);
        if (n <= 12) return launch_tile_group<4, 3, 8, 3>(a, t, s);
        if (n <= 16) return launch_tile_group<4, 4, 8, 3>(a, t, s);
        if (n <= 32) return launch_tile_group<8, 4, 8, 3>(a, t, s);
        return 0;
    }
    switch (t.slab_floats) {
    case 32: return launch_tile<8, 1, 2>(a, t, s);
    case 64: return launch_tile<8, 2, 2>(a, t, s);
    case 96: return launch_tile<8, 3, 2>(a, t, s);
    case 128: return launch_tile<8, 4, 1>(a, t, s);
    default: return 0;
    }
}

}  // namespace dory
