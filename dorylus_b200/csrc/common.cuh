// Shared device/host declarations of the engine (internal; not part of the C ABI).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

namespace dory {

// Row-major fp32 matrix resident in HBM.  `ld` (floats) is the padded row pitch: a multiple of 32
// floats (one 128 B line) for widths > 16, a multiple of 4 otherwise, so that every row starts
// 16 B-aligned for 128-bit loads and wide rows start on a cache-line boundary.  Padding columns are
// kept at zero by every producer.
struct DevMat {
    float *p = nullptr;
    uint64_t rows = 0;
    uint32_t cols = 0;
    uint32_t ld = 0;
    __host__ __device__ float *row(uint64_t r) const { return p + r * ld; }
    DevMat rows_from(uint64_t r0, uint64_t n) const {
        DevMat m = *this;
        m.p = p + r0 * ld;
        m.rows = n;
        return m;
    }
};

inline uint32_t padded_ld(uint32_t cols) {
    if (cols <= 16) return (cols + 3u) & ~3u;
    return (cols + 31u) & ~31u;
}

// ---- aggregation (spmm.cu) ---------------------------------------------------------------
enum SelfMode : int {
    SELF_NORM = 0,   // out = selfw[v] * src[v] + sum      (GCN, vtxDataVec)
    SELF_ONE = 1,    // out = src[v] + sum                 (GAT forward, unit self weight)
    SELF_ZERO = 2,   // out = sum                          (GAT backward, first term)
    SELF_ACCUM = 3   // out = out + sum                    (GAT backward, second term)
};

struct SpmmArgs {
    const uint64_t *ptrs;   // [V+1] columnPtrs (CSC, forward) or rowPtrs (CSR, backward); with source
                            // blocking [V*ptr_stride + 1]: row v, block b spans [ptrs[v*stride+b], ptrs[v*stride+b+1])
    uint32_t ptr_stride;    // number of source blocks the edge list of every row is grouped into (1 = none)
    uint32_t ptr_off;       // first source block this launch walks
    uint32_t ptr_span;      // number of consecutive source blocks it walks (>= 1)
    const uint32_t *idx;    // [E]   source row in `src` (local id, or V + ghost slot)
    const float *vals;      // [E]
    const float *selfw;     // [V]   vtxDataVec (SELF_NORM only)
    const float *src;       // [(V+G) x ld] local rows then ghost rows
    uint32_t src_rows;      // rows of the block `src` points at (local + ghost): bounds of its TMA tensor map
    float *out;             // [V x ld]
    uint32_t ld;            // common row pitch of src and out, in floats
    uint32_t nvec;          // row width in float4 units to process (data columns only, <= ld / 4)
    int self_mode;
    const uint32_t *heavy;  // row ids with degree >= heavy threshold, degree-descending (may be null)
    uint32_t n_heavy;
    uint32_t n_vheavy;      // the first n_vheavy entries of `heavy` are hub rows: a cluster of CTAs each
    const uint32_t *light;  // remaining row ids, degree-descending; null => rows low..low+n_light-1
    uint32_t n_light;
    uint32_t low;           // first row when `light` is null
    int cfg_lg, cfg_vec;    // kernel shape override (0 = derive from nvec)
    int cfg_unroll;         // gather instructions in flight per lane group (0 = default)
    int cfg_occ;            // CTAs per SM the kernel is compiled for (0 = default)
    int cfg_light;          // light rows: 0 auto, 1 a warp per row, 2 a lane group per row
    uint32_t light_avg_degree;  // mean edges per light row (picks the light-row kernel)
};

// Launches the aggregation; returns the number of kernels launched, or -1 on a launch error.
int launch_spmm(const SpmmArgs &a, cudaStream_t s);
// True for the (lanes per row, float4 per lane) shapes the row kernels are instantiated for (0, 0 = derive).
bool spmm_shape_supported(int lg, int vec);
// Heavy / light rows through the warp- and CTA-per-row kernels only (no lane-group selection).
int launch_spmm_rows(const SpmmArgs &a, cudaStream_t s);

// ---- shared-memory-staged aggregation (spmm_tile.cu, plan: tile_plan.h) ----------------------------
struct TilePlanDev {
    const uint64_t *ptrs;        // [2V + 1] row v: in-window edges [ptrs[2v], ptrs[2v+1]), the rest [ptrs[2v+1], ptrs[2v+2])
    const uint32_t *idx;         // [E] regrouped source rows
    const float *vals;           // [E]
    const uint32_t *rows;        // rows grouped by tile (team rows first, degree-descending)
    const uint32_t *tile_ptr;    // [n_tiles + 1]
    const uint32_t *tile_team;   // [n_tiles] leading rows walked by the whole CTA
    const uint32_t *tile_wlo;    // [n_tiles] first source row of the staged window
    const uint32_t *tile_wrows;  // [n_tiles] rows of the window (0: nothing staged)
    const uint64_t *tile_e0;     // [n_tiles] edge run [e0, e1) of the tile's rows in idx / vals (low-degree mode:
    const uint64_t *tile_e1;     //           the rows are consecutive, so the run is contiguous)
    uint32_t n_tiles;
    uint32_t max_wrows;
    uint32_t max_tile_rows;
    uint64_t max_tile_edges;
    uint32_t smem_ptr_off, smem_idx_off, smem_val_off;  // filled by the launcher (low-degree mode)
    uint32_t smem_win_off, stage_bytes;                 // pipelined kernel: layout of one stage
    int pipeline;                // low-degree mode: 1 = persistent CTAs with a two-stage TMA pipeline
    int low_degree;              // 1: lane group per row (rows fit one slab), 0: warp / CTA per row
    int slab_floats;             // high-degree mode: column slab width (32, 64, 96 or 128 floats)
};
// Returns kernels launched, 0 when the shape has no tile kernel (caller falls back), -1 on a launch error.
int launch_spmm_tile(const SpmmArgs &a, const TilePlanDev &t, cudaStream_t s);
// Dynamic shared memory one CTA of the tile kernel needs for rows of pitch ld / nvec float4 of data.
size_t tile_smem_bytes(uint32_t ld, uint32_t nvec, uint32_t windowRows, bool lowDegree, int slabFloats);
// ... plus, in low-degree mode, what the tile's staged offsets / ids / weights take.
size_t tile_edge_smem_bytes(uint64_t maxTileEdges, uint32_t maxTileRows);

// ---- dense apply (dense.cu) ----------------------------------------------------------------
enum GemmEpilogue : int { EPI_NONE = 0, EPI_TANH = 1 };

// C[M x N] = op(A) . op(B), fp32 SIMT.  transA: A stored [K x M]; transB: B stored [N x K].
// EPI_TANH additionally writes C2 = tanh(C).  For transA (the dW = AH^T . G reduction over all
// vertices) the K range is split over `ws` partial buffers and reduced in a fixed order.
struct GemmArgs {
    const float *A;
    uint32_t lda;
    const float *B;
    uint32_t ldb;
    float *C;
    uint32_t ldc;
    float *C2;  // EPI_TANH second output (same ldc), else null
    uint64_t M;
    uint32_t N;
    uint64_t K;
    bool transA, transB;
    int epilogue;
    float *ws;         // split-K workspace (transA only)
    size_t ws_floats;  // capacity of ws
    const float *Bh;   // narrow transA kernel only (M <= 64, ws set): B is read as B (*) (1 - Bh^2), Bh with B's pitch
};
int launch_gemm(const GemmArgs &g, cudaStream_t s);

// g = aTg (*) (1 - h^2)                              (CPU_comm.cpp:142-143)
int launch_tanh_backward(const float *aTg, const float *h, float *g, uint64_t n, cudaStream_t s);
// h = tanh(z) over n floats (n a multiple of 4)     (activate, CPU_comm.cpp:265-274)
int launch_tanh_forward(const float *z, float *h, uint64_t n, cudaStream_t s);

// Last-layer fused softmax / validation statistics / maskout / gradient scale
// (CPU_comm.cpp:108-121).  Writes d = (maskout(softmax(z)) - lab) / denom into `d`,
// optionally the un-masked predictions into `pred`, and acc/loss sums into stats[0..1].
struct SoftmaxCEArgs {
    const float *z;    // [V x ld] logits
    const float *lab;  // [V x ld] one-hot labels
    float *d;          // [V x ld]
    float *pred;       // [V x ld] or null
    uint32_t ld, C;
    uint32_t V;
    uint32_t trainEnd;    // (unsigned)(V * TRAIN_PORTION)
    uint32_t valEnd;      // trainEnd + (unsigned)(V * VAL_PORTION)
    uint64_t maskFloats;  // quirk Q6: number of FLOATS after row trainEnd overwritten by labels
    bool strictMask;      // DORY_FLAG_STRICT_MASK: overwrite whole rows >= trainEnd
    float denom;          // (float)(globalVtxCnt * TRAIN_PORTION)
    float *rowstat;       // [2 x V] scratch: per-row acc / loss contributions
    float *stats;         // [2] device: acc sum, loss sum
};
int launch_softmax_ce(const SoftmaxCEArgs &a, cudaStream_t s);
// The same, fused behind the logits product A[V x K] . W[K x ld] (a.z is not read); 0 = shape does not qualify.
int launch_gemm_softmax_ce(const float *A, uint32_t lda, const float *W, uint32_t ldw, uint64_t K, const SoftmaxCEArgs &a,
                           cudaStream_t s);
int launch_softmax_stats(const SoftmaxCEArgs &a, cudaStream_t s);  // the reduction alone (after launch_gemm_tc_softmax)

// Adam step on one weight matrix (AdamOptimizer.cpp:36-48); lr_t is computed on the host.
int launch_adam(float *w, const float *grad, float *m, float *v, size_t n, float lr_t, float beta1,
                float beta2, float eps, cudaStream_t s);

int launch_fill(float *p, size_t n, float value, cudaStream_t s);
// fp32 FMA micro-benchmark (bench.py's compute roof): blocks x 256 threads x iters x 8 FMAs
int launch_fma_peak(float *sink, int iters, unsigned blocks, cudaStream_t s);
// dst[r*ldd + c] = src[r*lds + c] for c < cols (changes the row pitch; either side may be dense).
int launch_repitch(const float *src, uint32_t lds, float *dst, uint32_t ldd, uint64_t rows, uint32_t cols,
                   cudaStream_t s);

// ---- ghost exchange (comm.cu) ---------------------------------------------------------------
// Packs rows `ids[i]` of src into dst[i] (row pitch ld floats, nvec float4 per row).
int launch_gather_rows(const float *src, const uint32_t *ids, uint32_t n, float *dst, uint32_t ld,
                       cudaStream_t s);

}  // namespace dory
