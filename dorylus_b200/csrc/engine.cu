// The engine behind include/dorylus_b200.h: partition state resident in HBM, the named-tensor
// table, the SAGA operators and the per-epoch state machine.
//
// Reference counterparts: Engine (src/graph-server/engine/engine.{hpp,cpp}, engine/utils.cpp,
// engine/ops/gcn_ops.cpp, gat_ops.cpp), ResourceComm / CPUComm (commmanager/resource_comm.cpp,
// CPU_comm.cpp) and the weight-server pieces a synchronous run needs (weight-server/
// weightserver.cpp:515-612, AdamOptimizer.cpp).  Design notes: DESIGN.md.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <memory>
#include <numeric>
#include <random>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dorylus_b200.h"
#include "comm.h"
#include "common.cuh"
#include "gat.cuh"
#include "gemm_tc.cuh"
#include "loader.h"
#include "partition.h"
#include "tile_plan.h"

using namespace dory;

namespace {

constexpr double kTrainPortion = 0.66;  // src/common/utils.hpp:60
constexpr double kValPortion = 0.1;     // src/common/utils.hpp:61
constexpr uint32_t kHeavyDegree = 1024; // default: rows with more edges get a whole CTA (spmm.cu)
constexpr int kNumEvents = 64;
// L2 budget of one aggregation pass (source rows x slab bytes).  Sweep on Reddit: 2 windows of 60 MB
// beat 1, 3 and 4 (profiles/round1_spmm_sweep_v4_srcblocks.json).
constexpr uint64_t kWindowBytes = 64ull << 20;

thread_local std::string g_create_error;

struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
    DevBuf() = default;
    DevBuf(const DevBuf &) = delete;
    DevBuf &operator=(const DevBuf &) = delete;
    DevBuf(DevBuf &&o) noexcept : p(o.p), bytes(o.bytes) {
        o.p = nullptr;
        o.bytes = 0;
    }
    ~DevBuf() { release(); }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    cudaError_t alloc(size_t n) {
        release();
        if (n == 0) n = 16;
        cudaError_t e = cudaMalloc(&p, n);
        if (e == cudaSuccess) bytes = n;
        return e;
    }
    template <class T>
    T *as() const { return static_cast<T *>(p); }
};

// Device copy of a tile plan (tile_plan.h) for the shared-memory-staged aggregation (spmm_tile.cu).
struct TilePlanBuf {
    DevBuf ptrs, idx, vals, rows, tptr, tteam, twlo, twrows, te0, te1;
    uint32_t n_tiles = 0, max_wrows = 0, tile_rows = 0, window_rows = 0, max_tile_rows = 0;
    uint64_t max_tile_edges = 0;
    double coverage = 0.0;
    bool low_degree = false;
    int slab_floats = 0;
};

struct Adjacency {
    std::unique_ptr<TilePlanBuf> tile;  // null: no plan (option off, or the graph has no locality to stage)
    DevBuf ptrs, idx, vals, heavy, light;
    uint64_t nnz = 0;
    uint32_t n_heavy = 0, n_light = 0;
    uint32_t n_src_rows = 0;  // rows of the block the indices address (local + ghost)
    uint32_t n_vheavy = 0;  // leading entries of `heavy` with degree >= hub_degree (a CTA cluster each)
    uint32_t light_avg_degree = 0;
    // Source-blocked copy (GCN): every row's edge list regrouped by source-row block so that one
    // launch only gathers from a (V+G)/nb-row window of the feature block (an L2-sized working set).
    DevBuf bptrs, bidx, bvals;  // [V*nb + 1], [E], [E]
    uint32_t nb = 1;
    // > 0: the first nb_local windows hold the partition's OWN rows, the others its ghost rows -- an
    // aggregation can then walk the local part while the ghost exchange is still in flight
    uint32_t nb_local = 0;
};

struct WeightSet {
    DevBuf w, dw, m, v;
    uint32_t rows = 0, cols = 0, ld = 0, prows = 0;  // prows: rows incl. zero padding (= ld of the input width)
    size_t floats() const { return (size_t)prows * ld; }
};

}  // namespace

struct dory_engine {
    dory_config cfg{};
    std::string err;
    cudaStream_t stream = nullptr;
    cudaEvent_t events[kNumEvents] = {};
    bool loaded = false;

    // graph (Graph, graph/graph.hpp:60-99)
    uint32_t V = 0, gV = 0, Gs = 0, Gd = 0;
    uint64_t Ein = 0, Eout = 0, Eglob = 0;
    Adjacency fwd, bwd;  // forwardAdj (CSC), backwardAdj (CSR)
    DevBuf norms;        // vtxDataVec
    std::vector<uint32_t> l2g;
    std::vector<std::vector<uint32_t>> sendIds[2];  // [dir][peer] local ids (host copy)

    // tensors: backing allocations + (layer, name) -> view
    std::vector<std::unique_ptr<DevBuf>> pool;
    std::map<std::pair<uint32_t, std::string>, DevMat> tensors;
    std::vector<WeightSet> W;    // "w" per layer
    std::vector<WeightSet> Ai;   // GAT "a_i" per layer (F' x 1)
    DevBuf scratchA, scratchB;   // V x max(ld) scratch (d / interGrad / pred)
    DevBuf gemm_ws;              // split-K partials
    DevBuf rowstat, stats_dev;   // softmax-CE reduction scratch
    DevBuf flush;                // L2 flush target
    DevBuf stage;                // dense staging for host <-> padded-row copies
    // input pipeline (dory_prefetch_tensor / dory_commit_prefetch): DMA on its own stream into a
    // per-tensor staging buffer, re-pitched into the tensor on the compute stream at commit time
    struct Prefetch {
        DevMat dst;
        DevBuf stage;
        cudaEvent_t staged = nullptr, consumed = nullptr;
        bool pending = false, used = false;
    };
    cudaStream_t copy_stream = nullptr;
    std::map<std::pair<uint32_t, std::string>, std::unique_ptr<Prefetch>> prefetch;
    int spmm_lg = 0, spmm_vec = 0, spmm_unroll = 0, spmm_occ = 0, spmm_light = 0;
    int row_order = 0;  // light-row issue order: 0 = decide from the graph, 1 = degree-descending, 2 = degree classes
    int tensor_cores = 1;  // tcgen05 path for H.W (option "tensor_cores")
    uint32_t src_blocks = 0;  // source windows per aggregation (0 = size from L2, 1 = off)
    // option "gat_windows": source windows for the GAT aggregations too (default on since round 2: Reddit GAT
    // epoch 18.0 -> 16.8 ms, profiles/round2_shape_reddit_gat*.json; parity at 602/128/41 in test_gpu_reddit_widths)
    int gat_windows = 1;
    uint32_t heavy_degree = kHeavyDegree;
    uint32_t hub_degree = 0;  // rows with more edges get a cluster of 8 CTAs (0 = from the partition's size)
    uint32_t locality_block = 0;  // rows per block of the locality-preserving row order (0 = from L2)
    // shared-memory-staged aggregation (options "tile*", include/dorylus_b200.h)
    // 0 off; 1 on whenever a plan can be built; 2 (default) on for HIGH-degree graphs whose plan serves enough
    // edges from shared memory -- measured (profiles/round2_tile_kernel.md): on the community-structured Reddit
    // shape the staged kernel is 14-17 % faster than the gather kernels, on the low-degree shapes it is slower
    // (there it needs an explicit tile=1), on a graph without locality no plan is kept at all.
    int tile_mode = 2;
    uint32_t tile_rows = 0, tile_window = 0, tile_smem_kb = 100, tile_slab = 0, tile_min_coverage = 50, tile_team = 4096;
    uint32_t tile_edges = 4096;   // low-degree mode: edges per tile (their ids / weights are staged too)
    int tile_pipe = 1;            // low-degree mode: persistent CTAs with a two-stage TMA pipeline
    int tn_small = 1;             // option "tn_small": narrow-M fp32 kernel for dW of layers with input width <= 64
    int fuse_softmax = 1;         // option "fuse_softmax": last-layer logits + soft-max / maskout in one kernel (C <= 64);
                                  // 1 = on tcgen05 when the shape qualifies, else fp32 SIMT; 2 = the SIMT kernel only
    int fuse_tanh_bwd = 1;        // option "fuse_tanh_bwd": layer-0 backward with an input width <= 32: tanh' applied inside the
                                  // narrow dW kernel's operand load instead of a pass of its own
    int tc_small = 1;             // option "tc_small": the small-tile tcgen05 kernel (several CTAs per SM) for that product, for
                                  // Z = A.W with K <= 128, N <= 64 and for grad = G.W^T; 0 = the round-1 kernels
    int tc_stages = 0;            // option "tc_stages": its shared-memory stages per CTA (0 = choose)
    // apply-first schedule (DORY_FLAG_APPLY_FIRST, include/dorylus_b200.h): af[l] != 0 -> layer l runs
    // A_hat . (in . W); decided in dory_load_partition from the flag / the "apply_first_mask" option
    std::vector<uint8_t> af;
    long af_mask = -1;
    bool apply_first(uint32_t l) const { return l < af.size() && af[l]; }

    // Adam (AdamOptimizer.hpp:69-84)
    float beta1 = .9f, beta2 = .999f, eps = 1e-07f, lr_t = 0.f;
    unsigned adam_epochs = 0;

    dory_stats stats{};
    // stream-ordered statistics read-back (dory_stats_enqueue / dory_stats_collect): pinned host
    // slots, one event each
    static constexpr uint32_t kStatSlots = 4;
    float *stats_host = nullptr;  // [kStatSlots][2], cudaHostAlloc
    cudaEvent_t stats_ready[kStatSlots] = {};
    dory_stats stats_snap[kStatSlots] = {};
    bool stats_pending[kStatSlots] = {};
    std::unique_ptr<dory::Comm> comm;
    // peer-memory exchange: local ghost tensor -> per-peer pointer to THEIR ghost tensor of the same
    // (layer, name), mapped with cudaIpcOpenMemHandle; ipc_bases are the mappings to close
    std::map<const float *, std::vector<float *>> peer_ghost;
    // what every peer said about ITS block when it exported it (rows of the ghost block), and which
    // (ghost tensor, direction) plans have been checked against that since the plan last changed
    std::map<const float *, std::vector<uint64_t>> peer_ghost_rows;
    std::set<std::pair<const float *, uint32_t>> p2p_validated;
    std::vector<void *> ipc_bases;
    int p2p = 1;
    // exchange / compute overlap (option "overlap"): a peer-memory exchange runs on its own high-priority
    // stream with its own communicator; the aggregation that consumes the ghost block walks the edges from
    // local rows first and only then waits for it (interior first, boundary after the receive)
    // Default 0: measured on the Amazon shape at 2 GPUs (profiles/round2_overlap_n2.md) the two passes over the
    // rows cost more than the hidden exchange returns.
    int overlap = 0;
    cudaStream_t comm_stream = nullptr;
    cudaEvent_t ev_compute = nullptr, ev_comm = nullptr;
    bool comm_pending = false;                 // work on comm_stream the compute stream has not waited for yet
    const float *comm_pending_ghost = nullptr; // the ghost block the pending exchange fills
    int p2p_variant = 0;        // option "p2p_rows": rows per warp of the store kernel (comm.cu)
    int p2p_ctas = 1;           // option "p2p_ctas": CTAs per SM of the store kernel in overlap mode
    int elide_pre_barrier = 1;  // option "p2p_elide_barrier"
    // ghost blocks an aggregation has read since this engine last took part in a collective: only
    // those need the barrier in front of a peer-memory exchange that overwrites them
    std::set<const float *> ghost_reads_pending;

    // widest row slab (bytes, <= 512) any aggregation of this model gathers: GCN aggregates widths
    // F_0 .. F_{L-1} (an apply-first layer l: F_{l+1}), GAT the layer outputs F_1 .. F_L
    uint32_t max_slab_bytes() const {
        uint32_t m = 16;
        const uint32_t lo = cfg.gnn_type == DORY_GCN ? 0 : 1, hi = cfg.gnn_type == DORY_GCN ? cfg.n_layers : cfg.n_layers + 1;
        for (uint32_t l = lo; l < hi; ++l) {
            const uint32_t w = apply_first(l) ? cfg.dims[l + 1] : cfg.dims[l];
            m = std::max(m, std::min<uint32_t>(padded_ld(w) * 4, 512));
        }
        return m;
    }
    uint32_t L() const { return cfg.n_layers; }
    uint32_t dim(uint32_t i) const { return cfg.dims[i]; }
};

namespace {

int fail(dory_engine *e, int code, const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (e) e->err = buf; else g_create_error = buf;
    return code;
}

#define CU(call)                                                                              \
    do {                                                                                      \
        cudaError_t _c = (call);                                                              \
        if (_c != cudaSuccess)                                                                \
            return fail(e, DORY_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_c), \
                        __FILE__, __LINE__);                                                  \
    } while (0)

#define LAUNCHED(expr)                                                                  \
    do {                                                                                \
        int _n = (expr);                                                                \
        if (_n < 0)                                                                     \
            return fail(e, DORY_ECUDA, "kernel launch failed: %s (%s:%d)",              \
                        cudaGetErrorString(cudaGetLastError()), __FILE__, __LINE__);    \
        e->stats.kernel_launches += (uint64_t)_n;                                       \
    } while (0)

DevMat new_tensor(dory_engine *e, uint64_t rows, uint32_t cols, cudaError_t &err) {
    DevMat m;
    m.rows = rows;
    m.cols = cols;
    m.ld = padded_ld(cols);
    auto buf = std::make_unique<DevBuf>();
    const size_t bytes = (size_t)std::max<uint64_t>(rows, 1) * m.ld * sizeof(float);
    err = buf->alloc(bytes);
    if (err != cudaSuccess) return m;
    err = cudaMemsetAsync(buf->p, 0, bytes, e->stream);
    m.p = buf->as<float>();
    e->pool.push_back(std::move(buf));
    return m;
}

const DevMat *find_tensor(const dory_engine *e, uint32_t layer, const char *name) {
    auto it = e->tensors.find({layer, std::string(name)});
    return it == e->tensors.end() ? nullptr : &it->second;
}

// Degree-descending row lists (longest-processing-time-first issue order for spmm.cu).
void build_row_lists(const std::vector<uint64_t> &ptrs, uint32_t heavyDegree, bool keepLocality, uint32_t blockRows,
                     std::vector<uint32_t> &heavy, std::vector<uint32_t> &light) {
    const uint32_t V = (uint32_t)ptrs.size() - 1;
    std::vector<uint32_t> order(V);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return (ptrs[a + 1] - ptrs[a]) > (ptrs[b + 1] - ptrs[b]);
    });
    heavy.clear();
    light.clear();
    // heavy rows (a CTA each): heaviest first, so the hubs start early and the tail back-fills.
    // light rows (a warp, or a lane group, each):
    //  - graph without locality in its vertex numbering: plain degree-descending order -- rows that
    //    share a CTA / warp then have near-equal trip counts (a CTA's registers are held until its
    //    longest row ends: natural order cost 13 % on the Reddit shape, degree classes 5 %);
    //  - graph WITH locality (most edges stay near the diagonal, e.g. community-ordered ids): blocks
    //    of `blockRows` consecutive vertices, inside a block power-of-two degree classes (heaviest
    //    first), inside a class natural id order.  The classes keep the trip counts of a warp's rows
    //    within 2x; the blocks keep a community's rows together in TIME: with classes over the whole
    //    graph a community's source rows were fetched from HBM once per class (7.1 GB of DRAM reads
    //    for 1.5 GB of compulsory bytes on the Friendster shape, profiles/round1_lowdeg.md), with
    //    blocks the source rows of a block (<= 24 MB) are still in L2 when its next class runs.
    for (uint32_t v : order)
        if ((ptrs[v + 1] - ptrs[v]) >= heavyDegree) heavy.push_back(v);
    if (!keepLocality) {
        for (uint32_t v : order)
            if ((ptrs[v + 1] - ptrs[v]) < heavyDegree) light.push_back(v);
        return;
    }
    auto cls = [&](uint32_t v) {
        uint64_t d = ptrs[v + 1] - ptrs[v];
        int c = 0;
        while (d >>= 1) ++c;
        return c;
    };
    blockRows = std::max(blockRows, 1u);
    std::vector<std::vector<uint32_t>> byClass(64);
    for (uint32_t b0 = 0; b0 < V; b0 += blockRows) {
        const uint32_t b1 = (uint32_t)std::min<uint64_t>((uint64_t)b0 + blockRows, V);
        for (auto &c : byClass) c.clear();
        for (uint32_t v = b0; v < b1; ++v)
            if ((ptrs[v + 1] - ptrs[v]) < heavyDegree) byClass[cls(v)].push_back(v);
        for (int c = 63; c >= 0; --c) light.insert(light.end(), byClass[c].begin(), byClass[c].end());
    }
}

int upload_adjacency(dory_engine *e, Adjacency &adj, const uint8_t *ptrs, const uint8_t *idx,
                     const uint8_t *vals, uint64_t nnz, uint32_t V, uint32_t nSrcRows) {
    adj.nnz = nnz;
    adj.n_src_rows = nSrcRows;
    std::vector<uint64_t> hp(V + 1);
    std::memcpy(hp.data(), ptrs, 8 * ((size_t)V + 1));
    for (uint32_t v = 0; v < V; ++v)
        if (hp[v + 1] < hp[v]) return fail(e, DORY_EFORMAT, "adjacency offsets decrease at vertex %u", v);
    CU(adj.ptrs.alloc(8 * ((size_t)V + 1)));
    CU(adj.idx.alloc(4 * nnz));
    CU(adj.vals.alloc(4 * nnz));
    CU(cudaMemcpyAsync(adj.ptrs.p, hp.data(), 8 * ((size_t)V + 1), cudaMemcpyHostToDevice, e->stream));
    if (nnz) {
        // validate indices on the host: an out-of-range id would be an out-of-bounds gather
        const uint8_t *p = idx;
        for (uint64_t i = 0; i < nnz; ++i, p += 4) {
            uint32_t s;
            std::memcpy(&s, p, 4);
            if (s >= nSrcRows) return fail(e, DORY_EFORMAT, "edge %llu references row %u >= %u",
                                           (unsigned long long)i, s, nSrcRows);
        }
        CU(cudaMemcpyAsync(adj.idx.p, idx, 4 * nnz, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(adj.vals.p, vals, 4 * nnz, cudaMemcpyHostToDevice, e->stream));
    }
    std::vector<uint32_t> heavy, light;
    // locality of the vertex numbering: share of edges whose source is within 1/64 of the row space
    // of its destination (a uniformly random graph has ~3 %)
    bool keepLocality = e->row_order == 2;
    if (e->row_order == 0 && nnz) {
        const uint64_t band = std::max<uint64_t>(1, nSrcRows / 64);
        const uint64_t stride = std::max<uint64_t>(1, nnz / (1u << 22));  // sample ~4 M edges
        uint64_t near = 0, seen = 0;
        uint32_t v = 0;
        for (uint64_t k = 0; k < nnz; k += stride) {
            while (hp[v + 1] <= k) ++v;
            uint32_t s;
            std::memcpy(&s, idx + 4 * k, 4);
            near += (s > v ? s - v : v - s) < band;
            ++seen;
        }
        keepLocality = near * 4 >= seen;  // >= 25 % of the edges are near-diagonal
    }
    // rows per locality block: their widest gathered slab must fit L2 with room to spare (24 MB)
    const uint32_t blockRows = e->locality_block ? e->locality_block
                                                 : std::max<uint32_t>(32768u, (24u << 20) / std::max(16u, e->max_slab_bytes()));
    build_row_lists(hp, e->heavy_degree, keepLocality, blockRows, heavy, light);
    adj.n_heavy = (uint32_t)heavy.size();
    adj.n_light = (uint32_t)light.size();
    // Hub rows.  The heavy launch keeps ~600 CTAs resident (148 SMs x 4); a row that holds more than
    // ~1/600 of the launch's edges cannot finish inside the launch's ideal duration even when it is
    // issued first, so rows above nnz / 2048 are split over a cluster of 8 CTAs.  The whole Reddit
    // shape has none (27 K < 114.6 M / 2048); one eighth of it has a few dozen per partition.
    const uint64_t hubDegree = e->hub_degree ? e->hub_degree : std::max<uint64_t>(2048, nnz / 2048);
    adj.n_vheavy = 0;  // `heavy` is degree-descending: the hubs are its prefix
    while (adj.n_vheavy < adj.n_heavy && hp[heavy[adj.n_vheavy] + 1] - hp[heavy[adj.n_vheavy]] >= hubDegree) ++adj.n_vheavy;
    {
        uint64_t lightEdges = 0;
        for (uint32_t v : light) lightEdges += hp[v + 1] - hp[v];
        adj.light_avg_degree = light.empty() ? 0 : (uint32_t)(lightEdges / light.size());
    }
    CU(adj.heavy.alloc(4 * heavy.size()));
    CU(adj.light.alloc(4 * light.size()));
    if (!heavy.empty())
        CU(cudaMemcpyAsync(adj.heavy.p, heavy.data(), 4 * heavy.size(), cudaMemcpyHostToDevice, e->stream));
    if (!light.empty())
        CU(cudaMemcpyAsync(adj.light.p, light.data(), 4 * light.size(), cudaMemcpyHostToDevice, e->stream));
    CU(cudaStreamSynchronize(e->stream));  // host staging vectors die at scope exit

    // ---- tile plan for the shared-memory-staged kernel (spmm_tile.cu): only kept when the vertex numbering
    // has enough locality for a window of source rows to serve a good share of a tile's edges
    adj.tile.reset();
    const bool tileLowDeg = V && nnz / V < 96;
    if (e->tile_mode && nnz && V >= 64 && !(e->tile_mode == 2 && tileLowDeg)) {
        const uint64_t avgDeg = nnz / V;
        const bool lowDeg = avgDeg < 96;
        // bytes per staged window row, widest layer this model aggregates
        uint32_t rowBytes = 0;
        int slabFloats = 0;
        const uint32_t lo = e->cfg.gnn_type == DORY_GCN ? 0 : 1, hi = e->cfg.gnn_type == DORY_GCN ? e->cfg.n_layers : e->cfg.n_layers + 1;
        if (lowDeg) {
            for (uint32_t l = lo; l < hi; ++l) {
                const uint32_t w = e->apply_first(l) ? e->cfg.dims[l + 1] : e->cfg.dims[l];
                const uint32_t nvec = (w + 3) / 4;
                if (nvec > 32) continue;  // that layer keeps the gather kernels
                rowBytes = std::max<uint32_t>(rowBytes, (uint32_t)(tile_smem_bytes(padded_ld(w), nvec, 64, true, 0) / 64));
            }
        } else {
            // 0 = per launch: 64-float slabs for rows up to 128 floats, 96-float slabs for wider ones (fewer
            // passes over the ids: 7 instead of 10 at F = 602); the window is sized for the wider of the two
            slabFloats = (int)e->tile_slab;
            rowBytes = (uint32_t)(slabFloats ? slabFloats : 96) * 4;
        }
        if (rowBytes) {
            TilePlanParams prm;
            prm.tileRows = e->tile_rows;
            prm.windowRows = e->tile_window;
            prm.maxWindowRows = std::max<uint32_t>(32, (e->tile_smem_kb * 1024u) / rowBytes);
            if (prm.windowRows > prm.maxWindowRows) prm.windowRows = prm.maxWindowRows;
            prm.teamDegree = lowDeg ? 0xffffffffu : e->tile_team;
            // rows the heavy (CTA-per-row) launch owns stay out of the tiles; a tile must be able to hold any other row
            prm.excludeDegree = lowDeg ? e->heavy_degree : 0;
            prm.edgeCap = lowDeg ? std::max(e->tile_edges, e->heavy_degree) : 0;
            prm.keepRowOrder = lowDeg;
            // mode 2: a cheap look first (every 16th tile, largest window the budget allows) -- a graph without
            // locality does not pay for the regrouped copy of its adjacency
            bool worth = true;
            if (e->tile_mode == 2) {
                uint32_t wmax = 32;
                while (wmax * 2 <= prm.maxWindowRows) wmax *= 2;
                const uint32_t w = prm.windowRows ? prm.windowRows : wmax;
                worth = estimate_tile_coverage(ptrs, idx, V, nSrcRows, std::max(16u, w / 2), w, 16) * 100.0 >= (double)e->tile_min_coverage;
            }
            TilePlanHost hp_;
            if (worth) build_tile_plan(ptrs, idx, vals, V, nSrcRows, prm, hp_);
            if (worth && (e->tile_mode == 1 || hp_.coverage() * 100.0 >= (double)e->tile_min_coverage)) {
                auto tb = std::make_unique<TilePlanBuf>();
                // 64 spare bytes behind every array: the low-degree kernel's bulk copies round their runs up to 16 B
                auto up = [&](DevBuf &b, const void *src, size_t bytes) -> cudaError_t {
                    cudaError_t c = b.alloc(bytes + 64);
                    if (c == cudaSuccess) c = cudaMemsetAsync(static_cast<uint8_t *>(b.p) + bytes, 0, 64, e->stream);
                    if (c == cudaSuccess && bytes) c = cudaMemcpyAsync(b.p, src, bytes, cudaMemcpyHostToDevice, e->stream);
                    return c;
                };
                CU(up(tb->ptrs, hp_.ptrs.data(), 8 * hp_.ptrs.size()));
                CU(up(tb->idx, hp_.idx.data(), 4 * hp_.idx.size()));
                CU(up(tb->vals, hp_.vals.data(), 4 * hp_.vals.size()));
                CU(up(tb->rows, hp_.rows.data(), 4 * hp_.rows.size()));
                CU(up(tb->tptr, hp_.tilePtr.data(), 4 * hp_.tilePtr.size()));
                CU(up(tb->tteam, hp_.tileTeam.data(), 4 * hp_.tileTeam.size()));
                CU(up(tb->twlo, hp_.tileWlo.data(), 4 * hp_.tileWlo.size()));
                CU(up(tb->twrows, hp_.tileWrows.data(), 4 * hp_.tileWrows.size()));
                CU(up(tb->te0, hp_.tileE0.data(), 8 * hp_.tileE0.size()));
                CU(up(tb->te1, hp_.tileE1.data(), 8 * hp_.tileE1.size()));
                tb->max_tile_rows = hp_.maxTileRows;
                tb->max_tile_edges = hp_.maxTileEdges;
                CU(cudaStreamSynchronize(e->stream));
                tb->n_tiles = (uint32_t)hp_.tileTeam.size();
                tb->max_wrows = hp_.maxWrows;
                tb->tile_rows = hp_.tileRows;
                tb->window_rows = hp_.windowRows;
                tb->coverage = hp_.coverage();
                tb->low_degree = lowDeg;
                tb->slab_floats = slabFloats;
                adj.tile = std::move(tb);
            }
        }
    }

    // ---- source-blocked copy.  A 128-float slab of the [V+G] x F block is (V+G) * 512 B; it only
    // stays L2-resident when that is well below the 126 MB L2 (the two L2 halves mirror lines that
    // both dies read, so the useful capacity for a chip-wide gather is about half).  Split the source
    // rows into nb windows of <= kWindowBytes and regroup each row's edges by window.
    // Windows are sized for the WIDEST aggregated row slab of this model (<= 512 B); layers with
    // narrower rows walk `span` consecutive windows per launch (aggregate_gcn), which works because a
    // row's windows are stored back to back.
    uint32_t nb = e->src_blocks;
    const uint64_t avgDeg = V ? nnz / V : 0;
    if (nb == 0) {
        nb = (uint32_t)(((uint64_t)nSrcRows * e->max_slab_bytes() + kWindowBytes - 1) / kWindowBytes);
        // every window costs one more pass over the rows (offsets, read-modify-write of `out`, a
        // partly filled 32-edge batch): only worth it while a (row, window) still holds ~64 edges.
        // Reddit (degree 492): 2 windows; Amazon / Friendster shapes (degree ~25): none
        // (tools/shape_bench.py: 10 windows made the Amazon-shape aggregation 4x slower).
        nb = (uint32_t)std::min<uint64_t>(nb, std::max<uint64_t>(1, avgDeg / 64));
    }
    nb = std::max(1u, std::min(nb, 64u));
    // Several partitions (GCN, option "overlap"): the window boundaries are laid so that one of them falls
    // between the partition's own rows [0, V) and its ghost rows [V, V + G).  nbL windows of local rows,
    // nbG of ghost rows, each sized like the plain windows (and at least one of either).
    uint32_t nbL = 0, nbG = 0;
    const uint32_t G = nSrcRows - V;
    if (e->cfg.num_nodes > 1 && e->overlap && e->cfg.gnn_type == DORY_GCN && G > 0 && nnz && e->src_blocks == 0) {
        const uint64_t cap = std::max<uint64_t>(1, avgDeg / 64);
        nbL = (uint32_t)std::min<uint64_t>(cap, std::max<uint64_t>(1, ((uint64_t)V * e->max_slab_bytes() + kWindowBytes - 1) / kWindowBytes));
        nbG = (uint32_t)std::min<uint64_t>(cap, std::max<uint64_t>(1, ((uint64_t)G * e->max_slab_bytes() + kWindowBytes - 1) / kWindowBytes));
        nb = nbL + nbG;
    }
    adj.nb = 1;
    adj.nb_local = 0;
    if (nb > 1 && nnz && (e->cfg.gnn_type == DORY_GCN || e->gat_windows)) {
        const uint32_t rowsPerBlock = (nSrcRows + nb - 1) / nb;
        const uint32_t rowsPerL = nbL ? (V + nbL - 1) / nbL : 1, rowsPerG = nbG ? (G + nbG - 1) / nbG : 1;
        auto block_of = [&](uint32_t src) -> uint32_t {
            if (!nbL) return src / rowsPerBlock;
            return src < V ? src / rowsPerL : nbL + (src - V) / rowsPerG;
        };
        std::vector<uint64_t> bp((size_t)V * nb + 1);
        std::vector<uint32_t> bi(nnz);
        std::vector<float> bv(nnz);
        const unsigned nt = std::max(1u, std::min(std::thread::hardware_concurrency(), 32u));
        // every (row, block) segment stays inside the row's original [hp[v], hp[v+1]) range, so rows
        // can be regrouped independently: stable counting sort by block id
        auto work = [&](unsigned t) {
            std::vector<uint64_t> cnt(nb);
            for (uint32_t v = (uint32_t)((uint64_t)V * t / nt); v < (uint32_t)((uint64_t)V * (t + 1) / nt); ++v) {
                std::fill(cnt.begin(), cnt.end(), 0);
                for (uint64_t k = hp[v]; k < hp[v + 1]; ++k) {
                    uint32_t s;
                    std::memcpy(&s, idx + 4 * k, 4);
                    ++cnt[block_of(s)];
                }
                uint64_t off = hp[v];
                for (uint32_t b = 0; b < nb; ++b) {
                    bp[(size_t)v * nb + b] = off;
                    const uint64_t c = cnt[b];
                    cnt[b] = off;  // becomes the write cursor
                    off += c;
                }
                for (uint64_t k = hp[v]; k < hp[v + 1]; ++k) {
                    uint32_t s;
                    float w;
                    std::memcpy(&s, idx + 4 * k, 4);
                    std::memcpy(&w, vals + 4 * k, 4);
                    const uint64_t pos = cnt[block_of(s)]++;
                    bi[pos] = s;
                    bv[pos] = w;
                }
            }
        };
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nt; ++t) th.emplace_back(work, t);
        for (auto &x : th) x.join();
        bp[(size_t)V * nb] = nnz;
        CU(adj.bptrs.alloc(8 * bp.size()));
        CU(adj.bidx.alloc(4 * nnz));
        CU(adj.bvals.alloc(4 * nnz));
        CU(cudaMemcpyAsync(adj.bptrs.p, bp.data(), 8 * bp.size(), cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(adj.bidx.p, bi.data(), 4 * nnz, cudaMemcpyHostToDevice, e->stream));
        CU(cudaMemcpyAsync(adj.bvals.p, bv.data(), 4 * nnz, cudaMemcpyHostToDevice, e->stream));
        CU(cudaStreamSynchronize(e->stream));
        adj.nb = nb;
        adj.nb_local = nbL;
    }
    return DORY_OK;
}

int alloc_weight(dory_engine *e, WeightSet &w, uint32_t rows, uint32_t cols) {
    w.rows = rows;
    w.cols = cols;
    w.ld = padded_ld(cols);
    w.prows = padded_ld(rows);
    const size_t bytes = w.floats() * sizeof(float);
    for (DevBuf *b : {&w.w, &w.dw, &w.m, &w.v}) {
        CU(b->alloc(bytes));
        CU(cudaMemsetAsync(b->p, 0, bytes, e->stream));
    }
    return DORY_OK;
}

// Engine::preallocateGCN, gcn_ops.cpp:27-93.  Local rows and ghost rows of one logical source
// tensor share ONE allocation ([V + G] rows) so that adjacency indices address it directly.
int preallocate_gcn(dory_engine *e) {
    const uint32_t L = e->L(), V = e->V;
    cudaError_t ce = cudaSuccess;
    auto T = [&](uint32_t layer, const char *name, const DevMat &m) { e->tensors[{layer, name}] = m; };
    DevMat x = new_tensor(e, (uint64_t)V + e->Gs, e->dim(0), ce);
    CU(ce);
    T(0, "x", x.rows_from(0, V));
    T(0, "fg", x.rows_from(V, e->Gs));
    for (uint32_t l = 0; l < L; ++l) {
        if (e->apply_first(l)) {
            // apply-first layer: t = in . W with its ghost block, dL/dz with ITS ghost block, and the
            // two aggregation outputs (z, u); "ah" / "grad" / "bg" of this layer are never formed
            const uint32_t nf = e->dim(l + 1);
            DevMat t = new_tensor(e, (uint64_t)V + e->Gs, nf, ce);
            CU(ce);
            T(l, "t", t.rows_from(0, V));
            T(l, "fg_t", t.rows_from(V, e->Gs));
            DevMat g = new_tensor(e, (uint64_t)V + e->Gd, nf, ce);
            CU(ce);
            T(l, "g", g.rows_from(0, V));
            T(l, "bg_g", g.rows_from(V, e->Gd));
            DevMat u = new_tensor(e, V, nf, ce);
            CU(ce);
            T(l, "u", u);
            if (l + 1 == L) {  // logits of the last layer (the reference order keeps them in scratch)
                DevMat z = new_tensor(e, V, nf, ce);
                CU(ce);
                T(l, "z", z);
            }
        } else {
            DevMat ah = new_tensor(e, V, e->dim(l), ce);
            CU(ce);
            T(l, "ah", ah);
        }
        if (l + 1 < L) {
            DevMat z = new_tensor(e, V, e->dim(l + 1), ce);
            CU(ce);
            DevMat h = new_tensor(e, (uint64_t)V + e->Gs, e->dim(l + 1), ce);
            CU(ce);
            T(l, "z", z);
            T(l, "h", h.rows_from(0, V));
            T(l + 1, "fg", h.rows_from(V, e->Gs));
        }
    }
    DevMat lab = new_tensor(e, V, e->dim(L), ce);
    CU(ce);
    T(L - 1, "lab", lab);
    for (uint32_t l = L - 1; l > 0; --l) {
        if (!e->apply_first(l)) {
            DevMat grad = new_tensor(e, (uint64_t)V + e->Gd, e->dim(l), ce);
            CU(ce);
            T(l, "grad", grad.rows_from(0, V));
            T(l - 1, "bg", grad.rows_from(V, e->Gd));
        }
        DevMat aTg = new_tensor(e, V, e->dim(l), ce);
        CU(ce);
        T(l - 1, "aTg", aTg);
    }
    return DORY_OK;
}

// Engine::preallocateGAT, gat_ops.cpp:27-115.
int preallocate_gat(dory_engine *e) {
    const uint32_t L = e->L(), V = e->V;
    cudaError_t ce = cudaSuccess;
    auto T = [&](uint32_t layer, const char *name, const DevMat &m) { e->tensors[{layer, name}] = m; };
    DevMat h0 = new_tensor(e, V, e->dim(0), ce);
    CU(ce);
    T(0, "h", h0);
    for (uint32_t l = 0; l < L; ++l) {
        const uint32_t nf = e->dim(l + 1);
        DevMat z = new_tensor(e, (uint64_t)V + e->Gs, nf, ce);
        CU(ce);
        T(l, "z", z.rows_from(0, V));
        T(l, "fg_z", z.rows_from(V, e->Gs));
        // per-edge vectors (E x 1 in the reference) are stored densely: ld = 1
        for (const char *nm : {"az", "dA"}) {
            DevMat m;
            auto buf = std::make_unique<DevBuf>();
            CU(buf->alloc(4 * (size_t)std::max<uint64_t>(e->Ein, 1)));
            CU(cudaMemsetAsync(buf->p, 0, buf->bytes, e->stream));
            m.p = buf->as<float>();
            m.rows = e->Ein;
            m.cols = 1;
            m.ld = 1;
            e->pool.push_back(std::move(buf));
            T(l, nm, m);
        }
        {   // "A" aliases forwardAdj.values for every layer (quirk Q12, gat_ops.cpp:61-64)
            DevMat m;
            m.p = e->fwd.vals.as<float>();
            m.rows = e->Ein;
            m.cols = 1;
            m.ld = 1;
            T(l, "A", m);
        }
        DevMat ah = new_tensor(e, V, nf, ce);
        CU(ce);
        T(l, "ah", ah);
        DevMat grad = new_tensor(e, (uint64_t)V + e->Gd, nf, ce);
        CU(ce);
        T(l, "grad", grad.rows_from(0, V));
        T(l, "bg_d", grad.rows_from(V, e->Gd));
        DevMat aTg = new_tensor(e, V, nf, ce);
        CU(ce);
        T(l, "aTg", aTg);
    }
    DevMat lab = new_tensor(e, V, e->dim(L), ce);
    CU(ce);
    T(L - 1, "lab", lab);
    return DORY_OK;
}

bool is_edge_vector(const DevMat &m) { return m.ld == 1 && m.cols == 1; }

// Host rows are dense (pitch = cols), HBM rows are padded (pitch = ld).  Large tensors cross PCIe as
// ONE contiguous DMA into a staging buffer and are re-pitched by a kernel: a strided
// cudaMemcpy2D of 2408-byte rows runs at a fraction of the link rate.
constexpr size_t kStageThreshold = 1u << 20;

int copy_in(dory_engine *e, const DevMat &m, const float *host) {
    if (m.rows == 0) return DORY_OK;
    const size_t dense = (size_t)m.rows * m.cols * 4;
    if (is_edge_vector(m) || m.ld == m.cols) {
        CU(cudaMemcpyAsync(m.p, host, dense, cudaMemcpyHostToDevice, e->stream));
    } else if (dense >= kStageThreshold) {
        if (e->stage.bytes < dense) CU(e->stage.alloc(dense));
        CU(cudaMemcpyAsync(e->stage.p, host, dense, cudaMemcpyHostToDevice, e->stream));
        LAUNCHED(launch_repitch(e->stage.as<float>(), m.cols, m.p, m.ld, m.rows, m.cols, e->stream));
    } else {
        CU(cudaMemcpy2DAsync(m.p, (size_t)m.ld * 4, host, (size_t)m.cols * 4, (size_t)m.cols * 4, m.rows,
                             cudaMemcpyHostToDevice, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));
    return DORY_OK;
}

int copy_out(dory_engine *e, const DevMat &m, float *host) {
    if (m.rows == 0) return DORY_OK;
    const size_t dense = (size_t)m.rows * m.cols * 4;
    if (is_edge_vector(m) || m.ld == m.cols) {
        CU(cudaMemcpyAsync(host, m.p, dense, cudaMemcpyDeviceToHost, e->stream));
    } else if (dense >= kStageThreshold) {
        if (e->stage.bytes < dense) CU(e->stage.alloc(dense));
        LAUNCHED(launch_repitch(m.p, m.ld, e->stage.as<float>(), m.cols, m.rows, m.cols, e->stream));
        CU(cudaMemcpyAsync(host, e->stage.p, dense, cudaMemcpyDeviceToHost, e->stream));
    } else {
        CU(cudaMemcpy2DAsync(host, (size_t)m.cols * 4, m.p, (size_t)m.ld * 4, (size_t)m.cols * 4, m.rows,
                             cudaMemcpyDeviceToHost, e->stream));
    }
    CU(cudaStreamSynchronize(e->stream));
    return DORY_OK;
}

// AdamOptimizer::nextIteration, AdamOptimizer.cpp:29-34 (float/double mix as in the reference:
// pow and sqrt are the double overloads, the results are narrowed into float members).
void adam_next_iteration(dory_engine *e) {
    ++e->adam_epochs;
    const float b1p = (float)std::pow((double)e->beta1, (double)e->adam_epochs);
    const float b2p = (float)std::pow((double)e->beta2, (double)e->adam_epochs);
    e->lr_t = (float)((double)e->cfg.learning_rate * std::sqrt((double)(1 - b2p)) / (double)(1 - b1p));
}

int join_comm(dory_engine *e);

// join: every operator but Gather and Scatter first waits for an exchange still in flight on the exchange
// stream (those two order themselves against it)
int check_loaded(dory_engine *e, bool join = true) {
    if (!e) return DORY_EINVAL;
    if (!e->loaded) return fail(e, DORY_ESTATE, "no partition loaded (call dory_load_partition first)");
    return join ? join_comm(e) : DORY_OK;
}

SpmmArgs spmm_args(const dory_engine *e, const Adjacency &adj, const float *selfw, int mode, const DevMat &src,
                   const DevMat &out, uint32_t low, uint32_t up, uint32_t V) {
    SpmmArgs a{};
    a.cfg_lg = e->spmm_lg;
    a.cfg_vec = e->spmm_vec;
    a.cfg_unroll = e->spmm_unroll;
    a.cfg_occ = e->spmm_occ;
    a.cfg_light = e->spmm_light;
    a.light_avg_degree = adj.light_avg_degree;
    a.ptrs = adj.ptrs.as<uint64_t>();
    a.ptr_stride = 1;
    a.ptr_off = 0;
    a.ptr_span = 1;
    a.idx = adj.idx.as<uint32_t>();
    a.vals = adj.vals.as<float>();
    a.selfw = selfw;
    a.src = src.p;
    a.src_rows = adj.n_src_rows;
    a.out = out.p;
    a.ld = src.ld;
    // float4 columns that carry data: the padding columns of a row (zero on both sides, never written)
    // are not gathered -- a quarter of every row at F = 48 / 41 / 100 (pitches 64 / 64 / 128)
    a.nvec = std::min<uint32_t>(src.ld / 4, (std::max(src.cols, out.cols) + 3) / 4);
    a.self_mode = mode;
    if (low == 0 && up == V) {
        a.heavy = adj.heavy.as<uint32_t>();
        a.n_heavy = adj.n_heavy;
        a.n_vheavy = adj.n_vheavy;
        a.light = adj.light.as<uint32_t>();
        a.n_light = adj.n_light;
        a.low = 0;
    } else {  // sub-range chunk (Lambda-style chunking): natural order, warp per row
        a.heavy = nullptr;
        a.n_heavy = 0;
        a.light = nullptr;
        a.n_light = up - low;
        a.low = low;
    }
    return a;
}

uint64_t edges_in_range(dory_engine *e, const Adjacency &adj, uint32_t low, uint32_t up) {
    if (low == 0 && up == e->V) return adj.nnz;
    return 0;  // partial chunks are not counted (would need a device read)
}

// ------------------------------------------------------------------ GCN operators
// One aggregation out = self(mode) + A . src over `adj`, rows [low, up), walking the source windows of
// the adjacency when it has them.  `vals` overrides the adjacency's edge values; with windows that is
// only valid for values that are constant along a destination row (the regrouped copy permutes a row's
// edges inside the row's own range) -- true of the GAT attention values and their gradients (quirk Q8).
// phase: 0 = the whole aggregation; 1 = self term + edges from the partition's own rows (the windows below
// nb_local); 2 = edges from ghost rows, accumulated onto phase 1's output.  Phases 1 and 2 need nb_local > 0.
int run_spmm(dory_engine *e, const Adjacency *adj, const float *selfw, int mode, const DevMat *src, const DevMat *out,
             uint32_t low, uint32_t up, const float *vals = nullptr, int phase = 0) {
    SpmmArgs a = spmm_args(e, *adj, selfw, mode, *src, *out, low, up, e->V);
    if (vals) a.vals = vals;
    if (!spmm_shape_supported(a.cfg_lg, a.cfg_vec))
        return fail(e, DORY_EINVAL, "options spmm_lg = %d, spmm_vec = %d: no aggregation kernel of that shape (lanes x float4 "
                    "per lane: 4 x {1,2,4}, 8 x {1,2,4}, 16 x {1,2}, 32 x {1,2,4}; set both or neither)", a.cfg_lg, a.cfg_vec);
    if (adj->tile && !vals && low == 0 && up == e->V && phase == 0) {
        // shared-memory-staged kernel (spmm_tile.cu) when this layer's rows fit its launch limits
        const TilePlanBuf &tb = *adj->tile;
        const bool shapeOk = tb.low_degree ? a.nvec <= 32 : true;
        size_t smem = tile_smem_bytes(a.ld, a.nvec, tb.max_wrows, tb.low_degree, tb.slab_floats ? tb.slab_floats : (a.nvec <= 32 ? 64 : 96)) +
                      (tb.low_degree ? tile_edge_smem_bytes(tb.max_tile_edges, tb.max_tile_rows) : 0);
        if (tb.low_degree && e->tile_pipe) smem = 2 * (smem + 256);  // two stages
        if (shapeOk && smem <= 200u * 1024u) {
            TilePlanDev t{};
            t.ptrs = tb.ptrs.as<uint64_t>();
            t.idx = tb.idx.as<uint32_t>();
            t.vals = tb.vals.as<float>();
            t.rows = tb.rows.as<uint32_t>();
            t.tile_ptr = tb.tptr.as<uint32_t>();
            t.tile_team = tb.tteam.as<uint32_t>();
            t.tile_wlo = tb.twlo.as<uint32_t>();
            t.tile_wrows = tb.twrows.as<uint32_t>();
            t.tile_e0 = tb.te0.as<uint64_t>();
            t.tile_e1 = tb.te1.as<uint64_t>();
            t.max_tile_rows = tb.max_tile_rows;
            t.max_tile_edges = tb.max_tile_edges;
            t.n_tiles = tb.n_tiles;
            t.max_wrows = tb.max_wrows;
            t.low_degree = tb.low_degree;
            t.pipeline = e->tile_pipe;
            t.slab_floats = tb.slab_floats ? tb.slab_floats : (a.nvec <= 32 ? 64 : 96);
            const int n = launch_spmm_tile(a, t, e->stream);
            LAUNCHED(n);
            if (n > 0) {  // 0: no tile kernel for this shape -> the gather kernels below
                if (tb.low_degree && adj->n_heavy) {  // rows the plan leaves out: CTA-per-row kernel of spmm.cu
                    SpmmArgs h = a;
                    h.n_light = 0;
                    LAUNCHED(launch_spmm_rows(h, e->stream));
                }
                e->stats.edges_aggregated += adj->nnz;
                return DORY_OK;
            }
        }
    }
    if (adj->nb > 1 && !(phase == 0 && adj->nb_local == 1 && adj->nb == 2 && !vals)) {
        // (the [own rows | ghost rows] pair alone is only walked in two passes when an exchange is in flight:
        // with nothing to wait for, the unsplit edge list below is one pass over the rows instead of two)
        // one pass per group of source windows; passes are separate launches (stream order) because
        // they accumulate into the same output rows.  A group holds as many windows as keep
        // rows x slab bytes within the L2 budget for THIS layer's row width.
        const uint32_t slab = std::min<uint32_t>(src->ld * 4, 512);
        const uint32_t span = e->src_blocks ? 1 : std::max(1u, e->max_slab_bytes() / slab);
        a.ptrs = adj->bptrs.as<uint64_t>();
        a.idx = adj->bidx.as<uint32_t>();
        a.vals = vals ? vals : adj->bvals.as<float>();
        a.ptr_stride = adj->nb;
        const uint32_t b_begin = phase == 2 ? adj->nb_local : 0, b_end = phase == 1 ? adj->nb_local : adj->nb;
        for (uint32_t b = b_begin; b < b_end;) {
            // a launch never spans the local / ghost boundary (the two sides may run at different times)
            const uint32_t limit = (adj->nb_local && b < adj->nb_local) ? adj->nb_local : b_end;
            a.ptr_off = b;
            a.ptr_span = std::min(span, limit - b);
            a.self_mode = b == 0 ? mode : SELF_ACCUM;
            LAUNCHED(launch_spmm(a, e->stream));
            b += a.ptr_span;
        }
    } else {
        LAUNCHED(launch_spmm(a, e->stream));
    }
    if (phase != 2) e->stats.edges_aggregated += edges_in_range(e, *adj, low, up);
    return DORY_OK;
}

// Makes the compute stream wait for whatever the exchange stream still has in flight.
int join_comm(dory_engine *e) {
    if (!e->comm_pending) return DORY_OK;
    CU(cudaStreamWaitEvent(e->stream, e->ev_comm, 0));
    e->comm_pending = false;
    e->comm_pending_ghost = nullptr;
    return DORY_OK;
}

int run_gcn_spmm(dory_engine *e, const Adjacency *adj, const DevMat *src, const DevMat *out, const DevMat *ghost, uint32_t low,
                 uint32_t up) {
    const float *nw = e->norms.as<float>();
    // the exchange that fills this aggregation's ghost block is still in flight: local rows first
    if (e->comm_pending && ghost && e->comm_pending_ghost == ghost->p && adj->nb_local > 0 && low == 0 && up == e->V) {
        int rc = run_spmm(e, adj, nw, SELF_NORM, src, out, low, up, nullptr, 1);
        if (rc) return rc;
        if ((rc = join_comm(e))) return rc;
        return run_spmm(e, adj, nw, SELF_NORM, src, out, low, up, nullptr, 2);
    }
    int rc = join_comm(e);
    if (rc) return rc;
    return run_spmm(e, adj, nw, SELF_NORM, src, out, low, up);
}

int softmax_ce_gcn(dory_engine *e, const float *logits, const DevMat &lab, float *d);

int aggregate_gcn(dory_engine *e, const dory_chunk *c) {
    const uint32_t L = e->L();
    if (c->upBound > e->V || c->lowBound > c->upBound) return fail(e, DORY_EINVAL, "chunk bounds out of range");
    const DevMat *src, *out, *ghost;
    const Adjacency *adj;
    const bool af = e->apply_first(c->layer);
    if (c->dir == DORY_FORWARD) {
        if (c->layer >= L) return fail(e, DORY_EINVAL, "aggregate: forward layer %u out of range", c->layer);
        if (af) {  // z = A_hat [t; fg_t], then the activation (include/dorylus_b200.h: apply-first schedule)
            if (c->lowBound != 0 || c->upBound != e->V)
                return fail(e, DORY_EINVAL, "aggregate: layer %u runs apply-first, which takes whole-partition chunks", c->layer);
            src = find_tensor(e, c->layer, "t");
            out = find_tensor(e, c->layer, "z");
            ghost = find_tensor(e, c->layer, "fg_t");
        } else {
            src = c->layer == 0 ? find_tensor(e, 0, "x") : find_tensor(e, c->layer - 1, "h");
            out = find_tensor(e, c->layer, "ah");
            ghost = find_tensor(e, c->layer, "fg");
        }
        adj = &e->fwd;
    } else {
        if (c->layer >= L || (c->layer == 0 && !af)) return fail(e, DORY_EINVAL, "aggregate: backward layer %u out of range", c->layer);
        if (af) {  // u = A_hat^T [g; bg_g]
            src = find_tensor(e, c->layer, "g");
            out = find_tensor(e, c->layer, "u");
            ghost = find_tensor(e, c->layer, "bg_g");
        } else {
            src = find_tensor(e, c->layer, "grad");
            out = find_tensor(e, c->layer - 1, "aTg");
            ghost = find_tensor(e, c->layer - 1, "bg");
        }
        adj = &e->bwd;
    }
    if (ghost) e->ghost_reads_pending.insert(ghost->p);
    int rc = run_gcn_spmm(e, adj, src, out, ghost, c->lowBound, c->upBound);
    if (rc || !af || c->dir != DORY_FORWARD) return rc;
    if (c->layer + 1 < L) {  // h = tanh(z), activate(): CPU_comm.cpp:265-274
        const DevMat &h = *find_tensor(e, c->layer, "h");
        LAUNCHED(launch_tanh_forward(out->p, h.p, (uint64_t)e->V * out->ld, e->stream));
        return DORY_OK;
    }
    return softmax_ce_gcn(e, out->p, *find_tensor(e, c->layer, "lab"), find_tensor(e, c->layer, "g")->p);
}

bool use_tensor_cores(const dory_engine *e) { return e->tensor_cores && !(e->cfg.flags & DORY_FLAG_NO_TENSOR_CORES); }

// Z = A . W (+ tanh): tcgen05 path when the shape qualifies, else fp32 SIMT.
int gemm_nn(dory_engine *e, const DevMat &A, const WeightSet &W, const DevMat &C, const DevMat *C2) {
    if (use_tensor_cores(e)) {
        int n = launch_gemm_tc(A.p, A.ld, A.rows, W.w.as<float>(), W.ld, W.prows, C.p, C2 ? C2->p : nullptr, C.ld,
                               C2 ? EPI_TANH : EPI_NONE, e->stream, e->tc_small ? e->tc_stages : -1);
        if (n > 0) {
            e->stats.kernel_launches += n;
            return DORY_OK;
        }
        if (n < 0) return fail(e, DORY_ECUDA, "tcgen05 GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        // n == 0: shape not supported by the tensor-core kernel -> SIMT path below
    }
    GemmArgs g{};
    g.A = A.p; g.lda = A.ld; g.B = W.w.as<float>(); g.ldb = W.ld; g.C = C.p; g.ldc = C.ld;
    g.C2 = C2 ? C2->p : nullptr;
    g.M = A.rows; g.N = C.ld; g.K = A.ld;
    g.transA = false; g.transB = false; g.epilogue = C2 ? EPI_TANH : EPI_NONE;
    LAUNCHED(launch_gemm(g, e->stream));
    return DORY_OK;
}

// C = G . W^T   (G: V x Fout, W: Fin x Fout -> C: V x Fin)
int gemm_nt(dory_engine *e, const float *G, uint32_t ldg, uint64_t rows, const WeightSet &W, const DevMat &C) {
    if (use_tensor_cores(e) && e->tc_small) {
        const int n = launch_gemm_nt_tc(G, ldg, rows, W.w.as<float>(), W.ld, W.prows, C.p, C.ld, e->tc_stages, e->stream);
        if (n > 0) {
            e->stats.kernel_launches += n;
            return DORY_OK;
        }
        if (n < 0) return fail(e, DORY_ECUDA, "tcgen05 GEMM (G . W^T) launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    GemmArgs g{};
    g.A = G; g.lda = ldg; g.B = W.w.as<float>(); g.ldb = W.ld; g.C = C.p; g.ldc = C.ld;
    g.M = rows; g.N = C.ld; g.K = W.ld;
    g.transA = false; g.transB = true; g.epilogue = EPI_NONE;
    LAUNCHED(launch_gemm(g, e->stream));
    return DORY_OK;
}

// dW = A^T . G   (A: V x Fin, G: V x Fout -> dW: Fin x Fout), deterministic split over vertices
// Gh != null (narrow kernel only, see tn_fusable): G is read as G (*) (1 - Gh^2), i.e. tanh' is applied on the way in.
bool tn_fusable(const dory_engine *e, const WeightSet &W) {
    return e->fuse_tanh_bwd && e->tn_small && W.prows <= 32 && e->gemm_ws.p != nullptr;
}

int gemm_tn(dory_engine *e, const DevMat &A, const float *G, uint32_t ldg, WeightSet &W, float *out,
            const float *Gh = nullptr) {
    // input width <= 32 (Friendster's 16): the narrow fp32 kernel (dense.cu: gemm_tn_small_kernel) -- the tcgen05
    // kernel needs a multiple of 32 input columns and its 128-row M tile would be three quarters padding.  From 64
    // columns on the tcgen05 kernel wins at every vertex count measured with CUDA events in the operator's own
    // sequence (profiles/round2_dense_skinny.md: 8.2 M x 64 x 64 -- 1.3 ms against 2.1 ms; the 2.7 ms of an earlier
    // cold ncu pass had put the rule at 4 M vertices).
    const bool narrow = e->tn_small && W.prows <= 32;
    if (Gh && !tn_fusable(e, W)) return fail(e, DORY_EINVAL, "internal: fused tanh' operand outside the narrow dW kernel");
    if (use_tensor_cores(e) && !narrow) {
        int n = launch_gemm_tn_tc(A.p, A.ld, W.prows, G, ldg, A.rows, out, W.ld, e->gemm_ws.as<float>(),
                                  e->gemm_ws.bytes / 4, e->stream);
        if (n > 0) {
            e->stats.kernel_launches += n;
            return DORY_OK;
        }
        if (n < 0) return fail(e, DORY_ECUDA, "tcgen05 dW GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
    }
    GemmArgs g{};
    g.A = A.p; g.lda = A.ld; g.B = G; g.ldb = ldg; g.C = out; g.ldc = W.ld;
    g.M = W.prows; g.N = W.ld; g.K = A.rows;
    g.transA = true; g.transB = false; g.epilogue = EPI_NONE;
    g.ws = e->gemm_ws.as<float>(); g.ws_floats = e->gemm_ws.bytes / 4;
    g.Bh = Gh;
    LAUNCHED(launch_gemm(g, e->stream));
    return DORY_OK;
}

// Last-layer soft-max, validation statistics, maskout and gradient scale (CPU_comm.cpp:108-121):
// d = (maskout(softmax(logits)) - lab) / (V_global * 0.66).
SoftmaxCEArgs softmax_args(dory_engine *e, const float *logits, const DevMat &lab, float *d);

int softmax_ce_gcn(dory_engine *e, const float *logits, const DevMat &lab, float *d) {
    SoftmaxCEArgs s = softmax_args(e, logits, lab, d);
    LAUNCHED(launch_softmax_ce(s, e->stream));
    e->stats.val_rows = s.valEnd - s.trainEnd;
    return DORY_OK;
}

SoftmaxCEArgs softmax_args(dory_engine *e, const float *logits, const DevMat &lab, float *d) {
    SoftmaxCEArgs s{};
    s.z = logits; s.lab = lab.p; s.d = d; s.pred = nullptr;
    s.ld = lab.ld; s.C = lab.cols; s.V = e->V;
    s.trainEnd = (unsigned)(e->V * kTrainPortion);
    s.valEnd = s.trainEnd + (unsigned)(e->V * kValPortion);
    s.maskFloats = e->V - s.trainEnd;  // CPU_comm.cpp:470: sizeof(FeatType) * (end - stt)
    s.strictMask = (e->cfg.flags & DORY_FLAG_STRICT_MASK) != 0;
    s.denom = (float)(e->gV * kTrainPortion);  // CPU_comm.cpp:121
    s.rowstat = e->rowstat.as<float>();
    s.stats = e->stats_dev.as<float>();
    return s;
}

// The rows a layer's dense product reads: "x" (layer 0) or the previous layer's "h", local rows.
const DevMat &gcn_layer_input(dory_engine *e, uint32_t layer) {
    return layer == 0 ? *find_tensor(e, 0, "x") : *find_tensor(e, layer - 1, "h");
}

// CPUComm::vtxNNForwardGCN, CPU_comm.cpp:98-135
int vtx_forward_gcn(dory_engine *e, uint32_t layer) {
    const uint32_t L = e->L();
    if (layer >= L) return fail(e, DORY_EINVAL, "apply_vertex: layer %u out of range", layer);
    WeightSet &W = e->W[layer];
    if (e->apply_first(layer))  // t = in . W; aggregation and activation follow (aggregate_gcn)
        return gemm_nn(e, gcn_layer_input(e, layer), W, *find_tensor(e, layer, "t"), nullptr);
    const DevMat &ah = *find_tensor(e, layer, "ah");
    if (layer + 1 < L) {
        return gemm_nn(e, ah, W, *find_tensor(e, layer, "z"), find_tensor(e, layer, "h"));
    }
    // last layer: logits -> softmax / stats / maskout / scale -> grad, dW
    const DevMat &lab = *find_tensor(e, layer, "lab");
    DevMat logits = lab;  // same shape
    logits.p = e->scratchA.as<float>();
    DevMat d = lab;
    d.p = e->scratchB.as<float>();
    int rc = DORY_OK;
    bool fused = false;
    if (e->fuse_softmax) {  // logits + soft-max in one kernel when a row's classes fit one tile (C <= 64)
        SoftmaxCEArgs sa = softmax_args(e, nullptr, lab, d.p);
        int n = 0;
        if (use_tensor_cores(e) && e->tc_small && e->fuse_softmax != 2) {  // tcgen05 logits, thread-per-row epilogue out of TMEM
            n = launch_gemm_tc_softmax(ah.p, ah.ld, W.w.as<float>(), W.ld, W.prows, sa, e->tc_stages, e->stream);
            if (n < 0) return fail(e, DORY_ECUDA, "tcgen05 soft-max GEMM launch failed: %s", cudaGetErrorString(cudaGetLastError()));
            if (n > 0) {
                e->stats.kernel_launches += n;
                n = launch_softmax_stats(sa, e->stream);
            }
        }
        if (n == 0) n = launch_gemm_softmax_ce(ah.p, ah.ld, W.w.as<float>(), W.ld, ah.ld, sa, e->stream);
        LAUNCHED(n);
        if (n > 0) {
            fused = true;
            e->stats.val_rows = sa.valEnd - sa.trainEnd;
        }
    }
    if (!fused) {
        rc = gemm_nn(e, ah, W, logits, nullptr);
        if (rc) return rc;
        if ((rc = softmax_ce_gcn(e, logits.p, lab, d.p))) return rc;
    }
    if (layer > 0) {
        rc = gemm_nt(e, d.p, d.ld, e->V, W, *find_tensor(e, layer, "grad"));
        if (rc) return rc;
    }
    return gemm_tn(e, ah, d.p, d.ld, W, W.dw.as<float>());
}

// CPUComm::vtxNNBackwardGCN, CPU_comm.cpp:137-159
int vtx_backward_gcn(dory_engine *e, uint32_t layer) {
    const uint32_t L = e->L();
    if (layer + 1 >= L) return fail(e, DORY_EINVAL, "apply_vertex backward: layer %u out of range", layer);
    const DevMat &aTg = *find_tensor(e, layer, "aTg");
    const DevMat &h = *find_tensor(e, layer, "h");
    if (e->apply_first(layer)) {  // dL/dz into "g"; its aggregation and dW follow in this layer's backward pass
        LAUNCHED(launch_tanh_backward(aTg.p, h.p, find_tensor(e, layer, "g")->p, (uint64_t)e->V * aTg.ld, e->stream));
        return DORY_OK;
    }
    const DevMat &ah = *find_tensor(e, layer, "ah");
    WeightSet &W = e->W[layer];
    // layer 0 wants dW only (no gradient flows further down): with a narrow input the dW kernel forms
    // g = aTg (*) (1 - h^2) while loading it, and the pass that writes g (and its re-read) disappears
    if (layer == 0 && tn_fusable(e, W)) return gemm_tn(e, ah, aTg.p, aTg.ld, W, W.dw.as<float>(), h.p);
    float *g = e->scratchA.as<float>();
    LAUNCHED(launch_tanh_backward(aTg.p, h.p, g, (uint64_t)e->V * aTg.ld, e->stream));
    int rc = gemm_tn(e, ah, g, aTg.ld, W, W.dw.as<float>());
    if (rc) return rc;
    if (layer != 0) rc = gemm_nt(e, g, aTg.ld, e->V, W, *find_tensor(e, layer, "grad"));
    return rc;
}

// Backward apply of an apply-first layer: dW = in^T . u; l > 0: aTg[l-1] = u . W^T (= dL/dh[l-1]).
int vtx_backward_apply_first(dory_engine *e, uint32_t layer) {
    const DevMat &u = *find_tensor(e, layer, "u");
    WeightSet &W = e->W[layer];
    int rc = gemm_tn(e, gcn_layer_input(e, layer), u.p, u.ld, W, W.dw.as<float>());
    if (rc || layer == 0) return rc;
    return gemm_nt(e, u.p, u.ld, e->V, W, *find_tensor(e, layer - 1, "aTg"));
}

// Engine::scatterGCN + ghostReceiverGCN: rows of `name`[srcLayer] -> peers' ghost block.
int exchange(dory_engine *e, uint32_t dir, const DevMat &local, const DevMat &ghost) {
    if (e->cfg.num_nodes <= 1) return DORY_OK;
    if (!e->comm) return fail(e, DORY_ESTATE, "scatter with %u partitions needs dory_comm_init", e->cfg.num_nodes);
    int launches = 0;
    std::string msg;
    auto pit = e->peer_ghost.find(ghost.p);
    bool p2p = e->p2p && pit != e->peer_ghost.end() && e->comm->p2p_ready((int)dir);
    if (p2p)
        for (uint32_t q = 0; q < e->cfg.num_nodes; ++q)
            if (q != e->cfg.node_id && !pit->second[q]) p2p = false;
    if (p2p && !e->p2p_validated.count({ghost.p, dir})) {
        // the store kernel writes peerGhost[q] + slot * ld over NVLink: a slot beyond the block the peer
        // exported (images of another partitioning, a peer whose schedule differs) would land in memory
        // that is not ours to write.  Checked once per (tensor, direction) plan.
        const auto &rows = e->peer_ghost_rows[ghost.p];
        for (uint32_t q = 0; q < e->cfg.num_nodes; ++q) {
            if (q == e->cfg.node_id) continue;
            const uint32_t bound = e->comm->send_slot_bound((int)dir, (int)q);
            if (bound > rows[q])
                return fail(e, DORY_EINVAL, "peer-memory exchange: rows shipped to partition %u land in ghost slots up to %u, "
                            "but its block has %llu rows (send plan and peer image disagree)", q, bound - 1,
                            (unsigned long long)rows[q]);
        }
        e->p2p_validated.insert({ghost.p, dir});
    }
    // Overlap: the peer-memory exchange goes to the exchange stream (its collectives to the second
    // communicator), ordered behind everything the compute stream has been given so far; whoever needs its
    // result waits for ev_comm (join_comm / run_gcn_spmm).
    const bool async = p2p && e->overlap && e->cfg.gnn_type == DORY_GCN && e->comm->has_aux();
    cudaStream_t xs = e->stream;
    if (async) {
        if (!e->comm_stream) {
            int lo = 0, hi = 0;
            CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
            CU(cudaStreamCreateWithPriority(&e->comm_stream, cudaStreamNonBlocking, hi));
            CU(cudaEventCreateWithFlags(&e->ev_compute, cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&e->ev_comm, cudaEventDisableTiming));
        }
        CU(cudaEventRecord(e->ev_compute, e->stream));
        CU(cudaStreamWaitEvent(e->comm_stream, e->ev_compute, 0));
        xs = e->comm_stream;
    } else {
        int rc = join_comm(e);
        if (rc) return rc;
    }
    if (p2p) {
        const bool pre = !e->elide_pre_barrier || e->ghost_reads_pending.count(ghost.p) != 0;
        e->comm->set_p2p_variant(e->p2p_variant);
        e->comm->set_p2p_ctas_per_sm(e->p2p_ctas);
        msg = e->comm->exchange_p2p((int)dir, local.p, pit->second.data(), local.ld, xs, launches, pre, local.cols, async);
        if (pre) e->ghost_reads_pending.erase(ghost.p);  // that barrier followed every rank's reads of this block
    } else {
        msg = e->comm->exchange((int)dir, local.p, ghost.p, local.ld, xs, launches);
    }
    if (!msg.empty()) return fail(e, DORY_ECOMM, "%s", msg.c_str());
    if (async) {
        CU(cudaEventRecord(e->ev_comm, e->comm_stream));
        e->comm_pending = true;
        e->comm_pending_ghost = ghost.p;
    } else if (p2p) {
        // Only the peer-memory path on the COMPUTE stream ends in a collective that every rank's earlier reads
        // precede (the "writers done" all-reduce); a grouped ncclSend/ncclRecv synchronises the pairs that
        // exchange rows and nobody else, and a collective on the exchange stream does not order the peers'
        // compute streams -- after either, a later peer-memory exchange still needs its "readers done" barrier
        // for blocks read since (the dW all-reduce at the end of the epoch clears them).
        e->ghost_reads_pending.clear();
    }
    e->stats.kernel_launches += launches;
    return DORY_OK;
}

int scatter_gcn(dory_engine *e, const dory_chunk *c) {
    const uint32_t L = e->L();
    const bool af = e->apply_first(c->layer);
    if (c->dir == DORY_FORWARD) {  // gcn_ops.cpp:207-209: h[layer-1] -> fg[layer]
        if (c->layer >= L) return fail(e, DORY_EINVAL, "scatter: forward layer %u out of range", c->layer);
        // apply-first layer: what its aggregation gathers is t = in . W
        if (af) return exchange(e, DORY_FORWARD, *find_tensor(e, c->layer, "t"), *find_tensor(e, c->layer, "fg_t"));
        // layer 0 (no reference counterpart): the reference fills the layer-0 ghost rows of EVERY
        // partition from the feature file (readFeaturesFile, engine/utils.cpp:486-552).  A caller that
        // streams features uploads only the rows it owns and ships x -> the peers' fg[0] over NVLink.
        if (c->layer == 0) return exchange(e, DORY_FORWARD, *find_tensor(e, 0, "x"), *find_tensor(e, 0, "fg"));
        return exchange(e, DORY_FORWARD, *find_tensor(e, c->layer - 1, "h"), *find_tensor(e, c->layer, "fg"));
    }
    if (c->layer >= L || (c->layer == 0 && !af)) return fail(e, DORY_EINVAL, "scatter: backward layer %u out of range", c->layer);
    if (af) return exchange(e, DORY_BACKWARD, *find_tensor(e, c->layer, "g"), *find_tensor(e, c->layer, "bg_g"));
    return exchange(e, DORY_BACKWARD, *find_tensor(e, c->layer, "grad"), *find_tensor(e, c->layer - 1, "bg"));
}

void inc_layer_gcn(const dory_engine *e, dory_chunk *c) {  // engine/utils.cpp:714-732
    if (c->dir == DORY_FORWARD) {
        c->layer++;
        if (c->layer == e->L()) {
            c->dir = DORY_BACKWARD;
            c->layer--;
        }
    } else if (c->layer == 0) {
        c->dir = DORY_FORWARD;
        c->epoch++;
    } else {
        c->layer--;
    }
}

void inc_layer_gat(const dory_engine *, dory_chunk *c) {  // engine/utils.cpp:734-748
    if (c->dir == DORY_FORWARD) {
        c->layer++;
    } else if (c->layer == 0) {
        c->dir = DORY_FORWARD;
        c->vertex = 1;
        c->epoch++;
    } else {
        c->layer--;
    }
}

// ------------------------------------------------------------------ GAT operators
int aggregate_gat(dory_engine *e, const dory_chunk *c) {
    const uint32_t L = e->L();
    if (c->layer == 0 || c->layer > L) return fail(e, DORY_EINVAL, "aggregateGAT: layer %u out of range", c->layer);
    if (c->upBound > e->V || c->lowBound > c->upBound) return fail(e, DORY_EINVAL, "chunk bounds out of range");
    const uint32_t fl = c->layer - 1;
    const DevMat &z = *find_tensor(e, fl, "z");
    if (const DevMat *gh = find_tensor(e, fl, "fg_z")) e->ghost_reads_pending.insert(gh->p);
    if (c->dir == DORY_BACKWARD)
        if (const DevMat *gh = find_tensor(e, fl, "bg_d")) e->ghost_reads_pending.insert(gh->p);
    // The edge values of both forward-adjacency terms live in arrays of the ORIGINAL edge order ("A"
    // aliases forwardAdj.values, "dA"); they are passed explicitly so that the windowed walk (option
    // "gat_windows") can use them with the regrouped ids.
    if (c->dir == DORY_FORWARD)  // gat_ops.cpp:201-220: ah = z + sum A[e] z_src
        return run_spmm(e, &e->fwd, nullptr, SELF_ONE, &z, find_tensor(e, fl, "ah"), c->lowBound, c->upBound,
                        e->fwd.vals.as<float>());
    // gat_ops.cpp:221-241: aTg = sum_out bvals * grad_dst  +  sum_in dA * z_src   (zero-initialised, Q11)
    const DevMat &aTg = *find_tensor(e, fl, "aTg");
    int rc = run_spmm(e, &e->bwd, nullptr, SELF_ZERO, find_tensor(e, fl, "grad"), &aTg, c->lowBound, c->upBound);
    if (rc) return rc;
    return run_spmm(e, &e->fwd, nullptr, SELF_ACCUM, &z, &aTg, c->lowBound, c->upBound, find_tensor(e, fl, "dA")->p);
}

const DevMat &gat_layer_input(dory_engine *e, uint32_t layer) {  // CPU_comm.cpp:162-164
    return layer == 0 ? *find_tensor(e, 0, "h") : *find_tensor(e, layer - 1, "ah");
}

int vtx_forward_gat(dory_engine *e, uint32_t layer) {  // CPU_comm.cpp:161-169
    if (layer >= e->L()) return fail(e, DORY_EINVAL, "apply_vertex: layer %u out of range", layer);
    return gemm_nn(e, gat_layer_input(e, layer), e->W[layer], *find_tensor(e, layer, "z"), nullptr);
}

int vtx_backward_gat(dory_engine *e, uint32_t layer) {  // CPU_comm.cpp:171-188
    if (layer >= e->L()) return fail(e, DORY_EINVAL, "apply_vertex backward: layer %u out of range", layer);
    const DevMat &aTg = *find_tensor(e, layer, "aTg");
    WeightSet &W = e->W[layer];
    int rc = gemm_tn(e, gat_layer_input(e, layer), aTg.p, aTg.ld, W, W.dw.as<float>());
    if (rc) return rc;
    if (layer != 0) rc = gemm_nt(e, aTg.p, aTg.ld, e->V, W, *find_tensor(e, layer - 1, "grad"));
    return rc;
}

int edge_forward_gat(dory_engine *e, uint32_t layer) {  // CPU_comm.cpp:190-203
    if (layer >= e->L()) return fail(e, DORY_EINVAL, "apply_edge: layer %u out of range", layer);
    const DevMat &z = *find_tensor(e, layer, "z");
    LAUNCHED(launch_gat_edge_forward(z.p, z.ld, z.cols, e->Ai[layer].w.as<float>(), e->fwd.ptrs.as<uint64_t>(),
                                     e->V, find_tensor(e, layer, "az")->p, e->fwd.vals.as<float>(), e->stream));
    return DORY_OK;
}

int edge_backward_gat(dory_engine *e, uint32_t layer) {  // CPU_comm.cpp:205-242
    if (layer >= e->L()) return fail(e, DORY_EINVAL, "apply_edge backward: layer %u out of range", layer);
    const DevMat &z = *find_tensor(e, layer, "z");
    const DevMat &grad = *find_tensor(e, layer, "grad");
    GatEdgeBackwardArgs a{};
    a.grad = grad.p; a.z = z.p; a.ld = z.ld; a.F = z.cols; a.V = e->V;
    a.a = e->Ai[layer].w.as<float>();
    a.az = find_tensor(e, layer, "az")->p;
    a.colPtrs = e->fwd.ptrs.as<uint64_t>();
    a.dA = find_tensor(e, layer, "dA")->p;
    a.da = e->Ai[layer].dw.as<float>();
    a.scratch = e->scratchA.as<float>();
    a.scratch_floats = e->scratchA.bytes / 4;
    a.ws = e->gemm_ws.as<float>();
    a.ws_floats = e->gemm_ws.bytes / 4;
    if (a.scratch_floats < (size_t)a.V + 2 * a.ld)
        return fail(e, DORY_EINVAL, "apply_edge backward: partition of %u vertices is too small for the edge-gradient workspace "
                    "(needs V * pitch >= V + 2 * pitch floats)", a.V);
    LAUNCHED(launch_gat_edge_backward(a, e->stream));
    return DORY_OK;
}

int scatter_gat(dory_engine *e, const dory_chunk *c) {  // gat_ops.cpp:277-333
    if (c->layer == 0 || c->layer > e->L()) return fail(e, DORY_EINVAL, "scatterGAT: layer %u out of range", c->layer);
    const uint32_t ol = c->layer - 1;
    if (c->dir == DORY_FORWARD)
        return exchange(e, DORY_FORWARD, *find_tensor(e, ol, "z"), *find_tensor(e, ol, "fg_z"));
    return exchange(e, DORY_BACKWARD, *find_tensor(e, ol, "grad"), *find_tensor(e, ol, "bg_d"));
}

int predict_gat(dory_engine *e, const dory_chunk *c) {  // gat_ops.cpp:247-265
    if (c->layer == 0 || c->layer > e->L()) return fail(e, DORY_EINVAL, "predictGAT: layer %u out of range", c->layer);
    if (c->upBound > e->V || c->lowBound > c->upBound) return fail(e, DORY_EINVAL, "chunk bounds out of range");
    const uint32_t fl = c->layer - 1;
    const DevMat *lab = find_tensor(e, fl, "lab");
    if (!lab) return fail(e, DORY_EINVAL, "predictGAT: no labels at layer %u", fl);
    const DevMat &grad = *find_tensor(e, fl, "grad");
    const float *logits;
    uint32_t ldl;
    if (e->cfg.flags & DORY_FLAG_GAT_PREDICT_AH) {
        const DevMat &ah = *find_tensor(e, fl, "ah");
        logits = ah.p;
        ldl = ah.ld;
    } else {  // quirk Q9: "az" (E x 1) reinterpreted as a dense V x C array
        if (e->Ein < (uint64_t)e->V * lab->cols)
            return fail(e, DORY_EINVAL, "predictGAT quirk mode reads az as V x C but E_in (%llu) < V*C; "
                        "the reference reads out of bounds here -- use DORY_FLAG_GAT_PREDICT_AH",
                        (unsigned long long)e->Ein);
        logits = find_tensor(e, fl, "az")->p;
        ldl = lab->cols;
    }
    LAUNCHED(launch_gat_predict(logits, ldl, lab->p, grad.p, lab->ld, lab->cols, c->lowBound, c->upBound, e->stream));
    return DORY_OK;
}

int apply_update_impl(dory_engine *e, uint32_t layer) {
    WeightSet &W = e->W[layer];
    int jrc = join_comm(e);  // no collective of the first communicator while the exchange stream is busy
    if (jrc) return jrc;
    if (e->comm && e->cfg.num_nodes > 1) {  // weighttensor.cpp:263-267: local + ghost updates summed
        std::string msg = e->comm->allreduce_sum(W.dw.as<float>(), W.floats(), e->stream);
        if (!msg.empty()) return fail(e, DORY_ECOMM, "%s", msg.c_str());
        e->ghost_reads_pending.clear();
    }
    if (e->cfg.gnn_type == DORY_GAT) return DORY_OK;  // tryApplyUpdateFake, weightserver.cpp:112-116 (Q10)
    LAUNCHED(launch_adam(W.w.as<float>(), W.dw.as<float>(), W.m.as<float>(), W.v.as<float>(), W.floats(),
                         e->lr_t, e->beta1, e->beta2, e->eps, e->stream));
    if (layer == 0) adam_next_iteration(e);  // AdamOptimizer.cpp:49-50
    return DORY_OK;
}

__global__ void publish_stats_kernel(const float *__restrict__ dev, volatile float *host) {
    if (threadIdx.x < 2) host[threadIdx.x] = dev[threadIdx.x];
    __threadfence_system();
}

int fetch_stats(dory_engine *e) {
    float hs[2] = {0.f, 0.f};
    CU(cudaMemcpyAsync(hs, e->stats_dev.p, sizeof hs, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->stats.acc_sum = hs[0];
    e->stats.loss_sum = hs[1];
    return DORY_OK;
}

}  // namespace

// =============================================================================== C ABI
extern "C" {

int dory_abi_version(void) { return DORY_ABI_VERSION; }

const char *dory_last_error(const dory_engine *e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int dory_create(dory_engine **out, const dory_config *cfg) {
    dory_engine *e = nullptr;  // for fail()
    if (!out || !cfg) return fail(e, DORY_EINVAL, "null argument");
    *out = nullptr;
    if (cfg->abi_version != DORY_ABI_VERSION) return fail(e, DORY_EINVAL, "ABI version %u != %u", cfg->abi_version, DORY_ABI_VERSION);
    if (cfg->n_layers < 1 || cfg->n_layers > DORY_MAX_LAYERS) return fail(e, DORY_EINVAL, "n_layers %u out of range", cfg->n_layers);
    if (cfg->gnn_type == DORY_GCN && cfg->n_layers < 2)
        return fail(e, DORY_EINVAL, "GCN needs >= 2 layers (the reference's last layer writes grad[layer], which only exists for layer > 0)");
    if (cfg->gnn_type != DORY_GCN && cfg->gnn_type != DORY_GAT) return fail(e, DORY_EINVAL, "unknown gnn_type %u", cfg->gnn_type);
    for (uint32_t i = 0; i <= cfg->n_layers; ++i)
        if (cfg->dims[i] == 0) return fail(e, DORY_EINVAL, "layer width %u is zero", i);
    if (cfg->dims[cfg->n_layers] > 256) return fail(e, DORY_EINVAL, "more than 256 classes not supported");
    if (cfg->num_nodes == 0 || cfg->node_id >= cfg->num_nodes) return fail(e, DORY_EINVAL, "node_id/num_nodes invalid");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(e, DORY_ENODEV, "no CUDA device visible; dorylus_b200 has no CPU fallback");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(e, DORY_EINVAL, "device %d out of range (%d visible)", cfg->device, ndev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return fail(e, DORY_ECUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10)
        return fail(e, DORY_ENODEV, "device %d is sm_%d%d; this library only carries sm_100a code", cfg->device, prop.major, prop.minor);
    if (cudaSetDevice(cfg->device) != cudaSuccess) return fail(e, DORY_ECUDA, "cudaSetDevice failed");
    std::unique_ptr<dory_engine> eng(new dory_engine());
    eng->cfg = *cfg;
    if (cudaStreamCreateWithFlags(&eng->stream, cudaStreamNonBlocking) != cudaSuccess)
        return fail(e, DORY_ECUDA, "cudaStreamCreate failed");
    for (auto &ev : eng->events)
        if (cudaEventCreate(&ev) != cudaSuccess) return fail(e, DORY_ECUDA, "cudaEventCreate failed");
    eng->adam_epochs = 0;
    adam_next_iteration(eng.get());  // AdamOptimizer ctor, AdamOptimizer.cpp:6-8
    *out = eng.release();
    return DORY_OK;
}

void dory_destroy(dory_engine *e) {
    if (!e) return;
    cudaSetDevice(e->cfg.device);
    if (e->copy_stream) cudaStreamSynchronize(e->copy_stream);
    if (e->comm_stream) cudaStreamSynchronize(e->comm_stream);
    if (e->stream) cudaStreamSynchronize(e->stream);
    e->comm.reset();
    for (void *b : e->ipc_bases) cudaIpcCloseMemHandle(b);
    for (auto &kv : e->prefetch) {
        if (kv.second->staged) cudaEventDestroy(kv.second->staged);
        if (kv.second->consumed) cudaEventDestroy(kv.second->consumed);
    }
    for (auto &ev : e->events)
        if (ev) cudaEventDestroy(ev);
    for (auto &ev : e->stats_ready)
        if (ev) cudaEventDestroy(ev);
    if (e->stats_host) cudaFreeHost(e->stats_host);
    if (e->copy_stream) cudaStreamDestroy(e->copy_stream);
    if (e->comm_stream) cudaStreamDestroy(e->comm_stream);
    if (e->ev_compute) cudaEventDestroy(e->ev_compute);
    if (e->ev_comm) cudaEventDestroy(e->ev_comm);
    if (e->stream) cudaStreamDestroy(e->stream);
    delete e;
}

int dory_prefetch_tensor(dory_engine *e, uint32_t layer, const char *name, const float *host, uint64_t rows,
                         uint32_t cols) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!name || !host) return fail(e, DORY_EINVAL, "null argument");
    const DevMat *m = find_tensor(e, layer, name);
    if (!m) return fail(e, DORY_EINVAL, "no tensor '%s' at layer %u", name, layer);
    if (m->rows != rows || m->cols != cols)
        return fail(e, DORY_EINVAL, "tensor '%s'[%u] is %llu x %u, caller passed %llu x %u", name, layer,
                    (unsigned long long)m->rows, m->cols, (unsigned long long)rows, cols);
    if (rows == 0) return DORY_OK;
    if (!e->copy_stream) CU(cudaStreamCreateWithFlags(&e->copy_stream, cudaStreamNonBlocking));
    auto &slot = e->prefetch[{layer, std::string(name)}];
    if (!slot) {
        slot.reset(new dory_engine::Prefetch());
        slot->dst = *m;
        CU(slot->stage.alloc((size_t)rows * cols * 4));
        CU(cudaEventCreateWithFlags(&slot->staged, cudaEventDisableTiming));
        CU(cudaEventCreateWithFlags(&slot->consumed, cudaEventDisableTiming));
    }
    if (slot->pending) return fail(e, DORY_ESTATE, "tensor '%s'[%u] already has an uncommitted prefetch", name, layer);
    // the staging buffer may still be read by the previous commit's re-pitch kernel
    if (slot->used) CU(cudaStreamWaitEvent(e->copy_stream, slot->consumed, 0));
    CU(cudaMemcpyAsync(slot->stage.p, host, (size_t)rows * cols * 4, cudaMemcpyHostToDevice, e->copy_stream));
    CU(cudaEventRecord(slot->staged, e->copy_stream));
    slot->pending = true;
    return DORY_OK;
}

int dory_commit_prefetch(dory_engine *e) {
    int rc = check_loaded(e);
    if (rc) return rc;
    for (auto &kv : e->prefetch) {
        dory_engine::Prefetch &p = *kv.second;
        if (!p.pending) continue;
        CU(cudaStreamWaitEvent(e->stream, p.staged, 0));
        if (is_edge_vector(p.dst) || p.dst.ld == p.dst.cols) {
            CU(cudaMemcpyAsync(p.dst.p, p.stage.p, (size_t)p.dst.rows * p.dst.cols * 4, cudaMemcpyDeviceToDevice, e->stream));
        } else {
            LAUNCHED(launch_repitch(p.stage.as<float>(), p.dst.cols, p.dst.p, p.dst.ld, p.dst.rows, p.dst.cols, e->stream));
        }
        CU(cudaEventRecord(p.consumed, e->stream));
        p.pending = false;
        p.used = true;
    }
    return DORY_OK;
}

int dory_sync(dory_engine *e) {
    if (!e) return DORY_EINVAL;
    if (e->comm_pending) {
        int rc = join_comm(e);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(e->stream));
    return DORY_OK;
}

void dory_free(void *p) { std::free(p); }

int dory_set_option(dory_engine *e, const char *key, const char *value) {
    if (!e) return DORY_EINVAL;
    if (!key || !value) return fail(e, DORY_EINVAL, "null argument");
    char *end = nullptr;
    const long v = std::strtol(value, &end, 10);
    if (end == value || v < 0) return fail(e, DORY_EINVAL, "option '%s': bad value '%s'", key, value);
    if (std::strcmp(key, "spmm_lg") == 0) {
        if (v != 0 && v != 4 && v != 8 && v != 16 && v != 32) return fail(e, DORY_EINVAL, "spmm_lg must be 0, 4, 8, 16 or 32");
        e->spmm_lg = (int)v;
    } else if (std::strcmp(key, "spmm_vec") == 0) {
        if (v > 4 || v == 3) return fail(e, DORY_EINVAL, "spmm_vec must be 0, 1, 2 or 4");
        e->spmm_vec = (int)v;
    } else if (std::strcmp(key, "p2p") == 0) {
        e->p2p = v != 0;
    } else if (std::strcmp(key, "overlap") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "overlap must be set before dory_load_partition");
        e->overlap = v != 0;
    } else if (std::strcmp(key, "p2p_ctas") == 0) {
        e->p2p_ctas = (int)std::max<long>(1, std::min<long>(v, 8));
    } else if (std::strcmp(key, "p2p_rows") == 0) {
        e->p2p_variant = (int)v;
    } else if (std::strcmp(key, "p2p_elide_barrier") == 0) {
        e->elide_pre_barrier = v != 0;
    } else if (std::strcmp(key, "src_blocks") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "src_blocks must be set before dory_load_partition");
        e->src_blocks = (uint32_t)v;
    } else if (std::strcmp(key, "gat_windows") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "gat_windows must be set before dory_load_partition");
        e->gat_windows = v != 0;
    } else if (std::strcmp(key, "locality_block") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "locality_block must be set before dory_load_partition");
        e->locality_block = (uint32_t)v;
    } else if (std::strcmp(key, "hub_degree") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "hub_degree must be set before dory_load_partition");
        e->hub_degree = (uint32_t)v;
    } else if (std::strcmp(key, "row_order") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "row_order must be set before dory_load_partition");
        if (v > 2) return fail(e, DORY_EINVAL, "row_order must be 0 (auto), 1 (degree-descending) or 2 (degree classes)");
        e->row_order = (int)v;
    } else if (std::strcmp(key, "spmm_light") == 0) {
        if (v > 2) return fail(e, DORY_EINVAL, "spmm_light must be 0 (auto), 1 (warp per row) or 2 (lane group per row)");
        e->spmm_light = (int)v;
    } else if (std::strcmp(key, "spmm_occ") == 0) {
        e->spmm_occ = (int)v;
    } else if (std::strcmp(key, "tn_small") == 0) {
        e->tn_small = v != 0;
    } else if (std::strcmp(key, "fuse_softmax") == 0) {
        e->fuse_softmax = (int)v;
    } else if (std::strcmp(key, "tc_stages") == 0) {
        e->tc_stages = (int)v;
    } else if (std::strcmp(key, "fuse_tanh_bwd") == 0) {
        e->fuse_tanh_bwd = v != 0;
    } else if (std::strcmp(key, "tc_small") == 0) {
        e->tc_small = v != 0;
    } else if (std::strcmp(key, "tensor_cores") == 0) {
        e->tensor_cores = v != 0;
    } else if (std::strcmp(key, "spmm_unroll") == 0) {
        if (v > 2) return fail(e, DORY_EINVAL, "spmm_unroll must be 0 (default), 1 or 2");
        e->spmm_unroll = (int)v;
    } else if (std::strcmp(key, "apply_first_mask") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "apply_first_mask must be set before dory_load_partition");
        if (e->cfg.gnn_type != DORY_GCN) return fail(e, DORY_EINVAL, "apply_first_mask is a GCN option");
        if (v >= (1L << e->cfg.n_layers)) return fail(e, DORY_EINVAL, "apply_first_mask has bits beyond layer %u", e->cfg.n_layers - 1);
        e->af_mask = v;
    } else if (std::strncmp(key, "tile", 4) == 0) {
        if (e->loaded && std::strcmp(key, "tile_pipe") != 0) return fail(e, DORY_ESTATE, "%s must be set before dory_load_partition", key);
        if (std::strcmp(key, "tile") == 0) {
            if (v > 2) return fail(e, DORY_EINVAL, "tile must be 0 (off), 1 (on) or 2 (on when the plan covers enough edges)");
            e->tile_mode = (int)v;
        } else if (std::strcmp(key, "tile_rows") == 0) {
            e->tile_rows = (uint32_t)v;
        } else if (std::strcmp(key, "tile_window") == 0) {
            e->tile_window = (uint32_t)v;
        } else if (std::strcmp(key, "tile_smem_kb") == 0) {
            if (v < 8 || v > 200) return fail(e, DORY_EINVAL, "tile_smem_kb must be 8..200");
            e->tile_smem_kb = (uint32_t)v;
        } else if (std::strcmp(key, "tile_slab") == 0) {
            if (v != 0 && v != 32 && v != 64 && v != 96 && v != 128) return fail(e, DORY_EINVAL, "tile_slab must be 0, 32, 64, 96 or 128");
            e->tile_slab = (uint32_t)v;
        } else if (std::strcmp(key, "tile_min_coverage") == 0) {
            if (v > 100) return fail(e, DORY_EINVAL, "tile_min_coverage is a percentage");
            e->tile_min_coverage = (uint32_t)v;
        } else if (std::strcmp(key, "tile_edges") == 0) {
            if (v < 256 || v > 16384) return fail(e, DORY_EINVAL, "tile_edges must be 256..16384");
            e->tile_edges = (uint32_t)v;
        } else if (std::strcmp(key, "tile_pipe") == 0) {
            e->tile_pipe = v != 0;
        } else if (std::strcmp(key, "tile_team") == 0) {
            e->tile_team = (uint32_t)std::max<long>(v, 1);
        } else {
            return fail(e, DORY_EINVAL, "unknown option '%s'", key);
        }
    } else if (std::strcmp(key, "heavy_degree") == 0) {
        if (e->loaded) return fail(e, DORY_ESTATE, "heavy_degree must be set before dory_load_partition");
        e->heavy_degree = (uint32_t)std::max<long>(v, 1);
    } else {
        return fail(e, DORY_EINVAL, "unknown option '%s'", key);
    }
    return DORY_OK;
}

int dory_preprocess_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, const int32_t *parts,
                          uint32_t n_vertices, uint32_t part, uint32_t n_parts, int undirected, void **image,
                          size_t *image_len) {
    dory_engine *e = nullptr;
    if (!parts || !image || !image_len || (n_edges && (!src || !dst))) return fail(e, DORY_EINVAL, "null argument");
    EdgeList el;
    el.src = src; el.dst = dst; el.stride = 1; el.n = n_edges;
    HostImage img;
    std::string msg = preprocess_partition(el, parts, n_vertices, part, n_parts, undirected != 0, img);
    if (!msg.empty()) return fail(e, DORY_EINVAL, "%s", msg.c_str());
    *image_len = img.size();
    *image = img.release();
    return DORY_OK;
}

int dory_preprocess_incident_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, const int32_t *parts,
                                   uint32_t n_vertices, uint32_t part, uint32_t n_parts, const uint32_t *in_degree,
                                   uint64_t global_edges, void **image, size_t *image_len) {
    dory_engine *e = nullptr;
    if (!parts || !image || !image_len || !in_degree || (n_edges && (!src || !dst))) return fail(e, DORY_EINVAL, "null argument");
    EdgeList el;
    el.src = src; el.dst = dst; el.stride = 1; el.n = n_edges;
    HostImage img;
    std::string msg = preprocess_partition(el, parts, n_vertices, part, n_parts, false, img, in_degree, global_edges);
    if (!msg.empty()) return fail(e, DORY_EINVAL, "%s", msg.c_str());
    *image_len = img.size();
    *image = img.release();
    return DORY_OK;
}

int dory_preprocess_dir(const char *dir, uint32_t part, uint32_t n_parts, int undirected) {
    dory_engine *e = nullptr;
    if (!dir) return fail(e, DORY_EINVAL, "null argument");
    const std::string d(dir);
    // graph.bsnap.edges: BSHeaderType {int32, uint32, uint64} then (src, dst) pairs (dataloader.hpp:11-15)
    FILE *f = std::fopen((d + "graph.bsnap.edges").c_str(), "rb");
    if (!f) return fail(e, DORY_EFORMAT, "cannot open %sgraph.bsnap.edges", dir);
    struct { int32_t sz; uint32_t nv; uint64_t ne; } hdr;
    if (std::fread(&hdr, sizeof hdr, 1, f) != 1 || hdr.sz != 4) {
        std::fclose(f);
        return fail(e, DORY_EFORMAT, "bad bsnap header");
    }
    std::vector<uint32_t> pairs;
    {
        std::fseek(f, 0, SEEK_END);
        long end = std::ftell(f);
        std::fseek(f, sizeof hdr, SEEK_SET);
        size_t n = ((size_t)end - sizeof hdr) / 8;
        pairs.resize(2 * n);
        if (n && std::fread(pairs.data(), 8, n, f) != n) {
            std::fclose(f);
            return fail(e, DORY_EFORMAT, "short read on edge file");
        }
    }
    std::fclose(f);
    // graph.bsnap.parts: one id per line; lines not starting with a digit are skipped (dataloader.cpp:66-68)
    std::vector<int32_t> parts;
    {
        FILE *pf = std::fopen((d + "graph.bsnap.parts").c_str(), "r");
        if (!pf) return fail(e, DORY_EFORMAT, "cannot open %sgraph.bsnap.parts", dir);
        char line[256];
        while (std::fgets(line, sizeof line, pf)) {
            if (line[0] < '0' || line[0] > '9') continue;
            parts.push_back((int32_t)std::strtol(line, nullptr, 10));
        }
        std::fclose(pf);
    }
    EdgeList el;
    el.src = pairs.data(); el.dst = pairs.data() + 1; el.stride = 2; el.n = pairs.size() / 2;
    HostImage img;
    std::string msg = preprocess_partition(el, parts.data(), (uint32_t)parts.size(), part, n_parts, undirected != 0, img);
    if (!msg.empty()) return fail(e, DORY_EINVAL, "%s", msg.c_str());
    char name[64];
    std::snprintf(name, sizeof name, "graph.%u.bin", part);
    FILE *of = std::fopen((d + name).c_str(), "wb");
    if (!of) return fail(e, DORY_EFORMAT, "cannot write %s%s", dir, name);
    const bool ok = std::fwrite(img.data(), 1, img.size(), of) == img.size();
    std::fclose(of);
    return ok ? DORY_OK : fail(e, DORY_EFORMAT, "short write on %s%s", dir, name);
}

int dory_read_features(const char *dataset_dir, const char *features_file, const void *graph_bin, size_t len,
                       uint32_t node_id, uint32_t n_features, float *local_rows, float *ghost_rows) {
    if (!dataset_dir || !features_file || !graph_bin || !local_rows) {
        g_create_error = "dory_read_features: null argument";
        return DORY_EINVAL;
    }
    dory::PartitionView g;
    std::string m = dory::parse_partition(graph_bin, len, g);
    if (m.empty() && g.srcGhostCnt && !ghost_rows) m = "dory_read_features: partition has ghost rows but ghost_rows is null";
    if (m.empty()) m = dory::read_features(dataset_dir, features_file, g, node_id, n_features, local_rows, ghost_rows);
    if (m.empty()) return DORY_OK;
    g_create_error = m;
    return DORY_EFORMAT;
}

int dory_read_labels(const char *labels_file, const void *graph_bin, size_t len, uint32_t kinds, float *onehot) {
    if (!labels_file || !graph_bin || !onehot) {
        g_create_error = "dory_read_labels: null argument";
        return DORY_EINVAL;
    }
    dory::PartitionView g;
    std::string m = dory::parse_partition(graph_bin, len, g);
    if (m.empty()) m = dory::read_labels(labels_file, g, kinds, onehot);
    if (m.empty()) return DORY_OK;
    g_create_error = m;
    return DORY_EFORMAT;
}

int dory_partition_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, uint32_t n_vertices,
                         uint32_t n_parts, uint32_t passes, int32_t *parts, uint64_t *edge_cut) {
    const std::string m = dory::partition_edges(src, dst, n_edges, n_vertices, n_parts, passes ? passes : 8, parts, edge_cut);
    if (m.empty()) return DORY_OK;
    g_create_error = m;
    return DORY_EINVAL;
}

int dory_partition_file(const char *bsnap_path, uint32_t n_parts, const char *out_dir) {
    if (!bsnap_path || !n_parts) {
        g_create_error = "dory_partition_file: bad argument";
        return DORY_EINVAL;
    }
    const std::string m = dory::partition_file(bsnap_path, n_parts, out_dir, 8);
    if (m.empty()) return DORY_OK;
    g_create_error = m;
    return DORY_EFORMAT;
}

int dory_load_partition(dory_engine *e, const void *graph_bin, size_t len) {
    if (!e || !graph_bin) return DORY_EINVAL;
    if (e->loaded) return fail(e, DORY_ESTATE, "a partition is already loaded on this engine");
    CU(cudaSetDevice(e->cfg.device));
    PartitionView pv;
    std::string msg = parse_partition(graph_bin, len, pv);
    if (!msg.empty()) return fail(e, DORY_EFORMAT, "%s", msg.c_str());
    if (pv.numNodes != e->cfg.num_nodes)
        return fail(e, DORY_EINVAL, "graph image was built for %u partitions, engine configured for %u", pv.numNodes, e->cfg.num_nodes);
    e->V = pv.localVtxCnt; e->gV = pv.globalVtxCnt; e->Gs = pv.srcGhostCnt; e->Gd = pv.dstGhostCnt;
    e->Ein = pv.localInEdgeCnt; e->Eout = pv.localOutEdgeCnt; e->Eglob = pv.globalEdgeCnt;
    if (e->V == 0) return fail(e, DORY_EINVAL, "partition has no local vertices");
    if (pv.fwdNnz != e->Ein || pv.bwdNnz != e->Eout) return fail(e, DORY_EFORMAT, "edge counts disagree with CSC/CSR nnz");
    e->l2g.resize(e->V);
    std::memcpy(e->l2g.data(), pv.localToGlobal, 4 * (size_t)e->V);
    for (int dir = 0; dir < 2; ++dir) {
        auto &lists = dir == 0 ? pv.fwdSend : pv.bwdSend;
        e->sendIds[dir].assign(pv.numNodes, {});
        for (uint32_t p = 0; p < pv.numNodes; ++p) {
            e->sendIds[dir][p].resize(lists[p].second);
            if (lists[p].second) std::memcpy(e->sendIds[dir][p].data(), lists[p].first, 4 * (size_t)lists[p].second);
            for (uint32_t id : e->sendIds[dir][p])
                if (id >= e->V) return fail(e, DORY_EFORMAT, "send list references vertex %u >= %u", id, e->V);
        }
    }
    e->af.assign(e->L(), 0);
    if (e->cfg.gnn_type == DORY_GCN)
        for (uint32_t l = 0; l < e->L(); ++l) {
            if (e->af_mask >= 0) e->af[l] = (e->af_mask >> l) & 1;
            else if (e->cfg.flags & DORY_FLAG_APPLY_FIRST) e->af[l] = padded_ld(e->dim(l + 1)) < padded_ld(e->dim(l));
        }
    int rc = upload_adjacency(e, e->fwd, pv.colPtrs, pv.rowIdxs, pv.fwdVals, pv.fwdNnz, e->V, e->V + e->Gs);
    if (rc) return rc;
    rc = upload_adjacency(e, e->bwd, pv.rowPtrs, pv.colIdxs, pv.bwdVals, pv.bwdNnz, e->V, e->V + e->Gd);
    if (rc) return rc;
    CU(e->norms.alloc(4 * (size_t)e->V));
    CU(cudaMemcpyAsync(e->norms.p, pv.norms, 4 * (size_t)e->V, cudaMemcpyHostToDevice, e->stream));

    rc = e->cfg.gnn_type == DORY_GCN ? preallocate_gcn(e) : preallocate_gat(e);
    if (rc) return rc;
    const uint32_t L = e->L();
    e->W.resize(L);
    uint32_t maxld = 0;
    for (uint32_t l = 0; l <= L; ++l) maxld = std::max(maxld, padded_ld(e->dim(l)));
    for (uint32_t l = 0; l < L; ++l) {
        rc = alloc_weight(e, e->W[l], e->dim(l), e->dim(l + 1));
        if (rc) return rc;
    }
    if (e->cfg.gnn_type == DORY_GAT) {
        e->Ai.resize(L);
        for (uint32_t l = 0; l < L; ++l) {
            rc = alloc_weight(e, e->Ai[l], e->dim(l + 1), 1);
            if (rc) return rc;
        }
    }
    const size_t scr = (size_t)e->V * maxld * sizeof(float);
    CU(e->scratchA.alloc(scr));
    CU(e->scratchB.alloc(scr));
    CU(cudaMemsetAsync(e->scratchA.p, 0, scr, e->stream));
    CU(cudaMemsetAsync(e->scratchB.p, 0, scr, e->stream));
    size_t wmax = 0;
    for (auto &w : e->W) wmax = std::max(wmax, w.floats());
    size_t ws_floats = std::max<size_t>(wmax * 64, (size_t)maxld * maxld * 64);
    // narrow layers (input width <= 64): gemm_tn_small_kernel splits the vertices 1184 ways (8 CTAs per SM)
    ws_floats = std::max<size_t>(ws_floats, (size_t)1184 * 64 * std::min<uint32_t>(maxld, 128));
    if (e->cfg.gnn_type == DORY_GAT)  // launch_gat_edge_backward: one partial row per 512 vertices + z^T z + its split-K
        ws_floats = std::max<size_t>(ws_floats, ((size_t)e->V / 512 + 1) * maxld + (size_t)maxld * maxld * 65);
    CU(e->gemm_ws.alloc(ws_floats * sizeof(float)));
    CU(e->rowstat.alloc(2 * (size_t)e->V * sizeof(float)));
    CU(e->stats_dev.alloc(2 * sizeof(float)));
    CU(cudaMemsetAsync(e->stats_dev.p, 0, 2 * sizeof(float), e->stream));
    CU(cudaStreamSynchronize(e->stream));
    e->loaded = true;
    return DORY_OK;
}

int dory_graph_counts(const dory_engine *e, uint64_t out[7]) {
    if (!e || !out || !e->loaded) return DORY_EINVAL;
    out[0] = e->V; out[1] = e->gV; out[2] = e->Gs; out[3] = e->Gd; out[4] = e->Ein; out[5] = e->Eout; out[6] = e->Eglob;
    return DORY_OK;
}

int dory_tensor_shape(const dory_engine *e, uint32_t layer, const char *name, uint64_t *rows, uint32_t *cols) {
    if (!e || !name || !e->loaded) return DORY_EINVAL;
    const DevMat *m = find_tensor(e, layer, name);
    if (!m) return DORY_EINVAL;
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    return DORY_OK;
}

int dory_tensor_device(const dory_engine *e, uint32_t layer, const char *name, void **dptr, uint64_t *rows,
                       uint32_t *cols, uint32_t *ld) {
    if (!e || !name || !e->loaded) return DORY_EINVAL;
    const DevMat *m = find_tensor(e, layer, name);
    if (!m) return DORY_EINVAL;
    if (dptr) *dptr = m->p;
    if (rows) *rows = m->rows;
    if (cols) *cols = m->cols;
    if (ld) *ld = m->ld;
    return DORY_OK;
}

int dory_set_tensor(dory_engine *e, uint32_t layer, const char *name, const float *host, uint64_t rows, uint32_t cols) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!name || !host) return fail(e, DORY_EINVAL, "null argument");
    const DevMat *m = find_tensor(e, layer, name);
    if (!m) return fail(e, DORY_EINVAL, "no tensor '%s' at layer %u", name, layer);
    if (m->rows != rows || m->cols != cols)
        return fail(e, DORY_EINVAL, "tensor '%s'[%u] is %llu x %u, caller passed %llu x %u", name, layer,
                    (unsigned long long)m->rows, m->cols, (unsigned long long)rows, cols);
    return copy_in(e, *m, host);
}

int dory_get_tensor(dory_engine *e, uint32_t layer, const char *name, float *host, uint64_t rows, uint32_t cols) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!name || !host) return fail(e, DORY_EINVAL, "null argument");
    const DevMat *m = find_tensor(e, layer, name);
    if (!m) return fail(e, DORY_EINVAL, "no tensor '%s' at layer %u", name, layer);
    if (m->rows != rows || m->cols != cols)
        return fail(e, DORY_EINVAL, "tensor '%s'[%u] is %llu x %u, caller passed %llu x %u", name, layer,
                    (unsigned long long)m->rows, m->cols, (unsigned long long)rows, cols);
    return copy_out(e, *m, host);
}

static int weight_ref(dory_engine *e, uint32_t layer, const char *name, WeightSet **w) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!name) return fail(e, DORY_EINVAL, "null argument");
    if (layer >= e->L()) return fail(e, DORY_EINVAL, "weight layer %u out of range", layer);
    if (std::strcmp(name, "w") == 0) *w = &e->W[layer];
    else if (std::strcmp(name, "a_i") == 0 && e->cfg.gnn_type == DORY_GAT) *w = &e->Ai[layer];
    else return fail(e, DORY_EINVAL, "unknown weight '%s'", name);
    return DORY_OK;
}

static int weight_copy(dory_engine *e, WeightSet *w, float *dev, float *host, uint32_t rows, uint32_t cols, bool to_dev) {
    if (!host) return fail(e, DORY_EINVAL, "null argument");
    if (rows != w->rows || cols != w->cols) return fail(e, DORY_EINVAL, "weight is %u x %u, caller passed %u x %u", w->rows, w->cols, rows, cols);
    if (to_dev)
        CU(cudaMemcpy2DAsync(dev, (size_t)w->ld * 4, host, (size_t)cols * 4, (size_t)cols * 4, rows, cudaMemcpyHostToDevice, e->stream));
    else
        CU(cudaMemcpy2DAsync(host, (size_t)cols * 4, dev, (size_t)w->ld * 4, (size_t)cols * 4, rows, cudaMemcpyDeviceToHost, e->stream));
    CU(cudaStreamSynchronize(e->stream));
    return DORY_OK;
}

int dory_set_weights(dory_engine *e, uint32_t layer, const char *name, const float *host, uint32_t rows, uint32_t cols) {
    WeightSet *w = nullptr;
    int rc = weight_ref(e, layer, name, &w);
    if (rc) return rc;
    return weight_copy(e, w, w->w.as<float>(), const_cast<float *>(host), rows, cols, true);
}

int dory_get_weights(dory_engine *e, uint32_t layer, const char *name, float *host, uint32_t rows, uint32_t cols) {
    WeightSet *w = nullptr;
    int rc = weight_ref(e, layer, name, &w);
    if (rc) return rc;
    return weight_copy(e, w, w->w.as<float>(), host, rows, cols, false);
}

int dory_get_weight_grad(dory_engine *e, uint32_t layer, const char *name, float *host, uint32_t rows, uint32_t cols) {
    WeightSet *w = nullptr;
    int rc = weight_ref(e, layer, name, &w);
    if (rc) return rc;
    return weight_copy(e, w, w->dw.as<float>(), host, rows, cols, false);
}

int dory_init_weights(dory_engine *e) {
    int rc = check_loaded(e);
    if (rc) return rc;
    // WeightServer::xavierInitializer (weightserver.cpp:567-585): U(-1,1) * sqrt(6/(d1+d2)) from a
    // std::default_random_engine seeded with 8888 for EVERY matrix; kaiming (:593-612) likewise with
    // N(0,1) * sqrt(2/d1).  libstdc++'s engine/distributions make this bit-reproducible.
    for (uint32_t l = 0; l < e->L(); ++l) {
        const uint32_t d1 = e->dim(l), d2 = e->dim(l + 1);
        std::vector<float> w((size_t)d1 * d2);
        {
            std::default_random_engine dre(8888);
            std::uniform_real_distribution<float> dist(-1, 1);
            for (auto &x : w) x = dist(dre);
            const float nf = std::sqrt(6.0 / (float(d1 + d2)));
            for (auto &x : w) x *= nf;
        }
        rc = dory_set_weights(e, l, "w", w.data(), d1, d2);
        if (rc) return rc;
        if (e->cfg.gnn_type == DORY_GAT) {
            std::vector<float> a(d2);
            std::default_random_engine dre(8888);
            std::normal_distribution<float> dist(0, 1);
            for (auto &x : a) x = dist(dre);
            const float nf = std::sqrt(2.0 / (float(d2)));
            for (auto &x : a) x *= nf;
            rc = dory_set_weights(e, l, "a_i", a.data(), d2, 1);
            if (rc) return rc;
        }
    }
    return DORY_OK;
}

int dory_apply_update(dory_engine *e, uint32_t layer) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (layer >= e->L()) return fail(e, DORY_EINVAL, "layer %u out of range", layer);
    return apply_update_impl(e, layer);
}

int dory_layer_schedule(const dory_engine *e, uint32_t layer, int *apply_first) {
    if (!e || !apply_first || !e->loaded || layer >= e->L()) return DORY_EINVAL;
    *apply_first = e->apply_first(layer) ? 1 : 0;
    return DORY_OK;
}

int dory_inc_layer(const dory_engine *e, dory_chunk *c) {
    if (!e || !c) return DORY_EINVAL;
    if (e->cfg.gnn_type == DORY_GCN) inc_layer_gcn(e, c); else inc_layer_gat(e, c);
    return DORY_OK;
}

int dory_aggregate(dory_engine *e, const dory_chunk *c) {
    int rc = check_loaded(e, e && e->cfg.gnn_type != DORY_GCN);  // GCN: run_gcn_spmm joins where it has to
    if (rc) return rc;
    if (!c) return fail(e, DORY_EINVAL, "null chunk");
    return e->cfg.gnn_type == DORY_GCN ? aggregate_gcn(e, c) : aggregate_gat(e, c);
}

int dory_apply_vertex(dory_engine *e, const dory_chunk *c) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!c) return fail(e, DORY_EINVAL, "null chunk");
    if (e->cfg.gnn_type == DORY_GCN) {
        if (c->dir == DORY_FORWARD) return vtx_forward_gcn(e, c->layer);
        if (c->layer >= e->L()) return fail(e, DORY_EINVAL, "apply_vertex: backward layer %u out of range", c->layer);
        if (e->apply_first(c->layer)) {  // this layer's own dense products come after ITS aggregation
            if ((rc = vtx_backward_apply_first(e, c->layer)) || c->layer == 0) return rc;
        }
        dory_chunk n = *c;  // applyVertexGCN, gcn_ops.cpp:198-200: inc layer first
        inc_layer_gcn(e, &n);
        if (n.dir != DORY_BACKWARD) return fail(e, DORY_EINVAL, "apply_vertex: backward chunk at layer 0 has nothing to apply");
        return vtx_backward_gcn(e, n.layer);
    }
    if (c->dir == DORY_FORWARD) return vtx_forward_gat(e, c->layer);
    dory_chunk n = *c;  // applyVertexGAT, gat_ops.cpp:271-273
    inc_layer_gat(e, &n);
    if (n.dir != DORY_BACKWARD) return fail(e, DORY_EINVAL, "apply_vertex: backward chunk at layer 0 has nothing to apply");
    return vtx_backward_gat(e, n.layer);
}

int dory_scatter(dory_engine *e, const dory_chunk *c) {
    int rc = check_loaded(e, false);  // exchange() orders itself against the exchange stream
    if (rc) return rc;
    if (!c) return fail(e, DORY_EINVAL, "null chunk");
    return e->cfg.gnn_type == DORY_GCN ? scatter_gcn(e, c) : scatter_gat(e, c);
}

int dory_apply_edge(dory_engine *e, const dory_chunk *c) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!c) return fail(e, DORY_EINVAL, "null chunk");
    if (e->cfg.gnn_type == DORY_GCN) return DORY_OK;  // applyEdgeGCN, gcn_ops.cpp:364-366
    if (c->layer == 0) return fail(e, DORY_EINVAL, "apply_edge: chunk layer 0 (NNCompute does layer--, CPU_comm.cpp:33)");
    return c->dir == DORY_FORWARD ? edge_forward_gat(e, c->layer - 1) : edge_backward_gat(e, c->layer - 1);
}

int dory_predict(dory_engine *e, const dory_chunk *c) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!c) return fail(e, DORY_EINVAL, "null chunk");
    if (e->cfg.gnn_type != DORY_GAT) return fail(e, DORY_EINVAL, "predict is a GAT operator");
    return predict_gat(e, c);
}

int dory_forward(dory_engine *e, uint32_t layer) {
    int rc = check_loaded(e);
    if (rc) return rc;
    dory_chunk c{0, e->cfg.node_id, 0, e->V, layer, DORY_FORWARD, 0, 1};
    if (e->cfg.gnn_type == DORY_GCN) {
        if (e->apply_first(layer)) {  // AV -> SC -> GA (+ activation)
            if ((rc = dory_apply_vertex(e, &c))) return rc;
            if ((rc = dory_scatter(e, &c))) return rc;
            if ((rc = dory_aggregate(e, &c))) return rc;
        } else {  // GA -> AV
            if ((rc = dory_aggregate(e, &c))) return rc;
            if ((rc = dory_apply_vertex(e, &c))) return rc;
        }
        inc_layer_gcn(e, &c);
        // -> SC -> AE for the next chunk; an apply-first layer ships its own t after its dense product
        if (c.dir == DORY_FORWARD && e->apply_first(c.layer)) return DORY_OK;
        if ((rc = dory_scatter(e, &c))) return rc;
        return dory_apply_edge(e, &c);
    }
    // GAT: AV -> SC -> AE -> GA (-> predict on the last layer)   
    if ((rc = dory_apply_vertex(e, &c))) return rc;
    inc_layer_gat(e, &c);
    if ((rc = dory_scatter(e, &c))) return rc;
    c.vertex = 0;
    if ((rc = dory_apply_edge(e, &c))) return rc;
    if ((rc = dory_aggregate(e, &c))) return rc;
    if (c.layer == e->L()) rc = dory_predict(e, &c);
    return rc;
}

int dory_backward(dory_engine *e, uint32_t layer) {
    int rc = check_loaded(e);
    if (rc) return rc;
    dory_chunk c{0, e->cfg.node_id, 0, e->V, layer, DORY_BACKWARD, 0, 1};
    if (e->cfg.gnn_type == DORY_GCN) {  // GA -> AV(B) -> SC -> AE for the next backward layer
        if ((rc = dory_aggregate(e, &c))) return rc;
        if ((rc = dory_apply_vertex(e, &c))) return rc;
        if (layer == 0) return DORY_OK;  // only reached when layer 0 is apply-first (dW[0] is done)
        inc_layer_gcn(e, &c);
        // vtxNNBackward(0) ended the epoch -- unless layer 0 is apply-first and still owes dW[0]
        if (c.layer == 0 && !e->apply_first(0)) return DORY_OK;
        if ((rc = dory_scatter(e, &c))) return rc;
        return dory_apply_edge(e, &c);
    }
    // GAT backward at chunk.layer = layer (feature layer layer-1): SC -> AE -> GA -> AV
    if ((rc = dory_scatter(e, &c))) return rc;
    c.vertex = 0;
    if ((rc = dory_apply_edge(e, &c))) return rc;
    if ((rc = dory_aggregate(e, &c))) return rc;
    c.vertex = 1;
    return dory_apply_vertex(e, &c);
}

int dory_epoch(dory_engine *e, dory_stats *stats) {
    int rc = check_loaded(e);
    if (rc) return rc;
    const uint32_t L = e->L();
    if (e->cfg.gnn_type == DORY_GCN) {
        for (uint32_t l = 0; l < L; ++l)
            if ((rc = dory_forward(e, l))) return rc;
        for (uint32_t l = L - 1; l > 0; --l)
            if ((rc = dory_backward(e, l))) return rc;
        if (e->apply_first(0) && (rc = dory_backward(e, 0))) return rc;
    } else {
        for (uint32_t l = 0; l < L; ++l)
            if ((rc = dory_forward(e, l))) return rc;
        for (uint32_t l = L; l > 0; --l)
            if ((rc = dory_backward(e, l))) return rc;
    }
    // weight updates in the order the weight server receives them: last layer first
    for (uint32_t l = L; l-- > 0;)
        if ((rc = apply_update_impl(e, l))) return rc;
    e->stats.epochs_done++;
    if (stats) return dory_get_stats(e, stats);
    return DORY_OK;
}

int dory_get_stats(dory_engine *e, dory_stats *stats) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!stats) return fail(e, DORY_EINVAL, "null argument");
    if ((rc = fetch_stats(e))) return rc;
    *stats = e->stats;
    return DORY_OK;
}

int dory_stats_enqueue(dory_engine *e, uint32_t slot) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (slot >= dory_engine::kStatSlots) return fail(e, DORY_EINVAL, "stats slot %u out of range", slot);
    if (e->stats_pending[slot]) return fail(e, DORY_ESTATE, "stats slot %u has an uncollected read-back", slot);
    if (!e->stats_host) CU(cudaHostAlloc(reinterpret_cast<void **>(&e->stats_host), sizeof(float) * 2 * dory_engine::kStatSlots, cudaHostAllocMapped));
    if (!e->stats_ready[slot]) CU(cudaEventCreateWithFlags(&e->stats_ready[slot], cudaEventDisableTiming));
    // Written by a one-warp kernel straight into the mapped pinned slot, not by a copy engine: an
    // 8-byte cudaMemcpyAsync queued behind the 0.6 GB input DMA of the next step on the same engine
    // and cost 2-8 ms per step.
    publish_stats_kernel<<<1, 32, 0, e->stream>>>(e->stats_dev.as<float>(), e->stats_host + 2 * slot);
    if (cudaGetLastError() != cudaSuccess) return fail(e, DORY_ECUDA, "stats publish kernel launch failed");
    e->stats.kernel_launches++;
    CU(cudaEventRecord(e->stats_ready[slot], e->stream));
    e->stats_snap[slot] = e->stats;  // host-side counters as of this point of the stream
    e->stats_pending[slot] = true;
    return DORY_OK;
}

int dory_stats_collect(dory_engine *e, uint32_t slot, dory_stats *stats) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!stats) return fail(e, DORY_EINVAL, "null argument");
    if (slot >= dory_engine::kStatSlots || !e->stats_pending[slot])
        return fail(e, DORY_ESTATE, "stats slot %u has no read-back in flight", slot);
    CU(cudaEventSynchronize(e->stats_ready[slot]));
    *stats = e->stats_snap[slot];
    stats->acc_sum = e->stats_host[2 * slot];
    stats->loss_sum = e->stats_host[2 * slot + 1];
    e->stats_pending[slot] = false;
    return DORY_OK;
}

int dory_comm_unique_id(void *id128) {
    dory_engine *e = nullptr;
    if (!id128) return fail(e, DORY_EINVAL, "null argument");
    std::string msg = dory::Comm::unique_id(id128);
    if (!msg.empty()) return fail(e, DORY_ECOMM, "%s", msg.c_str());
    return DORY_OK;
}

int dory_comm_init(dory_engine *e, const void *id128) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!id128) return fail(e, DORY_EINVAL, "null argument");
    if (e->comm) return fail(e, DORY_ESTATE, "communicator already initialised");
    auto comm = std::make_unique<dory::Comm>();
    std::string msg = comm->init(id128, (int)e->cfg.node_id, (int)e->cfg.num_nodes, e->cfg.device);
    if (!msg.empty()) return fail(e, DORY_ECOMM, "%s", msg.c_str());
    uint32_t maxld = 0;
    for (uint32_t l = 0; l <= e->L(); ++l) maxld = std::max(maxld, padded_ld(e->dim(l)));
    for (int dir = 0; dir < 2; ++dir) {
        msg = comm->set_send_lists(dir, e->sendIds[dir], maxld, e->stream);
        if (!msg.empty()) return fail(e, DORY_ECOMM, "%s", msg.c_str());
    }
    e->comm = std::move(comm);
    return DORY_OK;
}

int dory_comm_set_recv_slots(dory_engine *e, uint32_t dir, uint32_t peer, const uint32_t *slots, uint32_t n) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!e->comm) return fail(e, DORY_ESTATE, "dory_comm_init first");
    if (dir > 1 || peer >= e->cfg.num_nodes || (n && !slots)) return fail(e, DORY_EINVAL, "bad argument");
    const uint32_t G = dir == 0 ? e->Gs : e->Gd;
    for (uint32_t i = 0; i < n; ++i)
        if (slots[i] >= G) return fail(e, DORY_EINVAL, "recv slot %u >= ghost count %u", slots[i], G);
    uint32_t maxld = 0;
    for (uint32_t l = 0; l <= e->L(); ++l) maxld = std::max(maxld, padded_ld(e->dim(l)));
    std::string msg = e->comm->set_recv_slots((int)dir, (int)peer, slots, n, maxld, e->stream);
    if (!msg.empty()) return fail(e, DORY_ECOMM, "%s", msg.c_str());
    return DORY_OK;
}

int dory_ghost_slots(const void *my_bin, size_t my_len, uint32_t my_id, const void *peer_bin, size_t peer_len,
                     uint32_t dir, uint32_t *slots, uint32_t *n) {
    dory_engine *e = nullptr;
    if (!my_bin || !peer_bin || !n || dir > 1) return fail(e, DORY_EINVAL, "dory_ghost_slots: bad argument");
    PartitionView mine, peer;
    std::string msg = parse_partition(my_bin, my_len, mine);
    if (msg.empty()) msg = parse_partition(peer_bin, peer_len, peer);
    if (!msg.empty()) return fail(e, DORY_EFORMAT, "%s", msg.c_str());
    if (mine.numNodes != peer.numNodes || my_id >= peer.numNodes)
        return fail(e, DORY_EINVAL, "dory_ghost_slots: images are of %u and %u partitions, my_id %u", mine.numNodes, peer.numNodes, my_id);
    const auto &list = dir == 0 ? peer.fwdSend[my_id] : peer.bwdSend[my_id];
    *n = list.second;
    if (!slots) return DORY_OK;
    // my ghosts: (gvid, lvid) pairs; lvid = localVtxCnt + slot (graph/dataloader.cpp:311-322)
    const uint8_t *pairs = dir == 0 ? mine.srcGhostPairs : mine.dstGhostPairs;
    const uint32_t G = dir == 0 ? mine.srcGhostCnt : mine.dstGhostCnt;
    std::vector<std::pair<uint32_t, uint32_t>> ghosts(G);
    for (uint32_t k = 0; k < G; ++k) {
        std::memcpy(&ghosts[k].first, pairs + 8 * (size_t)k, 4);
        std::memcpy(&ghosts[k].second, pairs + 8 * (size_t)k + 4, 4);
        if (ghosts[k].second < mine.localVtxCnt || ghosts[k].second - mine.localVtxCnt >= G)
            return fail(e, DORY_EFORMAT, "dory_ghost_slots: ghost vertex id out of range");
    }
    std::sort(ghosts.begin(), ghosts.end());
    for (uint32_t i = 0; i < list.second; ++i) {
        uint32_t lvid, gvid;
        std::memcpy(&lvid, list.first + 4 * (size_t)i, 4);
        if (lvid >= peer.localVtxCnt) return fail(e, DORY_EFORMAT, "dory_ghost_slots: send list references vertex %u >= %u", lvid, peer.localVtxCnt);
        std::memcpy(&gvid, peer.localToGlobal + 4 * (size_t)lvid, 4);
        auto it = std::lower_bound(ghosts.begin(), ghosts.end(), std::make_pair(gvid, 0u));
        if (it == ghosts.end() || it->first != gvid)
            return fail(e, DORY_EINVAL, "dory_ghost_slots: peer sends vertex %u, which is not a ghost of this partition", gvid);
        slots[i] = it->second - mine.localVtxCnt;
    }
    return DORY_OK;
}

int dory_comm_set_send_slots(dory_engine *e, uint32_t dir, uint32_t peer, const uint32_t *slots, uint32_t n) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!e->comm) return fail(e, DORY_ESTATE, "dory_comm_init first");
    if (dir > 1 || peer >= e->cfg.num_nodes || peer == e->cfg.node_id || (n && !slots)) return fail(e, DORY_EINVAL, "bad argument");
    std::string msg = e->comm->set_send_slots((int)dir, (int)peer, slots, n);
    if (!msg.empty()) return fail(e, DORY_EINVAL, "%s", msg.c_str());
    e->p2p_validated.clear();  // checked against the peers' exported blocks before the next peer-memory exchange
    return DORY_OK;
}

namespace {
struct IpcBlob {
    cudaIpcMemHandle_t handle;
    uint64_t ghost_offset_bytes;
    uint32_t ghost_rows;
    uint32_t ld;  // row pitch of the exporter's block, floats
};
static_assert(sizeof(IpcBlob) <= DORY_IPC_BLOB_BYTES, "IPC blob does not fit DORY_IPC_BLOB_BYTES");

bool is_ghost_name(const char *n) {
    return !std::strcmp(n, "fg") || !std::strcmp(n, "bg") || !std::strcmp(n, "fg_z") || !std::strcmp(n, "bg_d") ||
           !std::strcmp(n, "fg_t") || !std::strcmp(n, "bg_g");
}
}  // namespace

int dory_comm_ipc_export(dory_engine *e, uint32_t layer, const char *ghost_name, void *blob80) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!ghost_name || !blob80 || !is_ghost_name(ghost_name)) return fail(e, DORY_EINVAL, "not a ghost tensor name");
    const DevMat *m = find_tensor(e, layer, ghost_name);
    if (!m) return fail(e, DORY_EINVAL, "no tensor '%s' at layer %u", ghost_name, layer);
    // ghost rows follow the V local rows inside one allocation
    IpcBlob b{};
    b.ghost_offset_bytes = (uint64_t)e->V * m->ld * 4;
    b.ghost_rows = (uint32_t)m->rows;
    b.ld = m->ld;
    void *base = reinterpret_cast<uint8_t *>(m->p) - b.ghost_offset_bytes;
    CU(cudaIpcGetMemHandle(&b.handle, base));
    std::memset(blob80, 0, DORY_IPC_BLOB_BYTES);
    std::memcpy(blob80, &b, sizeof b);
    return DORY_OK;
}

int dory_comm_ipc_import(dory_engine *e, uint32_t layer, const char *ghost_name, uint32_t peer, const void *blob80) {
    int rc = check_loaded(e);
    if (rc) return rc;
    if (!ghost_name || !blob80 || !is_ghost_name(ghost_name)) return fail(e, DORY_EINVAL, "not a ghost tensor name");
    if (peer >= e->cfg.num_nodes || peer == e->cfg.node_id) return fail(e, DORY_EINVAL, "bad peer");
    const DevMat *m = find_tensor(e, layer, ghost_name);
    if (!m) return fail(e, DORY_EINVAL, "no tensor '%s' at layer %u", ghost_name, layer);
    IpcBlob b;
    std::memcpy(&b, blob80, sizeof b);
    if (b.ld != m->ld)
        return fail(e, DORY_EINVAL, "peer %u exports '%s'[%u] with a row pitch of %u floats, ours is %u: the two engines were "
                    "configured with different layer widths or schedules", peer, ghost_name, layer, b.ld, m->ld);
    void *base = nullptr;
    CU(cudaIpcOpenMemHandle(&base, b.handle, cudaIpcMemLazyEnablePeerAccess));
    e->ipc_bases.push_back(base);
    auto &v = e->peer_ghost[m->p];
    if (v.size() != e->cfg.num_nodes) v.assign(e->cfg.num_nodes, nullptr);
    v[peer] = reinterpret_cast<float *>(reinterpret_cast<uint8_t *>(base) + b.ghost_offset_bytes);
    auto &rows = e->peer_ghost_rows[m->p];
    if (rows.size() != e->cfg.num_nodes) rows.assign(e->cfg.num_nodes, 0);
    rows[peer] = b.ghost_rows;
    e->p2p_validated.clear();
    return DORY_OK;
}

int dory_comm_send_gvids(const dory_engine *e, uint32_t dir, uint32_t peer, uint32_t *ids, uint32_t *n) {
    if (!e || !e->loaded || dir > 1 || peer >= e->cfg.num_nodes || !n) return DORY_EINVAL;
    const auto &l = e->sendIds[dir][peer];
    *n = (uint32_t)l.size();
    if (ids)
        for (size_t i = 0; i < l.size(); ++i) ids[i] = e->l2g[l[i]];
    return DORY_OK;
}

int dory_tile_info(const dory_engine *e, uint32_t dir, double *coverage, uint32_t *window_rows, uint32_t *tile_rows,
                   uint32_t *n_tiles) {
    if (!e || !e->loaded || dir > 1) return DORY_EINVAL;
    const Adjacency &adj = dir == 0 ? e->fwd : e->bwd;
    const TilePlanBuf *t = adj.tile.get();
    if (coverage) *coverage = t ? t->coverage : 0.0;
    if (window_rows) *window_rows = t ? t->window_rows : 0;
    if (tile_rows) *tile_rows = t ? t->tile_rows : 0;
    if (n_tiles) *n_tiles = t ? t->n_tiles : 0;
    return DORY_OK;
}

int dory_event_record(dory_engine *e, uint32_t slot) {
    if (!e || slot >= kNumEvents) return DORY_EINVAL;
    if (e->comm_pending) {  // a timing mark covers the exchange that was started before it
        int rc = join_comm(e);
        if (rc) return rc;
    }
    CU(cudaEventRecord(e->events[slot], e->stream));
    return DORY_OK;
}

int dory_event_elapsed_ms(dory_engine *e, uint32_t a, uint32_t b, float *ms) {
    if (!e || a >= kNumEvents || b >= kNumEvents || !ms) return DORY_EINVAL;
    CU(cudaEventSynchronize(e->events[b]));
    CU(cudaEventElapsedTime(ms, e->events[a], e->events[b]));
    return DORY_OK;
}

int dory_flush_l2(dory_engine *e, size_t bytes) {
    if (!e) return DORY_EINVAL;
    if (e->flush.bytes < bytes) CU(e->flush.alloc(bytes));
    LAUNCHED(launch_fill(e->flush.as<float>(), bytes / 4, 0.f, e->stream));
    return DORY_OK;
}

int dory_measure_fma_peak(dory_engine *e, float *tflops) {
    if (!e || !tflops) return DORY_EINVAL;
    if (e->flush.bytes < 256) CU(e->flush.alloc(256));
    int dev = 0, sms = 148;
    CU(cudaGetDevice(&dev));
    CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const unsigned blocks = (unsigned)sms * 8;  // 8 CTAs of 256 threads per SM: all 64 warp slots
    const int iters = 1 << 16;
    cudaEvent_t a, b;
    CU(cudaEventCreate(&a));
    CU(cudaEventCreate(&b));
    LAUNCHED(launch_fma_peak(e->flush.as<float>(), 1 << 10, blocks, e->stream));  // warm-up
    float best = 0.f;
    for (int rep = 0; rep < 3; ++rep) {
        CU(cudaEventRecord(a, e->stream));
        LAUNCHED(launch_fma_peak(e->flush.as<float>(), iters, blocks, e->stream));
        CU(cudaEventRecord(b, e->stream));
        CU(cudaEventSynchronize(b));
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, a, b));
        const double flops = 2.0 * 8.0 * iters * 256.0 * blocks;
        best = std::max(best, (float)(flops / (ms * 1e-3) / 1e12));
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    *tflops = best;
    return DORY_OK;
}

}  // extern "C"
