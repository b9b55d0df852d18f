/* dorylus_b200 — C ABI of the B200-native GNN aggregation engine.
 *
 * This is the drop-in boundary for the ONE hot path of uclasystem/dorylus that this repo
 * implements: per-layer Gather (normalised-adjacency SpMM, CSC forward / CSR
 * backward) -> ApplyVertex (H.W + activation / softmax-CE) -> Scatter (ghost rows) -> ApplyEdge.
 * Plain pointers and sizes only; no C++ / torch types cross it.  Every entry point names the
 * reference interface it stands in for (paths relative to the reference repo root).
 *
 * Conventions
 *   - every call returns DORY_OK (0) or a negative DORY_E* code; nothing calls exit()/abort()
 *     (the reference asserts or exit(EXIT_FAILURE)s: engine/engine.cpp:141-163);
 *     dory_last_error() returns the message of the last failing call on that engine.
 *   - compute calls ENQUEUE work on the engine's CUDA stream and return; dory_sync(),
 *     dory_get_tensor() and dory_get_stats() synchronise.
 *   - host pointers are caller-owned and only touched during the call.
 *   - one caller thread per engine (the reference drives one chunk per partition in CPU/GPU mode,
 *     run/run-onnode:62-70); one engine per GPU, one process per GPU.
 *   - tensors are addressed by (layer, name) exactly like Engine::savedNNTensors[layer][name]
 *     (engine/engine.hpp:157-158): GCN "x fg ah z h lab grad bg aTg" (engine/ops/gcn_ops.cpp:27-93),
 *     GAT "h z az fg_z A ah grad dA aTg bg_d lab" (engine/ops/gat_ops.cpp:27-115).
 *     They are row-major fp32 on the host side of this ABI; in HBM rows are padded.
 */
#ifndef DORYLUS_B200_H
#define DORYLUS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DORY_ABI_VERSION 1
#define DORY_MAX_LAYERS 8

enum {
    DORY_OK = 0,
    DORY_EINVAL = -1,   /* bad argument / unknown tensor / shape mismatch */
    DORY_ESTATE = -2,   /* call out of order (no partition loaded, no comm, ...) */
    DORY_ECUDA = -3,    /* a CUDA runtime call or kernel failed */
    DORY_ENOMEM = -4,
    DORY_ECOMM = -5,    /* NCCL / peer-memory failure */
    DORY_EFORMAT = -6,  /* malformed graph.<id>.bin / dataset file */
    DORY_ENODEV = -7    /* no usable sm_100 device: there is NO CPU fallback */
};

/* PROP_TYPE and GNN, src/common/utils.hpp:46,48 */
enum { DORY_FORWARD = 0, DORY_BACKWARD = 1 };
enum { DORY_GCN = 0, DORY_GAT = 1 };

/* == struct Chunk, src/common/utils.hpp:64-74 (same fields, fixed-width types). */
typedef struct dory_chunk {
    uint32_t localId;
    uint32_t globalId;
    uint32_t lowBound;
    uint32_t upBound;
    uint32_t layer;
    uint32_t dir;    /* DORY_FORWARD | DORY_BACKWARD */
    uint32_t epoch;
    uint8_t vertex;  /* 1: vertex NN (AV), 0: edge NN (AE) */
} dory_chunk;

/* flags */
#define DORY_FLAG_STRICT_MASK 0x1u  /* maskout whole non-train ROWS instead of replicating quirk Q6
                                       (CPU_comm.cpp:464-471 copies (end-stt) floats, not rows) */
#define DORY_FLAG_GAT_PREDICT_AH 0x2u /* predictGAT reads "ah" instead of replicating quirk Q9
                                       (gat_ops.cpp:252 reads "az") */
#define DORY_FLAG_NO_TENSOR_CORES 0x4u /* force the fp32 SIMT GEMM path for H.W */
#define DORY_FLAG_APPLY_FIRST 0x8u    /* GCN: run every layer whose output rows are narrower than its
                                       input rows as  A_hat . (in . W)  instead of the reference's
                                       (A_hat . in) . W  (gcn_ops.cpp:130-191 then CPU_comm.cpp:98-159).
                                       Same z / h / weight gradients up to fp32 rounding, F_out-wide
                                       gathers instead of F_in-wide ones (Reddit layer 0: 128 instead of
                                       602 floats per edge).  "ah", "grad" and "bg" of such a layer are not
                                       materialised; see "apply-first schedule" below.  Off by default:
                                       the default path keeps the reference's operator order. */

/* What Engine::init gathers from its CLI + layer config file (engine/utils.cpp:313-479). */
typedef struct dory_config {
    uint32_t abi_version;                /* DORY_ABI_VERSION */
    uint32_t gnn_type;                   /* DORY_GCN | DORY_GAT            (--gnn) */
    uint32_t n_layers;                   /* numLayers = #widths - 1 */
    uint32_t dims[DORY_MAX_LAYERS + 1];  /* layerConfig: F0 ... C */
    uint32_t node_id;                    /* partition / rank id            (nodeId) */
    uint32_t num_nodes;                  /* number of partitions           (numNodes) */
    int32_t device;                      /* CUDA device ordinal */
    float learning_rate;                 /* weight-server argv, run/run-onnode:226 (0.01) */
    uint32_t flags;
} dory_config;

/* Per-epoch results the reference logs ("batch Acc/Loss", CPU_comm.cpp:112-116) and the
 * counters bench.py reports. */
typedef struct dory_stats {
    float acc_sum;            /* getTrainStat acc over this partition's validation slice */
    float loss_sum;           /* getTrainStat loss (sum, not mean) */
    uint32_t val_rows;        /* (unsigned)(V_p * VAL_PORTION) */
    uint32_t epochs_done;
    uint64_t kernel_launches; /* kernels of THIS library launched since create */
    uint64_t edges_aggregated;/* edges walked by dory_aggregate since create */
} dory_stats;

typedef struct dory_engine dory_engine;

/* ---- lifecycle ---------------------------------------------------------------------------
 * dory_create        == Engine::init up to (not including) graph loading, engine/engine.cpp:40-61
 * dory_destroy       == Engine::destroy, engine/engine.cpp:316 ff. */
int dory_create(dory_engine **out, const dory_config *cfg);
void dory_destroy(dory_engine *e);
const char *dory_last_error(const dory_engine *e); /* e may be NULL: error of a failed dory_create */
int dory_abi_version(void);
int dory_sync(dory_engine *e);
/* Tuning knobs (no reference counterpart; results never depend on them beyond fp32 summation order).
 *   "spmm_lg" / "spmm_vec"  lanes per gathered row and float4 per lane of the aggregation kernel
 *                           (0 = choose from the row width); a slab narrower than the row walks the
 *                           adjacency once per slab.
 *   "spmm_unroll"           gather instructions issued back to back per lane group (0 = default).
 *   "spmm_occ"              CTAs per SM the aggregation kernel is compiled for (4, 5, 6 or 8).
 *   "spmm_light"            rows below the heavy threshold: 1 = a warp per row, 2 = a lane group per
 *                           row (several rows per warp; rows up to 128 floats), 0 = choose by degree.
 *   "src_blocks"            source-row windows per aggregation: each pass gathers only from a
 *                           (V+G)/n-row window so that it stays L2-resident (0 = size from the L2
 *                           capacity, 1 = off; set before dory_load_partition; GCN only).
 *   "gat_windows"           1: the GAT aggregations walk the source windows too (the attention values
 *                           "A" and their gradients "dA" are constant along a destination row -- quirk
 *                           Q8, one-sided score -- so the regrouped edge ids can be used with value
 *                           arrays kept in the original edge order; a caller that overwrites "A" / "dA"
 *                           with values that vary inside a row must turn this off).  Default 1; set
 *                           before dory_load_partition.
 *   "heavy_degree"          rows with at least this many edges get a whole CTA (set before
 *                           dory_load_partition).
 *   "hub_degree"            rows with at least this many edges get a thread-block cluster of 8 CTAs
 *                           (partials combined through distributed shared memory); 0 = E_p / 2048,
 *                           at least 2048 (set before dory_load_partition).
 *   "overlap"               1: with several partitions a GCN peer-memory exchange runs on its own
 *                           stream and the aggregation that consumes its ghost block walks the edges from the
 *                           partition's own rows first, waiting for the exchange only before the edges from
 *                           ghost rows (the adjacency is regrouped into [own rows | ghost rows] windows at
 *                           load time).  0 (default; the split's second pass over the rows cost more than the
 *                           hidden exchange returned where it was measured, profiles/round2_overlap_n2.md): the
 *                           exchange stays on the compute stream.  Set before load.
 *   "p2p_rows"              rows per warp of the peer-memory store kernel (0 = default 4, 1, 2;
 *                           9 = one row per warp with a system fence, the first version).
 *   "p2p_elide_barrier"     1 (default): skip the barrier in front of a peer-memory exchange when a
 *                           collective already separates it from the last read of that ghost block.
 *   "row_order"             issue order of the remaining rows: 1 = degree-descending, 2 = power-of-two
 *                           degree classes in vertex-id order (keeps the numbering's locality),
 *                           0 = decide from the share of near-diagonal edges (set before load).
 *   "locality_block"        rows per block of the locality-preserving order (row_order 2): degree
 *                           classes are formed inside blocks of this many consecutive vertices so that
 *                           a block's source rows stay in L2 across its classes (0 = 24 MB worth of
 *                           the widest gathered slab; set before load).
 *   "tensor_cores"          1: run H.W on tcgen05 (3xTF32, fp32-level accuracy) when the shape
 *                           qualifies; 0: always the fp32 CUDA-core GEMM.
 *   "tile"                  shared-memory-staged aggregation (spmm_tile.cu) for graphs whose vertex numbering
 *                           has locality: destination rows are cut into tiles, each tile's best window of
 *                           consecutive source rows is staged in shared memory by bulk TMA copies and the
 *                           edges into it never touch L2.  0 off, 1 on whenever a plan can be built, 2 (default) on
 *                           for graphs of average degree >= 96 whose plan serves at least "tile_min_coverage" % of
 *                           the edges from shared memory (measured 14-17 % faster than the gather kernels on the
 *                           community-structured Reddit shape, slower on the low-degree shapes, which therefore
 *                           need an explicit 1; profiles/round2_tile_kernel.md).  GCN aggregations of
 *                           whole-partition chunks; set before load.
 *   "tile_rows" / "tile_window"   destination rows per tile / source rows per window (0 = choose: the smallest
 *                           power-of-two window that keeps 92 % of the best coverage, tiles of half a window).
 *   "tile_smem_kb"          shared memory a CTA may spend on its window (default 100: two CTAs per SM).
 *   "tile_slab"             high-degree graphs: column slab in floats (32, 64, 96, 128; 0 = per launch: 64 for rows
 *                           up to 128 floats, 96 for wider ones).
 *   "tile_edges"            low-degree graphs: edges per tile (default 4096); a tile's offsets, ids and weights are
 *                           staged in shared memory beside its window.
 *   "tile_pipe"             low-degree graphs: 1 (default) persistent CTAs whose producer warp stages tile k+1 while
 *                           the other warps walk tile k (two-stage TMA pipeline); 0: one tile per CTA.
 *   "tile_team"             high-degree graphs: rows with at least this many edges are walked by the whole CTA.
 *   "fuse_softmax"          1 (default): the last layer's logits product and its soft-max / statistics / maskout /
 *                           gradient scale run as one kernel when the classes fit one 64-wide tile (the logits
 *                           never go to HBM) -- on tcgen05 with a thread-per-row epilogue out of TMEM when the
 *                           shape qualifies (K a multiple of 32, class pitch 32 or 64), else on the fp32 kernel;
 *                           2: the fp32 kernel only; 0: GEMM, then softmax_ce_kernel.
 *   "tc_small"              1 (default): the small-tile tcgen05 kernel (several CTAs per SM) for that product, for
 *                           Z = A.W with K <= 128 and N = 32 / 64, and for grad = G.W^T; 0: the deep-ring tcgen05
 *                           kernel / fp32 kernels of round 1.
 *   "tc_stages"             shared-memory stages per CTA of the small-tile kernel (0 = choose: 1 up to two K
 *                           blocks, else 2).
 *   "tn_small"              1 (default): narrow-M fp32 kernel for dW of layers whose input width is <= 32.
 *   "fuse_tanh_bwd"         1 (default): layer-0 backward with an input width <= 32 applies tanh' inside that kernel's
 *                           operand load instead of in a pass of its own (layer 0 needs dW only).
 *   "apply_first_mask"      bit l = 1: layer l runs apply-first (overrides the width rule of
 *                           DORY_FLAG_APPLY_FIRST; GCN only; set before dory_load_partition). */
int dory_set_option(dory_engine *e, const char *key, const char *value);

/* ---- dataset preprocessing (host only, no GPU needed) --------------------------------------
 * dory_preprocess_edges == DataLoader::preprocess (graph/dataloader.cpp:225-330) + RawGraph::dump
 * (graph/graph.cpp:200-273) on an in-memory edge list: returns a malloc'ed byte image that is
 * byte-identical to the reference's graph.<part>.bin.  `parts[v]` is the owner of global vertex v
 * (graph.bsnap.parts).  Free the image with dory_free().
 * dory_preprocess_dir  == the same, reading <dir>graph.bsnap.edges / .parts and writing
 * <dir>graph.<part>.bin like the reference (engine/engine.cpp:62-71).  `dir` ends with '/'. */
int dory_preprocess_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges,
                          const int32_t *parts, uint32_t n_vertices, uint32_t part,
                          uint32_t n_parts, int undirected, void **image, size_t *image_len);
int dory_preprocess_dir(const char *dir, uint32_t part, uint32_t n_parts, int undirected);
/* dory_preprocess_incident_edges == dory_preprocess_edges for a caller that holds only the edge records
 * INCIDENT to partition `part` (either endpoint owned by it), in edge-file order -- a pre-split edge
 * file, or a generator that runs per rank so that no process ever holds a 1.8 G-edge list.  The
 * reference reads the whole graph.bsnap.edges on every node (dataloader.cpp:241-275) and re-reads it
 * once more for the raw in-degrees of its ghost vertices (findGhostDegrees, :192-218); here those come
 * in as `in_degree[g]` (in-degree of global vertex g over the WHOLE graph, self loops excluded) and
 * `global_edges` (records in the whole graph).  The image equals the one dory_preprocess_edges builds
 * from the whole list (tests/test_loader.py). */
int dory_preprocess_incident_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges,
                                   const int32_t *parts, uint32_t n_vertices, uint32_t part, uint32_t n_parts,
                                   const uint32_t *in_degree, uint64_t global_edges, void **image,
                                   size_t *image_len);
void dory_free(void *p);

/* ---- dataset inputs (host only) ----------------------------------------------------------------
 * dory_read_features == Engine::readFeaturesFile (engine/utils.cpp:486-552) for the partition whose
 * graph.<id>.bin image is passed: if <dataset_dir>feats<F0>.<node_id>.bin exists it is read (local rows,
 * then source-ghost rows -- the reference's per-partition cache); otherwise the global features file
 * (uint32 numFeatures, then one row per global vertex) is streamed, the partition's local and ghost
 * rows are picked out and the cache file is written.  local_rows: [V_p x F0], ghost_rows: [Gs_p x F0]
 * (may be NULL when the partition has no source ghosts) -- what dory_set_tensor(e, 0, "x" / "fg") takes.
 * dory_read_labels   == Engine::readLabelsFile (engine/utils.cpp:559-596): one-hot [V_p x kinds] for
 * dory_set_tensor(e, L-1, "lab").  `dataset_dir` ends with '/'.  Errors: dory_last_error(NULL). */
int dory_read_features(const char *dataset_dir, const char *features_file, const void *graph_bin, size_t len,
                       uint32_t node_id, uint32_t n_features, float *local_rows, float *ghost_rows);
int dory_read_labels(const char *labels_file, const void *graph_bin, size_t len, uint32_t kinds, float *onehot);

/* ---- partitioning (host only) ------------------------------------------------------------------
 * dory_partition_edges == inputs/partitioner.cpp:63-111 (symmetrise the edge list, k-way edge-cut
 * partition with unit vertex weights) with METIS_PartGraphKway replaced by a deterministic
 * restreaming greedy partitioner that balances vertices AND in-edges per partition
 * (csrc/partition.cpp).  parts[v] receives the owner of global vertex v; *edge_cut (may be NULL) the
 * number of edge records whose endpoints have different owners.  passes = 0 picks the default.
 * dory_partition_file  == the whole tool: reads <bsnap_path> (graph.bsnap), writes
 * <out_dir>/<basename>.parts (one id per line, partitioner.cpp:123-128, read back by
 * DataLoader::readPartsFile, graph/dataloader.cpp:53-87) and <basename>.comm ("Communication cost: N").
 * Errors of these two calls are reported through dory_last_error(NULL). */
int dory_partition_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, uint32_t n_vertices,
                         uint32_t n_parts, uint32_t passes, int32_t *parts, uint64_t *edge_cut);
int dory_partition_file(const char *bsnap_path, uint32_t n_parts, const char *out_dir);

/* ---- partition + tensors --------------------------------------------------------------------
 * dory_load_partition == Graph::init (graph/graph.cpp:7-115) + Engine::preallocateGCN/GAT
 * (gcn_ops.cpp:27-93, gat_ops.cpp:27-115): parses a graph.<id>.bin image, uploads CSC/CSR/norms/
 * send lists to HBM and allocates every named tensor of every layer. */
int dory_load_partition(dory_engine *e, const void *graph_bin, size_t len);

/* Counts as Graph exposes them (graph/graph.hpp:70-77):
 * out[0..6] = localVtxCnt, globalVtxCnt, srcGhostCnt, dstGhostCnt, localInEdgeCnt, localOutEdgeCnt,
 * globalEdgeCnt. */
int dory_graph_counts(const dory_engine *e, uint64_t out[7]);

/* savedNNTensors[layer][name] <- host (readFeaturesFile / readLabelsFile, engine/utils.cpp:486-596)
 * and -> host.  rows/cols must match the tensor's shape (dory_tensor_shape). */
int dory_set_tensor(dory_engine *e, uint32_t layer, const char *name, const float *host,
                    uint64_t rows, uint32_t cols);
int dory_get_tensor(dory_engine *e, uint32_t layer, const char *name, float *host, uint64_t rows,
                    uint32_t cols);
int dory_tensor_shape(const dory_engine *e, uint32_t layer, const char *name, uint64_t *rows,
                      uint32_t *cols);
/* Input pipeline.  The reference loads features once from disk (readFeaturesFile,
 * engine/utils.cpp:486-552); a caller that streams new inputs every step can overlap the
 * host->device DMA with the previous step's compute:
 *   dory_prefetch_tensor  starts the DMA on the engine's copy stream and returns.  `host` should be
 *                         pinned memory and must stay valid and unchanged until the matching
 *                         dory_commit_prefetch has returned and dory_sync() has been called.
 *   dory_commit_prefetch  makes every prefetched tensor visible to the operators enqueued after it
 *                         (stream-ordered; operators enqueued before it still see the old values). */
int dory_prefetch_tensor(dory_engine *e, uint32_t layer, const char *name, const float *host,
                         uint64_t rows, uint32_t cols);
int dory_commit_prefetch(dory_engine *e);
/* Zero-copy view for callers that already hold data in HBM: device pointer + leading dimension
 * (in floats) of the padded row-major storage. */
int dory_tensor_device(const dory_engine *e, uint32_t layer, const char *name, void **dptr,
                       uint64_t *rows, uint32_t *cols, uint32_t *ld);

/* ---- weights (stand-in for the weight server the CPU/GPU backends talk to) -------------------
 * dory_init_weights  == WeightServer::initWeightsMasterGCN/GAT (weightserver.cpp:515-559):
 *                       xavier "w" per layer, kaiming "a_i" for GAT, seed 8888.
 * dory_set/get_weights == MessageService::prefetchWeightsMatrix / getWeightMatrix / getaMatrix
 *                       (commmanager/message_service.cpp:188-222); name is "w" or "a_i".
 * dory_get_weight_grad == the matrix handed to MessageService::sendWeightUpdate / sendaUpdate
 *                       (message_service.cpp:148-162); name "w" or "a_i".
 * dory_apply_update   == WeightTensor::tryApplyUpdate sync branch + AdamOptimizer::update
 *                       (weighttensor.cpp:263-284, AdamOptimizer.cpp:36-51): sums dW over all
 *                       partitions (all-reduce when a communicator exists) and steps Adam. */
int dory_init_weights(dory_engine *e);
int dory_set_weights(dory_engine *e, uint32_t layer, const char *name, const float *host,
                     uint32_t rows, uint32_t cols);
int dory_get_weights(dory_engine *e, uint32_t layer, const char *name, float *host, uint32_t rows,
                     uint32_t cols);
int dory_get_weight_grad(dory_engine *e, uint32_t layer, const char *name, float *host,
                         uint32_t rows, uint32_t cols);
int dory_apply_update(dory_engine *e, uint32_t layer);

/* ---- the SAGA operators (engine/engine.hpp:84-94) -------------------------------------------
 * dory_aggregate    == Engine::aggregateGCN / aggregateGAT   (gcn_ops.cpp:130-191, gat_ops.cpp:173-243)
 *                      honours chunk->lowBound/upBound (destination row range).
 * dory_apply_vertex == Engine::applyVertexGCN/GAT -> ResourceComm::NNCompute(vertex=true)
 *                      -> CPUComm::vtxNNForward{GCN,GAT} / vtxNNBackward{GCN,GAT}   (CPU_comm.cpp:98-188).
 *                      Like CPUComm it always processes the whole partition.
 * dory_scatter      == Engine::scatterGCN/GAT + ghostReceiver* + the scatter barrier
 *                      (gcn_ops.cpp:204-362, gat_ops.cpp:277-435, ops/pipeline.cpp:256-342).
 *                      Extension (GCN): a FORWARD chunk at layer 0 ships the owned rows of "x" into the
 *                      peers' "fg"[0] blocks.  The reference has no such step -- every partition reads
 *                      its layer-0 ghost rows from the feature file (engine/utils.cpp:486-552); with it
 *                      a caller that streams inputs uploads each feature row over PCIe once per box
 *                      instead of once per partition that has it as a ghost.
 * dory_apply_edge   == Engine::applyEdgeGCN/GAT -> NNCompute(vertex=false)
 *                      -> CPUComm::edgNNForwardGAT/edgNNBackwardGAT (CPU_comm.cpp:190-242); GCN: no-op.
 * dory_predict      == Engine::predictGAT (gat_ops.cpp:247-265).
 * dory_inc_layer    == Engine::incLayerGCN / incLayerGAT (engine/utils.cpp:714-748), in place. */
int dory_aggregate(dory_engine *e, const dory_chunk *c);
int dory_apply_vertex(dory_engine *e, const dory_chunk *c);
int dory_scatter(dory_engine *e, const dory_chunk *c);
int dory_apply_edge(dory_engine *e, const dory_chunk *c);
int dory_predict(dory_engine *e, const dory_chunk *c);
int dory_inc_layer(const dory_engine *e, dory_chunk *c);

/* Apply-first schedule (DORY_FLAG_APPLY_FIRST / option "apply_first_mask"; GCN).  For a layer l that
 * runs apply-first the operators keep their names but the order inside dory_forward / dory_backward
 * is AV -> SC -> GA (the order the reference's GAT uses), on these tensors:
 *   forward   dory_apply_vertex  "t"[l] = in . W[l]          (in = "x" or "h"[l-1], local rows only)
 *             dory_scatter       "t"[l] -> peers' "fg_t"[l]  (chunk.layer = l: fills the ghost block the
 *                                                             aggregation of layer l reads)
 *             dory_aggregate     "z"[l] = A_hat ["t"; "fg_t"], then "h"[l] = tanh("z"[l]); last layer:
 *                                soft-max / statistics / maskout / scale on "z" -> "g"[l]
 *                                (whole-partition chunks only)
 *   backward  dory_scatter       "g"[l] -> peers' "bg_g"[l]  ("g"[l] = dL/dz[l], chunk.layer = l)
 *             dory_aggregate     "u"[l] = A_hat^T ["g"; "bg_g"]
 *             dory_apply_vertex  dW[l] = in^T . "u"[l];  l > 0: "aTg"[l-1] = "u"[l] . W[l]^T, followed by
 *                                the activation derivative of layer l-1 exactly as in the reference order
 *                                (into "g"[l-1] when that layer is apply-first too)
 * "aTg"[l-1] is dL/dh[l-1] under either schedule.  A BACKWARD chunk at layer 0 is valid when layer 0 is
 * apply-first (its weight gradient needs one aggregation the reference order does not have).
 * dory_layer_schedule reports what a layer runs: *apply_first = 1 or 0. */
int dory_layer_schedule(const dory_engine *e, uint32_t layer, int *apply_first);

/* Coarser granularity named by BASELINE.json ("per-layer forward()/backward()"):
 * dory_forward(l)  = one chunk's GA->AV->SC->AE pass with dir == FORWARD at layer l,
 * dory_backward(l) = the same with dir == BACKWARD,
 * dory_epoch       = the whole state machine for one synchronous epoch incl. weight updates
 *                    (what Engine::runPipeline does for one epoch, engine/engine.cpp:237-314). */
int dory_forward(dory_engine *e, uint32_t layer);
int dory_backward(dory_engine *e, uint32_t layer);
int dory_epoch(dory_engine *e, dory_stats *stats /* may be NULL */);
int dory_get_stats(dory_engine *e, dory_stats *stats);
/* Stream-ordered read-back for callers that keep several steps in flight (dory_epoch(e, NULL) only
 * enqueues): dory_stats_enqueue copies the statistics of everything enqueued so far into pinned host
 * memory BEHIND that work and returns at once; dory_stats_collect waits for that copy alone (not for
 * later work) and returns it.  slot in 0..3; a slot must be collected before it is enqueued again. */
int dory_stats_enqueue(dory_engine *e, uint32_t slot);
int dory_stats_collect(dory_engine *e, uint32_t slot, dory_stats *stats);

/* ---- multi-GPU (replaces CommManager / NodeManager, commmanager/commmanager.cpp:11-279) -------
 * One process per GPU.  Rank 0 calls dory_comm_unique_id, the host distributes the 128 bytes
 * (bench.py uses torch.distributed), every rank calls dory_comm_init. */
#define DORY_UNIQUE_ID_BYTES 128
int dory_comm_unique_id(void *id128);
int dory_comm_init(dory_engine *e, const void *id128);
/* The send side of an exchange is in graph.<id>.bin (forwardLocalVtxDsts / backwardLocalVtxDsts,
 * graph/graph.cpp:50-65).  The receive side replaces the per-row gvid + std::map lookup of
 * ghostReceiverGCN (gcn_ops.cpp:310-318): once, at start-up, the host tells the engine for each
 * (direction, peer) which ghost slots (0-based inside the fg / bg block) that peer's rows land in,
 * in the order the peer sends them.  dorylus_b200.dist.GhostPlan computes it. */
int dory_comm_set_recv_slots(dory_engine *e, uint32_t dir, uint32_t peer, const uint32_t *slots,
                             uint32_t n);
/* Host-only helper for callers without a side channel between ranks (host/dorylus_b200_run): the
 * receive plan of one (direction, peer) computed from the two partition images alone -- on one box
 * every rank can read its peers' graph.<id>.bin from the dataset directory.  For the rows partition
 * `peer_bin` sends to partition `my_id` in direction `dir` (its forwardLocalVtxDsts[my_id] /
 * backwardLocalVtxDsts[my_id] list, in order), writes the ghost slot each one occupies in `my_bin`'s
 * fg / bg block: globalToGhostVtcs[gvid] - localVtxCnt, the lookup ghostReceiverGCN does per row
 * (gcn_ops.cpp:310-318).  slots may be NULL to query *n.  The same call with the roles swapped gives
 * dory_comm_set_send_slots its argument.  No GPU needed; errors through dory_last_error(NULL). */
int dory_ghost_slots(const void *my_bin, size_t my_len, uint32_t my_id, const void *peer_bin, size_t peer_len,
                     uint32_t dir, uint32_t *slots, uint32_t *n);
/* Peer-memory exchange (optional, same-node ranks): instead of pack -> NCCL send/recv -> unpack,
 * dory_scatter runs ONE kernel that reads each boundary row once and stores it straight into the
 * owning peers' ghost blocks through NVLink-mapped pointers; NCCL is only used for the two barriers
 * around it.  Set-up, once: (1) dory_comm_set_send_slots -- for each peer, the ghost slot on THAT
 * peer of every row we ship (the peer's dory_comm_set_recv_slots list for us); (2) every rank
 * exports each ghost-bearing tensor ("fg"/"bg", GAT "fg_z"/"bg_d") with dory_comm_ipc_export and
 * imports its peers' blobs with dory_comm_ipc_import.  A (tensor, peer set) that is fully imported
 * switches that exchange to the peer-memory path; option "p2p" = 0 forces NCCL. */
#define DORY_IPC_BLOB_BYTES 80
int dory_comm_set_send_slots(dory_engine *e, uint32_t dir, uint32_t peer, const uint32_t *slots,
                             uint32_t n);
int dory_comm_ipc_export(dory_engine *e, uint32_t layer, const char *ghost_name, void *blob80);
int dory_comm_ipc_import(dory_engine *e, uint32_t layer, const char *ghost_name, uint32_t peer,
                         const void *blob80);
/* Global ids of the rows this partition sends to `peer` in direction `dir` (send-list order);
 * returns the count through *n; ids may be NULL to query the count. */
int dory_comm_send_gvids(const dory_engine *e, uint32_t dir, uint32_t peer, uint32_t *ids,
                         uint32_t *n);

/* What the tile plan of direction `dir` (DORY_FORWARD: CSC, DORY_BACKWARD: CSR) looks like: share of the
 * edges served from shared memory, window / tile sizes in rows, number of tiles; all zero when the
 * staged kernel is not in use for that adjacency.  Any output pointer may be NULL. */
int dory_tile_info(const dory_engine *e, uint32_t dir, double *coverage, uint32_t *window_rows, uint32_t *tile_rows,
                   uint32_t *n_tiles);

/* ---- timing on the engine's own stream (bench.py's roofline leg) ------------------------------
 * CUDA events recorded on the stream the kernels are launched on; slots 0..63. */
int dory_event_record(dory_engine *e, uint32_t slot);
int dory_event_elapsed_ms(dory_engine *e, uint32_t slot_start, uint32_t slot_stop, float *ms);
/* Writes `bytes` of HBM on the engine's stream (bench.py flushes the 126 MB L2 between timed
 * iterations with this). */
int dory_flush_l2(dory_engine *e, size_t bytes);
/* Non-tensor fp32 FMA throughput of this GPU (TFLOP/s), measured with a register-only FMA loop on
 * every SM: the compute roof bench.py quotes beside the HBM roof. */
int dory_measure_fma_peak(dory_engine *e, float *tflops);

#ifdef __cplusplus
}
#endif
#endif /* DORYLUS_B200_H */
