/* TEST INFRASTRUCTURE ONLY — never linked into, or called by, the product path.
 *
 * extern "C" driver around the reference's OWN classes, compiled unmodified from
 * /root/reference (see oracle/build.py: build_ref).  It gives tests/ and the
 * fixture generator a way to run the real reference code for the pieces of the
 * hot path that build without Boost/ZeroMQ:
 *
 *   DataLoader::preprocess   src/graph-server/graph/dataloader.cpp:225-330
 *   Graph::init              src/graph-server/graph/graph.cpp:7-115
 *   Matrix::dot              src/common/matrix.cpp:263-315   (-> cblas_sgemm)
 *   AdamOptimizer            src/weight-server/AdamOptimizer.cpp:3-51
 *   WeightTensor             src/weight-server/weighttensor.cpp (the synchronous update: local and
 *                            ghost gradient sums, then Adam -- tryApplyUpdate :246-284)
 *   the Lambda functions' tensor ops (the reference's SECOND statement of the apply step):
 *     src/funcs/gcn/ops/forward_ops.cpp, backward_ops.cpp   softmax, tanh, tanhDerivative, maskout,
 *                                                            checkAccuracy, checkLoss
 *     src/funcs/gat/ops/forward_ops.cpp, backward_ops.cpp   leakyReLU(+Derivative), edgeMatMul,
 *                                                            expandDot, expandHadamardMul, reduce
 *   sequenced exactly as src/funcs/gcn/main.cpp:83-108 (finalLayer), :176-189 (backwardLayer),
 *   :244-250 (forwardLayer) and src/funcs/gat/main.cpp:84-101, :130-170 do -- main.cpp itself needs
 *   ZeroMQ and the AWS SDK and does not build here.
 *
 * Output: oracle/_ref/libdoryref.so (git-ignored, travels to the GPU box).
 */
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "graph-server/graph/dataloader.hpp"
#include "graph-server/graph/graph.hpp"
#include "common/matrix.hpp"
#include "weight-server/AdamOptimizer.hpp"
#include "weight-server/weighttensor.hpp"

#include <algorithm>
#include <mutex>
#include <cassert>
#include <cmath>

/* The two Lambda functions define the same global names (softmax, tanh, ...), so each pair of
 * translation units is compiled inside its own namespace, from the files where they lie.  Their
 * headers only pull common/matrix.hpp and common/utils.hpp, which are already included above. */
namespace ref_funcs_gcn {
#include "funcs/gcn/ops/forward_ops.cpp"
#include "funcs/gcn/ops/backward_ops.cpp"
}  // namespace ref_funcs_gcn
#undef __FWD_OPS_HPP__
#undef __BKWD_OPS_HPP__
namespace ref_funcs_gat {
#include "funcs/gat/ops/forward_ops.cpp"
#include "funcs/gat/ops/backward_ops.cpp"
}  // namespace ref_funcs_gat

extern "C" {

/* Runs the reference preprocessor: <dir>/graph.bsnap.edges + <dir>/graph.bsnap.parts
 * -> <dir>/graph.<nodeId>.bin.  `dir` must end with '/'. */
int ref_preprocess(const char *dir, unsigned nodeId, unsigned numNodes, int undirected) {
    DataLoader dl(std::string(dir), nodeId, numNodes, undirected != 0);
    dl.preprocess();
    return 0;
}

struct RefGraph {
    Graph g;
};

void *ref_graph_open(const char *file) {
    RefGraph *h = new RefGraph();
    h->g.init(std::string(file));
    return h;
}

void ref_graph_close(void *h) { delete static_cast<RefGraph *>(h); }

/* counts: [localVtxCnt, globalVtxCnt, srcGhostCnt, dstGhostCnt, inEdges, outEdges, globalEdges,
 *          fwd.nnz, bwd.nnz] */
void ref_graph_counts(void *h, uint64_t *out) {
    Graph &g = static_cast<RefGraph *>(h)->g;
    out[0] = g.localVtxCnt; out[1] = g.globalVtxCnt; out[2] = g.srcGhostCnt; out[3] = g.dstGhostCnt;
    out[4] = g.localInEdgeCnt; out[5] = g.localOutEdgeCnt; out[6] = g.globalEdgeCnt;
    out[7] = g.forwardAdj.nnz; out[8] = g.backwardAdj.nnz;
}

/* Copy the arrays Graph::init built.  Any pointer may be NULL to skip it. */
void ref_graph_arrays(void *h, unsigned *l2g, float *norms,
                      uint64_t *colPtrs, unsigned *rowIdxs, float *fvals,
                      uint64_t *rowPtrs, unsigned *colIdxs, float *bvals) {
    Graph &g = static_cast<RefGraph *>(h)->g;
    const unsigned V = g.localVtxCnt;
    if (l2g) std::memcpy(l2g, g.localToGlobalId.data(), sizeof(unsigned) * V);
    if (norms) std::memcpy(norms, g.vtxDataVec.data(), sizeof(float) * V);
    if (colPtrs) std::memcpy(colPtrs, g.forwardAdj.columnPtrs, sizeof(uint64_t) * (V + 1));
    if (rowIdxs) std::memcpy(rowIdxs, g.forwardAdj.rowIdxs, sizeof(unsigned) * g.forwardAdj.nnz);
    if (fvals) std::memcpy(fvals, g.forwardAdj.values, sizeof(float) * g.forwardAdj.nnz);
    if (rowPtrs) std::memcpy(rowPtrs, g.backwardAdj.rowPtrs, sizeof(uint64_t) * (V + 1));
    if (colIdxs) std::memcpy(colIdxs, g.backwardAdj.columnIdxs, sizeof(unsigned) * g.backwardAdj.nnz);
    if (bvals) std::memcpy(bvals, g.backwardAdj.values, sizeof(float) * g.backwardAdj.nnz);
}

/* Ghost maps as (gvid, lvid) pairs in std::map order (ascending gvid). which: 0 = src, 1 = dst. */
void ref_graph_ghosts(void *h, int which, unsigned *gvids, unsigned *lvids) {
    Graph &g = static_cast<RefGraph *>(h)->g;
    std::map<unsigned, unsigned> &m = which == 0 ? g.srcGhostVtcs : g.dstGhostVtcs;
    size_t i = 0;
    for (auto &kv : m) { gvids[i] = kv.first; lvids[i] = kv.second; ++i; }
}

/* Per-peer send lists. which: 0 = forwardLocalVtxDsts, 1 = backwardLocalVtxDsts.
 * Returns the list length; copies when out != NULL. */
unsigned ref_graph_sendlist(void *h, int which, unsigned peer, unsigned *out) {
    Graph &g = static_cast<RefGraph *>(h)->g;
    std::vector<std::vector<unsigned>> &l = which == 0 ? g.forwardLocalVtxDsts : g.backwardLocalVtxDsts;
    if (peer >= l.size()) return 0;
    if (out) std::memcpy(out, l[peer].data(), sizeof(unsigned) * l[peer].size());
    return (unsigned)l[peer].size();
}

/* Matrix::dot through the reference's own code path (row-major fp32, beta = 0). */
int ref_matrix_dot(const float *A, unsigned ra, unsigned ca, const float *B, unsigned rb, unsigned cb,
                   int tA, int tB, float scale, float *out) {
    Matrix a(ra, ca, const_cast<float *>(A));
    Matrix b(rb, cb, const_cast<float *>(B));
    Matrix c = a.dot(b, tA != 0, tB != 0, scale);
    std::memcpy(out, c.getData(), c.getDataSize());
    delete[] c.getData();
    return 0;
}

void *ref_adam_create(float lr, const unsigned *dims, unsigned ndims) {
    return new AdamOptimizer(lr, std::vector<unsigned>(dims, dims + ndims));
}
void ref_adam_update(void *h, unsigned layer, float *weight, float *grad) {
    static_cast<AdamOptimizer *>(h)->update(layer, weight, grad);
}
void ref_adam_destroy(void *h) { delete static_cast<AdamOptimizer *>(h); }


/* ---------------------------------------------------------------- src/funcs (Lambda) tensor ops
 * Inputs are copied into new[] buffers because the reference ops return freshly allocated
 * matrices and some (maskout, operator/=) write in place. */
static Matrix own_copy(const float *p, unsigned r, unsigned c) {
    float *d = new float[(size_t)r * c];
    std::memcpy(d, p, sizeof(float) * (size_t)r * c);
    return Matrix(r, c, d);
}
static void take(Matrix &m, float *out) {
    if (out) std::memcpy(out, m.getData(), m.getDataSize());
    delete[] m.getData();
}

/* forwardLayer, funcs/gcn/main.cpp:244-250: Z = AH.dot(W); H = tanh(Z). */
void ref_funcs_gcn_forward(const float *ah, const float *w, unsigned V, unsigned Fin, unsigned Fout,
                           float *z, float *h) {
    Matrix AH = own_copy(ah, V, Fin), W = own_copy(w, Fin, Fout);
    Matrix Z = AH.dot(W);
    Matrix H = ref_funcs_gcn::tanh(Z);
    take(H, h);
    take(Z, z);
    take(AH, nullptr);
    take(W, nullptr);
}

/* finalLayer, funcs/gcn/main.cpp:83-108: softmax, (accuracy / loss over all rows), maskout,
 * d_out = (preds - labels) / trainset_size, interGrad = d_out . W^T, d_weights = AH^T . d_out.
 * `scale` is what the caller divides by: the Lambda payload carries the integer
 * globalVtxCnt * TRAIN_PORTION (lambda_comm.cpp:156), CPUComm the float product (CPU_comm.cpp:121). */
void ref_funcs_gcn_final(const float *ah, const float *w, const float *lab, unsigned V, unsigned Fin,
                         unsigned C, float scale, float *preds_out, unsigned *correct_out, float *loss_out,
                         float *masked_out, float *d_out_out, float *grad, float *dW) {
    Matrix AH = own_copy(ah, V, Fin), W = own_copy(w, Fin, C), labels = own_copy(lab, V, C);
    Matrix Z = AH.dot(W);
    Matrix preds = ref_funcs_gcn::softmax(Z);
    take(Z, nullptr);
    if (preds_out) std::memcpy(preds_out, preds.getData(), preds.getDataSize());
    if (correct_out) *correct_out = ref_funcs_gcn::checkAccuracy(preds, labels);
    if (loss_out) *loss_out = ref_funcs_gcn::checkLoss(preds, labels);
    ref_funcs_gcn::maskout(preds, labels);
    if (masked_out) std::memcpy(masked_out, preds.getData(), preds.getDataSize());
    Matrix d_out = preds - labels;
    d_out /= scale;
    Matrix interGrad = d_out.dot(W, false, true);
    Matrix d_weights = AH.dot(d_out, true, false);
    take(interGrad, grad);
    take(d_weights, dW);
    take(d_out, d_out_out);
    take(preds, nullptr);
    take(labels, nullptr);
    take(AH, nullptr);
    take(W, nullptr);
}

/* backwardLayer, funcs/gcn/main.cpp:176-189. */
void ref_funcs_gcn_backward(const float *ah, const float *z, const float *aTg, const float *w, unsigned V,
                            unsigned Fin, unsigned Fout, float *resultGrad, float *dW) {
    Matrix AH = own_copy(ah, V, Fin), Z = own_copy(z, V, Fout), grad = own_copy(aTg, V, Fout),
           W = own_copy(w, Fin, Fout);
    Matrix actDeriv = ref_funcs_gcn::tanhDerivative(Z);
    Matrix interGrad = grad * actDeriv;
    Matrix result = interGrad.dot(W, false, true);
    Matrix d_weights = AH.dot(interGrad, true, false);
    take(result, resultGrad);
    take(d_weights, dW);
    take(interGrad, nullptr);
    take(actDeriv, nullptr);
    take(AH, nullptr);
    take(Z, nullptr);
    take(grad, nullptr);
    take(W, nullptr);
}

/* GAT edge forward, funcs/gat/main.cpp:84-94: az = edgeMatMul(eInfo, Z, a); A = leakyReLU(az). */
void ref_funcs_gat_edge_forward(const float *z, const float *a, const unsigned long long *edgePtrs, unsigned V,
                                unsigned F, unsigned nEdges, float *az, float *A) {
    Matrix Z = own_copy(z, V, F), av = own_copy(a, F, 1);
    EdgeInfo eInfo{V, nEdges, const_cast<unsigned long long *>(edgePtrs)};
    Matrix edgeValInputs = ref_funcs_gat::edgeMatMul(eInfo, Z, av);
    Matrix edgeVals = ref_funcs_gat::leakyReLU(edgeValInputs);
    take(edgeVals, A);
    take(edgeValInputs, az);
    take(Z, nullptr);
    take(av, nullptr);
}

/* GAT edge backward, funcs/gat/main.cpp:150-170: dLRelu = leakyReLUDerivative(az);
 * dAct = expandHadamardMul(grad, dLRelu); dA = dAct . a; da = (z^T z) . reduce(dAct)^T.
 * reduce() accumulates into an uninitialised new[] buffer in the reference (quirk Q11); the driver
 * cannot change that, so it is only safe for small dAct -- the caller passes small cases and the
 * result is checked for finiteness before it is used. */
void ref_funcs_gat_edge_backward(const float *grad, const float *az, const float *z, const float *a,
                                 const unsigned long long *edgePtrs, unsigned V, unsigned F, unsigned nEdges,
                                 float *dA, float *dAct_out) {
    Matrix G = own_copy(grad, V, F), AZ = own_copy(az, nEdges, 1), av = own_copy(a, F, 1);
    (void)z;
    EdgeInfo eInfo{V, nEdges, const_cast<unsigned long long *>(edgePtrs)};
    Matrix dLRelu = ref_funcs_gat::leakyReLUDerivative(AZ);
    Matrix dAct = ref_funcs_gat::expandHadamardMul(G, dLRelu, eInfo);
    Matrix dAm = dAct.dot(av);
    take(dAm, dA);
    take(dAct, dAct_out);
    take(dLRelu, nullptr);
    take(G, nullptr);
    take(AZ, nullptr);
    take(av, nullptr);
}

/* WeightTensor in sync mode (weighttensor.cpp): what one weight server does with the gradients of
 * a layer -- those it receives from graph servers ("local") and those relayed by the other weight
 * servers ("ghost") -- before and when it steps Adam. */
struct RefWeightTensor {
    std::mutex wmtx, umtx;
    WeightTensor *wt = nullptr;
};
void *ref_wt_create(const float *w, unsigned rows, unsigned cols, unsigned localTot, unsigned ghostTot) {
    RefWeightTensor *h = new RefWeightTensor();
    Matrix m = own_copy(w, rows, cols);
    h->wt = new WeightTensor(m, &h->wmtx, &h->umtx, true);
    h->wt->setLocalUpdTot(localTot);
    h->wt->setGhostUpdTot(ghostTot);
    return h;
}
unsigned ref_wt_local_update(void *h, const float *upd, unsigned n) {
    std::vector<float> u(upd, upd + n);
    return static_cast<RefWeightTensor *>(h)->wt->localUpdate(u.data());
}
unsigned ref_wt_ghost_update(void *h, const float *upd, unsigned n) {
    std::vector<float> u(upd, upd + n);
    return static_cast<RefWeightTensor *>(h)->wt->ghostUpdate(u.data());
}
/* Returns 1 when the update was applied (both counters had reached their totals), 0 otherwise. */
int ref_wt_try_apply(void *h, void *adam, unsigned layer, float *w_out) {
    WeightTensor *wt = static_cast<RefWeightTensor *>(h)->wt;
    const std::string info = wt->tryApplyUpdate(static_cast<AdamOptimizer *>(adam), layer);
    Matrix &m = wt->currMat();
    std::memcpy(w_out, m.getData(), m.getDataSize());
    return info.empty() ? 0 : 1;
}
void ref_wt_destroy(void *h) {
    RefWeightTensor *r = static_cast<RefWeightTensor *>(h);
    r->wt->free();
    delete r->wt;
    delete r;
}

/* Chunk::operator< (common/utils.hpp:76-89) -- the priority of the reference's chunk queues -- and
 * Chunk::isFirstLayer / isLastLayer, on chunks given as 8 unsigned fields
 * {localId, globalId, lowBound, upBound, layer, dir, epoch, vertex}. */
static Chunk to_chunk(const unsigned *f) {
    return Chunk{f[0], f[1], f[2], f[3], f[4], f[5] ? PROP_TYPE::BACKWARD : PROP_TYPE::FORWARD, f[6], f[7] != 0};
}
int ref_chunk_less(const unsigned *a, const unsigned *b) { return to_chunk(a) < to_chunk(b) ? 1 : 0; }
int ref_chunk_flags(const unsigned *a) {
    Chunk c = to_chunk(a);
    return (c.isFirstLayer() ? 1 : 0) | (c.isLastLayer() ? 2 : 0);
}

/* expandDot, funcs/gat/ops/backward_ops.cpp (same body as CPU_comm.cpp:299-319). */
void ref_funcs_gat_expand_dot(const float *m, const float *v, const unsigned long long *edgePtrs, unsigned V,
                              unsigned F, unsigned nEdges, float *out) {
    Matrix M = own_copy(m, V, F), vv = own_copy(v, F, 1);
    EdgeInfo eInfo{V, nEdges, const_cast<unsigned long long *>(edgePtrs)};
    Matrix r = ref_funcs_gat::expandDot(M, vv, eInfo);
    take(r, out);
    take(M, nullptr);
    take(vv, nullptr);
}

}  // extern "C"
