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

}  // extern "C"
