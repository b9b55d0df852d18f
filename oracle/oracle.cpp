/* TEST INFRASTRUCTURE ONLY — the CPU oracle.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library.  The product path (dorylus_b200/csrc, include/) never
 * links, imports or calls anything in oracle/.
 *
 * What this is: a line-faithful CPU restatement of the pieces of the reference's hot path
 * that cannot be compiled here (they include Boost / ZeroMQ headers that are absent):
 *
 *   Engine::aggregateGCN (CPU branch)      src/graph-server/engine/ops/gcn_ops.cpp:130-191
 *   Engine::aggregateGAT (CPU branch)      src/graph-server/engine/ops/gat_ops.cpp:173-243
 *   Engine::predictGAT / Engine::softmax   gat_ops.cpp:247-265, engine/ops/tensors.cpp:7-26
 *   Engine::srcVFeats2eFeats / dstVFeats2eFeats   engine/utils.cpp:655-705
 *   CPUComm::vtxNNForwardGCN / BackwardGCN        commmanager/CPU_comm.cpp:98-159
 *   CPUComm::vtxNNForwardGAT / BackwardGAT        CPU_comm.cpp:161-188
 *   CPUComm::edgNNForwardGAT / BackwardGAT        CPU_comm.cpp:190-242
 *   helpers activate/softmax/expandDot/...        CPU_comm.cpp:265-471
 *   WeightServer::xavierInitializer/kaiming       weight-server/weightserver.cpp:567-612
 *   AdamOptimizer                                 weight-server/AdamOptimizer.cpp:3-51
 *   WeightTensor::tryApplyUpdate (sync branch)    weight-server/weighttensor.cpp:263-284
 *
 * Dense contractions go through the same entry the reference uses, cblas_sgemm with
 * CblasRowMajor and beta = 0 (src/common/matrix.cpp:263-315); the OpenBLAS build is the one
 * inside scipy.libs (the reference pins no OpenBLAS version: gnnman/helpers/blas.install:13).
 *
 * Parity pinning (tests/test_oracle.py, fixtures under tests/golden/ written by make_golden.py):
 *   - xavier():   against miscs/dgl-non-sampling/data/raw0, raw1 (xavier.npz), and the seeded RNG
 *                 stream over 643,000 draws against miscs/check-correctness/weights-602-1000-41
 *   - adam, sgemm wrapper, graph arrays: against the compiled reference (oracle/_ref)
 *   - aggregate / apply (ah, z, h, soft-max, grad, aTg, dW): against the reference's dense-numpy GCN
 *                 miscs/numpy-gnn, IMPORTED and run by make_golden.py on a small symmetric graph
 *                 (numpy_gnn.npz), and against the dense statement (D^-1/2 A D^-1/2 + D^-1) X built
 *                 from the compiled reference loader's own arrays
 *   - apply step a second time, against the reference's own C++: the Lambda functions' tensor ops
 *                 (src/funcs/gcn/ops, src/funcs/gat/ops) compiled into oracle/_ref and sequenced as
 *                 funcs/gcn/main.cpp / funcs/gat/main.cpp do (funcs_ops.npz + live): soft-max, the
 *                 float-wise maskout (Q6), the scaled output gradient, tanh', weight gradients, the
 *                 GAT edge scores and their backward
 *   - still unpinned by the reference (nothing that states them runs here): CPUComm's validation
 *                 statistics getTrainStat (the Lambda twin sendAccLoss sits in a ZeroMQ translation
 *                 unit) and which of the two loss scales applies (float V*0.66 in CPUComm, its integer
 *                 truncation in the Lambda payload; the oracle follows CPUComm).
 */
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <vector>

#include <omp.h>

#include "cblas.h" /* oracle/shim/cblas.h */

typedef float FeatType;
typedef float EdgeType;

#define TRAIN_PORTION 0.66 /* src/common/utils.hpp:60 */
#define VAL_PORTION 0.1    /* src/common/utils.hpp:61 */

namespace {

inline FeatType *getVtxFeat(FeatType *base, unsigned lvid, unsigned featDim) {
    return base + (size_t)lvid * featDim; /* engine.hpp getVtxFeat */
}

/* Matrix::dot, src/common/matrix.cpp:263-315 (four transpose cases, beta = 0). */
void mdot(const float *A, unsigned ra, unsigned ca, const float *B, unsigned rb, unsigned cb,
          bool t1, bool t2, float scale, float *out, unsigned *m_out = nullptr, unsigned *n_out = nullptr) {
    unsigned m = 0, k = 0, n = 0;
    if (!t1 && !t2) {
        m = ra, k = ca, n = cb;
        cblas_sgemm(CblasRowMajor, CblasNoTrans, CblasNoTrans, m, n, k, scale, A, k, B, n, 0.0, out, n);
    } else if (t1 && t2) {
        m = ca, k = ra, n = rb;
        cblas_sgemm(CblasRowMajor, CblasTrans, CblasTrans, m, n, k, scale, A, m, B, k, 0.0, out, n);
    } else if (t1) {
        m = ca, k = ra, n = cb;
        cblas_sgemm(CblasRowMajor, CblasTrans, CblasNoTrans, m, n, k, scale, A, m, B, n, 0.0, out, n);
    } else {
        m = ra, k = ca, n = rb;
        cblas_sgemm(CblasRowMajor, CblasNoTrans, CblasTrans, m, n, k, scale, A, k, B, k, 0.0, out, n);
    }
    (void)rb; (void)k;
    if (m_out) *m_out = m;
    if (n_out) *n_out = n;
}

template <class It>
unsigned argmax(It first, It last) { /* graph-server utils argmax */
    It res = first;
    for (It it = first; it != last; ++it)
        if (*it > *res) res = it;
    return (unsigned)(res - first);
}

/* CPU_comm.cpp:276-297 (identical body in engine/ops/tensors.cpp:7-26). */
void softmax_rows(const FeatType *src, FeatType *dst, unsigned rows, unsigned length) {
#pragma omp parallel for
    for (unsigned r = 0; r < rows; ++r) {
        const FeatType *vecSrc = src + (size_t)r * length;
        FeatType *vecDst = dst + (size_t)r * length;
        FeatType denom = 1e-20;
        FeatType maxEle = *(std::max_element(vecSrc, vecSrc + length));
        for (unsigned c = 0; c < length; ++c) {
            vecDst[c] = std::exp(vecSrc[c] - maxEle);
            denom += vecDst[c];
        }
        for (unsigned c = 0; c < length; ++c) vecDst[c] /= denom;
    }
}

}  // namespace

extern "C" {

void orc_set_threads(int n) {
    omp_set_num_threads(n);
    scipy_openblas_set_num_threads(n);
}
int orc_get_threads(void) { return omp_get_max_threads(); }

/* ---------------------------------------------------------------- edge pointer tables
 * Engine::srcVFeats2eFeats (engine/utils.cpp:655-678) and dstVFeats2eFeats (:682-705):
 * one FeatType* per edge pointing at the source row (local tensor or ghost tensor).
 * Only the first half (the "src" pointers) is consumed by aggregate*. */
FeatType **orc_build_edge_table(const uint64_t *ptrs, const unsigned *idxs, unsigned vtcsCnt,
                                FeatType *vtcsTensor, FeatType *ghostTensor, unsigned featDim,
                                unsigned start, unsigned end) {
    /* [start, end) bounds the rows whose entries are filled (bench.py's bounded CPU sample); the
     * reference always fills all of them (start = 0, end = vtcsCnt).  The array keeps its full
     * size so that entries stay addressed by global edge id. */
    const uint64_t edgeCnt = ptrs[vtcsCnt];
    FeatType **eVtxFeatsBuf = new FeatType *[2 * edgeCnt];
    FeatType **eSrcVtxFeats = eVtxFeatsBuf;
    FeatType **eDstVtxFeats = eSrcVtxFeats + edgeCnt;
    unsigned long long edgeItr = ptrs[start];
    for (unsigned lvid = start; lvid < end; ++lvid) {
        for (unsigned long long eid = ptrs[lvid]; eid < ptrs[lvid + 1]; ++eid) {
            unsigned srcVid = idxs[eid];
            if (srcVid < vtcsCnt)
                eSrcVtxFeats[edgeItr] = getVtxFeat(vtcsTensor, srcVid, featDim);
            else
                eSrcVtxFeats[edgeItr] = getVtxFeat(ghostTensor, srcVid - vtcsCnt, featDim);
            eDstVtxFeats[edgeItr] = getVtxFeat(vtcsTensor, lvid, featDim);
            ++edgeItr;
        }
    }
    return eVtxFeatsBuf;
}
void orc_free_edge_table(FeatType **t) { delete[] t; }

/* ---------------------------------------------------------------- aggregateGCN
 * gcn_ops.cpp:130-191.  `ptrs/vals` are forwardAdj.columnPtrs/values for FORWARD and
 * backwardAdj.rowPtrs/values for BACKWARD (the two branches at :173-189 differ only in
 * which arrays they read).  `featTensor` is x / h[l-1] (FWD) or grad[l] (BWD); `inputTensor`
 * is the fedge / bedge table.  OpenMP over lvid as in the _CPU_ENABLED_ build (:159-161). */
void orc_aggregate_gcn(const uint64_t *ptrs, const EdgeType *vals, const EdgeType *vtxDataVec,
                       FeatType *featTensorBase, FeatType **inputTensor, unsigned featDim,
                       unsigned start, unsigned end, FeatType *outputTensor) {
    FeatType *featTensor = getVtxFeat(featTensorBase, start, featDim);
    FeatType *chunkPtr = getVtxFeat(outputTensor, start, featDim);
    std::memcpy(chunkPtr, featTensor, sizeof(FeatType) * (size_t)(end - start) * featDim);
#pragma omp parallel for
    for (unsigned lvid = start; lvid < end; lvid++) {
        FeatType *currDataDst = getVtxFeat(outputTensor, lvid, featDim);
        {
            const EdgeType normFactor = vtxDataVec[lvid];
            for (unsigned i = 0; i < featDim; ++i) currDataDst[i] *= normFactor;
        }
        for (uint64_t eid = ptrs[lvid]; eid < ptrs[lvid + 1]; ++eid) {
            EdgeType normFactor = vals[eid];
            for (unsigned j = 0; j < featDim; ++j) currDataDst[j] += inputTensor[eid][j] * normFactor;
        }
    }
}

/* ---------------------------------------------------------------- aggregateGAT
 * gat_ops.cpp:173-243.
 * FORWARD (:201-220): ah[start:end] = z[start:end]; ah[v] += sum_e A[e] * fedge[e].
 * BACKWARD (:221-241): aTg[v] += sum_{out e} bvals[e] * bedge[e] + sum_{in e} dA[e] * fedge[e].
 * Quirk Q11: the reference accumulates into a never-initialised aTg; the oracle zero-fills the
 * chunk rows first (the reference's GPU branch, :139-170, computes fresh values too). */
void orc_aggregate_gat_fwd(const uint64_t *colPtrs, const EdgeType *A, FeatType *zBase,
                           FeatType **inputFTensor, unsigned featDim, unsigned start, unsigned end,
                           FeatType *outputTensor) {
    std::memcpy(getVtxFeat(outputTensor, start, featDim), getVtxFeat(zBase, start, featDim),
                sizeof(FeatType) * (size_t)(end - start) * featDim);
#pragma omp parallel for
    for (unsigned lvid = start; lvid < end; lvid++) {
        FeatType *currDataDst = getVtxFeat(outputTensor, lvid, featDim);
        for (uint64_t eid = colPtrs[lvid]; eid < colPtrs[lvid + 1]; ++eid) {
            EdgeType edgeWeight = A[eid];
            for (unsigned j = 0; j < featDim; ++j) currDataDst[j] += inputFTensor[eid][j] * edgeWeight;
        }
    }
}

void orc_aggregate_gat_bwd(const uint64_t *colPtrs, const EdgeType *dA, FeatType **inputFTensor,
                           const uint64_t *rowPtrs, const EdgeType *bvals, FeatType **inputBTensor,
                           unsigned featDim, unsigned start, unsigned end, FeatType *outputTensor) {
    std::memset(getVtxFeat(outputTensor, start, featDim), 0,
                sizeof(FeatType) * (size_t)(end - start) * featDim); /* Q11 */
#pragma omp parallel for
    for (unsigned lvid = start; lvid < end; lvid++) {
        FeatType *currDataDst = getVtxFeat(outputTensor, lvid, featDim);
        for (uint64_t eid = rowPtrs[lvid]; eid < rowPtrs[lvid + 1]; ++eid) {
            EdgeType edgeWeight = bvals[eid];
            for (unsigned j = 0; j < featDim; ++j) currDataDst[j] += inputBTensor[eid][j] * edgeWeight;
        }
        for (uint64_t eid = colPtrs[lvid]; eid < colPtrs[lvid + 1]; ++eid) {
            EdgeType edgeGrad = dA[eid];
            for (unsigned j = 0; j < featDim; ++j) currDataDst[j] += inputFTensor[eid][j] * edgeGrad;
        }
    }
}

/* predictGAT, gat_ops.cpp:247-265: grad = softmax_rows(logits) - labels.
 * Quirk Q9: the reference passes savedNNTensors["az"] as `logits`; the caller chooses. */
void orc_predict_gat(const FeatType *logits, const FeatType *labels, unsigned rows, unsigned cols,
                     FeatType *outputDeriv) {
    softmax_rows(logits, outputDeriv, rows, cols);
    for (size_t i = 0; i < (size_t)rows * cols; ++i) outputDeriv[i] -= labels[i];
}

void orc_softmax(const FeatType *src, unsigned rows, unsigned cols, FeatType *dst) {
    softmax_rows(src, dst, rows, cols);
}

/* ---------------------------------------------------------------- Matrix::dot */
void orc_matrix_dot(const float *A, unsigned ra, unsigned ca, const float *B, unsigned rb, unsigned cb,
                    int tA, int tB, float scale, float *out) {
    mdot(A, ra, ca, B, rb, cb, tA != 0, tB != 0, scale, out);
}

/* ---------------------------------------------------------------- CPUComm GCN
 * vtxNNForwardGCN, hidden layer (CPU_comm.cpp:98-107): z = ah . W ; h = tanh(z). */
void orc_vtx_forward_gcn_hidden(const FeatType *ah, const FeatType *W, unsigned V, unsigned Fin,
                                unsigned Fout, FeatType *z, FeatType *h) {
    mdot(ah, V, Fin, W, Fin, Fout, false, false, 1.0f, z);
    const size_t n = (size_t)V * Fout;
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) h[i] = std::tanh(z[i]); /* activate(), :265-274 */
}

/* vtxNNForwardGCN, last layer (CPU_comm.cpp:108-133).
 * Outputs: predictions BEFORE maskout (pred_out, may be NULL), acc & loss sums over the
 * validation slice (getTrainStat :448-462), grad = d . W^T, dW = ah^T . d where
 * d = (maskout(P) - lab) / (globalVtxCnt * TRAIN_PORTION).
 * Quirk Q6 (maskout :464-471): memcpy of (end - stt) FLOATS, not rows. */
void orc_vtx_forward_gcn_last(const FeatType *ah, const FeatType *W, const FeatType *lab, unsigned V,
                              unsigned Fin, unsigned C, unsigned globalVtxCnt, FeatType *pred_out,
                              float *acc_out, float *loss_out, FeatType *grad, FeatType *dW,
                              FeatType *d_out) {
    const size_t n = (size_t)V * C;
    std::vector<FeatType> z(n), predictions(n);
    mdot(ah, V, Fin, W, Fin, C, false, false, 1.0f, z.data());
    softmax_rows(z.data(), predictions.data(), V, C);
    if (pred_out) std::memcpy(pred_out, predictions.data(), n * sizeof(FeatType));

    /* getTrainStat */
    float acc = 0.0, loss = 0.0;
    {
        unsigned featDim = C;
        unsigned valStt = (unsigned)(V * TRAIN_PORTION);
        unsigned valEnd = valStt + (unsigned)(V * VAL_PORTION);
        for (unsigned i = valStt; i < valEnd; i++) {
            const FeatType *currLabel = lab + (size_t)i * C;
            const FeatType *currPred = predictions.data() + (size_t)i * C;
            acc += currLabel[argmax(currPred, currPred + featDim)];
            loss -= std::log(currPred[argmax(currLabel, currLabel + featDim)]);
        }
    }
    if (acc_out) *acc_out = acc;
    if (loss_out) *loss_out = loss;

    /* maskout */
    {
        unsigned end = V;
        unsigned stt = (unsigned)(end * TRAIN_PORTION);
        std::memcpy(predictions.data() + (size_t)stt * C, lab + (size_t)stt * C,
                    sizeof(FeatType) * (end - stt));
    }
    /* hadamardSub (:424-435) then  d_output /= globalVtxCnt * TRAIN_PORTION  (:121; Matrix::operator/=
     * divides each element by the float rhs, src/common/matrix.cpp). */
    std::vector<FeatType> d(n);
    for (size_t ui = 0; ui < n; ++ui) d[ui] = predictions[ui] - lab[ui];
    {
        float rhs = globalVtxCnt * TRAIN_PORTION;
        for (size_t ui = 0; ui < n; ++ui) d[ui] /= rhs;
    }
    if (d_out) std::memcpy(d_out, d.data(), n * sizeof(FeatType));
    mdot(d.data(), V, C, W, Fin, C, false, true, 1.0f, grad); /* interGrad = d . W^T  (:123) */
    mdot(ah, V, Fin, d.data(), V, C, true, false, 1.0f, dW);  /* weightUpdates = ah^T . d (:128) */
}

/* vtxNNBackwardGCN (CPU_comm.cpp:137-159):
 * g = aTg (*) (1 - tanh(z)^2); dW = ah^T . g; if layer != 0: grad = g . W^T. */
void orc_vtx_backward_gcn(const FeatType *aTg, const FeatType *z, const FeatType *ah, const FeatType *W,
                          unsigned V, unsigned Fin, unsigned Fout, int layer_nonzero, FeatType *dW,
                          FeatType *grad) {
    const size_t n = (size_t)V * Fout;
    std::vector<FeatType> interGrad(n);
#pragma omp parallel for
    for (size_t i = 0; i < n; ++i) {
        FeatType actDeriv = 1 - std::pow(std::tanh(z[i]), 2); /* activateDerivative :437-446 */
        interGrad[i] = aTg[i] * actDeriv;                     /* Matrix::operator*(Matrix&) */
    }
    mdot(ah, V, Fin, interGrad.data(), V, Fout, true, false, 1.0f, dW);
    if (layer_nonzero) mdot(interGrad.data(), V, Fout, W, Fin, Fout, false, true, 1.0f, grad);
}

/* ---------------------------------------------------------------- CPUComm GAT
 * vtxNNForwardGAT (:161-169): z = feats . W. */
void orc_vtx_forward_gat(const FeatType *feats, const FeatType *W, unsigned V, unsigned Fin,
                         unsigned Fout, FeatType *z) {
    mdot(feats, V, Fin, W, Fin, Fout, false, false, 1.0f, z);
}
/* vtxNNBackwardGAT (:171-188): dW = h^T . aTg; if layer != 0: grad[layer-1] = aTg . W^T. */
void orc_vtx_backward_gat(const FeatType *h, const FeatType *aTg, const FeatType *W, unsigned V,
                          unsigned Fin, unsigned Fout, int layer_nonzero, FeatType *dW, FeatType *grad) {
    mdot(h, V, Fin, aTg, V, Fout, true, false, 1.0f, dW);
    if (layer_nonzero) mdot(aTg, V, Fout, W, Fin, Fout, false, true, 1.0f, grad);
}

/* edgNNForwardGAT (:190-203): az[e] = z[dst(e)] . a  (expandDot :299-319, accumulates j in order
 * into a zero-initialised slot), A[e] = leakyRelu(az[e]) (:384-395, alpha = 0.01). */
void orc_edg_forward_gat(const FeatType *z, const FeatType *a, const uint64_t *colPtrs, unsigned V,
                         unsigned featDim, FeatType *az, FeatType *A) {
    const uint64_t nnz = colPtrs[V];
    std::memset(az, 0, sizeof(FeatType) * nnz);
#pragma omp parallel for
    for (unsigned lvid = 0; lvid < V; lvid++) {
        const FeatType *mPtr = z + (size_t)lvid * featDim;
        for (unsigned long long eid = colPtrs[lvid]; eid < colPtrs[lvid + 1]; ++eid)
            for (unsigned j = 0; j < featDim; ++j) az[eid] += mPtr[j] * a[j];
    }
    FeatType alpha = 0.01;
#pragma omp parallel for
    for (uint64_t i = 0; i < nnz; ++i) A[i] = (az[i] > 0) ? az[i] : alpha * az[i];
}

/* edgNNBackwardGAT (:205-242):
 * dLRelu[e] = az[e] > 0 ? 1 : 0.01; dAct[e,:] = grad[dst(e),:] * dLRelu[e]  (E x F', materialised);
 * dA = dAct . a (E x 1, sgemm); dAct_reduce = column sums (Q11: reference sums into uninitialised
 * memory, the oracle starts from 0); zz = z^T . z; da = zz . dAct_reduce^T  (F' x 1). */
void orc_edg_backward_gat(const FeatType *grad, const FeatType *az, const FeatType *z, const FeatType *a,
                          const uint64_t *colPtrs, unsigned V, unsigned featDim, FeatType *dA,
                          FeatType *da) {
    const uint64_t nnz = colPtrs[V];
    FeatType alpha = 0.01;
    std::vector<FeatType> dLRelu(nnz);
    for (uint64_t i = 0; i < nnz; ++i) dLRelu[i] = (az[i] > 0) ? 1 : alpha;
    std::vector<FeatType> dAct((size_t)nnz * featDim, 0.0f);
#pragma omp parallel for
    for (unsigned lvid = 0; lvid < V; lvid++) {
        const FeatType *mPtr = grad + (size_t)lvid * featDim;
        for (unsigned long long eid = colPtrs[lvid]; eid < colPtrs[lvid + 1]; ++eid) {
            FeatType normFactor = dLRelu[eid];
            for (unsigned j = 0; j < featDim; ++j) dAct[eid * featDim + j] = mPtr[j] * normFactor;
        }
    }
    if (nnz > 0) mdot(dAct.data(), (unsigned)nnz, featDim, a, featDim, 1, false, false, 1.0f, dA);
    std::vector<FeatType> dAct_reduce(featDim, 0.0f);
    for (unsigned i = 0; i < featDim; i++)
        for (uint64_t eid = 0; eid < nnz; eid++) dAct_reduce[i] += dAct[eid * featDim + i];
    std::vector<FeatType> zz((size_t)featDim * featDim);
    mdot(z, V, featDim, z, V, featDim, true, false, 1.0f, zz.data());
    mdot(zz.data(), featDim, featDim, dAct_reduce.data(), 1, featDim, false, true, 1.0f, da);
}

/* ---------------------------------------------------------------- weight server pieces
 * xavierInitializer, weightserver.cpp:567-585: default_random_engine(8888) re-seeded per matrix. */
void orc_xavier(unsigned dim1, unsigned dim2, float *dptr) {
    std::default_random_engine dre(8888);
    std::uniform_real_distribution<float> dist(-1, 1);
    unsigned dataSize = dim1 * dim2;
    for (unsigned ui = 0; ui < dataSize; ++ui) dptr[ui] = dist(dre);
    float normFactor = std::sqrt(6.0 / (float(dim1 + dim2)));
    for (unsigned ui = 0; ui < dataSize; ++ui) dptr[ui] *= normFactor;
}
/* kaimingInitializer, weightserver.cpp:593-612. */
void orc_kaiming(unsigned dim1, unsigned dim2, float *dptr) {
    std::default_random_engine dre(8888);
    std::normal_distribution<float> dist(0, 1);
    unsigned dataSize = dim1 * dim2;
    for (unsigned ui = 0; ui < dataSize; ++ui) dptr[ui] = dist(dre);
    float normFactor = std::sqrt(2.0 / (float(dim1)));
    for (unsigned ui = 0; ui < dataSize; ++ui) dptr[ui] *= normFactor;
}

/* AdamOptimizer, AdamOptimizer.cpp:3-51 / AdamOptimizer.hpp:69-75. */
struct OrcAdam {
    float BETA1 = .9, BETA2 = .999, EPSILON = 1e-07;
    const float WEIGHT_DECAY = 0;
    float learning_rate, lr_t;
    unsigned epochs;
    std::vector<unsigned> dims;
    std::vector<std::vector<FeatType>> momentum, decay;
    void nextIteration() {
        ++epochs;
        float beta_1_power = pow(BETA1, epochs);
        float beta_2_power = pow(BETA2, epochs);
        lr_t = learning_rate * (sqrt(1 - beta_2_power)) / (1 - beta_1_power);
    }
};
void *orc_adam_create(float lr, const unsigned *dims, unsigned ndims) {
    OrcAdam *a = new OrcAdam();
    a->learning_rate = lr;
    a->dims.assign(dims, dims + ndims);
    a->epochs = 0;
    a->lr_t = 0;
    a->nextIteration();
    for (unsigned ui = 0; ui + 1 < ndims; ++ui) {
        a->momentum.emplace_back((size_t)dims[ui] * dims[ui + 1], 0.0f);
        a->decay.emplace_back((size_t)dims[ui] * dims[ui + 1], 0.0f);
    }
    return a;
}
void orc_adam_update(void *h, unsigned layer, float *weight, const float *gradient) {
    OrcAdam &o = *static_cast<OrcAdam *>(h);
    unsigned size = o.dims[layer] * o.dims[layer + 1];
    for (unsigned i = 0; i < size; ++i) {
        float gt = gradient[i] + o.WEIGHT_DECAY * weight[i];
        float prev_m = o.momentum[layer][i];
        float prev_d = o.decay[layer][i];
        o.momentum[layer][i] = o.BETA1 * prev_m + (1. - o.BETA1) * gt;
        o.decay[layer][i] = o.BETA2 * prev_d + (1. - o.BETA2) * gt * gt;
        float delta = o.lr_t * (o.momentum[layer][i]) / (sqrt(o.decay[layer][i]) + o.EPSILON);
        weight[i] -= delta;
    }
    if (layer == 0) o.nextIteration();
}
float orc_adam_lr_t(void *h) { return static_cast<OrcAdam *>(h)->lr_t; }
void orc_adam_destroy(void *h) { delete static_cast<OrcAdam *>(h); }

}  // extern "C"
