// TEST INFRASTRUCTURE ONLY.  The reference's own object code for the hot path, callable from Python:
//
//   Engine::preallocateGCN / aggregateGCN      src/graph-server/engine/ops/gcn_ops.cpp   (compiled in place)
//   Engine::preallocateGAT / aggregateGAT / predictGAT / applyVertexGAT / applyEdgeGAT
//                                              src/graph-server/engine/ops/gat_ops.cpp   (compiled in place)
//   CPUComm::NNCompute -> vtxNNForwardGCN / vtxNNBackwardGCN, vtxNN*GAT / edgNN*GAT and every helper they use
//   (activate, softmax, getTrainStat, maskout, hadamardSub, activateDerivative)
//                                              src/graph-server/commmanager/CPU_comm.cpp (compiled in place)
//   Graph::init, Matrix::dot                   graph/graph.cpp, common/matrix.cpp        (compiled in place)
//
// Those two translation units include <zmq.hpp> and Boost headers, which this image does not have: they are
// compiled against the stub headers of oracle/shim/ (types only; any transport call aborts).  What the rest
// of the graph server would supply at link time is supplied here instead -- none of it is arithmetic of the path:
//   * MessageService (commmanager/message_service.cpp: the ZeroMQ client of the weight server) becomes an
//     in-process endpoint: getWeightMatrix hands out the weights the caller installed, sendWeightUpdate /
//     sendAccloss keep what the reference would have sent.
//   * Engine::srcVFeats2eFeats / dstVFeats2eFeats / incLayerGCN (engine/utils.cpp:655-732; that file needs
//     boost::program_options for the command line) are restated below, line for line.
//   * ResourceComm::NNRecvCallback (hands the chunk to the next pipeline queue), CommManager::dataPushOut /
//     dataPullIn and Engine::verticesPushOut (ZeroMQ ghost exchange) are never reached by the calls made here.
// The reference is built with _CPU_ENABLED_ (CMakeLists.txt:23), -O3 -march=native -fopenmp like its CPU backend.
#include <omp.h>

#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "graph-server/commmanager/CPU_comm.hpp"
#include "graph-server/engine/engine.hpp"

// ------------------------------------------------------------------ in-process weight endpoint
namespace {
struct Endpoint {
    std::vector<std::vector<float>> w, dw, a, da;  // "w" and (GAT) "a_i" per layer, and what was sent back
    std::vector<unsigned> rows, cols;
    float acc = 0.f, loss = 0.f;
    unsigned updates = 0;
} g_ep;
}  // namespace

MessageService::MessageService(unsigned wPort_, unsigned nodeId_, unsigned numLayers_, GNN gnn_type_)
    : wctx(1), wsocket(wctx, ZMQ_DEALER), nodeId(nodeId_), wPort(wPort_), wsocktReady(false), gnn_type(gnn_type_),
      epoch(-1), numLayers(numLayers_) {
    weights.resize(numLayers);
    as.resize(numLayers);
}
void MessageService::setUpWeightSocket(char *) {}
void MessageService::prefetchWeightsMatrix() {  // message_service.cpp:188-222 without the wire
    epoch++;
    for (unsigned j = 0; j < numLayers && j < g_ep.w.size(); ++j) {
        weights[j] = Matrix(g_ep.rows[j], g_ep.cols[j], g_ep.w[j].data());
        if (j < g_ep.a.size() && !g_ep.a[j].empty()) as[j] = Matrix(g_ep.cols[j], 1, g_ep.a[j].data());
    }
}
Matrix MessageService::getWeightMatrix(unsigned layer) { return weights.at(layer); }
void MessageService::sendWeightUpdate(Matrix &matrix, unsigned layer) {  // :148-162: the sender owns and frees it
    g_ep.dw.at(layer).assign(matrix.getData(), matrix.getData() + matrix.getNumElemts());
    g_ep.updates++;
    deleteMatrix(matrix);
}
Matrix MessageService::getaMatrix(unsigned layer) { return as.at(layer); }
void MessageService::sendaUpdate(Matrix &matrix, unsigned layer) {
    g_ep.da.at(layer).assign(matrix.getData(), matrix.getData() + matrix.getNumElemts());
    deleteMatrix(matrix);
}
void MessageService::sendAccloss(float acc, float loss, unsigned) {
    g_ep.acc = acc;
    g_ep.loss = loss;
}

// ------------------------------------------------------------------ pieces of the graph server not on the path
void ResourceComm::NNRecvCallback(Engine *, Chunk &) {}
void CommManager::dataPushOut(unsigned, unsigned, unsigned, void *, unsigned) { std::abort(); }
bool CommManager::dataPullIn(unsigned *, unsigned *, void *, unsigned) { std::abort(); }
void Engine::verticesPushOut(unsigned, unsigned, unsigned *, FeatType *, unsigned, Chunk &) { std::abort(); }

// engine/utils.cpp:655-678
FeatType **Engine::srcVFeats2eFeats(FeatType *vtcsTensor, FeatType *ghostTensor, unsigned, unsigned featDim) {
    FeatType **eVtxFeatsBuf = new FeatType *[2 * graph.localInEdgeCnt];
    FeatType **eSrcVtxFeats = eVtxFeatsBuf;
    FeatType **eDstVtxFeats = eSrcVtxFeats + graph.localInEdgeCnt;
    unsigned long long edgeItr = 0;
    for (unsigned lvid = 0; lvid < graph.localVtxCnt; ++lvid) {
        for (unsigned long long eid = graph.forwardAdj.columnPtrs[lvid]; eid < graph.forwardAdj.columnPtrs[lvid + 1]; ++eid) {
            unsigned srcVid = graph.forwardAdj.rowIdxs[eid];
            eSrcVtxFeats[edgeItr] = srcVid < graph.localVtxCnt ? getVtxFeat(vtcsTensor, srcVid, featDim)
                                                               : getVtxFeat(ghostTensor, srcVid - graph.localVtxCnt, featDim);
            eDstVtxFeats[edgeItr] = getVtxFeat(vtcsTensor, lvid, featDim);
            ++edgeItr;
        }
    }
    return eVtxFeatsBuf;
}
// engine/utils.cpp:682-705
FeatType **Engine::dstVFeats2eFeats(FeatType *vtcsTensor, FeatType *ghostTensor, unsigned, unsigned featDim) {
    FeatType **eVtxFeatsBuf = new FeatType *[2 * graph.localOutEdgeCnt];
    FeatType **eSrcVtxFeats = eVtxFeatsBuf;
    FeatType **eDstVtxFeats = eSrcVtxFeats + graph.localOutEdgeCnt;
    unsigned long long edgeItr = 0;
    for (unsigned lvid = 0; lvid < graph.localVtxCnt; ++lvid) {
        for (unsigned long long eid = graph.backwardAdj.rowPtrs[lvid]; eid < graph.backwardAdj.rowPtrs[lvid + 1]; ++eid) {
            unsigned srcVid = graph.backwardAdj.columnIdxs[eid];
            eSrcVtxFeats[edgeItr] = srcVid < graph.localVtxCnt ? getVtxFeat(vtcsTensor, srcVid, featDim)
                                                               : getVtxFeat(ghostTensor, srcVid - graph.localVtxCnt, featDim);
            eDstVtxFeats[edgeItr] = getVtxFeat(vtcsTensor, lvid, featDim);
            ++edgeItr;
        }
    }
    return eVtxFeatsBuf;
}
// engine/utils.cpp:714-732
Chunk Engine::incLayerGCN(const Chunk &chunk) {
    Chunk nChunk = chunk;
    if (nChunk.dir == PROP_TYPE::FORWARD) {
        nChunk.layer++;
        if (nChunk.layer == numLayers) {
            nChunk.dir = PROP_TYPE::BACKWARD;
            nChunk.layer--;
        }
    } else if (nChunk.layer == 0) {
        nChunk.dir = PROP_TYPE::FORWARD;
        nChunk.epoch++;
    } else {
        nChunk.layer--;
    }
    return nChunk;
}

// engine/utils.cpp:734-748
Chunk Engine::incLayerGAT(const Chunk &chunk) {
    Chunk nChunk = chunk;
    if (nChunk.dir == PROP_TYPE::FORWARD) {
        nChunk.layer++;
    } else if (nChunk.layer == 0) {
        nChunk.dir = PROP_TYPE::FORWARD;
        nChunk.vertex = true;
        nChunk.epoch++;
    } else {
        nChunk.layer--;
    }
    return nChunk;
}

// ------------------------------------------------------------------ C interface for oracle/pyoracle.py
namespace {
struct RefEngine {
    Engine eng;
    CPUComm *comm = nullptr;
};
}  // namespace

extern "C" {

// graph_file: a graph.<id>.bin on disk (Graph::init reads it); dims[0..n_layers]; ws_file: a text file with one
// address line (CPUComm's constructor reads the weight-server list, CPU_comm.cpp:244-263).
void *refeng_create(const char *graph_file, const unsigned *dims, unsigned n_layers, const char *ws_file, int gat) {
    RefEngine *r = new RefEngine();
    Engine &e = r->eng;
    e.graph.init(std::string(graph_file));
    e.gnn_type = gat ? GNN::GAT : GNN::GCN;
    e.numLayers = n_layers;
    e.layerConfig.assign(dims, dims + n_layers + 1);
    e.nodeId = 0;
    e.numNodes = 1;
    e.weightserverPort = 0;
    e.weightserverIPFile = ws_file;
    e.forwardVerticesInitData = new FeatType[(size_t)e.getFeatDim(0) * e.graph.localVtxCnt]();
    e.forwardGhostInitData = new FeatType[(size_t)e.getFeatDim(0) * (e.graph.srcGhostCnt + 1)]();
    e.localVerticesLabels = new FeatType[(size_t)e.getFeatDim(n_layers) * e.graph.localVtxCnt]();
    e.savedNNTensors.resize(n_layers);   // engine.cpp:113-114
    e.savedEdgeTensors.resize(n_layers);
    if (gat) {
        e.preallocateGAT();
        // quirk Q11: the reference accumulates the backward aggregation into "aTg" as operator new[] left it
        // (gat_ops.cpp:103-104,221-241); like the oracle, start from zero
        for (unsigned l = 0; l < n_layers; ++l) {
            Matrix &m = e.savedNNTensors[l]["aTg"];
            std::memset(m.getData(), 0, m.getDataSize());
        }
    } else {
        e.preallocateGCN();              // the reference's own allocation + pointer tables
    }
    g_ep = Endpoint();
    g_ep.w.resize(n_layers);
    g_ep.dw.resize(n_layers);
    g_ep.a.resize(n_layers);
    g_ep.da.resize(n_layers);
    if (gat)
        for (unsigned l = 0; l < n_layers; ++l) g_ep.a[l].assign(dims[l + 1], 0.f);
    g_ep.rows.assign(dims, dims + n_layers);
    g_ep.cols.assign(dims + 1, dims + n_layers + 1);
    for (unsigned l = 0; l < n_layers; ++l) g_ep.w[l].assign((size_t)dims[l] * dims[l + 1], 0.f);
    r->comm = new CPUComm(&e);
    e.resComm = r->comm;
    return r;
}

float *refeng_tensor(void *h, unsigned layer, const char *name, unsigned *rows, unsigned *cols) {
    Engine &e = static_cast<RefEngine *>(h)->eng;
    if (layer >= e.savedNNTensors.size() || !e.savedNNTensors[layer].count(name)) return nullptr;
    Matrix &m = e.savedNNTensors[layer][name];
    *rows = m.getRows();
    *cols = m.getCols();
    return m.getData();
}

void refeng_set_weights(void *, unsigned layer, const float *w) {
    std::memcpy(g_ep.w.at(layer).data(), w, g_ep.w[layer].size() * sizeof(float));
}
int refeng_get_update(void *, unsigned layer, float *dw) {
    if (g_ep.dw.at(layer).empty()) return -1;
    std::memcpy(dw, g_ep.dw[layer].data(), g_ep.dw[layer].size() * sizeof(float));
    return 0;
}
void refeng_set_a(void *, unsigned layer, const float *a) {
    std::memcpy(g_ep.a.at(layer).data(), a, g_ep.a[layer].size() * sizeof(float));
}
int refeng_get_a_update(void *, unsigned layer, float *da) {
    if (g_ep.da.at(layer).empty()) return -1;
    std::memcpy(da, g_ep.da[layer].data(), g_ep.da[layer].size() * sizeof(float));
    return 0;
}
// The GAT operators of the reference on a whole-partition chunk with chunk.layer = `layer`:
// op 0 aggregateGAT, 1 applyVertexGAT, 2 applyEdgeGAT, 3 predictGAT (gat_ops.cpp:173-265,267-275,437-440)
void refeng_gat_op(void *h, int op, unsigned layer, int dir) {
    Engine &e = static_cast<RefEngine *>(h)->eng;
    Chunk c{0, e.nodeId, 0, e.graph.localVtxCnt, layer, dir == 0 ? PROP_TYPE::FORWARD : PROP_TYPE::BACKWARD, 1, op != 2};
    if (op == 0) e.aggregateGAT(c);
    else if (op == 1) e.applyVertexGAT(c);
    else if (op == 2) e.applyEdgeGAT(c);
    else e.predictGAT(c);
}
void refeng_stats(void *, float *acc, float *loss) {
    *acc = g_ep.acc;
    *loss = g_ep.loss;
}

static Chunk whole(Engine &e, unsigned layer, int dir, bool vertex) {
    return Chunk{0, e.nodeId, 0, e.graph.localVtxCnt, layer, dir == 0 ? PROP_TYPE::FORWARD : PROP_TYPE::BACKWARD, 1, vertex};
}
// Engine::aggregateGCN on a whole-partition chunk (or [low, up))
void refeng_aggregate(void *h, unsigned layer, int dir, unsigned low, unsigned up) {
    Engine &e = static_cast<RefEngine *>(h)->eng;
    Chunk c = whole(e, layer, dir, true);
    c.lowBound = low;
    c.upBound = up;
    e.aggregateGCN(c);
}
// Engine::applyVertexGCN -> CPUComm::NNCompute (forward at `layer`; backward: incLayer first, gcn_ops.cpp:194-202)
void refeng_apply_vertex(void *h, unsigned layer, int dir) {
    Engine &e = static_cast<RefEngine *>(h)->eng;
    Chunk c = whole(e, layer, dir, true);
    e.applyVertexGCN(c);
}
void refeng_set_threads(int n) { omp_set_num_threads(n > 0 ? n : 1); }

}  // extern "C"
