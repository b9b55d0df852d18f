// Host-side partition image handling: a zero-copy view over a graph.<id>.bin byte image
// (layout: reference graph/graph.cpp:200-273) and the array-based preprocessor that produces it.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <string>
#include <vector>

namespace dory {

// View into a graph.<id>.bin image; pointers alias the image (no copies).  Multi-byte fields in the
// image are not naturally aligned in general, so consumers memcpy / cudaMemcpy out of it.
struct PartitionView {
    uint32_t localVtxCnt = 0, globalVtxCnt = 0, srcGhostCnt = 0, dstGhostCnt = 0;
    uint64_t localInEdgeCnt = 0, localOutEdgeCnt = 0, globalEdgeCnt = 0;
    const uint8_t *localToGlobal = nullptr;  // u32[V]
    const uint8_t *norms = nullptr;          // f32[V]
    const uint8_t *srcGhostPairs = nullptr;  // (gvid u32, lvid u32)[Gs]
    const uint8_t *dstGhostPairs = nullptr;  // (gvid u32, lvid u32)[Gd]
    uint32_t numNodes = 0;
    std::vector<std::pair<const uint8_t *, uint32_t>> fwdSend, bwdSend;  // per peer: u32 ids, count
    uint64_t fwdNnz = 0, bwdNnz = 0;
    const uint8_t *fwdVals = nullptr, *colPtrs = nullptr, *rowIdxs = nullptr;  // CSC
    const uint8_t *bwdVals = nullptr, *rowPtrs = nullptr, *colIdxs = nullptr;  // CSR
};

// Returns "" on success, else an error message.
std::string parse_partition(const void *image, size_t len, PartitionView &out);

// Edge list accessor: either two separate arrays or interleaved (src,dst) pairs (the bsnap body).
struct EdgeList {
    const uint32_t *src = nullptr;
    const uint32_t *dst = nullptr;
    size_t stride = 1;  // 1: separate arrays, 2: interleaved pairs (dst == src + 1)
    uint64_t n = 0;
};

// A malloc'ed, UNINITIALISED byte image (a multi-GB image is written exactly once by the preprocessor;
// zero-filling it first and copying it out afterwards cost 12 s of a 45 s Friendster/8 partition).
struct HostImage {
    uint8_t *p = nullptr;
    size_t n = 0;
    HostImage() = default;
    HostImage(const HostImage &) = delete;
    HostImage &operator=(const HostImage &) = delete;
    ~HostImage() { std::free(p); }
    bool alloc(size_t bytes) {
        std::free(p);
        p = static_cast<uint8_t *>(std::malloc(bytes ? bytes : 1));
        n = p ? bytes : 0;
        return p != nullptr;
    }
    uint8_t *data() { return p; }
    size_t size() const { return n; }
    uint8_t *release() {  // the caller frees with free() (dory_free)
        uint8_t *r = p;
        p = nullptr;
        n = 0;
        return r;
    }
};

// == DataLoader::preprocess + RawGraph::dump.  Returns "" on success.
// `inDegree` / `globalEdges` (optional): for callers that hold only the edge records INCIDENT to the
// partition (either endpoint owned by `part`) instead of the whole edge file -- every in- and out-edge
// of a local vertex is among those, but the raw in-degree of a ghost vertex (findGhostDegrees re-reads
// the whole file for it, graph/dataloader.cpp:192-218) and the global edge count are not.  inDegree[g]
// = in-degree of global vertex g in the whole graph (self loops excluded), globalEdges = records in
// the whole graph.  With both null / 0 the list is taken to be the whole edge file.
std::string preprocess_partition(const EdgeList &edges, const int32_t *parts, uint32_t nVertices,
                                 uint32_t part, uint32_t nParts, bool undirected,
                                 HostImage &image, const uint32_t *inDegree = nullptr,
                                 uint64_t globalEdges = 0);

// == Engine::readFeaturesFile incl. the feats<F0>.<id>.bin cache / Engine::readLabelsFile
// (engine/utils.cpp:486-596) for one partition.  local: [V x F], ghost: [Gs x F], onehot: [V x kinds].
std::string read_features(const std::string &dir, const std::string &featuresFile, const PartitionView &g,
                          uint32_t nodeId, uint32_t F, float *local, float *ghost);
std::string read_labels(const std::string &labelsFile, const PartitionView &g, uint32_t kinds, float *onehot);

}  // namespace dory
