// Host-side partition image handling: a zero-copy view over a graph.<id>.bin byte image
// (layout: reference graph/graph.cpp:200-273) and the array-based preprocessor that produces it.
#pragma once
#include <cstddef>
#include <cstdint>
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

// == DataLoader::preprocess + RawGraph::dump.  Returns "" on success.
std::string preprocess_partition(const EdgeList &edges, const int32_t *parts, uint32_t nVertices,
                                 uint32_t part, uint32_t nParts, bool undirected,
                                 std::vector<uint8_t> &image);

// == Engine::readFeaturesFile incl. the feats<F0>.<id>.bin cache / Engine::readLabelsFile
// (engine/utils.cpp:486-596) for one partition.  local: [V x F], ghost: [Gs x F], onehot: [V x kinds].
std::string read_features(const std::string &dir, const std::string &featuresFile, const PartitionView &g,
                          uint32_t nodeId, uint32_t F, float *local, float *ghost);
std::string read_labels(const std::string &labelsFile, const PartitionView &g, uint32_t kinds, float *onehot);

}  // namespace dory
