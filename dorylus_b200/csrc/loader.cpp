// Partition preprocessor and graph.<id>.bin parser (host only; no CUDA).
//
// Behavioural contract = the reference's DataLoader::preprocess (graph/dataloader.cpp:225-330),
// findGhostDegrees (:192-218), setEdgeNormalizations (:153-185) and RawGraph::dump
// (graph/graph.cpp:200-273); the output image is byte-identical to the reference's
// graph.<id>.bin (tests/test_loader.py diffs it against the compiled reference loader).
//
// The reference builds per-vertex objects with std::map lookups per edge; here the same result is
// produced with two streaming passes over the edge list and flat arrays (counting sort keyed by
// local destination / source), so a 114 M-edge partition is built in seconds and the memory
// high-water mark is the image itself plus O(V_global) integers.
#include "loader.h"

#include <fstream>

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <limits>
#include <chrono>
#include <cstdio>
#include <thread>

namespace dory {
namespace {

// DORY_LOADER_VERBOSE=1: phase timings of preprocess_partition on stderr
struct PhaseTimer {
    bool on = std::getenv("DORY_LOADER_VERBOSE") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void mark(const char *what) {
        if (!on) return;
        auto n = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[loader] %-28s %.2f s\n", what, std::chrono::duration<double>(n - t).count());
        t = n;
    }
};

constexpr uint32_t kNone = std::numeric_limits<uint32_t>::max();

template <class T>
T rd(const uint8_t *p) {
    T v;
    std::memcpy(&v, p, sizeof(T));
    return v;
}

struct Cursor {
    const uint8_t *base;
    size_t len, off = 0;
    bool ok = true;
    const uint8_t *take(size_t bytes) {
        if (!ok || bytes > len - off) {
            ok = false;
            return base;
        }
        const uint8_t *p = base + off;
        off += bytes;
        return p;
    }
    template <class T>
    T get() {
        const uint8_t *p = take(sizeof(T));
        return ok ? rd<T>(p) : T();
    }
};

struct Writer {
    uint8_t *p;
    template <class T>
    void put(const T &v) {
        std::memcpy(p, &v, sizeof(T));
        p += sizeof(T);
    }
    void bytes(const void *src, size_t n) {
        if (n) std::memcpy(p, src, n);
        p += n;
    }
};

// (in-degree + 1)^-1/2 evaluated like the reference: pow in double, narrowed to float
// (dataloader.cpp:155-156: `float vtxNorm = std::pow(vtxDeg, -.5)` with an unsigned degree).
inline float inv_sqrt_deg(uint32_t degPlusOne) { return (float)std::pow((double)degPlusOne, -.5); }

unsigned loader_threads(uint64_t nEdges) {
    if (const char *s = std::getenv("DORY_LOADER_THREADS")) {
        int n = std::atoi(s);
        if (n > 0) return (unsigned)n;
    }
    if (nEdges < (1u << 20)) return 1;
    unsigned hw = std::thread::hardware_concurrency();
    return std::max(1u, std::min(hw ? hw : 1u, 32u));
}

void run_threads(unsigned n, const std::function<void(unsigned)> &fn) {
    if (n <= 1) {
        fn(0);
        return;
    }
    std::vector<std::thread> th;
    for (unsigned t = 0; t < n; ++t) th.emplace_back(fn, t);
    for (auto &x : th) x.join();
}

}  // namespace

std::string parse_partition(const void *image, size_t len, PartitionView &g) {
    Cursor c{static_cast<const uint8_t *>(image), len};
    g.localVtxCnt = c.get<uint32_t>();
    g.globalVtxCnt = c.get<uint32_t>();
    g.srcGhostCnt = c.get<uint32_t>();
    g.dstGhostCnt = c.get<uint32_t>();
    g.localInEdgeCnt = c.get<uint64_t>();
    g.localOutEdgeCnt = c.get<uint64_t>();
    g.globalEdgeCnt = c.get<uint64_t>();
    if (!c.ok) return "graph image truncated in header";
    const size_t V = g.localVtxCnt;
    g.localToGlobal = c.take(4 * V);
    g.norms = c.take(4 * V);
    g.srcGhostPairs = c.take(8 * (size_t)g.srcGhostCnt);
    g.dstGhostPairs = c.take(8 * (size_t)g.dstGhostCnt);
    g.numNodes = c.get<uint32_t>();
    if (!c.ok) return "graph image truncated in vertex section";
    if (g.numNodes > (1u << 20)) return "graph image: implausible numNodes";
    g.fwdSend.clear();
    g.bwdSend.clear();
    for (int dir = 0; dir < 2; ++dir) {
        for (uint32_t i = 0; i < g.numNodes; ++i) {
            uint32_t sz = c.get<uint32_t>();
            const uint8_t *p = c.take(4 * (size_t)sz);
            if (!c.ok) return "graph image truncated in send lists";
            (dir == 0 ? g.fwdSend : g.bwdSend).emplace_back(p, sz);
        }
    }
    uint32_t ccnt = c.get<uint32_t>();
    g.fwdNnz = c.get<uint64_t>();
    if (!c.ok) return "graph image truncated before CSC";
    if (g.fwdNnz > len) return "graph image: CSC nnz exceeds image size";
    g.fwdVals = c.take(4 * g.fwdNnz);
    g.colPtrs = c.take(8 * (V + 1));
    g.rowIdxs = c.take(4 * g.fwdNnz);
    uint32_t rcnt = c.get<uint32_t>();
    g.bwdNnz = c.get<uint64_t>();
    if (!c.ok) return "graph image truncated before CSR";
    if (g.bwdNnz > len) return "graph image: CSR nnz exceeds image size";
    g.bwdVals = c.take(4 * g.bwdNnz);
    g.rowPtrs = c.take(8 * (V + 1));
    g.colIdxs = c.take(4 * g.bwdNnz);
    if (!c.ok) return "graph image truncated in CSR";
    if (c.off != len) return "graph image has trailing bytes";
    if (ccnt != g.localVtxCnt || rcnt != g.localVtxCnt) return "graph image: CSC/CSR vertex count mismatch";
    if (rd<uint64_t>(g.colPtrs + 8 * V) != g.fwdNnz || rd<uint64_t>(g.rowPtrs + 8 * V) != g.bwdNnz)
        return "graph image: offsets do not end at nnz";
    return "";
}

std::string preprocess_partition(const EdgeList &el, const int32_t *parts, uint32_t nV, uint32_t me,
                                 uint32_t nParts, bool undirected, HostImage &image,
                                 const uint32_t *inDegree, uint64_t globalEdgesGiven) {
    if (me >= nParts) return "part id out of range";
    if (inDegree && undirected) return "incident-edge lists carry both directions explicitly (undirected = 0)";
    PhaseTimer pt_;
    // ---- local ids: order of appearance in the parts file (dataloader.cpp:66-83)
    std::vector<uint32_t> g2l(nV, kNone), l2g;
    for (uint32_t g = 0; g < nV; ++g) {
        if (parts[g] < 0 || (uint32_t)parts[g] >= nParts) return "partition id out of range in parts";
        if ((uint32_t)parts[g] == me) {
            g2l[g] = (uint32_t)l2g.size();
            l2g.push_back(g);
        }
    }
    const uint32_t V = (uint32_t)l2g.size();
    pt_.mark("local ids");

    // ---- pass 1: degrees, ghost discovery, boundary flags
    std::vector<uint64_t> inPtr(V + 1, 0), outPtr(V + 1, 0);
    // findGhostDegrees: raw records only, keyed by dst (or the caller's whole-graph degrees)
    std::vector<uint32_t> rawInDegOwn(inDegree ? 0 : nV, 0);
    uint32_t *rawInDegW = inDegree ? nullptr : rawInDegOwn.data();
    const uint32_t *rawInDeg = inDegree ? inDegree : rawInDegOwn.data();
    std::vector<uint32_t> srcGhostSlot(nV, kNone), dstGhostSlot(nV, kNone);
    std::vector<std::vector<uint8_t>> fwdFlag(nParts), bwdFlag(nParts);
    for (uint32_t p = 0; p < nParts; ++p)
        if (p != me) {
            fwdFlag[p].assign(V, 0);
            bwdFlag[p].assign(V, 0);
        }
    // Every update is keyed by one endpoint's global id; thread t applies exactly the updates whose
    // key falls in its id range [lo, hi), so the threads stream the whole edge list but never write
    // the same word.
    const unsigned nThreads = loader_threads(el.n);
    std::atomic<bool> bad{false};
    std::vector<uint64_t> edgeCount(nThreads, 0);
    run_threads(nThreads, [&](unsigned t) {
        const uint32_t lo = (uint32_t)((uint64_t)nV * t / nThreads), hi = (uint32_t)((uint64_t)nV * (t + 1) / nThreads);
        auto own = [&](uint32_t g) { return g >= lo && g < hi; };
        auto visit = [&](uint32_t from, uint32_t to) {  // processEdge, dataloader.cpp:94-146 (counting)
            if (!own(from) && !own(to)) return;  // before the two random reads of parts[]: most edges are not this thread's
            const uint32_t pf = (uint32_t)parts[from], pt = (uint32_t)parts[to];
            if (pf == me) {
                if (own(from)) {
                    const uint32_t lf = g2l[from];
                    ++outPtr[lf + 1];
                    if (pt != me) fwdFlag[pt][lf] = 1;
                }
                if (pt != me && own(to)) dstGhostSlot[to] = 0;  // discovered
            }
            if (pt == me) {
                if (own(to)) {
                    const uint32_t lt = g2l[to];
                    ++inPtr[lt + 1];
                    if (pf != me) bwdFlag[pf][lt] = 1;
                }
                if (pf != me && own(from)) srcGhostSlot[from] = 0;
            }
        };
        uint64_t cnt = 0;
        for (uint64_t i = 0; i < el.n; ++i) {
            const uint32_t s = el.src[i * el.stride], d = el.dst[i * el.stride];
            if (s >= nV || d >= nV) {
                bad = true;
                return;
            }
            if (s == d) continue;  // dataloader.cpp:268-269
            if (rawInDegW && own(d)) ++rawInDegW[d];
            visit(s, d);
            if (undirected) visit(d, s);
            ++cnt;
        }
        edgeCount[t] = cnt;
    });
    if (bad) return "edge endpoint out of range";
    pt_.mark("pass 1 (count, discover)");
    const uint64_t globalEdges = globalEdgesGiven ? globalEdgesGiven : edgeCount[0];
    for (uint32_t v = 0; v < V; ++v) {
        inPtr[v + 1] += inPtr[v];
        outPtr[v + 1] += outPtr[v];
    }
    const uint64_t nIn = inPtr[V], nOut = outPtr[V];

    // ---- ghost slots: ascending global id, starting at V (dataloader.cpp:311-322)
    std::vector<uint32_t> srcGhosts, dstGhosts;
    for (uint32_t g = 0; g < nV; ++g) {
        if (srcGhostSlot[g] != kNone) {
            srcGhostSlot[g] = V + (uint32_t)srcGhosts.size();
            srcGhosts.push_back(g);
        }
        if (dstGhostSlot[g] != kNone) {
            dstGhostSlot[g] = V + (uint32_t)dstGhosts.size();
            dstGhosts.push_back(g);
        }
    }

    pt_.mark("ghost slots");
    // ---- per-vertex (in-degree + 1)^-1/2.  Local: edges held (incl. undirected expansion);
    // ghost: raw-file in-degree (quirk Q3).
    std::vector<float> locNorm(V);
    for (uint32_t v = 0; v < V; ++v) locNorm[v] = inv_sqrt_deg((uint32_t)(inPtr[v + 1] - inPtr[v]) + 1);
    if (inDegree)  // the caller's degrees must agree with the in-edges it handed over for the local rows
        for (uint32_t v = 0; v < V; ++v)
            if (inDegree[l2g[v]] != inPtr[v + 1] - inPtr[v])
                return "inDegree[" + std::to_string(l2g[v]) + "] disagrees with the in-edges of that local vertex";

    // ---- image layout
    size_t bytes = 4 * 4 + 3 * 8 + 4 * (size_t)V * 2 + 8 * (srcGhosts.size() + dstGhosts.size()) + 4;
    std::vector<std::vector<uint32_t>> fwdList(nParts), bwdList(nParts);
    for (uint32_t p = 0; p < nParts; ++p) {
        if (p != me)
            for (uint32_t v = 0; v < V; ++v) {
                if (fwdFlag[p][v]) fwdList[p].push_back(v);
                if (bwdFlag[p][v]) bwdList[p].push_back(v);
            }
        bytes += 8 + 4 * (fwdList[p].size() + bwdList[p].size());
    }
    const size_t cscOff = bytes;
    bytes += 4 + 8 + 4 * nIn + 8 * ((size_t)V + 1) + 4 * nIn;
    const size_t csrOff = bytes;
    bytes += 4 + 8 + 4 * nOut + 8 * ((size_t)V + 1) + 4 * nOut;
    pt_.mark("norms, send lists");
    if (!image.alloc(bytes)) return "out of host memory for a " + std::to_string(bytes) + "-byte image";
    pt_.mark("image alloc");

    Writer w{image.data()};
    w.put(V);
    w.put(nV);
    w.put((uint32_t)srcGhosts.size());
    w.put((uint32_t)dstGhosts.size());
    w.put(nIn);
    w.put(nOut);
    w.put(globalEdges);
    w.bytes(l2g.data(), 4 * (size_t)V);
    for (uint32_t v = 0; v < V; ++v) w.put(locNorm[v] * locNorm[v]);  // dataloader.cpp:157
    for (uint32_t g : srcGhosts) {
        w.put(g);
        w.put(srcGhostSlot[g]);
    }
    for (uint32_t g : dstGhosts) {
        w.put(g);
        w.put(dstGhostSlot[g]);
    }
    w.put(nParts);
    for (uint32_t p = 0; p < nParts; ++p) {
        w.put((uint32_t)fwdList[p].size());
        w.bytes(fwdList[p].data(), 4 * fwdList[p].size());
    }
    for (uint32_t p = 0; p < nParts; ++p) {
        w.put((uint32_t)bwdList[p].size());
        w.bytes(bwdList[p].data(), 4 * bwdList[p].size());
    }
    // CSC header + offsets, CSR header + offsets; values / indices are filled in pass 2.
    uint8_t *cscVals = image.data() + cscOff + 12;
    uint8_t *cscPtrs = cscVals + 4 * nIn;
    uint8_t *cscIdx = cscPtrs + 8 * ((size_t)V + 1);
    uint8_t *csrVals = image.data() + csrOff + 12;
    uint8_t *csrPtrs = csrVals + 4 * nOut;
    uint8_t *csrIdx = csrPtrs + 8 * ((size_t)V + 1);
    {
        Writer h{image.data() + cscOff};
        h.put(V);
        h.put(nIn);
        std::memcpy(cscPtrs, inPtr.data(), 8 * ((size_t)V + 1));
        Writer h2{image.data() + csrOff};
        h2.put(V);
        h2.put(nOut);
        std::memcpy(csrPtrs, outPtr.data(), 8 * ((size_t)V + 1));
    }

    pt_.mark("image header");
    // ---- pass 2: place every in-/out-edge at its slot (insertion order == edge-file order, Q4)
    std::vector<uint64_t> inCur(inPtr.begin(), inPtr.end() - 1), outCur(outPtr.begin(), outPtr.end() - 1);
    run_threads(nThreads, [&](unsigned t) {
        const uint32_t lo = (uint32_t)((uint64_t)nV * t / nThreads), hi = (uint32_t)((uint64_t)nV * (t + 1) / nThreads);
        auto own = [&](uint32_t g) { return g >= lo && g < hi; };
        auto place = [&](uint32_t from, uint32_t to) {
            if (!own(from) && !own(to)) return;
            const uint32_t pf = (uint32_t)parts[from], pt = (uint32_t)parts[to];
            if (pf == me && own(from)) {
                const uint32_t lf = g2l[from];
                uint32_t id;
                float dn;
                if (pt == me) {
                    id = g2l[to];
                    dn = locNorm[id];
                } else {
                    id = dstGhostSlot[to];
                    dn = inv_sqrt_deg(rawInDeg[to] + 1);
                }
                const uint64_t k = outCur[lf]++;
                const float val = locNorm[lf] * dn;  // dataloader.cpp:177,181
                std::memcpy(csrIdx + 4 * k, &id, 4);
                std::memcpy(csrVals + 4 * k, &val, 4);
            }
            if (pt == me && own(to)) {
                const uint32_t lt = g2l[to];
                uint32_t id;
                float sn;
                if (pf == me) {
                    id = g2l[from];
                    sn = locNorm[id];
                } else {
                    id = srcGhostSlot[from];
                    sn = inv_sqrt_deg(rawInDeg[from] + 1);
                }
                const uint64_t k = inCur[lt]++;
                const float val = sn * locNorm[lt];  // dataloader.cpp:164,168
                std::memcpy(cscIdx + 4 * k, &id, 4);
                std::memcpy(cscVals + 4 * k, &val, 4);
            }
        };
        for (uint64_t i = 0; i < el.n; ++i) {
            const uint32_t s = el.src[i * el.stride], d = el.dst[i * el.stride];
            if (s == d) continue;
            place(s, d);
            if (undirected) place(d, s);
        }
    });
    pt_.mark("pass 2 (place)");
    return "";
}

// == Engine::readFeaturesFile (engine/utils.cpp:486-552) for the partition in `g`: the per-partition
// cache <dir>feats<F0>.<id>.bin (local rows, then source-ghost rows) wins when it exists; otherwise
// the global features file is streamed once (header {uint32 numFeatures}, rows in global-vertex
// order), rows of local vertices and of source ghosts are picked out, and the cache is written.
std::string read_features(const std::string &dir, const std::string &featuresFile, const PartitionView &g,
                          uint32_t nodeId, uint32_t F, float *local, float *ghost) {
    const size_t V = g.localVtxCnt, Gs = g.srcGhostCnt;
    const std::string cache = dir + "feats" + std::to_string(F) + "." + std::to_string(nodeId) + ".bin";
    {
        std::ifstream in(cache, std::ios::binary);
        if (in.good()) {
            in.read(reinterpret_cast<char *>(local), (std::streamsize)(sizeof(float) * V * F));
            if (Gs) in.read(reinterpret_cast<char *>(ghost), (std::streamsize)(sizeof(float) * Gs * F));
            if (!in.good()) return "feature cache " + cache + " is shorter than the partition needs";
            return "";
        }
    }
    std::ifstream in(featuresFile, std::ios::binary);
    if (!in.good()) return "cannot open features file " + featuresFile;
    uint32_t nf = 0;
    in.read(reinterpret_cast<char *>(&nf), 4);
    if (!in.good() || nf != F) return "features file: numFeatures does not match the layer config";
    // where does global vertex v go?  local id, or V + ghost slot (graph.srcGhostVtcs), or nowhere
    std::vector<uint32_t> where(g.globalVtxCnt, kNone);
    for (size_t l = 0; l < V; ++l) {
        const uint32_t gv = rd<uint32_t>(g.localToGlobal + 4 * l);
        if (gv >= g.globalVtxCnt) return "graph image: local vertex id out of range";
        where[gv] = (uint32_t)l;
    }
    for (size_t k = 0; k < Gs; ++k) {
        const uint32_t gv = rd<uint32_t>(g.srcGhostPairs + 8 * k), lv = rd<uint32_t>(g.srcGhostPairs + 8 * k + 4);
        if (gv >= g.globalVtxCnt || lv < V || lv - V >= Gs) return "graph image: ghost vertex id out of range";
        where[gv] = lv;  // the ghost test comes first in the reference (utils.cpp:520); ids are disjoint anyway
    }
    std::vector<float> row(F);
    uint32_t gvid = 0;
    while (in.read(reinterpret_cast<char *>(row.data()), (std::streamsize)(sizeof(float) * F))) {
        if (gvid >= g.globalVtxCnt) return "features file has more rows than the graph has vertices";
        const uint32_t w = where[gvid];
        if (w != kNone) std::memcpy((w < V ? local + (size_t)w * F : ghost + (size_t)(w - V) * F), row.data(), sizeof(float) * F);
        ++gvid;
    }
    if (gvid != g.globalVtxCnt) return "features file has fewer rows than the graph has vertices";
    std::ofstream out(cache, std::ios::binary);  // utils.cpp:537-551 (a failure to write is only logged there)
    if (out.good()) {
        out.write(reinterpret_cast<const char *>(local), (std::streamsize)(sizeof(float) * V * F));
        if (Gs) out.write(reinterpret_cast<const char *>(ghost), (std::streamsize)(sizeof(float) * Gs * F));
    }
    return "";
}

// == Engine::readLabelsFile (engine/utils.cpp:559-596): header {uint32 labelKinds}, one uint32 per
// global vertex; local vertices get a one-hot row of `kinds` floats.
std::string read_labels(const std::string &labelsFile, const PartitionView &g, uint32_t kinds, float *onehot) {
    std::ifstream in(labelsFile, std::ios::binary);
    if (!in.good()) return "cannot open labels file " + labelsFile;
    uint32_t k = 0;
    in.read(reinterpret_cast<char *>(&k), 4);
    if (!in.good() || k != kinds) return "labels file: labelKinds does not match the layer config";
    const size_t V = g.localVtxCnt;
    std::vector<uint32_t> g2l(g.globalVtxCnt, kNone);
    for (size_t l = 0; l < V; ++l) {
        const uint32_t gv = rd<uint32_t>(g.localToGlobal + 4 * l);
        if (gv >= g.globalVtxCnt) return "graph image: local vertex id out of range";
        g2l[gv] = (uint32_t)l;
    }
    std::memset(onehot, 0, sizeof(float) * V * kinds);
    uint32_t gvid = 0, cur = 0;
    while (in.read(reinterpret_cast<char *>(&cur), 4)) {
        if (gvid >= g.globalVtxCnt) return "labels file has more entries than the graph has vertices";
        if (g2l[gvid] != kNone) {
            if (cur >= kinds) return "label " + std::to_string(cur) + " out of range at vertex " + std::to_string(gvid);
            onehot[(size_t)g2l[gvid] * kinds + cur] = 1.f;
        }
        ++gvid;
    }
    if (gvid != g.globalVtxCnt) return "labels file has fewer entries than the graph has vertices";
    return "";
}

}  // namespace dory
