// Edge-cut partitioner: the stand-in for inputs/partitioner.cpp, which hands the symmetrised graph
// to METIS_PartGraphKway with unit vertex weights and writes one partition id per line
// (inputs/partitioner.cpp:63-128).  METIS is not available here, and a multilevel k-way cut is more
// than this path needs: what the GPUs care about is (1) few ghost rows per partition -- every cut
// edge is a row shipped over NVLink per layer and a row of the [local; ghost] block to gather from --
// and (2) equal vertex AND in-edge counts, because one aggregation is as slow as its heaviest
// partition.  Restreaming linear deterministic greedy does that in O(passes * E), deterministically:
//
//   order    breadth-first from the highest-degree vertex of every component (neighbours arrive
//            close together, so early placements see their community);
//   pass 0   v goes to the partition p maximising  |N(v) in p| * (1 - load_p)  among partitions with
//            room, ties to the least loaded (Stanton & Kliot's LDG);
//   pass k   the same stream again with everybody placed: v moves when another partition scores
//            higher and has room (restreaming, Nishimura & Ugander) until a pass moves < 0.1 %.
//
// load_p = max(vertices_p / vertex_cap, degree_p / degree_cap): both balance constraints at once.
#include "partition.h"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <numeric>
#include <string>
#include <vector>

namespace dory {

namespace {

struct SymCsr {
    std::vector<uint64_t> ptr;
    std::vector<uint32_t> adj;
};

// Both directions of every record (the reference inserts from->to and to->from into per-vertex
// sets, partitioner.cpp:69-73); duplicates are kept and simply weigh more.
SymCsr symmetrise(const uint32_t *src, const uint32_t *dst, uint64_t n, uint32_t V) {
    SymCsr g;
    g.ptr.assign((size_t)V + 1, 0);
    for (uint64_t e = 0; e < n; ++e) {
        ++g.ptr[src[e] + 1];
        ++g.ptr[dst[e] + 1];
    }
    for (uint32_t v = 0; v < V; ++v) g.ptr[v + 1] += g.ptr[v];
    g.adj.resize(g.ptr[V]);
    std::vector<uint64_t> cur(g.ptr.begin(), g.ptr.end() - 1);
    for (uint64_t e = 0; e < n; ++e) {
        g.adj[cur[src[e]]++] = dst[e];
        g.adj[cur[dst[e]]++] = src[e];
    }
    return g;
}

std::vector<uint32_t> bfs_order(const SymCsr &g, uint32_t V) {
    std::vector<uint32_t> byDeg(V);
    std::iota(byDeg.begin(), byDeg.end(), 0u);
    std::stable_sort(byDeg.begin(), byDeg.end(), [&](uint32_t a, uint32_t b) {
        return g.ptr[a + 1] - g.ptr[a] > g.ptr[b + 1] - g.ptr[b];
    });
    std::vector<uint8_t> seen(V, 0);
    std::vector<uint32_t> order;
    order.reserve(V);
    for (uint32_t root : byDeg) {
        if (seen[root]) continue;
        seen[root] = 1;
        size_t head = order.size();
        order.push_back(root);
        while (head < order.size()) {
            const uint32_t v = order[head++];
            for (uint64_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k) {
                const uint32_t u = g.adj[k];
                if (!seen[u]) {
                    seen[u] = 1;
                    order.push_back(u);
                }
            }
        }
    }
    return order;
}

}  // namespace

std::string partition_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, uint32_t V, uint32_t P,
                            uint32_t passes, int32_t *parts, uint64_t *edge_cut) {
    if (!parts || P == 0) return "partition: bad argument";
    if (n_edges && (!src || !dst)) return "partition: null edge arrays";
    for (uint64_t e = 0; e < n_edges; ++e)
        if (src[e] >= V || dst[e] >= V) return "partition: edge " + std::to_string(e) + " references a vertex >= numVertices";
    if (P == 1 || V == 0) {
        std::fill(parts, parts + V, 0);
        if (edge_cut) *edge_cut = 0;
        return "";
    }
    const SymCsr g = symmetrise(src, dst, n_edges, V);
    const std::vector<uint32_t> order = bfs_order(g, V);

    const double vcap = 1.03 * ((double)V / P) + 1.0;
    const double dcap = 1.08 * ((double)g.adj.size() / P) + 1.0;
    std::vector<double> nv(P, 0.0), nd(P, 0.0);
    std::vector<uint64_t> cnt(P, 0);
    std::vector<uint32_t> touched;
    std::fill(parts, parts + V, -1);

    auto load = [&](uint32_t p) { return std::max(nv[p] / vcap, nd[p] / dcap); };
    auto place = [&](uint32_t v, bool first) -> bool {
        const double deg = (double)(g.ptr[v + 1] - g.ptr[v]);
        const int32_t cur = parts[v];
        if (cur >= 0) {  // score the alternatives as if v had been taken out
            nv[cur] -= 1.0;
            nd[cur] -= deg;
        }
        touched.clear();
        for (uint64_t k = g.ptr[v]; k < g.ptr[v + 1]; ++k) {
            const int32_t q = parts[g.adj[k]];
            if (q < 0) continue;
            if (cnt[q]++ == 0) touched.push_back((uint32_t)q);
        }
        auto fits = [&](uint32_t p) { return nv[p] + 1.0 <= vcap && nd[p] + deg <= dcap; };
        int32_t best = -1;
        double bestScore = -1.0, bestLoad = 0.0;
        auto consider = [&](uint32_t p) {
            if (!fits(p) && (int32_t)p != cur) return;
            const double l = load(p), s = (double)cnt[p] * std::max(0.0, 1.0 - l);
            if (s > bestScore || (s == bestScore && (l < bestLoad || (l == bestLoad && (int32_t)p < best)))) {
                best = (int32_t)p;
                bestScore = s;
                bestLoad = l;
            }
        };
        for (uint32_t p : touched) consider(p);
        if (best < 0 || bestScore <= 0.0) {  // no placed neighbour with room: least loaded partition
            if (!first && cur >= 0) {
                best = cur;  // restreaming never moves a vertex without a reason
            } else {
                for (uint32_t p = 0; p < P; ++p) consider(p);
                if (best < 0) {  // every partition is at a cap (rounding): take the emptiest
                    best = 0;
                    for (uint32_t p = 1; p < P; ++p)
                        if (load(p) < load((uint32_t)best)) best = (int32_t)p;
                }
            }
        }
        for (uint32_t p : touched) cnt[p] = 0;
        nv[best] += 1.0;
        nd[best] += deg;
        const bool moved = best != cur;
        parts[v] = best;
        return moved;
    };

    for (uint32_t v : order) place(v, true);
    for (uint32_t it = 0; it < passes; ++it) {
        uint64_t moved = 0;
        for (uint32_t v : order) moved += place(v, false);
        if (moved * 1000 < (uint64_t)V) break;
    }
    if (edge_cut) {
        uint64_t cut = 0;
        for (uint64_t e = 0; e < n_edges; ++e) cut += parts[src[e]] != parts[dst[e]];
        *edge_cut = cut;
    }
    return "";
}

// <bsnap> = graph.bsnap (header {int32 4; uint32 numVertices; uint64 numEdges} + (src, dst) records,
// inputs/graphToBinary.cpp:15-20); writes <out_dir>/<basename>.parts and <basename>.comm like
// partitioner.cpp:113-128.
std::string partition_file(const char *bsnap, uint32_t P, const char *out_dir, uint32_t passes) {
    std::ifstream f(bsnap, std::ios::binary | std::ios::ate);
    if (!f.good()) return std::string("cannot open ") + bsnap;
    const uint64_t bytes = (uint64_t)f.tellg();
    f.seekg(0);
    struct {
        int32_t sizeOfVertexType;
        uint32_t numVertices;
        uint64_t numEdges;
    } hdr{};
    static_assert(sizeof(hdr) == 16, "BELHeaderType is 16 bytes");
    if (bytes < sizeof hdr || !f.read(reinterpret_cast<char *>(&hdr), sizeof hdr)) return "graph file shorter than its header";
    if (hdr.sizeOfVertexType != 4) return "graph file: sizeOfVertexType != 4";
    const uint64_t n = (bytes - sizeof hdr) / 8;  // the reference reads to EOF (partitioner.cpp:69)
    std::vector<uint32_t> rec(2 * n);
    if (n && !f.read(reinterpret_cast<char *>(rec.data()), (std::streamsize)(8 * n))) return "graph file: short read";
    std::vector<uint32_t> src(n), dst(n);
    for (uint64_t e = 0; e < n; ++e) {
        src[e] = rec[2 * e];
        dst[e] = rec[2 * e + 1];
    }
    rec.clear();
    rec.shrink_to_fit();
    std::vector<int32_t> parts(hdr.numVertices);
    uint64_t cut = 0;
    std::string m = partition_edges(src.data(), dst.data(), n, hdr.numVertices, P, passes, parts.data(), &cut);
    if (!m.empty()) return m;
    std::string dir = out_dir ? out_dir : ".";
    if (!dir.empty() && dir.back() != '/') dir += '/';
    std::string base = bsnap;
    const size_t slash = base.find_last_of('/');
    if (slash != std::string::npos) base = base.substr(slash + 1);
    {
        std::ofstream c(dir + base + ".comm");
        if (!c.good()) return "cannot write " + dir + base + ".comm";
        c << "Communication cost: " << cut << std::endl;
    }
    std::ofstream p(dir + base + ".parts");
    if (!p.good()) return "cannot write " + dir + base + ".parts";
    std::string buf;
    buf.reserve((size_t)hdr.numVertices * 3);
    for (int32_t q : parts) {
        buf += std::to_string(q);
        buf += '\n';
    }
    p.write(buf.data(), (std::streamsize)buf.size());
    return p.good() ? "" : "write failed";
}

}  // namespace dory
