// Host-side builder of the tile plan (tile_plan.h).  Pure C++, no CUDA.
#include "tile_plan.h"

#include <algorithm>
#include <atomic>
#include <cstring>
#include <numeric>
#include <thread>

namespace dory {
namespace {

inline uint64_t rd64(const uint8_t *p, size_t i) {
    uint64_t v;
    std::memcpy(&v, p + 8 * i, 8);
    return v;
}
inline uint32_t rd32(const uint8_t *p, size_t i) {
    uint32_t v;
    std::memcpy(&v, p + 4 * i, 4);
    return v;
}

unsigned plan_threads() { return std::max(1u, std::min(std::thread::hardware_concurrency(), 32u)); }

template <class F>
void parallel_tiles(uint32_t nTiles, F &&fn) {
    const unsigned nt = nTiles < 64 ? 1 : plan_threads();
    std::atomic<uint32_t> next{0};
    auto work = [&]() {
        std::vector<uint32_t> scratch;
        for (;;) {
            const uint32_t t0 = next.fetch_add(16);
            if (t0 >= nTiles) return;
            for (uint32_t t = t0; t < std::min(nTiles, t0 + 16); ++t) fn(t, scratch);
        }
    };
    if (nt == 1) {
        work();
        return;
    }
    std::vector<std::thread> th;
    for (unsigned i = 0; i < nt; ++i) th.emplace_back(work);
    for (auto &x : th) x.join();
}

// Best window of `W` consecutive source rows (start aligned to W / 8 rows) for the edges [e0, e1):
// returns the first row of the window and the number of edges it holds.
std::pair<uint32_t, uint64_t> best_window(const uint8_t *idx, uint64_t e0, uint64_t e1, uint32_t W, uint32_t nSrcRows,
                                          std::vector<uint32_t> &buckets) {
    const uint32_t g = std::max(1u, W / 8);
    const uint32_t span = (W + g - 1) / g;  // buckets per window
    buckets.clear();
    buckets.reserve(e1 - e0);
    for (uint64_t e = e0; e < e1; ++e) buckets.push_back(rd32(idx, e) / g);
    std::sort(buckets.begin(), buckets.end());
    uint64_t best = 0;
    uint32_t bestStart = 0;
    size_t lo = 0;
    for (size_t hi = 0; hi < buckets.size(); ++hi) {
        while (buckets[hi] - buckets[lo] >= span) ++lo;
        // window starting at bucket buckets[lo] holds entries lo..hi; prefer the window that ENDS at the
        // current bucket's upper edge only through its count: ties keep the first
        if (hi - lo + 1 > best) {
            best = hi - lo + 1;
            bestStart = buckets[lo];
        }
    }
    uint32_t wlo = bestStart * g;
    if (wlo + W > nSrcRows) wlo = nSrcRows > W ? nSrcRows - W : 0;  // keep the window full-sized at the end of the block
    return {wlo, best};
}

}  // namespace

double estimate_tile_coverage(const uint8_t *ptrs, const uint8_t *idx, uint32_t V, uint32_t nSrcRows, uint32_t tileRows,
                              uint32_t windowRows, uint32_t stride) {
    const uint32_t nTiles = (V + tileRows - 1) / tileRows;
    std::atomic<uint64_t> in{0}, all{0};
    stride = std::max(stride, 1u);
    parallel_tiles((nTiles + stride - 1) / stride, [&](uint32_t k, std::vector<uint32_t> &scratch) {
        const uint32_t t = k * stride;
        const uint32_t r0 = t * tileRows, r1 = (uint32_t)std::min<uint64_t>((uint64_t)r0 + tileRows, V);
        const uint64_t e0 = rd64(ptrs, r0), e1 = rd64(ptrs, r1);
        if (e1 == e0) return;
        auto w = best_window(idx, e0, e1, windowRows, nSrcRows, scratch);
        // exact count for the (possibly shifted) window
        uint64_t c = 0;
        for (uint64_t e = e0; e < e1; ++e) {
            const uint32_t s = rd32(idx, e);
            c += s >= w.first && s - w.first < windowRows;
        }
        in += c;
        all += e1 - e0;
    });
    return all ? (double)in / (double)all : 0.0;
}

void build_tile_plan(const uint8_t *ptrs, const uint8_t *idx, const uint8_t *vals, uint32_t V, uint32_t nSrcRows,
                     const TilePlanParams &prm, TilePlanHost &out) {
    const uint64_t E = rd64(ptrs, V);
    uint32_t W = prm.windowRows;
    if (W == 0) {
        // smallest power-of-two window that keeps >= 92 % of what the largest one reaches (every 8th tile)
        uint32_t wmax = 32;
        while (wmax * 2 <= prm.maxWindowRows) wmax *= 2;
        const uint32_t stride = 8;
        const double top = estimate_tile_coverage(ptrs, idx, V, nSrcRows, std::max(16u, wmax / 2), wmax, stride);
        W = wmax;
        for (uint32_t w = wmax / 2; w >= 32; w /= 2) {
            const double c = estimate_tile_coverage(ptrs, idx, V, nSrcRows, std::max(16u, w / 2), w, stride);
            if (c < 0.92 * top) break;
            W = w;
        }
    }
    W = std::max(1u, std::min(W, std::max(nSrcRows, 1u)));
    const uint32_t R = std::max(1u, prm.tileRows ? prm.tileRows : std::max(16u, W / 2));
    auto degree = [&](uint32_t v) { return rd64(ptrs, (size_t)v + 1) - rd64(ptrs, v); };
    // ---- tile boundaries: runs of consecutive rows, at most R rows and (edgeCap > 0) edgeCap edges each; a
    // row of excludeDegree edges or more belongs to no tile and ends the run it falls into
    std::vector<std::pair<uint32_t, uint32_t>> ranges;
    {
        uint32_t r0 = 0;
        uint64_t edges = 0;
        auto close = [&](uint32_t r1) {
            if (r1 > r0) ranges.emplace_back(r0, r1);
        };
        for (uint32_t v = 0; v < V; ++v) {
            const uint64_t d = degree(v);
            if (prm.excludeDegree && d >= prm.excludeDegree) {
                close(v);
                r0 = v + 1;
                edges = 0;
                continue;
            }
            if (v - r0 >= R || (prm.edgeCap && v > r0 && edges + d > prm.edgeCap)) {
                close(v);
                r0 = v;
                edges = 0;
            }
            edges += d;
        }
        close(V);
    }
    const uint32_t nTiles = (uint32_t)ranges.size();
    out.tileRows = R;
    out.windowRows = W;
    out.ptrs.assign(2 * (size_t)V + 1, 0);
    out.idx.resize(E);
    out.vals.resize(E);
    std::vector<uint32_t> natural(V);  // rows of a tile at [r0, r0 + count) in issue order
    std::vector<uint32_t> wlo(nTiles, 0), wrows(nTiles, 0), team(nTiles, 0);
    std::vector<uint64_t> tileEdges(nTiles, 0);
    std::atomic<uint64_t> inWin{0};
    // rows outside every tile keep their whole edge list in the "rest" part (nobody walks it through the plan)
    {
        std::vector<uint8_t> covered(V, 0);
        for (auto &rg : ranges)
            for (uint32_t v = rg.first; v < rg.second; ++v) covered[v] = 1;
        for (uint32_t v = 0; v < V; ++v)
            if (!covered[v]) {
                const uint64_t b = rd64(ptrs, v), e = rd64(ptrs, (size_t)v + 1);
                out.ptrs[2 * (size_t)v] = b;
                out.ptrs[2 * (size_t)v + 1] = b;
                for (uint64_t k = b; k < e; ++k) {
                    out.idx[k] = rd32(idx, k);
                    std::memcpy(&out.vals[k], vals + 4 * k, 4);
                }
            }
    }
    parallel_tiles(nTiles, [&](uint32_t t, std::vector<uint32_t> &scratch) {
        const uint32_t r0 = ranges[t].first, r1 = ranges[t].second;
        const uint64_t e0 = rd64(ptrs, r0), e1 = rd64(ptrs, r1);
        tileEdges[t] = e1 - e0;
        uint32_t lo = 0, n = 0;
        if (e1 > e0) {
            auto w = best_window(idx, e0, e1, W, nSrcRows, scratch);
            lo = w.first;
            n = std::min(W, nSrcRows - lo);
            // exact share of the (clipped / shifted) window
            uint64_t c = 0;
            for (uint64_t e = e0; e < e1; ++e) {
                const uint32_t s = rd32(idx, e);
                c += s >= lo && s - lo < n;
            }
            if ((double)c < prm.minTileCoverage * (double)(e1 - e0)) n = 0;
        }
        wlo[t] = lo;
        wrows[t] = n;
        uint64_t mine = 0;
        for (uint32_t v = r0; v < r1; ++v) {
            const uint64_t b = rd64(ptrs, v), e = rd64(ptrs, (size_t)v + 1);
            uint64_t cin = 0;
            for (uint64_t k = b; k < e; ++k) {
                const uint32_t s = rd32(idx, k);
                cin += n && s >= lo && s - lo < n;
            }
            out.ptrs[2 * (size_t)v] = b;
            out.ptrs[2 * (size_t)v + 1] = b + cin;
            uint64_t pi = b, po = b + cin;  // stable: both groups keep the edge-file order
            for (uint64_t k = b; k < e; ++k) {
                const uint32_t s = rd32(idx, k);
                float w;
                std::memcpy(&w, vals + 4 * k, 4);
                const bool in = n && s >= lo && s - lo < n;
                const uint64_t pos = in ? pi++ : po++;
                out.idx[pos] = s;
                out.vals[pos] = w;
            }
            mine += cin;
        }
        inWin += mine;
        uint32_t *rows = natural.data() + r0;
        std::iota(rows, rows + (r1 - r0), r0);
        uint32_t nt = 0;
        if (!prm.keepRowOrder) {
            // degree-descending (stable): the long rows start first, and the rows the whole CTA walks together
            // are a prefix
            std::stable_sort(rows, rows + (r1 - r0), [&](uint32_t a, uint32_t b) { return degree(a) > degree(b); });
            while (nt < r1 - r0 && degree(rows[nt]) >= prm.teamDegree) ++nt;
        }
        team[t] = nt;
    });
    out.ptrs[2 * (size_t)V] = E;
    out.inWindowEdges = inWin.load();
    // tiles heaviest first
    std::vector<uint32_t> order(nTiles);
    std::iota(order.begin(), order.end(), 0u);
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return tileEdges[a] > tileEdges[b]; });
    out.tilePtr.resize((size_t)nTiles + 1);
    out.tileTeam.resize(nTiles);
    out.tileWlo.resize(nTiles);
    out.tileWrows.resize(nTiles);
    out.tileE0.resize(nTiles);
    out.tileE1.resize(nTiles);
    out.rows.clear();
    out.rows.reserve(V);
    out.maxWrows = 0;
    out.maxTileEdges = 0;
    out.maxTileRows = 0;
    for (uint32_t i = 0; i < nTiles; ++i) {
        const uint32_t t = order[i];
        const uint32_t r0 = ranges[t].first, r1 = ranges[t].second;
        out.tilePtr[i] = (uint32_t)out.rows.size();
        out.rows.insert(out.rows.end(), natural.begin() + r0, natural.begin() + r1);
        out.tileTeam[i] = team[t];
        out.tileWlo[i] = wlo[t];
        out.tileWrows[i] = wrows[t];
        out.tileE0[i] = rd64(ptrs, r0);
        out.tileE1[i] = rd64(ptrs, r1);
        out.maxWrows = std::max(out.maxWrows, wrows[t]);
        out.maxTileEdges = std::max<uint64_t>(out.maxTileEdges, tileEdges[t]);
        out.maxTileRows = std::max(out.maxTileRows, r1 - r0);
    }
    out.tilePtr[nTiles] = (uint32_t)out.rows.size();
}

}  // namespace dory
