// Tile plan of the shared-memory-staged aggregation (spmm_tile.cu), built on the host at load time.
//
// The gather kernels of spmm.cu move E x F x 4 bytes from L2 into the SMs whatever HBM does; on a graph
// whose vertex numbering has locality (community-ordered ids: most in-edges of a run of destination rows
// come from one short range of source rows) that traffic can stay inside the SM instead.  The plan cuts
// the destination rows into TILES of `tileRows` consecutive rows and gives every tile one WINDOW of at
// most `windowRows` consecutive source rows (the range that holds most of the tile's edges, found with a
// bucketed histogram); every row's edge list is regrouped (stable) into [edges whose source lies in the
// tile's window | the rest].  The kernel stages the window's column slab in shared memory once per tile
// (bulk TMA copies) and serves the first group from there, the second from L2 as before.
#pragma once
#include <cstdint>
#include <vector>

namespace dory {

struct TilePlanHost {
    uint32_t tileRows = 0, windowRows = 0;  // parameters the plan was built with
    std::vector<uint64_t> ptrs;             // [2V + 1]: row v = [ptrs[2v], ptrs[2v+1]) in-window, [ptrs[2v+1], ptrs[2v+2]) rest
    std::vector<uint32_t> idx;              // [E] regrouped source rows
    std::vector<float> vals;                // [E]
    std::vector<uint32_t> rows;             // [V] rows grouped by tile: team rows first (degree-descending), then the others
    std::vector<uint32_t> tilePtr;          // [nTiles + 1] offsets into rows (tiles in heaviest-first order)
    std::vector<uint32_t> tileTeam;         // [nTiles] leading rows of the tile that the whole CTA walks together
    std::vector<uint32_t> tileWlo;          // [nTiles] first source row of the window
    std::vector<uint32_t> tileWrows;        // [nTiles] rows in the window (0 = none worth staging)
    std::vector<uint64_t> tileE0, tileE1;   // [nTiles] edge range [e0, e1) of the tile's rows (consecutive rows: one run)
    uint64_t inWindowEdges = 0;             // edges served from shared memory
    uint32_t maxWrows = 0;
    uint64_t maxTileEdges = 0;
    uint32_t maxTileRows = 0;
    double coverage() const { return idx.empty() ? 0.0 : (double)inWindowEdges / (double)idx.size(); }
};

struct TilePlanParams {
    uint32_t tileRows = 0;     // 0 = windowRows / 2
    uint32_t windowRows = 0;   // 0 = choose: the smallest power of two (<= maxWindowRows) that keeps >= 92 % of
                               //     the coverage the largest one reaches
    uint32_t maxWindowRows = 2048;
    uint32_t teamDegree = 512; // rows with at least this many edges are walked by the whole CTA
    uint32_t excludeDegree = 0; // rows with at least this many edges are left out of the tiles (0 = none): the
                                // low-degree kernel hands them to the CTA-per-row kernel of spmm.cu
    uint32_t edgeCap = 0;       // > 0: a tile holds at most this many edges (the low-degree kernel stages a tile's
                                // ids / weights in shared memory as well)
    bool keepRowOrder = false;  // true: rows of a tile stay in vertex order (the low-degree kernel addresses them as
                                // row0 + i); false: degree-descending with the team rows first
    double minTileCoverage = 0.25;  // a tile whose best window holds less than this share of its edges stages nothing
};

// ptrs / idx / vals: the adjacency as stored in graph.<id>.bin (u64 offsets [V+1], u32 source rows in
// [0, nSrcRows), fp32 values), possibly unaligned (byte pointers into the image).
void build_tile_plan(const uint8_t *ptrs, const uint8_t *idx, const uint8_t *vals, uint32_t V, uint32_t nSrcRows,
                     const TilePlanParams &prm, TilePlanHost &out);

// Share of the edges that the best window of every tile would hold, for a window of `windowRows` rows and
// tiles of `tileRows` rows, estimated on every `stride`-th tile.
double estimate_tile_coverage(const uint8_t *ptrs, const uint8_t *idx, uint32_t V, uint32_t nSrcRows, uint32_t tileRows,
                              uint32_t windowRows, uint32_t stride);

}  // namespace dory
