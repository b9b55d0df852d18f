// C entry points around build_tile_plan (dorylus_b200/csrc/tile_plan.h) for tests/test_tile_plan.py: the plan
// builder is pure host code, so its invariants are checked on CPU.  Test infrastructure, not product.
#include <cstdint>
#include <cstring>

#include "../../dorylus_b200/csrc/tile_plan.h"

extern "C" {

struct PlanSizes {
    uint32_t tileRows, windowRows, nTiles, nRows, maxWrows, maxTileRows;
    uint64_t E, inWindowEdges, maxTileEdges;
};

void *tp_build(const uint64_t *ptrs, const uint32_t *idx, const float *vals, uint32_t V, uint32_t nSrcRows, uint32_t tileRows,
               uint32_t windowRows, uint32_t maxWindowRows, uint32_t teamDegree, uint32_t excludeDegree, uint32_t edgeCap,
               int keepRowOrder, double minTileCoverage, PlanSizes *out) {
    dory::TilePlanParams prm;
    prm.tileRows = tileRows;
    prm.windowRows = windowRows;
    prm.maxWindowRows = maxWindowRows;
    prm.teamDegree = teamDegree;
    prm.excludeDegree = excludeDegree;
    prm.edgeCap = edgeCap;
    prm.keepRowOrder = keepRowOrder != 0;
    prm.minTileCoverage = minTileCoverage;
    auto *plan = new dory::TilePlanHost();
    dory::build_tile_plan(reinterpret_cast<const uint8_t *>(ptrs), reinterpret_cast<const uint8_t *>(idx),
                          reinterpret_cast<const uint8_t *>(vals), V, nSrcRows, prm, *plan);
    out->tileRows = plan->tileRows;
    out->windowRows = plan->windowRows;
    out->nTiles = (uint32_t)plan->tileTeam.size();
    out->nRows = (uint32_t)plan->rows.size();
    out->maxWrows = plan->maxWrows;
    out->maxTileRows = plan->maxTileRows;
    out->E = plan->idx.size();
    out->inWindowEdges = plan->inWindowEdges;
    out->maxTileEdges = plan->maxTileEdges;
    return plan;
}

// which: 0 ptrs (u64, 2V+1), 1 idx (u32, E), 2 vals (f32, E), 3 rows (u32), 4 tilePtr (u32, nTiles+1), 5 tileTeam,
// 6 tileWlo, 7 tileWrows (u32, nTiles), 8 tileE0, 9 tileE1 (u64, nTiles)
void tp_copy(void *h, int which, void *dst) {
    auto *p = static_cast<dory::TilePlanHost *>(h);
    auto cp = [&](const void *src, size_t bytes) { std::memcpy(dst, src, bytes); };
    switch (which) {
        case 0: cp(p->ptrs.data(), p->ptrs.size() * 8); break;
        case 1: cp(p->idx.data(), p->idx.size() * 4); break;
        case 2: cp(p->vals.data(), p->vals.size() * 4); break;
        case 3: cp(p->rows.data(), p->rows.size() * 4); break;
        case 4: cp(p->tilePtr.data(), p->tilePtr.size() * 4); break;
        case 5: cp(p->tileTeam.data(), p->tileTeam.size() * 4); break;
        case 6: cp(p->tileWlo.data(), p->tileWlo.size() * 4); break;
        case 7: cp(p->tileWrows.data(), p->tileWrows.size() * 4); break;
        case 8: cp(p->tileE0.data(), p->tileE0.size() * 8); break;
        case 9: cp(p->tileE1.data(), p->tileE1.size() * 8); break;
    }
}

void tp_free(void *h) { delete static_cast<dory::TilePlanHost *>(h); }

double tp_estimate(const uint64_t *ptrs, const uint32_t *idx, uint32_t V, uint32_t nSrcRows, uint32_t tileRows, uint32_t windowRows,
                   uint32_t stride) {
    return dory::estimate_tile_coverage(reinterpret_cast<const uint8_t *>(ptrs), reinterpret_cast<const uint8_t *>(idx), V,
                                        nSrcRows, tileRows, windowRows, stride);
}
}
