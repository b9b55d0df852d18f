// CPU unit test of host/saga_pipeline.hpp: recording stubs instead of the engine.
// Build: g++ -std=c++17 -O1 -Wall host/test_saga_pipeline.cpp -o host/test_saga_pipeline ; exit code 0 = pass.
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "saga_pipeline.hpp"

namespace {

int failures = 0;
#define CHECK(cond)                                                        \
    do {                                                                   \
        if (!(cond)) {                                                     \
            std::fprintf(stderr, "FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            ++failures;                                                    \
        }                                                                  \
    } while (0)

// incLayerGCN / incLayerGAT (engine/utils.cpp:714-748), restated for the stubs
void inc_gcn(dory_chunk &c, uint32_t L) {
    if (c.dir == DORY_FORWARD) {
        if (++c.layer == L) {
            c.dir = DORY_BACKWARD;
            --c.layer;
        }
    } else if (c.layer == 0) {
        c.dir = DORY_FORWARD;
        ++c.epoch;
    } else {
        --c.layer;
    }
}
void inc_gat(dory_chunk &c) {
    if (c.dir == DORY_FORWARD) {
        ++c.layer;
    } else if (c.layer == 0) {
        c.dir = DORY_FORWARD;
        c.vertex = 1;
        ++c.epoch;
    } else {
        --c.layer;
    }
}

struct Recorder {
    std::vector<std::string> log;
    uint32_t gnn, L;
    std::vector<float> accs;  // accuracy reported per epoch
    size_t statCalls = 0;
    std::string tag(const char *op, const dory_chunk &c) {
        char b[96];
        std::snprintf(b, sizeof b, "%s e%u %c%u c%u [%u,%u)", op, c.epoch, c.dir == DORY_FORWARD ? 'F' : 'B', c.layer,
                      c.localId, c.lowBound, c.upBound);
        return b;
    }
    saga::Ops ops() {
        saga::Ops o;
        o.aggregate = [this](const dory_chunk &c) { log.push_back(tag("GA", c)); return 0; };
        o.apply_vertex = [this](const dory_chunk &c) { log.push_back(tag("AV", c)); return 0; };
        o.scatter = [this](const dory_chunk &c) { log.push_back(tag("SC", c)); return 0; };
        o.apply_edge = [this](const dory_chunk &c) { log.push_back(tag("AE", c)); return 0; };
        o.predict = [this](const dory_chunk &c) { log.push_back(tag("PR", c)); return 0; };
        o.inc_layer = [this](dory_chunk &c) { gnn == DORY_GCN ? inc_gcn(c, L) : inc_gat(c); return 0; };
        o.apply_update = [this](uint32_t l) { log.push_back("UP W" + std::to_string(l)); return 0; };
        o.stats = [this](float *a, float *l, uint32_t *n) {
            *n = 100;
            *a = 100.f * (statCalls < accs.size() ? accs[statCalls] : 0.f);
            *l = 50.f;
            ++statCalls;
            return 0;
        };
        o.barrier = [this] { log.push_back("BARRIER"); };
        return o;
    }
};

std::vector<std::string> without_barriers(const std::vector<std::string> &v) {
    std::vector<std::string> r;
    for (auto &s : v)
        if (s != "BARRIER") r.push_back(s);
    return r;
}

void test_priority_order() {
    saga::ChunkQueue q;
    auto mk = [](uint32_t ep, uint32_t dir, uint32_t layer, uint8_t vtx, uint32_t id) {
        return dory_chunk{id, id, 0, 1, layer, dir, ep, vtx};
    };
    q.push(mk(2, DORY_FORWARD, 0, 1, 0));
    q.push(mk(1, DORY_BACKWARD, 0, 1, 0));
    q.push(mk(1, DORY_BACKWARD, 1, 1, 0));
    q.push(mk(1, DORY_BACKWARD, 1, 0, 0));
    q.push(mk(1, DORY_FORWARD, 1, 0, 0));
    q.push(mk(1, DORY_FORWARD, 1, 1, 1));
    q.push(mk(1, DORY_FORWARD, 1, 1, 0));
    q.push(mk(1, DORY_FORWARD, 0, 1, 0));
    // lowest epoch, forward first, shallow forward layers first, vertex before edge (forward),
    // lower ids first; backward: deeper layers first, edge before vertex
    const char *want[] = {"1F0v1#0", "1F1v1#0", "1F1v1#1", "1F1v0#0", "1B1v0#0", "1B1v1#0", "1B0v1#0", "2F0v1#0"};
    for (const char *w : want) {
        const dory_chunk c = q.top();
        q.pop();
        char b[32];
        std::snprintf(b, sizeof b, "%u%c%uv%u#%u", c.epoch, c.dir == DORY_FORWARD ? 'F' : 'B', c.layer, c.vertex, c.localId);
        CHECK(std::string(b) == w);
    }
}

void test_gcn_sequence() {
    Recorder r{{}, DORY_GCN, 2, {}};
    saga::Config cfg;
    cfg.gnn = DORY_GCN;
    cfg.numLayers = 2;
    cfg.numChunks = 2;
    cfg.numEpochs = 2;
    cfg.localVtxCnt = 9;
    cfg.log = nullptr;
    saga::Pipeline p(cfg, r.ops());
    CHECK(p.run() == 0);
    const std::vector<std::string> epoch1 = {
        "GA e1 F0 c0 [0,5)", "GA e1 F0 c1 [5,9)", "AV e1 F0 c0 [0,5)", "SC e1 F1 c0 [0,5)",
        "GA e1 F1 c0 [0,5)", "GA e1 F1 c1 [5,9)", "AV e1 F1 c0 [0,5)", "UP W1", "SC e1 B1 c0 [0,5)",
        "GA e1 B1 c0 [0,5)", "GA e1 B1 c1 [5,9)", "AV e1 B1 c0 [0,5)", "UP W0"};
    std::vector<std::string> got = without_barriers(r.log);
    CHECK(got.size() == 2 * epoch1.size());
    for (size_t i = 0; i < epoch1.size() && i < got.size(); ++i) CHECK(got[i] == epoch1[i]);
    if (got.size() == 2 * epoch1.size()) CHECK(got[epoch1.size()] == "GA e2 F0 c0 [0,5)");
    CHECK(p.epochs_run() == 2 && p.epoch_times().size() == 2 && r.statCalls == 2);
    // barriers: one per epoch boundary (3) + two around every scatter (2 per epoch x 2)
    size_t barriers = 0;
    for (auto &s : r.log) barriers += s == "BARRIER";
    CHECK(barriers == 3 + 2 * 2 * 2);
}

void test_three_layer_gcn_backward_scatters() {
    Recorder r{{}, DORY_GCN, 3, {}};
    saga::Config cfg;
    cfg.numLayers = 3;
    cfg.numEpochs = 1;
    cfg.localVtxCnt = 4;
    cfg.log = nullptr;
    saga::Pipeline p(cfg, r.ops());
    CHECK(p.run() == 0);
    const std::vector<std::string> want = {
        "GA e1 F0 c0 [0,4)", "AV e1 F0 c0 [0,4)", "SC e1 F1 c0 [0,4)", "GA e1 F1 c0 [0,4)", "AV e1 F1 c0 [0,4)",
        "SC e1 F2 c0 [0,4)", "GA e1 F2 c0 [0,4)", "AV e1 F2 c0 [0,4)", "UP W2", "SC e1 B2 c0 [0,4)",
        "GA e1 B2 c0 [0,4)", "AV e1 B2 c0 [0,4)", "UP W1", "SC e1 B1 c0 [0,4)", "GA e1 B1 c0 [0,4)",
        "AV e1 B1 c0 [0,4)", "UP W0"};
    CHECK(without_barriers(r.log) == want);
}

void test_gat_sequence() {
    Recorder r{{}, DORY_GAT, 2, {}};
    saga::Config cfg;
    cfg.gnn = DORY_GAT;
    cfg.numLayers = 2;
    cfg.numEpochs = 1;
    cfg.localVtxCnt = 4;
    cfg.log = nullptr;
    saga::Pipeline p(cfg, r.ops());
    CHECK(p.run() == 0);
    // forward AV -> SC -> AE -> GA per layer (+ predict), backward SC -> AE -> GA -> AV
    const std::vector<std::string> want = {
        "AV e1 F0 c0 [0,4)", "SC e1 F1 c0 [0,4)", "AE e1 F1 c0 [0,4)", "GA e1 F1 c0 [0,4)",
        "AV e1 F1 c0 [0,4)", "SC e1 F2 c0 [0,4)", "AE e1 F2 c0 [0,4)", "GA e1 F2 c0 [0,4)", "PR e1 F2 c0 [0,4)",
        "SC e1 B2 c0 [0,4)", "AE e1 B2 c0 [0,4)", "GA e1 B2 c0 [0,4)", "AV e1 B2 c0 [0,4)", "UP W1",
        "SC e1 B1 c0 [0,4)", "AE e1 B1 c0 [0,4)", "GA e1 B1 c0 [0,4)", "AV e1 B1 c0 [0,4)", "UP W0"};
    CHECK(without_barriers(r.log) == want);
    CHECK(r.statCalls == 1);
}

void test_early_stop() {
    Recorder r{{}, DORY_GCN, 2, {0.50f, 0.79f, 0.70f, 0.81f, 0.9f}};
    saga::Config cfg;
    cfg.numLayers = 2;
    cfg.numEpochs = 10;
    cfg.localVtxCnt = 4;
    cfg.targetAcc = 0.80f;
    cfg.switchThreshold = 0.02f;
    cfg.log = nullptr;
    saga::Pipeline p(cfg, r.ops());
    CHECK(p.run() == 0);
    // epoch 2 reaches CLOSE (0.79 >= 0.78); epoch 3 falls back but the state never goes backwards;
    // epoch 4 reaches DONE and the run stops at the next epoch boundary
    CHECK(p.converge_state() == saga::DONE);
    CHECK(p.epochs_run() == 4);
    CHECK(r.statCalls == 4);
}

void test_lost_chunk_is_an_error() {
    Recorder r{{}, DORY_GCN, 2, {}};
    saga::Ops o = r.ops();
    o.inc_layer = [](dory_chunk &) { return 0; };  // a chunk that never advances ends up nowhere new
    saga::Config cfg;
    cfg.numLayers = 2;
    cfg.numChunks = 2;
    cfg.numEpochs = 1;
    cfg.localVtxCnt = 1;  // second chunk is empty but still travels
    cfg.log = nullptr;
    saga::Pipeline p(cfg, o);
    // with inc_layer broken the forward chunk cycles GA -> AV -> SC -> AE at the same layer forever;
    // the run must still terminate: bound it by making scatter fail after a while
    int scatters = 0;
    o.scatter = [&](const dory_chunk &) { return ++scatters > 5 ? DORY_ESTATE : 0; };
    saga::Pipeline q(cfg, o);
    CHECK(q.run() == DORY_ESTATE);
}

}  // namespace

int main() {
    test_priority_order();
    test_gcn_sequence();
    test_three_layer_gcn_backward_scatters();
    test_gat_sequence();
    test_early_stop();
    test_lost_chunk_is_an_error();
    if (failures) {
        std::fprintf(stderr, "%d check(s) failed\n", failures);
        return 1;
    }
    std::puts("saga_pipeline: all checks passed");
    return 0;
}
