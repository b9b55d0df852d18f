// The reference's chunk pipeline as a host-side shell over the C ABI.
//
// Engine::run starts one thread pool per SAGA stage (engine/engine.cpp:237-314) and five worker
// functions hand Chunks from queue to queue (engine/ops/pipeline.cpp): scheduler -> Gather ->
// ApplyVertex -> Scatter (+ stash / barrier) -> ApplyEdge -> Gather ..., each queue a priority queue
// ordered by Chunk::operator< (common/utils.hpp:76-89); NNRecvCallback{GCN,GAT} decides where a chunk
// goes after an NN stage (commmanager/resource_comm.cpp:17-90); the weight server turns the summed
// accuracy into the EARLY -> CLOSE -> DONE state machine (weight-server/weightserver.cpp:230-294).
//
// Here the same routing runs on ONE host thread (the C ABI is one caller thread per engine and every
// operator only enqueues GPU work), in the synchronous mode that CPU/GPU backends use: the queues,
// their priority order, the per-epoch barrier, the Scatter barrier, the per-stage timers and the
// log lines ("Sync Epoch %u starts...", "Time for epoch %u: %.2lfms", "Epoch %u, acc: ...",
// "STATE switch: ...", "<EM>: ...") are the reference's.  Differences, all forced by the engine:
//   * dory_apply_vertex and dory_scatter cover the whole partition (like CPUComm / the scatter
//     barrier), so with several chunks per partition they run once per (epoch, dir, layer), when the
//     last chunk of that step arrives -- the "pre-barrier inside applyVertex" pipeline.cpp:226 notes;
//   * the bounded-staleness async mode is Lambda-only in the reference and is not built.
// The operator table is std::function so that the unit test can record the sequence without a GPU.
#pragma once

#include <algorithm>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <functional>
#include <queue>
#include <string>
#include <vector>

#include "../include/dorylus_b200.h"

namespace saga {

// == Chunk::operator< (common/utils.hpp:76-89): top() of a std::priority_queue is the chunk that is
// "largest", i.e. lowest epoch first, FORWARD before BACKWARD, shallower forward / deeper backward
// layers first, vertex before edge work forward (edge before vertex backward), then ids and bounds.
struct ChunkLess {
    bool operator()(const dory_chunk &a, const dory_chunk &b) const {
        if (a.epoch != b.epoch) return a.epoch > b.epoch;
        if (a.dir != b.dir) return a.dir > b.dir;
        if (a.layer != b.layer) return a.dir == DORY_FORWARD ? a.layer > b.layer : a.layer < b.layer;
        if (a.vertex != b.vertex) return a.dir == DORY_FORWARD ? (!a.vertex && b.vertex) : (a.vertex && !b.vertex);
        if (a.localId != b.localId) return a.localId > b.localId;
        if (a.globalId != b.globalId) return a.globalId > b.globalId;
        if (a.lowBound != b.lowBound) return a.lowBound > b.lowBound;
        return a.upBound > b.upBound;
    }
};
using ChunkQueue = std::priority_queue<dory_chunk, std::vector<dory_chunk>, ChunkLess>;

enum ConvergeState { EARLY = 0, CLOSE = 1, DONE = 2 };  // common/utils.hpp:50-53
inline const char *converge_name(ConvergeState s) { return s == EARLY ? "EARLY" : s == CLOSE ? "CLOSE" : "DONE"; }

struct Ops {
    std::function<int(const dory_chunk &)> aggregate, apply_vertex, scatter, apply_edge, predict;
    std::function<int(dory_chunk &)> inc_layer;
    std::function<int(uint32_t)> apply_update;
    // validation statistics of the epoch that just ran its last forward layer, summed over all
    // partitions (updateGlobalAccLoss, weightserver.cpp:230-262): acc sum, loss sum, vertex count
    std::function<int(float *, float *, uint32_t *)> stats;
    std::function<void()> barrier;  // NodeManager::barrier; may be empty on one node
    std::function<void()> sync;     // drain the device before a host timestamp; may be empty
};

struct Config {
    uint32_t gnn = DORY_GCN;
    uint32_t numLayers = 2;
    uint32_t numChunks = 1;  // numLambdasForward: chunks per partition (run/run-onnode:62-70 uses 1)
    uint32_t numEpochs = 1;
    uint32_t localVtxCnt = 0;
    uint32_t nodeId = 0;
    float targetAcc = 1.1f;         // weight-server argv[10]; > 1 never stops early
    float switchThreshold = 0.02f;  // weight-server argv[14]
    bool stageTimers = true;        // per-stage times for the <EM> report (one device sync per stage)
    FILE *log = stderr;
};

class Pipeline {
public:
    Pipeline(const Config &cfg, const Ops &ops) : c_(cfg), o_(ops) {
        tAgg_.assign(2 * c_.numLayers + 1, 0.0);
        tAv_ = tSc_ = tAe_ = tAgg_;
    }

    // == Engine::run for numEpochs synchronous epochs.  Returns 0 or the first operator error.
    int run() {
        load_chunks();
        while (!halt_) {
            const uint64_t before = steps_;
            int rc;
            if ((rc = schedule())) return rc;
            if ((rc = drain(GA_, &Pipeline::gather))) return rc;
            if ((rc = drain(AV_, &Pipeline::apply_vertex))) return rc;
            if ((rc = scatter_stage())) return rc;
            if ((rc = drain(AE_, &Pipeline::apply_edge))) return rc;
            if (steps_ == before && !halt_) return DORY_ESTATE;  // a chunk got lost: never spin
        }
        return 0;
    }

    // == the tail of Engine::printEngineMetrics (engine/utils.cpp:219-292) for a synchronous run
    void report() const {
        log("<EM>: Backend B200");
        log("<EM>: %u sync epochs and %u async epochs", numSyncEpochs_, 0u);
        log("<EM>: Using %u lambdas", c_.numChunks);
        const double denom = numSyncEpochs_ ? (double)numSyncEpochs_ : 1.0;
        log("<EM>: Forward:  Time per stage:");
        for (uint32_t i = 0; i < c_.numLayers; ++i) stage_lines(i, denom);
        log("<EM>: Backward: Time per stage:");
        for (uint32_t i = c_.numLayers; i < 2 * c_.numLayers; ++i) stage_lines(i, denom);
        log("<EM>: Final accuracy %.3lf", (double)lastAcc_);
        double sum = 0.0;
        for (double d : epochTimes_) sum += d;
        log("<EM>: Average  sync epoch time %.3lf ms", epochTimes_.empty() ? 0.0 : sum / epochTimes_.size());
        log("<EM>: Average async epoch time %.3lf ms", 0.0);
    }

    const std::vector<double> &epoch_times() const { return epochTimes_; }
    ConvergeState converge_state() const { return state_; }
    uint32_t epochs_run() const { return numSyncEpochs_; }
    float last_acc() const { return lastAcc_; }
    float last_loss() const { return lastLoss_; }

private:
    using Stage = int (Pipeline::*)(dory_chunk);
    static constexpr uint32_t kStartEpoch = 0;  // START_EPOCH, engine/engine.hpp

    static double now_ms() {
        return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
    }
    void log(const char *fmt, ...) const {  // printLog, graph-server/utils/utils.cpp:18-30
        if (!c_.log) return;
        std::fprintf(c_.log, "[ Node %3u ]  ", c_.nodeId);
        va_list ap;
        va_start(ap, fmt);
        std::vfprintf(c_.log, fmt, ap);
        va_end(ap);
        std::fputc('\n', c_.log);
    }
    void stage_lines(uint32_t i, double denom) const {
        log("<EM>    Aggregation   %2u  %.3lf ms", i, tAgg_[i] / denom);
        log("<EM>    ApplyVertex   %2u  %.3lf ms", i, tAv_[i] / denom);
        log("<EM>    Scatter       %2u  %.3lf ms", i, tSc_[i] / denom);
        log("<EM>    ApplyEdge     %2u  %.3lf ms", i, tAe_[i] / denom);
    }
    uint32_t abs_layer(const dory_chunk &c) const {  // Engine::getAbsLayer, engine/utils.cpp:706-710
        const uint32_t a = c.dir == DORY_FORWARD ? c.layer : 2 * c_.numLayers - 1 - c.layer;
        return a < tAgg_.size() ? a : (uint32_t)tAgg_.size() - 1;
    }
    bool is_last_layer(const dory_chunk &c) const {  // Engine::isLastLayer, engine/utils.cpp:749-752
        return c.dir == DORY_BACKWARD && c.layer == 0 && c.vertex;
    }
    template <class F>
    int timed(std::vector<double> &acc, const dory_chunk &c, F &&f) {
        if (!c_.stageTimers) return f();
        if (o_.sync) o_.sync();  // operators only enqueue: drain the device around each stage
        const double t0 = now_ms();
        const int rc = f();
        if (o_.sync) o_.sync();
        acc[abs_layer(c)] += now_ms() - t0;
        return rc;
    }

    void load_chunks() {  // Engine::loadChunks, engine/utils.cpp:598-609
        const uint32_t n = c_.numChunks ? c_.numChunks : 1;
        const uint32_t size = (c_.localVtxCnt + n - 1) / n;
        for (uint32_t cid = 0; cid < n; ++cid) {
            const uint32_t low = std::min(cid * size, c_.localVtxCnt), up = std::min(low + size, c_.localVtxCnt);
            SCH_.push(dory_chunk{cid, c_.nodeId * n + cid, low, up, 0, DORY_FORWARD, kStartEpoch + 1, 1});
        }
        currEpoch_ = kStartEpoch;
    }

    // == Engine::scheduleAsyncFunc in sync mode (pipeline.cpp:6-180): an epoch starts when every
    // chunk has come back; epoch 0 -> 1 is not timed (pipeline.cpp:117).
    int schedule() {
        if (SCH_.empty()) return 0;
        const dory_chunk top = SCH_.top();
        if (top.epoch > currEpoch_) {
            if (SCH_.size() < c_.numChunks) return 0;  // (4.1) wait for all chunks of the epoch
            if (o_.barrier) o_.barrier();
            if (o_.sync) o_.sync();
            const double t = now_ms();
            if (currEpoch_ > kStartEpoch) {  // (4.2) timing, skip epoch 0
                epochTimes_.push_back(t - epochStart_);
                log("Time for epoch %u: %.2lfms", currEpoch_, t - epochStart_);
            }
            epochStart_ = t;
            if (top.epoch > c_.numEpochs || state_ == DONE) {  // (1) all epochs finished / early stop
                halt_ = true;
                return 0;
            }
            ++currEpoch_;
            ++numSyncEpochs_;
            log("Sync Epoch %u starts...", currEpoch_);
        }
        while (!SCH_.empty() && SCH_.top().epoch == currEpoch_) {
            ++steps_;
            (c_.gnn == DORY_GCN ? GA_ : AV_).push(SCH_.top());
            SCH_.pop();
        }
        return 0;
    }

    int drain(ChunkQueue &q, Stage stage) {
        while (!q.empty()) {
            const dory_chunk c = q.top();
            q.pop();
            ++steps_;
            if (int rc = (this->*stage)(c)) return rc;
        }
        return 0;
    }

    // == Engine::gatherWorkFunc (pipeline.cpp:184-221)
    int gather(dory_chunk c) {
        if (int rc = timed(tAgg_, c, [&] { return o_.aggregate(c); })) return rc;
        if (c_.gnn == DORY_GAT && c.dir == DORY_FORWARD && c.layer == c_.numLayers) {  // last forward layer
            if (o_.predict)
                if (int rc = o_.predict(c)) return rc;
            if (++arrived_ == c_.numChunks) {
                arrived_ = 0;
                if (int rc = epoch_statistics(c.epoch)) return rc;
            }
            c.dir = DORY_BACKWARD;  // switch direction
            SC_.push(c);
            return 0;
        }
        AV_.push(c);
        return 0;
    }

    // == applyVertexWorkFunc + applyVertex{GCN,GAT} + NNRecvCallback{GCN,GAT}.  The NN runs once per
    // step, on the chunk that completes it.
    int apply_vertex(dory_chunk c) {
        c.vertex = 1;
        pendingAv_.push_back(c);
        if (pendingAv_.size() < c_.numChunks) return 0;
        const dory_chunk first = pendingAv_.front();
        if (int rc = timed(tAv_, first, [&] { return o_.apply_vertex(first); })) return rc;
        // which layer's weights got their gradient: forward = the last GCN layer; backward = the
        // layer the incremented chunk names (applyVertexGCN / applyVertexGAT increment first)
        dory_chunk nn = first;
        if (first.dir == DORY_BACKWARD) o_.inc_layer(nn);
        const bool gcnLastForward = c_.gnn == DORY_GCN && first.dir == DORY_FORWARD && first.layer + 1 == c_.numLayers;
        if (gcnLastForward) {
            if (int rc = epoch_statistics(first.epoch)) return rc;
        }
        if (o_.apply_update && (gcnLastForward || first.dir == DORY_BACKWARD))
            if (int rc = o_.apply_update(nn.layer)) return rc;
        for (dory_chunk ch : pendingAv_) {
            dory_chunk next = ch;
            if (ch.dir == DORY_BACKWARD) {  // NNCompute ran on the incremented chunk
                o_.inc_layer(next);
                if (is_last_layer(next)) {  // end of the epoch: increment again into the next one
                    o_.inc_layer(next);
                    SCH_.push(next);
                } else {
                    SC_.push(next);
                }
            } else {  // forward: increment after the NN
                o_.inc_layer(next);
                SC_.push(next);
            }
        }
        pendingAv_.clear();
        return 0;
    }

    // == scatterWorkFunc for CPU/GPU backends (pipeline.cpp:256-342): barrier when every chunk has
    // arrived, scatter, stash, barrier again once the ghosts are in, then on to ApplyEdge.
    int scatter_stage() {
        if (SC_.empty()) return 0;
        if (SC_.size() < c_.numChunks) return 0;
        if (o_.barrier) o_.barrier();
        ++steps_;
        const dory_chunk first = SC_.top();
        if (int rc = timed(tSc_, first, [&] { return o_.scatter(first); })) return rc;
        if (o_.barrier) o_.barrier();
        while (!SC_.empty()) {
            AE_.push(SC_.top());
            SC_.pop();
        }
        return 0;
    }

    // == applyEdgeWorkFunc (pipeline.cpp:362-392); GCN: nothing to compute, on to Gather
    // (applyEdgeGCN, gcn_ops.cpp:364-366); GAT: the edge NN covers the partition, once per step.
    int apply_edge(dory_chunk c) {
        c.vertex = 0;
        if (c_.gnn == DORY_GCN) {
            GA_.push(c);
            return 0;
        }
        pendingAe_.push_back(c);
        if (pendingAe_.size() < c_.numChunks) return 0;
        const dory_chunk first = pendingAe_.front();
        if (int rc = timed(tAe_, first, [&] { return o_.apply_edge(first); })) return rc;
        for (const dory_chunk &ch : pendingAe_) GA_.push(ch);  // NNRecvCallbackGAT: AE & AEB -> GAQueue
        pendingAe_.clear();
        return 0;
    }

    // == WeightServer::updateGlobalAccLoss + tryEarlyStop (weightserver.cpp:230-294)
    int epoch_statistics(uint32_t epoch) {
        if (!o_.stats) return 0;
        float acc = 0.f, loss = 0.f;
        uint32_t cnt = 0;
        if (int rc = o_.stats(&acc, &loss, &cnt)) return rc;
        if (cnt) {
            acc /= cnt;
            loss /= cnt;
        }
        lastAcc_ = acc;
        lastLoss_ = loss;
        if (c_.nodeId == 0) log("Epoch %u, acc: %.4f, loss: %.4f", epoch, acc, loss);
        const ConvergeState cur = acc >= c_.targetAcc ? DONE : acc >= c_.targetAcc - c_.switchThreshold ? CLOSE : EARLY;
        if (cur > state_) {  // transitions only go EARLY -> CLOSE -> DONE
            if (c_.nodeId == 0) log("STATE switch: %s -> %s at epoch %u", converge_name(state_), converge_name(cur), epoch);
            state_ = cur;
        }
        return 0;
    }

    Config c_;
    Ops o_;
    ChunkQueue SCH_, GA_, AV_, SC_, AE_;
    std::vector<dory_chunk> pendingAv_, pendingAe_;
    uint32_t arrived_ = 0;
    uint64_t steps_ = 0;
    uint32_t currEpoch_ = 0, numSyncEpochs_ = 0;
    bool halt_ = false;
    double epochStart_ = 0.0;
    std::vector<double> epochTimes_, tAgg_, tAv_, tSc_, tAe_;
    ConvergeState state_ = EARLY;
    float lastAcc_ = 0.f, lastLoss_ = 0.f;
};

// The operator table bound to one engine.
inline Ops engine_ops(dory_engine *e) {
    Ops o;
    o.aggregate = [e](const dory_chunk &c) { return dory_aggregate(e, &c); };
    o.apply_vertex = [e](const dory_chunk &c) { return dory_apply_vertex(e, &c); };
    o.scatter = [e](const dory_chunk &c) { return dory_scatter(e, &c); };
    o.apply_edge = [e](const dory_chunk &c) { return dory_apply_edge(e, &c); };
    o.predict = [e](const dory_chunk &c) { return dory_predict(e, &c); };
    o.inc_layer = [e](dory_chunk &c) { return dory_inc_layer(e, &c); };
    o.apply_update = [e](uint32_t l) { return dory_apply_update(e, l); };
    o.stats = [e](float *acc, float *loss, uint32_t *cnt) {
        dory_stats s{};
        const int rc = dory_get_stats(e, &s);
        *acc = s.acc_sum;
        *loss = s.loss_sum;
        *cnt = s.val_rows;
        return rc;
    };
    o.sync = [e] { dory_sync(e); };
    return o;
}

}  // namespace saga
