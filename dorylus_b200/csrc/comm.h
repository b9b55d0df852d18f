// Ghost-row exchange between the partitions of one 8-GPU box (internal).
//
// Replaces the reference's ZeroMQ data channel (commmanager/commmanager.cpp:214-279), the per-row
// (gvid, features) messages of Engine::verticesPushOut (engine/utils.cpp:623-650), the std::map
// lookups of ghostReceiverGCN (engine/ops/gcn_ops.cpp:310-318), the per-message ACKs and the
// scatter barrier (ops/pipeline.cpp:262-281) with one all-to-all-v over NVLink per (layer, dir).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

namespace dory {

class Comm {
public:
    Comm() = default;
    ~Comm();
    Comm(const Comm &) = delete;
    Comm &operator=(const Comm &) = delete;

    // All methods return "" on success or an error message.
    static std::string unique_id(void *id128);
    std::string init(const void *id128, int rank, int nranks, int device);
    // Send side: per peer, the local row ids to ship (graph.<id>.bin send lists).
    std::string set_send_lists(int dir, const std::vector<std::vector<uint32_t>> &ids, uint32_t maxld,
                               cudaStream_t s);
    // Receive side: per peer, the ghost slots its rows land in, in arrival order.
    std::string set_recv_slots(int dir, int peer, const uint32_t *slots, uint32_t n, uint32_t maxld,
                               cudaStream_t s);
    // Ships rows of `local` ([V x ld]) to the peers and fills `ghost` ([G x ld]).
    std::string exchange(int dir, const float *local, float *ghost, uint32_t ld, cudaStream_t s, int &launches);
    std::string allreduce_sum(float *buf, size_t n, cudaStream_t s);

    // ---- peer-memory path: pack + ship in ONE kernel that stores rows straight into the peers'
    // ghost blocks over NVLink (no staging buffer, no unpack); NCCL only provides the two barriers.
    // Send side additionally needs, per peer, the ghost slot each shipped row lands in.
    std::string set_send_slots(int dir, int peer, const uint32_t *slots, uint32_t n);
    // peerGhost[q]: device pointer (mapped with cudaIpcOpenMemHandle) to peer q's ghost block for
    // this exchange; entries for q == rank are ignored.
    // pre_barrier = false skips the barrier in front of the stores (see exchange_p2p).
    // cols: data columns of a row (0 = the whole pitch): only those cross NVLink, the padding stays zero.
    // aux: the barriers go to the second communicator (has_aux()), for an exchange issued on its own stream
    // while collectives of the first one (dW all-reduce) are issued on the compute stream.
    std::string exchange_p2p(int dir, const float *local, float *const *peerGhost, uint32_t ld, cudaStream_t s,
                             int &launches, bool pre_barrier = true, uint32_t cols = 0, bool aux = false);
    bool has_aux() const { return nccl2_ != nullptr; }
    // 0: 4 rows per warp (default), 1 / 2: that many rows per warp, 9: one row + system fence per warp
    void set_p2p_variant(int v) { p2p_variant_ = v; }
    // CTAs per SM of the store kernel when it runs beside an aggregation (aux = true)
    void set_p2p_ctas_per_sm(int v) { p2p_ctas_per_sm_ = v; }
    bool p2p_ready(int dir) const;
    // Largest ghost slot any row shipped to `peer` lands in (+1); 0 when nothing is shipped there.
    uint32_t send_slot_bound(int dir, int peer) const;
    int rank() const { return rank_; }
    int nranks() const { return nranks_; }

private:
    struct Plan {
        std::vector<uint32_t> sendCount, sendOff;  // per peer, rows
        std::vector<uint32_t> recvCount, recvOff;
        std::vector<std::vector<uint32_t>> recvSlots;  // host copy per peer
        uint32_t sendTotal = 0, recvTotal = 0;
        uint32_t *dSendIds = nullptr;    // concatenated send ids, peer order
        uint32_t *dRecvSlots = nullptr;  // concatenated recv slots, peer order
        float *sendStage = nullptr, *recvStage = nullptr;
        size_t sendStageFloats = 0, recvStageFloats = 0;
        bool recvIdentity = false;  // concatenated slots == 0..G-1: receive straight into the ghost block
        bool recvDirty = true;
        // peer-memory path
        std::vector<std::vector<uint32_t>> sendSlots;  // per peer, ghost slot on that peer per shipped row
        uint32_t *dSendSlots = nullptr;                // concatenated, same order as dSendIds
        uint8_t *dSendPeer = nullptr;                  // peer id of every shipped row
        uint32_t *dSendOrder = nullptr;                // issue order that interleaves the peers
        bool sendSlotsDirty = true;
    };
    float *barrier_buf_ = nullptr;
    int p2p_variant_ = 0;
    int p2p_ctas_per_sm_ = 1;
    std::string finalize_recv(Plan &p, uint32_t maxld, cudaStream_t s);

    void *nccl_ = nullptr;   // ncclComm_t
    void *nccl2_ = nullptr;  // ncclCommSplit of it (same ranks): the exchange stream's barriers
    std::string allreduce_on(void *comm, float *buf, size_t n, cudaStream_t s);
    int rank_ = 0, nranks_ = 1, device_ = 0;
    Plan plan_[2];
};

}  // namespace dory
