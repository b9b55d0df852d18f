// Ghost-row exchange over NCCL (send/recv grouped into one all-to-all-v per exchange).
// NCCL is bound at run time with dlopen so that the library loads on machines without NCCL and
// shares the copy the host process (e.g. torch) has already loaded.
#include "comm.h"

#include <dlfcn.h>
#include <nccl.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace dory {
namespace {

struct NcclApi {
    void *handle = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    decltype(&ncclAllReduce) AllReduce = nullptr;
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclCommSplit) CommSplit = nullptr;  // optional (NCCL >= 2.18)
    std::string error;
};

NcclApi &api() {
    static NcclApi a;
    static std::once_flag once;
    std::call_once(once, [] {
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            a.handle = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.handle) break;
        }
        if (!a.handle) {
            a.error = std::string("cannot dlopen libnccl.so.2: ") + dlerror();
            return;
        }
#define DORY_SYM(field, sym)                                                   \
    a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, #sym));      \
    if (!a.field) {                                                            \
        a.error = "libnccl is missing symbol " #sym;                           \
        return;                                                                \
    }
        DORY_SYM(GetUniqueId, ncclGetUniqueId)
        DORY_SYM(CommInitRank, ncclCommInitRank)
        DORY_SYM(CommDestroy, ncclCommDestroy)
        DORY_SYM(GroupStart, ncclGroupStart)
        DORY_SYM(GroupEnd, ncclGroupEnd)
        DORY_SYM(Send, ncclSend)
        DORY_SYM(Recv, ncclRecv)
        DORY_SYM(AllReduce, ncclAllReduce)
        DORY_SYM(GetErrorString, ncclGetErrorString)
#undef DORY_SYM
        a.CommSplit = reinterpret_cast<decltype(a.CommSplit)>(dlsym(a.handle, "ncclCommSplit"));
    });
    return a;
}

#define NC(call)                                                                         \
    do {                                                                                 \
        ncclResult_t _r = (call);                                                        \
        if (_r != ncclSuccess) return std::string(#call " failed: ") + api().GetErrorString(_r); \
    } while (0)
#define CUS(call)                                                                        \
    do {                                                                                 \
        cudaError_t _c = (call);                                                         \
        if (_c != cudaSuccess) return std::string(#call " failed: ") + cudaGetErrorString(_c); \
    } while (0)

__global__ void scatter_rows_kernel(const float4 *__restrict__ src, const uint32_t *__restrict__ slots,
                                    uint32_t n, float4 *__restrict__ dst, uint32_t ld4) {
    const uint32_t r = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (r >= n) return;
    const float4 *s = src + (size_t)r * ld4;
    float4 *d = dst + (size_t)slots[r] * ld4;
    for (uint32_t c = threadIdx.x & 31; c < ld4; c += 32) d[c] = s[c];
}

static_assert(sizeof(ncclUniqueId) == 128, "DORY_UNIQUE_ID_BYTES must match ncclUniqueId");

}  // namespace

Comm::~Comm() {
    for (Plan &p : plan_) {
        if (p.dSendIds) cudaFree(p.dSendIds);
        if (p.dRecvSlots) cudaFree(p.dRecvSlots);
        if (p.sendStage) cudaFree(p.sendStage);
        if (p.recvStage) cudaFree(p.recvStage);
        if (p.dSendSlots) cudaFree(p.dSendSlots);
        if (p.dSendPeer) cudaFree(p.dSendPeer);
        if (p.dSendOrder) cudaFree(p.dSendOrder);
    }
    if (barrier_buf_) cudaFree(barrier_buf_);
    if (nccl2_ && api().CommDestroy) api().CommDestroy(static_cast<ncclComm_t>(nccl2_));
    if (nccl_ && api().CommDestroy) api().CommDestroy(static_cast<ncclComm_t>(nccl_));
}

std::string Comm::unique_id(void *id128) {
    NcclApi &a = api();
    if (!a.error.empty()) return a.error;
    ncclUniqueId id;
    NC(a.GetUniqueId(&id));
    std::memcpy(id128, &id, sizeof id);
    return "";
}

std::string Comm::init(const void *id128, int rank, int nranks, int device) {
    NcclApi &a = api();
    if (!a.error.empty()) return a.error;
    rank_ = rank;
    nranks_ = nranks;
    device_ = device;
    CUS(cudaSetDevice(device));
    ncclUniqueId id;
    std::memcpy(&id, id128, sizeof id);
    ncclComm_t c;
    NC(a.CommInitRank(&c, nranks, id, rank));
    nccl_ = c;
    if (a.CommSplit && !std::getenv("DORY_NO_AUX_COMM")) {  // second communicator over the same ranks (collective call)
        ncclComm_t c2 = nullptr;
        const ncclResult_t r = a.CommSplit(c, 0, rank, &c2, nullptr);
        if (r == ncclSuccess && c2) nccl2_ = c2;
        else if (rank == 0) fprintf(stderr, "[dorylus_b200] ncclCommSplit: %s -- exchanges stay on the compute stream\n", a.GetErrorString(r));
    } else if (rank == 0 && !a.CommSplit) {
        fprintf(stderr, "[dorylus_b200] this NCCL has no ncclCommSplit -- exchanges stay on the compute stream\n");
    }
    if (rank == 0 && std::getenv("DORY_VERBOSE")) fprintf(stderr, "[dorylus_b200] second communicator: %s\n", nccl2_ ? "yes" : "no");
    for (Plan &p : plan_) {
        p.sendCount.assign(nranks, 0);
        p.sendOff.assign(nranks, 0);
        p.recvCount.assign(nranks, 0);
        p.recvOff.assign(nranks, 0);
        p.recvSlots.assign(nranks, {});
    }
    return "";
}

std::string Comm::set_send_lists(int dir, const std::vector<std::vector<uint32_t>> &ids, uint32_t maxld,
                                 cudaStream_t s) {
    Plan &p = plan_[dir];
    std::vector<uint32_t> flat;
    for (int q = 0; q < nranks_; ++q) {
        p.sendOff[q] = (uint32_t)flat.size();
        p.sendCount[q] = q == rank_ ? 0 : (uint32_t)ids[q].size();
        if (q != rank_) flat.insert(flat.end(), ids[q].begin(), ids[q].end());
    }
    p.sendTotal = (uint32_t)flat.size();
    CUS(cudaMalloc(&p.dSendIds, std::max<size_t>(flat.size(), 1) * 4));
    if (!flat.empty()) CUS(cudaMemcpyAsync(p.dSendIds, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, s));
    p.sendStageFloats = (size_t)std::max<uint32_t>(p.sendTotal, 1) * maxld;
    CUS(cudaMalloc(&p.sendStage, p.sendStageFloats * 4));
    CUS(cudaStreamSynchronize(s));
    return "";
}

std::string Comm::set_recv_slots(int dir, int peer, const uint32_t *slots, uint32_t n, uint32_t, cudaStream_t) {
    Plan &p = plan_[dir];
    if (peer == rank_ && n) return "a partition does not receive ghost rows from itself";
    p.recvSlots[peer].assign(slots, slots + n);
    p.recvDirty = true;
    return "";
}

std::string Comm::finalize_recv(Plan &p, uint32_t maxld, cudaStream_t s) {
    std::vector<uint32_t> flat;
    for (int q = 0; q < nranks_; ++q) {
        p.recvOff[q] = (uint32_t)flat.size();
        p.recvCount[q] = (uint32_t)p.recvSlots[q].size();
        flat.insert(flat.end(), p.recvSlots[q].begin(), p.recvSlots[q].end());
    }
    p.recvTotal = (uint32_t)flat.size();
    p.recvIdentity = true;
    for (uint32_t i = 0; i < flat.size(); ++i)
        if (flat[i] != i) {
            p.recvIdentity = false;
            break;
        }
    if (p.dRecvSlots) cudaFree(p.dRecvSlots), p.dRecvSlots = nullptr;
    if (p.recvStage) cudaFree(p.recvStage), p.recvStage = nullptr;
    CUS(cudaMalloc(&p.dRecvSlots, std::max<size_t>(flat.size(), 1) * 4));
    if (!flat.empty()) CUS(cudaMemcpyAsync(p.dRecvSlots, flat.data(), flat.size() * 4, cudaMemcpyHostToDevice, s));
    if (!p.recvIdentity) {
        p.recvStageFloats = (size_t)std::max<uint32_t>(p.recvTotal, 1) * maxld;
        CUS(cudaMalloc(&p.recvStage, p.recvStageFloats * 4));
    }
    CUS(cudaStreamSynchronize(s));
    p.recvDirty = false;
    return "";
}

std::string Comm::exchange(int dir, const float *local, float *ghost, uint32_t ld, cudaStream_t s, int &launches) {
    NcclApi &a = api();
    Plan &p = plan_[dir];
    if (p.recvDirty) {
        // staging is sized for the widest layer the caller will ever exchange; ld of this call bounds it
        std::string m = finalize_recv(p, (uint32_t)(p.sendStageFloats / std::max<uint32_t>(p.sendTotal, 1)), s);
        if (!m.empty()) return m;
    }
    if ((size_t)p.sendTotal * ld > p.sendStageFloats) return "exchange: row pitch exceeds staging capacity";
    launches = 0;
    if (p.sendTotal) {  // pack: one launch for all peers
        if (launch_gather_rows(local, p.dSendIds, p.sendTotal, p.sendStage, ld, s) < 0) return "pack kernel launch failed";
        ++launches;
    }
    float *recvBase = p.recvIdentity ? ghost : p.recvStage;
    ncclComm_t c = static_cast<ncclComm_t>(nccl_);
    NC(a.GroupStart());
    for (int q = 0; q < nranks_; ++q) {
        if (q == rank_) continue;
        if (p.sendCount[q])
            NC(a.Send(p.sendStage + (size_t)p.sendOff[q] * ld, (size_t)p.sendCount[q] * ld, ncclFloat, q, c, s));
        if (p.recvCount[q])
            NC(a.Recv(recvBase + (size_t)p.recvOff[q] * ld, (size_t)p.recvCount[q] * ld, ncclFloat, q, c, s));
    }
    NC(a.GroupEnd());
    if (!p.recvIdentity && p.recvTotal) {
        scatter_rows_kernel<<<(p.recvTotal + 7) / 8, 256, 0, s>>>(reinterpret_cast<const float4 *>(p.recvStage),
                                                                  p.dRecvSlots, p.recvTotal,
                                                                  reinterpret_cast<float4 *>(ghost), ld / 4);
        if (cudaGetLastError() != cudaSuccess) return "unpack kernel launch failed";
        ++launches;
    }
    return "";
}

std::string Comm::allreduce_on(void *comm, float *buf, size_t n, cudaStream_t s) {
    NcclApi &a = api();
    NC(a.AllReduce(buf, buf, n, ncclFloat, ncclSum, static_cast<ncclComm_t>(comm), s));
    return "";
}

std::string Comm::allreduce_sum(float *buf, size_t n, cudaStream_t s) { return allreduce_on(nccl_, buf, n, s); }

// ------------------------------------------------------------------ peer-memory exchange
namespace {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
    float4 *p[kMaxPeers];
};

// One warp per shipped row: read the local row once, store it into the owner-of-the-ghost's HBM
// through the NVLink-mapped pointer.  `order` interleaves the peers so that all links are busy from
// the first wave on.  RPW rows per warp: all their loads are issued before the first remote store
// (a 512 B row is one float4 per lane, so RPW rows = RPW independent loads in flight per lane).
// No fence inside the kernel: the stores only have to be visible to the peer once the collective
// that follows the kernel on this stream has completed, and kernel completion orders them before it.
template <int RPW>
__global__ void __launch_bounds__(256)
p2p_scatter_kernel(const float4 *__restrict__ local, const uint32_t *__restrict__ ids,
                   const uint32_t *__restrict__ slots, const uint8_t *__restrict__ peer,
                   const uint32_t *__restrict__ order, uint32_t n, PeerPtrs pp, uint32_t ld4, uint32_t n4) {
    const uint32_t lane = threadIdx.x & 31;
    // grid-stride over the shipped rows: a launch that shares the chip with an aggregation (overlap mode) is
    // sized to a CTA per SM -- the NVLink stores are posted, a few hundred warps keep the links busy -- instead
    // of filling every SM slot with CTAs that mostly wait for the fabric
    for (uint32_t v0 = (blockIdx.x * 8 + (threadIdx.x >> 5)) * RPW; v0 < n; v0 += gridDim.x * 8 * RPW) {
        const float4 *s[RPW];
        float4 *d[RPW];
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            s[k] = nullptr;
            d[k] = nullptr;
            if (v0 + k < n) {
                const uint32_t r = order[v0 + k];
                s[k] = local + (size_t)ids[r] * ld4;
                d[k] = pp.p[peer[r]] + (size_t)slots[r] * ld4;
            }
        }
        for (uint32_t c = lane; c < n4; c += 32) {  // data columns only: the padding of a row stays zero on both sides
            float4 x[RPW];
#pragma unroll
            for (int k = 0; k < RPW; ++k)
                if (s[k]) x[k] = __ldg(s[k] + c);
#pragma unroll
            for (int k = 0; k < RPW; ++k)
                if (s[k]) d[k][c] = x[k];
        }
    }
}

// A warp per shipped row with a system-scope fence per warp (the first version; kept selectable
// for the comparison in profiles/).
__global__ void __launch_bounds__(256)
p2p_scatter_fenced_kernel(const float4 *__restrict__ local, const uint32_t *__restrict__ ids,
                          const uint32_t *__restrict__ slots, const uint8_t *__restrict__ peer,
                          const uint32_t *__restrict__ order, uint32_t n, PeerPtrs pp, uint32_t ld4) {
    const uint32_t v = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (v >= n) return;
    const uint32_t r = order[v];
    const float4 *s = local + (size_t)ids[r] * ld4;
    float4 *d = pp.p[peer[r]] + (size_t)slots[r] * ld4;
    for (uint32_t c = threadIdx.x & 31; c < ld4; c += 32) d[c] = s[c];
    __threadfence_system();
}

}  // namespace

std::string Comm::set_send_slots(int dir, int peer, const uint32_t *slots, uint32_t n) {
    Plan &p = plan_[dir];
    if (p.sendSlots.size() != (size_t)nranks_) p.sendSlots.assign(nranks_, {});
    if (n != p.sendCount[peer]) return "send slots: count differs from the send list of that peer";
    p.sendSlots[peer].assign(slots, slots + n);
    p.sendSlotsDirty = true;
    return "";
}

uint32_t Comm::send_slot_bound(int dir, int peer) const {
    const Plan &p = plan_[dir];
    if (p.sendSlots.size() != (size_t)nranks_) return 0;
    uint32_t b = 0;
    for (uint32_t sl : p.sendSlots[peer]) b = std::max(b, sl + 1);
    return b;
}

bool Comm::p2p_ready(int dir) const {
    const Plan &p = plan_[dir];
    if (nranks_ > kMaxPeers || p.sendSlots.size() != (size_t)nranks_) return false;
    for (int q = 0; q < nranks_; ++q)
        if (q != rank_ && p.sendSlots[q].size() != p.sendCount[q]) return false;
    return true;
}

std::string Comm::exchange_p2p(int dir, const float *local, float *const *peerGhost, uint32_t ld, cudaStream_t s,
                               int &launches, bool pre_barrier, uint32_t cols, bool aux) {
    Plan &p = plan_[dir];
    launches = 0;
    void *bcomm = aux && nccl2_ ? nccl2_ : nccl_;
    float *bbuf_off = nullptr;  // set below: the two communicators use different words of barrier_buf_
    if (!p2p_ready(dir)) return "peer-memory exchange: send slots not installed";
    if (p.sendSlotsDirty) {
        std::vector<uint32_t> slots(p.sendTotal), order;
        std::vector<uint8_t> peer(p.sendTotal);
        uint32_t longest = 0;
        for (int q = 0; q < nranks_; ++q) {
            if (q == rank_) continue;
            for (uint32_t i = 0; i < p.sendCount[q]; ++i) {
                slots[p.sendOff[q] + i] = p.sendSlots[q][i];
                peer[p.sendOff[q] + i] = (uint8_t)q;
            }
            longest = std::max(longest, p.sendCount[q]);
        }
        order.reserve(p.sendTotal);
        for (uint32_t i = 0; i < longest; ++i)
            for (int q = 0; q < nranks_; ++q)
                if (q != rank_ && i < p.sendCount[q]) order.push_back(p.sendOff[q] + i);
        if (p.dSendSlots) cudaFree(p.dSendSlots), p.dSendSlots = nullptr;
        if (p.dSendPeer) cudaFree(p.dSendPeer), p.dSendPeer = nullptr;
        if (p.dSendOrder) cudaFree(p.dSendOrder), p.dSendOrder = nullptr;
        const size_t n = std::max<uint32_t>(p.sendTotal, 1);
        CUS(cudaMalloc(&p.dSendSlots, n * 4));
        CUS(cudaMalloc(&p.dSendPeer, n));
        CUS(cudaMalloc(&p.dSendOrder, n * 4));
        if (p.sendTotal) {
            CUS(cudaMemcpyAsync(p.dSendSlots, slots.data(), n * 4, cudaMemcpyHostToDevice, s));
            CUS(cudaMemcpyAsync(p.dSendPeer, peer.data(), n, cudaMemcpyHostToDevice, s));
            CUS(cudaMemcpyAsync(p.dSendOrder, order.data(), n * 4, cudaMemcpyHostToDevice, s));
        }
        CUS(cudaStreamSynchronize(s));
        p.sendSlotsDirty = false;
    }
    if (!barrier_buf_) {
        CUS(cudaMalloc(&barrier_buf_, 32));
        CUS(cudaMemsetAsync(barrier_buf_, 0, 32, s));
    }
    bbuf_off = barrier_buf_ + (bcomm == nccl_ ? 0 : 4);
    // barrier 1: every peer has finished reading the ghost block we are about to overwrite.  The
    // caller elides it when a collective already separates those reads from this call (every rank
    // runs the same operator sequence, so a collective that follows MY reads follows the peers' too).
    if (pre_barrier) {
        std::string m = allreduce_on(bcomm, bbuf_off, 1, s);
        if (!m.empty()) return m;
    }
    if (p.sendTotal) {
        PeerPtrs pp{};
        for (int q = 0; q < nranks_; ++q) pp.p[q] = q == rank_ ? nullptr : reinterpret_cast<float4 *>(peerGhost[q]);
        const float4 *l4 = reinterpret_cast<const float4 *>(local);
        const uint32_t n = p.sendTotal;
        const uint32_t n4 = cols ? std::min(ld / 4, (cols + 3) / 4) : ld / 4;  // float4 per row that carry data
        static int sms = 0;
        if (!sms) {
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device_);
            if (sms <= 0) sms = 148;
        }
        const uint32_t cap = aux ? (uint32_t)sms * (uint32_t)std::max(1, p2p_ctas_per_sm_) : 0xffffffffu;
        auto grid = [n, cap](uint32_t rpw) { return std::min(cap, (n + 8 * rpw - 1) / (8 * rpw)); };
        switch (p2p_variant_) {
        case 1: p2p_scatter_kernel<1><<<grid(1), 256, 0, s>>>(l4, p.dSendIds, p.dSendSlots, p.dSendPeer, p.dSendOrder, n, pp, ld / 4, n4); break;
        case 2: p2p_scatter_kernel<2><<<grid(2), 256, 0, s>>>(l4, p.dSendIds, p.dSendSlots, p.dSendPeer, p.dSendOrder, n, pp, ld / 4, n4); break;
        case 9: p2p_scatter_fenced_kernel<<<grid(1), 256, 0, s>>>(l4, p.dSendIds, p.dSendSlots, p.dSendPeer, p.dSendOrder, n, pp, ld / 4); break;
        default: p2p_scatter_kernel<4><<<grid(4), 256, 0, s>>>(l4, p.dSendIds, p.dSendSlots, p.dSendPeer, p.dSendOrder, n, pp, ld / 4, n4); break;
        }
        if (cudaGetLastError() != cudaSuccess) return "peer-memory scatter kernel launch failed";
        ++launches;
    }
    // barrier 2: every peer's stores into OUR ghost block are complete (their kernels have finished)
    return allreduce_on(bcomm, bbuf_off + 1, 1, s);
}

}  // namespace dory
