// TEST INFRASTRUCTURE ONLY -- the comm-check variant (tests/hostcheck/build.py: build_commcheck) links
// the product's REAL communicator (dorylus_b200/_obj/comm_cu.o: send / receive plans, staging, the grouped
// all-to-all-v, the peer-memory store kernel's plan) and gives it, inside this library only:
//   * dlopen / dlsym that resolve "libnccl.so.2" to the functions below,
//   * an NCCL whose ranks are THREADS of this process (a collective is a rendezvous of those threads,
//     the wire is a memcpy),
//   * CUDA IPC handles that carry the pointer itself (one address space),
//   * scalar statements of the three kernels comm.cu launches.
#include <cuda_runtime.h>

#include <condition_variable>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

namespace {

struct World {
    std::mutex m;
    std::condition_variable cv;
    int nranks = 0, arrived = 0;
    uint64_t generation = 0;
    // mail[src][dst] = (pointer, floats) posted by src's ncclSend to dst inside the current group
    std::vector<std::vector<std::pair<const float *, size_t>>> mail;
    std::vector<float *> buf;
    void barrier() {
        std::unique_lock<std::mutex> lk(m);
        const uint64_t g = generation;
        if (++arrived == nranks) {
            arrived = 0;
            ++generation;
            cv.notify_all();
        } else {
            cv.wait(lk, [&] { return generation != g; });
        }
    }
};
struct FakeComm {
    World *w;
    int rank;
};
struct Op {
    bool send;
    void *p;
    size_t count;
    int peer;
    FakeComm *c;
};
std::mutex g_mutex;
std::map<uint64_t, World> g_worlds;
uint64_t g_next = 1;
thread_local std::vector<Op> t_group;
thread_local int t_depth = 0;

struct UniqueId {
    char internal[128];
};

int run_group() {
    if (t_group.empty()) return 0;
    FakeComm *c = t_group[0].c;
    World &w = *c->w;
    for (const Op &o : t_group)
        if (o.send) w.mail[c->rank][o.peer] = {static_cast<const float *>(o.p), o.count};
    w.barrier();  // every send of this collective is posted
    int rc = 0;
    for (const Op &o : t_group)
        if (!o.send) {
            const auto &m = w.mail[o.peer][c->rank];
            if (m.second != o.count) rc = 5;  // ncclInvalidArgument: send and receive sizes disagree
            else std::memcpy(o.p, m.first, o.count * sizeof(float));
        }
    w.barrier();  // nobody reuses its send buffer while a peer still copies from it
    t_group.clear();
    return rc;
}

}  // namespace

extern "C" {

int ncclGetUniqueId(UniqueId *id) {
    std::lock_guard<std::mutex> lk(g_mutex);
    std::memset(id, 0, sizeof *id);
    const uint64_t v = g_next++;
    std::memcpy(id->internal, &v, sizeof v);
    return 0;
}
int ncclCommInitRank(void **comm, int nranks, UniqueId id, int rank) {
    uint64_t v;
    std::memcpy(&v, id.internal, sizeof v);
    World *w;
    {
        std::lock_guard<std::mutex> lk(g_mutex);
        w = &g_worlds[v];
        std::lock_guard<std::mutex> lk2(w->m);
        if (w->nranks == 0) {
            w->nranks = nranks;
            w->mail.assign(nranks, std::vector<std::pair<const float *, size_t>>(nranks, {nullptr, 0}));
            w->buf.assign(nranks, nullptr);
        }
        if (w->nranks != nranks || rank < 0 || rank >= nranks) return 5;
    }
    *comm = new FakeComm{w, rank};
    w->barrier();
    return 0;
}
int ncclCommDestroy(void *comm) {
    delete static_cast<FakeComm *>(comm);
    return 0;
}
int ncclGroupStart(void) {
    ++t_depth;
    return 0;
}
int ncclGroupEnd(void) { return --t_depth == 0 ? run_group() : 0; }
int ncclSend(const void *p, size_t count, int, int peer, void *comm, cudaStream_t) {
    t_group.push_back(Op{true, const_cast<void *>(p), count, peer, static_cast<FakeComm *>(comm)});
    return t_depth ? 0 : run_group();
}
int ncclRecv(void *p, size_t count, int, int peer, void *comm, cudaStream_t) {
    t_group.push_back(Op{false, p, count, peer, static_cast<FakeComm *>(comm)});
    return t_depth ? 0 : run_group();
}
int ncclAllReduce(const void *send, void *recv, size_t count, int, int, void *comm, cudaStream_t) {
    FakeComm *c = static_cast<FakeComm *>(comm);
    World &w = *c->w;
    w.buf[c->rank] = const_cast<float *>(static_cast<const float *>(send));
    w.barrier();
    std::vector<float> sum(count, 0.f);
    for (int q = 0; q < w.nranks; ++q)
        for (size_t i = 0; i < count; ++i) sum[i] += w.buf[q][i];
    w.barrier();
    std::memcpy(recv, sum.data(), count * sizeof(float));
    w.barrier();
    return 0;
}
const char *ncclGetErrorString(int) { return "hostcheck: emulated NCCL"; }

// ---- the dynamic loader, as seen from inside this library (-Bsymbolic)
void *dlopen(const char *name, int) {
    static int self;
    return name && std::strstr(name, "libnccl") ? &self : nullptr;
}
char *dlerror(void) {
    static char msg[] = "hostcheck: only libnccl can be dlopen'ed here";
    return msg;
}
void *dlsym(void *, const char *sym) {
    static const std::map<std::string, void *> table = {
        {"ncclGetUniqueId", (void *)&ncclGetUniqueId},   {"ncclCommInitRank", (void *)&ncclCommInitRank},
        {"ncclCommDestroy", (void *)&ncclCommDestroy},   {"ncclGroupStart", (void *)&ncclGroupStart},
        {"ncclGroupEnd", (void *)&ncclGroupEnd},         {"ncclSend", (void *)&ncclSend},
        {"ncclRecv", (void *)&ncclRecv},                 {"ncclAllReduce", (void *)&ncclAllReduce},
        {"ncclGetErrorString", (void *)&ncclGetErrorString},
    };
    auto it = table.find(sym);
    return it == table.end() ? nullptr : it->second;
}

}  // extern "C"

// ---- the kernels comm.cu launches, by their (mangled) names; called from fake_cudart.cpp's cudaLaunchKernel
namespace {
struct PeerPtrs {
    float *p[16];
};
}  // namespace

bool hostcheck_comm_kernel(const std::string &name, void **args) {
    if (name.find("scatter_rows_kernel") != std::string::npos && name.find("p2p") == std::string::npos) {
        // (const float4 *src, const uint32_t *slots, uint32_t n, float4 *dst, uint32_t ld4)
        const float *src = *static_cast<const float **>(args[0]);
        const uint32_t *slots = *static_cast<const uint32_t **>(args[1]);
        const uint32_t n = *static_cast<uint32_t *>(args[2]);
        float *dst = *static_cast<float **>(args[3]);
        const uint32_t ld = 4 * *static_cast<uint32_t *>(args[4]);
        for (uint32_t r = 0; r < n; ++r) std::memcpy(dst + (size_t)slots[r] * ld, src + (size_t)r * ld, sizeof(float) * ld);
        return true;
    }
    if (name.find("p2p_scatter") != std::string::npos) {
        // (const float4 *local, const uint32_t *ids, const uint32_t *slots, const uint8_t *peer,
        //  const uint32_t *order, uint32_t n, PeerPtrs pp, uint32_t ld4)
        const float *local = *static_cast<const float **>(args[0]);
        const uint32_t *ids = *static_cast<const uint32_t **>(args[1]);
        const uint32_t *slots = *static_cast<const uint32_t **>(args[2]);
        const uint8_t *peer = *static_cast<const uint8_t **>(args[3]);
        const uint32_t *order = *static_cast<const uint32_t **>(args[4]);
        const uint32_t n = *static_cast<uint32_t *>(args[5]);
        const PeerPtrs &pp = *static_cast<PeerPtrs *>(args[6]);
        const uint32_t ld = 4 * *static_cast<uint32_t *>(args[7]);
        for (uint32_t v = 0; v < n; ++v) {
            const uint32_t r = order[v];
            std::memcpy(pp.p[peer[r]] + (size_t)slots[r] * ld, local + (size_t)ids[r] * ld, sizeof(float) * ld);
        }
        return true;
    }
    return false;
}
