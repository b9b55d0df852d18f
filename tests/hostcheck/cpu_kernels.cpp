// TEST INFRASTRUCTURE ONLY -- scalar CPU statements of the launchers the engine calls
// (dorylus_b200/csrc/common.cuh, gat.cuh, gemm_tc.cuh, comm.h), with the SAME argument contracts as the
// CUDA translation units they stand in for (spmm.cu, dense.cu, gat.cu, gemm_tc.cu, comm.cu): padded row
// pitches, row lists, source windows, self modes, split workspaces.  Together with fake_cudart.cpp they
// let the product's engine object (engine_cu.o) run its host logic -- tensor tables, operator order,
// schedules, window passes, error paths -- on a machine without a GPU, against the oracle.
// They check the ENGINE, not the kernels: the CUDA kernels are checked on the GPU (-m gpu).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include "../../dorylus_b200/csrc/comm.h"
#include "../../dorylus_b200/csrc/common.cuh"
#include "../../dorylus_b200/csrc/gat.cuh"
#include "../../dorylus_b200/csrc/gemm_tc.cuh"

namespace dory {

// ------------------------------------------------------------------ aggregation (spmm.cu)
static void spmm_row(const SpmmArgs &a, uint32_t row) {
    const uint64_t pbase = (uint64_t)row * a.ptr_stride + a.ptr_off;
    const uint64_t e0 = a.ptrs[pbase], e1 = a.ptrs[pbase + a.ptr_span];
    const uint32_t n = a.nvec * 4;  // data columns (whole float4 units); padding columns are not touched
    std::vector<float> acc(n, 0.f);
    for (uint64_t e = e0; e < e1; ++e) {
        const float w = a.vals[e];
        const float *s = a.src + (size_t)a.idx[e] * a.ld;
        for (uint32_t c = 0; c < n; ++c) acc[c] += w * s[c];
    }
    float *o = a.out + (size_t)row * a.ld;
    const float *self = a.src + (size_t)row * a.ld;
    for (uint32_t c = 0; c < n; ++c) {
        float sv = 0.f;
        if (a.self_mode == SELF_NORM) sv = self[c] * a.selfw[row];
        else if (a.self_mode == SELF_ONE) sv = self[c];
        else if (a.self_mode == SELF_ACCUM) sv = o[c];
        o[c] = sv + acc[c];
    }
}

int launch_spmm(const SpmmArgs &a, cudaStream_t) {
    if (a.nvec > a.ld / 4 || a.ptr_span == 0 || a.ptr_off + a.ptr_span > a.ptr_stride) return -1;
    int launches = 0;
    if (a.heavy) {
        for (uint32_t i = 0; i < a.n_heavy; ++i) spmm_row(a, a.heavy[i]);
        launches += a.n_heavy ? 1 : 0;
    }
    if (a.light) {
        for (uint32_t i = 0; i < a.n_light; ++i) spmm_row(a, a.light[i]);
    } else {
        for (uint32_t i = 0; i < a.n_light; ++i) spmm_row(a, a.low + i);
    }
    launches += a.n_light ? 1 : 0;
    return launches;
}

// ------------------------------------------------------------------ dense apply (dense.cu)
int launch_gemm(const GemmArgs &g, cudaStream_t) {
    // C[M x N] = op(A) . op(B); transA: A stored [K x M]; transB: B stored [N x K]
    for (uint64_t m = 0; m < g.M; ++m)
        for (uint32_t n = 0; n < g.N; ++n) {
            double s = 0.0;  // any summation order is within the 1e-5 bar; double keeps this side exact
            for (uint64_t k = 0; k < g.K; ++k) {
                const float av = g.transA ? g.A[k * g.lda + m] : g.A[m * g.lda + k];
                const float bv = g.transB ? g.B[(size_t)n * g.ldb + k] : g.B[k * g.ldb + n];
                s += (double)av * bv;
            }
            g.C[m * g.ldc + n] = (float)s;
            if (g.epilogue == EPI_TANH && g.C2) g.C2[m * g.ldc + n] = std::tanh((float)s);
        }
    return 1;
}

int launch_tanh_backward(const float *aTg, const float *h, float *g, uint64_t n, cudaStream_t) {
    for (uint64_t i = 0; i < n / 4 * 4; ++i) g[i] = aTg[i] * (1.f - h[i] * h[i]);
    return n >= 4 ? 1 : 0;
}

int launch_tanh_forward(const float *z, float *h, uint64_t n, cudaStream_t) {
    for (uint64_t i = 0; i < n / 4 * 4; ++i) h[i] = std::tanh(z[i]);
    return n >= 4 ? 1 : 0;
}

int launch_softmax_ce(const SoftmaxCEArgs &a, cudaStream_t) {
    float acc = 0.f, loss = 0.f;
    std::vector<float> p(a.C);
    const uint64_t maskBeg = (uint64_t)a.trainEnd * a.C;
    for (uint32_t row = 0; row < a.V; ++row) {
        const float *z = a.z + (size_t)row * a.ld, *lab = a.lab + (size_t)row * a.ld;
        float mx = -INFINITY;
        for (uint32_t c = 0; c < a.C; ++c) mx = std::max(mx, z[c]);
        float sum = 0.f;
        for (uint32_t c = 0; c < a.C; ++c) sum += (p[c] = std::exp(z[c] - mx));
        const float denom = 1e-20f + sum;
        uint32_t pi = 0, li = 0;
        for (uint32_t c = 0; c < a.C; ++c) {
            p[c] /= denom;
            if (p[c] > p[pi]) pi = c;
            if (lab[c] > lab[li]) li = c;
        }
        if (row >= a.trainEnd && row < a.valEnd) {
            acc += lab[pi];
            loss -= std::log(p[li]);
        }
        for (uint32_t c = 0; c < a.C; ++c) {
            if (a.pred) a.pred[(size_t)row * a.ld + c] = p[c];
            const uint64_t flat = (uint64_t)row * a.C + c;
            const bool masked = a.strictMask ? row >= a.trainEnd : (flat >= maskBeg && flat < maskBeg + a.maskFloats);
            a.d[(size_t)row * a.ld + c] = ((masked ? lab[c] : p[c]) - lab[c]) / a.denom;
        }
    }
    a.stats[0] = acc;
    a.stats[1] = loss;
    return 2;
}

int launch_adam(float *w, const float *grad, float *m, float *v, size_t n, float lr_t, float beta1, float beta2,
                float eps, cudaStream_t) {
    for (size_t i = 0; i < n; ++i) {  // AdamOptimizer.cpp:36-48 (float / double mix as written there)
        const float gt = grad[i];
        m[i] = beta1 * m[i] + (1. - beta1) * gt;
        v[i] = beta2 * v[i] + (1. - beta2) * gt * gt;
        w[i] -= lr_t * m[i] / (std::sqrt(v[i]) + eps);
    }
    return 1;
}

int launch_fill(float *p, size_t n, float value, cudaStream_t) {
    std::fill(p, p + n, value);
    return 1;
}

int launch_fma_peak(float *, int, unsigned, cudaStream_t) { return 1; }

int launch_repitch(const float *src, uint32_t lds, float *dst, uint32_t ldd, uint64_t rows, uint32_t cols, cudaStream_t) {
    for (uint64_t r = 0; r < rows; ++r) std::memmove(dst + r * ldd, src + r * lds, sizeof(float) * cols);
    return 1;
}

int launch_gather_rows(const float *src, const uint32_t *ids, uint32_t n, float *dst, uint32_t ld, cudaStream_t) {
    for (uint32_t i = 0; i < n; ++i) std::memcpy(dst + (size_t)i * ld, src + (size_t)ids[i] * ld, sizeof(float) * ld);
    return 1;
}

// ------------------------------------------------------------------ tcgen05 paths: "shape not supported"
int launch_gemm_tc(const float *, uint32_t, uint64_t, const float *, uint32_t, uint32_t, float *, float *, uint32_t, int,
                   cudaStream_t) {
    return 0;
}
int launch_gemm_tn_tc(const float *, uint32_t, uint32_t, const float *, uint32_t, uint64_t, float *, uint32_t, float *,
                      size_t, cudaStream_t) {
    return 0;
}

// ------------------------------------------------------------------ GAT edge operators (gat.cu)
constexpr uint32_t kALd = 4;  // a_i is an F x 1 weight with row pitch 4
constexpr float kAlpha = 0.01f;

int launch_gat_edge_forward(const float *z, uint32_t ld, uint32_t F, const float *a, const uint64_t *colPtrs, uint32_t V,
                            float *az, float *A, cudaStream_t) {
    for (uint32_t v = 0; v < V; ++v) {
        float s = 0.f;
        for (uint32_t j = 0; j < F; ++j) s += z[(size_t)v * ld + j] * a[(size_t)j * kALd];
        for (uint64_t e = colPtrs[v]; e < colPtrs[v + 1]; ++e) {
            az[e] = s;
            A[e] = s > 0.f ? s : kAlpha * s;
        }
    }
    return V ? 1 : 0;
}

int launch_gat_edge_backward(const GatEdgeBackwardArgs &a, cudaStream_t) {
    if (a.V == 0) return 0;
    if (a.scratch_floats < (size_t)a.V + 2 * a.ld) return -1;
    std::vector<double> reduced(a.F, 0.0);
    for (uint32_t v = 0; v < a.V; ++v) {
        float t = 0.f, c = 0.f;
        for (uint32_t j = 0; j < a.F; ++j) t += a.grad[(size_t)v * a.ld + j] * a.a[(size_t)j * kALd];
        for (uint64_t e = a.colPtrs[v]; e < a.colPtrs[v + 1]; ++e) {
            const float d = a.az[e] > 0.f ? 1.f : kAlpha;
            a.dA[e] = t * d;
            c += d;
        }
        for (uint32_t j = 0; j < a.F; ++j) reduced[j] += (double)a.grad[(size_t)v * a.ld + j] * c;
    }
    for (uint32_t i = 0; i < a.F; ++i) {  // da = (z^T z) . dAct_reduce
        double s = 0.0;
        for (uint32_t j = 0; j < a.F; ++j) {
            double zz = 0.0;
            for (uint32_t v = 0; v < a.V; ++v) zz += (double)a.z[(size_t)v * a.ld + i] * a.z[(size_t)v * a.ld + j];
            s += zz * reduced[j];
        }
        a.da[(size_t)i * kALd] = (float)s;
    }
    return 6;
}

int launch_gat_predict(const float *logits, uint32_t ldl, const float *lab, float *grad, uint32_t ld, uint32_t C,
                       uint32_t low, uint32_t up, cudaStream_t) {
    for (uint32_t row = low; row < up; ++row) {
        const float *z = logits + (size_t)row * ldl;
        float mx = -INFINITY, sum = 0.f;
        for (uint32_t c = 0; c < C; ++c) mx = std::max(mx, z[c]);
        std::vector<float> p(C);
        for (uint32_t c = 0; c < C; ++c) sum += (p[c] = std::exp(z[c] - mx));
        for (uint32_t c = 0; c < C; ++c) grad[(size_t)row * ld + c] = p[c] / (1e-20f + sum) - lab[(size_t)row * ld + c];
    }
    return up > low ? 1 : 0;
}

// ------------------------------------------------------------------ Comm: ranks are THREADS of this process
// Stand-in for comm.cu's NCCL communicator: every rank is an engine driven by its own host thread, a
// collective is a rendezvous of those threads, and "the wire" is a memcpy between their buffers.  It
// implements the pack -> all-to-all-v -> unpack path (Comm::exchange) and the dW all-reduce; the
// peer-memory path needs CUDA IPC and reports itself unavailable, so the engine falls back to it.
}  // namespace dory

#include <condition_variable>
#include <map>
#include <mutex>

namespace dory {
namespace {
struct World {
    std::mutex m;
    std::condition_variable cv;
    int nranks = 0, arrived = 0;
    uint64_t generation = 0;
    std::vector<Comm *> members;
    std::vector<const float *> local;  // what each rank currently offers (exchange) / reduces (all-reduce)
    std::vector<float *> buf;
    std::vector<uint32_t> ld;
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
std::mutex g_worlds_mutex;
std::map<uint64_t, World> g_worlds;
uint64_t g_next_world = 1;
}  // namespace

Comm::~Comm() {
    for (auto &p : plan_) std::free(p.dSendIds);
}

std::string Comm::unique_id(void *id128) {
    std::lock_guard<std::mutex> lk(g_worlds_mutex);
    std::memset(id128, 0, 128);
    const uint64_t id = g_next_world++;
    std::memcpy(id128, &id, sizeof id);
    return "";
}

std::string Comm::init(const void *id128, int rank, int nranks, int device) {
    uint64_t id;
    std::memcpy(&id, id128, sizeof id);
    World *w;
    {
        std::lock_guard<std::mutex> lk(g_worlds_mutex);
        w = &g_worlds[id];
        std::lock_guard<std::mutex> lk2(w->m);
        if (w->nranks == 0) {
            w->nranks = nranks;
            w->members.assign(nranks, nullptr);
            w->local.assign(nranks, nullptr);
            w->buf.assign(nranks, nullptr);
            w->ld.assign(nranks, 0);
        }
        if (w->nranks != nranks || rank < 0 || rank >= nranks || w->members[rank]) return "hostcheck comm: inconsistent init";
        w->members[rank] = this;
    }
    nccl_ = w;
    rank_ = rank;
    nranks_ = nranks;
    device_ = device;
    w->barrier();
    return "";
}

std::string Comm::set_send_lists(int dir, const std::vector<std::vector<uint32_t>> &ids, uint32_t, cudaStream_t) {
    Plan &p = plan_[dir];
    p.sendCount.assign(nranks_, 0);
    p.sendOff.assign(nranks_, 0);
    p.recvSlots.assign(nranks_, {});
    p.sendTotal = 0;
    for (int q = 0; q < nranks_; ++q) {
        p.sendOff[q] = p.sendTotal;
        p.sendCount[q] = (uint32_t)ids[q].size();
        p.sendTotal += p.sendCount[q];
    }
    std::free(p.dSendIds);
    p.dSendIds = static_cast<uint32_t *>(std::malloc(4 * (size_t)std::max(1u, p.sendTotal)));
    for (int q = 0; q < nranks_; ++q)
        if (!ids[q].empty()) std::memcpy(p.dSendIds + p.sendOff[q], ids[q].data(), 4 * ids[q].size());
    return "";
}

std::string Comm::set_recv_slots(int dir, int peer, const uint32_t *slots, uint32_t n, uint32_t, cudaStream_t) {
    if (peer < 0 || peer >= nranks_) return "hostcheck comm: bad peer";
    plan_[dir].recvSlots[peer].assign(slots, slots + n);
    return "";
}

std::string Comm::exchange(int dir, const float *local, float *ghost, uint32_t ld, cudaStream_t, int &launches) {
    World &w = *static_cast<World *>(nccl_);
    w.local[rank_] = local;
    w.ld[rank_] = ld;
    w.barrier();  // every rank has published the tensor it ships
    std::string err;
    for (int q = 0; q < nranks_; ++q) {
        if (q == rank_) continue;
        const Plan &theirs = w.members[q]->plan_[dir];
        const std::vector<uint32_t> &slots = plan_[dir].recvSlots[q];
        if (theirs.sendCount[rank_] != slots.size() || w.ld[q] != ld) {
            err = "hostcheck comm: send list and receive plan disagree";
            continue;
        }
        const uint32_t *ids = theirs.dSendIds + theirs.sendOff[rank_];
        for (size_t i = 0; i < slots.size(); ++i)
            std::memcpy(ghost + (size_t)slots[i] * ld, w.local[q] + (size_t)ids[i] * ld, sizeof(float) * ld);
    }
    w.barrier();  // nobody overwrites its tensor while a peer still reads it
    launches += 2;
    return err;
}

std::string Comm::allreduce_sum(float *buf, size_t n, cudaStream_t) {
    World &w = *static_cast<World *>(nccl_);
    w.buf[rank_] = buf;
    w.barrier();
    std::vector<float> sum(n, 0.f);
    for (int q = 0; q < nranks_; ++q)  // rank order on every rank: identical bits everywhere
        for (size_t i = 0; i < n; ++i) sum[i] += w.buf[q][i];
    w.barrier();
    std::memcpy(buf, sum.data(), sizeof(float) * n);
    w.barrier();
    return "";
}

std::string Comm::set_send_slots(int, int, const uint32_t *, uint32_t) { return ""; }
std::string Comm::exchange_p2p(int, const float *, float *const *, uint32_t, cudaStream_t, int &, bool) {
    return "hostcheck comm: the peer-memory path needs CUDA IPC";
}
bool Comm::p2p_ready(int) const { return false; }

}  // namespace dory
