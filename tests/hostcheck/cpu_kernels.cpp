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

#ifndef DORY_LAUNCHCHECK  // the launch-check build links the product's own spmm.cu / dense.cu / gat.cu objects
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

// entry points spmm.cu gained in round 2 (engine.cu references them): same scalar statement / always-true
int launch_spmm(const SpmmArgs &a, cudaStream_t);
int launch_spmm_rows(const SpmmArgs &a, cudaStream_t s) { return launch_spmm(a, s); }
bool spmm_shape_supported(int, int) { return true; }

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
                float bv = g.transB ? g.B[(size_t)n * g.ldb + k] : g.B[k * g.ldb + n];
                if (g.Bh) bv = bv * (1.f - g.Bh[k * g.ldb + n] * g.Bh[k * g.ldb + n]);  // fused tanh' operand (transA, !transB)
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

// round 2: the fused logits + soft-max kernel is not emulated ("shape does not qualify" -> GEMM, then the statement below)
int launch_gemm_softmax_ce(const float *, uint32_t, const float *, uint32_t, uint64_t, const SoftmaxCEArgs &, cudaStream_t) { return 0; }
int launch_softmax_stats(const SoftmaxCEArgs &, cudaStream_t) { return 0; }

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

#endif  // !DORY_LAUNCHCHECK

// The shared-memory-staged kernel (spmm_tile.cu; product option "tile", off by default) is not emulated:
// "no kernel for this shape" sends the engine back to the gather path.
int launch_spmm_tile(const SpmmArgs &, const TilePlanDev &, cudaStream_t) { return 0; }
size_t tile_smem_bytes(uint32_t, uint32_t, uint32_t, bool, int) { return 0; }
size_t tile_edge_smem_bytes(uint64_t, uint32_t) { return 0; }


// ------------------------------------------------------------------ tcgen05 paths: "shape not supported"
int launch_gemm_nt_tc(const float *, uint32_t, uint64_t, const float *, uint32_t, uint32_t, float *, uint32_t, int, cudaStream_t) { return 0; }
int launch_gemm_tc_softmax(const float *, uint32_t, const float *, uint32_t, uint32_t, const SoftmaxCEArgs &, int, cudaStream_t) { return 0; }
int launch_gemm_tc(const float *, uint32_t, uint64_t, const float *, uint32_t, uint32_t, float *, float *, uint32_t, int,
                   cudaStream_t, int) {
    return 0;
}
int launch_gemm_tn_tc(const float *, uint32_t, uint32_t, const float *, uint32_t, uint64_t, float *, uint32_t, float *,
                      size_t, cudaStream_t) {
    return 0;
}

#ifndef DORY_LAUNCHCHECK
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

#endif  // !DORY_LAUNCHCHECK

#ifndef DORY_COMMCHECK  // the comm-check build links the product's own comm.cu object (fake_nccl.cpp)
// ------------------------------------------------------------------ Comm: collectives through a directory
// Stand-in for comm.cu's NCCL communicator.  Ranks may be threads of one process (the Python tests) or
// separate processes (host/run_onnode.sh): the "unique id" is the path of a fresh directory, a barrier
// is one marker file per rank and generation, and the wire is a file per (sender, receiver).  It
// implements the pack -> all-to-all-v -> unpack path (Comm::exchange) and the dW all-reduce; the
// peer-memory path needs CUDA IPC and reports itself unavailable, so the engine falls back to it.
}  // namespace dory

#include <chrono>
#include <cstdio>
#include <fstream>
#include <thread>

#include <sys/stat.h>
#include <unistd.h>

namespace dory {
namespace {
struct World {
    std::string dir;
    uint64_t gen = 0;
    int rank = 0, nranks = 1;
    std::string path(const char *kind, uint64_t g, int a, int b = -1) const {
        return dir + "/" + kind + "." + std::to_string(g) + "." + std::to_string(a) + (b >= 0 ? "." + std::to_string(b) : "");
    }
    static void put(const std::string &p, const void *data, size_t n) {
        const std::string tmp = p + ".tmp";
        FILE *f = std::fopen(tmp.c_str(), "wb");
        if (f) {
            if (n) std::fwrite(data, 1, n, f);
            std::fclose(f);
            std::rename(tmp.c_str(), p.c_str());
        }
    }
    static bool get(const std::string &p, std::vector<char> &out) {
        std::ifstream f(p, std::ios::binary | std::ios::ate);
        if (!f.good()) return false;
        out.resize((size_t)f.tellg());
        f.seekg(0);
        if (!out.empty()) f.read(out.data(), (std::streamsize)out.size());
        return true;
    }
    bool barrier() {  // false on timeout (a peer died)
        const uint64_t g = ++gen;
        put(path("b", g, rank), "", 0);
        const auto t0 = std::chrono::steady_clock::now();
        for (int q = 0; q < nranks; ++q) {
            struct stat st;
            while (::stat(path("b", g, q).c_str(), &st) != 0) {
                if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.0) return false;
                std::this_thread::sleep_for(std::chrono::microseconds(200));
            }
        }
        if (g > 2) ::unlink(path("b", g - 2, rank).c_str());  // everybody is past generation g-2 by now
        return true;
    }
};
}  // namespace

Comm::~Comm() {
    for (auto &p : plan_) std::free(p.dSendIds);
    delete static_cast<World *>(nccl_);
}

std::string Comm::unique_id(void *id128) {
    char tmpl[] = "/tmp/dory_hostcheck_XXXXXX";
    if (!::mkdtemp(tmpl)) return "hostcheck comm: mkdtemp failed";
    std::memset(id128, 0, 128);
    std::memcpy(id128, tmpl, sizeof tmpl);
    return "";
}

std::string Comm::init(const void *id128, int rank, int nranks, int device) {
    World *w = new World();
    w->dir = std::string(static_cast<const char *>(id128));
    w->rank = rank;
    w->nranks = nranks;
    nccl_ = w;
    rank_ = rank;
    nranks_ = nranks;
    device_ = device;
    return w->barrier() ? "" : "hostcheck comm: a rank did not show up";
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
    const Plan &p = plan_[dir];
    const uint64_t g = w.gen + 1;  // the generation of the barrier that follows the writes
    std::vector<float> pack;
    for (int q = 0; q < nranks_; ++q) {
        if (q == rank_) continue;
        pack.resize((size_t)p.sendCount[q] * ld);
        for (uint32_t i = 0; i < p.sendCount[q]; ++i)
            std::memcpy(pack.data() + (size_t)i * ld, local + (size_t)p.dSendIds[p.sendOff[q] + i] * ld, sizeof(float) * ld);
        World::put(w.path("x", g, rank_, q), pack.data(), pack.size() * sizeof(float));
    }
    if (!w.barrier()) return "hostcheck comm: exchange timed out";
    std::string err;
    std::vector<char> in;
    for (int q = 0; q < nranks_; ++q) {
        if (q == rank_) continue;
        const std::vector<uint32_t> &slots = p.recvSlots[q];
        if (!World::get(w.path("x", g, q, rank_), in) || in.size() != slots.size() * ld * sizeof(float)) {
            err = "hostcheck comm: send list and receive plan disagree";
            continue;
        }
        for (size_t i = 0; i < slots.size(); ++i)
            std::memcpy(ghost + (size_t)slots[i] * ld, in.data() + i * ld * sizeof(float), sizeof(float) * ld);
    }
    if (!w.barrier()) return "hostcheck comm: exchange timed out";
    for (int q = 0; q < nranks_; ++q)
        if (q != rank_) ::unlink(w.path("x", g, rank_, q).c_str());
    launches += 2;
    return err;
}

std::string Comm::allreduce_sum(float *buf, size_t n, cudaStream_t) {
    World &w = *static_cast<World *>(nccl_);
    const uint64_t g = w.gen + 1;
    World::put(w.path("r", g, rank_), buf, n * sizeof(float));
    if (!w.barrier()) return "hostcheck comm: all-reduce timed out";
    std::vector<float> sum(n, 0.f);
    std::vector<char> in;
    for (int q = 0; q < nranks_; ++q) {  // rank order on every rank: identical bits everywhere
        if (!World::get(w.path("r", g, q), in) || in.size() != n * sizeof(float)) return "hostcheck comm: all-reduce size mismatch";
        const float *v = reinterpret_cast<const float *>(in.data());
        for (size_t i = 0; i < n; ++i) sum[i] += v[i];
    }
    if (!w.barrier()) return "hostcheck comm: all-reduce timed out";
    ::unlink(w.path("r", g, rank_).c_str());
    std::memcpy(buf, sum.data(), sizeof(float) * n);
    return "";
}

std::string Comm::set_send_slots(int, int, const uint32_t *, uint32_t) { return ""; }
std::string Comm::exchange_p2p(int, const float *, float *const *, uint32_t, cudaStream_t, int &, bool, uint32_t, bool) {
    return "hostcheck comm: the peer-memory path needs CUDA IPC";
}
bool Comm::p2p_ready(int) const { return false; }
uint32_t Comm::send_slot_bound(int, int) const { return 0; }

}  // namespace dory
#else
}  // namespace dory
#endif  // !DORY_COMMCHECK
