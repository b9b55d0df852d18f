// C++ host driver over the C ABI: what src/graph-server/main.cpp + Engine::init/run do for a
// synchronous single-partition run in CPU/GPU mode (reference engine/engine.cpp:40-167,223-314),
// minus the ZeroMQ / weight-server processes.  Reads the reference's dataset directory:
//
//   <datasetdir>/graph.bsnap.edges, graph.bsnap.parts   (preprocessed into graph.<id>.bin if absent,
//                                                        like engine.cpp:62-71)
//   --featuresfile features.bsnap   --labelsfile labels.bsnap   --layerfile <one width per line>
//
//   dorylus_b200_run --datasetdir D/ --featuresfile F --labelsfile L --layerfile C
//                    [--numepochs 10] [--lr 0.01] [--gnn GCN|GAT] [--undirected 0]
//                    [--pipeline 1 [--numlambdas N] [--targetacc A] [--switchthreshold T]] [--dry-run 1]
//                    [--apply-first 1]
//                    [--numnodes N --nodeid I [--device D] [--rendezvous DIR | --run-id ID] [--exchange p2p|nccl]
//                     [--rendezvous-timeout SECONDS]]
//
// --pipeline 1 drives the epochs through host/saga_pipeline.hpp -- the reference's chunk queues,
// priority order, barriers, early-stop state machine and <EM> report (engine/ops/pipeline.cpp) --
// instead of dory_epoch; --numlambdas is the number of chunks per partition (numLambdasForward).
// --apply-first 1 sets DORY_FLAG_APPLY_FIRST (GCN layers that narrow run A_hat.(in.W)); that schedule is
// driven by dory_epoch, not by the reference's queue order, so it excludes --pipeline 1.
// --numnodes N --nodeid I runs partition I of N (one process per GPU, like one graph server per
// machine in run/run-onnode; host/run_onnode.sh starts all N on one box).  The ranks need no network
// side channel: the receive plan of every exchange is computed from the partition images in the dataset
// directory (dory_ghost_slots), and the two things that must travel -- rank 0's NCCL id and, for the
// peer-memory exchange, every rank's CUDA IPC handles -- go through small files under --rendezvous
// (default <datasetdir>.rendezvous/).  Statistics are per partition, as every graph server logs its own.
// --dry-run 1 stops after the host-side half (preprocess, features / labels incl. the reference's
// feats<F0>.<id>.bin cache) and prints what it read: no GPU needed.
//
// Prints one line per epoch in the weight server's format (weightserver.cpp:258-262:
// "Epoch %u, acc: %.3f, loss: %.3f") plus the epoch time the graph server reports
// (ops/pipeline.cpp:117-136).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>

#include "../include/dorylus_b200.h"
#include "saga_pipeline.hpp"

namespace {

bool read_file(const std::string &path, std::vector<char> &out) {
    std::ifstream f(path, std::ios::binary | std::ios::ate);
    if (!f.good()) return false;
    out.resize((size_t)f.tellg());
    f.seekg(0);
    f.read(out.data(), (std::streamsize)out.size());
    return f.good() || f.eof();
}

[[noreturn]] void die(dory_engine *e, const char *what) {
    std::fprintf(stderr, "dorylus_b200_run: %s: %s\n", what, dory_last_error(e));
    std::exit(EXIT_FAILURE);
}

// ---- file rendezvous between the ranks of one box (N > 1): a blob appears atomically (write to a
// temporary name, rename) and readers poll for it.
void publish(const std::string &path, const void *data, size_t n) {
    const std::string tmp = path + ".tmp";
    FILE *f = std::fopen(tmp.c_str(), "wb");
    if (!f || std::fwrite(data, 1, n, f) != n) {
        std::fprintf(stderr, "dorylus_b200_run: cannot write %s\n", tmp.c_str());
        std::exit(EXIT_FAILURE);
    }
    std::fclose(f);
    if (std::rename(tmp.c_str(), path.c_str()) != 0) {
        std::fprintf(stderr, "dorylus_b200_run: cannot rename %s\n", tmp.c_str());
        std::exit(EXIT_FAILURE);
    }
}

double g_rendezvous_timeout_s = 3600.0;  // --rendezvous-timeout: a peer may still be preprocessing a large partition

void await(const std::string &path, std::vector<char> &out, size_t expect, double timeout_s = g_rendezvous_timeout_s) {
    const auto t0 = std::chrono::steady_clock::now();
    for (;;) {
        if (read_file(path, out) && out.size() == expect) return;
        if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
            std::fprintf(stderr, "dorylus_b200_run: timed out waiting for %s\n", path.c_str());
            std::exit(EXIT_FAILURE);
        }
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
    }
}

// (layer, name) of every ghost block that takes part in an exchange, in an order all ranks agree on
// (the Python mirror's Engine.ghost_tensors).
std::vector<std::pair<uint32_t, std::string>> ghost_tensors(const dory_engine *e, const dory_config &cfg) {
    std::vector<std::pair<uint32_t, std::string>> out;
    const uint32_t L = cfg.n_layers;
    if (cfg.gnn_type == DORY_GCN) {
        out.push_back({0, "fg"});
        for (uint32_t l = 0; l < L; ++l) {
            int af = 0;
            dory_layer_schedule(e, l, &af);
            if (af) {
                out.push_back({l, "fg_t"});
                out.push_back({l, "bg_g"});
            } else if (l > 0) {
                out.push_back({l, "fg"});
                out.push_back({l - 1, "bg"});
            }
        }
    } else {
        for (uint32_t l = 0; l < L; ++l) out.push_back({l, "fg_z"});
        for (uint32_t l = 0; l < L; ++l) out.push_back({l, "bg_d"});
    }
    return out;
}

// Slots of one (receiver, sender, direction) from the two partition images (host only).
std::vector<uint32_t> plan_slots(const std::vector<char> &recvImage, uint32_t recvId, const std::vector<char> &sendImage, uint32_t dir) {
    uint32_t n = 0;
    if (dory_ghost_slots(recvImage.data(), recvImage.size(), recvId, sendImage.data(), sendImage.size(), dir, nullptr, &n) != DORY_OK)
        die(nullptr, "dory_ghost_slots");
    std::vector<uint32_t> s(n);
    if (n && dory_ghost_slots(recvImage.data(), recvImage.size(), recvId, sendImage.data(), sendImage.size(), dir, s.data(), &n) != DORY_OK)
        die(nullptr, "dory_ghost_slots");
    return s;
}

}  // namespace

int main(int argc, char **argv) {
    std::string dir, featuresFile, labelsFile, layerFile, gnn = "GCN", rendezvous, runId, exchange = "p2p";
    unsigned epochs = 10, undirected = 0, pipeline = 0, numLambdas = 1, dryRun = 0, applyFirst = 0;
    unsigned numNodes = 1, nodeId = 0;
    int device = -1;
    float lr = 0.01f, targetAcc = 1.1f, switchThreshold = 0.02f;
    for (int i = 1; i + 1 < argc; i += 2) {
        const std::string k = argv[i], v = argv[i + 1];
        if (k == "--datasetdir") dir = v;
        else if (k == "--featuresfile") featuresFile = v;
        else if (k == "--labelsfile") labelsFile = v;
        else if (k == "--layerfile") layerFile = v;
        else if (k == "--numepochs") epochs = (unsigned)std::atoi(v.c_str());
        else if (k == "--lr") lr = (float)std::atof(v.c_str());
        else if (k == "--gnn") gnn = v;
        else if (k == "--undirected") undirected = (unsigned)std::atoi(v.c_str());
        else if (k == "--pipeline") pipeline = (unsigned)std::atoi(v.c_str());
        else if (k == "--dry-run") dryRun = (unsigned)std::atoi(v.c_str());
        else if (k == "--apply-first") applyFirst = (unsigned)std::atoi(v.c_str());
        else if (k == "--numnodes") numNodes = (unsigned)std::max(1, std::atoi(v.c_str()));
        else if (k == "--nodeid") nodeId = (unsigned)std::atoi(v.c_str());
        else if (k == "--device") device = std::atoi(v.c_str());
        else if (k == "--rendezvous") rendezvous = v;
        else if (k == "--run-id") runId = v;
        else if (k == "--exchange") exchange = v;
        else if (k == "--rendezvous-timeout") g_rendezvous_timeout_s = std::atof(v.c_str());
        else if (k == "--numlambdas") numLambdas = (unsigned)std::max(1, std::atoi(v.c_str()));
        else if (k == "--targetacc") targetAcc = (float)std::atof(v.c_str());
        else if (k == "--switchthreshold") switchThreshold = (float)std::atof(v.c_str());
        else {
            std::fprintf(stderr, "unknown flag %s\n", k.c_str());
            return EXIT_FAILURE;
        }
    }
    if (dir.empty() || featuresFile.empty() || labelsFile.empty() || layerFile.empty()) {
        std::fprintf(stderr, "usage: %s --datasetdir D/ --featuresfile F --labelsfile L --layerfile C "
                             "[--numepochs N] [--lr 0.01] [--gnn GCN|GAT] [--undirected 0]\n", argv[0]);
        return EXIT_FAILURE;
    }
    if (dir.back() != '/') dir += '/';
    if (nodeId >= numNodes || (exchange != "p2p" && exchange != "nccl")) {
        std::fprintf(stderr, "need --nodeid < --numnodes and --exchange p2p|nccl\n");
        return EXIT_FAILURE;
    }
    if (numNodes > 1 && pipeline) {
        std::fprintf(stderr, "--pipeline 1 drives one partition (the shell's barriers are per process)\n");
        return EXIT_FAILURE;
    }
    // The ranks of a run meet through files.  Files of an EARLIER run in the same directory (its NCCL id, its
    // CUDA IPC handles, its image.<q>.ready markers) would be accepted as this run's, so several partitions
    // need either a directory of their own (--rendezvous, what host/run_onnode.sh passes: mktemp -d) or a
    // --run-id that every rank of the run shares and that becomes part of every file name.
    if (numNodes > 1 && rendezvous.empty() && runId.empty() && !dryRun) {
        std::fprintf(stderr, "--numnodes %u needs --rendezvous <fresh directory> or --run-id <unique per run>: the default "
                             "directory may hold the files of a previous run\n", numNodes);
        return EXIT_FAILURE;
    }
    if (rendezvous.empty()) rendezvous = dir.substr(0, dir.size() - 1) + ".rendezvous";
    if (rendezvous.back() != '/') rendezvous += '/';
    const std::string rdvTag = runId.empty() ? "" : "." + runId;

    // readLayerConfigFile, engine/utils.cpp:460-479
    dory_config cfg{};
    cfg.abi_version = DORY_ABI_VERSION;
    cfg.gnn_type = gnn == "GAT" ? DORY_GAT : DORY_GCN;
    {
        std::ifstream f(layerFile);
        std::string line;
        unsigned n = 0;
        while (std::getline(f, line)) {
            if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
            if (n > DORY_MAX_LAYERS) break;
            cfg.dims[n++] = (uint32_t)std::stoul(line);
        }
        if (n < 2) {
            std::fprintf(stderr, "layer file needs at least two widths\n");
            return EXIT_FAILURE;
        }
        cfg.n_layers = n - 1;
    }
    cfg.node_id = nodeId;
    cfg.num_nodes = numNodes;
    cfg.device = device >= 0 ? device : (int)nodeId;
    cfg.learning_rate = lr;
    if (applyFirst) {
        if (pipeline || cfg.gnn_type != DORY_GCN) {
            std::fprintf(stderr, "--apply-first 1 is a GCN schedule driven by dory_epoch (not with --pipeline 1 / GAT)\n");
            return EXIT_FAILURE;
        }
        cfg.flags |= DORY_FLAG_APPLY_FIRST;
    }

    // Engine::init order (engine/engine.cpp:62-100): partition image (preprocess when absent), then
    // features and labels.  Everything up to dory_create is host-only.
    const std::string imageName = "graph." + std::to_string(nodeId) + ".bin";
    std::vector<char> image;
    if (!read_file(dir + imageName, image)) {
        std::fprintf(stderr, "[ Node %3u ]  Preprocessing... Output to %s%s\n", nodeId, dir.c_str(), imageName.c_str());
        if (dory_preprocess_dir(dir.c_str(), nodeId, numNodes, (int)undirected) != DORY_OK) die(nullptr, "dory_preprocess_dir");
        if (!read_file(dir + imageName, image)) die(nullptr, "cannot read the partition image");
    }
    if (image.size() < 16) die(nullptr, "partition image is too short");
    uint32_t hdr[4];  // localVtxCnt, globalVtxCnt, srcGhostCnt, dstGhostCnt (graph/graph.cpp:204-207)
    std::memcpy(hdr, image.data(), sizeof hdr);
    const uint64_t V = hdr[0], Gs = hdr[2];
    const uint32_t F0 = cfg.dims[0], C = cfg.dims[cfg.n_layers];

    // readFeaturesFile (with its feats<F0>.<id>.bin cache) / readLabelsFile, engine/utils.cpp:486-596
    std::vector<float> feats(V * F0), ghostFeats(Gs * F0), onehot(V * C);
    if (dory_read_features(dir.c_str(), featuresFile.c_str(), image.data(), image.size(), nodeId, F0, feats.data(),
                           Gs ? ghostFeats.data() : nullptr) != DORY_OK)
        die(nullptr, "dory_read_features");
    if (dory_read_labels(labelsFile.c_str(), image.data(), image.size(), C, onehot.data()) != DORY_OK)
        die(nullptr, "dory_read_labels");
    // N > 1: the receive plan (which ghost slot every incoming row lands in) and its mirror (where MY
    // rows land on each peer) from the partition images alone.  A peer's image is read once that peer
    // says it is complete; a dry run has no peers and preprocesses what is missing itself.
    std::vector<std::vector<uint32_t>> recvSlots[2], sendSlots[2];
    if (numNodes > 1) {
        if (!dryRun) {
            ::mkdir(rendezvous.c_str(), 0777);
            publish(rendezvous + "image." + std::to_string(nodeId) + ".ready" + rdvTag, "1", 1);
        }
        for (int d = 0; d < 2; ++d) {
            recvSlots[d].resize(numNodes);
            sendSlots[d].resize(numNodes);
        }
        for (unsigned q = 0; q < numNodes; ++q) {
            if (q == nodeId) continue;
            const std::string peerName = dir + "graph." + std::to_string(q) + ".bin";
            std::vector<char> peerImage, marker;
            if (!dryRun) await(rendezvous + "image." + std::to_string(q) + ".ready" + rdvTag, marker, 1);
            if (!read_file(peerName, peerImage)) {
                if (!dryRun) die(nullptr, "peer partition image missing");
                if (dory_preprocess_dir(dir.c_str(), q, numNodes, (int)undirected) != DORY_OK) die(nullptr, "dory_preprocess_dir(peer)");
                if (!read_file(peerName, peerImage)) die(nullptr, "cannot read the peer's partition image");
            }
            for (uint32_t d = 0; d < 2; ++d) {
                recvSlots[d][q] = plan_slots(image, nodeId, peerImage, d);
                sendSlots[d][q] = plan_slots(peerImage, q, image, d);
            }
        }
    }
    if (dryRun) {  // host-side half only: what was read, without touching a GPU
        double fs = 0, ls = 0;
        for (float x : feats) fs += x;
        for (size_t i = 0; i < onehot.size(); ++i) ls += onehot[i] * (double)(i % C);
        std::printf("dry run: V %llu ghosts %llu F0 %u classes %u feature_sum %.6f label_sum %.1f\n",
                    (unsigned long long)V, (unsigned long long)Gs, F0, C, fs, ls);
        for (unsigned q = 0; q < numNodes && numNodes > 1; ++q) {
            if (q == nodeId) continue;
            for (int d = 0; d < 2; ++d) {
                unsigned long long rs = 0, ss = 0;
                for (size_t i = 0; i < recvSlots[d][q].size(); ++i) rs += (unsigned long long)(recvSlots[d][q][i] + 1) * (i + 1);
                for (size_t i = 0; i < sendSlots[d][q].size(); ++i) ss += (unsigned long long)(sendSlots[d][q][i] + 1) * (i + 1);
                std::printf("plan: dir %d peer %u recv %zu %llu send %zu %llu\n", d, q, recvSlots[d][q].size(), rs,
                            sendSlots[d][q].size(), ss);
            }
        }
        return EXIT_SUCCESS;
    }

    dory_engine *e = nullptr;
    if (dory_create(&e, &cfg) != DORY_OK) die(nullptr, "dory_create");
    if (dory_load_partition(e, image.data(), image.size()) != DORY_OK) die(e, "dory_load_partition");
    const char *in_name = cfg.gnn_type == DORY_GCN ? "x" : "h";
    if (dory_set_tensor(e, 0, in_name, feats.data(), V, F0) != DORY_OK) die(e, "dory_set_tensor(features)");
    if (dory_set_tensor(e, cfg.n_layers - 1, "lab", onehot.data(), V, C) != DORY_OK) die(e, "dory_set_tensor(labels)");
    if (dory_init_weights(e) != DORY_OK) die(e, "dory_init_weights");

    if (numNodes > 1) {
        // layer-0 ghost features come from the feature file, like every graph server reads its own
        // (readFeaturesFile, engine/utils.cpp:486-552); GAT's first exchange ships z, so it has none
        if (cfg.gnn_type == DORY_GCN && Gs && dory_set_tensor(e, 0, "fg", ghostFeats.data(), Gs, F0) != DORY_OK)
            die(e, "dory_set_tensor(ghost features)");
        // communicator: rank 0's NCCL id travels through the rendezvous directory
        std::vector<char> id(DORY_UNIQUE_ID_BYTES);
        if (nodeId == 0) {
            if (dory_comm_unique_id(id.data()) != DORY_OK) die(nullptr, "dory_comm_unique_id");
            publish(rendezvous + "nccl_id" + rdvTag, id.data(), id.size());
        } else {
            await(rendezvous + "nccl_id" + rdvTag, id, DORY_UNIQUE_ID_BYTES);
        }
        if (dory_comm_init(e, id.data()) != DORY_OK) die(e, "dory_comm_init");
        for (uint32_t d = 0; d < 2; ++d)
            for (unsigned q = 0; q < numNodes; ++q) {
                if (q == nodeId) continue;
                if (dory_comm_set_recv_slots(e, d, q, recvSlots[d][q].data(), (uint32_t)recvSlots[d][q].size()) != DORY_OK)
                    die(e, "dory_comm_set_recv_slots");
                if (exchange == "p2p" &&
                    dory_comm_set_send_slots(e, d, q, sendSlots[d][q].data(), (uint32_t)sendSlots[d][q].size()) != DORY_OK)
                    die(e, "dory_comm_set_send_slots");
            }
        if (exchange == "p2p") {  // map every peer's ghost blocks (CUDA IPC handles through the rendezvous directory)
            const auto ghosts = ghost_tensors(e, cfg);
            std::vector<char> blobs(ghosts.size() * DORY_IPC_BLOB_BYTES);
            for (size_t i = 0; i < ghosts.size(); ++i)
                if (dory_comm_ipc_export(e, ghosts[i].first, ghosts[i].second.c_str(), blobs.data() + i * DORY_IPC_BLOB_BYTES) != DORY_OK)
                    die(e, "dory_comm_ipc_export");
            publish(rendezvous + "ipc." + std::to_string(nodeId) + rdvTag, blobs.data(), blobs.size());
            for (unsigned q = 0; q < numNodes; ++q) {
                if (q == nodeId) continue;
                std::vector<char> theirs;
                await(rendezvous + "ipc." + std::to_string(q) + rdvTag, theirs, blobs.size());
                for (size_t i = 0; i < ghosts.size(); ++i)
                    if (dory_comm_ipc_import(e, ghosts[i].first, ghosts[i].second.c_str(), q, theirs.data() + i * DORY_IPC_BLOB_BYTES) != DORY_OK)
                        die(e, "dory_comm_ipc_import");
            }
        }
    }

    if (pipeline) {
        saga::Config pc;
        pc.gnn = cfg.gnn_type;
        pc.numLayers = cfg.n_layers;
        pc.numChunks = numLambdas;
        pc.numEpochs = epochs;
        pc.localVtxCnt = (uint32_t)V;
        pc.nodeId = nodeId;
        pc.targetAcc = targetAcc;
        pc.switchThreshold = switchThreshold;
        saga::Pipeline pipe(pc, saga::engine_ops(e));
        if (pipe.run() != DORY_OK) die(e, "pipeline");
        pipe.report();
        std::printf("Final: epochs %u, state %s, acc: %.3f, loss: %.3f\n", pipe.epochs_run(),
                    saga::converge_name(pipe.converge_state()), pipe.last_acc(), pipe.last_loss());
        dory_destroy(e);
        return EXIT_SUCCESS;
    }

    double total_ms = 0;
    for (unsigned ep = 1; ep <= epochs; ++ep) {
        const auto t0 = std::chrono::steady_clock::now();
        dory_stats st{};
        if (dory_epoch(e, &st) != DORY_OK) die(e, "dory_epoch");
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        if (ep > 1) total_ms += ms;  // the reference skips the first epoch (pipeline.cpp:117)
        const double denom = st.val_rows ? st.val_rows : 1;
        if (numNodes > 1) std::printf("[ Node %3u ]  ", nodeId);
        std::printf("Epoch %u, acc: %.3f, loss: %.3f, time: %.3f ms\n", ep, st.acc_sum / denom, st.loss_sum / denom, ms);
    }
    if (numNodes > 1) std::printf("[ Node %3u ]  ", nodeId);
    if (epochs > 1) std::printf("<EM>: Average epoch time %.3f ms\n", total_ms / (epochs - 1));
    dory_destroy(e);
    return EXIT_SUCCESS;
}
