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
//
// --pipeline 1 drives the epochs through host/saga_pipeline.hpp -- the reference's chunk queues,
// priority order, barriers, early-stop state machine and <EM> report (engine/ops/pipeline.cpp) --
// instead of dory_epoch; --numlambdas is the number of chunks per partition (numLambdasForward).
// --apply-first 1 sets DORY_FLAG_APPLY_FIRST (GCN layers that narrow run A_hat.(in.W)); that schedule is
// driven by dory_epoch, not by the reference's queue order, so it excludes --pipeline 1.
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
#include <vector>

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

}  // namespace

int main(int argc, char **argv) {
    std::string dir, featuresFile, labelsFile, layerFile, gnn = "GCN";
    unsigned epochs = 10, undirected = 0, pipeline = 0, numLambdas = 1, dryRun = 0, applyFirst = 0;
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
    cfg.node_id = 0;
    cfg.num_nodes = 1;
    cfg.device = 0;
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
    std::vector<char> image;
    if (!read_file(dir + "graph.0.bin", image)) {
        std::fprintf(stderr, "[ Node   0 ]  Preprocessing... Output to %sgraph.0.bin\n", dir.c_str());
        if (dory_preprocess_dir(dir.c_str(), 0, 1, (int)undirected) != DORY_OK) die(nullptr, "dory_preprocess_dir");
        if (!read_file(dir + "graph.0.bin", image)) die(nullptr, "cannot read graph.0.bin");
    }
    if (image.size() < 16) die(nullptr, "graph.0.bin is too short");
    uint32_t hdr[4];  // localVtxCnt, globalVtxCnt, srcGhostCnt, dstGhostCnt (graph/graph.cpp:204-207)
    std::memcpy(hdr, image.data(), sizeof hdr);
    const uint64_t V = hdr[0], Gs = hdr[2];
    const uint32_t F0 = cfg.dims[0], C = cfg.dims[cfg.n_layers];

    // readFeaturesFile (with its feats<F0>.<id>.bin cache) / readLabelsFile, engine/utils.cpp:486-596
    std::vector<float> feats(V * F0), ghostFeats(Gs * F0), onehot(V * C);
    if (dory_read_features(dir.c_str(), featuresFile.c_str(), image.data(), image.size(), 0, F0, feats.data(),
                           Gs ? ghostFeats.data() : nullptr) != DORY_OK)
        die(nullptr, "dory_read_features");
    if (dory_read_labels(labelsFile.c_str(), image.data(), image.size(), C, onehot.data()) != DORY_OK)
        die(nullptr, "dory_read_labels");
    if (dryRun) {  // host-side half only: what was read, without touching a GPU
        double fs = 0, ls = 0;
        for (float x : feats) fs += x;
        for (size_t i = 0; i < onehot.size(); ++i) ls += onehot[i] * (double)(i % C);
        std::printf("dry run: V %llu ghosts %llu F0 %u classes %u feature_sum %.6f label_sum %.1f\n",
                    (unsigned long long)V, (unsigned long long)Gs, F0, C, fs, ls);
        return EXIT_SUCCESS;
    }

    dory_engine *e = nullptr;
    if (dory_create(&e, &cfg) != DORY_OK) die(nullptr, "dory_create");
    if (dory_load_partition(e, image.data(), image.size()) != DORY_OK) die(e, "dory_load_partition");
    const char *in_name = cfg.gnn_type == DORY_GCN ? "x" : "h";
    if (dory_set_tensor(e, 0, in_name, feats.data(), V, F0) != DORY_OK) die(e, "dory_set_tensor(features)");
    if (dory_set_tensor(e, cfg.n_layers - 1, "lab", onehot.data(), V, C) != DORY_OK) die(e, "dory_set_tensor(labels)");
    if (dory_init_weights(e) != DORY_OK) die(e, "dory_init_weights");

    if (pipeline) {
        saga::Config pc;
        pc.gnn = cfg.gnn_type;
        pc.numLayers = cfg.n_layers;
        pc.numChunks = numLambdas;
        pc.numEpochs = epochs;
        pc.localVtxCnt = (uint32_t)V;
        pc.nodeId = 0;
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
        std::printf("Epoch %u, acc: %.3f, loss: %.3f, time: %.3f ms\n", ep, st.acc_sum / denom, st.loss_sum / denom, ms);
    }
    if (epochs > 1) std::printf("<EM>: Average epoch time %.3f ms\n", total_ms / (epochs - 1));
    dory_destroy(e);
    return EXIT_SUCCESS;
}
