// Edge-cut partitioner (internal; the C ABI wraps it as dory_partition_edges / dory_partition_file).
#pragma once
#include <cstdint>
#include <string>

namespace dory {

// parts[v] <- owner of global vertex v, in [0, n_parts).  Returns "" or an error message.
std::string partition_edges(const uint32_t *src, const uint32_t *dst, uint64_t n_edges, uint32_t n_vertices,
                            uint32_t n_parts, uint32_t passes, int32_t *parts, uint64_t *edge_cut);
std::string partition_file(const char *bsnap_path, uint32_t n_parts, const char *out_dir, uint32_t passes);

}  // namespace dory
