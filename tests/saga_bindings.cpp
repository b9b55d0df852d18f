// Test-only C bindings of host/saga_pipeline.hpp pieces that are compared with the reference's own
// code (tests/test_oracle.py: Chunk::operator<).
#include "../host/saga_pipeline.hpp"

extern "C" int saga_chunk_less(const unsigned *a, const unsigned *b) {
    auto mk = [](const unsigned *f) { return dory_chunk{f[0], f[1], f[2], f[3], f[4], f[5], f[6], (uint8_t)(f[7] != 0)}; };
    return saga::ChunkLess()(mk(a), mk(b)) ? 1 : 0;
}
