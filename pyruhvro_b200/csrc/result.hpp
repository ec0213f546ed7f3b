// Output-side bookkeeping shared by the CUDA engine and the host emulation used in tests:
// the exact-size arena layout of every Arrow buffer of every output batch, and the Arrow C
// Data Interface export of a decoded batch (what FieldDecoder::finish + RecordBatch::try_new
// produce in the reference, ruhvro/src/fast_decode.rs:536-567,829-834).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>

#include "arrow_c.h"
#include "plan.hpp"

namespace rv {

struct ChunkOut {
    int64_t rows = 0;
    std::vector<int64_t> space_rows;   // rows per row space
    std::vector<size_t> slot_off;      // arena offset per slot
    std::vector<int64_t> slot_bytes;   // exact logical bytes per slot
    std::vector<int64_t> null_count;   // per slot (meaningful for validity slots)
};

struct Layout {
    std::vector<ChunkOut> chunks;
    size_t zero_bytes = 0;    // leading region that must be zero-initialised (atomicOr targets)
    size_t total_bytes = 0;
};

// clamp_chunks (ruhvro/src/deserialize.rs:53-55)
int64_t clamp_chunks(int64_t num_chunks, int64_t n);

// chunk_tot is [k][max(S,1)] stream totals per chunk (all zero when n == 0).
Layout compute_layout(const Plan& plan, int64_t n, int k, const unsigned long long* chunk_tot);

// Exports one batch as a struct array whose children are the top-level columns.
// `base` is the arena (host or device); `keep` is retained until the array is released.
void export_batch(const Plan& plan, const ChunkOut& c, const uint8_t* base, std::shared_ptr<void> keep, ArrowArray* out);

// Sum of the exact byte lengths of every exported buffer (B_out).
int64_t exported_bytes(const Plan& plan, const std::vector<ChunkOut>& chunks);

}  // namespace rv
