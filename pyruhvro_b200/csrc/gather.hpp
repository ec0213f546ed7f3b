// Gathering shard-local batches into single RecordBatches (multi-GPU; BASELINE.json configs[4]).
//
// Records shard by message: every rank decodes a contiguous range into its own device-resident batch, exactly the
// reference's per-chunk batches (ruhvro/src/deserialize.rs:57-68,115-119).  When ONE batch over all rows is wanted the
// ranks exchange a few counts per rank (GatherMeta), every rank computes the same plan from them, and each rank then
// PUSHES its Arrow buffers straight into the gathered arena on the group's leader GPU (peer memory over NVLink, one
// kernel per rank) with the fix-ups fused into the copy: offsets are rebased by the totals of the ranks before it,
// bitmaps are shifted to their bit position (seam words merged with atomic OR), everything else lands at its prefix
// offset.  Arrow's i32 offsets cap a batch at 2^31-1 rows / bytes per column, so consecutive ranks are grouped into as
// few batches as fit ("gather into ceil(.) batches").
//
// This file is the host-side planning (no CUDA): shared by the engine and by the tests' host emulation.
#pragma once
#include <cstdint>
#include <vector>

#include "plan.hpp"
#include "result.hpp"

namespace rv {

// What one rank's batch contributes, as int64 words: [rows of every row space][total of every stream][null count of
// every validity slot].
int64_t gather_meta_len(const Plan& plan);
void gather_meta_of(const Plan& plan, const ChunkOut& c, int64_t* out);

enum GatherKind : int32_t {
    GK_RAW = 0,      // `count` bytes
    GK_OFFSETS = 1,  // `count` rows: dst[1 + i] = src[1 + i] + param (param = what the earlier ranks hold); the first rank also writes dst[0] = 0
    GK_BITS = 2      // `count` bits, placed at bit `param` of the destination bitmap
};

struct GatherJob {
    int32_t kind;
    int32_t slot;
    int64_t dst_off;  // bytes from the gathered arena's base
    int64_t count;
    int64_t param;
};

struct GatherGroup {
    int first_rank = 0, n_ranks = 0;
    ChunkOut out;                              // the gathered batch: exact layout inside its arena, rows, null counts
    size_t arena_bytes = 0;
    std::vector<std::vector<GatherJob>> jobs;  // per member rank (index = rank - first_rank)
};

struct GatherPlan {
    std::vector<GatherGroup> groups;
    std::vector<int> group_of_rank;
};

// metas: [world][gather_meta_len].  Throws std::runtime_error when a single rank's batch is itself beyond the i32 limits.
GatherPlan plan_gather(const Plan& plan, const int64_t* metas, int world);

}  // namespace rv
