// Kernel parameter blocks and launch wrappers (sm_100a).  See kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "plan.hpp"

namespace rv {

constexpr int kBlock = 256;  // records per tile == threads per CTA (one record per lane)

struct DecodeParams {
    // input: packed Avro records (BinaryArray layout, deserialize.rs:90) with i64 offsets
    const uint8_t* data;
    const int64_t* offsets;
    int64_t n;            // records
    int64_t chunk_rows;   // n / k (last chunk takes the remainder, deserialize.rs:57-68)
    int32_t k;            // chunks (= output batches)
    int32_t tiles_per_chunk;
    int32_t n_tiles;
    // plan
    const DNode* nodes;
    int32_t n_nodes;
    int32_t n_streams;
    int32_t n_slots;
    const int32_t* sym_off;
    const uint8_t* sym_bytes;
    // scan scratch
    uint32_t* tile_agg;    // [n_streams][n_tiles] per-tile totals
    uint32_t* tile_base;   // [n_streams][n_tiles] exclusive prefix within the chunk
    unsigned long long* chunk_tot;  // [k][n_streams]
    unsigned long long* err;        // min over (record << 8 | code); ~0 = none
    // output
    void* const* bufs;     // [k][n_slots]
    uint32_t smem_data_cap;  // bytes of shared memory available for staging a tile's bytes
};

struct NullCountJob {
    const uint32_t* bitmap;
    int64_t n_bits;
};

size_t decode_smem_bytes(int n_nodes, int n_streams, uint32_t data_cap);
cudaError_t prepare_kernels();  // opt in to large dynamic shared memory (once per process)

void launch_count(const DecodeParams& p, size_t smem, cudaStream_t s);
void launch_scan(const DecodeParams& p, cudaStream_t s);
void launch_emit(const DecodeParams& p, size_t smem, cudaStream_t s);
void launch_null_count(const NullCountJob* jobs, int n_jobs, long long* out, cudaStream_t s);

}  // namespace rv
