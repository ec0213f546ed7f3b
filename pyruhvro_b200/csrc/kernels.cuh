// Launch wrappers of the statically compiled kernels (sm_100a).  See kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "dev_types.h"

namespace rv {

struct NullCountJob {
    const uint32_t* bitmap;
    int64_t n_bits;                            // used when n_bits_dev is null
    const unsigned long long* n_bits_dev;      // device word holding the bit count (rows of a deeper row space)
};

struct CompactJob {
    const uint8_t* src;
    uint8_t* dst;      // both 16-byte aligned
    int64_t bytes;
};

cudaError_t prepare_kernels();  // opt in to large dynamic shared memory (once per device)

void launch_fused(const DecodeParams& p, size_t smem, cudaStream_t s);   // interpreter walker, one CTA per tile
void launch_rebase_i32(int32_t* dst, const int32_t* src, long long n, int32_t add, cudaStream_t s);
void launch_concat_bits(uint32_t* dst, long long dst_bit, const uint32_t* src, long long nbits, cudaStream_t s);
void launch_null_count(const NullCountJob* jobs, int n_jobs, long long* out, cudaStream_t s);
void launch_compact(const CompactJob* jobs, int n_jobs, int parts, cudaStream_t s);

}  // namespace rv
