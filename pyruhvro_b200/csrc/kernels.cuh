// Launch wrappers of the statically compiled (interpreter) kernels (sm_100a).  See kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "dev_types.h"

namespace rv {

struct NullCountJob {
    const uint32_t* bitmap;
    int64_t n_bits;
};

cudaError_t prepare_kernels();  // opt in to large dynamic shared memory (once per device)

void launch_count(const DecodeParams& p, int n_ctas, size_t smem, cudaStream_t s);   // interpreter walker
void launch_emit(const DecodeParams& p, int n_ctas, size_t smem, cudaStream_t s);    // interpreter walker
void launch_scan(const DecodeParams& p, cudaStream_t s);
void launch_tile_span_max(const DecodeParams& p, unsigned long long* ctrl, cudaStream_t s);   // ctrl[1] = max tile input bytes, ctrl[4..5] = first/last offset
void launch_tile_utf8_max(const DecodeParams& p, unsigned long long* out_max, cudaStream_t s);   // [0] = max tile Utf8 staging bytes
void launch_rebase_i32(int32_t* dst, const int32_t* src, long long n, int32_t add, cudaStream_t s);
void launch_concat_bits(uint32_t* dst, long long dst_bit, const uint32_t* src, long long nbits, cudaStream_t s);
void launch_null_count(const NullCountJob* jobs, int n_jobs, long long* out, cudaStream_t s);

}  // namespace rv
