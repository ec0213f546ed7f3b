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

// One push of the multi-GPU gather (gather.hpp): `dst` may be peer memory (NVLink).
struct PushJob {
    const uint8_t* src;   // this rank's buffer (64-byte aligned)
    uint8_t* dst;         // where it lands in the gathered arena (RAW: any alignment; OFFSETS: element `row_base`; BITS: the bitmap's first word)
    int64_t count;        // RAW: bytes; OFFSETS: rows; BITS: bits
    int64_t param;        // OFFSETS: add; BITS: destination bit
    int32_t kind;         // GatherKind
    int32_t pad;
};

// Object container files: where every record of every block starts (ocf.hpp).
struct OcfBlockDev { long long data_off, size, count, rec_base; };
struct OcfParams {
    const uint8_t* data;          // the file's bytes on the device
    const OcfBlockDev* blocks;
    int32_t n_blocks;
    const DNode* nodes;
    int32_t n_nodes;
    const int32_t* sym_off;
    const uint8_t* sym_bytes;
    int64_t* offsets;             // [n_records + 1], relative to `data`
    int64_t n_records;
    long long end_off;
    unsigned long long* err;      // min over (record << 8 | code); ~0 = none
};

cudaError_t prepare_kernels();  // opt in to large dynamic shared memory (once per device)

void launch_fused(const DecodeParams& p, size_t smem, cudaStream_t s);   // interpreter walker, one CTA per tile
void launch_rebase_i32(int32_t* dst, const int32_t* src, long long n, int32_t add, cudaStream_t s);
void launch_concat_bits(uint32_t* dst, long long dst_bit, const uint32_t* src, long long nbits, cudaStream_t s);
void launch_null_count(const NullCountJob* jobs, int n_jobs, long long* out, cudaStream_t s);
void launch_compact(const CompactJob* jobs, int n_jobs, int parts, cudaStream_t s);
void launch_ocf_offsets(const OcfParams& q, cudaStream_t s);
void launch_gather_push(const PushJob* jobs, int n_jobs, int parts, cudaStream_t s);

}  // namespace rv
