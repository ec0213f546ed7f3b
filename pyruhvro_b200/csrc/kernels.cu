// sm_100a kernels of the Avro -> Arrow direct decode.
//
//   count_kernel  one CTA per 256-record tile; the tile's contiguous byte window is staged into
//                 shared memory with coalesced 128-bit loads; each lane walks its record (COUNT)
//                 and the CTA reduces per-stream totals -> tile_agg.  Validates every record.
//   scan_kernel   per (stream, chunk): exclusive scan of the tile totals (each chunk is its own
//                 chain so Arrow offsets restart at 0 in every output batch) + chunk totals.
//   emit_kernel   same staging; COUNT walk, CTA-wide exclusive scan of the lane counts + the
//                 tile's base, then the EMIT walk writes values / offsets / validity / bytes
//                 straight into the Arrow buffers.
//   null_count_kernel  popcount of validity bitmaps (decides lazy validity export).
//
// Integer/byte work bounded by HBM bandwidth; no tensor-core use.
#include "kernels.cuh"

#include <algorithm>

#include "dev_kernels.cuh"
#include "interp.cuh"

namespace rv {
namespace {

// Generic (interpreter) kernels.  Normal mode: one CTA per tile.  Overflow mode (p.tile_list set): a
// fixed grid strides over the tiles the schema-specialised kernels skipped; the list length lives on
// the device (p.overflow[0]) so no host round trip is needed to size the launch.
__global__ void __launch_bounds__(kBlock) count_kernel(const DecodeParams p) {
    if (!p.tile_list) { count_body<InterpWalker, true>(p, int(blockIdx.x)); return; }
    const int n = p.overflow[0];
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        count_body<InterpWalker, true>(p, p.tile_list[i]);
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kBlock) emit_kernel(const DecodeParams p) {
    if (!p.tile_list) { emit_body<InterpWalker, true>(p, int(blockIdx.x)); return; }
    const int n = p.overflow[0];
    for (int i = blockIdx.x; i < n; i += gridDim.x) {
        emit_body<InterpWalker, true>(p, p.tile_list[i]);
        __syncthreads();
    }
}

// grid = k * n_streams CTAs.  Each thread owns a contiguous run of the chunk's tiles.
__global__ void scan_kernel(const DecodeParams p) {
    __shared__ unsigned long long s_part[32];
    const int s = int(blockIdx.x % unsigned(p.n_streams)), j = int(blockIdx.x / unsigned(p.n_streams));
    const int t_begin = j * p.tiles_per_chunk;
    const int t_end = (j == p.k - 1) ? p.n_tiles : t_begin + p.tiles_per_chunk;
    const int T = t_end - t_begin;
    const int nthr = blockDim.x;
    const int per = (T + nthr - 1) / nthr;
    const int a = t_begin + min(T, int(threadIdx.x) * per);
    const int b = t_begin + min(T, (int(threadIdx.x) + 1) * per);
    const uint32_t* agg = p.tile_agg + size_t(s) * p.n_tiles;
    uint32_t* base = p.tile_base + size_t(s) * p.n_tiles;
    unsigned long long local = 0;
    for (int i = a; i < b; ++i) local += agg[i];
    // block exclusive scan of `local`
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned long long incl = local;
    for (int d = 1; d < 32; d <<= 1) {
        unsigned long long v = __shfl_up_sync(0xFFFFFFFFu, incl, d);
        if (lane >= d) incl += v;
    }
    if (lane == 31) s_part[warp] = incl;
    __syncthreads();
    if (warp == 0) {
        unsigned long long v = (lane < (nthr + 31) / 32) ? s_part[lane] : 0ull;
        unsigned long long w = v;
        for (int d = 1; d < 32; d <<= 1) {
            unsigned long long u = __shfl_up_sync(0xFFFFFFFFu, w, d);
            if (lane >= d) w += u;
        }
        s_part[lane] = w - v;  // exclusive prefix of warp totals
        if (lane == 31) {
            const unsigned long long total = w;
            p.chunk_tot[size_t(j) * p.n_streams + s] = total;
            if (total > 0x7FFFFFFFull) report(p, int64_t(j) * p.chunk_rows, E_OVERFLOW);
        }
    }
    __syncthreads();
    unsigned long long run = s_part[warp] + (incl - local);
    for (int i = a; i < b; ++i) {
        base[i] = uint32_t(run);
        run += agg[i];
    }
}

__global__ void null_count_kernel(const NullCountJob* jobs, long long* ones_out) {
    const NullCountJob job = jobs[blockIdx.x];
    const int64_t n_words = (job.n_bits + 31) >> 5;
    long long ones = 0;
    for (int64_t w = int64_t(blockIdx.y) * blockDim.x + threadIdx.x; w < n_words; w += int64_t(gridDim.y) * blockDim.x) {
        uint32_t v = __ldg(job.bitmap + w);
        const int64_t rem = job.n_bits - (w << 5);
        if (rem < 32) v &= (1u << rem) - 1u;
        ones += __popc(v);
    }
    for (int d = 16; d; d >>= 1) ones += __shfl_xor_sync(0xFFFFFFFFu, ones, d);
    __shared__ long long s_w[32];
    if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = ones;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long tot = 0;
        for (int i = 0; i < int(blockDim.x >> 5); ++i) tot += s_w[i];
        if (tot) atomicAdd(reinterpret_cast<unsigned long long*>(ones_out + blockIdx.x), static_cast<unsigned long long>(tot));
    }
}

// ---- window sizing: the largest tile decides how much shared memory a CTA needs ---------------------
// max over tiles of the tile's input byte span (one thread per tile).
// Also notes the first and last input offset (the input's byte span) in the control block.
__global__ void tile_span_max_kernel(const DecodeParams p, unsigned long long* ctrl) {
    unsigned long long* out_max = ctrl + 1;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        ctrl[4] = static_cast<unsigned long long>(p.offsets[0]);
        ctrl[5] = static_cast<unsigned long long>(p.offsets[p.n]);
    }
    unsigned long long m = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.n_tiles; t += gridDim.x * blockDim.x) {
        const Tile tl = tile_of(p, t);
        const long long span = p.offsets[tl.r0 + tl.nrec] - p.offsets[tl.r0];
        if (span > 0 && (unsigned long long)span > m) m = (unsigned long long)span;
    }
    for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, m, d); if (o > m) m = o; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out_max, m);
}

// max over tiles of the Utf8 bytes the tile stages for write-out (sum over byte streams, + alignment slack).
__global__ void tile_utf8_max_kernel(const DecodeParams p, unsigned long long* out_max) {
    unsigned long long m = 0;
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < p.n_tiles; t += gridDim.x * blockDim.x) {
        unsigned long long sum = 0;
        for (int s = 0; s < p.n_streams; ++s)
            if (p.stream_slot[s] >= 0) sum += p.tile_agg[size_t(s) * p.n_tiles + t] + 31ull;
        if (sum > m) m = sum;
    }
    for (int d = 16; d; d >>= 1) { const unsigned long long o = __shfl_xor_sync(0xFFFFFFFFu, m, d); if (o > m) m = o; }
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out_max, m);
}

// ---- fix-ups for gathering shard-local Arrow buffers into one batch (multi-GPU) ----
// dst[i] = src[i] + add: rebases a shard's i32 offsets by the total of the shards before it.
__global__ void rebase_i32_kernel(int32_t* dst, const int32_t* src, long long n, int32_t add) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) dst[i] = src[i] + add;
}

// ORs `nbits` bits of `src` (bit 0 first) into `dst` starting at bit `dst_bit`; dst must be zero where
// nothing was written yet.  One thread per destination word; seam words are shared between shards, hence atomicOr.
__global__ void concat_bits_kernel(uint32_t* dst, long long dst_bit, const uint32_t* src, long long nbits) {
    const long long w0 = dst_bit >> 5, w1 = (dst_bit + nbits + 31) >> 5;
    const unsigned sh = unsigned(dst_bit & 31);
    const long long n_src_words = (nbits + 31) >> 5;
    for (long long w = w0 + blockIdx.x * (long long)blockDim.x + threadIdx.x; w < w1; w += (long long)gridDim.x * blockDim.x) {
        const long long k = w - w0;  // source word whose low bits land at bit `sh` of this word
        uint32_t lo = 0, hi = 0;
        if (k < n_src_words) {
            hi = src[k];
            const long long rem = nbits - (k << 5);
            if (rem < 32) hi &= (1u << rem) - 1u;
        }
        if (k >= 1 && sh) {
            lo = src[k - 1];
            const long long rem = nbits - ((k - 1) << 5);
            if (rem < 32) lo &= (1u << rem) - 1u;
        }
        const uint32_t v = sh ? ((hi << sh) | (lo >> (32u - sh))) : hi;
        if (v) atomicOr(dst + w, v);
    }
}

}  // namespace

void launch_tile_span_max(const DecodeParams& p, unsigned long long* ctrl, cudaStream_t s) {
    const int blocks = std::max(1, std::min((p.n_tiles + 255) / 256, 148 * 4));
    tile_span_max_kernel<<<blocks, 256, 0, s>>>(p, ctrl);
}

void launch_tile_utf8_max(const DecodeParams& p, unsigned long long* out_max, cudaStream_t s) {
    const int blocks = std::max(1, std::min((p.n_tiles + 255) / 256, 148 * 4));
    tile_utf8_max_kernel<<<blocks, 256, 0, s>>>(p, out_max);
}

void launch_rebase_i32(int32_t* dst, const int32_t* src, long long n, int32_t add, cudaStream_t s) {
    if (n <= 0) return;
    const int blocks = int(std::min<long long>((n + 255) / 256, 148 * 8));
    rebase_i32_kernel<<<blocks, 256, 0, s>>>(dst, src, n, add);
}

void launch_concat_bits(uint32_t* dst, long long dst_bit, const uint32_t* src, long long nbits, cudaStream_t s) {
    if (nbits <= 0) return;
    const long long words = ((dst_bit + nbits + 31) >> 5) - (dst_bit >> 5);
    const int blocks = int(std::min<long long>((words + 255) / 256, 148 * 8));
    concat_bits_kernel<<<blocks, 256, 0, s>>>(dst, dst_bit, src, nbits);
}

cudaError_t prepare_kernels() {
    cudaError_t e = cudaFuncSetAttribute(count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(count_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(emit_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    return e;
}

void launch_count(const DecodeParams& p, int n_ctas, size_t smem, cudaStream_t s) {
    count_kernel<<<n_ctas, kBlock, smem, s>>>(p);
}

void launch_scan(const DecodeParams& p, cudaStream_t s) {
    const int t_max = p.n_tiles - (p.k - 1) * p.tiles_per_chunk > p.tiles_per_chunk ? p.n_tiles - (p.k - 1) * p.tiles_per_chunk : p.tiles_per_chunk;
    int threads = 32;
    while (threads < 1024 && threads < t_max) threads <<= 1;
    if (p.n_streams == 0) return;
    scan_kernel<<<unsigned(p.n_streams) * unsigned(p.k), threads, 0, s>>>(p);
}

void launch_emit(const DecodeParams& p, int n_ctas, size_t smem, cudaStream_t s) {
    emit_kernel<<<n_ctas, kBlock, smem, s>>>(p);
}

void launch_null_count(const NullCountJob* jobs, int n_jobs, long long* out, cudaStream_t s) {
    if (n_jobs <= 0) return;
    null_count_kernel<<<dim3(n_jobs, 16), 256, 0, s>>>(jobs, out);
}

}  // namespace rv
