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
#include "walker.cuh"

namespace rv {
namespace {

constexpr int kWarps = kBlock / 32;

struct Tile {
    int chunk;
    int local_tile;
    int64_t r0;          // first record of the tile
    int nrec;            // records in the tile
    int64_t chunk_len;   // rows in the chunk
};

__device__ __forceinline__ Tile tile_of(const DecodeParams& p, int tile) {
    Tile t;
    int j = tile / p.tiles_per_chunk;
    if (j > p.k - 1) j = p.k - 1;
    t.chunk = j;
    t.local_tile = tile - j * p.tiles_per_chunk;
    const int64_t cs = int64_t(j) * p.chunk_rows;
    const int64_t ce = (j == p.k - 1) ? p.n : cs + p.chunk_rows;
    t.chunk_len = ce - cs;
    t.r0 = cs + int64_t(t.local_tile) * kBlock;
    const int64_t left = ce - t.r0;
    t.nrec = left < kBlock ? int(left) : kBlock;
    return t;
}

struct Smem {
    DNode* nodes;
    uint32_t* cur;    // [n_streams][kBlock]
    uint32_t* wtot;   // [n_streams][kWarps]
    uint8_t* data;    // staged tile bytes
};

__device__ __forceinline__ Smem carve(uint8_t* base, const DecodeParams& p) {
    Smem s;
    s.nodes = reinterpret_cast<DNode*>(base);
    s.cur = reinterpret_cast<uint32_t*>(base + size_t(p.n_nodes) * sizeof(DNode));
    s.wtot = s.cur + size_t(p.n_streams) * kBlock;
    s.data = reinterpret_cast<uint8_t*>(s.wtot + size_t(p.n_streams) * kWarps);
    return s;
}

// Loads the plan and the tile's byte window into shared memory and builds the lane's context.
// Returns the record index of this lane.
__device__ __forceinline__ int64_t setup(const DecodeParams& p, const Tile& t, const Smem& sm, WalkCtx& c) {
    const int tid = threadIdx.x;
    for (int i = tid; i < p.n_nodes * int(sizeof(DNode) / 16); i += kBlock)
        reinterpret_cast<uint4*>(sm.nodes)[i] = __ldg(reinterpret_cast<const uint4*>(p.nodes) + i);
    for (int s = 0; s < p.n_streams; ++s) sm.cur[s * kBlock + tid] = 0;

    const int64_t t0 = __ldg(p.offsets + t.r0);
    const int64_t t1 = __ldg(p.offsets + t.r0 + t.nrec);
    const int64_t span = t1 - t0;
    const uint32_t mis = uint32_t(reinterpret_cast<uintptr_t>(p.data + t0) & 15u);
    const bool staged = span >= 0 && uint64_t(span) + mis <= uint64_t(p.smem_data_cap);
    if (staged) {
        const uint4* g = reinterpret_cast<const uint4*>(p.data + t0 - mis);
        const int nvec = int((span + mis + 15) >> 4);
        uint4* d = reinterpret_cast<uint4*>(sm.data);
        for (int i = tid; i < nvec; i += kBlock) d[i] = __ldg(g + i);
    }

    c.nodes = sm.nodes;
    c.cur = sm.cur + tid;
    c.cur_stride = kBlock;
    c.sym_off = p.sym_off;
    c.sym_bytes = p.sym_bytes;
    c.bufs = p.bufs ? p.bufs + size_t(t.chunk) * p.n_slots : nullptr;
    c.err = 0;
    c.in_range = tid < t.nrec;
    c.row0 = uint32_t(t.local_tile) * kBlock + tid;
    c.store_word = (tid & 31) == 0 && int64_t(c.row0) < t.chunk_len;
    c.base = sm.data;
    c.pos = c.end = 0;
    const int64_t r = t.r0 + tid;
    if (c.in_range) {
        const int64_t o0 = __ldg(p.offsets + r), o1 = __ldg(p.offsets + r + 1);
        if (o1 < o0 || o1 - o0 > int64_t(0xFFFFFFF0u)) c.err = E_OVERFLOW;  // malformed offsets / >4 GiB record
        else if (staged) {
            if (o0 < t0 || o1 > t1) c.err = E_OVERFLOW;
            else { c.pos = uint32_t(o0 - t0) + mis; c.end = uint32_t(o1 - t0) + mis; }
        } else {
            c.base = p.data + o0;
            c.end = uint32_t(o1 - o0);
        }
    }
    return r;
}

__device__ __forceinline__ void report(const DecodeParams& p, int64_t record, uint32_t code) {
    atomicMin(p.err, (static_cast<unsigned long long>(record) << 8) | code);
}

extern __shared__ __align__(16) uint8_t g_smem[];

__global__ void __launch_bounds__(kBlock) count_kernel(const DecodeParams p) {
    const Tile t = tile_of(p, blockIdx.x);
    const Smem sm = carve(g_smem, p);
    WalkCtx c;
    const int64_t r = setup(p, t, sm, c);
    __syncthreads();
    walk_record<WM_COUNT>(c, p.n_nodes);
    if (c.in_range && c.err) report(p, r, c.err);
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int s = warp; s < p.n_streams; s += kWarps) {
        unsigned long long sum = 0;
        for (int i = lane; i < kBlock; i += 32) sum += sm.cur[s * kBlock + i];
        for (int d = 16; d; d >>= 1) sum += __shfl_xor_sync(0xFFFFFFFFu, sum, d);
        if (lane == 0) {
            if (sum > 0x7FFFFFFFull) { report(p, t.r0, E_OVERFLOW); sum = 0x7FFFFFFFull; }
            p.tile_agg[size_t(s) * p.n_tiles + blockIdx.x] = uint32_t(sum);
        }
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

__global__ void __launch_bounds__(kBlock) emit_kernel(const DecodeParams p) {
    const Tile t = tile_of(p, blockIdx.x);
    const Smem sm = carve(g_smem, p);
    WalkCtx c;
    (void)setup(p, t, sm, c);
    __syncthreads();
    const uint32_t pos0 = c.pos;
    const uint8_t* base0 = c.base;
    walk_record<WM_COUNT>(c, p.n_nodes);
    const uint32_t count_err = c.err;
    __syncthreads();

    // CTA-wide exclusive scan of every stream's lane counts, offset by the tile's base.
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int s = 0; s < p.n_streams; ++s) {
        const uint32_t v = sm.cur[s * kBlock + tid];
        uint32_t incl = v;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += u;
        }
        if (lane == 31) sm.wtot[s * kWarps + warp] = incl;
        sm.cur[s * kBlock + tid] = incl - v;
    }
    __syncthreads();
    for (int s = 0; s < p.n_streams; ++s) {
        uint32_t b = __ldg(p.tile_base + size_t(s) * p.n_tiles + blockIdx.x);
        for (int w = 0; w < warp; ++w) b += sm.wtot[s * kWarps + w];
        sm.cur[s * kBlock + tid] += b;
    }
    // offsets[0] = 0 of every offsets buffer of this chunk (first tile of the chunk only)
    if (t.local_tile == 0) {
        for (int i = tid; i < p.n_nodes; i += kBlock) {
            const DNode nd = sm.nodes[i];
            if (nd.kind == NK_STR || nd.kind == NK_ENUM || nd.kind == NK_LIST || nd.kind == NK_MAP)
                static_cast<int32_t*>(c.bufs[nd.slot_a])[0] = 0;
        }
    }
    c.pos = pos0;
    c.base = base0;
    c.err = count_err;  // a record that failed validation emits only null slots
    walk_record<WM_EMIT>(c, p.n_nodes);
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

}  // namespace

size_t decode_smem_bytes(int n_nodes, int n_streams, uint32_t data_cap) {
    return size_t(n_nodes) * sizeof(DNode) + size_t(n_streams) * kBlock * 4 + size_t(n_streams) * kWarps * 4 + data_cap;
}

cudaError_t prepare_kernels() {
    cudaError_t e = cudaFuncSetAttribute(count_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(emit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

void launch_count(const DecodeParams& p, size_t smem, cudaStream_t s) {
    count_kernel<<<p.n_tiles, kBlock, smem, s>>>(p);
}

void launch_scan(const DecodeParams& p, cudaStream_t s) {
    const int t_max = p.n_tiles - (p.k - 1) * p.tiles_per_chunk > p.tiles_per_chunk ? p.n_tiles - (p.k - 1) * p.tiles_per_chunk : p.tiles_per_chunk;
    int threads = 32;
    while (threads < 1024 && threads < t_max) threads <<= 1;
    if (p.n_streams == 0) return;
    scan_kernel<<<unsigned(p.n_streams) * unsigned(p.k), threads, 0, s>>>(p);
}

void launch_emit(const DecodeParams& p, size_t smem, cudaStream_t s) {
    emit_kernel<<<p.n_tiles, kBlock, smem, s>>>(p);
}

void launch_null_count(const NullCountJob* jobs, int n_jobs, long long* out, cudaStream_t s) {
    if (n_jobs <= 0) return;
    null_count_kernel<<<dim3(n_jobs, 16), 256, 0, s>>>(jobs, out);
}

}  // namespace rv
