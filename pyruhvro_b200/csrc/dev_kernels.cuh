// Kernel bodies of the decode, templated over the walker (InterpWalker or a generated,
// schema-specialised walker).  Device-only; compiled by nvcc (kernels.cu) and by NVRTC (jit.cpp).
//
//   count_body  one CTA per 256-record tile: one TMA bulk copy (cp.async.bulk + mbarrier) stages the tile's
//               contiguous byte window in shared memory, every lane COUNT-walks its record (full validation),
//               a warp-per-stream scan turns the lane counts into in-tile prefixes -> lane_off, tile_agg.
//   emit_body   same staging; the lanes' cursors (tile base from scan_kernel + in-tile prefix from count) are
//               loaded while the window is in flight -> EMIT walk.  Utf8 bytes are assembled per column in a
//               shared-memory staging area and leave through TMA bulk stores (16-byte aligned body) plus a few
//               head/tail bytes; fixed-width values / offsets are stored row-aligned; space-0 validity is one
//               ballot word per warp.
//
// Shared-memory map (dynamic, rv_smem; smem_map() in dev_types.h):
//   [nodes n_nodes*32][wtot S*8*4 (emit: tile bases)][tot (S+1)*4][adj S*4][mbar 8][ptrs n_slots*8][cur S*256*4][in: smem_data_cap][out: smem_stage_cap]
//   (register-cursor walkers: `cur` overlays `in`)
#pragma once
#include "dev_core.cuh"

namespace rv {

struct Tile {
    int chunk;
    int local_tile;
    int64_t r0;          // first record of the tile
    int nrec;            // records in the tile
    int64_t chunk_len;   // rows in the chunk
};

__device__ __forceinline__ Tile tile_of(const DecodeParams& p, int tile) {
    Tile t;
    int j = 0;
    if (p.k > 1) {  // tile / tiles_per_chunk without the ~40-instruction integer division: estimate and correct
        j = int(__fdividef(float(tile), float(p.tiles_per_chunk)));
        while (j > 0 && j * p.tiles_per_chunk > tile) --j;
        while ((j + 1) * p.tiles_per_chunk <= tile) ++j;
        if (j > p.k - 1) j = p.k - 1;
    }
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

struct TileWindow {
    bool staged;
    bool tma;         // the window is in flight on the copy engine: wait_window() before reading it
    int64_t t0, t1;   // byte range of the tile in `data`
    uint32_t mis;     // (data + t0) & 15: the window starts at the aligned-down address
    int64_t o0, o1;   // this lane's record: byte range in `data`
    int64_t f0, f1;   // (one lane) byte range of the tile `prefetch_dist` ahead, -1: none
};

// L2 prefetch: a CTA's first act is a dependent chain of DRAM round trips (offsets -> input window);
// the CTA that ran `prefetch_dist` tiles earlier has already pulled those lines into L2.
__device__ __forceinline__ void l2_prefetch_line(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
constexpr int kPrefetchLane = 32;  // the lane (first of warp 1) that carries the look-ahead

// ---- TMA (bulk async copy) staging of the input window -----------------------------------------------
// One thread hands the whole window to the copy engine (cp.async.bulk global -> shared, completion counted in
// bytes on an mbarrier) instead of every thread moving its share through registers (LDG.128 + STS.128): the
// prologue loses ~40 issue slots per thread and the loads no longer occupy registers.
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
#if !defined(RV_NO_TMA_STAGE)
__device__ __forceinline__ void mbar_init(uint32_t mbar, uint32_t arrivals) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(arrivals) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");  // visible to the async proxy
}
__device__ __forceinline__ void bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(mbar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "RV_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra RV_DONE;\n"
        "bra RV_WAIT;\n"
        "RV_DONE:\n"
        "}" ::"r"(mbar), "r"(parity) : "memory");
}
#endif

// Loads the plan and the tile's byte window into shared memory.
// EMIT: the lanes' cursors (tile base + in-tile prefix) and, further down, zeroing of the Utf8 staging
// area are issued here too, so their latency overlaps the input loads.
template <bool EMIT, class W>
__device__ __forceinline__ TileWindow stage_in(const DecodeParams& p, const Tile& t, const int tile_id, const SmemMap& m, typename W::Cur& q) {
    const int tid = threadIdx.x;
#if !defined(RV_NO_TMA_STAGE)
    if (tid == 0) mbar_init(smem_addr(rv_smem + m.mbar), 1);
#endif
    if (p.n_nodes) {
        uint4* dn = reinterpret_cast<uint4*>(rv_smem + m.nodes);
        for (int i = tid; i < p.n_nodes * 2; i += kBlock) dn[i] = __ldg(reinterpret_cast<const uint4*>(p.nodes) + i);
    }
    uint32_t* cur = reinterpret_cast<uint32_t*>(rv_smem + m.cur);
    if (EMIT) {
        // this chunk's Arrow buffer pointers (see buf_ptr)
        void** sp = reinterpret_cast<void**>(rv_smem + m.ptrs);
        void* const* gp = p.bufs + size_t(t.chunk) * p.n_slots;
        for (int i = tid; i < p.n_slots; i += kBlock) sp[i] = gp[i];
        const uint32_t* lo = p.lane_off + size_t(tile_id) * p.n_streams * kBlock;
        if constexpr (W::kRegCursors) {
#pragma unroll
            for (int s = 0; s < W::kStreams; ++s)
                q.v[s] = __ldg(lo + s * kBlock + tid) + __ldg(p.tile_base + size_t(s) * p.n_tiles + tile_id);
        } else {
            for (int s = 0; s < p.n_streams; ++s)
                cur[s * kBlock + tid] = __ldg(lo + s * kBlock + tid) + __ldg(p.tile_base + size_t(s) * p.n_tiles + tile_id);
        }
        uint4* z = reinterpret_cast<uint4*>(rv_smem + m.out);
        for (uint32_t i = tid; i < (p.smem_stage_cap >> 4); i += kBlock) z[i] = make_uint4(0u, 0u, 0u, 0u);
    } else if constexpr (!W::kRegCursors) {
        for (int s = 0; s < p.n_streams; ++s) cur[s * kBlock + tid] = 0;
    }
    (void)cur;
    TileWindow w;
    // One round trip: the tile's bounds, this lane's record bounds and (one lane) the bounds of the tile a
    // later CTA will work on are all requested before anything waits.
    w.t0 = __ldg(p.offsets + t.r0);
    w.t1 = __ldg(p.offsets + t.r0 + t.nrec);
    w.o0 = w.o1 = 0;
    if (tid < t.nrec) {
        w.o0 = __ldg(p.offsets + t.r0 + tid);
        w.o1 = __ldg(p.offsets + t.r0 + tid + 1);
    }
    w.f0 = w.f1 = -1;
    const int64_t ft = int64_t(tile_id) + p.prefetch_dist;
    if (p.prefetch_dist > 0 && !p.tile_list && ft < p.n_tiles) {
        const Tile f = tile_of(p, int(ft));
        if (tid < 17 && f.r0 + 16 * tid <= p.n) l2_prefetch_line(p.offsets + f.r0 + 16 * tid);  // the 2 KiB (+8 B) of offsets of that tile
        if (tid == kPrefetchLane) {
            w.f0 = __ldg(p.offsets + f.r0);
            w.f1 = __ldg(p.offsets + f.r0 + f.nrec);
        }
        if (EMIT && tid == kPrefetchLane + 1)
            l2_prefetch_bulk(p.lane_off + size_t(ft) * p.n_streams * kBlock, uint32_t(p.n_streams) * kBlock * 4u);
    }
    const int64_t span = w.t1 - w.t0;
    w.mis = uint32_t(reinterpret_cast<uintptr_t>(p.data + w.t0) & 15u);
    w.staged = span >= 0 && uint64_t(span) + w.mis <= uint64_t(p.smem_data_cap);
    w.tma = false;
    if (w.staged) {
        const uint4* g = reinterpret_cast<const uint4*>(p.data + w.t0 - w.mis);
        const int nvec = int((span + w.mis + 15) >> 4);
        uint4* d = reinterpret_cast<uint4*>(rv_smem + m.in);
#if !defined(RV_NO_TMA_STAGE)
        w.tma = nvec > 0;
        if (w.tma && tid == 0) bulk_load(smem_addr(d), g, uint32_t(nvec) << 4, smem_addr(rv_smem + m.mbar));
#else
        for (int i = tid; i < nvec; i += kBlock) d[i] = __ldg(g + i);
#endif
    }
    return w;
}

// After the CTA barrier that follows stage_in (which also publishes the mbarrier's initialisation).
__device__ __forceinline__ void wait_window(const TileWindow& w, const SmemMap& m) {
#if !defined(RV_NO_TMA_STAGE)
    if (w.tma) mbar_wait(smem_addr(rv_smem + m.mbar), 0);
#else
    (void)w; (void)m;
#endif
}

template <class C>
__device__ __forceinline__ int64_t init_ctx(C& c, const DecodeParams& p, const Tile& t, const SmemMap& m, const TileWindow& w) {
    const int tid = threadIdx.x;
    c.nodes = reinterpret_cast<const DNode*>(rv_smem + m.nodes);
    c.cur = reinterpret_cast<uint32_t*>(rv_smem + m.cur) + tid;
    c.sym_off = p.sym_off;
    c.sym_bytes = p.sym_bytes;
    c.bufs = p.bufs ? p.bufs + size_t(t.chunk) * p.n_slots : nullptr;
    c.ptrs_soff = p.bufs ? m.ptrs : 0u;
    c.err = 0;
    c.pm = 0;
    c.usel = 0;
    c.stage_on = false;
    c.stage_soff = m.out;
    c.stage_adj = reinterpret_cast<const uint32_t*>(rv_smem + m.adj);
    c.in_range = tid < t.nrec;
    c.row0 = uint32_t(t.local_tile) * kBlock + tid;
    c.store_word = (tid & 31) == 0 && int64_t(c.row0) < t.chunk_len;
    c.base = p.data;
    c.soff = m.in;
    c.pos = c.end = 0;
    const int64_t r = t.r0 + tid;
    if (c.in_range) {
        const int64_t o0 = w.o0, o1 = w.o1;
        if (o1 < o0 || o1 - o0 > int64_t(0xFFFFFFF0u)) c.err = E_OVERFLOW;  // malformed offsets / >4 GiB record
        else if (C::kShared) {
            if (o0 < w.t0 || o1 > w.t1) c.err = E_OVERFLOW;
            else { c.pos = uint32_t(o0 - w.t0) + w.mis; c.end = uint32_t(o1 - w.t0) + w.mis; }
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

// The look-ahead lane asks L2 for the input window of the tile `prefetch_dist` ahead (its bounds arrived
// long ago, with this tile's own offsets).
__device__ __forceinline__ void prefetch_window(const DecodeParams& p, const TileWindow& w) {
    if (threadIdx.x == kPrefetchLane && w.f0 >= 0 && w.f1 > w.f0 && w.f1 - w.f0 < (int64_t(1) << 20)) {
        const uintptr_t a = reinterpret_cast<uintptr_t>(p.data + w.f0) & ~uintptr_t(15);
        const uintptr_t e = (reinterpret_cast<uintptr_t>(p.data + w.f1) + 15) & ~uintptr_t(15);
        l2_prefetch_bulk(reinterpret_cast<const void*>(a), uint32_t(e - a));
    }
}

// ---- count ----------------------------------------------------------------------------------
template <class W, bool SM>
__device__ __forceinline__ void count_walk(const DecodeParams& p, const Tile& t, const SmemMap& m, const TileWindow& w, typename W::Cur& q) {
    WalkCtx<SM> c;
    const int64_t r = init_ctx(c, p, t, m, w);
    if constexpr (W::kRegCursors) {
#pragma unroll
        for (int s = 0; s < W::kStreams; ++s) q.v[s] = 0;
    }
#if !defined(RV_ABL_NOCOUNTWALK)
    W::template walk<WM_COUNT>(c, p.n_nodes, q);
#endif
    if (c.in_range && c.err) report(p, r, c.err);
}

// GENERIC = the walker can read records straight from global memory (interpreter).  A specialised
// walker is only instantiated for the shared-memory window; tiles that do not fit are handed to the
// interpreter kernels through p.overflow.
template <class W, bool GENERIC>
__device__ __forceinline__ void count_body(const DecodeParams& p, const int tile_id) {
    const Tile t = tile_of(p, tile_id);
    const SmemMap m = smem_map(p.n_nodes, p.n_streams, p.n_slots, p.smem_data_cap, W::kRegCursors);
    typename W::Cur q;
    const TileWindow w = stage_in<false, W>(p, t, tile_id, m, q);
    __syncthreads();
    wait_window(w, m);
    if (w.staged) count_walk<W, true>(p, t, m, w, q);
    else if constexpr (GENERIC) count_walk<W, false>(p, t, m, w, q);
    else {
        if (threadIdx.x == 0) p.overflow_list[atomicAdd(p.overflow, 1)] = tile_id;
        return;  // uniform: the whole CTA leaves
    }
    __syncthreads();
    // CTA-wide exclusive scan of every stream's lane counts.  The per-record prefixes are saved so the
    // emit kernel does not have to walk the records a second time just to learn where they write.
    // One WARP scans one stream: each lane takes 8 consecutive records (two 128-bit loads), sums them
    // serially, and a single 5-step shuffle scan joins the 32 lane totals — instead of every warp running a
    // shuffle scan per stream plus a cross-warp pass.
    constexpr int kPerLane = kBlock / 32;  // 4, 8, ...: a multiple of 4, so every lane moves whole uint4
    uint32_t* cur = reinterpret_cast<uint32_t*>(rv_smem + m.cur);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if constexpr (W::kRegCursors) {  // the scan area overlays the (now dead) input window
#pragma unroll
        for (int s = 0; s < W::kStreams; ++s) cur[s * kBlock + tid] = q.v[s];
    }
    __syncthreads();
    for (int s = warp; s < p.n_streams; s += kWarps) {
        uint32_t* cl = cur + s * kBlock + lane * kPerLane;
        uint32_t v[kPerLane];
        uint32_t any = 0;
#pragma unroll
        for (int i = 0; i < kPerLane; i += 4) {
            const uint4 x = *reinterpret_cast<const uint4*>(cl + i);
            v[i] = x.x; v[i + 1] = x.y; v[i + 2] = x.z; v[i + 3] = x.w;
            any |= x.x | x.y | x.z | x.w;
        }
        // kBlock values below 2^31 / kBlock cannot overflow 31 bits; anything bigger (a tile of huge zero-width
        // lists) takes the exact 64-bit path below
        if (__any_sync(0xFFFFFFFFu, any >= (0x80000000u / uint32_t(kBlock)))) {
            if (lane == 0) {
                uint32_t* cs = cur + s * kBlock;
                unsigned long long run = 0;
                for (int i = 0; i < kBlock; ++i) {
                    const uint32_t x = cs[i];
                    cs[i] = uint32_t(run);
                    run += x;
                }
                if (run > 0x7FFFFFFFull) { report(p, t.r0, E_OVERFLOW); run = 0x7FFFFFFFull; }
                p.tile_agg[size_t(s) * p.n_tiles + tile_id] = uint32_t(run);
            }
            __syncwarp();
            continue;
        }
        uint32_t tot = 0;
#pragma unroll
        for (int i = 0; i < kPerLane; ++i) {  // lane-local exclusive prefix
            const uint32_t x = v[i];
            v[i] = tot;
            tot += x;
        }
        uint32_t incl = tot;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += u;
        }
        const uint32_t base = incl - tot;
#pragma unroll
        for (int i = 0; i < kPerLane; i += 4)
            *reinterpret_cast<uint4*>(cl + i) = make_uint4(base + v[i], base + v[i + 1], base + v[i + 2], base + v[i + 3]);
        if (lane == 31) p.tile_agg[size_t(s) * p.n_tiles + tile_id] = incl;  // < 2^31 by construction
    }
    __syncthreads();
    uint32_t* lo = p.lane_off + size_t(tile_id) * p.n_streams * kBlock;
    if constexpr (W::kRegCursors) {
#pragma unroll
        for (int s = 0; s < W::kStreams; ++s) lo[s * kBlock + tid] = cur[s * kBlock + tid];
    } else {
        for (int s = 0; s < p.n_streams; ++s) lo[s * kBlock + tid] = cur[s * kBlock + tid];
    }
    prefetch_window(p, w);
}

// ---- emit: staging map ------------------------------------------------------------------------
// Stream s's Utf8 bytes of this tile occupy [tile_base, tile_base + tot) of its Arrow data buffer; in shared
// memory its region starts at a 16-byte boundary plus the destination's misalignment, so the write-out can use
// aligned uint4.  Warp 0 computes the map with a shuffle scan.  Its inputs do not depend on the tile's bytes, so
// they are requested at the very top of the CTA (map_preload) and consumed after the window loads were issued
// (map_finish): one barrier and one exposed L2 round trip less than computing the map after the window arrived.
struct MapPre { uint32_t tb, tt, ga, region; int slot; };

__device__ __forceinline__ MapPre map_preload(const DecodeParams& p, const Tile& t, const int tile_id, const int s) {
    MapPre r{0u, 0u, 0u, 0u, -1};
    if (s < p.n_streams) {
        r.tb = __ldg(p.tile_base + size_t(s) * p.n_tiles + tile_id);
        r.tt = __ldg(p.tile_agg + size_t(s) * p.n_tiles + tile_id);
        r.slot = p.stream_slot[s];
        if (r.slot >= 0) {
            const uint8_t* b = static_cast<const uint8_t*>(p.bufs[size_t(t.chunk) * p.n_slots + r.slot]);
            r.ga = uint32_t(reinterpret_cast<uintptr_t>(b + r.tb) & 15u);
            r.region = (r.tt + r.ga + 15u) & ~15u;
        }
    }
    return r;
}

// warp 0 only
__device__ __forceinline__ void map_finish(const DecodeParams& p, const Tile& t, const int tile_id, const SmemMap& m, MapPre pre) {
    uint32_t* tbase = reinterpret_cast<uint32_t*>(rv_smem + m.wtot);
    uint32_t* tot = reinterpret_cast<uint32_t*>(rv_smem + m.tot);
    uint32_t* adj = reinterpret_cast<uint32_t*>(rv_smem + m.adj);
    const int lane = threadIdx.x & 31;
    uint32_t carry = 0;
    for (int s0 = 0; s0 < p.n_streams; s0 += 32) {
        const int s = s0 + lane;
        if (s0 > 0) pre = map_preload(p, t, tile_id, s);
        uint32_t incl = pre.region;
        for (int d = 1; d < 32; d <<= 1) {
            const uint32_t u = __shfl_up_sync(0xFFFFFFFFu, incl, d);
            if (lane >= d) incl += u;
        }
        if (s < p.n_streams) {
            tbase[s] = pre.tb;
            tot[s] = pre.tt;
            adj[s] = pre.slot >= 0 ? (carry + incl - pre.region + pre.ga) - pre.tb : 0u;  // staging offset of chunk-relative byte o = adj + o
        }
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) tot[p.n_streams] = (p.smem_stage_cap > 0 && carry <= p.smem_stage_cap) ? 1u : 0u;
}

// ---- emit -----------------------------------------------------------------------------------
template <class W, bool SM>
__device__ __forceinline__ void emit_walks(const DecodeParams& p, const Tile& t, const int tile_id, const SmemMap& m, const TileWindow& w, typename W::Cur& q) {
    WalkCtx<SM> c;
    (void)init_ctx(c, p, t, m, w);
    uint32_t* tbase = reinterpret_cast<uint32_t*>(rv_smem + m.wtot);  // reused: [S] tile bases, then [S] region alignments
    uint32_t* tot = reinterpret_cast<uint32_t*>(rv_smem + m.tot);
    uint32_t* adj = reinterpret_cast<uint32_t*>(rv_smem + m.adj);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // (each record's first row / first byte per stream was loaded into `cur` by stage_in: the chunk-relative
    // tile base from scan_kernel + the record's prefix inside the tile from the count kernel; the staging map
    // was written by warp 0 before the barrier that published the window)
    const bool stage_on = p.n_streams > 0 && tot[p.n_streams] != 0;
    // offsets[0] = 0 of every offsets buffer of this chunk (first tile of the chunk only)
    if (t.local_tile == 0) {
        if (p.n_nodes) {
            for (int i = tid; i < p.n_nodes; i += kBlock) {
                const DNode nd = c.nodes[i];
                if (nd.kind == NK_STR || nd.kind == NK_ENUM || nd.kind == NK_LIST || nd.kind == NK_MAP)
                    static_cast<int32_t*>(buf_ptr(c, nd.slot_a))[0] = 0;
            }
        } else {
            W::zero_offsets(c, tid);
        }
    }
    c.stage_on = stage_on;
#if !defined(RV_ABL_NOWALK)
    W::template walk<WM_EMIT>(c, p.n_nodes, q);
#endif
    prefetch_window(p, w);

#if defined(RV_ABL_NOWRITEOUT)
    if (false) {
#else
    if (stage_on) {  // coalesced write-out of the staged Utf8 bytes, one warp per stream at a time
#endif
#if !defined(RV_NO_TMA_STORE)
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // staged bytes -> visible to the copy engine
#endif
        __syncthreads();
        for (int s = warp; s < p.n_streams; s += kWarps) {
            const int slot = p.stream_slot[s];
            const uint32_t n = tot[s];
            if (slot < 0 || n == 0) continue;
            const uint32_t tb = tbase[s];
            uint8_t* g = static_cast<uint8_t*>(buf_ptr(c, slot)) + tb;
            const uint32_t so = m.out + adj[s] + tb;  // rv_smem offset of the region's first byte
            const uint32_t head = min(n, (16u - uint32_t(reinterpret_cast<uintptr_t>(g) & 15u)) & 15u);
            for (uint32_t i = lane; i < head; i += 32) g[i] = rv_smem[so + i];
            const uint32_t nvec = (n - head) >> 4;
#if !defined(RV_NO_TMA_STORE)
            // the 16-byte aligned body leaves through the copy engine (shared -> global bulk store): one
            // instruction per column instead of a store loop, and the warp does not wait for the data to drain
            if (lane == 0 && nvec > 0)
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                             ::"l"(g + head), "r"(smem_addr(rv_smem + so + head)), "r"(nvec << 4) : "memory");
#else
            const uint4* sv = reinterpret_cast<const uint4*>(rv_smem + so + head);
            uint4* gv = reinterpret_cast<uint4*>(g + head);
            for (uint32_t i = lane; i < nvec; i += 32) gv[i] = sv[i];
#endif
            const uint32_t done = head + (nvec << 4);
            for (uint32_t i = done + lane; i < n; i += 32) g[i] = rv_smem[so + i];
        }
#if !defined(RV_NO_TMA_STORE)
        if (lane == 0) {  // shared memory must outlive the engine's reads
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
#endif
    }
}

template <class W, bool GENERIC>
__device__ __forceinline__ void emit_body(const DecodeParams& p, const int tile_id) {
    const Tile t = tile_of(p, tile_id);
    const SmemMap m = smem_map(p.n_nodes, p.n_streams, p.n_slots, p.smem_data_cap, W::kRegCursors);
    typename W::Cur q;
    const bool map_warp = threadIdx.x < 32;
    MapPre pre{0u, 0u, 0u, 0u, -1};
    if (map_warp) pre = map_preload(p, t, tile_id, int(threadIdx.x));
    const TileWindow w = stage_in<true, W>(p, t, tile_id, m, q);
    if (map_warp) map_finish(p, t, tile_id, m, pre);
    __syncthreads();
    wait_window(w, m);
    if constexpr (!GENERIC) {
        // specialised walker: staged input AND staged output only.  (!w.staged: the count pass already put the
        // tile on the overflow list.)
        if (!w.staged) return;
        const uint32_t* tot = reinterpret_cast<const uint32_t*>(rv_smem + m.tot);
        if (p.n_utf8 > 0 && tot[p.n_streams] == 0) {
            if (threadIdx.x == 0) p.overflow_list[atomicAdd(p.overflow, 1)] = tile_id;
            return;  // uniform
        }
        emit_walks<W, true>(p, t, tile_id, m, w, q);
    } else {
        if (w.staged) emit_walks<W, true>(p, t, tile_id, m, w, q);
        else emit_walks<W, false>(p, t, tile_id, m, w, q);
    }
}

}  // namespace rv
