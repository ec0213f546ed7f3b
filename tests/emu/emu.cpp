// TEST INFRASTRUCTURE ONLY — host emulation of the GPU decode pipeline.
//
// Compiles the product's walker (pyruhvro_b200/csrc/walker.cuh), plan, layout and Arrow export
// for the CPU and steps through the same count -> per-chunk scan -> exact layout -> emit ->
// null-count sequence the kernels run, one 256-"lane" tile at a time.  It exists so the decode
// LOGIC can be checked against the oracle in the GPU-less build container; it is never loaded by
// the product (pyruhvro_b200 has no CPU path) and is not what the -m gpu tests exercise.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <stdexcept>
#include <algorithm>
#include <memory>
#include <string>
#include <vector>

#include "../../pyruhvro_b200/csrc/plan.hpp"
#include "../../pyruhvro_b200/csrc/result.hpp"
#include "../../pyruhvro_b200/csrc/schema.hpp"
#include "../../pyruhvro_b200/csrc/interp.cuh"
#ifdef EMU_GEN_WALKER
#include EMU_GEN_WALKER   // schema-specialised walker source produced by rv_schema_walker_source()
using EmuWalker = rv::gen::Walker;
#else
using EmuWalker = rv::InterpWalker;
#endif

using namespace rv;

namespace {
constexpr int kTile = rv::kBlock;

struct Tile { int chunk; int64_t r0; int nrec; int local; };

std::vector<Tile> make_tiles(int64_t n, int k) {
    std::vector<Tile> t;
    const int64_t cr = n / k;
    for (int j = 0; j < k; ++j) {
        const int64_t cs = int64_t(j) * cr, ce = (j == k - 1) ? n : cs + cr;
        int local = 0;
        for (int64_t r = cs; r < ce; r += kTile) t.push_back(Tile{j, r, int(std::min<int64_t>(kTile, ce - r)), local++});
    }
    return t;
}

using Ctx = WalkCtx<false>;      // PRECISE flavour: reads the record where it lies
using FastCtx = WalkCtx<true>;   // FAST flavour: reads a padded copy of the tile's byte window, like the staged shared-memory window

// register-cursor walkers exchange their cursors with the emulation's cur[stream][lane] table
template <class Q>
void store_cursors(const Q& q, uint32_t* cur_lane, int S) {
    if constexpr (EmuWalker::kRegCursors) for (int s = 0; s < S; ++s) cur_lane[size_t(s) * kTile] = q.v[s];
}
template <class Q>
void load_cursors(Q& q, const uint32_t* cur_lane, int S) {
    if constexpr (EmuWalker::kRegCursors) for (int s = 0; s < S; ++s) q.v[s] = cur_lane[size_t(s) * kTile];
}

template <class C>
void init_common(C& c, const Plan& plan, const Tile& t, int lane, uint32_t* cur, void* const* bufs) {
    std::memset(&c, 0, sizeof c);
    c.nodes = plan.nodes.data();
    c.cur = cur + lane;
    c.cur_stride = kTile;
    c.sym_off = plan.sym_off.data();
    c.sym_bytes = plan.sym_bytes.data();
    c.bufs = bufs;
    c.in_range = lane < t.nrec;
    c.row0 = uint32_t(t.local) * kTile + uint32_t(lane);
    c.store_word = false;
}

void init_ctx(Ctx& c, const Plan& plan, const uint8_t* data, const int64_t* off, const Tile& t, int lane,
              uint32_t* cur, int S, void* const* bufs) {
    init_common(c, plan, t, lane, cur, bufs);
    (void)S;
    if (c.in_range) {
        const int64_t r = t.r0 + lane;
        c.base = data + off[r];
        c.pos = 0;
        c.end = uint32_t(off[r + 1] - off[r]);
    }
}

// The tile's bytes followed by kWindowPad bytes of junk (the device leaves whatever was in shared memory there).
std::vector<uint8_t> make_window(const uint8_t* data, const int64_t* off, const Tile& t) {
    const int64_t t0 = off[t.r0], t1 = off[t.r0 + t.nrec];
    std::vector<uint8_t> w(size_t(t1 - t0) + kWindowPad, uint8_t(0xA5));
    if (t1 > t0) std::memcpy(w.data(), data + t0, size_t(t1 - t0));
    return w;
}

void init_fast(FastCtx& c, const Plan& plan, const std::vector<uint8_t>& window, const int64_t* off, const Tile& t, int lane,
               uint32_t* cur, void* const* bufs) {
    init_common(c, plan, t, lane, cur, bufs);
    c.base = window.data();
    if (c.in_range) {
        const int64_t r = t.r0 + lane, t0 = off[t.r0];
        c.pos = uint32_t(off[r] - t0);
        c.end = uint32_t(off[r + 1] - t0);
    }
}

// COUNT of one tile the way fused_body does it: every lane walks FAST; a lane whose record is not plain repeats it
// PRECISE (which alone decides errors); a 32-lane "warp" with any such lane is marked for the precise EMIT.
// Returns the first error's lane or -1.
int count_tile(const Plan& plan, const uint8_t* data, const int64_t* off, const Tile& t, int S, std::vector<uint32_t>& cur,
               bool* warp_precise, uint32_t* err_code, long long* fast_lanes, long long* precise_lanes) {
    std::fill(cur.begin(), cur.end(), 0u);
    const std::vector<uint8_t> window = make_window(data, off, t);
    for (int w = 0; w < kTile / 32; ++w) warp_precise[w] = false;
    for (int lane = 0; lane < kTile; ++lane) {
        FastCtx f;
        init_fast(f, plan, window, off, t, lane, cur.data(), nullptr);
        EmuWalker::Cur q{};
        EmuWalker::walk<WM_COUNT>(f, int(plan.nodes.size()), q);
        if (f.in_range && f.pos > f.end + kWindowPad) { std::fprintf(stderr, "emu: fast reader ran %u bytes past the record\n", f.pos - f.end); std::abort(); }
        if (f.in_range && f.err) {
            warp_precise[lane / 32] = true;
            for (int s = 0; s < S; ++s) cur[size_t(s) * kTile + lane] = 0;
            Ctx c;
            init_ctx(c, plan, data, off, t, lane, cur.data(), S, nullptr);
            q = EmuWalker::Cur{};
            EmuWalker::walk<WM_COUNT>(c, int(plan.nodes.size()), q);
            store_cursors(q, cur.data() + lane, S);
            if (precise_lanes) ++*precise_lanes;
            if (c.err) { *err_code = c.err; return lane; }
        } else {
            store_cursors(q, cur.data() + lane, S);
            if (fast_lanes && f.in_range) ++*fast_lanes;
        }
    }
    return -1;
}
}  // namespace

extern "C" {

// Returns 0 on success; on a data error returns the error code and sets *err_record.
// out_batches must have room for clamp_chunks(num_chunks, n) ArrowArrays.
int emu_decode(const char* json, size_t len, const uint8_t* data, const int64_t* off, int64_t n, int64_t num_chunks,
               ArrowArray* out_batches, ArrowSchema* out_schema, int64_t* k_out, int64_t* err_record, char* msg, size_t msg_cap) {
    try {
        auto avro = parse_avro_schema(json, len);
        std::string why;
        if (!is_supported(*avro, &why)) throw std::runtime_error("unsupported: " + why);
        auto fields = to_arrow_fields(*avro);
        auto plan_sp = std::make_shared<Plan>(build_plan(*avro, fields));
        const Plan& plan = *plan_sp;
        const int S = int(plan.streams.size()), S1 = std::max(S, 1);
        const int k = int(clamp_chunks(num_chunks, n));
        *k_out = k;
        auto tiles = make_tiles(n, k);
        std::vector<uint32_t> cur(size_t(S1) * kTile);
        std::vector<std::vector<uint32_t>> tile_agg(tiles.size(), std::vector<uint32_t>(size_t(S1), 0));
        // ---- count pass ----
        std::vector<char> precise_warp(tiles.size() * size_t(kTile / 32), 0);
        long long fast_lanes = 0, precise_lanes = 0;
        for (size_t ti = 0; ti < tiles.size(); ++ti) {
            bool wp[kTile / 32];
            uint32_t code = 0;
            const int bad = count_tile(plan, data, off, tiles[ti], S, cur, wp, &code, &fast_lanes, &precise_lanes);
            if (bad >= 0) { *err_record = tiles[ti].r0 + bad; return int(code); }
            for (int w = 0; w < kTile / 32; ++w) precise_warp[ti * size_t(kTile / 32) + size_t(w)] = wp[w];
            for (int s = 0; s < S; ++s) {
                uint64_t sum = 0;
                for (int lane = 0; lane < kTile; ++lane) sum += cur[size_t(s) * kTile + lane];
                if (sum > 0x7FFFFFFFull) { *err_record = tiles[ti].r0; return int(E_OVERFLOW); }
                tile_agg[ti][size_t(s)] = uint32_t(sum);
            }
        }
        if (std::getenv("EMU_TRACE")) std::fprintf(stderr, "emu: %lld fast lanes, %lld precise lanes\n", fast_lanes, precise_lanes);
        // ---- per-chunk scan ----
        std::vector<unsigned long long> chunk_tot(size_t(k) * size_t(S1), 0ull);
        std::vector<std::vector<uint32_t>> tile_base(tiles.size(), std::vector<uint32_t>(size_t(S1), 0));
        for (size_t ti = 0; ti < tiles.size(); ++ti)
            for (int s = 0; s < S; ++s) {
                unsigned long long& tot = chunk_tot[size_t(tiles[ti].chunk) * size_t(S1) + size_t(s)];
                tile_base[ti][size_t(s)] = uint32_t(tot);
                tot += tile_agg[ti][size_t(s)];
                if (tot > 0x7FFFFFFFull) { *err_record = tiles[ti].r0; return int(E_OVERFLOW); }
            }
        // ---- layout + arena ----
        Layout L = compute_layout(plan, n, k, chunk_tot.data());
        struct Keep { std::shared_ptr<Plan> plan; std::vector<ChunkOut> chunks; uint8_t* arena; ~Keep() { std::free(arena); } };
        auto keep = std::make_shared<Keep>();
        keep->plan = plan_sp;
        keep->arena = static_cast<uint8_t*>(std::calloc(std::max<size_t>(L.total_bytes, 64), 1));
        // poison everything that the device does not zero, to catch missing writes
        std::memset(keep->arena + L.zero_bytes, 0, L.total_bytes - L.zero_bytes);
        const int n_slots = int(plan.slots.size());
        std::vector<void*> bufs(size_t(k) * size_t(std::max(n_slots, 1)));
        for (int j = 0; j < k; ++j)
            for (int sl = 0; sl < n_slots; ++sl) bufs[size_t(j) * size_t(n_slots) + size_t(sl)] = keep->arena + L.chunks[size_t(j)].slot_off[size_t(sl)];
        // ---- emit pass ----
        for (size_t ti = 0; ti < tiles.size(); ++ti) {
            const Tile& t = tiles[ti];
            void* const* cb = bufs.data() + size_t(t.chunk) * size_t(n_slots);
            bool wp[kTile / 32];
            uint32_t code = 0;
            (void)count_tile(plan, data, off, t, S, cur, wp, &code, nullptr, nullptr);
            for (int s = 0; s < S; ++s) {  // exclusive scan across lanes + tile base
                uint32_t run = tile_base[ti][size_t(s)];
                for (int lane = 0; lane < kTile; ++lane) { uint32_t v = cur[size_t(s) * kTile + lane]; cur[size_t(s) * kTile + lane] = run; run += v; }
            }
            if (t.local == 0)
                for (const DNode& nd : plan.nodes)
                    if (nd.kind == NK_STR || nd.kind == NK_ENUM || nd.kind == NK_LIST || nd.kind == NK_MAP || nd.kind == NK_BYTES) static_cast<int32_t*>(cb[nd.slot_a])[0] = 0;
            const std::vector<uint8_t> window = make_window(data, off, t);
            for (int lane = 0; lane < kTile; ++lane) {
                EmuWalker::Cur q{};
                load_cursors(q, cur.data() + lane, S);
                if (wp[lane / 32]) {  // a warp with a record that is not plain emits with the precise walker
                    Ctx c;
                    init_ctx(c, plan, data, off, t, lane, cur.data(), S, cb);
                    EmuWalker::walk<WM_EMIT>(c, int(plan.nodes.size()), q);
                } else {
                    FastCtx f;
                    init_fast(f, plan, window, off, t, lane, cur.data(), cb);
                    EmuWalker::walk<WM_EMIT>(f, int(plan.nodes.size()), q);
                }
            }
        }
        // ---- null counts ----
        for (int j = 0; j < k; ++j)
            for (int sl : plan.validity_slots) {
                ChunkOut& c = L.chunks[size_t(j)];
                const int64_t bits = c.space_rows[size_t(plan.slots[size_t(sl)].space)];
                const uint8_t* bm = keep->arena + c.slot_off[size_t(sl)];
                int64_t ones = 0;
                for (int64_t i = 0; i < bits; ++i) ones += (bm[i >> 3] >> (i & 7)) & 1;
                c.null_count[size_t(sl)] = bits - ones;
            }
        keep->chunks = L.chunks;
        for (int j = 0; j < k; ++j) export_batch(plan, keep->chunks[size_t(j)], keep->arena, keep, &out_batches[j]);
        if (out_schema) export_arrow_schema(fields, out_schema);
        return 0;
    } catch (const std::exception& e) {
        std::snprintf(msg, msg_cap, "%s", e.what());
        return -1;
    }
}

}  // extern "C"
