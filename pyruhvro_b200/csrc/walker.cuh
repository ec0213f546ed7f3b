// Lock-step record walker: the per-record byte walk of the reference
// (FieldDecoder::decode / append_null, ruhvro/src/fast_decode.rs:420-534, and the
// Record/Union/List/Map decoders :595-799; wire primitives :845-922), re-shaped for a
// warp in which every lane owns one record.
//
// Design: all lanes step through the SAME pre-order node list (the program counter is
// warp-uniform); what differs per lane is a presence predicate.  A node that is absent
// for a lane (null branch, unselected union variant, null parent struct) is exactly the
// reference's `append_null` on that column: the lane still writes the null slot.  Only
// array/map item loops have lane-dependent trip counts.  The walk runs twice per record:
//   COUNT  no stores; accumulates per-stream contributions (bytes of each Utf8 column,
//          rows of each array/map child space) into cur[stream].
//   EMIT   cur[stream] now holds this record's first byte / first row (after the scans);
//          values, offsets, validity and string bytes are written to the Arrow buffers.
//
// The same source compiles for the device (kernels.cu) and, for logic tests without a GPU,
// for the host (tests/emu) — the host build is test infrastructure only.
#pragma once
#include <cstdint>

#include "plan.hpp"

#if defined(__CUDACC__)
#define RV_HD __host__ __device__ __forceinline__
#else
#define RV_HD inline
#endif

namespace rv {


enum WalkMode : int { WM_COUNT = 0, WM_EMIT = 1 };

struct WalkCtx {
    const uint8_t* base;   // window holding this record (shared-memory tile or global)
    uint32_t pos, end;     // cursor / record end, relative to base
    uint32_t err;          // first error of this record
    uint32_t pm;           // presence by tree level: bit L = node at level L is present & valid
    uint64_t usel;         // selected variant per union nesting level (8 bits each)
    const DNode* nodes;
    uint32_t* cur;         // per-lane cursors: cur[stream * cur_stride]
    uint32_t cur_stride;
    void* const* bufs;     // slot -> buffer of this chunk
    const int32_t* sym_off;
    const uint8_t* sym_bytes;
    uint32_t row0;         // chunk-local row of this record
    bool in_range;         // the lane owns a record
    bool store_word;       // space-0 bitmaps: this lane stores the warp's ballot word
};

RV_HD void fail(WalkCtx& c, uint32_t code) {
    if (!c.err) c.err = code;
}

// read_zigzag_long, fast_decode.rs:854-869
RV_HD int64_t rd_varint(WalkCtx& c) {
    uint64_t r = 0;
    uint32_t shift = 0;
    for (;;) {
        if (c.pos >= c.end) { fail(c, E_EOF); return 0; }
        const uint32_t b = c.base[c.pos++];
        r |= uint64_t(b & 0x7Fu) << shift;
        if (!(b & 0x80u)) break;
        shift += 7;
        if (shift >= 64) { fail(c, E_VARINT); return 0; }
    }
    return int64_t(r >> 1) ^ -int64_t(r & 1);
}

RV_HD void copy_bytes(uint8_t* dst, const uint8_t* src, uint32_t n) {
    for (uint32_t i = 0; i < n; ++i) dst[i] = src[i];
}

// Bit `row` of a validity / boolean buffer.  Space 0: rows are lane-aligned, so the warp
// ballots and one lane stores a whole 32-bit word.  Deeper spaces: rows are lane-private
// cursors, so set bits go through atomicOr into a zero-initialised buffer.
template <int D>
RV_HD void put_bit(WalkCtx& c, int slot, uint32_t row, bool bit) {
#if defined(__CUDA_ARCH__)
    if (D == 0) {
        const unsigned w = __ballot_sync(0xFFFFFFFFu, bit);
        if (c.store_word) static_cast<uint32_t*>(c.bufs[slot])[row >> 5] = w;
    } else {
        if (bit) atomicOr(static_cast<unsigned int*>(c.bufs[slot]) + (row >> 5), 1u << (row & 31));
    }
#else
    if (bit && (D > 0 || c.in_range)) static_cast<uint8_t*>(c.bufs[slot])[row >> 3] |= uint8_t(1u << (row & 7));
#endif
}

template <int MODE, int D>
RV_HD void walk_range(WalkCtx& c, int pc, const int end, const uint32_t row) {
    const bool st = (D > 0) || c.in_range;  // this lane may store at `row`
    while (pc < end) {
        const DNode nd = c.nodes[pc];
        // --- presence: parent present, and (inside a union) this variant selected ---
        bool present = (c.err == 0) && ((c.pm >> (nd.level - 1)) & 1u);
        if (nd.variant != 0xFF) present = present && (uint32_t((c.usel >> (8 * (nd.ulevel - 1))) & 0xFF) == nd.variant);
        // --- 2-variant null union: union_branch, fast_decode.rs:585-593 ---
        bool valid = present;
        if ((nd.flags & NF_NULLABLE) && present) {
            const int64_t idx = rd_varint(c);
            if (c.err) valid = false;
            else if (idx == 0 || idx == 1) valid = (idx == 1) == ((nd.flags & NF_NULL_FIRST) != 0);
            else { fail(c, E_BRANCH); valid = false; }
        }
        switch (nd.kind) {
            case NK_I32: {
                int32_t v = 0;
                if (valid) { const int64_t x = rd_varint(c); if (!c.err) v = int32_t(x); else valid = false; }
                if (MODE == WM_EMIT && st) static_cast<int32_t*>(c.bufs[nd.slot_a])[row] = v;
                break;
            }
            case NK_I64: {
                int64_t v = 0;
                if (valid) { const int64_t x = rd_varint(c); if (!c.err) v = x; else valid = false; }
                if (MODE == WM_EMIT && st) static_cast<int64_t*>(c.bufs[nd.slot_a])[row] = v;
                break;
            }
            case NK_F32: {  // read_f32 :871-879
                uint32_t v = 0;
                if (valid) {
                    if (c.end - c.pos < 4) { fail(c, E_EOF); valid = false; }
                    else {
                        const uint8_t* p = c.base + c.pos;
                        v = uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
                        c.pos += 4;
                    }
                }
                if (MODE == WM_EMIT && st) static_cast<uint32_t*>(c.bufs[nd.slot_a])[row] = v;
                break;
            }
            case NK_F64: {  // read_f64 :881-891
                uint64_t v = 0;
                if (valid) {
                    if (c.end - c.pos < 8) { fail(c, E_EOF); valid = false; }
                    else {
                        const uint8_t* p = c.base + c.pos;
                        const uint32_t lo = uint32_t(p[0]) | (uint32_t(p[1]) << 8) | (uint32_t(p[2]) << 16) | (uint32_t(p[3]) << 24);
                        const uint32_t hi = uint32_t(p[4]) | (uint32_t(p[5]) << 8) | (uint32_t(p[6]) << 16) | (uint32_t(p[7]) << 24);
                        v = uint64_t(lo) | (uint64_t(hi) << 32);
                        c.pos += 8;
                    }
                }
                if (MODE == WM_EMIT && st) static_cast<uint64_t*>(c.bufs[nd.slot_a])[row] = v;
                break;
            }
            case NK_BOOL: {  // read_bool :893-900
                bool v = false;
                if (valid) {
                    if (c.pos >= c.end) { fail(c, E_EOF); valid = false; }
                    else {
                        const uint32_t b = c.base[c.pos++];
                        if (b > 1) { fail(c, E_BOOL); valid = false; } else v = b != 0;
                    }
                }
                if (MODE == WM_EMIT) put_bit<D>(c, nd.slot_a, row, v);
                break;
            }
            case NK_STR:    // read_string :902-922 (no UTF-8 validation, like the reference)
            case NK_ENUM: { // append_enum :570-578
                uint32_t len = 0;
                const uint8_t* src = c.base;
                if (valid) {
                    const int64_t l = rd_varint(c);
                    if (c.err) valid = false;
                    else if (nd.kind == NK_STR) {
                        if (l < 0) { fail(c, E_NEG_LEN); valid = false; }
                        else if (uint64_t(l) > uint64_t(c.end - c.pos)) { fail(c, E_EOF); valid = false; }
                        else { len = uint32_t(l); src = c.base + c.pos; c.pos += len; }
                    } else {
                        if (uint64_t(l) >= uint64_t(uint32_t(nd.aux2))) { fail(c, E_ENUM); valid = false; }
                        else {
                            const int32_t b0 = c.sym_off[nd.aux + int32_t(l)];
                            len = uint32_t(c.sym_off[nd.aux + int32_t(l) + 1] - b0);
                            src = c.sym_bytes + b0;
                        }
                    }
                }
                uint32_t& cur = c.cur[uint32_t(nd.stream) * c.cur_stride];
                if (MODE == WM_COUNT) {
                    const uint32_t nxt = cur + len;
                    if (nxt < cur) fail(c, E_OVERFLOW);
                    cur = nxt;
                } else {
                    uint32_t o = cur;
                    if (len) copy_bytes(static_cast<uint8_t*>(c.bufs[nd.slot_b]) + o, src, len);
                    o += len;
                    if (st) static_cast<int32_t*>(c.bufs[nd.slot_a])[row + 1] = int32_t(o);
                    cur = o;
                }
                break;
            }
            case NK_NULL:
                break;
            case NK_REC: {  // decode_present / append_null :597-616 — children follow in pre-order
                c.pm = (c.pm & ~(1u << nd.level)) | (uint32_t(valid) << nd.level);
                if (MODE == WM_EMIT && (nd.flags & NF_VALIDITY)) put_bit<D>(c, nd.slot_v, row, valid);
                ++pc;
                continue;
            }
            case NK_UNION: {  // UnionDecoder::decode / append_null :643-668
                uint32_t sel = 0xFE;  // matches no variant: every child appends null
                int8_t tid = 0;
                bool ok = false;
                if (valid) {
                    const int64_t idx = rd_varint(c);
                    if (!c.err) {
                        if (idx < 0 || idx >= int64_t(nd.aux)) fail(c, E_BRANCH);
                        else { sel = uint32_t(idx); tid = int8_t(idx); ok = true; }
                    }
                }
                c.pm = (c.pm & ~(1u << nd.level)) | (uint32_t(ok) << nd.level);
                c.usel = (c.usel & ~(uint64_t(0xFF) << (8 * nd.ulevel))) | (uint64_t(sel) << (8 * nd.ulevel));
                if (MODE == WM_EMIT && st) static_cast<int8_t*>(c.bufs[nd.slot_a])[row] = tid;
                ++pc;
                continue;
            }
            case NK_LIST:
            case NK_MAP: {  // ListDecoder / MapDecoder :703-727, :745-770; read_block_count :689-700
                uint32_t& cur = c.cur[uint32_t(nd.stream) * c.cur_stride];
                const uint32_t r = (MODE == WM_EMIT) ? cur : 0u;  // first child row of this list
                uint32_t total = 0;
                if (valid) {
                    c.pm |= (1u << nd.level);
                    int64_t rem = 0;
                    for (;;) {
                        if (rem == 0) {
                            int64_t n = rd_varint(c);
                            if (c.err) break;
                            if (n < 0) {
                                (void)rd_varint(c);  // block byte size: ignored
                                if (c.err) break;
                                n = int64_t(0 - uint64_t(n));
                                if (n < 0) continue;  // i64::MIN: empty range in the reference
                            }
                            if (n == 0) break;
                            if (nd.flags & NF_ZERO_ITEMS) {  // items are zero bytes wide and own no buffers
                                if (n > int64_t(0x7FFFFFFF) - int64_t(total)) { fail(c, E_OVERFLOW); break; }
                                total += uint32_t(n);
                                continue;
                            }
                            rem = n;
                        }
                        if constexpr (D < kMaxListDepth) walk_range<MODE, D + 1>(c, pc + 1, nd.end, r + total);
                        if (c.err) break;
                        ++total;
                        --rem;
                    }
                }
                if (MODE == WM_COUNT) {
                    const uint32_t nxt = cur + total;
                    if (nxt < cur || nxt > 0x7FFFFFFFu) fail(c, E_OVERFLOW);
                    cur = nxt;
                } else {
                    cur = r + total;
                    if (st) static_cast<int32_t*>(c.bufs[nd.slot_a])[row + 1] = int32_t(r + total);
                    if (nd.flags & NF_VALIDITY) put_bit<D>(c, nd.slot_v, row, valid);
                }
                pc = nd.end;
                continue;
            }
            default:
                fail(c, E_SCHEMA);
                break;
        }
        // leaves: lazily-exported validity (arrow-rs NullBufferBuilder semantics)
        if (MODE == WM_EMIT && (nd.flags & NF_VALIDITY)) put_bit<D>(c, nd.slot_v, row, valid);
        ++pc;
    }
}

// One record, from the top-level record's children (decode_with_arrow_schema, :824-828).
template <int MODE>
RV_HD void walk_record(WalkCtx& c, int n_nodes) {
    c.pm = (c.in_range && c.err == 0) ? 1u : 0u;
    c.usel = 0;
    walk_range<MODE, 0>(c, 0, n_nodes, c.row0);
}

}  // namespace rv
