// Types shared by the host planner, the statically compiled kernels, the NVRTC-compiled
// schema-specialised kernels and the tests-only host emulation.  Must stay free of standard
// library includes when compiled by NVRTC (__CUDACC_RTC__).
#pragma once

#if defined(__CUDACC_RTC__)
typedef signed char int8_t;
typedef unsigned char uint8_t;
typedef short int16_t;
typedef unsigned short uint16_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef unsigned long long uintptr_t;
typedef unsigned long size_t;
#else
#include <cstddef>
#include <cstdint>
#endif

namespace rv {

// ---- device-visible node of the generic (interpreted) walker ------------------------------
enum NodeKind : uint8_t {
    NK_I32 = 0,   // int, date            (varint -> i32, `as i32` truncation, fast_decode.rs:424,430)
    NK_I64 = 1,   // long, timestamp-*    (varint -> i64)
    NK_F32 = 2,   // float   (4 raw LE bytes, :871-879)
    NK_F64 = 3,   // double  (8 raw LE bytes, :881-891)
    NK_BOOL = 4,  // boolean (:893-900), LSB-first bit column
    NK_STR = 5,   // string  (:902-922)
    NK_ENUM = 6,  // enum -> Utf8 symbol text (:570-578)
    NK_NULL = 7,  // null (:480)
    NK_REC = 8,   // record (:597-616)
    NK_UNION = 9, // N-variant sparse union (:643-668)
    NK_LIST = 10, // array (:703-727)
    NK_MAP = 11   // map (:745-770); children = [keys (NK_STR), values]
};

enum NodeFlags : uint8_t {
    NF_NULLABLE = 1,     // wrapped in a 2-variant null union (Nullable* variants, :94-119)
    NF_NULL_FIRST = 2,   // which branch index is null (:404-414)
    NF_VALIDITY = 4,     // a validity bitmap is written for this node
    NF_ZERO_ITEMS = 8    // list/map whose items occupy zero bytes and own no buffers
};

struct DNode {
    uint8_t kind;
    uint8_t flags;
    uint8_t level;     // depth in the node tree (root record's children are level 1)
    uint8_t ulevel;    // number of NK_UNION ancestors
    uint8_t variant;   // index within the parent union, 0xFF otherwise
    uint8_t space;     // row space of this node's rows
    uint8_t pad0, pad1;
    int32_t end;       // one past the last node of this subtree (pre-order)
    int16_t slot_v;    // validity bitmap slot (-1: none)
    int16_t slot_a;    // values / offsets / type_ids slot (-1: none)
    int16_t slot_b;    // Utf8 data slot (-1: none)
    int16_t stream;    // NK_STR/NK_ENUM: byte stream; NK_LIST/NK_MAP: child-row stream; else -1
    int32_t aux;       // NK_ENUM: first entry in the symbol-offset table; NK_UNION: variant count
    int32_t aux2;      // NK_ENUM: symbol count
    int32_t pad2;
};
static_assert(sizeof(DNode) == 32, "DNode layout is shared with the kernels");

constexpr int kMaxListDepth = 3;   // row-space nesting the kernels are instantiated for
constexpr int kMaxLevel = 31;      // per-lane presence mask is 32 bits
constexpr int kMaxUnionLevel = 8;  // per-lane union selections packed 8 x 8 bits
constexpr int kMaxStreams = 120;
constexpr int kMaxNodes = 1024;

// Per-record error categories (the bail!/anyhow! sites of fast_decode.rs); values match rv_status.
enum ErrCode : uint32_t {
    E_OK = 0,
    E_EOF = 1,       // "unexpected end of buffer" (:849,874,884,910)
    E_VARINT = 2,    // "zigzag varint too long" (:866)
    E_BOOL = 3,      // "invalid boolean byte" (:898)
    E_NEG_LEN = 4,   // "negative string length" (:906)
    E_BRANCH = 5,    // "invalid union branch index" / "out of range" (:591,646)
    E_ENUM = 6,      // "enum index out of range" (:575)
    E_SCHEMA = 7,
    E_OVERFLOW = 8   // i32 Arrow offset overflow (arrow-rs panics; reported as an error)
};

#ifndef RV_KBLOCK
#define RV_KBLOCK 256
#endif
constexpr int kBlock = RV_KBLOCK;  // records per tile == threads per CTA (one record per lane)
static_assert(kBlock % 128 == 0 && kBlock <= 1024, "tiles are whole groups of 4 warps");
constexpr int kWarps = kBlock / 32;

// Kernel parameter block (count / scan / emit).
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
    const int16_t* stream_slot;  // [n_streams] Utf8 data slot of a byte stream, -1 for a row stream
    // scan scratch
    uint32_t* tile_agg;    // [n_streams][n_tiles] per-tile totals
    uint32_t* tile_base;   // [n_streams][n_tiles] exclusive prefix within the chunk
    uint32_t* lane_off;    // [n_tiles][n_streams][kBlock] each record's exclusive prefix inside its tile
    unsigned long long* chunk_tot;  // [k][n_streams]
    unsigned long long* err;        // min over (record << 8 | code); ~0 = none
    // output
    void* const* bufs;     // [k][n_slots]
    // Tile routing between the two walkers.  The schema-specialised kernels only handle tiles whose
    // bytes fit the shared-memory window; they append the others to `overflow_list` (`overflow[0]` counts them) and the generic interpreter kernels are launched over that list (tile_list != nullptr).
    const int32_t* tile_list;  // interpreter pass over overflow tiles: blockIdx.x -> tile id
    int32_t* overflow;         // specialised count pass: number of tiles it skipped ...
    int32_t* overflow_list;    // ... and their ids (the specialised emit pass appends tiles whose strings do not fit
                               // its staging area; the interpreter emit pass handles both kinds)
    int32_t n_utf8;            // Utf8 byte streams in the plan
    int32_t prefetch_dist;    // CTAs resident on the device: a CTA prefetches (into L2) the tile that far ahead
    uint32_t smem_data_cap;   // bytes of shared memory for staging a tile's input bytes
    uint32_t smem_stage_cap;  // bytes of shared memory for staging a tile's Utf8 output bytes (emit)
};

// Dynamic shared-memory map of count/emit CTAs (byte offsets inside the CTA's shared memory):
//   [nodes n_nodes*32][wtot S*8*4][tot (S+1)*4][adj S*4][ptrs n_slots*8][cur S*256*4][in: data_cap][out: stage_cap]
// alias_cur (walkers that keep their cursors in registers): the scan area `cur` overlays the input window
// (it is only written after every lane finished reading the window) and costs no extra shared memory.
struct SmemMap {
    uint32_t nodes, wtot, tot, adj, mbar, ptrs, cur, in, out;
};

#if defined(__CUDACC__)
__host__ __device__
#endif
inline SmemMap smem_map(int n_nodes, int n_streams, int n_slots, uint32_t data_cap, bool alias_cur) {
    SmemMap m;
    m.nodes = 0;
    m.wtot = uint32_t(n_nodes) * 32u;
    m.tot = m.wtot + uint32_t(n_streams) * kWarps * 4u;
    m.adj = m.tot + uint32_t(n_streams + 1) * 4u;
    m.mbar = (m.adj + uint32_t(n_streams) * 4u + 7u) & ~7u;  // mbarrier of the bulk-copy staging (8 bytes)
    m.ptrs = (m.mbar + 8u + 15u) & ~15u;
    m.cur = (m.ptrs + uint32_t(n_slots) * 8u + 15u) & ~15u;
    const uint32_t cur_bytes = uint32_t(n_streams) * kBlock * 4u;
    if (alias_cur) {
        m.in = m.cur;
        const uint32_t win = data_cap > cur_bytes ? data_cap : cur_bytes;
        m.out = (m.in + win + 15u) & ~15u;
    } else {
        m.in = (m.cur + cur_bytes + 15u) & ~15u;
        m.out = (m.in + data_cap + 15u) & ~15u;
    }
    return m;
}

}  // namespace rv
