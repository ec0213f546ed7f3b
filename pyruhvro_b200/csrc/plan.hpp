// Decode plan: the GPU analogue of the reference's FieldDecoder tree
// (ruhvro/src/fast_decode.rs:73-167, built by make_* at :176-414).
//
// The tree is flattened in pre-order into `DNode`s that every lane of a warp
// steps through in lock step.  Three derived notions drive the kernels:
//
//   row space   space 0 = the records of a chunk; every list/map node opens a new
//               space whose rows are its items.  A node's rows live in one space.
//   stream      a quantity that needs a prefix sum across records: the row count a
//               record contributes to a space (>0), or the bytes a record contributes
//               to one Utf8 column.  Per-record counts are scanned to get each
//               record's first row / first byte in the Arrow buffers.
//   slot        one Arrow buffer of one node (validity / values / offsets / data /
//               type_ids); the kernels address output memory as bufs[slot].
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "schema.hpp"

namespace rv {

// ---- device-visible node ----------------------------------------------------
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

// ---- host-side description ----------------------------------------------------
enum class SlotRole : uint8_t { Validity, Bits, Values32, Values64, Offsets, Data, TypeIds };

struct Slot {
    SlotRole role;
    int node;      // owning DNode index
    int space;     // row space whose row count sizes this buffer (Data: unused)
    int stream;    // Data: the byte stream that sizes it
    bool zero_init;  // written with atomicOr (bit buffers in spaces > 0): must start zeroed
};

struct Stream {
    bool is_rows;  // true: rows of `space`; false: bytes of Utf8 node `node`
    int space;
    int node;
};

// One Arrow array of the output tree (what finish() returns, fast_decode.rs:536-567).
struct OutArray {
    AT type;
    int node;          // DNode that owns the buffers (-1 for a map's synthetic entries struct)
    int space;         // row space giving its length
    bool always_validity;  // nullable record/list/map: bitmap always exported (:629,739,790)
    int slot_v, slot_a, slot_b;
    std::vector<int> children;  // indices into Plan::arrays
};

struct Plan {
    std::vector<DNode> nodes;
    std::vector<Slot> slots;
    std::vector<Stream> streams;
    std::vector<int> space_stream;   // space -> its row stream (space 0: -1)
    std::vector<int> space_depth;    // space -> list nesting depth (space 0: 0)
    std::vector<int32_t> sym_off;    // concatenated per-enum prefix offsets into sym_bytes (n+1 entries per enum)
    std::vector<uint8_t> sym_bytes;
    std::vector<OutArray> arrays;
    std::vector<int> top_arrays;     // one per top-level column
    std::vector<int> validity_slots; // slots with role Validity (null counts are computed for these)
    int n_spaces = 1;
    int max_depth = 0;
};

// Throws std::runtime_error when the schema exceeds a documented limit.
Plan build_plan(const AvroNode& top, const std::vector<ArrowField>& fields);

}  // namespace rv
