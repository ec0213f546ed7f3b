#include "result.hpp"

#include <algorithm>

namespace rv {
namespace {

size_t pad64(size_t v) { return (v + 63) & ~size_t(63); }

struct ExportNode {
    std::shared_ptr<void> keep;
    std::vector<const void*> buffers;
    std::vector<ArrowArray> child_storage;
    std::vector<ArrowArray*> child_ptrs;
};

void release_array(ArrowArray* a) {
    if (!a || !a->release) return;
    for (int64_t i = 0; i < a->n_children; ++i)
        if (a->children[i] && a->children[i]->release) a->children[i]->release(a->children[i]);
    delete static_cast<ExportNode*>(a->private_data);
    a->release = nullptr;
}

struct Exporter {
    const Plan& plan;
    const ChunkOut& c;
    const uint8_t* base;
    const std::shared_ptr<void>& keep;

    const void* buf(int slot) const { return slot < 0 ? nullptr : base + c.slot_off[size_t(slot)]; }

    void finish(ExportNode* p, ArrowArray* out, int64_t len, int64_t nulls, size_t n_children) const {
        out->length = len;
        out->null_count = nulls;
        out->offset = 0;
        out->n_buffers = int64_t(p->buffers.size());
        out->n_children = int64_t(n_children);
        out->buffers = p->buffers.empty() ? nullptr : p->buffers.data();
        out->children = p->child_ptrs.empty() ? nullptr : p->child_ptrs.data();
        out->dictionary = nullptr;
        out->release = release_array;
        out->private_data = p;
    }

    void fill(int ai, ArrowArray* out) const {
        const OutArray& a = plan.arrays[size_t(ai)];
        auto* p = new ExportNode();
        p->keep = keep;
        const int64_t len = c.space_rows[size_t(a.space)];
        int64_t nulls = 0;
        const void* validity = nullptr;
        if (a.slot_v >= 0) {
            nulls = c.null_count[size_t(a.slot_v)];
            // explicit BooleanBufferBuilder (nullable record/list/map) -> always present;
            // NullBufferBuilder (primitive/string/bool builders) -> only once a null was appended
            if (a.always_validity || nulls > 0) validity = buf(a.slot_v);
        }
        switch (a.type) {
            case AT::Null: nulls = len; break;  // NullArray::new(len), no buffers (:558)
            case AT::SparseUnion: p->buffers = {buf(a.slot_a)}; nulls = 0; break;
            case AT::Struct: p->buffers = {validity}; break;
            case AT::List: case AT::Map: p->buffers = {validity, buf(a.slot_a)}; break;
            case AT::Utf8: case AT::Binary: p->buffers = {validity, buf(a.slot_a), buf(a.slot_b)}; break;
            default: p->buffers = {validity, buf(a.slot_a)}; break;
        }
        p->child_storage.resize(a.children.size());
        p->child_ptrs.resize(a.children.size());
        for (size_t i = 0; i < a.children.size(); ++i) {
            fill(a.children[i], &p->child_storage[i]);
            p->child_ptrs[i] = &p->child_storage[i];
        }
        finish(p, out, len, nulls, a.children.size());
    }
};

}  // namespace

int64_t clamp_chunks(int64_t num_chunks, int64_t n) {
    return std::min<int64_t>(std::max<int64_t>(num_chunks, 1), std::max<int64_t>(n, 1));
}

Layout compute_layout(const Plan& plan, int64_t n, int k, const unsigned long long* chunk_tot) {
    Layout L;
    const int S = std::max<int>(int(plan.streams.size()), 1);
    const int n_slots = int(plan.slots.size());
    const int64_t chunk_rows = n / k;  // build_slices (deserialize.rs:57-68): the last chunk takes the remainder
    L.chunks.resize(size_t(k));
    size_t total = 0;
    for (int pass = 0; pass < 2; ++pass) {  // pass 0: zero-initialised bit buffers; pass 1: everything else
        for (int j = 0; j < k; ++j) {
            ChunkOut& c = L.chunks[size_t(j)];
            if (pass == 0) {
                c.rows = (j == k - 1) ? n - chunk_rows * (k - 1) : chunk_rows;
                c.space_rows.assign(size_t(plan.n_spaces), 0);
                c.space_rows[0] = c.rows;
                for (int sp = 1; sp < plan.n_spaces; ++sp)
                    c.space_rows[size_t(sp)] = int64_t(chunk_tot[size_t(j) * size_t(S) + size_t(plan.space_stream[size_t(sp)])]);
                c.slot_off.assign(size_t(n_slots), 0);
                c.slot_bytes.assign(size_t(n_slots), 0);
                c.null_count.assign(size_t(n_slots), 0);
            }
            for (int sl = 0; sl < n_slots; ++sl) {
                const Slot& slot = plan.slots[size_t(sl)];
                if (slot.zero_init != (pass == 0)) continue;
                const int64_t rows = c.space_rows[size_t(slot.space)];
                int64_t logical = 0;
                size_t alloc = 0;
                switch (slot.role) {
                    case SlotRole::Validity: case SlotRole::Bits:
                        logical = (rows + 7) / 8; alloc = size_t((rows + 31) / 32) * 4; break;  // written as 32-bit words
                    case SlotRole::Values32: logical = rows * 4; alloc = size_t(logical); break;
                    case SlotRole::Values64: logical = rows * 8; alloc = size_t(logical); break;
                    case SlotRole::ValuesW: logical = rows * slot.width; alloc = size_t(logical); break;
                    case SlotRole::Offsets: logical = (rows + 1) * 4; alloc = size_t(logical); break;
                    case SlotRole::TypeIds: logical = rows; alloc = size_t(logical); break;
                    case SlotRole::Data:
                        logical = int64_t(chunk_tot[size_t(j) * size_t(S) + size_t(slot.stream)]); alloc = size_t(logical); break;
                }
                c.slot_off[size_t(sl)] = total;
                c.slot_bytes[size_t(sl)] = logical;
                total += pad64(std::max<size_t>(alloc, 1));
            }
        }
        if (pass == 0) L.zero_bytes = total;
    }
    L.total_bytes = total;
    return L;
}

void export_batch(const Plan& plan, const ChunkOut& c, const uint8_t* base, std::shared_ptr<void> keep, ArrowArray* out) {
    Exporter ex{plan, c, base, keep};
    auto* p = new ExportNode();
    p->keep = keep;
    p->buffers = {nullptr};  // RecordBatch as a non-nullable struct array
    p->child_storage.resize(plan.top_arrays.size());
    p->child_ptrs.resize(plan.top_arrays.size());
    for (size_t i = 0; i < plan.top_arrays.size(); ++i) {
        ex.fill(plan.top_arrays[i], &p->child_storage[i]);
        p->child_ptrs[i] = &p->child_storage[i];
    }
    ex.finish(p, out, c.rows, 0, plan.top_arrays.size());
}

int64_t exported_bytes(const Plan& plan, const std::vector<ChunkOut>& chunks) {
    int64_t total = 0;
    for (const ChunkOut& c : chunks) {
        for (const OutArray& a : plan.arrays) {
            if (a.slot_v >= 0 && (a.always_validity || c.null_count[size_t(a.slot_v)] > 0)) total += c.slot_bytes[size_t(a.slot_v)];
            if (a.slot_a >= 0) total += c.slot_bytes[size_t(a.slot_a)];
            if (a.slot_b >= 0) total += c.slot_bytes[size_t(a.slot_b)];
        }
    }
    return total;
}

}  // namespace rv
