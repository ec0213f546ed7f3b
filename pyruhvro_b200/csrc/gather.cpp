// See gather.hpp.
#include "gather.hpp"

#include <stdexcept>

namespace rv {

int64_t gather_meta_len(const Plan& plan) {
    return int64_t(plan.n_spaces) + int64_t(plan.streams.size()) + int64_t(plan.validity_slots.size());
}

void gather_meta_of(const Plan& plan, const ChunkOut& c, int64_t* out) {
    int64_t* o = out;
    for (int sp = 0; sp < plan.n_spaces; ++sp) *o++ = c.space_rows[size_t(sp)];
    for (size_t s = 0; s < plan.streams.size(); ++s) {
        const Stream& st = plan.streams[s];
        if (st.is_rows) *o++ = c.space_rows[size_t(st.space)];
        else *o++ = c.slot_bytes[size_t(plan.nodes[size_t(st.node)].slot_b)];
    }
    for (int sl : plan.validity_slots) *o++ = c.null_count[size_t(sl)];
}

namespace {

constexpr int64_t kI32Max = 0x7FFFFFFF;

int width_of(const Slot& s) {
    switch (s.role) {
        case SlotRole::Values32: return 4;
        case SlotRole::Values64: return 8;
        case SlotRole::ValuesW: return s.width;
        case SlotRole::TypeIds: return 1;
        default: return 0;
    }
}

}  // namespace

GatherPlan plan_gather(const Plan& plan, const int64_t* metas, int world) {
    const int n_sp = plan.n_spaces, S = int(plan.streams.size()), nv = int(plan.validity_slots.size());
    const int64_t M = gather_meta_len(plan);
    auto rows_of = [&](int r, int sp) { return metas[int64_t(r) * M + sp]; };
    auto tot_of = [&](int r, int s) { return metas[int64_t(r) * M + n_sp + s]; };
    auto nulls_of = [&](int r, int v) { return metas[int64_t(r) * M + n_sp + S + v]; };

    GatherPlan gp;
    gp.group_of_rank.assign(size_t(world), 0);
    int r = 0;
    while (r < world) {
        // ---- as many consecutive ranks as fit Arrow's i32 limits
        std::vector<int64_t> rows(size_t(n_sp), 0), tot(size_t(std::max(S, 1)), 0);
        int e = r;
        for (; e < world; ++e) {
            bool fits = true;
            for (int sp = 0; sp < n_sp && fits; ++sp) fits = rows[size_t(sp)] + rows_of(e, sp) <= kI32Max;
            for (int s = 0; s < S && fits; ++s) fits = tot[size_t(s)] + tot_of(e, s) <= kI32Max;
            if (!fits) {
                if (e == r) throw std::runtime_error("a single rank's batch is beyond Arrow's i32 offsets");
                break;
            }
            for (int sp = 0; sp < n_sp; ++sp) rows[size_t(sp)] += rows_of(e, sp);
            for (int s = 0; s < S; ++s) tot[size_t(s)] += tot_of(e, s);
        }
        GatherGroup g;
        g.first_rank = r;
        g.n_ranks = e - r;
        std::vector<unsigned long long> ctot(size_t(std::max(S, 1)), 0ull);
        for (int s = 0; s < S; ++s) ctot[size_t(s)] = static_cast<unsigned long long>(tot[size_t(s)]);
        Layout L = compute_layout(plan, rows[0], 1, ctot.data());
        g.out = std::move(L.chunks[0]);
        g.arena_bytes = L.total_bytes;
        for (int v = 0; v < nv; ++v) {
            int64_t n = 0;
            for (int q = r; q < e; ++q) n += nulls_of(q, v);
            g.out.null_count[size_t(plan.validity_slots[size_t(v)])] = n;
        }
        // ---- every member's pushes
        std::vector<int64_t> row_base(size_t(n_sp), 0), tot_base(size_t(std::max(S, 1)), 0);
        for (int q = r; q < e; ++q) {
            gp.group_of_rank[size_t(q)] = int(gp.groups.size());
            std::vector<GatherJob> jobs;
            for (int sl = 0; sl < int(plan.slots.size()); ++sl) {
                const Slot& slot = plan.slots[size_t(sl)];
                const int64_t dst = int64_t(g.out.slot_off[size_t(sl)]);
                const int64_t my_rows = rows_of(q, slot.space), base = row_base[size_t(slot.space)];
                switch (slot.role) {
                    case SlotRole::Validity: case SlotRole::Bits:
                        if (my_rows) jobs.push_back(GatherJob{GK_BITS, sl, dst, my_rows, base});
                        break;
                    case SlotRole::Offsets: {
                        const int st = plan.nodes[size_t(slot.node)].stream;
                        jobs.push_back(GatherJob{GK_OFFSETS, sl, dst + 4 * base, my_rows, tot_base[size_t(st)]});
                        break;
                    }
                    case SlotRole::Data: {
                        const int64_t bytes = tot_of(q, slot.stream);
                        if (bytes) jobs.push_back(GatherJob{GK_RAW, sl, dst + tot_base[size_t(slot.stream)], bytes, 0});
                        break;
                    }
                    default: {
                        const int w = width_of(slot);
                        if (my_rows && w) jobs.push_back(GatherJob{GK_RAW, sl, dst + base * w, my_rows * w, 0});
                        break;
                    }
                }
            }
            g.jobs.push_back(std::move(jobs));
            for (int sp = 0; sp < n_sp; ++sp) row_base[size_t(sp)] += rows_of(q, sp);
            for (int s = 0; s < S; ++s) tot_base[size_t(s)] += tot_of(q, s);
        }
        gp.groups.push_back(std::move(g));
        r = e;
    }
    return gp;
}

}  // namespace rv
