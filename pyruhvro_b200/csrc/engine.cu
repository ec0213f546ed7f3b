// Host orchestration + C ABI of the decoder (include/ruhvro_b200.h).
//
// Replaces the L3 layer of the reference for the direct-decode path:
//   ruhvro/src/deserialize.rs:25-30,53-121 (dispatch, clamp_chunks, build_slices, fan-out)
// with: plan upload -> count kernel -> per-chunk scan -> exact-size Arrow arena -> emit kernel ->
// null counts -> Arrow C Data Interface export.  One CUDA stream per call; device memory comes
// from the stream-ordered pool, host output memory from a pinned-slab cache.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "../../include/ruhvro_b200.h"
#include "arrow_c.h"
#include "jit.hpp"
#include "kernels.cuh"
#include "plan.hpp"
#include "result.hpp"
#include "schema.hpp"

using namespace rv;

// ------------------------------------------------------------------------------------------
// thread-local diagnostics
// ------------------------------------------------------------------------------------------
namespace {

thread_local std::string t_error;
thread_local float t_timings[6] = {0, 0, 0, 0, 0, 0};
thread_local int t_launches = 0;
thread_local const char* t_walker = "none";
thread_local long long t_overflow_tiles = 0;

double env_double(const char* name, double dflt) {
    const char* v = std::getenv(name);
    return v && *v ? std::atof(v) : dflt;
}

rv_status fail(rv_status st, const std::string& msg) {
    t_error = msg;
    return st;
}

#define RV_CUDA(expr)                                                                                  \
    do {                                                                                               \
        cudaError_t e_ = (expr);                                                                       \
        if (e_ != cudaSuccess)                                                                         \
            return fail(RV_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));              \
    } while (0)

// ------------------------------------------------------------------------------------------
// pinned host slabs (cached: cudaHostAlloc of GiB-sized blocks costs hundreds of ms)
// ------------------------------------------------------------------------------------------
class PinnedCache {
  public:
    void* get(size_t bytes, size_t* actual) {
        size_t want = round(bytes);
        {
            std::lock_guard<std::mutex> g(mu_);
            auto it = free_.lower_bound(want);
            if (it != free_.end() && it->first <= want + want / 4 + (1u << 20)) {
                void* p = it->second;
                *actual = it->first;
                cached_ -= it->first;
                free_.erase(it);
                return p;
            }
        }
        void* p = nullptr;
        if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) {
            trim(0);
            if (cudaHostAlloc(&p, want, cudaHostAllocDefault) != cudaSuccess) return nullptr;
        }
        *actual = want;
        return p;
    }
    void put(void* p, size_t actual) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> g(mu_);
            if (cached_ + actual <= kMaxCached) {
                free_.emplace(actual, p);
                cached_ += actual;
                return;
            }
        }
        cudaFreeHost(p);
    }
    void trim(size_t keep) {
        std::lock_guard<std::mutex> g(mu_);
        while (cached_ > keep && !free_.empty()) {
            auto it = std::prev(free_.end());
            cudaFreeHost(it->second);
            cached_ -= it->first;
            free_.erase(it);
        }
    }

  private:
    static constexpr size_t kMaxCached = size_t(24) << 30;
    static size_t round(size_t b) {
        size_t g = b >= (size_t(64) << 20) ? (size_t(16) << 20) : (b >= (1u << 20) ? (1u << 20) : 65536);
        return ((b ? b : 1) + g - 1) / g * g;
    }
    std::mutex mu_;
    std::multimap<size_t, void*> free_;
    size_t cached_ = 0;
};

// ------------------------------------------------------------------------------------------
// device memory: size-bucketed cache over cudaMalloc.  Blocks are only returned once the work that
// used them has been synchronised (every decode path syncs its stream before its buffers die; arenas
// die after their batches are released), so reuse across streams and threads needs no stream ordering.
// ------------------------------------------------------------------------------------------
class DeviceCache {
  public:
    void* get(size_t bytes, int device, size_t* actual) {
        const size_t want = round(bytes);
        {
            std::lock_guard<std::mutex> g(mu_);
            auto& fl = free_[device];
            auto it = fl.lower_bound(want);
            if (it != fl.end() && it->first <= want + want / 4 + (size_t(1) << 20)) {
                void* p = it->second;
                *actual = it->first;
                cached_ -= it->first;
                fl.erase(it);
                return p;
            }
        }
        void* p = nullptr;
        if (cudaMalloc(&p, want) != cudaSuccess) {
            (void)cudaGetLastError();
            trim(device);
            if (cudaMalloc(&p, want) != cudaSuccess) { (void)cudaGetLastError(); return nullptr; }
        }
        *actual = want;
        return p;
    }
    void put(void* p, size_t actual, int device) {
        if (!p) return;
        {
            std::lock_guard<std::mutex> g(mu_);
            if (cached_ + actual <= kMaxCached) {
                free_[device].emplace(actual, p);
                cached_ += actual;
                return;
            }
        }
        cudaFree(p);  // the cache is full: give the block back to the driver
    }
    void trim(int device) {
        std::lock_guard<std::mutex> g(mu_);
        auto& fl = free_[device];
        for (auto& kv : fl) { cudaFree(kv.second); cached_ -= kv.first; }
        fl.clear();
    }

  private:
    static constexpr size_t kMaxCached = size_t(96) << 30;  // of the B200's 180 GB
    static size_t round(size_t b) {
        const size_t g = b >= (size_t(32) << 20) ? (size_t(4) << 20) : (b >= (size_t(1) << 20) ? (size_t(256) << 10) : 4096);
        return ((b ? b : 1) + g - 1) / g * g;
    }
    std::mutex mu_;
    std::map<int, std::multimap<size_t, void*>> free_;
    size_t cached_ = 0;
};

DeviceCache& devmem() {
    static DeviceCache* c = new DeviceCache();  // intentionally leaked
    return *c;
}

// Worker streams of the chunk pipeline (created once per device).
cudaStream_t worker_stream(int device, int w) {
    static std::mutex mu;
    static std::map<std::pair<int, int>, cudaStream_t> streams;
    std::lock_guard<std::mutex> g(mu);
    auto key = std::make_pair(device, w);
    auto it = streams.find(key);
    if (it != streams.end()) return it->second;
    cudaStream_t s = nullptr;
    if (cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking) != cudaSuccess) s = nullptr;
    streams[key] = s;
    return s;
}

std::mutex g_host_mu;
std::map<void*, size_t> g_host_sizes;  // rv_host_alloc blocks -> slab size

PinnedCache& pinned() {
    static PinnedCache* c = new PinnedCache();  // intentionally leaked: outlives static destructors
    return *c;
}

// ------------------------------------------------------------------------------------------
// per-process CUDA init
// ------------------------------------------------------------------------------------------
rv_status ensure_cuda(int* device) {
    static std::mutex mu;
    static std::vector<char> inited;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        return fail(RV_ERR_CUDA, std::string("no CUDA device available (this library has no CPU fallback): ") + cudaGetErrorString(e));
    RV_CUDA(cudaGetDevice(device));
    std::lock_guard<std::mutex> g(mu);
    if (inited.size() < size_t(n)) inited.resize(size_t(n), 0);
    if (!inited[size_t(*device)]) {
        RV_CUDA(prepare_kernels());
        cudaMemPool_t pool;
        RV_CUDA(cudaDeviceGetDefaultMemPool(&pool, *device));
        unsigned long long thr = ~0ull;  // keep freed blocks in the pool: allocation becomes a free-list pop
        RV_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        inited[size_t(*device)] = 1;
    }
    return RV_OK;
}

struct DevBuf {  // cached device allocation, returned on scope exit (after the owning stream was synchronised)
    void* p = nullptr;
    size_t actual = 0;
    int device = 0;
    cudaError_t alloc(size_t bytes, cudaStream_t) {
        if (cudaGetDevice(&device) != cudaSuccess) return cudaErrorInvalidDevice;
        p = devmem().get(bytes, device, &actual);
        return p ? cudaSuccess : cudaErrorMemoryAllocation;
    }
    ~DevBuf() { if (p) devmem().put(p, actual, device); }
};

}  // namespace

// ------------------------------------------------------------------------------------------
// schema handle
// ------------------------------------------------------------------------------------------
struct DevicePlan {
    DNode* nodes = nullptr;
    int32_t* sym_off = nullptr;
    uint8_t* sym_bytes = nullptr;
    int16_t* stream_slot = nullptr;
};

// Schema-specialised kernels (NVRTC), shared by all devices of the process.
struct JitState {
    bool tried = false;
    bool ok = false;
    std::string status = "not compiled";
    cudaLibrary_t lib = nullptr;
    cudaKernel_t count = nullptr, emit = nullptr;
};

struct rv_schema {
    std::atomic<int> refs{1};
    std::unique_ptr<AvroNode> avro;
    bool supported = false;
    std::string why;              // why it is unsupported / why no plan
    std::vector<ArrowField> fields;
    bool has_fields = false;
    Plan plan;
    bool has_plan = false;
    std::mutex mu;
    std::map<int, DevicePlan> dev;  // device id -> uploaded plan
    JitState jit;
};

namespace {

rv_status device_plan(rv_schema* s, int device, DevicePlan* out) {
    std::lock_guard<std::mutex> g(s->mu);
    auto it = s->dev.find(device);
    if (it != s->dev.end()) { *out = it->second; return RV_OK; }
    DevicePlan d;
    const Plan& p = s->plan;
    RV_CUDA(cudaMalloc(&d.nodes, std::max<size_t>(1, p.nodes.size()) * sizeof(DNode)));
    RV_CUDA(cudaMalloc(&d.sym_off, std::max<size_t>(1, p.sym_off.size()) * 4));
    RV_CUDA(cudaMalloc(&d.sym_bytes, std::max<size_t>(1, p.sym_bytes.size())));
    std::vector<int16_t> sslot(std::max<size_t>(1, p.streams.size()), int16_t(-1));
    for (size_t i = 0; i < p.streams.size(); ++i)
        if (!p.streams[i].is_rows) sslot[i] = p.nodes[size_t(p.streams[i].node)].slot_b;
    RV_CUDA(cudaMalloc(&d.stream_slot, sslot.size() * 2));
    RV_CUDA(cudaMemcpy(d.stream_slot, sslot.data(), sslot.size() * 2, cudaMemcpyHostToDevice));
    RV_CUDA(cudaMemcpy(d.nodes, p.nodes.data(), p.nodes.size() * sizeof(DNode), cudaMemcpyHostToDevice));
    if (!p.sym_off.empty()) RV_CUDA(cudaMemcpy(d.sym_off, p.sym_off.data(), p.sym_off.size() * 4, cudaMemcpyHostToDevice));
    if (!p.sym_bytes.empty()) RV_CUDA(cudaMemcpy(d.sym_bytes, p.sym_bytes.data(), p.sym_bytes.size(), cudaMemcpyHostToDevice));
    s->dev[device] = d;
    *out = d;
    return RV_OK;
}

constexpr int kOverflowGrid = 592;  // 4 CTAs per SM striding over the (normally empty) overflow list

std::atomic<int> g_jit_override{-1};  // -1: follow RV_JIT; 0/1: rv_set_jit_enabled()

bool jit_enabled() {
    const int o = g_jit_override.load(std::memory_order_relaxed);
    if (o >= 0) return o != 0;
    const char* e = std::getenv("RV_JIT");
    return !(e && e[0] == '0');
}

std::string device_arch(int device) {
    if (const char* e = std::getenv("RV_JIT_ARCH")) return e;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return "sm_100a";
    return "sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ((prop.major >= 9) ? "a" : "");
}

// Compiles + loads the schema-specialised kernels once per schema.  On any failure the generic
// interpreter kernels (also on the GPU) are used and the reason is kept in jit.status.
const JitState& ensure_jit(rv_schema* s, int device) {
    static const JitState disabled = [] { JitState d; d.tried = true; d.status = "disabled (RV_JIT=0 / rv_set_jit_enabled(0))"; return d; }();
    if (!jit_enabled()) return disabled;
    std::lock_guard<std::mutex> g(s->mu);
    JitState& j = s->jit;
    if (j.tried) return j;
    j.tried = true;
    std::vector<char> cubin;
    std::string log;
    if (!jit_cubin(generate_kernel_source(s->plan), device_arch(device), &cubin, &log)) { j.status = "NVRTC: " + log; return j; }
    cudaError_t e = cudaLibraryLoadData(&j.lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0);
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&j.count, j.lib, "rvj_count");
    if (e == cudaSuccess) e = cudaLibraryGetKernel(&j.emit, j.lib, "rvj_emit");
    if (e == cudaSuccess) e = cudaFuncSetAttribute(reinterpret_cast<const void*>(j.count), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(reinterpret_cast<const void*>(j.emit), cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    // occupancy is bounded by shared memory: ask for the largest carveout instead of the driver's guess
    if (e == cudaSuccess) e = cudaFuncSetAttribute(reinterpret_cast<const void*>(j.count), cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(reinterpret_cast<const void*>(j.emit), cudaFuncAttributePreferredSharedMemoryCarveout, 100);
    if (e != cudaSuccess) { j.status = std::string("loading the compiled walker failed: ") + cudaGetErrorString(e); (void)cudaGetLastError(); return j; }
    j.ok = true;
    j.status = "ok";
    return j;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// result
// ------------------------------------------------------------------------------------------
namespace {

struct Arena {  // one decode call's output memory; shared by the result and every exported batch
    void* dev = nullptr;
    size_t dev_actual = 0;
    size_t bytes = 0;
    void* host = nullptr;
    size_t host_actual = 0;
    int device = 0;
    ~Arena() {
        if (dev) devmem().put(dev, dev_actual, device);
        if (host) pinned().put(host, host_actual);
    }
};

}  // namespace

struct rv_result {
    rv_schema* schema = nullptr;
    std::vector<ChunkOut> chunks;
    std::vector<std::shared_ptr<Arena>> arenas;  // arenas[i] backs chunks[i] (one arena may back many chunks)
    int64_t arrow_bytes = 0;
    ~rv_result() { if (schema) rv_schema_release(schema); }
};

namespace {

const char* err_text(uint32_t code) {
    switch (code) {
        case E_EOF: return "unexpected end of buffer";
        case E_VARINT: return "zigzag varint too long";
        case E_BOOL: return "invalid boolean byte";
        case E_NEG_LEN: return "negative string length";
        case E_BRANCH: return "invalid union branch index";
        case E_ENUM: return "enum index out of range";
        case E_OVERFLOW: return "Arrow i32 offset overflow (or malformed input offsets)";
        default: return "decode error";
    }
}

// Small pinned host block per calling thread: the template the control words are initialised from and the
// landing zone of their read-back.  (Copies to pageable memory block the host once per copy; a decode call used
// to make eight of them.)
struct HostScratch {
    uint8_t* p = nullptr;
    size_t cap = 0;
    ~HostScratch() { if (p) { cudaFreeHost(p); (void)cudaGetLastError(); } }  // worker threads of the chunk pipeline end with the call
    uint8_t* get(size_t bytes) {
        if (bytes > cap) {
            if (p) cudaFreeHost(p);
            p = nullptr;
            cap = 0;
            const size_t want = std::max<size_t>(bytes * 2, 1 << 16);
            if (cudaHostAlloc(reinterpret_cast<void**>(&p), want, cudaHostAllocDefault) != cudaSuccess) { (void)cudaGetLastError(); p = nullptr; return nullptr; }
            cap = want;
        }
        return p;
    }
};
thread_local HostScratch t_scratch;

struct EventPool {  // cudaEventCreate/Destroy per call is measurable at small batch sizes
    cudaEvent_t ev[8] = {};
    int device = -1;
    ~EventPool() { for (auto& e : ev) if (e) cudaEventDestroy(e); (void)cudaGetLastError(); }
    cudaError_t get(int dev, cudaEvent_t** out) {
        if (device != dev) {
            for (auto& e : ev) { if (e) cudaEventDestroy(e); e = nullptr; }
            for (auto& e : ev) { const cudaError_t r = cudaEventCreate(&e); if (r != cudaSuccess) return r; }
            device = dev;
        }
        *out = ev;
        return cudaSuccess;
    }
};
thread_local EventPool t_events;

// device control block of one decode call, in 64-bit words
enum CtrlWord : int { CW_ERR = 0, CW_MAX_SPAN = 1, CW_MAX_UTF8 = 2, CW_OVERFLOW = 3, CW_OFF_FIRST = 4, CW_OFF_LAST = 5, CW_CHUNK_TOT = 8 };

// ---- the decode call ------------------------------------------------------------------------
rv_status decode_on_device(rv_schema* s, const uint8_t* d_data, const int64_t* d_offsets, int64_t n, int64_t num_chunks,
                           int64_t total_bytes_hint, cudaStream_t stream, int device, rv_result** out, int64_t record_base = 0) {
    const Plan& plan = s->plan;
    const int S = int(plan.streams.size());
    const int n_slots = int(plan.slots.size());
    // clamp_chunks (deserialize.rs:53-55)
    int64_t k64 = clamp_chunks(num_chunks, n);
    if (k64 > (int64_t(1) << 24)) return fail(RV_ERR_INVALID, "num_chunks above 2^24 is not supported");
    const int k = int(k64);
    const int64_t chunk_rows = n / k;  // build_slices (:57-68)
    const int64_t last_rows = n - chunk_rows * (k - 1);

    auto res = std::make_unique<rv_result>();
    res->schema = rv_schema_retain(s);
    auto arena_sp = std::make_shared<Arena>();
    arena_sp->device = device;

    for (int i = 0; i < 4; ++i) t_timings[i] = 0;
    t_launches = 0;
    // RV_TRACE=1: host-side phase times of this call on stderr (development aid)
    static const bool trace = std::getenv("RV_TRACE") && std::getenv("RV_TRACE")[0] == '1';
    auto t_prev = std::chrono::steady_clock::now();
    std::string trace_line;
    auto mark = [&](const char* what) {
        if (!trace) return;
        const auto now = std::chrono::steady_clock::now();
        trace_line += std::string(what) + "=" + std::to_string(std::chrono::duration<double, std::micro>(now - t_prev).count()).substr(0, 6) + "us ";
        t_prev = now;
    };

    std::vector<unsigned long long> chunk_tot(size_t(k) * size_t(std::max(S, 1)), 0ull);
    DecodeParams p{};
    DevBuf tile_agg, tile_base, lane_off, d_ctrl, d_bufs, d_overflow;
    DecodeParams pi{};   // interpreter pass over the tiles the specialised kernels skipped
    size_t smem_interp = 0;
    cudaEvent_t* ev = nullptr;
    RV_CUDA(t_events.get(device, &ev));
    const size_t ctrl_words = size_t(CW_CHUNK_TOT) + chunk_tot.size();
    unsigned long long* h_ctrl = nullptr;  // pinned: [template][read-back]
    size_t smem_count = 0, smem_emit = 0, smem_room_out = 0;
    const size_t pad_count = size_t(env_double("RV_COUNT_SMEM_PAD", 0)), pad_emit = size_t(env_double("RV_EMIT_SMEM_PAD", 0));
    unsigned long long max_utf8 = 0;
    bool use_jit = false;
    cudaKernel_t jit_count = nullptr, jit_emit = nullptr;

    mark("setup");
    if (n > 0) {
        DevicePlan dp;
        rv_status st = device_plan(s, device, &dp);
        if (st) return st;
        const int64_t tpc = std::max<int64_t>(1, (chunk_rows + kBlock - 1) / kBlock);
        const int64_t tiles_last = (last_rows + kBlock - 1) / kBlock;
        const int64_t n_tiles = tpc * (k - 1) + tiles_last;
        if (n_tiles > 0x7FFFFFF0ll) return fail(RV_ERR_INVALID, "too many records for one call");

        // tiling is known: fill it in first so the sizing kernel can use it
        p.data = d_data; p.offsets = d_offsets; p.n = n; p.chunk_rows = chunk_rows; p.k = k;
        p.tiles_per_chunk = int32_t(tpc); p.n_tiles = int32_t(n_tiles);
        // control block: error word, window maxima, overflow count, input span, per-chunk stream totals —
        // initialised by ONE copy from the pinned template and read back by one copy per phase
        const size_t ones_words = size_t(k) * std::max<size_t>(plan.validity_slots.size(), 1);
        h_ctrl = reinterpret_cast<unsigned long long*>(t_scratch.get((ctrl_words * 2 + ones_words) * 8));
        if (!h_ctrl) return fail(RV_ERR_CUDA, "pinned allocation of the control block failed");
        unsigned long long* h_back = h_ctrl + ctrl_words;
        std::memset(h_ctrl, 0, ctrl_words * 8);
        h_ctrl[CW_ERR] = ~0ull;
        RV_CUDA(d_ctrl.alloc(ctrl_words * 8, stream));
        unsigned long long* ctrl = static_cast<unsigned long long*>(d_ctrl.p);
        RV_CUDA(cudaMemcpyAsync(ctrl, h_ctrl, ctrl_words * 8, cudaMemcpyHostToDevice, stream));
        launch_tile_span_max(p, ctrl, stream);  // CW_MAX_SPAN, CW_OFF_FIRST, CW_OFF_LAST
        RV_CUDA(cudaGetLastError());
        t_launches += 1;
        RV_CUDA(cudaMemcpyAsync(h_back, ctrl, size_t(CW_CHUNK_TOT) * 8, cudaMemcpyDeviceToHost, stream));
        RV_CUDA(cudaStreamSynchronize(stream));
        const unsigned long long max_span = h_back[CW_MAX_SPAN];
        int64_t total_bytes = total_bytes_hint;
        if (total_bytes < 0) total_bytes = int64_t(h_back[CW_OFF_LAST]) - int64_t(h_back[CW_OFF_FIRST]);
        mark("span_sync");
        // Walker: schema-specialised (NVRTC) when available, else the generic interpreter.
        const JitState& jit = ensure_jit(s, device);
        use_jit = jit.ok;
        jit_count = jit.count;
        jit_emit = jit.emit;
        t_walker = use_jit ? "jit" : "interp";
        const int plan_nodes = use_jit ? 0 : int(plan.nodes.size());  // the generated walker has the plan baked in
        // shared-memory budget: [plan +] cursors are fixed; then the tile's input bytes; then (emit) the
        // staging area in which the tile's Utf8 output is assembled for coalesced write-out
        const size_t fixed = smem_map(plan_nodes, S, n_slots, 0, use_jit).in;
        const size_t limit = 227 * 1024;
        if (fixed + 2048 > limit) return fail(RV_ERR_SCHEMA, "schema too wide for the shared-memory cursor table");
        const double avg = total_bytes > 0 ? double(total_bytes) / double(n) : 16.0;
        // The window is sized for the LARGEST tile (measured above), so no tile needs the interpreter overflow
        // pass; outliers beyond 1.5x the mean tile are not allowed to shrink everyone's occupancy and do take it.
        size_t want = std::min<size_t>(size_t(max_span), size_t(avg * kBlock * env_double("RV_IN_CLAMP", 1.5))) + 48;
        if (use_jit) want = std::max<size_t>(want, size_t(S) * kBlock * 4);  // the scan area overlays the window
        want = (want + 63) & ~size_t(63);
        want = std::max<size_t>(want, 2048);
        const size_t room = (limit - fixed - 64) & ~size_t(15);
        size_t cap_in = std::min(want, room);
        smem_count = smem_map(plan_nodes, S, n_slots, uint32_t(cap_in), use_jit).out;
        smem_room_out = room > cap_in ? room - cap_in : 0;
        p.stream_slot = dp.stream_slot;
        const int n_nodes_param = plan_nodes;
        p.nodes = dp.nodes; p.n_nodes = int32_t(n_nodes_param); p.n_streams = S; p.n_slots = n_slots;
        p.sym_off = dp.sym_off; p.sym_bytes = dp.sym_bytes;
        p.n_utf8 = 0;
        for (int st_ = 0; st_ < S; ++st_) p.n_utf8 += plan.streams[size_t(st_)].is_rows ? 0 : 1;
        p.smem_data_cap = uint32_t(cap_in);
        p.prefetch_dist = 0;
        if (!(std::getenv("RV_NO_PREFETCH") && std::getenv("RV_NO_PREFETCH")[0] == '1')) {
            // CTAs resident on the device ~ how far ahead the tile a finishing CTA's successor will take is
            const size_t per_cta = smem_count + 1024;
            const int ctas_per_sm = int(std::max<size_t>(1, std::min<size_t>(8, (228 * 1024) / per_cta)));
            int sms = 148;
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
            p.prefetch_dist = sms * ctas_per_sm;
        }
        RV_CUDA(tile_agg.alloc(size_t(std::max(S, 1)) * size_t(n_tiles) * 4, stream));
        RV_CUDA(tile_base.alloc(size_t(std::max(S, 1)) * size_t(n_tiles) * 4, stream));
        RV_CUDA(lane_off.alloc(size_t(std::max(S, 1)) * size_t(n_tiles) * kBlock * 4, stream));
        RV_CUDA(d_overflow.alloc((size_t(n_tiles) + 1) * 4, stream));
        p.tile_agg = static_cast<uint32_t*>(tile_agg.p);
        p.tile_base = static_cast<uint32_t*>(tile_base.p);
        p.lane_off = static_cast<uint32_t*>(lane_off.p);
        p.chunk_tot = ctrl + CW_CHUNK_TOT;
        p.err = ctrl + CW_ERR;
        p.bufs = nullptr;

        mark("allocs");
        // development knobs: extra (unused) dynamic shared memory per CTA, to measure the kernels' sensitivity to occupancy
        smem_count = std::min<size_t>(smem_count + pad_count, limit);
        RV_CUDA(cudaEventRecord(ev[0], stream));
        p.tile_list = nullptr;
        p.overflow = reinterpret_cast<int32_t*>(ctrl + CW_OVERFLOW);
        p.overflow_list = static_cast<int32_t*>(d_overflow.p);
        if (use_jit) {
            void* args[] = {&p};
            RV_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(jit_count), dim3(unsigned(p.n_tiles)), dim3(kBlock), args, smem_count, stream));
            pi = p;
            pi.n_nodes = int32_t(plan.nodes.size());
            pi.tile_list = static_cast<const int32_t*>(d_overflow.p);
            pi.smem_stage_cap = 0;
            pi.smem_data_cap = uint32_t(std::min<size_t>(cap_in, (limit - smem_map(pi.n_nodes, S, n_slots, 0, false).in - 64) & ~size_t(15)));
            smem_interp = smem_map(pi.n_nodes, S, n_slots, pi.smem_data_cap, false).out;
            launch_count(pi, kOverflowGrid, smem_interp, stream);
            t_launches += 1;
        } else {
            launch_count(p, p.n_tiles, smem_count, stream);
        }
        RV_CUDA(cudaEventRecord(ev[1], stream));
        launch_scan(p, stream);
        RV_CUDA(cudaEventRecord(ev[2], stream));
        launch_tile_utf8_max(p, ctrl + CW_MAX_UTF8, stream);
        t_launches += 1;
        RV_CUDA(cudaGetLastError());
        t_launches += S > 0 ? 2 : 1;

        RV_CUDA(cudaMemcpyAsync(h_back, ctrl, ctrl_words * 8, cudaMemcpyDeviceToHost, stream));
        mark("count_launched");
        RV_CUDA(cudaStreamSynchronize(stream));
        mark("count_sync");
        const unsigned long long err_word = h_back[CW_ERR];
        const int overflow_n = int(uint32_t(h_back[CW_OVERFLOW]));
        max_utf8 = h_back[CW_MAX_UTF8];
        std::memcpy(chunk_tot.data(), h_back + CW_CHUNK_TOT, chunk_tot.size() * 8);
        t_overflow_tiles = overflow_n;
        if (err_word != ~0ull) {
            const uint32_t code = uint32_t(err_word & 0xFF);
            return fail(rv_status(code), std::string(err_text(code)) + " (record " + std::to_string(int64_t(err_word >> 8) + record_base) + ")");
        }
    }

    // ---- exact arena layout -------------------------------------------------------------
    Layout L = compute_layout(plan, n, k, chunk_tot.data());
    res->chunks = std::move(L.chunks);
    const size_t zero_bytes = L.zero_bytes, total = L.total_bytes;
    arena_sp->bytes = total;
    arena_sp->dev = devmem().get(std::max<size_t>(total, 64), device, &arena_sp->dev_actual);
    if (!arena_sp->dev) return fail(RV_ERR_CUDA, "device allocation of the Arrow buffer arena failed (" + std::to_string(total) + " bytes)");
    uint8_t* arena = static_cast<uint8_t*>(arena_sp->dev);
    res->arenas.assign(res->chunks.size(), arena_sp);

    if (n > 0) {
        if (zero_bytes) RV_CUDA(cudaMemsetAsync(arena, 0, zero_bytes, stream));
        // pointer table
        std::vector<void*> h_bufs(size_t(k) * size_t(n_slots));
        for (int j = 0; j < k; ++j)
            for (int sl = 0; sl < n_slots; ++sl) h_bufs[size_t(j) * size_t(n_slots) + size_t(sl)] = arena + res->chunks[size_t(j)].slot_off[size_t(sl)];
        RV_CUDA(d_bufs.alloc(h_bufs.size() * sizeof(void*), stream));
        RV_CUDA(cudaMemcpyAsync(d_bufs.p, h_bufs.data(), h_bufs.size() * sizeof(void*), cudaMemcpyHostToDevice, stream));
        p.bufs = static_cast<void* const*>(d_bufs.p);

        // Utf8 staging area of the emit CTAs, sized from the now-known string totals: a tile's share of
        // every Utf8 column (+12%) plus 16 bytes of alignment slack per column.
        {
            unsigned long long utf8 = 0;
            int n_utf8 = 0;
            for (int st_ = 0; st_ < S; ++st_) {
                if (plan.streams[size_t(st_)].is_rows) continue;
                ++n_utf8;
                for (int j = 0; j < k; ++j) utf8 += chunk_tot[size_t(j) * size_t(S) + size_t(st_)];
            }
            size_t cap_out = 0;
            if (n_utf8 > 0) {
                const double per_tile = double(utf8) / double(p.n_tiles);
                // the largest tile's Utf8 bytes (measured), clamped at 1.5x the mean for outliers: a tile that
                // outgrows the staging area writes its strings straight to global memory
                cap_out = std::min<size_t>(size_t(max_utf8), size_t(per_tile * env_double("RV_OUT_CLAMP", 1.5)) + size_t(n_utf8) * 31) + 64;
                cap_out = (cap_out + 63) & ~size_t(63);
                if (cap_out > smem_room_out) cap_out = smem_room_out & ~size_t(15);
                if (cap_out < 256) cap_out = 0;  // no room at all: strings go straight to global (interpreter pass)
            }
            if (const char* ev_ = std::getenv("RV_NO_STAGE_OUT")) if (ev_[0] == '1') cap_out = 0;
            p.smem_stage_cap = uint32_t(cap_out);
            smem_emit = std::min<size_t>(smem_count - pad_count + cap_out + pad_emit, 227 * 1024);
            if (p.prefetch_dist > 0) {
                int sms = 148;
                cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
                p.prefetch_dist = sms * int(std::max<size_t>(1, std::min<size_t>(8, (228 * 1024) / (smem_emit + 1024))));
            }
        }
        mark("layout");
        RV_CUDA(cudaEventRecord(ev[3], stream));
        if (use_jit) {
            void* args[] = {&p};
            RV_CUDA(cudaLaunchKernel(reinterpret_cast<const void*>(jit_emit), dim3(unsigned(p.n_tiles)), dim3(kBlock), args, smem_emit, stream));
            pi.bufs = p.bufs;
            launch_emit(pi, kOverflowGrid, smem_interp, stream);
            t_launches += 1;
        } else {
            launch_emit(p, p.n_tiles, smem_emit, stream);
        }
        RV_CUDA(cudaEventRecord(ev[4], stream));
        RV_CUDA(cudaGetLastError());
        t_launches += 1;

        // null counts of every validity bitmap
        const int nv = int(plan.validity_slots.size());
        std::vector<long long> ones(size_t(k) * size_t(std::max(nv, 1)), 0);
        DevBuf d_jobs, d_ones;
        if (nv > 0) {
            std::vector<NullCountJob> jobs(size_t(k) * size_t(nv));
            for (int j = 0; j < k; ++j)
                for (int v = 0; v < nv; ++v) {
                    const int sl = plan.validity_slots[size_t(v)];
                    const ChunkOut& c = res->chunks[size_t(j)];
                    jobs[size_t(j) * size_t(nv) + size_t(v)] =
                        NullCountJob{reinterpret_cast<const uint32_t*>(arena + c.slot_off[size_t(sl)]), c.space_rows[size_t(plan.slots[size_t(sl)].space)]};
                }
            RV_CUDA(d_jobs.alloc(jobs.size() * sizeof(NullCountJob), stream));
            RV_CUDA(d_ones.alloc(ones.size() * 8, stream));
            RV_CUDA(cudaMemcpyAsync(d_jobs.p, jobs.data(), jobs.size() * sizeof(NullCountJob), cudaMemcpyHostToDevice, stream));
            RV_CUDA(cudaMemsetAsync(d_ones.p, 0, ones.size() * 8, stream));
            launch_null_count(static_cast<const NullCountJob*>(d_jobs.p), int(jobs.size()), static_cast<long long*>(d_ones.p), stream);
            RV_CUDA(cudaGetLastError());
            t_launches += 1;
            RV_CUDA(cudaEventRecord(ev[5], stream));
            RV_CUDA(cudaMemcpyAsync(h_ctrl + ctrl_words * 2, d_ones.p, ones.size() * 8, cudaMemcpyDeviceToHost, stream));
        } else {
            RV_CUDA(cudaEventRecord(ev[5], stream));
        }
        // the overflow counter again: the specialised emit pass may have added tiles whose strings did not fit
        if (use_jit) RV_CUDA(cudaMemcpyAsync(h_ctrl + ctrl_words, d_ctrl.p ? static_cast<unsigned long long*>(d_ctrl.p) + CW_OVERFLOW : nullptr, 8, cudaMemcpyDeviceToHost, stream));
        mark("emit_launched");
        RV_CUDA(cudaStreamSynchronize(stream));
        mark("emit_sync");
        if (use_jit) t_overflow_tiles = int(uint32_t(h_ctrl[ctrl_words]));
        if (nv > 0) std::memcpy(ones.data(), h_ctrl + ctrl_words * 2, ones.size() * 8);
        for (int j = 0; j < k; ++j)
            for (int v = 0; v < nv; ++v) {
                const int sl = plan.validity_slots[size_t(v)];
                ChunkOut& c = res->chunks[size_t(j)];
                c.null_count[size_t(sl)] = c.space_rows[size_t(plan.slots[size_t(sl)].space)] - ones[size_t(j) * size_t(nv) + size_t(v)];
            }
        cudaEventElapsedTime(&t_timings[0], ev[0], ev[1]);
        cudaEventElapsedTime(&t_timings[1], ev[1], ev[2]);
        cudaEventElapsedTime(&t_timings[2], ev[3], ev[4]);
        cudaEventElapsedTime(&t_timings[3], ev[4], ev[5]);
    } else {
        // n == 0: one empty batch; offsets buffers hold the single 0 entry
        RV_CUDA(cudaMemsetAsync(arena, 0, std::max<size_t>(total, 64), stream));
        RV_CUDA(cudaStreamSynchronize(stream));
    }
    res->arrow_bytes = exported_bytes(plan, res->chunks);
    mark("finish");
    if (trace) std::fprintf(stderr, "[rv trace] %s\n", trace_line.c_str());
    *out = res.release();
    return RV_OK;
}

rv_status check_decodable(const rv_schema* s) {
    if (!s) return fail(RV_ERR_INVALID, "null schema handle");
    if (!s->supported)
        return fail(RV_ERR_SCHEMA, "schema is outside the direct-decode subset (" + s->why + "); this library has no Value-tree CPU fallback");
    if (!s->has_plan) return fail(RV_ERR_SCHEMA, s->why);
    return RV_OK;
}

}  // namespace

// ------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------
extern "C" {

rv_status rv_schema_parse(const char* json, size_t len, rv_schema** out) {
    if (!json || !out) return fail(RV_ERR_INVALID, "null argument");
    *out = nullptr;
    try {
        auto s = std::make_unique<rv_schema>();
        s->avro = parse_avro_schema(json, len);
        s->supported = is_supported(*s->avro, &s->why);
        if (s->avro->k == AK::Record) {
            try {
                s->fields = to_arrow_fields(*s->avro);
                s->has_fields = true;
            } catch (const std::exception& e) {
                if (s->supported) { s->supported = false; s->why = e.what(); }
            }
        }
        if (s->supported && s->has_fields) {
            try {
                s->plan = build_plan(*s->avro, s->fields);
                s->has_plan = true;
            } catch (const std::exception& e) {
                s->why = e.what();
            }
        }
        *out = s.release();
        return RV_OK;
    } catch (const std::exception& e) {
        return fail(RV_ERR_SCHEMA, e.what());
    }
}

rv_schema* rv_schema_retain(rv_schema* s) {
    if (s) s->refs.fetch_add(1, std::memory_order_relaxed);
    return s;
}

void rv_schema_release(rv_schema* s) {
    if (!s) return;
    if (s->refs.fetch_sub(1, std::memory_order_acq_rel) == 1) {
        for (auto& kv : s->dev) {
            cudaFree(kv.second.nodes);
            cudaFree(kv.second.sym_off);
            cudaFree(kv.second.sym_bytes);
            cudaFree(kv.second.stream_slot);
        }
        if (s->jit.lib) cudaLibraryUnload(s->jit.lib);
        delete s;
    }
}

int rv_schema_is_supported(const rv_schema* s) { return s && s->supported && s->has_plan ? 1 : 0; }

rv_status rv_schema_export_arrow(const rv_schema* s, struct ArrowSchema* out) {
    if (!s || !out) return fail(RV_ERR_INVALID, "null argument");
    if (!s->has_fields) return fail(RV_ERR_SCHEMA, s->why.empty() ? "top-level schema is not a record" : s->why);
    try {
        export_arrow_schema(s->fields, out);
        return RV_OK;
    } catch (const std::exception& e) {
        return fail(RV_ERR_SCHEMA, e.what());
    }
}

rv_status rv_decode_device(const rv_schema* s, const uint8_t* d_data, const int64_t* d_offsets, int64_t n,
                           int64_t num_chunks, void* cuda_stream, rv_result** out) {
    if (!out) return fail(RV_ERR_INVALID, "null argument");
    *out = nullptr;
    rv_status st = check_decodable(s);
    if (st) return st;
    if (n < 0 || (n > 0 && (!d_data || !d_offsets))) return fail(RV_ERR_INVALID, "bad input pointers");
    if (reinterpret_cast<uintptr_t>(d_data) & 15u) return fail(RV_ERR_INVALID, "d_data must be 16-byte aligned");
    int device = 0;
    st = ensure_cuda(&device);
    if (st) return st;
    try {
        return decode_on_device(const_cast<rv_schema*>(s), d_data, d_offsets, n, num_chunks, -1, static_cast<cudaStream_t>(cuda_stream), device, out);
    } catch (const std::exception& e) {
        return fail(RV_ERR_INVALID, e.what());
    }
}

}  // extern "C"

namespace {

// Copies one arena to a pinned host slab on `stream` (synchronises the stream).
rv_status arena_to_host(Arena& a, cudaStream_t stream, float* ms) {
    if (a.host) return RV_OK;
    size_t actual = 0;
    void* h = pinned().get(std::max<size_t>(a.bytes, 64), &actual);
    if (!h) return fail(RV_ERR_CUDA, "pinned host allocation failed");
    cudaEvent_t e0, e1;
    RV_CUDA(cudaEventCreate(&e0));
    RV_CUDA(cudaEventCreate(&e1));
    cudaEventRecord(e0, stream);
    cudaError_t e = cudaMemcpyAsync(h, a.dev, a.bytes, cudaMemcpyDeviceToHost, stream);
    cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    float t = 0;
    cudaEventElapsedTime(&t, e0, e1);
    if (ms) *ms += t;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    if (e != cudaSuccess) {
        pinned().put(h, actual);
        return fail(RV_ERR_CUDA, std::string("device->host copy: ") + cudaGetErrorString(e));
    }
    a.host = h;
    a.host_actual = actual;
    return RV_OK;
}

// H2D of rows [r0, r1) of the caller's packed input, decode into `num_chunks` batches, D2H.
rv_status decode_host_range(rv_schema* s, const uint8_t* data, const int64_t* offsets, int64_t r0, int64_t r1, int64_t num_chunks,
                            cudaStream_t stream, int device, rv_result** out, float* h2d_ms, float* d2h_ms) {
    const int64_t n = r1 - r0;
    DevBuf d_data, d_off;
    int64_t total = 0;
    const uint8_t* base = nullptr;
    if (n > 0) {
        const int64_t b0 = offsets[r0];
        total = offsets[r1] - b0;
        if (total < 0) return fail(RV_ERR_INVALID, "offsets are not monotonic");
        RV_CUDA(d_data.alloc(size_t(total) + 64, stream));
        RV_CUDA(d_off.alloc(size_t(n + 1) * 8, stream));
        cudaEvent_t e0, e1;
        RV_CUDA(cudaEventCreate(&e0));
        RV_CUDA(cudaEventCreate(&e1));
        cudaEventRecord(e0, stream);
        // The device copy keeps the caller's absolute offsets: the base pointer is biased so that
        // base + offsets[i] addresses record i (kept 16-byte aligned by the b0 & 15 shift).
        cudaError_t e = cudaMemcpyAsync(static_cast<uint8_t*>(d_data.p) + (b0 & 15), data + b0, size_t(total), cudaMemcpyHostToDevice, stream);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_off.p, offsets + r0, size_t(n + 1) * 8, cudaMemcpyHostToDevice, stream);
        cudaEventRecord(e1, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        float t = 0;
        cudaEventElapsedTime(&t, e0, e1);
        if (h2d_ms) *h2d_ms += t;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
        if (e != cudaSuccess) return fail(RV_ERR_CUDA, std::string("host->device copy: ") + cudaGetErrorString(e));
        base = static_cast<const uint8_t*>(d_data.p) + (b0 & 15) - b0;
    }
    rv_status st;
    try {
        st = decode_on_device(s, base, static_cast<const int64_t*>(d_off.p), n, num_chunks, total, stream, device, out, r0);
    } catch (const std::exception& e) {
        return fail(RV_ERR_INVALID, e.what());
    }
    if (st) return st;
    st = arena_to_host(*(*out)->arenas[0], stream, d2h_ms);
    if (st) { rv_result_free(*out); *out = nullptr; }
    return st;
}

bool pipeline_enabled() {
    const char* e = std::getenv("RV_PIPELINE");
    return !(e && e[0] == '0');
}

}  // namespace

extern "C" {

rv_status rv_result_to_host(rv_result* r) {
    if (!r) return fail(RV_ERR_INVALID, "null result");
    float ms = 0;
    for (auto& a : r->arenas) {
        rv_status st = arena_to_host(*a, nullptr, &ms);
        if (st) return st;
    }
    t_timings[5] = ms;
    return RV_OK;
}

rv_status rv_decode_host(const rv_schema* s_, const uint8_t* data, const int64_t* offsets, int64_t n,
                         int64_t num_chunks, rv_result** out) {
    if (!out) return fail(RV_ERR_INVALID, "null argument");
    *out = nullptr;
    rv_schema* s = const_cast<rv_schema*>(s_);
    rv_status st = check_decodable(s);
    if (st) return st;
    if (n < 0 || (n > 0 && (!data || !offsets))) return fail(RV_ERR_INVALID, "bad input pointers");
    int device = 0;
    st = ensure_cuda(&device);
    if (st) return st;
    const int64_t k = clamp_chunks(num_chunks, n);
    float h2d = 0, d2h = 0;
    // pipelining pays when every chunk is big enough to amortise its own launches and copies; many small
    // chunks go through ONE launch set that handles all chunks at once
    if (k < 2 || n / k < 16384 || !pipeline_enabled()) {
        st = decode_host_range(s, data, offsets, 0, n, num_chunks, nullptr, device, out, &h2d, &d2h);
        t_timings[4] = h2d;
        t_timings[5] = d2h;
        return st;
    }
    // Chunks are independent batches (deserialize.rs:57-68,92-119): decode them on a few worker streams so
    // the H2D copy of chunk i+1 overlaps the kernels and the D2H copy of chunk i (full-duplex PCIe).  This
    // is the GPU-side analogue of the reference fanning chunks out to its thread pool.
    const int64_t chunk_rows = n / k;
    std::vector<rv_result*> parts(size_t(k), nullptr);
    std::vector<rv_status> status(size_t(k), RV_OK);
    std::vector<std::string> message{size_t(k), std::string()};
    std::atomic<int64_t> next{0};
    std::mutex acc_mu;
    float acc[6] = {0, 0, 0, 0, 0, 0};
    int acc_launches = 0;
    long long acc_overflow = 0;
    const char* walker = "none";
    auto worker = [&](int wid) {
        cudaSetDevice(device);
        cudaStream_t stream = worker_stream(device, wid);
        for (;;) {
            const int64_t i = next.fetch_add(1);
            if (i >= k) break;
            const int64_t r0 = i * chunk_rows, r1 = (i == k - 1) ? n : r0 + chunk_rows;
            float hm = 0, dm = 0;
            status[size_t(i)] = decode_host_range(s, data, offsets, r0, r1, 1, stream, device, &parts[size_t(i)], &hm, &dm);
            if (status[size_t(i)]) message[size_t(i)] = t_error;
            std::lock_guard<std::mutex> g(acc_mu);
            for (int q = 0; q < 4; ++q) acc[q] += t_timings[q];
            acc[4] += hm;
            acc[5] += dm;
            acc_launches += t_launches;
            acc_overflow += t_overflow_tiles;
            walker = t_walker;
        }
    };
    int want_workers = 4;
    if (const char* ev = std::getenv("RV_WORKERS")) want_workers = std::max(1, std::atoi(ev));
    const int n_workers = int(std::min<int64_t>(k, want_workers));
    std::vector<std::thread> threads;
    for (int w = 0; w < n_workers; ++w) threads.emplace_back(worker, w);
    for (auto& t : threads) t.join();
    for (int q = 0; q < 6; ++q) t_timings[q] = acc[q];
    t_launches = acc_launches;
    t_overflow_tiles = acc_overflow;
    t_walker = walker;
    auto res = std::make_unique<rv_result>();
    res->schema = rv_schema_retain(s);
    rv_status first = RV_OK;
    for (int64_t i = 0; i < k; ++i) {
        if (status[size_t(i)] && !first) { first = status[size_t(i)]; t_error = message[size_t(i)]; }  // first failing chunk wins (:115-119)
    }
    for (int64_t i = 0; i < k; ++i) {
        rv_result* part = parts[size_t(i)];
        if (!part) continue;
        if (!first) {
            res->chunks.push_back(std::move(part->chunks[0]));
            res->arenas.push_back(part->arenas[0]);
            res->arrow_bytes += part->arrow_bytes;
        }
        rv_result_free(part);
    }
    if (first) return first;
    *out = res.release();
    return RV_OK;
}

int64_t rv_result_num_batches(const rv_result* r) { return r ? int64_t(r->chunks.size()) : 0; }
int64_t rv_result_num_rows(const rv_result* r, int64_t batch) {
    if (!r || batch < 0 || batch >= int64_t(r->chunks.size())) return -1;
    return r->chunks[size_t(batch)].rows;
}
int64_t rv_result_arrow_bytes(const rv_result* r) { return r ? r->arrow_bytes : 0; }
int64_t rv_result_buffer_bytes(const rv_result* r) {
    if (!r) return 0;
    int64_t total = 0;
    const Arena* last = nullptr;
    for (auto& a : r->arenas) {
        if (a.get() != last) total += int64_t(a->bytes);
        last = a.get();
    }
    return total;
}

rv_status rv_result_export(rv_result* r, int64_t batch, struct ArrowArray* out_array, struct ArrowSchema* out_schema) {
    if (!r || !out_array) return fail(RV_ERR_INVALID, "null argument");
    if (batch < 0 || batch >= int64_t(r->chunks.size())) return fail(RV_ERR_INVALID, "batch index out of range");
    if (!r->arenas[size_t(batch)]->host) return fail(RV_ERR_INVALID, "result is device-resident: call rv_result_to_host() or rv_result_export_device()");
    if (out_schema) {
        rv_status st = rv_schema_export_arrow(r->schema, out_schema);
        if (st) return st;
    }
    export_batch(r->schema->plan, r->chunks[size_t(batch)], static_cast<const uint8_t*>(r->arenas[size_t(batch)]->host), r->arenas[size_t(batch)], out_array);
    return RV_OK;
}

rv_status rv_result_export_device(rv_result* r, int64_t batch, struct ArrowDeviceArray* out_array, struct ArrowSchema* out_schema) {
    if (!r || !out_array) return fail(RV_ERR_INVALID, "null argument");
    if (batch < 0 || batch >= int64_t(r->chunks.size())) return fail(RV_ERR_INVALID, "batch index out of range");
    if (out_schema) {
        rv_status st = rv_schema_export_arrow(r->schema, out_schema);
        if (st) return st;
    }
    export_batch(r->schema->plan, r->chunks[size_t(batch)], static_cast<const uint8_t*>(r->arenas[size_t(batch)]->dev), r->arenas[size_t(batch)], &out_array->array);
    out_array->device_id = r->arenas[size_t(batch)]->device;
    out_array->device_type = ARROW_DEVICE_CUDA;
    out_array->sync_event = nullptr;  // the decode call synchronised its stream before returning
    out_array->reserved[0] = out_array->reserved[1] = out_array->reserved[2] = 0;
    return RV_OK;
}

void rv_result_free(rv_result* r) { delete r; }

void* rv_host_alloc(size_t bytes) {
    size_t actual = 0;
    void* p = pinned().get(bytes, &actual);
    if (!p) {
        t_error = "cudaHostAlloc failed (no CUDA device, or out of pinnable memory)";
        return nullptr;
    }
    std::lock_guard<std::mutex> g(g_host_mu);
    g_host_sizes[p] = actual;
    return p;
}
void rv_host_free(void* p) {
    if (!p) return;
    size_t actual = 0;
    {
        std::lock_guard<std::mutex> g(g_host_mu);
        auto it = g_host_sizes.find(p);
        if (it == g_host_sizes.end()) return;
        actual = it->second;
        g_host_sizes.erase(it);
    }
    pinned().put(p, actual);  // back to the slab cache: re-pinning GiB-sized blocks costs ~100s of ms
}

int rv_last_timings(float* out_ms, int cap) {
    int n = cap < 6 ? cap : 6;
    for (int i = 0; i < n; ++i) out_ms[i] = t_timings[i];
    return n;
}
int rv_last_launch_count(void) { return t_launches; }
rv_status rv_dev_rebase_i32(int32_t* d_dst, const int32_t* d_src, int64_t n, int32_t add, void* cuda_stream) {
    if (n < 0 || (n > 0 && (!d_dst || !d_src))) return fail(RV_ERR_INVALID, "bad argument");
    launch_rebase_i32(d_dst, d_src, n, add, static_cast<cudaStream_t>(cuda_stream));
    RV_CUDA(cudaGetLastError());
    return RV_OK;
}

rv_status rv_dev_concat_bits(uint32_t* d_dst_words, int64_t dst_bit, const uint32_t* d_src_words, int64_t nbits, void* cuda_stream) {
    if (nbits < 0 || dst_bit < 0 || (nbits > 0 && (!d_dst_words || !d_src_words))) return fail(RV_ERR_INVALID, "bad argument");
    launch_concat_bits(d_dst_words, dst_bit, d_src_words, nbits, static_cast<cudaStream_t>(cuda_stream));
    RV_CUDA(cudaGetLastError());
    return RV_OK;
}

// used by encode.cu (same library, separate translation unit)
void* rv_internal_dev_get(size_t bytes, int device, size_t* actual) { return devmem().get(bytes, device, actual); }
void rv_internal_dev_put(void* p, size_t actual, int device) { devmem().put(p, actual, device); }
const void* rv_schema_avro_root(const rv_schema* s) { return s ? s->avro.get() : nullptr; }
void rv_set_last_error(const char* msg) { t_error = msg ? msg : ""; }

const char* rv_last_walker(void) { return t_walker; }
const char* rv_schema_jit_status(const rv_schema* s) {
    if (!s) return "null schema";
    std::lock_guard<std::mutex> g(const_cast<rv_schema*>(s)->mu);
    t_error = s->jit.tried ? s->jit.status : "not attempted yet";
    return t_error.c_str();
}
long long rv_last_overflow_tiles(void) { return t_overflow_tiles; }
void rv_set_jit_enabled(int enabled) { g_jit_override.store(enabled < 0 ? -1 : (enabled ? 1 : 0)); }

int64_t rv_schema_walker_source(const rv_schema* s, char* buf, size_t cap) {
    if (!s || !s->has_plan) return -1;
    const std::string src = generate_walker_source(s->plan);
    if (buf && cap) {
        const size_t n = std::min(cap - 1, src.size());
        std::memcpy(buf, src.data(), n);
        buf[n] = 0;
    }
    return int64_t(src.size());
}

rv_status rv_schema_precompile(const rv_schema* s, const char* arch) {
    rv_status st = check_decodable(s);
    if (st) return st;
    std::vector<char> cubin;
    std::string log;
    if (!jit_cubin(generate_kernel_source(s->plan), arch && *arch ? arch : "sm_100a", &cubin, &log)) return fail(RV_ERR_CUDA, "NVRTC: " + log);
    return RV_OK;
}
const char* rv_last_error(void) { return t_error.c_str(); }
const char* rv_version(void) { return "pyruhvro_b200 0.1.0 (sm_100a)"; }

}  // extern "C"
