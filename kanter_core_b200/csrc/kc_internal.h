// kc_internal.h — internal types shared by the host code and the CUDA kernels
// of libkanter_b200.so.  Nothing here is part of the ABI (include/kanter_b200.h).
#pragma once

#include <cuda_runtime.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <initializer_list>
#include <map>
#include <set>
#include <memory>
#include <mutex>
#include <string>
#include <chrono>
#include <new>
#include <stdexcept>
#include <tuple>
#include <vector>

#include "../../include/kanter_b200.h"

// ---- error plumbing --------------------------------------------------------
void kc_set_error(const char* fmt, ...);
#define KC_FAIL(code, ...)        \
    do {                          \
        kc_set_error(__VA_ARGS__); \
        return (code);            \
    } while (0)
#define KC_CUDA(expr)                                                                  \
    do {                                                                               \
        cudaError_t e__ = (expr);                                                      \
        if (e__ != cudaSuccess) {                                                      \
            kc_set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__, \
                         __LINE__, cudaGetErrorString(e__));                           \
            return KC_ERR_CUDA;                                                        \
        }                                                                              \
    } while (0)
#define KC_TRY(expr)              \
    do {                          \
        int32_t rc__ = (expr);    \
        if (rc__ != KC_OK) return rc__; \
    } while (0)

// Nothing may unwind through the C ABI (the caller is Rust): every int32_t entry point is a function-try-block
// that ends in this.  Host memory running out, or any other C++ exception, becomes KC_ERR_GENERIC.
int32_t kc_fail_exception(const char* what);
#define KC_ABI_CATCH                                                              \
    catch (const std::bad_alloc&) { return kc_fail_exception(nullptr); }          \
    catch (const std::exception& e__) { return kc_fail_exception(e__.what()); }   \
    catch (...) { return kc_fail_exception("unknown exception"); }

#include "kc_tape.h"

// ---- resize weight tables (host-built, device-resident) ---------------------
struct KcAxisTable {
    uint32_t src_len = 0, dst_len = 0, max_taps = 0;
    int filter = 0;
    // device arrays: left[dst_len], count[dst_len], weights[max_taps][dst_len] (tap-major)
    uint32_t* d_left = nullptr;
    uint32_t* d_count = nullptr;
    float* d_weights = nullptr;
    float* d_weights_eo[3] = {nullptr, nullptr, nullptr};   // block-wise copies for the long-window horizontal pass (256 / 128 / 64 outputs per block), built on first use
    // the same three arrays as ONE 2-D table of 32-bit words, [2 + max_taps][dst_len]: row 0 = left, row 1 = count,
    // rows 2.. = the weights tap by tap -- the TMA resize kernel fetches a group's slice of it with one tensor load
    uint32_t* d_vtab = nullptr;
    std::vector<uint32_t> h_left, h_count;
    std::vector<float> h_weights;  // [dst_len][max_taps] on the host
    // marching tables for long windows (downsampling), built on first use (kc_resize.cu):
    // for every SOURCE index r and ring slot s = o mod 8: the weight of r in the one output o of that
    // slot whose window holds r (NaN: none), and the output that is complete after r (-1: none)
    int march_state = 0;           // 0 not built, 1 usable, -1 this axis cannot march (falls back)
    float* d_march_w = nullptr;    // [src_len][8]
    float* d_march_w2 = nullptr;   // [src_len][8][2]: every weight twice (the TMA-fed march reads FFMA2 operand pairs)
    int32_t* d_march_o = nullptr;  // [src_len][8]
    int32_t* d_march_info = nullptr;  // [src_len][12]: the eight output ids, a flag word (completing slots | tap slots << 8), padding
};

// ---- planes -------------------------------------------------------------------
// SPILLED: pixels moved to pinned host memory by the spill queue (kc_context.cu); forcing the plane brings them back
enum KcPlaneKind { KC_PLANE_DEVICE = 0, KC_PLANE_CONST = 1, KC_PLANE_EXPR = 2, KC_PLANE_SPILLED = 3 };

struct kc_plane {
    std::atomic<int> refs{1};
    kc_context* ctx = nullptr;
    uint32_t w = 0, h = 0;
    int kind = KC_PLANE_DEVICE;
    // DEVICE
    float* dptr = nullptr;
    bool owned = true;
    // spill queue: host copy while SPILLED, recency stamp, and a pin count held while a launch is being assembled
    float* host_copy = nullptr;
    bool host_borrowed = false;   // host_copy is the caller's (pinned) memory: a plane whose upload is deferred until somebody reads it
    uint64_t last_use = 0;
    int pins = 0;
    int resident_idx = -1;        // position in ctx->resident while the plane owns device storage
    // CONST
    float value = 0.0f;
    // EXPR: a lazily evaluated  a (op) b ; operands are retained
    int op = 0;
    kc_plane* a = nullptr;
    kc_plane* b = nullptr;
    // scratch for the fusion planner
    int mark = 0;          // visit stamp of the last cone collection
    int uses_in_cone = 0;  // operand references from inside that cone
    int ktag = 0;          // stamp of the kernel plan this node is computed in
    int tmp_slot = -1;     // temporary register currently holding the value
    int out_slot = -1;     // output slot when the node is a kernel output
    int remaining = 0;     // in-kernel uses not yet emitted
    int need = 0;          // size of the in-kernel sub-expression
    bool is_out = false;
    bool computed = false;

    size_t count() const { return (size_t)w * h; }
    size_t bytes() const { return count() * sizeof(float); }
};

// Device memory of a frozen evaluation plan (kc_exec.cu, kc_live_graph replay): every allocation the evaluation makes
// gets a slot of its own, the same one on every replay, so the captured kernels' pointer arguments stay valid.
struct KcArena {
    void* base = nullptr;
    std::vector<void*> slots;
    std::vector<size_t> bytes;
    size_t next = 0;          // the slot the next allocation of the capturing evaluation gets
    int live = 0;             // buffers handed out and not yet released
    bool orphaned = false;    // the plan is gone: the memory goes when `live` reaches 0
};

// What one kernel launch reads and writes, logged while an evaluation is captured for replay: the captured graph is a chain
// (one stream), the log lets the plan replace the chain by the true dependencies so that independent branches run side by side.
struct KcSpan { const void* p; size_t n; };
struct KcFootprint { std::vector<KcSpan> reads, writes; };

struct kc_context {
    // Planes and live graphs point back at their context.  Each of them, and the caller's own
    // handle, holds one count; kc_context_destroy releases the device side at once (`closed`)
    // and the struct itself goes with the last count, so images and graphs that outlive the
    // context -- the reference's SlotImages outlive its Engine -- can still be released.
    std::atomic<int> handles{1};
    bool closed = false;
    int device = 0;
    cudaStream_t stream = nullptr;           // every kernel, and every copy that is not one of the two below
    bool own_stream = true;                  // false: the caller's stream (kc_context_create_on_stream)
    // copy engines next to the compute stream: planes built from host memory are uploaded on
    // `upload_stream`, RGBA8 results leave on `download_stream`, each tied to `stream` by events,
    // so the upload of the next evaluation overlaps the download of the previous one
    cudaStream_t upload_stream = nullptr, download_stream = nullptr;
    cudaEvent_t ev_up_wait = nullptr, ev_up_done = nullptr, ev_dl_wait = nullptr, ev_dl_done = nullptr;
    bool dl_pending = false;                 // ev_dl_done was recorded and nobody waited for it yet
    void* dl_staging = nullptr;              // device RGBA8 buffer the download stream reads; guarded by ev_dl_done
    size_t dl_staging_bytes = 0;
    int sm_count = 148;
    kc_options opts{};
    std::recursive_mutex mu;
    uint64_t kernel_launches = 0;
    uint64_t bytes_live = 0;
    // spill queue (TransientBufferQueue, src/transient_buffer.rs:250-411): above `memory_threshold` bytes of
    // live planes the least recently used ones move to pinned host memory and come back on access
    uint64_t memory_threshold = UINT64_MAX;
    uint64_t use_tick = 0, bytes_spilled = 0, n_spills = 0, n_reloads = 0;
    uint64_t planes_on_host = 0;              // spilled + deferred planes alive: evaluations must look for them
    uint64_t bytes_h2d = 0, bytes_d2h = 0;    // what crossed PCIe on behalf of the caller (kc_context_transfer_stats)
    std::vector<kc_plane*> resident;                       // owned DEVICE planes, candidates for spilling
    std::map<size_t, std::vector<void*>> host_free_lists;  // pinned host buffers kept for reuse
    // per-request accounting (reset by the live graph)
    uint64_t run_kernels = 0, run_groups = 0, run_bytes = 0;
    std::map<std::tuple<uint32_t, uint32_t, int>, std::shared_ptr<KcAxisTable>> axis_tables;
    // 1 -> len broadcasts whose single normalised tap is exactly 1.0 (kc_exec.cu, plane_resize)
    std::map<std::pair<uint32_t, int>, bool> unit_broadcast;
    std::atomic<bool> cancel{false};
    // ProcessPackManager::max_count (src/process_pack.rs:27, set_max_processing_nodes src/texture_processor.rs:111-114):
    // how many nodes one engine turn admits; num_cpus::get() by default
    size_t max_processing_nodes = 1;
    // kernels whose dynamic shared-memory limit has been raised on THIS device (the attribute is per device)
    std::set<const void*> smem_attr_done;
    // exact-size recycling of device buffers on top of the stream-ordered pool: a plane freed
    // by one evaluation is handed to the next one of the same size without touching the driver
    std::map<size_t, std::vector<void*>> free_lists;
    uint64_t bytes_cached = 0;
    // optional per-launch device timing (kc_context_set_timing)
    bool timing = false;
    struct TimedLaunch { int kind; cudaEvent_t start, stop; };
    std::vector<TimedLaunch> timed;
    std::vector<cudaEvent_t> event_pool;
    // peer halo mailboxes (kc_h2n.cu): once a context has waited on a peer's flag, every synchronising call
    // checks the device's time-out counter and fails instead of handing back a strip computed from a stale row
    // > 0 while the operand cone of a stencil is being evaluated (KcExactScope): the tape kernels then use the
    // EXACT arithmetic whatever opts.math_mode says.  HeightToNormal differentiates its input -- on a smooth
    // 4096-wide map an input error e becomes an output error of about e * W / 2 -- so the 4e-7 of FAST pow would
    // leave the north star's 1e-5 / 1e-6; bit-identical inputs keep FAST HeightToNormal inside it.
    int exact_scope = 0;
    // evaluation replay: a log of allocation sizes while a plan is being recorded, the arena that serves allocations
    // while one is being captured, and every arena that still has buffers out
    std::vector<size_t>* alloc_log = nullptr;
    KcArena* arena_active = nullptr;
    std::vector<KcArena*> arenas;
    // concurrent section (kc_context_concurrent_begin/end): replays of DIFFERENT plans go to side streams ("lanes"), so that
    // independent graphs overlap on the device; a plan keeps its lane (its replays write the same arena and stay in order)
    int lanes_open = 0;                    // lanes of the open section, 0: none
    int lane_next = 0;
    bool lanes_dirty = false;              // a lane has work the compute stream has not been made to wait for
    std::vector<cudaStream_t> lane_streams;
    std::vector<cudaEvent_t> lane_events;
    std::vector<char> lane_used;
    cudaEvent_t lane_fork = nullptr;
    std::vector<KcFootprint>* capture_log = nullptr;   // one entry per kernel launch, in launch order (only while capturing)
    bool capturing = false;                    // ctx->stream is in CUDA stream capture: nothing but kernel launches may be enqueued
    unsigned int* d_halo_timeouts = nullptr;   // device counter of waits that gave up (this context's kernels only)
    bool halo_used = false;
    uint32_t halo_timeouts_seen = 0;
};

// kernel kinds for kc_context_timing_read
enum { KC_KERNEL_TAPE = 0, KC_KERNEL_FILL, KC_KERNEL_FROM_U8, KC_KERNEL_H2N, KC_KERNEL_RESIZE_V, KC_KERNEL_RESIZE_H, KC_KERNEL_KINDS };

// brackets one kernel launch with CUDA events on the context's stream when timing is on
struct KcTimed {
    kc_context* ctx;
    cudaEvent_t stop = nullptr;
    KcTimed(kc_context* c, int kind) : ctx(c) {
        if (!ctx->timing) return;
        cudaEvent_t ev[2];
        for (int i = 0; i < 2; ++i) {
            if (!ctx->event_pool.empty()) { ev[i] = ctx->event_pool.back(); ctx->event_pool.pop_back(); }
            else cudaEventCreate(&ev[i]);
        }
        cudaEventRecord(ev[0], ctx->stream);
        stop = ev[1];
        ctx->timed.push_back({kind, ev[0], ev[1]});
    }
    ~KcTimed() {
        if (stop) cudaEventRecord(stop, ctx->stream);
    }
};

inline bool kc_tape_exact(const kc_context* ctx) { return ctx->opts.math_mode == KC_MATH_EXACT || ctx->exact_scope > 0; }
struct KcExactScope {
    kc_context* ctx;
    bool on;
    explicit KcExactScope(kc_context* c, bool enable = true) : ctx(c), on(enable) { if (on) ++ctx->exact_scope; }
    ~KcExactScope() { if (on) --ctx->exact_scope; }
    KcExactScope(const KcExactScope&) = delete;
    KcExactScope& operator=(const KcExactScope&) = delete;
};

inline void kc_log_launch(kc_context* ctx, std::initializer_list<KcSpan> reads, std::initializer_list<KcSpan> writes) {
    if (!ctx->capture_log) return;
    KcFootprint f;
    for (const KcSpan& s : reads) if (s.p && s.n) f.reads.push_back(s);
    for (const KcSpan& s : writes) if (s.p && s.n) f.writes.push_back(s);
    ctx->capture_log->push_back(std::move(f));
}

// RAII device selection + context lock
inline void kc_ctx_ref(kc_context* c) { c->handles.fetch_add(1, std::memory_order_relaxed); }
inline void kc_ctx_unref(kc_context* c) {
    if (c->handles.fetch_sub(1, std::memory_order_acq_rel) == 1) delete c;
}
struct KcGuard {   // holds a count of its own: the last plane may go while the lock is held
    kc_context* ctx;
    int prev = -1;
    explicit KcGuard(kc_context* c) : ctx(c) {
        kc_ctx_ref(ctx);
        ctx->mu.lock();
        cudaGetDevice(&prev);
        if (prev != ctx->device) cudaSetDevice(ctx->device);
    }
    ~KcGuard() {
        if (prev >= 0 && prev != ctx->device) cudaSetDevice(prev);
        ctx->mu.unlock();
        kc_ctx_unref(ctx);
    }
    KcGuard(const KcGuard&) = delete;
    KcGuard& operator=(const KcGuard&) = delete;
};
// every kc_plane is made and unmade here (the context count goes with it)
inline kc_plane* kcp_alloc(kc_context* ctx) {
    auto* p = new kc_plane();
    p->ctx = ctx;
    if (ctx) kc_ctx_ref(ctx);   // NULL: a constant descriptor that belongs to no context (kc_plane_from_value)
    return p;
}
inline void kcp_dealloc(kc_plane* p) {
    kc_context* c = p->ctx;
    delete p;
    if (c) kc_ctx_unref(c);
}

// ---- host-side profile (KC_HOST_PROFILE=1): wall time of the evaluator's stages, printed when a context is destroyed
enum { KC_HP_EVALUATE = 0, KC_HP_PROCESS_NODE, KC_HP_FORCE, KC_HP_LAUNCH_SEGMENTS, KC_HP_LAUNCH_TAPE, KC_HP_RESIZE, KC_HP_H2N, KC_HP_COUNT };
extern std::atomic<uint64_t> g_kc_hp_ns[KC_HP_COUNT], g_kc_hp_calls[KC_HP_COUNT];
extern bool g_kc_hp_on;
struct KcHostTimer {
    int k;
    std::chrono::steady_clock::time_point t0;
    explicit KcHostTimer(int kind) : k(kind) { if (g_kc_hp_on) t0 = std::chrono::steady_clock::now(); }
    ~KcHostTimer() {
        if (!g_kc_hp_on) return;
        g_kc_hp_ns[k] += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
        g_kc_hp_calls[k]++;
    }
};

// ---- tuning knobs (kc_context.cu): 0 = let the library choose.  Set from the environment
// (KC_TILE_V, KC_CTAS, KC_STAGES, KC_SRC_SOFT_CAP, KC_RESIZE_THREADS) at load time or through
// kc_debug_set_tuning; used by the sweep scripts under scripts/ ----------------------------
struct KcTuning {
    int tile_v = 0;          // float4s per thread per tile of the fused elementwise kernel (1, 2, 4)
    int ctas = 0;            // resident CTAs per SM of that kernel (1..3)
    int stages = 0;          // its pipeline depth (2..4)
    int smem_cap_kb = 0;     // shared memory per SM that kernel sizes itself for (default 227): less leaves room for a kernel of another lane
    int smem_cap_exact_kb = 0;  // the same for launches of a glibc-exact cone inside a concurrent section (default 113, like the others)
    int src_soft_cap = 0;    // merge independent outputs that only share SOURCES while the union has <= this many
    int resize_threads = 0;  // threads per CTA of the fused resize kernel (32, 64, 128)
    int jit = 0;             // per-tape specialisation of the fused kernel: 0 auto (hot, long tapes on large planes), 1 always, -1 never
    int resize_tma = 0;      // fused upsample kernel with tensor-map loads/stores: 0 auto (on), -1 off (the cp.async / STG kernel)
    int resize_g = 0;        // its output rows per group (8, 16; default 16)
    int resize_rc = 0;       // rows per accumulator chunk of its horizontal pass (4, 8, 16)
    int resize_store = 0;    // 0 auto (every warp stores its own RC x 128 tiles); -1: one tensor store per G x 256 half of the block's tile
    int resize_minb = 0;     // resident CTAs per SM it is compiled for (6, 8; groups of 8 rows only)
};
extern KcTuning g_kc_tuning;
// kc_jit.cu
uint64_t kcj_generation();   // bumped whenever a specialised kernel becomes available: captured plans older than that are re-captured
int32_t kcj_try_launch(kc_context* ctx, const KcTapeArgs& args, int ns_max, int v, int ctas, int stages, bool* launched);

// raise a kernel's dynamic shared-memory limit once per context (= per device)
inline int32_t kc_ensure_smem_attr(kc_context* ctx, const void* fn, int bytes) {
    if (ctx->smem_attr_done.count(fn)) return KC_OK;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
    if (e != cudaSuccess) { kc_set_error("cudaFuncSetAttribute failed: %s", cudaGetErrorString(e)); return KC_ERR_CUDA; }
    ctx->smem_attr_done.insert(fn);
    return KC_OK;
}

// ---- device buffers (kc_context.cu): stream-ordered, recycled by exact size ----
int32_t kc_dev_alloc(kc_context* ctx, size_t bytes, void** out);
void kc_dev_free(kc_context* ctx, void* p, size_t bytes);
void kc_dev_trim(kc_context* ctx);
void kc_arena_orphan(kc_context* ctx, KcArena* a);

// ---- NUMA placement of pinned host memory (kc_numa.cu) ---------------------------------
int kc_device_numa_node(int device);                                  // -1: unknown
int32_t kc_host_alloc_on_node(int node, size_t bytes, void** out);    // page-locked; node < 0: wherever cudaHostAlloc puts it
int32_t kc_host_release(void* p);                                     // frees either kind

// ---- PNG codec on the host (kc_png.cu) ------------------------------------------
int32_t kc_png_decode_vec(const uint8_t* data, size_t n, std::vector<uint8_t>& samples, uint32_t& w, uint32_t& h, uint32_t& ch);
int32_t kc_png_decode_file_vec(const char* path, std::vector<uint8_t>& samples, uint32_t& w, uint32_t& h, uint32_t& ch);
int32_t kc_png_write_file(const char* path, const uint8_t* px, uint32_t w, uint32_t h, int ch);

// ---- spill queue (kc_context.cu) ---------------------------------------------------
void kcp_touch(kc_plane* p);                       // mark as most recently used
void kcp_adopt_storage(kc_plane* p, float* dptr);  // p becomes an owned DEVICE plane over storage taken from kcp_take_storage
float* kcp_take_storage(kc_plane* p);              // detach p's device storage (p is about to be deleted)
int32_t kcp_reload(kc_context* ctx, kc_plane* p);  // SPILLED -> DEVICE
int32_t kcp_prefetch_leaves(kc_context* ctx, kc_plane* const* roots, size_t n);  // kc_fusion.cu: reload every host-resident leaf under roots
int32_t kc_enforce_threshold(kc_context* ctx);     // spill unpinned LRU planes until bytes_live <= memory_threshold
void kcp_retain(kc_plane* p);
void kcp_release(kc_plane* p);
// Keeps planes in HBM while a launch that reads/writes them is assembled.  A pin is also a reference:
// the launch's outputs drop their operands once it is enqueued, and a source that only the
// expression held would otherwise be gone before the pin is.
struct KcPin {
    std::vector<kc_plane*> v;
    void add(kc_plane* p) { if (p) { kcp_retain(p); ++p->pins; v.push_back(p); } }
    ~KcPin() { for (kc_plane* p : v) { --p->pins; kcp_release(p); } }
    KcPin() = default;
    KcPin(const KcPin&) = delete;
    KcPin& operator=(const KcPin&) = delete;
};

// ---- plane helpers (kc_context.cu) -------------------------------------------
int32_t kcp_new_device(kc_context* ctx, uint32_t w, uint32_t h, kc_plane** out);
kc_plane* kcp_new_const(kc_context* ctx, uint32_t w, uint32_t h, float v);
kc_plane* kcp_new_expr(kc_context* ctx, int op, kc_plane* a, kc_plane* b);  // retains a, b
void kcp_retain(kc_plane* p);
void kcp_release(kc_plane* p);
// make every plane in `roots` device-resident, fusing the lazy expressions
// feeding them into as few kernels as possible (kc_fusion.cu)
int32_t kcp_force(kc_context* ctx, kc_plane* const* roots, size_t n);
// RGBA8 export of an image, fusing the conversion into the producing kernel
int32_t kcp_export_rgba8(kc_context* ctx, const kc_image* img, int srgb, uint32_t* d_out);

inline void kci_clear(kc_image* im) {
    im->kind = KC_IMAGE_GRAY;
    im->width = im->height = 0;
    for (int c = 0; c < 4; ++c) im->planes[c] = nullptr;
}
inline int kci_nplanes(const kc_image* im) { return im->kind == KC_IMAGE_RGBA ? 4 : 1; }
void kci_retain(const kc_image* im);
void kci_release(kc_image* im);

// ---- kernel launchers (defined in the .cu files) -----------------------------
int32_t kck_launch_tape(kc_context* ctx, const KcTapeArgs& args);
int32_t kck_from_u8(kc_context* ctx, const uint8_t* d_samples, uint32_t channels, size_t n,
                    float* const planes[4]);
// h_full/halo: for a horizontal strip of a taller image, the full height and the row above
// the strip (NULL halo + h_full == h: the whole image, toroidal wrap)
// halo mailboxes in peer memory (kc_h2n.cu)
struct kc_halo_link;
// publish_to / ack_to: the fused exchange -- the kernel also publishes the strip's last row into `publish_to` and
// acknowledges the halo it read in `ack_to` (kc_height_to_normal_strip_exchange)
int32_t kck_height_to_normal(kc_context* ctx, const float* hgt, uint32_t w, uint32_t h, uint32_t h_full,
                             const float* halo, float* r, float* g, float* b,
                             const unsigned long long* peer_flag = nullptr, unsigned long long peer_step = 0,
                             const kc_halo_link* publish_to = nullptr, const kc_halo_link* ack_to = nullptr);
int32_t kck_halo_read_args(const kc_halo_link* inbox, uint64_t step, const float** halo, const unsigned long long** flag);
int32_t kck_halo_ack(kc_context* ctx, const kc_halo_link* inbox, uint64_t step);
uint32_t kck_halo_width(const kc_halo_link* l);
// after a synchronisation of ctx->stream: KC_ERR_CUDA if a wait on a peer's flag gave up since the last check
int32_t kck_halo_check_timeouts(kc_context* ctx);
// the compute stream waits for everything the lanes of a concurrent section were given (no-op when there is nothing)
int32_t kc_lanes_join(kc_context* ctx);
// the lane of a plan inside a concurrent section (made to wait for what the compute stream holds so far)
int32_t kc_lane_acquire(kc_context* ctx, int* plan_lane, cudaStream_t* out);
int32_t kck_resize_plane(kc_context* ctx, const float* src, uint32_t sw, uint32_t sh, float* dst,
                         uint32_t dw, uint32_t dh, int filter);
int32_t kck_resize_plane_rows(kc_context* ctx, const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t dw,
                              uint32_t dh, int filter, uint32_t row0, uint32_t nrows);
// up to four planes of one geometry in one launch; *done = false: not a configuration the tensor-map kernel takes
int32_t kck_resize_planes_rows_batched(kc_context* ctx, const float* const* srcs, float* const* dsts, int n, uint32_t sw, uint32_t sh,
                                       uint32_t dw, uint32_t dh, int filter, uint32_t row0, uint32_t nrows, bool* done);
// host-side weight table exactly as image 0.24.0 computes it (kc_resize.cu)
void kc_resize_axis_host(uint32_t src_len, uint32_t dst_len, int filter, std::vector<uint32_t>& left,
                         std::vector<uint32_t>& count, std::vector<float>& weights, uint32_t& max_taps);

// host evaluation of one mix op (constant folding; glibc powf == the reference's)
float kc_host_mix(int op, float l, float r);

// ---- per-node semantics shared by kc_process_node and the live graph ---------
struct KcSlotData {
    uint32_t node_id = 0, slot_id = 0;
    kc_image image{};
};
