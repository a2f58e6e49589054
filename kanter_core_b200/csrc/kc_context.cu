// kc_context.cu — context, device-resident planes and SlotImage handles.
//
// Replaces the reference's pixel-buffer layer: `Buffer`/`SlotImage`
// (src/slot_image.rs:12-19), `SlotData` (src/slot_data.rs:35-39) and the
// in-memory half of `TransientBuffer{,Container}` (src/transient_buffer.rs:28-31,
// 188-247).  Planes are f32, row-major, one allocation per channel, in HBM,
// immutable once written, reference counted.  Constant planes (`vec![v; n]`)
// stay descriptors until somebody needs their pixels.
#include <algorithm>
#include <cstdarg>
#include <cstdlib>
#include <cstring>

#include <cstdlib>
#include <string>

#include <thread>

#include "kc_internal.h"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

void kc_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

std::atomic<uint64_t> g_kc_hp_ns[KC_HP_COUNT], g_kc_hp_calls[KC_HP_COUNT];
bool g_kc_hp_on = getenv("KC_HOST_PROFILE") != nullptr;
static void host_profile_report() {
    static const char* names[KC_HP_COUNT] = {"evaluate", "process_node", "kcp_force", "launch_segments", "kck_launch_tape", "plane_resize", "img_h2n"};
    fprintf(stderr, "[kc host profile] stage: calls, total ms, us per call (stages nest: evaluate > process_node, kcp_force > launch_segments > kck_launch_tape)\n");
    for (int i = 0; i < KC_HP_COUNT; ++i) {
        const uint64_t c = g_kc_hp_calls[i].load(), ns = g_kc_hp_ns[i].load();
        if (c) fprintf(stderr, "[kc host profile] %-16s %8llu  %9.3f  %8.2f\n", names[i], (unsigned long long)c, ns / 1e6, ns / 1e3 / c);
    }
}

int32_t kc_fail_exception(const char* what) {
    if (what) kc_set_error("internal error: %s", what);
    else kc_set_error("out of host memory");
    return KC_ERR_GENERIC;
}

extern "C" const char* kc_last_error(void) { return g_err; }
extern "C" int32_t kc_abi_version(void) { return KC_ABI_VERSION; }
extern "C" void kc_free(void* p) { free(p); }

extern "C" const char* kc_error_string(int32_t code) {
    // Display for TexProError, src/error.rs:37-64
    switch (code) {
        case KC_OK: return "Ok";
        case KC_ERR_GENERIC: return "Something went wrong";
        case KC_ERR_CANCELED: return "Node processing was canceled";
        case KC_ERR_IMAGE: return "Image error";
        case KC_ERR_INVALID_BUFFER_COUNT: return "Invalid number of channels";
        case KC_ERR_INVALID_NODE_ID: return "Invalid `NodeId`";
        case KC_ERR_INVALID_NODE_TYPE: return "Invalid `NodeType`";
        case KC_ERR_INVALID_SLOT_ID: return "Invalid `SlotId`";
        case KC_ERR_INVALID_SLOT_TYPE: return "Invalid `SlotType`";
        case KC_ERR_INVALID_EDGE: return "Invalid `Edge`";
        case KC_ERR_NO_SLOT_DATA: return "Could not find a `SlotData`";
        case KC_ERR_SLOT_OCCUPIED: return "`SlotId` is already in use";
        case KC_ERR_SLOT_NOT_OCCUPIED: return "`SlotId` is not in use";
        case KC_ERR_UNABLE_TO_LOCK: return "Unable to get a lock";
        case KC_ERR_NODE_PROCESSING: return "Error during node processing";
        case KC_ERR_POISON: return "Error with poisoned lock";
        case KC_ERR_TRY_LOCK: return "Error when trying to lock";
        case KC_ERR_NODE_DIRTY: return "The node is not up to date";
        case KC_ERR_IO: return "I/O error";
        case KC_ERR_INVALID_NAME:
            return "Invalid name, can only contain lowercase letters, numbers and underscores";
        case KC_ERR_CUDA: return "CUDA error";
        case KC_ERR_INVALID_ARGUMENT: return "Invalid argument";
        default: return "Unknown error";
    }
}

extern "C" int32_t kc_host_alloc(size_t bytes, void** out) try {
    if (!out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "kc_host_alloc: out is NULL");
    // on the NUMA node of the CURRENT device (kc_numa.cu); plain cudaHostAlloc where the topology is unknown
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) { cudaGetLastError(); dev = 0; }
    return kc_host_alloc_on_node(kc_device_numa_node(dev), bytes, out);
} KC_ABI_CATCH
extern "C" int32_t kc_host_free(void* p) try {
    return kc_host_release(p);
} KC_ABI_CATCH

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
extern "C" void kc_options_default(kc_options* o) {
    if (!o) return;
    memset(o, 0, sizeof *o);
    o->math_mode = KC_MATH_EXACT;
    o->fuse = 1;
}

static int32_t context_create(int32_t device, const kc_options* opts, cudaStream_t external, bool have_external, kc_context** out) {
    if (!out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "kc_context_create: out is NULL");
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0)
        KC_FAIL(KC_ERR_CUDA, "no CUDA device available (%s); this backend has no CPU fallback",
                e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    if (device < 0 || device >= n) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "device %d out of range [0,%d)", device, n);
    cudaDeviceProp prop;
    KC_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        KC_FAIL(KC_ERR_CUDA, "device %d is sm_%d%d; this library carries sm_100a code only", device,
                prop.major, prop.minor);
    auto* ctx = new kc_context();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->max_processing_nodes = std::max(1u, std::thread::hardware_concurrency());   // num_cpus::get(), src/process_pack.rs:27
    kc_options_default(&ctx->opts);
    if (opts) ctx->opts = *opts;
    int prev = 0;
    cudaGetDevice(&prev);
    auto setup = [&]() -> int32_t {
        KC_CUDA(cudaSetDevice(device));
        if (have_external) {
            ctx->stream = external;          // the caller's stream: kernels are ordered with the caller's own work; never destroyed here
            ctx->own_stream = false;
        } else {
            KC_CUDA(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        }
        KC_CUDA(cudaStreamCreateWithFlags(&ctx->upload_stream, cudaStreamNonBlocking));
        KC_CUDA(cudaStreamCreateWithFlags(&ctx->download_stream, cudaStreamNonBlocking));
        for (cudaEvent_t* ev : {&ctx->ev_up_wait, &ctx->ev_up_done, &ctx->ev_dl_wait, &ctx->ev_dl_done})
            KC_CUDA(cudaEventCreateWithFlags(ev, cudaEventDisableTiming));
        // keep freed planes in the stream-ordered pool instead of returning them to the driver
        cudaMemPool_t pool;
        KC_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
        uint64_t thr = UINT64_MAX;
        KC_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
        return KC_OK;
    };
    const int32_t rc = setup();
    cudaSetDevice(prev);
    if (rc != KC_OK) {   // whatever was created so far goes again
        if (ctx->own_stream && ctx->stream) cudaStreamDestroy(ctx->stream);
        if (ctx->upload_stream) cudaStreamDestroy(ctx->upload_stream);
        if (ctx->download_stream) cudaStreamDestroy(ctx->download_stream);
        for (cudaEvent_t ev : {ctx->ev_up_wait, ctx->ev_up_done, ctx->ev_dl_wait, ctx->ev_dl_done})
            if (ev) cudaEventDestroy(ev);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return KC_OK;
}
extern "C" int32_t kc_context_create(int32_t device, const kc_options* opts, kc_context** out) try {
    return context_create(device, opts, nullptr, false, out);
} KC_ABI_CATCH
extern "C" int32_t kc_context_create_on_stream(int32_t device, const kc_options* opts, void* cuda_stream, kc_context** out) try {
    // every kernel of the context is enqueued on the caller's stream (NULL: the legacy default stream),
    // so the evaluation is ordered with the caller's other CUDA work without events in between
    return context_create(device, opts, (cudaStream_t)cuda_stream, true, out);
} KC_ABI_CATCH

static void axis_table_free(KcAxisTable& t) {
    if (t.d_left) cudaFree(t.d_left);
    if (t.d_count) cudaFree(t.d_count);
    if (t.d_weights) cudaFree(t.d_weights);
    for (float*& p : t.d_weights_eo) {
        if (p) cudaFree(p);
        p = nullptr;
    }
    if (t.d_march_w) cudaFree(t.d_march_w);
    if (t.d_march_o) cudaFree(t.d_march_o);
    if (t.d_march_w2) cudaFree(t.d_march_w2);
    t.d_march_w2 = nullptr;
    if (t.d_march_info) cudaFree(t.d_march_info);
    t.d_march_info = nullptr;
    if (t.d_vtab) cudaFree(t.d_vtab);
    t.d_vtab = nullptr;
    t.d_march_w = nullptr;
    t.d_march_o = nullptr;
    t.d_left = t.d_count = nullptr;
    t.d_weights = nullptr;
}

extern "C" int32_t kc_context_destroy(kc_context* ctx) try {
    if (!ctx) return KC_OK;
    if (ctx->closed) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "the context was destroyed already");
    {
        KcGuard g(ctx);
        ctx->lanes_open = 0;
        kc_lanes_join(ctx);
        cudaStreamSynchronize(ctx->upload_stream);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamSynchronize(ctx->download_stream);
        if (ctx->dl_staging) cudaFreeAsync(ctx->dl_staging, ctx->stream);
        if (ctx->d_halo_timeouts) { cudaFree(ctx->d_halo_timeouts); ctx->d_halo_timeouts = nullptr; }
        for (cudaStream_t st : ctx->lane_streams) cudaStreamDestroy(st);
        for (cudaEvent_t ev : ctx->lane_events) cudaEventDestroy(ev);
        if (ctx->lane_fork) cudaEventDestroy(ctx->lane_fork);
        ctx->lane_streams.clear();
        ctx->lane_events.clear();
        ctx->lane_used.clear();
        ctx->lane_fork = nullptr;
        kc_dev_trim(ctx);
        for (auto& kv : ctx->axis_tables) axis_table_free(*kv.second);
        ctx->axis_tables.clear();
        for (auto& t : ctx->timed) { cudaEventDestroy(t.start); cudaEventDestroy(t.stop); }
        for (cudaEvent_t e : ctx->event_pool) cudaEventDestroy(e);
        for (cudaEvent_t e : {ctx->ev_up_wait, ctx->ev_up_done, ctx->ev_dl_wait, ctx->ev_dl_done}) cudaEventDestroy(e);
        cudaStreamSynchronize(ctx->stream);
        cudaStreamDestroy(ctx->upload_stream);
        cudaStreamDestroy(ctx->download_stream);
        if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
        ctx->timed.clear();
        ctx->event_pool.clear();
        ctx->dl_staging = nullptr;
        ctx->stream = ctx->upload_stream = ctx->download_stream = nullptr;
        ctx->ev_up_wait = ctx->ev_up_done = ctx->ev_dl_wait = ctx->ev_dl_done = nullptr;
        ctx->closed = true;
    }
    if (g_kc_hp_on) host_profile_report();
    kc_ctx_unref(ctx);   // planes and live graphs still alive keep the bookkeeping until they are released
    return KC_OK;
} KC_ABI_CATCH

// ---- concurrent sections ---------------------------------------------------------------------------------------
int32_t kc_lanes_join(kc_context* ctx) {
    if (!ctx->lanes_dirty) return KC_OK;
    for (size_t i = 0; i < ctx->lane_streams.size(); ++i)
        if (ctx->lane_used[i]) {
            KC_CUDA(cudaEventRecord(ctx->lane_events[i], ctx->lane_streams[i]));
            KC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->lane_events[i], 0));
            ctx->lane_used[i] = 0;
        }
    ctx->lanes_dirty = false;
    return KC_OK;
}

int32_t kc_lane_acquire(kc_context* ctx, int* plan_lane, cudaStream_t* out) {
    const int n = ctx->lanes_open;
    while ((int)ctx->lane_streams.size() < n) {
        cudaStream_t st = nullptr;
        cudaEvent_t ev = nullptr;
        KC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
        if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
            cudaStreamDestroy(st);
            KC_FAIL(KC_ERR_CUDA, "could not create the event of a lane: %s", cudaGetErrorString(cudaGetLastError()));
        }
        ctx->lane_streams.push_back(st);
        ctx->lane_events.push_back(ev);
        ctx->lane_used.push_back(0);
    }
    if (!ctx->lane_fork) KC_CUDA(cudaEventCreateWithFlags(&ctx->lane_fork, cudaEventDisableTiming));
    if (*plan_lane < 0 || *plan_lane >= n) *plan_lane = ctx->lane_next++ % n;
    const int l = *plan_lane;
    // everything the compute stream holds so far (uploads of this plan's inputs included) comes first
    KC_CUDA(cudaEventRecord(ctx->lane_fork, ctx->stream));
    KC_CUDA(cudaStreamWaitEvent(ctx->lane_streams[l], ctx->lane_fork, 0));
    ctx->lane_used[l] = 1;
    ctx->lanes_dirty = true;
    *out = ctx->lane_streams[l];
    return KC_OK;
}

extern "C" int32_t kc_context_concurrent_begin(kc_context* ctx, int32_t lanes) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (lanes < 1 || lanes > 8) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "lanes must be 1..8, got %d", lanes);
    KcGuard g(ctx);
    KC_TRY(kc_lanes_join(ctx));
    ctx->lanes_open = lanes > 1 ? lanes : 0;
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_context_concurrent_end(kc_context* ctx) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    KcGuard g(ctx);
    ctx->lanes_open = 0;
    return kc_lanes_join(ctx);
} KC_ABI_CATCH

extern "C" int32_t kc_context_synchronize(kc_context* ctx) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    KcGuard g(ctx);
    KC_TRY(kc_lanes_join(ctx));
    // uploads are waited for by `stream` (event), downloads are not
    KC_CUDA(cudaStreamSynchronize(ctx->stream));
    KC_CUDA(cudaStreamSynchronize(ctx->download_stream));
    ctx->dl_pending = false;
    return kck_halo_check_timeouts(ctx);
} KC_ABI_CATCH
extern "C" int32_t kc_context_device(const kc_context* ctx, int32_t* device) try {
    if (!ctx || !device) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *device = ctx->device;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_stream(const kc_context* ctx, void** stream) try {
    if (!ctx || !stream) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *stream = (void*)ctx->stream;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_set_math_mode(kc_context* ctx, int32_t mode) try {
    if (!ctx || (mode != KC_MATH_EXACT && mode != KC_MATH_FAST)) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "bad math mode");
    ctx->opts.math_mode = mode;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_set_fuse(kc_context* ctx, int32_t fuse) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    ctx->opts.fuse = fuse ? 1 : 0;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_set_resize_unclamped(kc_context* ctx, int32_t unclamped) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    ctx->opts.resize_unclamped = unclamped ? 1 : 0;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_stats(const kc_context* ctx, uint64_t* kernel_launches, uint64_t* bytes_live) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (kernel_launches) *kernel_launches = ctx->kernel_launches;
    if (bytes_live) *bytes_live = ctx->bytes_live;
    return KC_OK;
} KC_ABI_CATCH

// ---------------------------------------------------------------------------
// device buffers.  All work of a context is ordered on ONE stream, so a buffer
// released by the host may be handed out again immediately: whatever still reads
// it was enqueued earlier on that stream than whatever will write it next.
// ---------------------------------------------------------------------------
int32_t kc_dev_alloc(kc_context* ctx, size_t bytes, void** out) try {
    if (ctx->closed) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "the context was destroyed");
    if (KcArena* a = ctx->arena_active) {   // an evaluation is being captured: allocation i is slot i of the plan's arena
        if (a->next >= a->slots.size() || a->bytes[a->next] != bytes)
            KC_FAIL(KC_ERR_GENERIC, "evaluation replay: allocation %zu (%zu bytes) does not match the recorded plan", a->next, bytes);
        *out = a->slots[a->next++];
        a->live++;
        return KC_OK;
    }
    if (ctx->alloc_log) ctx->alloc_log->push_back(bytes);
    auto it = ctx->free_lists.find(bytes);
    if (it != ctx->free_lists.end() && !it->second.empty()) {
        *out = it->second.back();
        it->second.pop_back();
        ctx->bytes_cached -= bytes;
        return KC_OK;
    }
    cudaError_t e = cudaMallocAsync(out, bytes, ctx->stream);
    if (e != cudaSuccess && ctx->bytes_cached) {  // give the cache back and retry once
        cudaGetLastError();
        kc_dev_trim(ctx);
        e = cudaMallocAsync(out, bytes, ctx->stream);
    }
    if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "cudaMallocAsync(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
    return KC_OK;
} KC_ABI_CATCH

static void arena_destroy(KcArena* a) {
    if (a->base) cudaFree(a->base);
    delete a;
}
void kc_arena_orphan(kc_context* ctx, KcArena* a) {   // the plan that owned it is gone
    if (!a) return;
    if (a->live == 0) {
        ctx->arenas.erase(std::remove(ctx->arenas.begin(), ctx->arenas.end(), a), ctx->arenas.end());
        if (!ctx->closed) kc_lanes_join(ctx);              // a replay on a lane of a concurrent section may still write into it
        if (!ctx->closed) cudaStreamSynchronize(ctx->stream);
        arena_destroy(a);
    } else {
        a->orphaned = true;
    }
}

void kc_dev_free(kc_context* ctx, void* p, size_t bytes) {
    if (!p) return;
    for (size_t i = 0; i < ctx->arenas.size(); ++i) {   // (empty unless a live graph replays its evaluations)
        KcArena* a = ctx->arenas[i];
        if (p < a->base || p >= (char*)a->base + ((char*)a->slots.back() - (char*)a->base) + a->bytes.back()) continue;
        if (--a->live == 0 && a->orphaned) {
            ctx->arenas.erase(ctx->arenas.begin() + i);
            if (!ctx->closed) kc_lanes_join(ctx);
            if (!ctx->closed) cudaStreamSynchronize(ctx->stream);
            arena_destroy(a);
        }
        return;                                          // arena memory never enters the free lists
    }
    if (ctx->closed) {   // a plane that outlived its context: the streams are gone, free synchronously
        cudaFree(p);
        return;
    }
    ctx->free_lists[bytes].push_back(p);
    ctx->bytes_cached += bytes;
}

void kc_dev_trim(kc_context* ctx) {
    for (auto& kv : ctx->free_lists)
        for (void* p : kv.second) cudaFreeAsync(p, ctx->stream);
    ctx->free_lists.clear();
    ctx->bytes_cached = 0;
    bool any = false;
    for (auto& kv : ctx->host_free_lists) any |= !kv.second.empty();
    if (any) {
        cudaStreamSynchronize(ctx->stream);   // copies out of / into these buffers may still be in flight
        for (auto& kv : ctx->host_free_lists)
            for (void* p : kv.second) kc_host_release(p);
        ctx->host_free_lists.clear();
    }
}

// ---------------------------------------------------------------------------
// planes
// ---------------------------------------------------------------------------
static void resident_add(kc_context* ctx, kc_plane* p);
static inline size_t plane_alloc_bytes(const kc_plane* p) {
    // round up to a whole float4 so the vector kernels' last access stays in bounds
    size_t bytes = ((p->bytes() + 15) / 16) * 16;
    return bytes ? bytes : 16;
}
int32_t kcp_new_device(kc_context* ctx, uint32_t w, uint32_t h, kc_plane** out) {
    auto* p = kcp_alloc(ctx);
    p->w = w;
    p->h = h;
    p->kind = KC_PLANE_DEVICE;
    const size_t bytes = plane_alloc_bytes(p);
    int32_t rc = kc_dev_alloc(ctx, bytes, (void**)&p->dptr);
    if (rc != KC_OK) {
        kcp_dealloc(p);
        return rc;
    }
    ctx->bytes_live += bytes;
    p->last_use = ++ctx->use_tick;
    resident_add(ctx, p);
    if (ctx->bytes_live > ctx->memory_threshold) {
        ++p->pins;                       // never the plane being handed out
        rc = kc_enforce_threshold(ctx);
        --p->pins;
        if (rc != KC_OK) { kcp_release(p); return rc; }
    }
    *out = p;
    return KC_OK;
}

// ---------------------------------------------------------------------------
// spill queue.  Everything here is ordered on the context's compute stream: the device-to-host copy
// of a spill is enqueued before the device buffer goes back to the recycler, the host-to-device
// copy of a reload before anything that reads the plane, and host buffers are recycled, never
// freed, while the context lives -- so no call in this section has to wait for the GPU.
// ---------------------------------------------------------------------------
// the list of spill candidates; each plane knows its position, so joining and leaving are O(1)
// (a context with a thousand live planes releases dozens per evaluation)
static void resident_add(kc_context* ctx, kc_plane* p) {
    p->resident_idx = (int)ctx->resident.size();
    ctx->resident.push_back(p);
}
static void resident_remove(kc_context* ctx, kc_plane* p) {
    const int i = p->resident_idx;
    if (i < 0 || (size_t)i >= ctx->resident.size() || ctx->resident[i] != p) return;
    kc_plane* last = ctx->resident.back();
    ctx->resident[i] = last;
    last->resident_idx = i;
    ctx->resident.pop_back();
    p->resident_idx = -1;
}
void kcp_touch(kc_plane* p) {
    if (p && p->ctx) p->last_use = ++p->ctx->use_tick;
}
float* kcp_take_storage(kc_plane* p) {
    float* d = p->dptr;
    resident_remove(p->ctx, p);
    p->dptr = nullptr;
    p->owned = false;
    return d;
}
void kcp_adopt_storage(kc_plane* p, float* dptr) {
    p->kind = KC_PLANE_DEVICE;
    p->dptr = dptr;
    p->owned = true;
    p->last_use = ++p->ctx->use_tick;
    resident_add(p->ctx, p);
}
static int32_t host_alloc(kc_context* ctx, size_t bytes, void** out) {
    auto it = ctx->host_free_lists.find(bytes);
    if (it != ctx->host_free_lists.end() && !it->second.empty()) {
        *out = it->second.back();
        it->second.pop_back();
        return KC_OK;
    }
    return kc_host_alloc_on_node(kc_device_numa_node(ctx->device), bytes, out);   // spilled planes stay near their GPU
}
static int32_t spill_one(kc_context* ctx, kc_plane* p) {
    const size_t bytes = plane_alloc_bytes(p);
    void* h = nullptr;
    KC_TRY(host_alloc(ctx, bytes, &h));
    cudaError_t e = cudaMemcpyAsync(h, p->dptr, p->bytes(), cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) {
        ctx->host_free_lists[bytes].push_back(h);
        KC_FAIL(KC_ERR_CUDA, "spill copy failed: %s", cudaGetErrorString(e));
    }
    kc_dev_free(ctx, p->dptr, bytes);
    resident_remove(ctx, p);
    p->dptr = nullptr;
    p->host_copy = (float*)h;
    p->kind = KC_PLANE_SPILLED;
    ctx->bytes_live -= bytes;
    ctx->bytes_spilled += bytes;
    ctx->n_spills++;
    ctx->planes_on_host++;
    return KC_OK;
}
int32_t kc_enforce_threshold(kc_context* ctx) try {
    while (ctx->bytes_live > ctx->memory_threshold) {
        kc_plane* victim = nullptr;
        for (kc_plane* q : ctx->resident)
            if (q->pins == 0 && q->owned && q->dptr && (!victim || q->last_use < victim->last_use)) victim = q;
        if (!victim) break;              // everything left is in use right now
        KC_TRY(spill_one(ctx, victim));
    }
    return KC_OK;
} KC_ABI_CATCH
int32_t kcp_reload(kc_context* ctx, kc_plane* p) {
    if (p->kind != KC_PLANE_SPILLED) return KC_OK;
    if (ctx->capturing) KC_FAIL(KC_ERR_GENERIC, "a plane would have to come back from host memory during a stream capture");
    const size_t bytes = plane_alloc_bytes(p);
    void* d = nullptr;
    KC_TRY(kc_dev_alloc(ctx, bytes, &d));
    cudaError_t e;
    if (p->host_borrowed) {
        // a deferred upload: the caller's pinned memory -> HBM on the upload stream, ordered like kc_plane_from_host
        e = cudaEventRecord(ctx->ev_up_wait, ctx->stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->upload_stream, ctx->ev_up_wait, 0);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d, p->host_copy, p->bytes(), cudaMemcpyHostToDevice, ctx->upload_stream);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_up_done, ctx->upload_stream);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_up_done, 0);
    } else {
        e = cudaMemcpyAsync(d, p->host_copy, p->bytes(), cudaMemcpyHostToDevice, ctx->stream);
    }
    if (e != cudaSuccess) {
        kc_dev_free(ctx, d, bytes);
        KC_FAIL(KC_ERR_CUDA, "reload copy failed: %s", cudaGetErrorString(e));
    }
    if (p->host_borrowed) {
        p->host_borrowed = false;
        ctx->bytes_h2d += p->bytes();
    } else {
        ctx->host_free_lists[bytes].push_back(p->host_copy);   // reusable at once: later copies into it are stream-ordered after this one
        ctx->bytes_spilled -= bytes;
        ctx->n_reloads++;
    }
    p->host_copy = nullptr;
    p->dptr = (float*)d;
    p->kind = KC_PLANE_DEVICE;
    p->last_use = ++ctx->use_tick;
    resident_add(ctx, p);
    ctx->bytes_live += bytes;
    ctx->planes_on_host--;
    if (ctx->bytes_live > ctx->memory_threshold) {
        ++p->pins;
        int32_t rc = kc_enforce_threshold(ctx);
        --p->pins;
        return rc;
    }
    return KC_OK;
}

extern "C" int32_t kc_context_set_memory_threshold(kc_context* ctx, uint64_t bytes) try {
    // TextureProcessor::memory_threshold, src/texture_processor.rs:19; 0 = no limit
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    KcGuard g(ctx);
    ctx->memory_threshold = bytes ? bytes : UINT64_MAX;
    return kc_enforce_threshold(ctx);
} KC_ABI_CATCH
extern "C" int32_t kc_context_set_max_processing_nodes(kc_context* ctx, size_t count) try {
    // TextureProcessor::set_max_processing_nodes, src/texture_processor.rs:111-114 -> ProcessPackManager::max_count
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    KcGuard g(ctx);
    ctx->max_processing_nodes = count ? count : 1;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_max_processing_nodes(const kc_context* ctx, size_t* count) try {
    if (!ctx || !count) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *count = ctx->max_processing_nodes;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_spill_stats(const kc_context* ctx, uint64_t* bytes_spilled, uint64_t* spills, uint64_t* reloads) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (bytes_spilled) *bytes_spilled = ctx->bytes_spilled;
    if (spills) *spills = ctx->n_spills;
    if (reloads) *reloads = ctx->n_reloads;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_plane_in_memory(const kc_plane* p, int32_t* in_memory) try {
    // TransientBufferContainer::in_memory: constants and lazy planes hold no pixels, so they count as resident
    if (!p || !in_memory) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *in_memory = p->kind != KC_PLANE_SPILLED;
    return KC_OK;
} KC_ABI_CATCH

kc_plane* kcp_new_const(kc_context* ctx, uint32_t w, uint32_t h, float v) {
    auto* p = kcp_alloc(ctx);
    p->w = w;
    p->h = h;
    p->kind = KC_PLANE_CONST;
    p->value = v;
    return p;
}

kc_plane* kcp_new_expr(kc_context* ctx, int op, kc_plane* a, kc_plane* b) {
    auto* p = kcp_alloc(ctx);
    // constants broadcast; the size comes from whichever side has pixels
    const kc_plane* sz = (a->kind != KC_PLANE_CONST) ? a : b;
    p->w = sz->w;
    p->h = sz->h;
    p->kind = KC_PLANE_EXPR;
    p->op = op;
    p->a = a;
    p->b = b;
    kcp_retain(a);
    kcp_retain(b);
    return p;
}

void kcp_retain(kc_plane* p) {
    if (p) p->refs.fetch_add(1, std::memory_order_relaxed);
}

void kcp_release(kc_plane* p) {
    // iterative so that long lazy chains cannot overflow the stack
    std::vector<kc_plane*> work;
    if (p) work.push_back(p);
    while (!work.empty()) {
        kc_plane* q = work.back();
        work.pop_back();
        if (q->refs.fetch_sub(1, std::memory_order_acq_rel) != 1) continue;
        if (q->kind == KC_PLANE_DEVICE && q->owned && q->dptr) {
            KcGuard g(q->ctx);
            const size_t bytes = plane_alloc_bytes(q);
            kc_dev_free(q->ctx, q->dptr, bytes);
            resident_remove(q->ctx, q);
            q->ctx->bytes_live -= bytes;
        } else if (q->kind == KC_PLANE_SPILLED && q->host_copy) {
            KcGuard g(q->ctx);
            if (!q->host_borrowed) {
                const size_t bytes = plane_alloc_bytes(q);
                if (q->ctx->closed) kc_host_release(q->host_copy);
                else q->ctx->host_free_lists[bytes].push_back(q->host_copy);
                q->ctx->bytes_spilled -= bytes;
            }
            q->ctx->planes_on_host--;
        } else if (q->kind == KC_PLANE_EXPR) {
            if (q->a) work.push_back(q->a);
            if (q->b) work.push_back(q->b);
        }
        kcp_dealloc(q);
    }
}

void kci_retain(const kc_image* im) {
    for (int c = 0; c < 4; ++c)
        if (im->planes[c]) kcp_retain(im->planes[c]);
}
void kci_release(kc_image* im) {
    for (int c = 0; c < 4; ++c) {
        if (im->planes[c]) kcp_release(im->planes[c]);
        im->planes[c] = nullptr;
    }
}

extern "C" int32_t kc_plane_create(kc_context* ctx, uint32_t w, uint32_t h, kc_plane** out) try {
    if (!ctx || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    return kcp_new_device(ctx, w, h, out);
} KC_ABI_CATCH

extern "C" int32_t kc_plane_from_value(kc_context* ctx, uint32_t w, uint32_t h, float v, kc_plane** out) try {
    // ctx may be NULL: a descriptor that is only measured (kc_plane_size, kc_calculate_size) needs no device
    if (!out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    *out = kcp_new_const(ctx, w, h, v);
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_plane_from_host(kc_context* ctx, uint32_t w, uint32_t h, const float* host, kc_plane** out) try {
    if (!ctx || !out || !host) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    kc_plane* p = nullptr;
    KC_TRY(kcp_new_device(ctx, w, h, &p));
    // The buffer may be a recycled one that kernels already enqueued on `stream` still read, and a
    // fresh one exists only in `stream` order: the copy waits for the stream's current tail, then
    // the stream waits for the copy.  What is behind the tail at this point is at most the kernels
    // of the previous evaluation -- its RGBA8 download runs on the download stream -- so this copy
    // overlaps that download (PCIe is full duplex).
    cudaError_t e = cudaEventRecord(ctx->ev_up_wait, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->upload_stream, ctx->ev_up_wait, 0);
    if (e == cudaSuccess) e = cudaMemcpyAsync(p->dptr, host, p->bytes(), cudaMemcpyHostToDevice, ctx->upload_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_up_done, ctx->upload_stream);
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->stream, ctx->ev_up_done, 0);
    if (e != cudaSuccess) {
        kcp_release(p);
        KC_FAIL(KC_ERR_CUDA, "upload failed: %s", cudaGetErrorString(e));
    }
    ctx->bytes_h2d += p->bytes();
    *out = p;
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_plane_from_host_deferred(kc_context* ctx, uint32_t w, uint32_t h, const float* host, kc_plane** out) try {
    // The pixels stay in the caller's (pinned) memory until something reads the plane; a plane nobody
    // reads -- the alpha of an image that only goes through Mix, say -- never crosses PCIe.  `host` must
    // stay valid and unchanged until the context has been synchronised after the plane's last use.
    if (!ctx || !out || !host) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    auto* p = kcp_alloc(ctx);
    p->w = w;
    p->h = h;
    p->kind = KC_PLANE_SPILLED;
    p->host_copy = const_cast<float*>(host);
    p->host_borrowed = true;
    ctx->planes_on_host++;
    *out = p;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_transfer_stats(const kc_context* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes) try {
    // bytes copied host->device (plane uploads, u8 samples) and device->host (plane / RGBA8 downloads) so far
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    if (h2d_bytes) *h2d_bytes = ctx->bytes_h2d;
    if (d2h_bytes) *d2h_bytes = ctx->bytes_d2h;
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_plane_wrap_device(kc_context* ctx, uint32_t w, uint32_t h, void* device_ptr, kc_plane** out) try {
    if (!ctx || !out || !device_ptr) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (((uintptr_t)device_ptr & 15) != 0) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "device pointer must be 16-byte aligned");
    auto* p = kcp_alloc(ctx);
    p->w = w;
    p->h = h;
    p->kind = KC_PLANE_DEVICE;
    p->dptr = (float*)device_ptr;
    p->owned = false;
    *out = p;
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_plane_retain(kc_plane* p) try {
    if (!p) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "plane is NULL");
    kcp_retain(p);
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_plane_release(kc_plane* p) try {
    if (!p) return KC_OK;
    kc_context* ctx = p->ctx;
    if (!ctx) {             // a context-less constant descriptor (kc_plane_from_value(NULL, ..)): no device state to guard
        kcp_release(p);
        return KC_OK;
    }
    KcGuard g(ctx);
    kcp_release(p);
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_plane_size(const kc_plane* p, uint32_t* w, uint32_t* h) try {
    if (!p) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "plane is NULL");
    if (w) *w = p->w;
    if (h) *h = p->h;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_plane_is_constant(const kc_plane* p, int32_t* is_const, float* value) try {
    if (!p) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "plane is NULL");
    if (is_const) *is_const = p->kind == KC_PLANE_CONST;
    if (value) *value = p->kind == KC_PLANE_CONST ? p->value : 0.0f;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_plane_device_ptr(kc_plane* p, void** device_ptr) try {
    if (!p || !device_ptr) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(p->ctx);
    KC_TRY(kcp_force(p->ctx, &p, 1));
    ++p->pins;   // a raw pointer has left the library: this plane is never spilled again
    *device_ptr = p->dptr;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_plane_upload(kc_plane* p, const float* host) try {
    if (!p || !host) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (p->kind != KC_PLANE_DEVICE && p->kind != KC_PLANE_SPILLED) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "plane has no device storage");
    KcGuard g(p->ctx);
    KC_TRY(kcp_reload(p->ctx, p));
    kcp_touch(p);
    KC_CUDA(cudaMemcpyAsync(p->dptr, host, p->bytes(), cudaMemcpyHostToDevice, p->ctx->stream));
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_plane_download(kc_plane* p, float* host) try {
    if (!p || !host) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (p->kind == KC_PLANE_CONST) {  // a descriptor: fill on the host, no device work
        size_t n = p->count();
        for (size_t i = 0; i < n; ++i) host[i] = p->value;
        return KC_OK;
    }
    KcGuard g(p->ctx);
    KC_TRY(kc_lanes_join(p->ctx));
    KC_TRY(kcp_force(p->ctx, &p, 1));
    KC_CUDA(cudaMemcpyAsync(host, p->dptr, p->bytes(), cudaMemcpyDeviceToHost, p->ctx->stream));
    KC_CUDA(cudaStreamSynchronize(p->ctx->stream));
    p->ctx->bytes_d2h += p->bytes();
    return kck_halo_check_timeouts(p->ctx);
} KC_ABI_CATCH

// ---------------------------------------------------------------------------
// images
// ---------------------------------------------------------------------------
extern "C" int32_t kc_image_from_u8(kc_context* ctx, const uint8_t* samples, uint32_t w, uint32_t h,
                                    uint32_t channels, kc_image* out) try {
    // deconstruct_image, src/shared.rs:16-56: u8/255 per sample, channels dealt
    // round-robin, absent colour planes 0.0, absent alpha 1.0.  read_slot_image
    // (:218-261) always ends up with four planes, i.e. an Rgba image.
    if (!ctx || !samples || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (channels < 1 || channels > 4) KC_FAIL(KC_ERR_INVALID_BUFFER_COUNT, "channels must be 1..4, got %u", channels);
    KcGuard g(ctx);
    if (ctx->capturing) KC_FAIL(KC_ERR_GENERIC, "an image would have to be uploaded during a stream capture");
    kci_clear(out);
    out->kind = KC_IMAGE_RGBA;
    out->width = w;
    out->height = h;
    size_t n = (size_t)w * h;
    uint8_t* d_samples = nullptr;
    const size_t staging = ((n * channels + 15) / 16) * 16 + 16;
    KC_TRY(kc_dev_alloc(ctx, staging, (void**)&d_samples));
    cudaError_t e = cudaMemcpyAsync(d_samples, samples, n * channels, cudaMemcpyHostToDevice, ctx->stream);
    ctx->bytes_h2d += n * channels;
    float* ptrs[4] = {nullptr, nullptr, nullptr, nullptr};
    KcPin pin;   // the planes allocated first stay in HBM while the later ones are allocated
    int32_t rc = e == cudaSuccess ? KC_OK : KC_ERR_CUDA;
    for (uint32_t c = 0; c < 4 && rc == KC_OK; ++c) {
        if (c < channels) {
            rc = kcp_new_device(ctx, w, h, &out->planes[c]);
            if (rc == KC_OK) { ptrs[c] = out->planes[c]->dptr; pin.add(out->planes[c]); }
        } else {
            out->planes[c] = kcp_new_const(ctx, w, h, c == 3 ? 1.0f : 0.0f);
        }
    }
    if (rc == KC_OK) rc = kck_from_u8(ctx, d_samples, channels, n, ptrs);
    kc_dev_free(ctx, d_samples, staging);
    if (rc != KC_OK) {
        kci_release(out);
        if (e != cudaSuccess) kc_set_error("upload failed: %s", cudaGetErrorString(e));
        return rc;
    }
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_image_from_host_planes(kc_context* ctx, int32_t kind, uint32_t w, uint32_t h,
                                             const float* const* planes, kc_image* out) try {
    if (!ctx || !planes || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    kci_clear(out);
    out->kind = kind == KC_IMAGE_RGBA ? KC_IMAGE_RGBA : KC_IMAGE_GRAY;
    out->width = w;
    out->height = h;
    int np = kci_nplanes(out);
    for (int c = 0; c < np; ++c) {
        int32_t rc = kc_plane_from_host(ctx, w, h, planes[c], &out->planes[c]);
        if (rc != KC_OK) {
            kci_release(out);
            return rc;
        }
    }
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_image_from_host_planes_deferred(kc_context* ctx, int32_t kind, uint32_t w, uint32_t h,
                                                      const float* const* planes, kc_image* out) try {
    if (!ctx || !planes || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    kci_clear(out);
    out->kind = kind == KC_IMAGE_RGBA ? KC_IMAGE_RGBA : KC_IMAGE_GRAY;
    out->width = w;
    out->height = h;
    int np = kci_nplanes(out);
    for (int c = 0; c < np; ++c) {
        int32_t rc = kc_plane_from_host_deferred(ctx, w, h, planes[c], &out->planes[c]);
        if (rc != KC_OK) {
            kci_release(out);
            return rc;
        }
    }
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_image_from_value(kc_context* ctx, uint32_t w, uint32_t h, float v, int32_t rgba, kc_image* out) try {
    // SlotImage::from_value, src/slot_image.rs:28-64: alpha is 1.0 whatever v is
    if (!ctx || !out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    kci_clear(out);
    out->width = w;
    out->height = h;
    if (rgba) {
        out->kind = KC_IMAGE_RGBA;
        for (int c = 0; c < 3; ++c) out->planes[c] = kcp_new_const(ctx, w, h, v);
        out->planes[3] = kcp_new_const(ctx, w, h, 1.0f);
    } else {
        out->kind = KC_IMAGE_GRAY;
        out->planes[0] = kcp_new_const(ctx, w, h, v);
    }
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_image_retain(const kc_image* img) try {
    if (!img) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "image is NULL");
    kci_retain(img);
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_image_release(kc_image* img) try {
    if (!img) return KC_OK;
    kc_context* ctx = nullptr;
    for (int c = 0; c < 4; ++c)
        if (img->planes[c]) ctx = img->planes[c]->ctx;
    if (!ctx) {             // only context-less constant descriptors (or nothing): release them all the same
        kci_release(img);
        return KC_OK;
    }
    KcGuard g(ctx);
    kci_release(img);
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_image_download(kc_context* ctx, const kc_image* in, float* const* host_planes) try {
    if (!ctx || !in || !host_planes) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    for (int c = 0; c < kci_nplanes(in); ++c)
        if (!in->planes[c] || !host_planes[c]) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "image has no plane %d (released?) or no destination for it", c);
    KcGuard g(ctx);
    KC_TRY(kc_lanes_join(ctx));
    int np = kci_nplanes(in);
    // one fused launch for whatever is still lazy, then the copies
    std::vector<kc_plane*> lazy;
    for (int c = 0; c < np; ++c)
        if (in->planes[c]->kind == KC_PLANE_EXPR) lazy.push_back(in->planes[c]);
    if (!lazy.empty()) KC_TRY(kcp_force(ctx, lazy.data(), lazy.size()));
    for (int c = 0; c < np; ++c) {
        kc_plane* p = in->planes[c];
        if (p->kind == KC_PLANE_CONST) {
            size_t n = p->count();
            for (size_t i = 0; i < n; ++i) host_planes[c][i] = p->value;
        } else {
            KC_CUDA(cudaMemcpyAsync(host_planes[c], p->dptr, p->bytes(), cudaMemcpyDeviceToHost, ctx->stream));
            ctx->bytes_d2h += p->bytes();
        }
    }
    KC_CUDA(cudaStreamSynchronize(ctx->stream));
    return kck_halo_check_timeouts(ctx);
} KC_ABI_CATCH

extern "C" int32_t kc_image_materialize(kc_context* ctx, const kc_image* in, int32_t include_constants) try {
    // make the image's planes real pixels in HBM with one fused launch (lazy
    // expression planes always; constant descriptors only on request)
    if (!ctx || !in) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    std::vector<kc_plane*> roots;
    for (int c = 0; c < kci_nplanes(in); ++c) {
        kc_plane* p = in->planes[c];
        if (!p) KC_FAIL(KC_ERR_INVALID_BUFFER_COUNT, "plane %d is NULL", c);
        const bool want = p->kind == KC_PLANE_EXPR || (p->kind == KC_PLANE_CONST && include_constants);
        if (want && std::find(roots.begin(), roots.end(), p) == roots.end()) roots.push_back(p);
    }
    return roots.empty() ? KC_OK : kcp_force(ctx, roots.data(), roots.size());
} KC_ABI_CATCH

extern "C" int32_t kc_image_to_u8_device(kc_context* ctx, const kc_image* in, int32_t srgb, void* device_rgba8) try {
    if (!ctx || !in || !device_rgba8) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    if (((uintptr_t)device_rgba8 & 15) != 0) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "device pointer must be 16-byte aligned");
    KcGuard g(ctx);
    return kcp_export_rgba8(ctx, in, srgb, (uint32_t*)device_rgba8);
} KC_ABI_CATCH

// RGBA8 export: the conversion kernel runs on `stream` into a device buffer the context keeps,
// the copy to the host runs on the download stream.  The buffer is rewritten only after the
// previous download has finished (ev_dl_done).
static int32_t to_u8_enqueue(kc_context* ctx, const kc_image* in, int32_t srgb, uint8_t* host_rgba8) {
    KC_TRY(kc_lanes_join(ctx));
    for (int c = 0; c < kci_nplanes(in); ++c)
        if (!in->planes[c]) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "image has no plane %d (released?)", c);
    const size_t n = (size_t)in->planes[0]->w * in->planes[0]->h;
    const size_t staging = ((n * 4 + 15) / 16) * 16 + 16;
    // deferred inputs start their uploads BEFORE the compute stream is made to wait for the previous
    // download: an upload waits for the compute stream's tail, and that wait must not be part of it
    {
        kc_plane* roots[4] = {in->planes[0], in->planes[1], in->planes[2], in->planes[3]};
        KC_TRY(kcp_prefetch_leaves(ctx, roots, (size_t)kci_nplanes(in)));
    }
    if (ctx->dl_pending) KC_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->ev_dl_done, 0));
    if (ctx->dl_staging_bytes < staging) {
        if (ctx->dl_staging) KC_CUDA(cudaFreeAsync(ctx->dl_staging, ctx->stream));   // ordered after the wait above
        ctx->dl_staging = nullptr;
        ctx->dl_staging_bytes = 0;
        KC_CUDA(cudaMallocAsync(&ctx->dl_staging, staging, ctx->stream));
        ctx->dl_staging_bytes = staging;
    }
    KC_TRY(kcp_export_rgba8(ctx, in, srgb, (uint32_t*)ctx->dl_staging));
    KC_CUDA(cudaEventRecord(ctx->ev_dl_wait, ctx->stream));
    KC_CUDA(cudaStreamWaitEvent(ctx->download_stream, ctx->ev_dl_wait, 0));
    cudaError_t e = cudaMemcpyAsync(host_rgba8, ctx->dl_staging, n * 4, cudaMemcpyDeviceToHost, ctx->download_stream);
    if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "download failed: %s", cudaGetErrorString(e));
    KC_CUDA(cudaEventRecord(ctx->ev_dl_done, ctx->download_stream));
    ctx->dl_pending = true;
    ctx->bytes_d2h += n * 4;
    return KC_OK;
}

extern "C" int32_t kc_image_to_u8(kc_context* ctx, const kc_image* in, int32_t srgb, uint8_t* host_rgba8) try {
    // SlotImage::to_u8 / to_u8_srgb, src/slot_image.rs:142-207
    if (!ctx || !in || !host_rgba8) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    KC_TRY(to_u8_enqueue(ctx, in, srgb, host_rgba8));
    cudaError_t e = cudaEventSynchronize(ctx->ev_dl_done);
    if (e != cudaSuccess) KC_FAIL(KC_ERR_CUDA, "download failed: %s", cudaGetErrorString(e));
    return kck_halo_check_timeouts(ctx);
} KC_ABI_CATCH

extern "C" int32_t kc_image_to_u8_async(kc_context* ctx, const kc_image* in, int32_t srgb, uint8_t* host_rgba8) try {
    if (!ctx || !in || !host_rgba8) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    return to_u8_enqueue(ctx, in, srgb, host_rgba8);
} KC_ABI_CATCH

// ---------------------------------------------------------------------------
// device-side timing on the context's stream (what bench.py uses: CUDA events
// recorded on the stream the kernels are launched on)
// ---------------------------------------------------------------------------
extern "C" int32_t kc_event_create(void** out) try {
    if (!out) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "out is NULL");
    cudaEvent_t e;
    KC_CUDA(cudaEventCreate(&e));
    *out = (void*)e;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_event_destroy(void* ev) try {
    if (ev) KC_CUDA(cudaEventDestroy((cudaEvent_t)ev));
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_event_record(kc_context* ctx, void* ev) try {
    if (!ctx || !ev) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    KC_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->stream));
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_event_record_download(kc_context* ctx, void* ev) try {
    // completes when every RGBA8 download enqueued so far (kc_image_to_u8_async) has landed
    if (!ctx || !ev) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KcGuard g(ctx);
    KC_CUDA(cudaEventRecord((cudaEvent_t)ev, ctx->download_stream));
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_event_synchronize(void* ev) try {
    if (!ev) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KC_CUDA(cudaEventSynchronize((cudaEvent_t)ev));
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_event_elapsed_ms(void* start, void* stop, float* ms) try {
    if (!start || !stop || !ms) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "NULL argument");
    KC_CUDA(cudaEventSynchronize((cudaEvent_t)stop));
    KC_CUDA(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
    return KC_OK;
} KC_ABI_CATCH

// per-launch kernel timing: every kernel this library launches on the context is
// bracketed by two events on the context's stream while timing is on
extern "C" int32_t kc_context_set_timing(kc_context* ctx, int32_t on) try {
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    KcGuard g(ctx);
    ctx->timing = on != 0;
    return KC_OK;
} KC_ABI_CATCH
extern "C" int32_t kc_context_timing_read(kc_context* ctx, int32_t kind, double* total_ms, uint64_t* launches) try {
    // waits for the stream, sums the launches of `kind` (-1: all kinds) recorded so far and forgets them
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    KcGuard g(ctx);
    KC_CUDA(cudaStreamSynchronize(ctx->stream));
    double sum = 0.0;
    uint64_t n = 0;
    std::vector<kc_context::TimedLaunch> keep;
    for (auto& t : ctx->timed) {
        if (kind >= 0 && t.kind != kind) { keep.push_back(t); continue; }
        float ms = 0.0f;
        KC_CUDA(cudaEventElapsedTime(&ms, t.start, t.stop));
        sum += ms;
        ++n;
        ctx->event_pool.push_back(t.start);
        ctx->event_pool.push_back(t.stop);
    }
    ctx->timed.swap(keep);
    if (total_ms) *total_ms = sum;
    if (launches) *launches = n;
    return KC_OK;
} KC_ABI_CATCH

static int env_int(const char* name) {
    const char* v = getenv(name);
    return v ? atoi(v) : 0;
}
static KcTuning tuning_from_env() {
    KcTuning t;
    t.tile_v = env_int("KC_TILE_V");
    t.ctas = env_int("KC_CTAS");
    t.stages = env_int("KC_STAGES");
    t.src_soft_cap = env_int("KC_SRC_SOFT_CAP");
    t.resize_threads = env_int("KC_RESIZE_THREADS");
    t.jit = env_int("KC_JIT");
    t.resize_tma = getenv("KC_RESIZE_NO_TMA") ? -1 : 0;
    t.resize_g = env_int("KC_RESIZE_G");
    t.resize_rc = env_int("KC_RESIZE_RC");
    t.resize_minb = env_int("KC_RESIZE_MINB");
    return t;
}
KcTuning g_kc_tuning = tuning_from_env();

extern "C" int32_t kc_debug_set_tuning(const char* key, int32_t value) try {
    if (!key) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "key is NULL");
    const std::string k(key);
    if (k == "tile_v") g_kc_tuning.tile_v = value;
    else if (k == "ctas") g_kc_tuning.ctas = value;
    else if (k == "stages") g_kc_tuning.stages = value;
    else if (k == "smem_cap_kb") g_kc_tuning.smem_cap_kb = value;
    else if (k == "smem_cap_exact_kb") g_kc_tuning.smem_cap_exact_kb = value;
    else if (k == "src_soft_cap") g_kc_tuning.src_soft_cap = value;
    else if (k == "resize_threads") g_kc_tuning.resize_threads = value;
    else if (k == "jit") g_kc_tuning.jit = value;
    else if (k == "resize_tma") g_kc_tuning.resize_tma = value;
    else if (k == "resize_g") g_kc_tuning.resize_g = value;
    else if (k == "resize_rc") g_kc_tuning.resize_rc = value;
    else if (k == "resize_minb") g_kc_tuning.resize_minb = value;
    else if (k == "resize_store") g_kc_tuning.resize_store = value;
    else KC_FAIL(KC_ERR_INVALID_ARGUMENT, "unknown tuning key '%s'", key);
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_context_pcie_probe(kc_context* ctx, void* pinned_host, size_t bytes, int32_t reps, int32_t direction,
                                         double* h2d_gbs, double* d2h_gbs) try {
    // Plain cudaMemcpyAsync between `pinned_host` and a scratch device buffer, `reps` copies of `bytes` each way, timed
    // with events: the PCIe ceiling the end-to-end number is judged against.  direction: 0 host->device, 1 device->host,
    // 2 both at once (upload and download streams, full duplex).
    if (!ctx || !pinned_host || bytes == 0 || reps <= 0) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "bad argument");
    KcGuard g(ctx);
    void *d_up = nullptr, *d_dn = nullptr;
    KC_CUDA(cudaMalloc(&d_up, bytes));
    if (cudaMalloc(&d_dn, bytes) != cudaSuccess) { cudaFree(d_up); KC_FAIL(KC_ERR_CUDA, "probe buffer allocation failed"); }
    cudaEvent_t e[4];
    for (auto& ev : e) cudaEventCreate(&ev);
    cudaStreamSynchronize(ctx->stream);
    const bool up = direction == 0 || direction == 2, dn = direction == 1 || direction == 2;
    // the second half of the host buffer receives the downloads when both directions run, so they do not share pages
    char* h_dn = (char*)pinned_host + (direction == 2 ? bytes / 2 : 0);
    const size_t n_up = direction == 2 ? bytes / 2 : bytes, n_dn = n_up;
    cudaError_t err = cudaSuccess;
    if (up) cudaEventRecord(e[0], ctx->upload_stream);
    if (dn) cudaEventRecord(e[2], ctx->download_stream);
    for (int i = 0; i < reps && err == cudaSuccess; ++i) {
        if (up) err = cudaMemcpyAsync(d_up, pinned_host, n_up, cudaMemcpyHostToDevice, ctx->upload_stream);
        if (dn && err == cudaSuccess) err = cudaMemcpyAsync(h_dn, d_dn, n_dn, cudaMemcpyDeviceToHost, ctx->download_stream);
    }
    if (up) cudaEventRecord(e[1], ctx->upload_stream);
    if (dn) cudaEventRecord(e[3], ctx->download_stream);
    cudaStreamSynchronize(ctx->upload_stream);
    cudaStreamSynchronize(ctx->download_stream);
    float ms = 0;
    if (h2d_gbs) { *h2d_gbs = 0; if (up && cudaEventElapsedTime(&ms, e[0], e[1]) == cudaSuccess && ms > 0) *h2d_gbs = (double)n_up * reps / (ms / 1e3) / 1e9; }
    if (d2h_gbs) { *d2h_gbs = 0; if (dn && cudaEventElapsedTime(&ms, e[2], e[3]) == cudaSuccess && ms > 0) *d2h_gbs = (double)n_dn * reps / (ms / 1e3) / 1e9; }
    for (auto& ev : e) cudaEventDestroy(ev);
    cudaFree(d_up);
    cudaFree(d_dn);
    KC_CUDA(err);
    return KC_OK;
} KC_ABI_CATCH

extern "C" int32_t kc_context_trim(kc_context* ctx) try {
    // hand the recycled device buffers back to the driver's pool
    if (!ctx) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "ctx is NULL");
    KcGuard g(ctx);
    kc_dev_trim(ctx);
    return KC_OK;
} KC_ABI_CATCH
