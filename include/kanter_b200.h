/*
 * kanter_b200.h — C ABI of the B200-native evaluation backend for the
 * per-pixel path of lukors/kanter_core (crate `vismut_core` 0.10.0).
 *
 * The reference is pure safe Rust with no FFI of its own, so
 * this header IS the boundary a Rust `vismut_core` shim binds to.  Every entry
 * point names the reference item (file:line under the reference root) whose
 * behaviour it replaces.  INTEGRATION.md shows the Rust `extern "C"` block and
 * where each call goes inside `process_node` / `engine::process_loop`.
 *
 * Conventions
 *  - plain C types only: pointers, sizes, POD structs; no C++/torch types.
 *  - every function returns an int32 status: 0 = Ok, 1 + the discriminant of
 *    `TexProError` (src/error.rs:5-27) otherwise, KC_ERR_CUDA for device errors;
 *    kc_last_error() returns a thread-local message for the last failure.
 *  - pixel data are row-major planar f32, one plane per channel, exactly the
 *    reference's `SlotImage::{Gray(plane), Rgba([plane;4])}`
 *    (src/slot_image.rs:12-19).  Planes live in HBM, are immutable once
 *    written and reference-counted (the reference's `Arc<TransientBufferContainer>`).
 *  - work is enqueued on the context's CUDA stream; functions that hand data
 *    back to the host synchronise that stream, nothing else does.
 *  - there is no CPU fallback: without a usable sm_100 device
 *    kc_context_create fails with KC_ERR_CUDA.
 *  - nothing unwinds through this boundary: a C++ exception inside the library
 *    (host memory running out included) comes back as KC_ERR_GENERIC.
 */
#ifndef KANTER_B200_H
#define KANTER_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define KC_ABI_VERSION 1

/* ---- enums ---------------------------------------------------------------- */

/* status codes: 1 + discriminant of TexProError, src/error.rs:5-27 */
enum {
    KC_OK = 0, KC_ERR_GENERIC = 1, KC_ERR_CANCELED = 2, KC_ERR_IMAGE = 3,
    KC_ERR_INVALID_BUFFER_COUNT = 4, KC_ERR_INVALID_NODE_ID = 5,
    KC_ERR_INVALID_NODE_TYPE = 6, KC_ERR_INVALID_SLOT_ID = 7,
    KC_ERR_INVALID_SLOT_TYPE = 8, KC_ERR_INVALID_EDGE = 9, KC_ERR_NO_SLOT_DATA = 10,
    KC_ERR_SLOT_OCCUPIED = 11, KC_ERR_SLOT_NOT_OCCUPIED = 12, KC_ERR_UNABLE_TO_LOCK = 13,
    KC_ERR_NODE_PROCESSING = 14, KC_ERR_POISON = 15, KC_ERR_TRY_LOCK = 16,
    KC_ERR_NODE_DIRTY = 17, KC_ERR_IO = 18, KC_ERR_INVALID_NAME = 19,
    KC_ERR_CUDA = 100, KC_ERR_INVALID_ARGUMENT = 101
};

/* `enum NodeType`, src/node/node_type.rs:14-28 (same order) */
enum {
    KC_NODE_INPUT_GRAY = 0, KC_NODE_INPUT_RGBA, KC_NODE_OUTPUT_GRAY, KC_NODE_OUTPUT_RGBA,
    KC_NODE_GRAPH, KC_NODE_IMAGE, KC_NODE_EMBED, KC_NODE_WRITE, KC_NODE_VALUE, KC_NODE_MIX,
    KC_NODE_HEIGHT_TO_NORMAL, KC_NODE_SEPARATE_RGBA, KC_NODE_COMBINE_RGBA
};
/* `enum MixType`, src/node/mix.rs:20-26 */
enum { KC_MIX_ADD = 0, KC_MIX_SUBTRACT, KC_MIX_MULTIPLY, KC_MIX_DIVIDE, KC_MIX_POW };
/* `enum ResizePolicy`, src/node/mod.rs:33-40 */
enum { KC_POLICY_MOST_PIXELS = 0, KC_POLICY_LEAST_PIXELS, KC_POLICY_LARGEST_AXES,
       KC_POLICY_SMALLEST_AXES, KC_POLICY_SPECIFIC_SLOT, KC_POLICY_SPECIFIC_SIZE };
/* `enum ResizeFilter`, src/node/mod.rs:63-69 */
enum { KC_FILTER_NEAREST = 0, KC_FILTER_TRIANGLE, KC_FILTER_CATMULL_ROM, KC_FILTER_GAUSSIAN,
       KC_FILTER_LANCZOS3 };
/* `enum SlotType`, src/node/mod.rs:197-202 */
enum { KC_SLOT_GRAY = 0, KC_SLOT_RGBA, KC_SLOT_GRAY_OR_RGBA };
/* `enum Side`, src/node/mod.rs:101-105 */
enum { KC_SIDE_INPUT = 0, KC_SIDE_OUTPUT };
/* `enum NodeState`, src/live_graph.rs:23-37 */
enum { KC_STATE_CLEAN = 0, KC_STATE_DIRTY, KC_STATE_REQUESTED, KC_STATE_PRIORITISED,
       KC_STATE_PROCESSING, KC_STATE_PROCESSING_DIRTY };
/* `SlotImage` discriminant, src/slot_image.rs:16-19 */
enum { KC_IMAGE_GRAY = 0, KC_IMAGE_RGBA = 1 };

/* arithmetic mode of the device kernels */
enum {
    /* IEEE add/sub/mul/div/sqrt in the reference's order, no FMA contraction,
     * pow through fp64: reproduces the reference's CPU results bit for bit on
     * everything its goldens pin. */
    KC_MATH_EXACT = 0,
    /* same results within 1e-5 relative / 1e-6 absolute (BASELINE.json
     * north_star): MUFU-based pow, algebraically reduced HeightToNormal, FMA
     * in the resize accumulation. */
    KC_MATH_FAST = 1
};

/* ---- handles and PODs ----------------------------------------------------- */

typedef struct kc_context kc_context;        /* one device + stream + plane pool */
typedef struct kc_plane kc_plane;            /* Arc<TransientBufferContainer>, src/transient_buffer.rs:188-191 */
typedef struct kc_graph kc_graph;            /* NodeGraph, src/node_graph.rs:17-22 */
typedef struct kc_live_graph kc_live_graph;  /* LiveGraph, src/live_graph.rs:63-74 */

typedef struct kc_options {
    int32_t math_mode;        /* KC_MATH_*            (default KC_MATH_EXACT) */
    int32_t fuse;             /* 1: fuse chains of elementwise nodes into one kernel (default 1) */
    /* 0 (default): the second (horizontal) pass of a resize clamps to [0,1], as image-0.24's
     * horizontal_sample does for f32.  No golden of the reference pins that clamp, so it
     * can be switched off: 1 leaves overshoot (Lanczos3, CatmullRom) and out-of-range inputs as they are. */
    int32_t resize_unclamped;
    int32_t reserved[5];
} kc_options;

/* SlotImage, src/slot_image.rs:16-19.  planes[1..3] are NULL for Gray.  A
 * kc_image owns one reference on each plane; drop it with kc_image_release. */
typedef struct kc_image {
    int32_t kind;             /* KC_IMAGE_GRAY | KC_IMAGE_RGBA */
    uint32_t width, height;   /* of planes[0] (SlotImage::size, src/slot_image.rs:116-121) */
    kc_plane* planes[4];
} kc_image;

/* SlotData, src/slot_data.rs:35-39 */
typedef struct kc_slot_data {
    uint32_t node_id, slot_id;
    kc_image image;
} kc_slot_data;

/* EmbeddedSlotData, src/node/embed.rs:15-20 */
typedef struct kc_embedded_slot_data {
    uint32_t slot_data_id, slot_id;
    kc_image image;
} kc_embedded_slot_data;

/* Edge, src/edge.rs:9-14 */
typedef struct kc_edge {
    uint32_t output_id, input_id, output_slot, input_slot;
} kc_edge;

/* Node, src/node/mod.rs:114-123 (priority/cancel are engine state, not data) */
typedef struct kc_node_desc {
    uint32_t node_id;
    int32_t node_type;        /* KC_NODE_* */
    float value;              /* Value(f32) */
    int32_t mix_type;         /* Mix(MixType) */
    const char* name;         /* Input.. / Output.. name, Image/Write path; may be NULL */
    const kc_graph* graph;    /* Graph(NodeGraph) payload; copied by the callee */
    uint32_t embed_id;        /* Embed(EmbeddedSlotDataId) */
    int32_t resize_policy;    /* KC_POLICY_* */
    uint32_t policy_slot;     /* SpecificSlot(SlotId) */
    uint32_t policy_width, policy_height; /* SpecificSize(Size) */
    int32_t resize_filter;    /* KC_FILTER_* */
} kc_node_desc;

/* Slot, src/node/mod.rs:223-238 */
typedef struct kc_slot {
    char name[32];
    uint32_t slot_id;
    int32_t slot_type;        /* KC_SLOT_* */
} kc_slot;

/* ---- library -------------------------------------------------------------- */

int32_t kc_abi_version(void);
const char* kc_last_error(void);
const char* kc_error_string(int32_t code);  /* Display for TexProError, src/error.rs:37-64 */
void kc_free(void* p);                      /* frees memory the library malloc'ed for the caller */
/* pinned host staging memory for uploads/downloads */
int32_t kc_host_alloc(size_t bytes, void** out);          /* on the NUMA node of the current CUDA device, where the host tells */
int32_t kc_host_free(void* p);
/* NUMA placement (no reference counterpart; it is what the 8-GPU end-to-end curve hangs on): page-locked memory on the
 * node `device` is attached to (mmap + mbind + cudaHostRegister; plain cudaHostAlloc when the topology is unknown or
 * KC_NO_NUMA is set), the topology as seen from here, and moving the calling thread next to the device. */
int32_t kc_host_alloc_near_device(int32_t device, size_t bytes, void** out);   /* free with kc_host_free */
int32_t kc_numa_info(int32_t device, int32_t* device_node, int32_t* thread_node, int32_t* nodes);
int32_t kc_bind_thread_near_device(int32_t device, int32_t* bound);
/* the PCIe ceiling next to it: `reps` plain cudaMemcpyAsync copies of `bytes` between pinned_host and a scratch device buffer,
 * timed with events; direction 0 host->device, 1 device->host, 2 both at once (half the buffer each way) */
int32_t kc_context_pcie_probe(kc_context* ctx, void* pinned_host, size_t bytes, int32_t reps, int32_t direction, double* h2d_gbs, double* d2h_gbs);

/* ---- context: replaces TextureProcessor's worker threads,
 *      src/texture_processor.rs:34-56 (engine + transient-buffer queue) ------ */
void kc_options_default(kc_options* o);
int32_t kc_context_create(int32_t device, const kc_options* opts, kc_context** out);
/* the same, with the caller's CUDA stream (a cudaStream_t; it must belong to `device` and outlive the context)
 * as the stream every kernel is enqueued on: evaluations are ordered with the caller's other work on that stream.
 * Uploads from host memory and RGBA8 downloads still use the context's two copy streams, tied in by events. */
int32_t kc_context_create_on_stream(int32_t device, const kc_options* opts, void* cuda_stream, kc_context** out);
/* Waits for the streams, then frees the streams, the buffer cache and the spill buffers.  Planes,
 * images and live graphs made from the context may be released AFTER this call (the reference's
 * Arc-owned SlotImages outlive its Engine the same way); every other use of them fails with
 * KC_ERR_INVALID_ARGUMENT, and `ctx` itself must not be passed to anything again. */
int32_t kc_context_destroy(kc_context* ctx);
int32_t kc_context_synchronize(kc_context* ctx);
int32_t kc_context_device(const kc_context* ctx, int32_t* device);
/* the cudaStream_t all work is enqueued on (as void*), for callers that time it */
int32_t kc_context_stream(const kc_context* ctx, void** stream);
int32_t kc_context_set_math_mode(kc_context* ctx, int32_t mode);
int32_t kc_context_set_fuse(kc_context* ctx, int32_t fuse);
int32_t kc_context_set_resize_unclamped(kc_context* ctx, int32_t unclamped);   /* kc_options.resize_unclamped */
/* counters: kernels launched by this library on the context since creation,
 * bytes currently held by live planes (TransientBufferQueue::bytes_memory,
 * src/transient_buffer.rs:413-420) */
int32_t kc_context_stats(const kc_context* ctx, uint64_t* kernel_launches, uint64_t* bytes_live);

/* ---- PNG codec on the host (zlib), replaces the `image` crate at the three call sites that
 *      touch files: read_slot_image src/shared.rs:218-261, Image node src/node/image.rs:10-26,
 *      Write node src/node/write.rs:5-21.  Decoding yields the interleaved 8-bit samples
 *      `DynamicImage::as_flat_samples_u8` would (1..4 channels); buffers are freed with kc_free. */
int32_t kc_png_decode(const uint8_t* data, size_t n, uint8_t** samples, uint32_t* w, uint32_t* h, uint32_t* channels);
int32_t kc_png_decode_file(const char* path, uint8_t** samples, uint32_t* w, uint32_t* h, uint32_t* channels);
int32_t kc_png_encode(const uint8_t* samples, uint32_t w, uint32_t h, uint32_t channels, uint8_t** png, size_t* n);
int32_t kc_png_encode_file(const char* path, const uint8_t* samples, uint32_t w, uint32_t h, uint32_t channels);

/* tuning knobs for the sweep scripts (0 = library default): "tile_v", "ctas", "stages",
 * "src_soft_cap", "resize_threads"; process-wide */
int32_t kc_debug_set_tuning(const char* key, int32_t value);
/* NVRTC-only check of the kernel generator (kc_jit.cu): compiles the specialised kernel of a one-segment tape
 * (planner words op | arg << 8) for sm_100a; needs no GPU */
int32_t kc_debug_jit_compile(const uint32_t* instr, uint32_t n_instr, int32_t exact, int32_t v, int32_t ctas, size_t* cubin_bytes);
/* (tile float4s per thread, resident CTAs per SM, pipeline stages) of the last fused elementwise launch; V < 0: the
 * specialised kernel ran */
int32_t kc_debug_last_tile_config(int32_t* v, int32_t* ctas, int32_t* stages);
/* Under the automatic policy a hot tape is compiled on a background thread while the interpreter keeps serving it
 * (no stall in the caller's loop; KC_JIT_SYNC=1 compiles in the launching thread instead).  This waits until no
 * compile is running, at most timeout_ms; *still_running = compiles left.  Benchmarks call it to end their warm-up. */
int32_t kc_debug_jit_wait(int32_t timeout_ms, int32_t* still_running);
/* ---- spill queue: TransientBufferQueue, src/transient_buffer.rs:250-411 + TextureProcessor::memory_threshold,
 *      src/texture_processor.rs:19.  Above `bytes` of live planes in HBM the least recently used ones move to
 *      pinned host memory (the reference writes them to disk) and come back when something reads them.
 *      0 = no limit (the default: 180 GB of HBM). */
int32_t kc_context_set_memory_threshold(kc_context* ctx, uint64_t bytes);
/* ---- priority admission: ProcessPackManager, src/process_pack.rs:33-96 + PriorityPropagator, src/priority.rs:101-127.
 *      An engine turn (kc_live_graph_update_turn, and every turn kc_live_graph_await_clean takes on an auto_update graph)
 *      admits at most this many closest-processable nodes, highest PROPAGATED priority first.
 *      set_max_processing_nodes, src/texture_processor.rs:111-114; default = the host's logical CPUs (num_cpus::get()). */
int32_t kc_context_set_max_processing_nodes(kc_context* ctx, size_t count);

/* Concurrent section: the reference's engine runs up to `max_processing_nodes` ready nodes side by side on its thread
 * pool (src/process_pack.rs:27, src/engine.rs:288), and nodes of DIFFERENT live graphs are the easy case.  Between
 * begin and end, evaluations that are served by evaluation replay (kc_live_graph_set_replay) are launched on `lanes`
 * side streams instead of the context's stream -- a plan keeps its lane, so replays of the same live graph stay in
 * order -- and graphs whose kernels are bound by different units (a glibc-exact pow cone on the fp64 pipe next to
 * HBM-bound kernels) overlap on the device.  The caller's side of the contract: the live graphs evaluated inside one
 * section do not consume each other's results, and nothing else reads or overwrites their inputs or results before the
 * section ends.  kc_context_concurrent_end -- and, as a safety net, kc_context_synchronize, every download, and every
 * call that computes on the context's stream -- makes that stream wait for the lanes.  lanes = 1 is the ordinary
 * behaviour. */
int32_t kc_context_concurrent_begin(kc_context* ctx, int32_t lanes);
int32_t kc_context_concurrent_end(kc_context* ctx);
int32_t kc_context_max_processing_nodes(const kc_context* ctx, size_t* count);
int32_t kc_context_spill_stats(const kc_context* ctx, uint64_t* bytes_spilled, uint64_t* spills, uint64_t* reloads);
int32_t kc_plane_in_memory(const kc_plane* p, int32_t* in_memory);   /* TransientBufferContainer::in_memory */
/* hand the device buffers the context keeps for reuse back to the driver's pool */
int32_t kc_context_trim(kc_context* ctx);
/* per-launch device timing: while on, every kernel the library launches on the
 * context is bracketed by CUDA events on the context's stream.  kind: 0 fused
 * elementwise tape, 1 fill, 2 u8->f32, 3 HeightToNormal, 4 resize vertical,
 * 5 resize horizontal, -1 all.  Reading waits for the stream and resets the sum. */
int32_t kc_context_set_timing(kc_context* ctx, int32_t on);
int32_t kc_context_timing_read(kc_context* ctx, int32_t kind, double* total_ms, uint64_t* launches);
/* CUDA events on the context's stream, for device-side timing */
int32_t kc_event_create(void** out_event);
int32_t kc_event_destroy(void* event);
int32_t kc_event_record(kc_context* ctx, void* event);
/* on the context's download stream: completes when every kc_image_to_u8_async /
 * kc_live_graph_read_rgba_async enqueued so far has delivered its bytes */
int32_t kc_event_record_download(kc_context* ctx, void* event);
int32_t kc_event_synchronize(void* event);
int32_t kc_event_elapsed_ms(void* start_event, void* stop_event, float* ms);  /* waits for stop */

/* ---- planes: TransientBufferContainer / Buffer, src/slot_image.rs:12,
 *      src/transient_buffer.rs:188-247 --------------------------------------- */
int32_t kc_plane_create(kc_context* ctx, uint32_t w, uint32_t h, kc_plane** out);          /* uninitialised */
/* vec![v; n], kept as a descriptor until pixels are needed.  ctx may be NULL for a descriptor that is
 * only ever measured (kc_plane_size, kc_calculate_size): no device is touched. */
int32_t kc_plane_from_value(kc_context* ctx, uint32_t w, uint32_t h, float v, kc_plane** out);
int32_t kc_plane_from_host(kc_context* ctx, uint32_t w, uint32_t h, const float* host, kc_plane** out);
/* adopt caller-owned device memory (16-byte aligned, w*h floats); never freed by the library */
int32_t kc_plane_wrap_device(kc_context* ctx, uint32_t w, uint32_t h, void* device_ptr, kc_plane** out);
int32_t kc_plane_retain(kc_plane* p);
int32_t kc_plane_release(kc_plane* p);
int32_t kc_plane_size(const kc_plane* p, uint32_t* w, uint32_t* h);
int32_t kc_plane_is_constant(const kc_plane* p, int32_t* is_const, float* value);
/* device address of the pixels (materialises a constant descriptor) */
int32_t kc_plane_device_ptr(kc_plane* p, void** device_ptr);
int32_t kc_plane_upload(kc_plane* p, const float* host);      /* only before first use */
int32_t kc_plane_download(kc_plane* p, float* host);          /* synchronises */

/* ---- images: SlotImage, src/slot_image.rs ---------------------------------- */
/* deconstruct_image + read_slot_image (src/shared.rs:16-56,218-261): decoded
 * interleaved u8 samples -> Rgba planes (sample/255; absent colour 0, absent alpha 1) */
int32_t kc_image_from_u8(kc_context* ctx, const uint8_t* samples, uint32_t w, uint32_t h,
                         uint32_t channels, kc_image* out);
/* host f32 planes (1 for Gray, 4 for Rgba) -> device image */
int32_t kc_image_from_host_planes(kc_context* ctx, int32_t kind, uint32_t w, uint32_t h,
                                  const float* const* planes, kc_image* out);
/* SlotImage::from_value, src/slot_image.rs:28-64 */
/* deferred upload: the planes stay in the caller's pinned memory until something reads them (then they go up on
 * the upload stream); a plane no node reads -- the alpha of an image that only feeds Mix -- never crosses PCIe.
 * The host memory must stay valid and unchanged until the context has been synchronised after the last use. */
int32_t kc_plane_from_host_deferred(kc_context* ctx, uint32_t w, uint32_t h, const float* host, kc_plane** out);
int32_t kc_image_from_host_planes_deferred(kc_context* ctx, int32_t kind, uint32_t w, uint32_t h,
                                           const float* const* planes, kc_image* out);
/* bytes copied host->device and device->host on behalf of the caller since the context was created */
int32_t kc_context_transfer_stats(const kc_context* ctx, uint64_t* h2d_bytes, uint64_t* d2h_bytes);
int32_t kc_image_from_value(kc_context* ctx, uint32_t w, uint32_t h, float v, int32_t rgba, kc_image* out);
/* SlotImage::as_type, src/slot_image.rs:212-256 */
int32_t kc_image_as_type(kc_context* ctx, const kc_image* in, int32_t rgba, kc_image* out);
/* SlotImage::to_u8 / to_u8_srgb, src/slot_image.rs:142-207: host_rgba8 gets w*h*4 bytes */
int32_t kc_image_to_u8(kc_context* ctx, const kc_image* in, int32_t srgb, uint8_t* host_rgba8);
/* same, but returns once the conversion kernel and the copy to (pinned) host memory are enqueued;
 * host_rgba8 is valid after kc_context_synchronize.  The copy runs on the context's download
 * stream, so uploads of the next evaluation (kc_plane_from_host) overlap it. */
int32_t kc_image_to_u8_async(kc_context* ctx, const kc_image* in, int32_t srgb, uint8_t* host_rgba8);
/* same, result left in device memory (w*h*4 bytes, caller-owned) */
int32_t kc_image_to_u8_device(kc_context* ctx, const kc_image* in, int32_t srgb, void* device_rgba8);
int32_t kc_image_download(kc_context* ctx, const kc_image* in, float* const* host_planes);
/* turn the image's lazily evaluated planes into pixels in HBM with one fused launch;
 * constant planes stay descriptors unless include_constants is set */
int32_t kc_image_materialize(kc_context* ctx, const kc_image* in, int32_t include_constants);
int32_t kc_image_retain(const kc_image* img);
int32_t kc_image_release(kc_image* img);

/* ---- per-node operators: the bodies of the files under src/node/ ------------------------ */
/* mix::process, src/node/mix.rs:51-134.  left/right may be NULL (unconnected). */
int32_t kc_mix(kc_context* ctx, int32_t mix_type, const kc_image* left, const kc_image* right, kc_image* out);
/* height_to_normal::process, src/node/height_to_normal.rs:16-77 (Gray in, Rgba out) */
int32_t kc_height_to_normal(kc_context* ctx, const kc_image* in, kc_image* out);
/* HeightToNormal on a horizontal strip of a taller image (multi-GPU tiling):
 * `strip` holds rows [y0, y0+h) of a Gray image `full_height` tall, `halo_row` (w x 1) is the
 * row above the strip -- row y0-1, or the image's LAST row for the strip that starts at 0
 * (wrapping_sample_subtract, src/node/process_shared.rs:52-60).  Bit-identical to the
 * corresponding rows of kc_height_to_normal on the whole image. */
int32_t kc_height_to_normal_strip(kc_context* ctx, const kc_image* strip, kc_plane* halo_row, uint32_t full_height, kc_image* out);
/* ---- halo rows through peer memory (one process per GPU; SURVEY.md section 8e: "NVLink peer copies only for
 *      halo exchange").  The GPU that owns the strip above publishes its last row into a mailbox in its own
 *      HBM; the GPU below maps the mailbox with CUDA IPC and its HeightToNormal kernel waits for the step's
 *      flag and reads the row straight out of peer memory.  Steps count 1, 2, 3, ...; a mailbox holds two. */
typedef struct kc_halo_link kc_halo_link;
int32_t kc_halo_outbox_create(kc_context* ctx, uint32_t width, kc_halo_link** out);
int32_t kc_halo_outbox_handle(const kc_halo_link* outbox, uint8_t handle[64]);              /* cudaIpcMemHandle_t bytes: send to the rank below */
int32_t kc_halo_inbox_open(kc_context* ctx, const uint8_t handle[64], uint32_t width, kc_halo_link** out);
int32_t kc_halo_inbox_local(kc_context* ctx, const kc_halo_link* outbox, kc_halo_link** out); /* same process: ring of one rank, tests */
int32_t kc_halo_publish(kc_halo_link* outbox, kc_plane* plane, uint32_t row, uint64_t step);
int32_t kc_height_to_normal_strip_peer(kc_context* ctx, const kc_image* strip, const kc_halo_link* inbox, uint64_t step,
                                       uint32_t full_height, kc_image* out);
/* the whole exchange of a step in ONE launch: the stencil kernel publishes the strip's last row into `outbox`, reads the row
 * above out of `inbox` and acknowledges it there (no kc_halo_publish, no separate ack) */
int32_t kc_height_to_normal_strip_exchange(kc_context* ctx, const kc_image* strip, kc_halo_link* outbox, const kc_halo_link* inbox,
                                           uint64_t step, uint32_t full_height, kc_image* out);
int32_t kc_halo_timeouts(kc_context* ctx, uint32_t* count);   /* waits that gave up after 2 s; 0 in a healthy run */
int32_t kc_halo_link_destroy(kc_halo_link* link);
/* device-to-device copy of whole rows between planes of equal width (halo rows; works
 * across devices with peer access: one cudaMemcpyAsync over NVLink) */
int32_t kc_plane_copy_rows(kc_context* ctx, kc_plane* dst, uint32_t dst_row, kc_plane* src, uint32_t src_row, uint32_t rows);
/* resize_buffers' per-plane imageops::resize, src/shared.rs:155-201 */
int32_t kc_resize(kc_context* ctx, const kc_image* in, uint32_t w, uint32_t h, int32_t filter, kc_image* out);
/* rows [row_begin, row_begin + row_count) of that result (out is w x row_count): the share of one GPU when
 * a resize is tiled over GPUs by output rows; bit-identical to the same rows of kc_resize */
int32_t kc_resize_rows(kc_context* ctx, const kc_image* in, uint32_t w, uint32_t h, int32_t filter,
                       uint32_t row_begin, uint32_t row_count, kc_image* out);
/* separate_rgba::process / combine_rgba::process (plane aliasing),
 * src/node/separate_rgba.rs:38-69, src/node/combine_rgba.rs:14-97.
 * in may be NULL; channels[i] may be NULL. */
int32_t kc_separate_rgba(kc_context* ctx, const kc_image* in, kc_image out[4]);
int32_t kc_combine_rgba(kc_context* ctx, const kc_image* const channels[4], kc_image* out);
/* calculate_size, src/shared.rs:61-139 (sizes only; edges sorted by the callee) */
int32_t kc_calculate_size(const kc_slot_data* slot_datas, size_t n_slot_datas, const kc_edge* edges,
                          size_t n_edges, int32_t policy, uint32_t policy_slot, uint32_t policy_w,
                          uint32_t policy_h, uint32_t* out_w, uint32_t* out_h);
/* THE drop-in seam: process_node, src/node/node_type.rs:213-248, called from the
 * engine at src/engine.rs:288-296.  slot_datas[i] belongs to edges[i].  Writes up
 * to out_cap results (each owning its plane references) and their count. */
int32_t kc_process_node(kc_context* ctx, const kc_node_desc* node,
                        const kc_slot_data* slot_datas, size_t n_slot_datas,
                        const kc_embedded_slot_data* embedded, size_t n_embedded,
                        const kc_slot_data* input_slot_datas, size_t n_input_slot_datas,
                        const kc_edge* edges, size_t n_edges,
                        kc_slot_data* out, size_t out_cap, size_t* n_out);

/* ---- NodeGraph: src/node_graph.rs, src/node/node_type.rs:141-211 ------------ */
int32_t kc_graph_create(kc_graph** out);
int32_t kc_graph_destroy(kc_graph* g);
int32_t kc_graph_clone(const kc_graph* g, kc_graph** out);
int32_t kc_graph_from_json(const char* json_text, kc_graph** out);       /* serde schema, data/invert_graph.json */
int32_t kc_graph_from_path(const char* path, kc_graph** out);            /* NodeGraph::from_path, :33-46 */
int32_t kc_graph_export_json(const kc_graph* g, char** out_text);        /* NodeGraph::export_json, :98-102; kc_free the text */
int32_t kc_graph_export_json_path(const kc_graph* g, const char* path);
int32_t kc_graph_add_node(kc_graph* g, const kc_node_desc* node, uint32_t* out_node_id);  /* add_node, :332-337 */
int32_t kc_graph_add_node_with_id(kc_graph* g, const kc_node_desc* node);                 /* add_node_with_id, :339-348 */
int32_t kc_graph_remove_node(kc_graph* g, uint32_t node_id);                              /* remove_node, :476-485 */
int32_t kc_graph_connect(kc_graph* g, uint32_t output_id, uint32_t input_id, uint32_t output_slot, uint32_t input_slot);     /* connect, :416-446 */
int32_t kc_graph_try_connect(kc_graph* g, uint32_t output_id, uint32_t input_id, uint32_t output_slot, uint32_t input_slot); /* try_connect, :394-413 */
int32_t kc_graph_disconnect_slot(kc_graph* g, uint32_t node_id, int32_t side, uint32_t slot_id);                             /* disconnect_slot, :500-520 */
int32_t kc_graph_remove_edge(kc_graph* g, const kc_edge* e);                              /* remove_edge, :464-474 */
int32_t kc_graph_can_connect(const kc_graph* g, uint32_t output_id, uint32_t input_id, uint32_t output_slot, uint32_t input_slot); /* can_connect, :376-393 */
int32_t kc_graph_connected_edges(const kc_graph* g, uint32_t node_id, int32_t side, uint32_t slot_id, kc_edge* edges, size_t cap, size_t* n); /* :518-537 */
int32_t kc_graph_new_id(kc_graph* g, uint32_t* out);                                          /* new_id, :86-96 */
int32_t kc_graph_rename_output_node(kc_graph* g, uint32_t node_id, const char* new_name, char** old_name); /* :232-270; kc_free old_name */
int32_t kc_graph_node_count(const kc_graph* g, size_t* n);
int32_t kc_graph_node_at(const kc_graph* g, size_t index, kc_node_desc* out);  /* borrowed name/graph pointers */
int32_t kc_graph_node(const kc_graph* g, uint32_t node_id, kc_node_desc* out); /* node, :129-135 */
int32_t kc_graph_set_node(kc_graph* g, const kc_node_desc* node);              /* replace the node with node->node_id */
/* Node.priority (src/node/mod.rs:120): Priority::set_priority / priority / propagated_priority, src/priority.rs:33-45.
 * `propagated` is the value PriorityPropagator::update leaves behind: max(own, children's propagated). */
int32_t kc_graph_set_node_priority(kc_graph* g, uint32_t node_id, int8_t priority);
int32_t kc_graph_node_priority(const kc_graph* g, uint32_t node_id, int8_t* priority, int8_t* propagated);
int32_t kc_graph_edge_count(const kc_graph* g, size_t* n);
int32_t kc_graph_edge_at(const kc_graph* g, size_t index, kc_edge* out);
int32_t kc_graph_input_slot_id_with_name(const kc_graph* g, const char* name, uint32_t* slot_id);   /* :285-290 */
int32_t kc_graph_output_slot_id_with_name(const kc_graph* g, const char* name, uint32_t* slot_id);  /* :292-297 */
int32_t kc_graph_output_ids(const kc_graph* g, uint32_t* ids, size_t cap, size_t* n);                /* :351-357 */
int32_t kc_graph_input_ids(const kc_graph* g, uint32_t* ids, size_t cap, size_t* n);                 /* :359-365 */
/* Node::input_slots / output_slots, src/node/node_type.rs:141-211 */
int32_t kc_node_input_slots(const kc_node_desc* node, kc_slot* slots, size_t cap, size_t* n);
int32_t kc_node_output_slots(const kc_node_desc* node, kc_slot* slots, size_t cap, size_t* n);

/* ---- LiveGraph + engine: src/live_graph.rs, src/engine.rs ------------------- */
int32_t kc_live_graph_create(kc_context* ctx, kc_live_graph** out);          /* TextureProcessor::new_live_graph, src/texture_processor.rs:58-63 */
int32_t kc_live_graph_destroy(kc_live_graph* lg);
int32_t kc_live_graph_set_node_graph(kc_live_graph* lg, const kc_graph* g);  /* set_node_graph (copies g; everything dirty) */
/* the LiveGraph's NodeGraph, for read-only queries with the kc_graph_* getters */
int32_t kc_live_graph_node_graph(const kc_live_graph* lg, const kc_graph** out);
int32_t kc_live_graph_set_use_cache(kc_live_graph* lg, int32_t use_cache);   /* pub use_cache, src/live_graph.rs:72 */
int32_t kc_live_graph_set_auto_update(kc_live_graph* lg, int32_t auto_update); /* pub auto_update, :71 */
/* graph edits that also dirty the affected nodes (src/live_graph.rs:422-566) */
int32_t kc_live_graph_add_node(kc_live_graph* lg, const kc_node_desc* node, uint32_t* out_node_id);
int32_t kc_live_graph_add_node_with_id(kc_live_graph* lg, const kc_node_desc* node);
int32_t kc_live_graph_remove_node(kc_live_graph* lg, uint32_t node_id);
int32_t kc_live_graph_connect(kc_live_graph* lg, uint32_t output_id, uint32_t input_id, uint32_t output_slot, uint32_t input_slot);
int32_t kc_live_graph_disconnect_slot(kc_live_graph* lg, uint32_t node_id, int32_t side, uint32_t slot_id);
int32_t kc_live_graph_set_node(kc_live_graph* lg, const kc_node_desc* node);
/* add_input_slot_data (:347-350) / embed_slot_data_with_id (:324-341) */
int32_t kc_live_graph_add_input_slot_data(kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, const kc_image* image);
int32_t kc_live_graph_clear_input_slot_data(kc_live_graph* lg);
int32_t kc_live_graph_embed_slot_data_with_id(kc_live_graph* lg, const kc_image* image, uint32_t slot_id, uint32_t embed_id);
int32_t kc_live_graph_replace_embedded(kc_live_graph* lg, const kc_image* image, uint32_t embed_id);
/* decoded pixels for an Image(path) node; the codec stays on the host side of
 * the boundary (src/node/image.rs:10-26).  Without data the node yields the
 * reference's 1x1 magenta failure image. */
int32_t kc_live_graph_set_image_data_u8(kc_live_graph* lg, uint32_t node_id, const uint8_t* samples,
                                        uint32_t w, uint32_t h, uint32_t channels);
/* request (:219-227) + the engine's work (src/engine.rs:128-307): evaluate every
 * dirty ancestor of the given nodes and the nodes themselves on the context's
 * stream.  Asynchronous with respect to the host. */
int32_t kc_live_graph_request(kc_live_graph* lg, const uint32_t* node_ids, size_t n);
/* await_clean_read (:181-195): request + wait until the node's data exist */
int32_t kc_live_graph_await_clean(kc_live_graph* lg, uint32_t node_id);
int32_t kc_live_graph_cancel(kc_live_graph* lg);                              /* Node.cancel / shutdown flags, src/node/process_shared.rs:67-69 */
int32_t kc_live_graph_node_state(const kc_live_graph* lg, uint32_t node_id, int32_t* state); /* node_state, :243-249 */
int32_t kc_live_graph_slot_data(const kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, kc_image* out);   /* slot_data, :415-420 (retains) */
int32_t kc_live_graph_slot_in_memory(const kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, int32_t* in_memory); /* :410-412 */
int32_t kc_live_graph_slot_data_size(const kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, uint32_t* w, uint32_t* h); /* :407-409 */
int32_t kc_live_graph_node_slot_ids(const kc_live_graph* lg, uint32_t node_id, uint32_t* slot_ids, size_t cap, size_t* n);    /* node_slot_datas, :389-405 */
int32_t kc_live_graph_buffer_rgba(kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, uint8_t* host_rgba8, size_t cap);   /* buffer_rgba, :93-95 */
int32_t kc_live_graph_buffer_srgba(kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, uint8_t* host_rgba8, size_t cap);  /* try_buffer_srgba's to_u8_srgb, :127-153 */
/* ---- LiveGraph bookkeeping, src/live_graph.rs (id lists: pass ids == NULL to get the count) ---- */
int32_t kc_live_graph_changed_consume(kc_live_graph* lg, uint32_t* ids, size_t cap, size_t* n);                 /* changed_consume, :156-160 */
int32_t kc_live_graph_node_ids_with_state(const kc_live_graph* lg, int32_t state, int32_t without, uint32_t* ids, size_t cap, size_t* n); /* :261-277 */
int32_t kc_live_graph_get_closest_processable(const kc_live_graph* lg, uint32_t node_id, uint32_t* ids, size_t cap, size_t* n);       /* :279-311 */
int32_t kc_live_graph_mark(kc_live_graph* lg, uint32_t node_id, int32_t state);   /* request :219-227 / prioritise :229-237: state change only */
int32_t kc_live_graph_update(kc_live_graph* lg, size_t* n_processed);             /* one engine turn for this graph, src/engine.rs:128-183 */
/* ONE engine turn with priority admission (src/engine.rs:128-307 + src/process_pack.rs:33-96): the closest processable
 * ancestors of the wanted nodes, at most max_processing_nodes of them, highest propagated priority first; the ids of the
 * nodes that ran come back in that order (admitted == NULL: just the count). */
/* Evaluation replay (no reference counterpart: the engine re-runs dirty nodes one by one, src/engine.rs:128-307).  With it on,
 * the second identical kc_live_graph_request over an unchanged graph and the same input PLANES (new pixel content in the same
 * buffers, or the same images embedded again) is captured into one executable CUDA graph and later ones replay it: no planner,
 * no per-node work, one launch.  Results are those of the ordinary evaluation, bit for bit.  Off by default. */
int32_t kc_live_graph_set_replay(kc_live_graph* lg, int32_t on);
int32_t kc_live_graph_replay_stats(const kc_live_graph* lg, uint64_t* captures, uint64_t* replays);
int32_t kc_live_graph_update_turn(kc_live_graph* lg, uint32_t* admitted, size_t cap, size_t* n_admitted);
int32_t kc_live_graph_set_priority(kc_live_graph* lg, uint32_t node_id, int8_t priority);   /* node(id)?.priority.set_priority(v) */
int32_t kc_live_graph_remove_edge(kc_live_graph* lg, const kc_edge* e);           /* remove_edge, :551-566 */
int32_t kc_live_graph_rename_output_node(kc_live_graph* lg, uint32_t node_id, const char* new_name, char** old_name); /* :625-627 */
int32_t kc_live_graph_new_id(kc_live_graph* lg, uint32_t* out);                   /* new_id, :422-424 */
/* await_clean_read + buffer_rgba in one call, so the f32 -> RGBA8 conversion
 * fuses into the kernel that produces the node's planes (never stored as f32) */
int32_t kc_live_graph_read_rgba(kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, int32_t srgb, uint8_t* host_rgba8, size_t cap);
/* same through kc_image_to_u8_async: the bytes are valid after kc_context_synchronize */
int32_t kc_live_graph_read_rgba_async(kc_live_graph* lg, uint32_t node_id, uint32_t slot_id, int32_t srgb, uint8_t* host_rgba8, size_t cap);
/* what the last request did: kernels launched, fused elementwise groups, and the
 * compulsory HBM bytes (inputs read once + outputs written once) of those kernels */
int32_t kc_live_graph_last_run_stats(const kc_live_graph* lg, uint64_t* kernels, uint64_t* fused_groups,
                                     uint64_t* algorithmic_bytes);

#ifdef __cplusplus
}
#endif
#endif /* KANTER_B200_H */
