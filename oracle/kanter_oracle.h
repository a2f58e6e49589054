/*
 * kanter_oracle.h — C ABI of the CPU ORACLE.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a CPU restatement of the per-pixel
 * evaluation path of lukors/kanter_core (crate `vismut_core` 0.10.0), written
 * from the reference's semantics (file:line citations are in the .cpp).  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load it.  The product (kanter_core_b200/) never does.
 *
 * Parity status: PINNED for everything the reference's own goldens exercise
 * (all 22 golden checks of tests/integration_tests.rs reproduce byte-exactly,
 * see tests/test_oracle_goldens.py).  UNPINNED for the sRGB export (to_u8_srgb:
 * no test of the reference calls it; checked against the formula in float64),
 * for the Nearest / CatmullRom /
 * Gaussian / Lanczos3 resize filters, for Triangle down-sampling and for the
 * [0,1] clamp of the resize's horizontal pass: the reference delegates those to
 * the un-vendored third-party crate image 0.24.0 (Cargo.lock:237-240) and has
 * no golden that exercises them; the oracle restates that crate's published
 * algorithm (imageops/sample.rs).  Independent evidence short of a pin:
 * Triangle, CatmullRom and Lanczos3 agree with Pillow's resampler (another
 * implementation of the same separable design) to float rounding, up- and
 * down-sampling (tests/test_oracle_goldens.py).
 */
#ifndef KANTER_ORACLE_H
#define KANTER_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* node types: order of `enum NodeType`, src/node/node_type.rs:14-28 */
enum {
    KO_INPUT_GRAY = 0, KO_INPUT_RGBA, KO_OUTPUT_GRAY, KO_OUTPUT_RGBA, KO_GRAPH,
    KO_IMAGE, KO_EMBED, KO_WRITE, KO_VALUE, KO_MIX, KO_HEIGHT_TO_NORMAL,
    KO_SEPARATE_RGBA, KO_COMBINE_RGBA
};
/* src/node/mix.rs:20-26 */
enum { KO_ADD = 0, KO_SUBTRACT, KO_MULTIPLY, KO_DIVIDE, KO_POW };
/* src/node/mod.rs:33-40 */
enum { KO_MOST_PIXELS = 0, KO_LEAST_PIXELS, KO_LARGEST_AXES, KO_SMALLEST_AXES,
       KO_SPECIFIC_SLOT, KO_SPECIFIC_SIZE };
/* src/node/mod.rs:63-69 */
enum { KO_NEAREST = 0, KO_TRIANGLE, KO_CATMULL_ROM, KO_GAUSSIAN, KO_LANCZOS3 };

/* error codes = 1 + discriminant of TexProError (src/error.rs:5-27); 0 = Ok */
enum {
    KO_OK = 0, KO_ERR_GENERIC = 1, KO_ERR_CANCELED = 2, KO_ERR_IMAGE = 3,
    KO_ERR_INVALID_BUFFER_COUNT = 4, KO_ERR_INVALID_NODE_ID = 5,
    KO_ERR_INVALID_NODE_TYPE = 6, KO_ERR_INVALID_SLOT_ID = 7,
    KO_ERR_INVALID_SLOT_TYPE = 8, KO_ERR_INVALID_EDGE = 9, KO_ERR_NO_SLOT_DATA = 10,
    KO_ERR_SLOT_OCCUPIED = 11, KO_ERR_SLOT_NOT_OCCUPIED = 12, KO_ERR_UNABLE_TO_LOCK = 13,
    KO_ERR_NODE_PROCESSING = 14, KO_ERR_POISON = 15, KO_ERR_TRY_LOCK = 16,
    KO_ERR_NODE_DIRTY = 17, KO_ERR_IO = 18, KO_ERR_INVALID_NAME = 19
};

/* ---- per-plane / per-image primitives (row-major f32 planes) ------------- */

/* src/shared.rs:16-56  deconstruct_image: interleaved u8 -> 4 f32 planes */
void ko_deconstruct_u8(const uint8_t* samples, uint32_t w, uint32_t h, uint32_t channels,
                       float* r, float* g, float* b, float* a);
/* src/slot_image.rs:142-207  to_u8 / to_u8_srgb.  g,b,a == NULL => Gray image */
void ko_to_u8(const float* r, const float* g, const float* b, const float* a,
              uint64_t n, int srgb, uint8_t* out_rgba8);
/* src/node/mix.rs:136-192  one plane */
void ko_mix_plane(int op, const float* l, const float* r, uint64_t n, float* out);
/* src/slot_image.rs:242-253  Rgba -> Gray */
void ko_rgb_to_gray(const float* r, const float* g, const float* b, uint64_t n, float* out);
/* src/node/height_to_normal.rs:16-77 */
void ko_height_to_normal(const float* hgt, uint32_t w, uint32_t h,
                         float* out_r, float* out_g, float* out_b);
/* the same on rows [y0, y0+h) of an image h_full rows tall; halo_row = row (y0-1) mod h_full */
void ko_height_to_normal_strip(const float* strip, uint32_t w, uint32_t h, uint32_t h_full,
                               const float* halo_row, float* out_r, float* out_g, float* out_b);
/* src/shared.rs:159-199 -> image 0.24.0 imageops::resize on one Luma<f32> plane */
void ko_resize_plane(const float* src, uint32_t sw, uint32_t sh,
                     float* dst, uint32_t dw, uint32_t dh, int filter);
/* image 0.24.0 imageops/sample.rs weight table for one axis.  Writes, for every
 * output index o, left[o], count[o] and count[o] normalised weights at
 * weights[o*max_taps ..].  Returns the largest tap count (call with
 * weights == NULL to query it). */
uint32_t ko_resize_weights(uint32_t src_len, uint32_t dst_len, int filter,
                           uint32_t* left, uint32_t* count, float* weights, uint32_t max_taps);

/* ---- graph evaluation (src/node/node_type.rs:213-267 process_node, run over
 *      the whole DAG in the order the engine would, src/engine.rs:128-307) -- */
typedef struct ko_graph ko_graph;

ko_graph* ko_graph_new(void);
void ko_graph_free(ko_graph*);
/* nested is copied (the reference clones the NodeGraph, src/node/graph.rs:22) */
int ko_graph_add_node(ko_graph*, uint32_t node_id, int node_type, float value, int mix_type,
                      const char* name, const ko_graph* nested, uint32_t embed_id,
                      int resize_policy, uint32_t policy_slot, uint32_t policy_w,
                      uint32_t policy_h, int resize_filter);
int ko_graph_add_edge(ko_graph*, uint32_t output_id, uint32_t input_id,
                      uint32_t output_slot, uint32_t input_slot);
/* decoded pixels of an Image node (src/node/image.rs); none => 1x1 magenta */
int ko_graph_set_image_u8(ko_graph*, uint32_t node_id, const uint8_t* samples,
                          uint32_t w, uint32_t h, uint32_t channels);
/* LiveGraph::add_input_slot_data (src/live_graph.rs:347-350); planes: 1 or 4 */
int ko_graph_add_input_f32(ko_graph*, uint32_t node_id, int is_rgba, uint32_t w, uint32_t h,
                           const float* const* planes);
/* LiveGraph::embed_slot_data_with_id (src/live_graph.rs:324-341) */
int ko_graph_embed_f32(ko_graph*, uint32_t embed_id, int is_rgba, uint32_t w, uint32_t h,
                       const float* const* planes);
/* Evaluate every node (each node's own loop is single-threaded, as in the
 * reference); up to max_threads ready nodes run concurrently
 * (src/process_pack.rs:27).  Returns 0 or the first error code. */
int ko_graph_eval(ko_graph*, int max_threads);
/* slot data of an evaluated node.  planes[] receive borrowed pointers. */
int ko_graph_slot(const ko_graph*, uint32_t node_id, uint32_t slot_id, int* is_rgba,
                  uint32_t* w, uint32_t* h, const float** planes);
/* number of slot datas a node produced (and their slot ids) */
int ko_graph_slot_ids(const ko_graph*, uint32_t node_id, uint32_t* slot_ids, int cap);

/* Evaluate `copies` independent clones of the graph concurrently on up to
 * max_threads threads (one thread per ready node): the reference engine's
 * best case on a multi-core host.  Returns seconds of wall time, <0 on error. */
double ko_graph_eval_batch(const ko_graph*, int copies, int max_threads);

#ifdef __cplusplus
}
#endif
#endif
