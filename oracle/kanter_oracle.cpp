// kanter_oracle.cpp — CPU ORACLE (test infrastructure; see kanter_oracle.h).
//
// A CPU restatement of the per-pixel evaluation path of lukors/kanter_core
// (vismut_core 0.10.0).  Every function cites the reference file:line whose
// semantics it restates.  All pixel arithmetic is IEEE binary32, evaluated in
// the reference's order; build with -ffp-contract=off (Rust never contracts
// a*b+c into an FMA) and link glibc's powf/sinf/expf/sqrtf, which is what
// Rust's f32::powf/sin/exp/sqrt lower to on Linux.
//
// Parity status: see the header comment of kanter_oracle.h (pinned by the 22
// reference golden checks; the four non-Triangle resize filters are unpinned).
#include "kanter_oracle.h"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// Pixel buffers: src/slot_image.rs:12-19, src/slot_data.rs:5-39
// ---------------------------------------------------------------------------
struct Plane {
    uint32_t w = 0, h = 0;
    std::vector<float> px;
    Plane(uint32_t w_, uint32_t h_, float v) : w(w_), h(h_), px((size_t)w_ * h_, v) {}
    Plane(uint32_t w_, uint32_t h_) : w(w_), h(h_), px((size_t)w_ * h_) {}
};
using PlaneP = std::shared_ptr<const Plane>;

struct Image {  // SlotImage
    bool rgba = false;
    PlaneP p[4];
    uint32_t w() const { return p[0]->w; }  // size() reads plane 0, slot_image.rs:116-121
    uint32_t h() const { return p[0]->h; }
};

struct SlotData {
    uint32_t node_id, slot_id;
    Image image;
};

PlaneP const_plane(uint32_t w, uint32_t h, float v) { return std::make_shared<Plane>(w, h, v); }
PlaneP pixel_buffer(float v) { return const_plane(1, 1, v); }  // src/node/mod.rs:240-244

// SlotImage::from_value, src/slot_image.rs:28-64 (alpha plane is always 1.0)
Image image_from_value(uint32_t w, uint32_t h, float v, bool rgba) {
    Image im;
    im.rgba = rgba;
    if (rgba) {
        im.p[0] = const_plane(w, h, v);
        im.p[1] = const_plane(w, h, v);
        im.p[2] = const_plane(w, h, v);
        im.p[3] = const_plane(w, h, 1.0f);
    } else {
        im.p[0] = const_plane(w, h, v);
    }
    return im;
}

// ---------------------------------------------------------------------------
// Export: src/slot_image.rs:142-207, src/slot_data.rs:100-109
// ---------------------------------------------------------------------------
// Rust f32::clamp: NaN stays NaN.
inline float rust_clamp01(float v) {
    if (v < 0.0f) return 0.0f;
    if (v > 1.0f) return 1.0f;
    return v;
}
// Rust f32::min: if one side is NaN the other is returned.
inline float rust_min(float a, float b) {
    if (a != a) return b;
    if (b != b) return a;
    return a < b ? a : b;
}
// Rust `as u8`: saturating, truncating, NaN -> 0.
inline uint8_t rust_as_u8(float v) {
    if (v != v) return 0;
    if (v <= 0.0f) return 0;
    if (v >= 255.0f) return 255;
    return (uint8_t)v;
}
inline uint8_t f32_to_u8(float v) {  // slot_image.rs:142-145
    return rust_as_u8(rust_min(rust_clamp01(v) * 255.0f, 255.0f));
}
inline float srgb_to_linear(float v) {  // slot_data.rs:100-109
    if (v <= 0.0f) return v;
    if (v <= 0.04045f) return v / 12.92f;
    return powf((v + 0.055f) / 1.055f, 2.4f);
}
inline uint8_t f32_to_u8_srgb(float v) {  // slot_image.rs:173-176
    return rust_as_u8(rust_min(srgb_to_linear(rust_clamp01(v)) * 255.0f, 255.0f));
}

// ---------------------------------------------------------------------------
// Mix: src/node/mix.rs:136-192
// ---------------------------------------------------------------------------
inline float mix_op(int op, float l, float r) {
    switch (op) {
        case KO_ADD: return l + r;
        case KO_SUBTRACT: return l - r;
        case KO_MULTIPLY: return l * r;
        case KO_DIVIDE: return l / r;
        default: return powf(l, r);
    }
}

void mix_plane(int op, const float* l, const float* r, uint64_t n, float* out) {
    // one switch outside the loop so the scalar loop is what the reference's
    // per-op `from_fn` closures are
    switch (op) {
        case KO_ADD: for (uint64_t i = 0; i < n; ++i) out[i] = l[i] + r[i]; break;
        case KO_SUBTRACT: for (uint64_t i = 0; i < n; ++i) out[i] = l[i] - r[i]; break;
        case KO_MULTIPLY: for (uint64_t i = 0; i < n; ++i) out[i] = l[i] * r[i]; break;
        case KO_DIVIDE: for (uint64_t i = 0; i < n; ++i) out[i] = l[i] / r[i]; break;
        default: for (uint64_t i = 0; i < n; ++i) out[i] = powf(l[i], r[i]); break;
    }
}

// SlotImage::as_type, src/slot_image.rs:212-256
Image as_type(const Image& im, bool rgba) {
    if (im.rgba == rgba) return im;
    Image out;
    out.rgba = rgba;
    uint32_t w = im.w(), h = im.h();
    if (!im.rgba) {
        out.p[0] = im.p[0];
        out.p[1] = im.p[0];
        out.p[2] = im.p[0];
        out.p[3] = const_plane(w, h, 1.0f);
    } else {
        auto g = std::make_shared<Plane>(w, h);
        const float *r = im.p[0]->px.data(), *gg = im.p[1]->px.data(), *b = im.p[2]->px.data();
        size_t n = (size_t)w * h;
        for (size_t i = 0; i < n; ++i) g->px[i] = ((r[i] + gg[i]) + b[i]) / 3.0f;  // :247-250
        out.p[0] = g;
    }
    return out;
}

// ---------------------------------------------------------------------------
// HeightToNormal: src/node/height_to_normal.rs:16-77 with nalgebra 0.29
// Vector3::{normalize, cross} semantics and wrapping_sample_subtract
// (src/node/process_shared.rs:52-60).
// ---------------------------------------------------------------------------
struct V3 { float x, y, z; };
inline V3 v3_normalize(V3 v) {
    float n = sqrtf((v.x * v.x + v.y * v.y) + v.z * v.z);
    return V3{v.x / n, v.y / n, v.z / n};
}
inline V3 v3_cross(V3 a, V3 b) {
    return V3{a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}

// h_full / halo describe a horizontal strip of a taller image (rows [y0, y0+h) of h_full rows,
// halo = row y0-1 with wrap): the multi-GPU tiling of SURVEY.md 8(e).  halo == NULL and
// h_full == h is the reference's whole-image case.
void height_to_normal(const float* hgt, uint32_t w, uint32_t h, uint32_t h_full, const float* halo, float* o0, float* o1,
                      float* o2) {
    const float dx = 1.0f / (float)w;       // :29
    const float dy = 1.0f / (float)h_full;  // :30
    for (uint32_t y = 0; y < h; ++y) {
        const float* up_row = (y != 0) ? hgt + (size_t)(y - 1) * w : (halo ? halo : hgt + (size_t)(h - 1) * w);
        for (uint32_t x = 0; x < w; ++x) {
            const uint32_t xl = (x == 0) ? w - 1 : x - 1;
            const float px = hgt[(size_t)y * w + x];
            const float up = up_row[x];
            const float lf = hgt[(size_t)y * w + xl];
            V3 t = v3_normalize(V3{dx, 0.0f, px - lf});   // :58
            V3 b = v3_normalize(V3{0.0f, dy, up - px});   // :59
            V3 n = v3_normalize(v3_cross(t, b));          // :60
            size_t i = (size_t)y * w + x;
            o0[i] = n.x * 0.5f + 0.5f;                    // :63
            o1[i] = n.y * 0.5f + 0.5f;
            o2[i] = n.z * 0.5f + 0.5f;
        }
    }
}

// ---------------------------------------------------------------------------
// Resize: src/shared.rs:159-199 calls image::imageops::resize (image 0.24.0,
// Cargo.lock:237-240; third-party, source not under /root/reference).  This
// restates that crate's published imageops/sample.rs: filter kernels, the
// per-output tap window, per-output renormalisation, vertical pass first
// (unclamped f32 intermediate) and horizontal pass second with a clamp to
// [f32::DEFAULT_MIN_VALUE, DEFAULT_MAX_VALUE] = [0, 1].
// ---------------------------------------------------------------------------
const float PI_F = 3.14159265358979323846f;  // f32::consts::PI

inline float sinc(float t) {
    float a = t * PI_F;
    if (t == 0.0f) return 1.0f;
    return sinf(a) / a;
}
inline float lanczos(float x, float t) {
    if (fabsf(x) < t) return sinc(x) * sinc(x / t);
    return 0.0f;
}
inline float bc_cubic_spline(float x, float b, float c) {
    float a = fabsf(x);
    float k;
    if (a < 1.0f) {
        k = (12.0f - 9.0f * b - 6.0f * c) * (a * a * a) + (-18.0f + 12.0f * b + 6.0f * c) * (a * a) +
            (6.0f - 2.0f * b);
    } else if (a < 2.0f) {
        k = (-b - 6.0f * c) * (a * a * a) + (6.0f * b + 30.0f * c) * (a * a) +
            (-12.0f * b - 48.0f * c) * a + (8.0f * b + 24.0f * c);
    } else {
        k = 0.0f;
    }
    return k / 6.0f;
}
inline float gaussian(float x, float r) {
    return (1.0f / (sqrtf(2.0f * PI_F) * r)) * expf(-(x * x) / (2.0f * (r * r)));
}
inline float filter_kernel(int filter, float x) {
    switch (filter) {
        case KO_NEAREST: return 1.0f;                                   // box_kernel
        case KO_TRIANGLE: return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f;
        case KO_CATMULL_ROM: return bc_cubic_spline(x, 0.0f, 0.5f);
        case KO_GAUSSIAN: return gaussian(x, 0.5f);
        default: return lanczos(x, 3.0f);
    }
}
inline float filter_support(int filter) {
    switch (filter) {
        case KO_NEAREST: return 0.0f;
        case KO_TRIANGLE: return 1.0f;
        case KO_CATMULL_ROM: return 2.0f;
        case KO_GAUSSIAN: return 3.0f;
        default: return 3.0f;
    }
}

struct Taps {
    uint32_t left;
    std::vector<float> w;
};

// The window/weights computation shared by vertical_sample and
// horizontal_sample of image 0.24.0 (identical in both).
Taps taps_for(uint32_t o, uint32_t src_len, uint32_t dst_len, int filter) {
    const float ratio = (float)src_len / (float)dst_len;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float src_support = filter_support(filter) * sratio;
    float input = ((float)o + 0.5f) * ratio;
    int64_t left = (int64_t)floorf(input - src_support);
    left = std::min<int64_t>(std::max<int64_t>(left, 0), (int64_t)src_len - 1);
    int64_t right = (int64_t)ceilf(input + src_support);
    right = std::min<int64_t>(std::max<int64_t>(right, left + 1), (int64_t)src_len);
    input = input - 0.5f;
    Taps t;
    t.left = (uint32_t)left;
    float sum = 0.0f;
    for (int64_t i = left; i < right; ++i) {
        float w = filter_kernel(filter, ((float)i - input) / sratio);
        t.w.push_back(w);
        sum += w;
    }
    for (float& w : t.w) w /= sum;
    return t;
}

inline float image_clamp(float a, float lo, float hi) {  // image::math::utils::clamp
    if (a < lo) return lo;
    if (a > hi) return hi;
    return a;
}

void resize_plane(const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t dw, uint32_t dh,
                  int filter) {
    // vertical_sample: sw x sh -> sw x dh, f32, no clamp
    std::vector<float> tmp((size_t)sw * dh);
    for (uint32_t oy = 0; oy < dh; ++oy) {
        Taps t = taps_for(oy, sh, dh, filter);
        for (uint32_t x = 0; x < sw; ++x) {
            float acc = 0.0f;
            for (size_t i = 0; i < t.w.size(); ++i)
                acc += src[(size_t)(t.left + i) * sw + x] * t.w[i];
            tmp[(size_t)oy * sw + x] = acc;
        }
    }
    // horizontal_sample: sw x dh -> dw x dh, clamp to [0,1]
    for (uint32_t ox = 0; ox < dw; ++ox) {
        Taps t = taps_for(ox, sw, dw, filter);
        for (uint32_t y = 0; y < dh; ++y) {
            float acc = 0.0f;
            const float* row = &tmp[(size_t)y * sw + t.left];
            for (size_t i = 0; i < t.w.size(); ++i) acc += row[i] * t.w[i];
            dst[(size_t)y * dw + ox] = image_clamp(acc, 0.0f, 1.0f);
        }
    }
}

PlaneP resized(const PlaneP& p, uint32_t w, uint32_t h, int filter) {
    auto out = std::make_shared<Plane>(w, h);
    resize_plane(p->px.data(), p->w, p->h, out->px.data(), w, h, filter);
    return out;
}

// ---------------------------------------------------------------------------
// Graph model: src/node/mod.rs:114-123, src/edge.rs:9-14, src/node_graph.rs:17-22
// ---------------------------------------------------------------------------
struct Node {
    uint32_t node_id = 0;
    int type = 0;
    float value = 0.0f;
    int mix_type = 0;
    std::string name;
    std::shared_ptr<ko_graph> nested;
    uint32_t embed_id = 0;
    int policy = KO_MOST_PIXELS;
    uint32_t policy_slot = 0, policy_w = 0, policy_h = 0;
    int filter = KO_TRIANGLE;
};
struct Edge {
    uint32_t output_id, input_id, output_slot, input_slot;
};
struct ImageU8 {
    uint32_t w, h, ch;
    std::vector<uint8_t> px;
};

}  // namespace

struct ko_graph {
    std::vector<Node> nodes;
    std::vector<Edge> edges;
    std::map<uint32_t, ImageU8> images;
    std::vector<SlotData> inputs;                       // LiveGraph.input_slot_datas
    std::vector<std::pair<uint32_t, Image>> embeds;     // LiveGraph.embedded_slot_datas
    std::vector<SlotData> slot_datas;                   // LiveGraph.slot_datas (results)
};

namespace {

// number of input slots with a given name: Node::input_slots, node_type.rs:141-175
// (only the slot ids are needed here; names map to fixed ids)
int output_slot_count(const Node& n) {  // Node::output_slots, node_type.rs:177-211
    switch (n.type) {
        case KO_OUTPUT_GRAY: case KO_OUTPUT_RGBA: return 0;
        case KO_SEPARATE_RGBA: return 4;
        case KO_GRAPH: {
            int c = 0;
            for (const Node& m : n.nested->nodes)
                if (m.type == KO_OUTPUT_GRAY || m.type == KO_OUTPUT_RGBA) ++c;
            return c;
        }
        default: return 1;
    }
}

const SlotData* with_slot(const std::vector<SlotData>& v, uint32_t slot) {  // process_shared.rs:22-30
    for (const SlotData& s : v)
        if (s.slot_id == slot) return &s;
    return nullptr;
}

struct Size { uint32_t w, h; };
inline uint32_t pixel_count(Size s) { return s.w * s.h; }  // u32 multiply, slot_data.rs:27-29

// calculate_size, src/shared.rs:61-139.  `edges` sorted by input slot.
Size calculate_size(const std::vector<SlotData>& sd, const std::vector<Edge>& edges, const Node& n) {
    switch (n.policy) {
        case KO_MOST_PIXELS: {
            if (sd.empty()) return Size{1, 1};
            size_t best = 0;  // Iterator::max_by keeps the LAST maximum
            for (size_t i = 1; i < sd.size(); ++i)
                if (pixel_count(Size{sd[i].image.w(), sd[i].image.h()}) >=
                    pixel_count(Size{sd[best].image.w(), sd[best].image.h()}))
                    best = i;
            return Size{sd[best].image.w(), sd[best].image.h()};
        }
        case KO_LEAST_PIXELS: {
            size_t best = 0;  // Iterator::min_by keeps the FIRST minimum
            for (size_t i = 1; i < sd.size(); ++i)
                if (pixel_count(Size{sd[i].image.w(), sd[i].image.h()}) <
                    pixel_count(Size{sd[best].image.w(), sd[best].image.h()}))
                    best = i;
            return Size{sd[best].image.w(), sd[best].image.h()};
        }
        case KO_LARGEST_AXES: {
            Size s{0, 0};
            for (const SlotData& d : sd) {
                s.w = std::max(s.w, d.image.w());
                s.h = std::max(s.h, d.image.h());
            }
            return s;
        }
        case KO_SMALLEST_AXES: {
            Size s{UINT32_MAX, UINT32_MAX};
            for (const SlotData& d : sd) {
                s.w = std::min(s.w, d.image.w());
                s.h = std::min(s.h, d.image.h());
            }
            return s;
        }
        case KO_SPECIFIC_SLOT: {
            const Edge* e = nullptr;
            for (const Edge& c : edges)
                if (c.input_slot == n.policy_slot) { e = &c; break; }
            if (!e && !edges.empty()) e = &edges.front();
            if (!e) return Size{1, 1};
            for (const SlotData& d : sd)
                if (d.slot_id == e->output_slot && d.node_id == e->output_id)
                    return Size{d.image.w(), d.image.h()};
            return Size{1, 1};  // unreachable in the reference (`expect`)
        }
        default: return Size{n.policy_w, n.policy_h};
    }
}

// resize_buffers, src/shared.rs:141-216: every plane of every input whose size
// differs from the target is resampled independently.
std::vector<SlotData> resize_buffers(const std::vector<SlotData>& sd, const std::vector<Edge>& edges,
                                     const Node& n) {
    if (sd.empty()) return sd;
    Size size = calculate_size(sd, edges, n);
    std::vector<SlotData> out;
    for (const SlotData& d : sd) {
        if (d.image.w() != size.w || d.image.h() != size.h) {
            SlotData r = d;
            int np = d.image.rgba ? 4 : 1;
            for (int c = 0; c < np; ++c) r.image.p[c] = resized(d.image.p[c], size.w, size.h, n.filter);
            out.push_back(r);
        } else {
            out.push_back(d);
        }
    }
    return out;
}

int eval_graph(ko_graph* g, int max_threads);

// One node: process_node + process_node_internal, src/node/node_type.rs:98-138,213-267
int process_node(ko_graph* g, const Node& node, const std::vector<SlotData>& input_data,
                 const std::vector<Edge>& node_edges, std::vector<SlotData>& out) {
    std::vector<Edge> edges = node_edges;
    std::stable_sort(edges.begin(), edges.end(),
                     [](const Edge& a, const Edge& b) { return a.input_slot < b.input_slot; });
    std::vector<SlotData> rs = resize_buffers(input_data, edges, node);
    // assign_slot_ids, :250-267
    std::vector<SlotData> sd;
    for (const Edge& e : edges) {
        const SlotData* f = nullptr;
        for (const SlotData& d : rs)
            if (e.output_slot == d.slot_id && e.output_id == d.node_id) { f = &d; break; }
        if (!f) return KO_ERR_GENERIC;  // `.unwrap()` panic
        sd.push_back(SlotData{e.input_id, e.input_slot, f->image});
    }

    out.clear();
    switch (node.type) {
        case KO_INPUT_RGBA: {  // src/node/input_rgba.rs:7-13
            if (g->inputs.empty()) return KO_ERR_GENERIC;  // index panic
            out.push_back(SlotData{node.node_id, 0, g->inputs[0].image});
            break;
        }
        case KO_INPUT_GRAY: {  // src/node/input_gray.rs:7-16
            for (const SlotData& d : g->inputs)
                if (d.node_id == node.node_id) { out.push_back(d); break; }
            break;
        }
        case KO_OUTPUT_RGBA:
        case KO_OUTPUT_GRAY: {  // src/node/output.rs:12-33
            if (!sd.empty()) {
                out.push_back(SlotData{node.node_id, 0, sd[0].image});
            } else {
                Image im;
                if (node.type == KO_OUTPUT_RGBA) {
                    im.rgba = true;
                    im.p[0] = pixel_buffer(0.0f);
                    im.p[1] = pixel_buffer(0.0f);
                    im.p[2] = pixel_buffer(0.0f);
                    im.p[3] = pixel_buffer(1.0f);
                } else {
                    im.p[0] = pixel_buffer(0.0f);
                }
                out.push_back(SlotData{node.node_id, 0, im});
            }
            break;
        }
        case KO_GRAPH: {  // src/node/graph.rs:14-51
            ko_graph inner;
            inner.nodes = node.nested->nodes;
            inner.edges = node.nested->edges;
            inner.images = node.nested->images;
            inner.embeds = node.nested->embeds;
            for (const SlotData& d : sd) inner.inputs.push_back(SlotData{d.slot_id, 0, d.image});
            int rc = eval_graph(&inner, 1);
            if (rc) return rc;
            for (const Node& m : inner.nodes) {
                if (m.type != KO_OUTPUT_GRAY && m.type != KO_OUTPUT_RGBA) continue;
                for (const SlotData& d : inner.slot_datas)
                    if (d.node_id == m.node_id) out.push_back(SlotData{node.node_id, m.node_id, d.image});
            }
            break;
        }
        case KO_IMAGE: {  // src/node/image.rs:10-26, src/shared.rs:16-56,218-261
            Image im;
            im.rgba = true;
            auto it = g->images.find(node.node_id);
            if (it == g->images.end()) {  // decode failure => 1x1 magenta
                im.p[0] = pixel_buffer(1.0f);
                im.p[1] = pixel_buffer(0.0f);
                im.p[2] = pixel_buffer(1.0f);
                im.p[3] = pixel_buffer(1.0f);
            } else {
                const ImageU8& u = it->second;
                std::shared_ptr<Plane> pl[4];
                for (int c = 0; c < 4; ++c) pl[c] = std::make_shared<Plane>(u.w, u.h);
                ko_deconstruct_u8(u.px.data(), u.w, u.h, u.ch, pl[0]->px.data(), pl[1]->px.data(),
                                  pl[2]->px.data(), pl[3]->px.data());
                for (int c = 0; c < 4; ++c) im.p[c] = pl[c];
            }
            out.push_back(SlotData{node.node_id, 0, im});
            break;
        }
        case KO_EMBED: {  // src/node/embed.rs:33-50
            const Image* f = nullptr;
            for (auto& e : g->embeds)
                if (e.first == node.embed_id) { f = &e.second; break; }
            if (!f) return KO_ERR_NODE_PROCESSING;
            out.push_back(SlotData{node.node_id, 0, *f});
            break;
        }
        case KO_WRITE:  // src/node/write.rs:5-21 (file I/O; out of scope, yields no data)
            break;
        case KO_VALUE: {  // src/node/value.rs:14-26
            Image im;
            im.p[0] = pixel_buffer(node.value);
            out.push_back(SlotData{node.node_id, 0, im});
            break;
        }
        case KO_MIX: {  // src/node/mix.rs:51-134
            const SlotData* l = with_slot(sd, 0);
            const SlotData* r = with_slot(sd, 1);
            Image il, ir;
            if (l) {
                il = l->image;
                ir = r ? as_type(r->image, il.rgba) : image_from_value(il.w(), il.h(), 0.0f, il.rgba);
            } else if (r) {
                ir = r->image;
                il = image_from_value(ir.w(), ir.h(), 0.0f, ir.rgba);
            } else {
                out.push_back(SlotData{node.node_id, 0, image_from_value(1, 1, 0.0f, false)});
                break;
            }
            uint32_t w = il.w(), h = il.h();
            size_t n = (size_t)w * h;
            Image res;
            res.rgba = il.rgba;
            int np = il.rgba ? 3 : 1;
            for (int c = 0; c < np; ++c) {
                auto p = std::make_shared<Plane>(w, h);
                mix_plane(node.mix_type, il.p[c]->px.data(), ir.p[c]->px.data(), n, p->px.data());
                res.p[c] = p;
            }
            if (il.rgba) res.p[3] = const_plane(w, h, 1.0f);  // :203-212
            out.push_back(SlotData{node.node_id, 0, res});
            break;
        }
        case KO_HEIGHT_TO_NORMAL: {  // src/node/height_to_normal.rs:16-77
            const SlotData* in = with_slot(sd, 0);
            if (!in || in->image.rgba) break;  // Ok(Vec::new()) => InvalidBufferCount below
            uint32_t w = in->image.w(), h = in->image.h();
            std::shared_ptr<Plane> o[3];
            for (int c = 0; c < 3; ++c) o[c] = std::make_shared<Plane>(w, h);
            height_to_normal(in->image.p[0]->px.data(), w, h, h, nullptr, o[0]->px.data(), o[1]->px.data(),
                             o[2]->px.data());
            Image res;
            res.rgba = true;
            for (int c = 0; c < 3; ++c) res.p[c] = o[c];
            res.p[3] = const_plane(w, h, 1.0f);  // from_buffers_rgb, slot_image.rs:90-102
            out.push_back(SlotData{node.node_id, 0, res});
            break;
        }
        case KO_SEPARATE_RGBA: {  // src/node/separate_rgba.rs:38-69
            if (!sd.empty() && sd[0].image.rgba) {
                for (uint32_t c = 0; c < 4; ++c) {
                    Image im;
                    im.p[0] = sd[0].image.p[c];
                    out.push_back(SlotData{node.node_id, c, im});
                }
            } else {
                for (uint32_t c = 0; c < 4; ++c) {
                    Image im;
                    im.p[0] = pixel_buffer(0.0f);
                    out.push_back(SlotData{node.node_id, c, im});
                }
            }
            break;
        }
        case KO_COMBINE_RGBA: {  // src/node/combine_rgba.rs:14-97
            uint32_t w = 1, h = 1;
            if (!sd.empty()) { w = sd[0].image.w(); h = sd[0].image.h(); }
            PlaneP zero;
            Image res;
            res.rgba = true;
            for (uint32_t c = 0; c < 4; ++c) {
                const SlotData* s = with_slot(sd, c);
                if (s) {
                    if (s->image.rgba) return KO_ERR_GENERIC;  // panic!, :25
                    res.p[c] = s->image.p[0];
                } else if (c == 3) {
                    res.p[c] = const_plane(w, h, 1.0f);
                } else {
                    if (!zero) zero = const_plane(w, h, 0.0f);
                    res.p[c] = zero;
                }
            }
            out.push_back(SlotData{node.node_id, 0, res});
            break;
        }
        default: return KO_ERR_INVALID_NODE_TYPE;
    }

    // output count check, node_type.rs:124-137
    if (node.type != KO_OUTPUT_GRAY && node.type != KO_OUTPUT_RGBA &&
        (int)out.size() != output_slot_count(node))
        return KO_ERR_INVALID_BUFFER_COUNT;
    return KO_OK;
}

// The engine (src/engine.rs:128-307, src/process_pack.rs:27-96) reduced to what
// determines results: a node runs once all its parents are Clean; up to
// max_threads nodes run at once, one OS thread per node, each node's pixel loop
// single-threaded.  (The 1 ms polling and priority pre-emption are omitted;
// this favours the CPU baseline.)
struct Sched {
    std::mutex mu;
    std::condition_variable cv;
};

int eval_graph(ko_graph* g, int max_threads) {
    const size_t N = g->nodes.size();
    g->slot_datas.clear();
    std::vector<int> state(N, 0);  // 0 dirty, 1 processing, 2 clean
    auto index_of = [&](uint32_t id) -> int {
        for (size_t i = 0; i < N; ++i)
            if (g->nodes[i].node_id == id) return (int)i;
        return -1;
    };
    for (const Edge& e : g->edges)
        if (index_of(e.output_id) < 0 || index_of(e.input_id) < 0) return KO_ERR_INVALID_NODE_ID;

    if (max_threads < 1) max_threads = 1;
    Sched s;
    int running = 0, err = KO_OK;
    size_t done = 0;
    std::vector<std::thread> threads;

    auto ready = [&](size_t i) {
        if (state[i] != 0) return false;
        for (const Edge& e : g->edges)
            if (e.input_id == g->nodes[i].node_id && state[index_of(e.output_id)] != 2) return false;
        return true;
    };

    std::unique_lock<std::mutex> lk(s.mu);
    while (done < N && err == KO_OK) {
        bool launched = false;
        for (size_t i = 0; i < N && running < max_threads; ++i) {
            if (!ready(i)) continue;
            state[i] = 1;
            ++running;
            launched = true;
            // gather inputs in graph-edge order, engine.rs:217-262
            std::vector<Edge> ne;
            std::vector<SlotData> in;
            bool missing = false;
            for (const Edge& e : g->edges) {
                if (e.input_id != g->nodes[i].node_id) continue;
                const SlotData* f = nullptr;
                for (const SlotData& d : g->slot_datas)
                    if (d.node_id == e.output_id && d.slot_id == e.output_slot) { f = &d; break; }
                if (!f) { missing = true; break; }
                ne.push_back(e);
                in.push_back(*f);
            }
            if (missing) { err = KO_ERR_NO_SLOT_DATA; --running; break; }
            auto work = [g, i, ne, in, &s, &state, &running, &done, &err]() {
                std::vector<SlotData> out;
                int rc = process_node(g, g->nodes[i], in, ne, out);
                std::lock_guard<std::mutex> l(s.mu);
                if (rc != KO_OK && err == KO_OK) err = rc;
                for (SlotData& d : out) g->slot_datas.push_back(std::move(d));
                state[i] = 2;
                --running;
                ++done;
                s.cv.notify_all();
            };
            if (max_threads == 1) {
                lk.unlock();
                work();
                lk.lock();
            } else {
                threads.emplace_back(work);
            }
        }
        if (err != KO_OK) break;
        if (!launched) {
            if (running == 0) { err = KO_ERR_NODE_DIRTY; break; }  // cycle: never becomes clean
            s.cv.wait(lk);
        }
    }
    while (running > 0) s.cv.wait(lk);
    lk.unlock();
    for (std::thread& t : threads) t.join();
    return err;
}

Image image_from_planes(int is_rgba, uint32_t w, uint32_t h, const float* const* planes) {
    Image im;
    im.rgba = is_rgba != 0;
    int np = is_rgba ? 4 : 1;
    for (int c = 0; c < np; ++c) {
        auto p = std::make_shared<Plane>(w, h);
        std::memcpy(p->px.data(), planes[c], (size_t)w * h * sizeof(float));
        im.p[c] = p;
    }
    return im;
}

}  // namespace

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
extern "C" {

void ko_deconstruct_u8(const uint8_t* s, uint32_t w, uint32_t h, uint32_t ch, float* r, float* g,
                       float* b, float* a) {
    // src/shared.rs:16-56: sample/255, channels dealt round-robin; absent colour
    // channels are 0, absent alpha is 1.
    float* pl[4] = {r, g, b, a};
    size_t n = (size_t)w * h;
    for (size_t i = 0; i < n; ++i)
        for (uint32_t c = 0; c < ch && c < 4; ++c) pl[c][i] = (float)s[i * ch + c] / 255.0f;
    for (uint32_t c = ch; c < 4; ++c)
        for (size_t i = 0; i < n; ++i) pl[c][i] = (c == 3) ? 1.0f : 0.0f;
}

void ko_to_u8(const float* r, const float* g, const float* b, const float* a, uint64_t n, int srgb,
              uint8_t* out) {
    if (!g) {  // Gray => [v,v,v,255]
        for (uint64_t i = 0; i < n; ++i) {
            uint8_t v = srgb ? f32_to_u8_srgb(r[i]) : f32_to_u8(r[i]);
            out[4 * i + 0] = v; out[4 * i + 1] = v; out[4 * i + 2] = v; out[4 * i + 3] = 255;
        }
        return;
    }
    for (uint64_t i = 0; i < n; ++i) {
        out[4 * i + 0] = srgb ? f32_to_u8_srgb(r[i]) : f32_to_u8(r[i]);
        out[4 * i + 1] = srgb ? f32_to_u8_srgb(g[i]) : f32_to_u8(g[i]);
        out[4 * i + 2] = srgb ? f32_to_u8_srgb(b[i]) : f32_to_u8(b[i]);
        out[4 * i + 3] = f32_to_u8(a[i]);  // alpha stays linear, slot_image.rs:201
    }
}

void ko_mix_plane(int op, const float* l, const float* r, uint64_t n, float* out) {
    mix_plane(op, l, r, n, out);
}

void ko_rgb_to_gray(const float* r, const float* g, const float* b, uint64_t n, float* out) {
    for (uint64_t i = 0; i < n; ++i) out[i] = ((r[i] + g[i]) + b[i]) / 3.0f;
}

void ko_height_to_normal(const float* hgt, uint32_t w, uint32_t h, float* o0, float* o1, float* o2) {
    height_to_normal(hgt, w, h, h, nullptr, o0, o1, o2);
}

void ko_height_to_normal_strip(const float* strip, uint32_t w, uint32_t h, uint32_t h_full, const float* halo_row,
                               float* o0, float* o1, float* o2) {
    height_to_normal(strip, w, h, h_full, halo_row, o0, o1, o2);
}

void ko_resize_plane(const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t dw,
                     uint32_t dh, int filter) {
    resize_plane(src, sw, sh, dst, dw, dh, filter);
}

uint32_t ko_resize_weights(uint32_t src_len, uint32_t dst_len, int filter, uint32_t* left,
                           uint32_t* count, float* weights, uint32_t max_taps) {
    uint32_t mx = 0;
    for (uint32_t o = 0; o < dst_len; ++o) {
        Taps t = taps_for(o, src_len, dst_len, filter);
        mx = std::max<uint32_t>(mx, (uint32_t)t.w.size());
        if (weights) {
            left[o] = t.left;
            count[o] = (uint32_t)t.w.size();
            for (size_t i = 0; i < t.w.size() && i < max_taps; ++i) weights[(size_t)o * max_taps + i] = t.w[i];
        }
    }
    return mx;
}

ko_graph* ko_graph_new(void) { return new ko_graph(); }
void ko_graph_free(ko_graph* g) { delete g; }

int ko_graph_add_node(ko_graph* g, uint32_t node_id, int node_type, float value, int mix_type,
                      const char* name, const ko_graph* nested, uint32_t embed_id, int policy,
                      uint32_t policy_slot, uint32_t policy_w, uint32_t policy_h, int filter) {
    for (const Node& n : g->nodes)
        if (n.node_id == node_id) return KO_ERR_INVALID_NODE_ID;
    Node n;
    n.node_id = node_id;
    n.type = node_type;
    n.value = value;
    n.mix_type = mix_type;
    n.name = name ? name : "";
    if (nested) n.nested = std::make_shared<ko_graph>(*nested);
    if (node_type == KO_GRAPH && !nested) return KO_ERR_INVALID_NODE_TYPE;
    n.embed_id = embed_id;
    n.policy = policy;
    n.policy_slot = policy_slot;
    n.policy_w = policy_w;
    n.policy_h = policy_h;
    n.filter = filter;
    g->nodes.push_back(n);
    return KO_OK;
}

int ko_graph_add_edge(ko_graph* g, uint32_t output_id, uint32_t input_id, uint32_t output_slot,
                      uint32_t input_slot) {
    g->edges.push_back(Edge{output_id, input_id, output_slot, input_slot});
    return KO_OK;
}

int ko_graph_set_image_u8(ko_graph* g, uint32_t node_id, const uint8_t* s, uint32_t w, uint32_t h,
                          uint32_t ch) {
    ImageU8 u;
    u.w = w; u.h = h; u.ch = ch;
    u.px.assign(s, s + (size_t)w * h * ch);
    g->images[node_id] = std::move(u);
    return KO_OK;
}

int ko_graph_add_input_f32(ko_graph* g, uint32_t node_id, int is_rgba, uint32_t w, uint32_t h,
                           const float* const* planes) {
    g->inputs.push_back(SlotData{node_id, 0, image_from_planes(is_rgba, w, h, planes)});
    return KO_OK;
}

int ko_graph_embed_f32(ko_graph* g, uint32_t embed_id, int is_rgba, uint32_t w, uint32_t h,
                       const float* const* planes) {
    for (auto& e : g->embeds)
        if (e.first == embed_id) return KO_ERR_INVALID_SLOT_ID;  // live_graph.rs:329-340
    g->embeds.emplace_back(embed_id, image_from_planes(is_rgba, w, h, planes));
    return KO_OK;
}

int ko_graph_eval(ko_graph* g, int max_threads) { return eval_graph(g, max_threads); }

int ko_graph_slot(const ko_graph* g, uint32_t node_id, uint32_t slot_id, int* is_rgba, uint32_t* w,
                  uint32_t* h, const float** planes) {
    for (const SlotData& d : g->slot_datas) {
        if (d.node_id != node_id || d.slot_id != slot_id) continue;
        *is_rgba = d.image.rgba ? 1 : 0;
        *w = d.image.w();
        *h = d.image.h();
        int np = d.image.rgba ? 4 : 1;
        for (int c = 0; c < np; ++c) planes[c] = d.image.p[c]->px.data();
        return KO_OK;
    }
    return KO_ERR_NO_SLOT_DATA;
}

int ko_graph_slot_ids(const ko_graph* g, uint32_t node_id, uint32_t* slot_ids, int cap) {
    int n = 0;
    for (const SlotData& d : g->slot_datas) {
        if (d.node_id != node_id) continue;
        if (n < cap) slot_ids[n] = d.slot_id;
        ++n;
    }
    return n;
}

double ko_graph_eval_batch(const ko_graph* g, int copies, int max_threads) {
    // `copies` independent graphs pushed into one TextureProcessor: the engine
    // admits up to max_threads ready nodes across all of them.  Each copy is a
    // chain of dependent nodes here, so run each copy's DAG on its own thread,
    // capped at max_threads concurrent copies.
    if (copies < 1) copies = 1;
    if (max_threads < 1) max_threads = 1;
    std::vector<std::unique_ptr<ko_graph>> gs;
    for (int i = 0; i < copies; ++i) gs.emplace_back(new ko_graph(*g));
    std::mutex mu;
    int next = 0, err = 0;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<std::thread> th;
    int nt = std::min(copies, max_threads);
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&]() {
            for (;;) {
                int i;
                {
                    std::lock_guard<std::mutex> l(mu);
                    if (next >= copies) return;
                    i = next++;
                }
                int rc = eval_graph(gs[i].get(), 1);
                if (rc) { std::lock_guard<std::mutex> l(mu); err = rc; }
            }
        });
    for (std::thread& t : th) t.join();
    double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return err ? -(double)err : dt;
}

}  // extern "C"
