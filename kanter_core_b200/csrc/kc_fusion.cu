// kc_fusion.cu — the elementwise fusion planner.
//
// The reference evaluates a graph node by node, each Mix/as_type materialising
// whole planes (src/node/mix.rs:136-302, src/slot_image.rs:212-256), scheduled
// by a polling engine (src/engine.rs:128-307).  Here a per-pixel node only
// records a lazy expression plane (KC_PLANE_EXPR).  When pixels are really
// needed (a stencil/resize input, a requested node, a download, the RGBA8
// export) kcp_force() collects the expression DAG feeding the wanted planes,
// cuts it into groups that fit the tape kernel's limits, compiles each group to
// an op tape for the accumulator machine of kc_kernels.cu and launches ONE
// kernel per group: sources are read once, results written once, every
// intermediate lives in registers.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iterator>

#include "kc_internal.h"

float kc_host_mix(int op, float l, float r) {
    // constant folding uses the host libm: Rust's f32::powf is glibc powf too
    switch (op) {
        case KC_MIX_ADD: return l + r;
        case KC_MIX_SUBTRACT: return l - r;
        case KC_MIX_MULTIPLY: return l * r;
        case KC_MIX_DIVIDE: return l / r;
        default: return powf(l, r);
    }
}

namespace {

constexpr int CAP_NODES = 20;   // expression nodes per kernel
constexpr int MARK_BASE = 1;

struct Cone {
    std::vector<kc_plane*> order;  // EXPR nodes, operands before users
};

inline bool is_expr(const kc_plane* p) { return p->kind == KC_PLANE_EXPR; }

// iterative post-order DFS over EXPR operands
void collect(kc_plane* const* roots, size_t n, int mark, Cone& c) {
    struct Frame { kc_plane* p; int stage; };
    std::vector<Frame> st;
    for (size_t i = 0; i < n; ++i) {
        if (!is_expr(roots[i]) || roots[i]->mark == mark) continue;
        st.push_back({roots[i], 0});
        roots[i]->mark = mark;
        roots[i]->uses_in_cone = 0;
        while (!st.empty()) {
            Frame& f = st.back();
            kc_plane* p = f.p;
            if (f.stage < 2) {
                kc_plane* ch = f.stage == 0 ? p->a : p->b;
                f.stage++;
                if (is_expr(ch)) {
                    if (ch->mark != mark) {
                        ch->mark = mark;
                        ch->uses_in_cone = 0;
                        st.push_back({ch, 0});
                    }
                }
            } else {
                c.order.push_back(p);
                st.pop_back();
            }
        }
    }
    for (kc_plane* p : c.order) {
        if (is_expr(p->a)) p->a->uses_in_cone++;
        if (is_expr(p->b)) p->b->uses_in_cone++;
    }
}

// ---- tape code generation for one segment ------------------------------------
struct Gen {
    std::vector<uint32_t> instr;
    std::vector<float> imm;
    std::vector<kc_plane*> srcs;   // DEVICE planes bound to S[k]
    bool tmp_used[KC_MAX_TMP] = {false, false, false, false, false, false};
    int max_tmp = 0;               // temporaries the tape touches (0..KC_MAX_TMP)
    bool ok = true;
    int in_kernel_mark = 0;

    void emit(uint32_t op, uint32_t arg, float v = 0.0f) {
        if ((int)instr.size() >= KC_MAX_TAPE) { ok = false; return; }
        instr.push_back(op | (arg << 8));
        imm.push_back(v);
    }
    void touch_tmp(int j) { max_tmp = std::max(max_tmp, j + 1); }
    int alloc_tmp() {
        for (int j = 0; j < KC_MAX_TMP; ++j)
            if (!tmp_used[j]) { tmp_used[j] = true; touch_tmp(j); return j; }
        ok = false;
        return 0;
    }
    int src_slot(kc_plane* p) {
        for (size_t k = 0; k < srcs.size(); ++k)
            if (srcs[k] == p) return (int)k;
        if ((int)srcs.size() >= KC_MAX_SRC) { ok = false; return 0; }
        srcs.push_back(p);
        return (int)srcs.size() - 1;
    }
    bool in_kernel(const kc_plane* p) const { return is_expr(p) && p->ktag == in_kernel_mark; }
    // a value that can be named as an instruction operand right now
    bool simple(const kc_plane* p) const { return !in_kernel(p) || p->tmp_slot >= 0; }

    // operand reference; consumes one use of p
    void operand(kc_plane* p, uint32_t& arg, float& v) {
        v = 0.0f;
        if (in_kernel(p)) {
            arg = KC_ARG_TMP0 + p->tmp_slot;
            consume(p);
        } else if (p->kind == KC_PLANE_CONST) {
            arg = KC_ARG_IMM;
            v = p->value;
        } else {
            arg = (uint32_t)src_slot(p);
        }
    }
    void consume(kc_plane* p) {
        if (!in_kernel(p)) return;
        if (--p->remaining == 0 && p->tmp_slot >= 0) {
            tmp_used[p->tmp_slot] = false;
            p->tmp_slot = -1;
        }
    }
    // leave p's value in acc (consumes one use of p)
    void value(kc_plane* p) {
        if (simple(p)) {
            uint32_t arg; float v;
            operand(p, arg, v);
            emit(TOP_LD, arg, v);
        } else {
            compute(p);
            consume(p);
        }
    }
    static uint32_t fwd(int mix) {
        switch (mix) {
            case KC_MIX_ADD: return TOP_ADD;
            case KC_MIX_SUBTRACT: return TOP_SUB;
            case KC_MIX_MULTIPLY: return TOP_MUL;
            case KC_MIX_DIVIDE: return TOP_DIV;
            default: return TOP_POW;
        }
    }
    static uint32_t rev(int mix) {  // acc holds the RIGHT operand
        switch (mix) {
            case KC_MIX_ADD: return TOP_ADD;
            case KC_MIX_SUBTRACT: return TOP_RSUB;
            case KC_MIX_MULTIPLY: return TOP_MUL;
            case KC_MIX_DIVIDE: return TOP_RDIV;
            default: return TOP_RPOW;
        }
    }
    // compute an in-kernel EXPR node into acc; saves it to a temp when it has
    // further uses, stores it when it is a kernel output.  Does NOT consume.
    void compute(kc_plane* x) {
        if (!ok) return;
        kc_plane *a = x->a, *b = x->b;
        uint32_t arg; float v;
        if (simple(b)) {
            value(a);
            operand(b, arg, v);
            emit(fwd(x->op), arg, v);
        } else if (simple(a)) {
            value(b);
            operand(a, arg, v);
            emit(rev(x->op), arg, v);
        } else {
            // both sides still need computing: the bigger cone first, parked in a temp
            const bool a_first = a->need >= b->need;
            kc_plane* first = a_first ? a : b;
            kc_plane* second = a_first ? b : a;
            compute(first);
            if (first->tmp_slot < 0) {  // not already saved as a multi-use value
                int t = alloc_tmp();
                emit(TOP_ST_TMP, (uint32_t)t);
                first->tmp_slot = t;
            }
            value(second);
            operand(first, arg, v);
            emit(a_first ? rev(x->op) : fwd(x->op), arg, v);
        }
        if (x->is_out) emit(TOP_ST_OUT, (uint32_t)x->out_slot);
        x->computed = true;
        if (x->remaining > 1 && x->tmp_slot < 0) {
            int t = alloc_tmp();
            emit(TOP_ST_TMP, (uint32_t)t);
            x->tmp_slot = t;
        }
    }
};

// One segment: a set of outputs whose expression cones share nodes or sources.
struct SegPlan {
    std::vector<kc_plane*> nodes;  // EXPR nodes, topological order
    std::vector<kc_plane*> outs;   // EXPR outputs (become DEVICE planes) or one CONST plane to fill
    bool is_const_fill = false;
    // filled by gen_segment
    Gen gen;
    size_t n_px = 0;
    std::vector<float*> out_ptrs;
};

// pack: 0 none, 1 RGBA8 from pack_planes[4], 2 gray from pack_planes[0]
bool gen_segment(SegPlan& k, int pack, kc_plane* const* pack_planes, int srgb) {
    static std::atomic<int> kernel_serial{1000};
    Gen& g = k.gen;
    g.in_kernel_mark = ++kernel_serial;
    if (k.is_const_fill) {
        kc_plane* c = k.outs[0];
        k.n_px = c->count();
        g.emit(TOP_LD, KC_ARG_IMM, c->value);
        g.emit(TOP_ST_OUT, 0);
        return g.ok;
    }
    for (kc_plane* p : k.nodes) {
        p->ktag = g.in_kernel_mark;
        p->remaining = 0;
        p->tmp_slot = -1;
        p->out_slot = -1;
        p->is_out = false;
        p->computed = false;
    }
    for (kc_plane* p : k.nodes) {
        if (g.in_kernel(p->a)) p->a->remaining++;
        if (g.in_kernel(p->b)) p->b->remaining++;
        int na = g.in_kernel(p->a) ? p->a->need : 0;
        int nb = g.in_kernel(p->b) ? p->b->need : 0;
        p->need = 1 + na + nb;
    }
    uint32_t w = 0, h = 0;
    if (!k.nodes.empty()) { w = k.nodes[0]->w; h = k.nodes[0]->h; }
    if (pack) {
        w = pack_planes[0]->w; h = pack_planes[0]->h;
        for (int c = 0; c < (pack == 1 ? 4 : 1); ++c)
            if (g.in_kernel(pack_planes[c])) pack_planes[c]->remaining++;
        if (pack == 1) { g.tmp_used[0] = g.tmp_used[1] = g.tmp_used[2] = true; g.touch_tmp(2); }
    }
    k.n_px = (size_t)w * h;
    if ((int)k.outs.size() > KC_MAX_OUT) return false;
    for (size_t m = 0; m < k.outs.size(); ++m) {
        k.outs[m]->is_out = true;
        k.outs[m]->out_slot = (int)m;
    }
    // outputs in topological order; one that is interior to a later output's
    // expression stays parked in a temporary until its last use
    for (kc_plane* o : k.outs) {
        if (o->computed) continue;
        o->remaining++;  // the root itself counts as a use while it is computed
        g.compute(o);
        g.consume(o);
    }
    if (pack == 1) {
        for (int c = 0; c < 3; ++c) {
            g.value(pack_planes[c]);
            g.emit(TOP_ST_TMP, (uint32_t)c);
        }
        g.value(pack_planes[3]);
        g.emit(TOP_PACK_RGBA, srgb ? 1u : 0u);
    } else if (pack == 2) {
        g.value(pack_planes[0]);
        g.emit(TOP_PACK_GRAY, srgb ? 1u : 0u);
    }
    return g.ok;
}

int32_t alloc_plane_storage(kc_context* ctx, kc_plane* like, float** out) {
    kc_plane* tmp = nullptr;
    KC_TRY(kcp_new_device(ctx, like->w, like->h, &tmp));
    *out = kcp_take_storage(tmp);  // the storage is adopted by the caller (kcp_adopt_storage)
    delete tmp;
    return KC_OK;
}

// Launch a set of generated, mutually independent segments: grouped by pixel
// count, up to KC_MAX_SEG per launch.
int32_t launch_segments(kc_context* ctx, std::vector<SegPlan*>& segs, uint32_t* d_rgba8) {
    KcHostTimer hp(KC_HP_LAUNCH_SEGMENTS);
    KcPin pin;   // allocating the outputs may push the spill queue over its threshold: the sources stay
    for (SegPlan* k : segs)
        for (kc_plane* sp : k->gen.srcs) { pin.add(sp); kcp_touch(sp); }
    for (SegPlan* k : segs) {
        for (kc_plane* o : k->outs) {
            float* d = nullptr;
            KC_TRY(alloc_plane_storage(ctx, o, &d));
            k->out_ptrs.push_back(d);
        }
    }
    std::vector<bool> done(segs.size(), false);
    for (size_t i = 0; i < segs.size(); ++i) {
        if (done[i]) continue;
        KcTapeArgs args;
        memset(&args, 0, sizeof args);
        args.n = segs[i]->n_px;
        args.variant = 0;  // becomes the number of temporaries the launch needs
        uint32_t pc = 0;
        for (size_t j = i; j < segs.size() && args.n_seg < (uint32_t)KC_MAX_SEG; ++j) {
            SegPlan* k = segs[j];
            if (done[j] || k->n_px != args.n) continue;
            if (pc + k->gen.instr.size() > (size_t)KC_MAX_TAPE) continue;
            KcSegment& sg = args.seg[args.n_seg++];
            sg.tape_begin = pc;
            for (size_t t = 0; t < k->gen.instr.size(); ++t) {
                args.instr[pc] = k->gen.instr[t];
                args.imm[pc] = k->gen.imm[t];
                ++pc;
            }
            sg.tape_end = pc;
            args.variant = std::max<uint32_t>(args.variant, (uint32_t)k->gen.max_tmp);
            sg.n_src = (uint32_t)k->gen.srcs.size();
            for (size_t q = 0; q < k->gen.srcs.size(); ++q) sg.src[q] = k->gen.srcs[q]->dptr;
            for (size_t m = 0; m < k->out_ptrs.size(); ++m) sg.out[m] = k->out_ptrs[m];
            sg.out_rgba8 = d_rgba8;
            done[j] = true;
            ctx->run_groups++;
            ctx->run_bytes += (uint64_t)(k->gen.srcs.size() + k->out_ptrs.size()) * k->n_px * 4 + (d_rgba8 ? k->n_px * 4 : 0);
        }
        static const bool trace = getenv("KC_TRACE_FUSION") != nullptr;
        if (trace) {
            fprintf(stderr, "[kc fusion] launch n=%zu segments=%u tape=%u tmps=%u:", (size_t)args.n, args.n_seg, pc, args.variant);
            for (uint32_t q = 0; q < args.n_seg; ++q) {
                int outs = 0;
                for (int m = 0; m < KC_MAX_OUT; ++m) outs += args.seg[q].out[m] != nullptr;
                fprintf(stderr, " [src=%u out=%d ops=%u%s]", args.seg[q].n_src, outs, args.seg[q].tape_end - args.seg[q].tape_begin,
                        args.seg[q].out_rgba8 ? " rgba8" : "");
            }
            fprintf(stderr, "\n");
        }
        int32_t rc = kck_launch_tape(ctx, args);
        if (rc != KC_OK) return rc;
    }
    // the outputs become device planes; their operand references are dropped
    for (SegPlan* k : segs) {
        for (size_t m = 0; m < k->outs.size(); ++m) {
            kc_plane* o = k->outs[m];
            kc_plane *a = o->a, *b = o->b;
            kcp_adopt_storage(o, k->out_ptrs[m]);
            o->a = o->b = nullptr;
            if (a) kcp_release(a);
            if (b) kcp_release(b);
        }
    }
    return KC_OK;
}

// distinct DEVICE sources and node count of the sub-cone rooted at each node
struct Est {
    int nodes = 0;
    std::vector<kc_plane*> srcs;
};

int32_t force_impl(kc_context* ctx, kc_plane* const* roots, size_t n, int pack, int srgb, uint32_t* d_rgba8, int depth);

// Shrink cones that cannot fit one kernel by materialising an interior node.
int32_t split_oversized(kc_context* ctx, Cone& c, bool& changed, int depth) {
    changed = false;
    std::map<kc_plane*, Est> est;
    for (kc_plane* p : c.order) {
        Est e;
        e.nodes = 1;
        kc_plane* ops[2] = {p->a, p->b};
        for (kc_plane* o : ops) {
            if (is_expr(o)) {
                const Est& s = est[o];
                e.nodes += s.nodes;
                for (kc_plane* q : s.srcs)
                    if (std::find(e.srcs.begin(), e.srcs.end(), q) == e.srcs.end()) e.srcs.push_back(q);
            } else if (o->kind == KC_PLANE_DEVICE) {
                if (std::find(e.srcs.begin(), e.srcs.end(), o) == e.srcs.end()) e.srcs.push_back(o);
            }
        }
        if (e.nodes > CAP_NODES || (int)e.srcs.size() > KC_MAX_SRC) {
            // p is the first (deepest) violator: its operands fit; materialise the larger one
            kc_plane* pick = nullptr;
            int best = -1;
            for (kc_plane* o : ops)
                if (is_expr(o) && est[o].nodes > best) { best = est[o].nodes; pick = o; }
            if (!pick) KC_FAIL(KC_ERR_GENERIC, "fusion planner: unsplittable node");
            KC_TRY(force_impl(ctx, &pick, 1, 0, 0, nullptr, depth + 1));
            changed = true;
            return KC_OK;
        }
        est[p] = std::move(e);
    }
    return KC_OK;
}

// all EXPR nodes and DEVICE sources below `o`
void cone_of(kc_plane* o, std::vector<kc_plane*>& nodes, std::vector<kc_plane*>& srcs) {
    std::vector<kc_plane*> st{o};
    while (!st.empty()) {
        kc_plane* p = st.back();
        st.pop_back();
        if (is_expr(p)) {
            if (std::find(nodes.begin(), nodes.end(), p) != nodes.end()) continue;
            nodes.push_back(p);
            st.push_back(p->a);
            st.push_back(p->b);
        } else if (p->kind == KC_PLANE_DEVICE) {
            if (std::find(srcs.begin(), srcs.end(), p) == srcs.end()) srcs.push_back(p);
        }
    }
}

// With a memory threshold in force: bring every spilled plane under the roots back into HBM and pin
// every pixel-holding leaf, so that nothing this evaluation still has to read can be pushed out by
// the allocations it makes (the pins go away with the caller's KcPin).
int32_t reload_and_pin_leaves(kc_context* ctx, kc_plane* const* roots, size_t n, KcPin& pin) {
    std::vector<kc_plane*> st(std::make_reverse_iterator(roots + n), std::make_reverse_iterator(roots)), seen;   // roots are visited (and stamped) in order
    while (!st.empty()) {
        kc_plane* p = st.back();
        st.pop_back();
        if (!p || std::find(seen.begin(), seen.end(), p) != seen.end()) continue;
        seen.push_back(p);
        if (p->kind == KC_PLANE_DEVICE) {
            pin.add(p);
            kcp_touch(p);
        } else if (p->kind == KC_PLANE_SPILLED) {
            pin.add(p);                 // pinned first: the reload's own threshold pass must not pick it
            KC_TRY(kcp_reload(ctx, p));
        } else if (p->kind == KC_PLANE_EXPR) {
            st.push_back(p->a);
            st.push_back(p->b);
        }
    }
    return KC_OK;
}

int32_t force_impl(kc_context* ctx, kc_plane* const* roots, size_t n, int pack, int srgb, uint32_t* d_rgba8, int depth) {
    static std::atomic<int> mark_serial{MARK_BASE};
    if (depth > 64) KC_FAIL(KC_ERR_GENERIC, "fusion planner: recursion too deep");
    KcPin leaves;
    const bool spilling = ctx->memory_threshold != UINT64_MAX || ctx->planes_on_host != 0;
    for (int round = 0; round < 100000; ++round) {
        // every round: planes materialised by the previous one are leaves now
        if (spilling) KC_TRY(reload_and_pin_leaves(ctx, roots, n, leaves));
        Cone c;
        collect(roots, n, ++mark_serial, c);
        bool changed = false;
        KC_TRY(split_oversized(ctx, c, changed, depth));
        if (changed) continue;
        std::map<kc_plane*, int> pos;
        for (size_t i = 0; i < c.order.size(); ++i) pos[c.order[i]] = (int)i;

        if (pack) {
            // the export kernel computes whatever is still lazy itself; nothing is stored as f32
            SegPlan k;
            k.nodes = c.order;
            if (!gen_segment(k, pack, roots, srgb)) {
                // too big for one kernel: materialise the channels first, then export them
                KC_TRY(force_impl(ctx, roots, n, 0, 0, nullptr, depth + 1));
                continue;
            }
            std::vector<SegPlan*> one{&k};
            return launch_segments(ctx, one, d_rgba8);
        }

        // outputs: lazy roots, plus interior values somebody outside the cone still holds;
        // constant roots become fill segments of the same launch
        std::vector<kc_plane*> outs;
        for (kc_plane* p : c.order) {
            bool root = false;
            for (size_t i = 0; i < n; ++i) root |= (roots[i] == p);
            if (root || p->refs.load() > p->uses_in_cone) outs.push_back(p);
        }
        std::vector<kc_plane*> const_roots;
        for (size_t i = 0; i < n; ++i)
            if (roots[i]->kind == KC_PLANE_CONST && std::find(const_roots.begin(), const_roots.end(), roots[i]) == const_roots.end())
                const_roots.push_back(roots[i]);
        if (outs.empty() && const_roots.empty()) return KC_OK;

        // components: outputs whose cones share an expression node always go together (else the
        // node would be computed twice); outputs that only share SOURCE planes are merged -- the
        // shared planes are then read once -- while the merged kernel keeps few enough sources
        // for the big-tile configurations of the tile VM (src_soft_cap, see pick_config)
        const int soft_cap = g_kc_tuning.src_soft_cap > 0 ? std::min(g_kc_tuning.src_soft_cap, KC_MAX_SRC) : KC_SRC_SOFT_CAP;
        const size_t no = outs.size();
        std::vector<std::vector<kc_plane*>> cn(no), cs(no);
        for (size_t i = 0; i < no; ++i) cone_of(outs[i], cn[i], cs[i]);
        std::vector<int> comp(no);
        for (size_t i = 0; i < no; ++i) comp[i] = (int)i;
        auto find = [&](int x) { while (comp[x] != x) x = comp[x] = comp[comp[x]]; return x; };
        for (size_t i = 0; i < no; ++i)
            for (size_t j = i + 1; j < no; ++j) {
                if (find((int)i) == find((int)j)) continue;
                bool share = false;
                for (kc_plane* p : cn[i]) if (std::find(cn[j].begin(), cn[j].end(), p) != cn[j].end()) { share = true; break; }
                if (share && outs[i]->w == outs[j]->w && outs[i]->h == outs[j]->h) comp[find((int)j)] = find((int)i);
            }
        // second pass: source sharing, bounded by the soft cap on the union of sources
        {
            auto comp_srcs = [&](int root) {
                std::vector<kc_plane*> u;
                for (size_t i = 0; i < no; ++i)
                    if (find((int)i) == root)
                        for (kc_plane* p : cs[i]) if (std::find(u.begin(), u.end(), p) == u.end()) u.push_back(p);
                return u;
            };
            for (size_t i = 0; i < no; ++i)
                for (size_t j = i + 1; j < no; ++j) {
                    const int ri = find((int)i), rj = find((int)j);
                    if (ri == rj || outs[i]->w != outs[j]->w || outs[i]->h != outs[j]->h) continue;
                    std::vector<kc_plane*> ui = comp_srcs(ri), uj = comp_srcs(rj);
                    size_t shared = 0;
                    for (kc_plane* p : uj) shared += std::find(ui.begin(), ui.end(), p) != ui.end();
                    if (shared == 0) continue;
                    if ((int)(ui.size() + uj.size() - shared) <= soft_cap) comp[rj] = ri;
                }
        }
        // one segment per component: the longest prefix of its outputs (topological
        // order) that fits the machine; the rest waits for the next round
        std::vector<std::unique_ptr<SegPlan>> plans;
        for (size_t r = 0; r < no; ++r) {
            if (find((int)r) != (int)r) continue;
            auto k = std::make_unique<SegPlan>();
            std::vector<kc_plane*> srcs;
            for (size_t i = 0; i < no; ++i) {
                if (find((int)i) != (int)r) continue;
                std::vector<kc_plane*> nodes = k->nodes, s2 = srcs;
                for (kc_plane* p : cn[i]) if (std::find(nodes.begin(), nodes.end(), p) == nodes.end()) nodes.push_back(p);
                for (kc_plane* p : cs[i]) if (std::find(s2.begin(), s2.end(), p) == s2.end()) s2.push_back(p);
                const bool fits = (int)nodes.size() <= CAP_NODES && (int)s2.size() <= KC_MAX_SRC && (int)k->outs.size() + 1 <= KC_MAX_OUT;
                if (!fits && !k->outs.empty()) break;
                k->nodes.swap(nodes);
                srcs.swap(s2);
                k->outs.push_back(outs[i]);
            }
            std::sort(k->nodes.begin(), k->nodes.end(), [&](kc_plane* a, kc_plane* b) { return pos[a] < pos[b]; });
            plans.push_back(std::move(k));
        }
        bool failed = false;
        for (auto& k : plans) {
            if (gen_segment(*k, 0, nullptr, 0)) continue;
            // register / tape pressure: materialise an operand of the last output and re-plan
            kc_plane* o = k->outs.back();
            kc_plane* pick = is_expr(o->a) ? o->a : (is_expr(o->b) ? o->b : nullptr);
            if (!pick) KC_FAIL(KC_ERR_GENERIC, "fusion planner: kernel generation failed on a leaf group");
            KC_TRY(force_impl(ctx, &pick, 1, 0, 0, nullptr, depth + 1));
            failed = true;
            break;
        }
        if (failed) continue;
        for (kc_plane* cr : const_roots) {
            auto k = std::make_unique<SegPlan>();
            k->is_const_fill = true;
            k->outs.push_back(cr);
            gen_segment(*k, 0, nullptr, 0);
            plans.push_back(std::move(k));
        }
        std::vector<SegPlan*> segs;
        for (auto& k : plans) segs.push_back(k.get());
        KC_TRY(launch_segments(ctx, segs, nullptr));
        // anything that did not fit this round is picked up by the next one
    }
    KC_FAIL(KC_ERR_GENERIC, "fusion planner did not converge");
}

}  // namespace

int32_t kcp_prefetch_leaves(kc_context* ctx, kc_plane* const* roots, size_t n) {
    // start the uploads of every host-resident plane under the roots now (they are reloaded, not pinned)
    if (ctx->planes_on_host == 0) return KC_OK;
    KcPin scratch;
    return reload_and_pin_leaves(ctx, roots, n, scratch);
}

int32_t kcp_force(kc_context* ctx, kc_plane* const* roots, size_t n) {
    KcHostTimer hp(KC_HP_FORCE);
    KC_TRY(kc_lanes_join(ctx));        // whoever asks for pixels on the compute stream comes after the lanes of a concurrent section
    int32_t rc = force_impl(ctx, roots, n, 0, 0, nullptr, 0);
    // the evaluation held its operands in HBM (pins); now that they are released the queue may settle --
    // except for the planes the caller asked for: it is about to read them
    if (rc == KC_OK && ctx->bytes_live > ctx->memory_threshold) {
        KcPin keep;
        for (size_t i = 0; i < n; ++i) keep.add(roots[i]);
        rc = kc_enforce_threshold(ctx);
    }
    return rc;
}

int32_t kcp_export_rgba8(kc_context* ctx, const kc_image* img, int srgb, uint32_t* d_out) {
    // SlotImage::to_u8 / to_u8_srgb (src/slot_image.rs:142-207) as the tail of
    // whatever still has to be computed for this image.
    const int np = kci_nplanes(img);
    kc_plane* roots[4] = {img->planes[0], img->planes[1], img->planes[2], img->planes[3]};
    int32_t rc = force_impl(ctx, roots, (size_t)np, np == 4 ? 1 : 2, srgb, d_out, 0);
    if (rc == KC_OK && ctx->bytes_live > ctx->memory_threshold) rc = kc_enforce_threshold(ctx);
    return rc;
}
