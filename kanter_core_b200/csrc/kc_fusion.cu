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
#include <cstring>

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

// ---- tape code generation for one kernel -------------------------------------
struct Gen {
    KcTapeArgs args{};
    std::vector<kc_plane*> srcs;   // DEVICE planes bound to S[k]
    bool tmp_used[KC_MAX_TMP] = {false, false, false, false, false, false};
    bool ok = true;
    int in_kernel_mark = 0;

    void emit(uint32_t op, uint32_t arg, float imm = 0.0f) {
        if (args.n_instr >= (uint32_t)KC_MAX_TAPE) { ok = false; return; }
        args.instr[args.n_instr] = op | (arg << 8);
        args.imm[args.n_instr] = imm;
        args.n_instr++;
    }
    int alloc_tmp() {
        for (int j = 0; j < KC_MAX_TMP; ++j)
            if (!tmp_used[j]) { tmp_used[j] = true; return j; }
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
    void operand(kc_plane* p, uint32_t& arg, float& imm) {
        imm = 0.0f;
        if (in_kernel(p)) {
            arg = KC_ARG_TMP0 + p->tmp_slot;
            consume(p);
        } else if (p->kind == KC_PLANE_CONST) {
            arg = KC_ARG_IMM;
            imm = p->value;
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
            uint32_t arg; float imm;
            operand(p, arg, imm);
            emit(TOP_LD, arg, imm);
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
        uint32_t arg; float imm;
        if (simple(b)) {
            value(a);
            operand(b, arg, imm);
            emit(fwd(x->op), arg, imm);
        } else if (simple(a)) {
            value(b);
            operand(a, arg, imm);
            emit(rev(x->op), arg, imm);
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
            operand(first, arg, imm);
            emit(a_first ? rev(x->op) : fwd(x->op), arg, imm);
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

struct KernelPlan {
    std::vector<kc_plane*> nodes;  // in topo order
    std::vector<kc_plane*> outs;
};

// Build + launch one kernel.  pack: 0 none, 1 RGBA8 from `pack_planes[4]`, 2 gray.
// Returns KC_OK, or KC_ERR_GENERIC with gen_failed=true when the group does not
// fit the machine (caller splits and retries).
int32_t run_kernel(kc_context* ctx, KernelPlan& k, int pack, kc_plane* const* pack_planes, int srgb,
                   uint32_t* d_rgba8, bool& gen_failed) {
    gen_failed = false;
    static std::atomic<int> kernel_serial{1000};
    Gen g;
    g.in_kernel_mark = ++kernel_serial;
    // in-kernel use counts and Sethi-Ullman-ish sizes
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
        if (pack == 1) g.tmp_used[0] = g.tmp_used[1] = g.tmp_used[2] = true;
    }
    const size_t n_px = (size_t)w * h;

    if ((int)k.outs.size() > KC_MAX_OUT) { gen_failed = true; return KC_ERR_GENERIC; }
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
    if (!g.ok) { gen_failed = true; return KC_ERR_GENERIC; }

    std::vector<float*> out_ptrs;
    for (kc_plane* o : k.outs) {
        kc_plane* tmp = nullptr;
        int32_t rc = kcp_new_device(ctx, o->w, o->h, &tmp);
        if (rc != KC_OK) return rc;
        out_ptrs.push_back(tmp->dptr);
        tmp->owned = false;  // storage moves into `o` below
        tmp->dptr = nullptr;
        delete tmp;
    }
    g.args.n = n_px;
    g.args.n_src = (uint32_t)g.srcs.size();
    for (size_t s = 0; s < g.srcs.size(); ++s) g.args.src[s] = g.srcs[s]->dptr;
    for (size_t m = 0; m < out_ptrs.size(); ++m) g.args.out[m] = out_ptrs[m];
    g.args.out_rgba8 = d_rgba8;
    int32_t rc = kck_launch_tape(ctx, g.args);
    if (rc != KC_OK) {
        for (float* p : out_ptrs) cudaFreeAsync(p, ctx->stream);
        return rc;
    }
    ctx->run_groups++;
    ctx->run_bytes += (uint64_t)(g.srcs.size() + out_ptrs.size()) * n_px * 4 + (pack ? n_px * 4 : 0);
    // the outputs become device planes; their operand references are dropped
    for (size_t m = 0; m < k.outs.size(); ++m) {
        kc_plane* o = k.outs[m];
        kc_plane *a = o->a, *b = o->b;
        o->kind = KC_PLANE_DEVICE;
        o->dptr = out_ptrs[m];
        o->owned = true;
        o->a = o->b = nullptr;
        kcp_release(a);
        kcp_release(b);
    }
    return KC_OK;
}

int32_t materialise_const(kc_context* ctx, kc_plane* p) {
    kc_plane* tmp = nullptr;
    KC_TRY(kcp_new_device(ctx, p->w, p->h, &tmp));
    float* d = tmp->dptr;
    tmp->owned = false;
    tmp->dptr = nullptr;
    delete tmp;
    int32_t rc = kck_fill(ctx, d, p->count(), p->value);
    if (rc != KC_OK) { cudaFreeAsync(d, ctx->stream); return rc; }
    ctx->run_bytes += p->bytes();
    p->kind = KC_PLANE_DEVICE;
    p->dptr = d;
    p->owned = true;
    return KC_OK;
}

// distinct DEVICE sources of the sub-cone rooted at each node, capped
struct Est {
    int nodes = 0;
    std::vector<kc_plane*> srcs;
};

int32_t force_impl(kc_context* ctx, kc_plane* const* roots, size_t n, int pack, int srgb, uint32_t* d_rgba8, int depth);

// Shrink cones that cannot fit one kernel by materialising an interior node.
// Returns true when something was forced (caller must re-plan).
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

int32_t force_impl(kc_context* ctx, kc_plane* const* roots, size_t n, int pack, int srgb, uint32_t* d_rgba8, int depth) {
    static std::atomic<int> mark_serial{MARK_BASE};
    if (depth > 64) KC_FAIL(KC_ERR_GENERIC, "fusion planner: recursion too deep");
    for (int attempt = 0; attempt < 100000; ++attempt) {
        Cone c;
        collect(roots, n, ++mark_serial, c);
        if (c.order.empty() && !pack) return KC_OK;
        bool changed = false;
        KC_TRY(split_oversized(ctx, c, changed, depth));
        if (changed) continue;

        // outputs: the roots, plus interior values somebody outside the cone still holds
        std::vector<kc_plane*> outs;
        if (!pack) {
            for (kc_plane* p : c.order) {
                bool root = false;
                for (size_t i = 0; i < n; ++i) root |= (roots[i] == p);
                if (root || p->refs.load() > p->uses_in_cone) outs.push_back(p);
            }
        }
        // greedy partition of the outputs (topological order) into kernels
        std::vector<KernelPlan> plans;
        {
            // membership: node -> index of the kernel that computes it (-1 none yet)
            std::map<kc_plane*, int> owner;
            auto cone_of = [&](kc_plane* o, int kidx, std::vector<kc_plane*>& add, std::vector<kc_plane*>& srcs) {
                // nodes of o's cone not materialised by an earlier kernel
                std::vector<kc_plane*> st{o};
                std::vector<kc_plane*> seen;
                while (!st.empty()) {
                    kc_plane* p = st.back();
                    st.pop_back();
                    if (std::find(seen.begin(), seen.end(), p) != seen.end()) continue;
                    seen.push_back(p);
                    auto it = owner.find(p);
                    bool earlier_out = it != owner.end() && it->second < kidx &&
                                       std::find(outs.begin(), outs.end(), p) != outs.end();
                    if (is_expr(p) && !earlier_out) {
                        if (it == owner.end() || it->second != kidx) add.push_back(p);
                        st.push_back(p->a);
                        st.push_back(p->b);
                    } else if (p->kind != KC_PLANE_CONST) {
                        if (std::find(srcs.begin(), srcs.end(), p) == srcs.end()) srcs.push_back(p);
                    }
                }
            };
            KernelPlan cur;
            std::vector<kc_plane*> cur_srcs;
            int kidx = 0;
            auto flush = [&]() {
                if (cur.outs.empty()) return;
                plans.push_back(cur);
                cur = KernelPlan();
                cur_srcs.clear();
                kidx++;
            };
            for (kc_plane* o : outs) {
                std::vector<kc_plane*> add, srcs = cur_srcs;
                cone_of(o, kidx, add, srcs);
                bool fits = (int)(cur.nodes.size() + add.size()) <= CAP_NODES && (int)srcs.size() <= KC_MAX_SRC &&
                            (int)cur.outs.size() + 1 <= KC_MAX_OUT;
                if (!fits && !cur.outs.empty()) {
                    flush();
                    add.clear();
                    srcs.clear();
                    cone_of(o, kidx, add, srcs);
                }
                for (kc_plane* p : add) { cur.nodes.push_back(p); owner[p] = kidx; }
                cur.outs.push_back(o);
                cur_srcs = srcs;
            }
            flush();
        }
        // nodes inside each plan must be in topological order
        std::map<kc_plane*, int> pos;
        for (size_t i = 0; i < c.order.size(); ++i) pos[c.order[i]] = (int)i;
        bool failed = false;
        for (KernelPlan& k : plans) {
            std::sort(k.nodes.begin(), k.nodes.end(), [&](kc_plane* a, kc_plane* b) { return pos[a] < pos[b]; });
            std::sort(k.outs.begin(), k.outs.end(), [&](kc_plane* a, kc_plane* b) { return pos[a] < pos[b]; });
            bool gen_failed = false;
            int32_t rc = run_kernel(ctx, k, 0, nullptr, 0, nullptr, gen_failed);
            if (gen_failed) {
                // register/tape pressure: materialise the deepest operand of the last output and re-plan
                kc_plane* o = k.outs.back();
                kc_plane* pick = is_expr(o->a) ? o->a : (is_expr(o->b) ? o->b : nullptr);
                if (!pick) KC_FAIL(KC_ERR_GENERIC, "fusion planner: kernel generation failed on a leaf group");
                KC_TRY(force_impl(ctx, &pick, 1, 0, 0, nullptr, depth + 1));
                failed = true;
                break;
            }
            if (rc != KC_OK) return rc;
        }
        if (failed) continue;
        if (pack) {
            // everything the export needs that is still lazy goes into the export kernel itself
            Cone pc;
            collect(roots, n, ++mark_serial, pc);
            KernelPlan k;
            k.nodes = pc.order;
            bool gen_failed = false;
            int32_t rc = run_kernel(ctx, k, pack, roots, srgb, d_rgba8, gen_failed);
            if (gen_failed) {
                // too big for one export kernel: materialise the channels first, then export
                KC_TRY(force_impl(ctx, roots, n, 0, 0, nullptr, depth + 1));
                continue;
            }
            return rc;
        }
        return KC_OK;
    }
    KC_FAIL(KC_ERR_GENERIC, "fusion planner did not converge");
}

}  // namespace

int32_t kcp_force(kc_context* ctx, kc_plane* const* roots, size_t n) {
    // constants among the roots become real planes; lazy ones are fused
    for (size_t i = 0; i < n; ++i)
        if (roots[i]->kind == KC_PLANE_CONST) KC_TRY(materialise_const(ctx, roots[i]));
    return force_impl(ctx, roots, n, 0, 0, nullptr, 0);
}

int32_t kcp_export_rgba8(kc_context* ctx, const kc_image* img, int srgb, uint32_t* d_out) {
    // SlotImage::to_u8 / to_u8_srgb (src/slot_image.rs:142-207) as the tail of
    // whatever still has to be computed for this image.
    const int np = kci_nplanes(img);
    kc_plane* roots[4] = {img->planes[0], img->planes[1], img->planes[2], img->planes[3]};
    return force_impl(ctx, roots, (size_t)np, np == 4 ? 1 : 2, srgb, d_out, 0);
}
