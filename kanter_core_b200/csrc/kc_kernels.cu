// kc_kernels.cu — the elementwise device kernels and their launch code.
//
//  * kc_tile_vm_kernel: ONE kernel for every chain/tree of per-pixel node work:
//    Mix add/subtract/multiply/divide/pow (src/node/mix.rs:136-302), the
//    Rgba->Gray average of SlotImage::as_type (src/slot_image.rs:242-253),
//    constant fills (SlotImage::from_value, :28-64) and the f32->RGBA8 export of
//    SlotImage::to_u8 / to_u8_srgb (:142-207).  Persistent CTAs walk (segment, tile)
//    work items -- up to four independent segments per launch, e.g. the R, G and B
//    chains of an Rgba graph -- with the source tiles arriving by TMA bulk copies;
//    a fused group reads each source plane once and writes each result plane once,
//    intermediates never touch HBM.  The device code lives in kc_tile_vm.cuh (shared
//    with the NVRTC-specialised variant, kc_jit.cu); here the body is instantiated
//    with the tape interpreter, and kck_launch_tape picks the configuration and the
//    variant.
//  * kc_from_u8_kernel: deconstruct_image's u8/255 de-interleave (src/shared.rs:16-56).
//  (constant planes somebody insists on seeing as pixels are a two-instruction fill segment of the tape kernel: no kernel of their own)
#include <cstdlib>

#include "kc_internal.h"

namespace {

#include "kc_tile_vm.cuh"

template <bool EXACT, int V, int MINB>
__global__ void __launch_bounds__(TVM_THREADS, MINB)
    kc_tile_vm_kernel(const __grid_constant__ KcTapeArgs A, int stages, int ns_max, uint32_t tiles_per_plane, uint32_t total_work) {
    kc_tile_vm_body<EXACT, V, KcInterp>(A, stages, ns_max, tiles_per_plane, total_work);
}

// byte / 255.0f, correctly rounded, without the generic div.rn expansion: q = b*y, q += (b - 255 q)*y
// with y = RN(1/255).  All 256 inputs are checked against the oracle's `as f32 / 255.` by
// tests/test_gpu_ops.py::test_u8_roundtrip_every_value.
// The integer -> float step is the 2^23 trick: byte k of the word is dropped into the mantissa of 8388608.0f by ONE
// byte permute and the bias subtracted (both full-rate; I2F.U8 runs on the quarter-rate conversion unit and was where
// 74 % of the kernel's stall samples sat, profiles/ncu_full_from_u8_r02.txt).
template <int K>
__device__ __forceinline__ float kc_u8_over_255(uint32_t word) {
    const float a = __fsub_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, 0x7650u | K)), 8388608.0f);   // exact: 0..255
    const float y = 0x1.010102p-8f;
    const float q = __fmul_rn(a, y);
    return __fmaf_rn(__fmaf_rn(-255.0f, q, a), y, q);
}

// deconstruct_image, src/shared.rs:27-33: plane_c[i] = samples[i*C + c] as f32 / 255.
// One thread converts 4 consecutive pixels per item (4*C bytes in, one float4 per channel plane out) and keeps
// FU8_ITEMS items in flight, their loads issued together: the kernel is a pure stream and lives on bytes in flight.
constexpr int FU8_ITEMS = 2;
template <int C>
__global__ void __launch_bounds__(256, 6) kc_from_u8_kernel(const uint8_t* __restrict__ s, size_t n,
                                                          float* __restrict__ p0, float* __restrict__ p1,
                                                          float* __restrict__ p2, float* __restrict__ p3) {
    const size_t nvec = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float* planes[4] = {p0, p1, p2, p3};
    for (size_t v0 = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v0 < nvec; v0 += stride * FU8_ITEMS) {
        uint32_t wds[FU8_ITEMS][C];
#pragma unroll
        for (int it = 0; it < FU8_ITEMS; ++it) {
            const size_t v = v0 + it * stride;
            if (v < nvec) {
                // 4*C bytes starting at a 4-byte aligned offset: one 128-bit load when C == 4, C 32-bit loads otherwise
                if (C == 4) {
                    const uint4 q = __ldcs(reinterpret_cast<const uint4*>(s) + v);
                    wds[it][0] = q.x; wds[it][1 % C] = q.y; wds[it][2 % C] = q.z; wds[it][3 % C] = q.w;
                } else {
                    const uint32_t* src = reinterpret_cast<const uint32_t*>(s + 4 * C * v);
#pragma unroll
                    for (int k = 0; k < C; ++k) wds[it][k] = __ldcs(src + k);
                }
            }
        }
#pragma unroll
        for (int it = 0; it < FU8_ITEMS; ++it) {
            const size_t v = v0 + it * stride;
            if (v >= nvec) break;
            float vals[4][C];
#pragma unroll
            for (int b = 0; b < 4 * C; ++b) {
                const uint32_t w = wds[it][b >> 2];
                float f;
                switch (b & 3) {
                    case 0: f = kc_u8_over_255<0>(w); break;
                    case 1: f = kc_u8_over_255<1>(w); break;
                    case 2: f = kc_u8_over_255<2>(w); break;
                    default: f = kc_u8_over_255<3>(w); break;
                }
                vals[b / C][b % C] = f;
            }
#pragma unroll
            for (int c = 0; c < C; ++c)
                __stcs(reinterpret_cast<float4*>(planes[c]) + v, make_float4(vals[0][c], vals[1][c], vals[2][c], vals[3][c]));
        }
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        for (size_t i = 4 * nvec; i < n; ++i)
            for (int c = 0; c < C; ++c) planes[c][i] = __fdiv_rn((float)s[i * C + c], 255.0f);
    }
}

inline int grid_for(kc_context* ctx, size_t work_items, int block, int ctas_per_sm) {
    size_t want = (work_items + block - 1) / block;
    size_t cap = (size_t)ctx->sm_count * ctas_per_sm;
    if (want < 1) want = 1;
    return (int)(want < cap ? want : cap);
}

template <bool EXACT, int V, int MINB>
int32_t launch_tile_vm(kc_context* ctx, const KcTapeArgs& a, int stages, int ns_max, int nt_max, int ctas_per_sm) {
    constexpr int TILE_PX = 1024 * V;
    const size_t smem = 128 + (size_t)(stages * ns_max + nt_max) * TILE_PX * 4;
    KC_TRY(kc_ensure_smem_attr(ctx, (const void*)kc_tile_vm_kernel<EXACT, V, MINB>, 227 * 1024));
    const uint64_t tiles = (a.n + TILE_PX - 1) / TILE_PX;
    const uint64_t total = tiles * a.n_seg;
    if (total > 0xffffffffull) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "plane too large for one launch");
    // persistent grid: one CTA per resident slot, each walks work items w, w+grid, ...
    const uint64_t grid = std::min<uint64_t>(total, (uint64_t)ctx->sm_count * ctas_per_sm);
    // resolve operands to byte offsets for this tile size (see KC_R_*)
    KcTapeArgs r = a;
    uint32_t n_instr = 0;
    for (uint32_t q = 0; q < a.n_seg; ++q) n_instr = std::max(n_instr, a.seg[q].tape_end);
    for (uint32_t pc = 0; pc < n_instr; ++pc) {
        const uint32_t op = a.instr[pc] & 0xffu, arg = (a.instr[pc] >> 8) & 0xffu;
        uint32_t w = op;
        if (op <= TOP_RPOW) {
            if (arg == (uint32_t)KC_ARG_IMM) w |= KC_R_IMM;
            else if (arg >= (uint32_t)KC_ARG_TMP0) w |= KC_R_TMP | (((arg - KC_ARG_TMP0) * (uint32_t)TILE_PX * 4u) << 8);
            else w |= (arg * (uint32_t)TILE_PX * 4u) << 8;
        } else if (op == TOP_ST_TMP) {
            w |= (arg * (uint32_t)TILE_PX * 4u) << 8;
        } else {
            w |= arg << 8;
        }
        r.instr[pc] = w;
    }
    kc_tile_vm_kernel<EXACT, V, MINB><<<(unsigned)grid, TVM_THREADS, smem, ctx->stream>>>(r, stages, ns_max, (uint32_t)tiles, (uint32_t)total);
    return KC_OK;
}

struct TvmConfig { int v, ctas, stages; };

// Tile size, CTAs per SM and pipeline depth for a launch with ns_max sources and
// nt_max shared-memory temporaries per segment.  Measured on B200 (profiles/):
// 4096-pixel tiles beat smaller ones (the dispatch is amortised over 16 pixels per
// thread), 3 resident CTAs beat 2 beat 1, and 2 stages are as good as 4.
TvmConfig pick_config(int ns_max, int nt_max, unsigned long long n, int shared_cap_kb) {
    const int force_v = g_kc_tuning.tile_v, force_stages = g_kc_tuning.stages, force_ctas = g_kc_tuning.ctas;
    static const TvmConfig order[] = {{4, 3, 0}, {4, 2, 0}, {2, 3, 0}, {2, 2, 0}, {1, 3, 0}, {1, 2, 0}, {4, 1, 0}, {2, 1, 0}, {1, 1, 0}};
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 honours the tuning overrides, pass 1 ignores them
        for (const TvmConfig& c : order) {
            if (pass == 0 && ((force_v && c.v != force_v) || (force_ctas && c.ctas != force_ctas))) continue;
            // do not use a tile (much) bigger than the plane -- nor tiles so big that a small plane is a handful of them:
            // below ~150 tiles the launch is latency-bound and every SM that gets a tile shortens it (a 256^2 plane is
            // 16 tiles of 4096 pixels but 64 of 1024; with glibc-exact pow in the tape that was 6 us of fp64 per block)
            if (c.v > 1 && n <= (unsigned long long)512 * c.v) continue;
            if (pass == 0 && !force_v && c.v > 1 && (n + 1024ull * c.v - 1) / (1024ull * c.v) < 148ull) continue;
            // inside a concurrent section (kc_context_concurrent_begin) a launch sizes itself for half an SM's shared memory,
            // so that a kernel of another lane finds room beside it: the persistent grid would otherwise hold every SM to
            // itself until its last tile (configs[4], 3 lanes: 0.345 ms per graph with the whole SM, 0.332 with half)
            const size_t cap_kb = g_kc_tuning.smem_cap_kb >= 16 && g_kc_tuning.smem_cap_kb <= 227 ? (size_t)g_kc_tuning.smem_cap_kb : shared_cap_kb > 0 ? (size_t)shared_cap_kb : 227;
            const size_t budget = cap_kb * 1024 / (size_t)c.ctas - 1024 - 128;
            const size_t tile_b = (size_t)4096 * c.v;
            for (int st = 4; st >= (force_stages == 1 ? 1 : 2); --st) {
                if (pass == 0 && force_stages && st != force_stages) continue;
                if (!force_stages && ns_max > 0 && st > 2 && c.ctas >= 2 && (size_t)(st * ns_max + nt_max) * tile_b > budget) continue;
                if ((size_t)(st * ns_max + nt_max) * tile_b <= budget) return TvmConfig{c.v, c.ctas, ns_max == 0 ? 2 : st};
            }
        }
    }
    return TvmConfig{1, 1, 2};
}

}  // namespace

int g_kc_last_tile_config[3] = {0, 0, 0};
extern "C" int32_t kc_debug_last_tile_config(int32_t* v, int32_t* ctas, int32_t* stages) try {
    if (v) *v = g_kc_last_tile_config[0];
    if (ctas) *ctas = g_kc_last_tile_config[1];
    if (stages) *stages = g_kc_last_tile_config[2];
    return KC_OK;
} KC_ABI_CATCH

int32_t kck_launch_tape(kc_context* ctx, const KcTapeArgs& args) {
    if (args.n == 0 || args.n_seg == 0) return KC_OK;
    if (ctx->capture_log) {   // what this launch touches (either kernel below is ONE launch)
        KcFootprint f;
        for (uint32_t q = 0; q < args.n_seg; ++q) {
            const KcSegment& sg = args.seg[q];
            for (uint32_t k = 0; k < sg.n_src; ++k) f.reads.push_back({sg.src[k], (size_t)args.n * 4});
            for (int m = 0; m < KC_MAX_OUT; ++m) if (sg.out[m]) f.writes.push_back({sg.out[m], (size_t)args.n * 4});
            if (sg.out_rgba8) f.writes.push_back({sg.out_rgba8, (size_t)args.n * 4});
        }
        ctx->capture_log->push_back(std::move(f));
    }
    KcHostTimer hp(KC_HP_LAUNCH_TAPE);
    int ns_max = 0;
    for (uint32_t s = 0; s < args.n_seg; ++s) ns_max = std::max<int>(ns_max, (int)args.seg[s].n_src);
    const int nt_max = (int)args.variant;  // temporaries the tapes touch (set by the planner)
    const bool exact = kc_tape_exact(ctx);   // EXACT mode, or the operand cone of a stencil (KcExactScope)
    const int shared_cap = ctx->lanes_open > 1 ? (exact && g_kc_tuning.smem_cap_exact_kb > 0 ? g_kc_tuning.smem_cap_exact_kb : 113) : 0;
    const TvmConfig c = pick_config(ns_max, nt_max, args.n, shared_cap);
    g_kc_last_tile_config[0] = c.v; g_kc_last_tile_config[1] = c.ctas; g_kc_last_tile_config[2] = c.stages;
    KcTimed timed(ctx, KC_KERNEL_TAPE);
    int32_t rc;
    {
        // a kernel specialised for this tape (kc_jit.cu): its temporaries are registers, so the launch
        // configuration is chosen for zero shared-memory temporaries, V bounded by the register budget
        TvmConfig j = pick_config(ns_max, 0, args.n, shared_cap);
        while (j.v > 1 && (2 + nt_max) * 4 * j.v > 96) j.v >>= 1;
        bool launched = false;
        KC_TRY(kcj_try_launch(ctx, args, ns_max, j.v, j.ctas, j.stages, &launched));
        if (launched) {
            g_kc_last_tile_config[0] = -j.v; g_kc_last_tile_config[1] = j.ctas; g_kc_last_tile_config[2] = j.stages;   // negative V: the specialised kernel ran
            ctx->kernel_launches++;
            ctx->run_kernels++;
            return KC_OK;
        }
    }
    const bool three = c.ctas >= 3;
#define KC_TVM(E, VV) (three ? launch_tile_vm<E, VV, 3>(ctx, args, c.stages, ns_max, nt_max, c.ctas) : launch_tile_vm<E, VV, 2>(ctx, args, c.stages, ns_max, nt_max, c.ctas))
    if (c.v == 4) rc = exact ? KC_TVM(true, 4) : KC_TVM(false, 4);
    else if (c.v == 2) rc = exact ? KC_TVM(true, 2) : KC_TVM(false, 2);
    else rc = exact ? KC_TVM(true, 1) : KC_TVM(false, 1);
#undef KC_TVM
    KC_TRY(rc);
    KC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    ctx->run_kernels++;
    return KC_OK;
}

int32_t kck_from_u8(kc_context* ctx, const uint8_t* d_samples, uint32_t channels, size_t n,
                    float* const planes[4]) {
    if (n == 0) return KC_OK;
    const int grid = grid_for(ctx, (n + 3) >> 2, 256, 6);   // 6 CTAs of 256 threads are resident per SM (<= 40 registers): one full wave
    KcTimed timed(ctx, KC_KERNEL_FROM_U8);
    switch (channels) {
        case 1: kc_from_u8_kernel<1><<<grid, 256, 0, ctx->stream>>>(d_samples, n, planes[0], planes[1], planes[2], planes[3]); break;
        case 2: kc_from_u8_kernel<2><<<grid, 256, 0, ctx->stream>>>(d_samples, n, planes[0], planes[1], planes[2], planes[3]); break;
        case 3: kc_from_u8_kernel<3><<<grid, 256, 0, ctx->stream>>>(d_samples, n, planes[0], planes[1], planes[2], planes[3]); break;
        default: kc_from_u8_kernel<4><<<grid, 256, 0, ctx->stream>>>(d_samples, n, planes[0], planes[1], planes[2], planes[3]); break;
    }
    KC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    ctx->run_kernels++;
    return KC_OK;
}
