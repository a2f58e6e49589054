// kc_kernels.cu — the elementwise device kernels.
//
//  * kc_tape_kernel: ONE kernel for every chain/tree of per-pixel node work:
//    Mix add/subtract/multiply/divide/pow (src/node/mix.rs:136-302), the
//    Rgba->Gray average of SlotImage::as_type (src/slot_image.rs:242-253),
//    constant fills (SlotImage::from_value, :28-64) and the f32->RGBA8 export of
//    SlotImage::to_u8 / to_u8_srgb (:142-207).  It interprets a short op tape
//    (kc_internal.h) once per float4 of pixels with sources, temporaries and the
//    accumulator all in registers, so a fused group reads each source plane once
//    and writes each result plane once; intermediates never touch HBM.
//  * kc_from_u8_kernel: deconstruct_image's u8/255 de-interleave (src/shared.rs:16-56).
//  * kc_fill_kernel: materialise a constant plane.
//
// HBM-bound streaming work: 16-byte coalesced accesses, streaming cache hints
// (every byte is touched once), all source loads of a pixel group issued before
// the first use, grid = SM count x resident CTAs.
#include "kc_internal.h"

namespace {

// ---------------------------------------------------------------------------
// arithmetic.  EXACT mode spells every operation with a round-to-nearest
// intrinsic so that no compiler flag can contract or reassociate it.
// ---------------------------------------------------------------------------
__device__ __forceinline__ float kc_pow_exact(float a, float b) {
    // Rust f32::powf == glibc powf, which is correctly rounded in all but ~1e-9
    // of cases.  fp64 pow (<= 2 ulp of double) rounded to f32 matches it.
    return (float)pow((double)a, (double)b);
}

__device__ __forceinline__ float kc_pow_fast(float x, float y) {
    // x^y = 2^(y*log2 x) for positive normal x and moderate y, with the exponent
    // product split so its rounding error does not scale with |log2 x|:
    // x = m * 2^e, m in [sqrt(.5), sqrt(2));  t = y*e + y*log2(m).
    const uint32_t ix = __float_as_uint(x);
    if (ix - 0x00800000u < 0x7f000000u && fabsf(y) <= 16.0f) {
        const int e = (int)(ix - 0x3f3504f3u) >> 23;
        const float m = __uint_as_float(ix - ((uint32_t)e << 23));
        const float ef = (float)e;
        const float lm = __log2f(m);              // |lm| <= 0.5, abs err 2^-22
        const float p1 = y * ef;
        const float r1 = fmaf(y, ef, -p1);        // exact residual of the product
        const float nf = rintf(p1);
        if (fabsf(nf) < 100.0f) {
            const float f = (p1 - nf) + fmaf(y, lm, r1);   // |f| <= 0.5 + 8
            const float s = __uint_as_float((uint32_t)((int)nf + 127) << 23);
            return exp2f(f) * s;
        }
    }
    return powf(x, y);
}

template <bool EXACT>
__device__ __forceinline__ float kc_pow(float a, float b) {
    return EXACT ? kc_pow_exact(a, b) : kc_pow_fast(a, b);
}

// SlotImage::f32_to_u8, src/slot_image.rs:142-145:
//   ((v.clamp(0,1) * 255.).min(255.)) as u8
// Rust's clamp keeps NaN, min(NaN,255) = 255, `as u8` truncates and saturates.
__device__ __forceinline__ uint32_t kc_to_u8(float v) {
    float c = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    float m = __fmul_rn(c, 255.0f);
    m = (m != m) ? 255.0f : fminf(m, 255.0f);
    return __float2uint_rz(m);
}
// srgb_to_linear, src/slot_data.rs:100-109, applied to the clamped value (:173-176)
template <bool EXACT>
__device__ __forceinline__ uint32_t kc_to_u8_srgb(float v) {
    float c = v < 0.0f ? 0.0f : (v > 1.0f ? 1.0f : v);
    float l;
    if (c <= 0.0f) l = c;
    else if (c <= 0.04045f) l = __fdiv_rn(c, 12.92f);
    else l = kc_pow<EXACT>(__fdiv_rn(__fadd_rn(c, 0.055f), 1.055f), 2.4f);
    float m = __fmul_rn(l, 255.0f);
    m = (m != m) ? 255.0f : fminf(m, 255.0f);
    return __float2uint_rz(m);
}

__device__ __forceinline__ float4 ld_stream(const float* p) {
    return __ldcs(reinterpret_cast<const float4*>(p));
}
__device__ __forceinline__ void st_stream(float* p, float4 v) {
    __stcs(reinterpret_cast<float4*>(p), v);
}

#define KC_LANES(expr_x, expr_y, expr_z, expr_w) make_float4(expr_x, expr_y, expr_z, expr_w)

template <bool EXACT>
__device__ __forceinline__ float4 tape_binary(uint32_t op, float4 a, float4 x) {
    switch (op) {
        case TOP_ADD: return KC_LANES(__fadd_rn(a.x, x.x), __fadd_rn(a.y, x.y), __fadd_rn(a.z, x.z), __fadd_rn(a.w, x.w));
        case TOP_SUB: return KC_LANES(__fsub_rn(a.x, x.x), __fsub_rn(a.y, x.y), __fsub_rn(a.z, x.z), __fsub_rn(a.w, x.w));
        case TOP_RSUB: return KC_LANES(__fsub_rn(x.x, a.x), __fsub_rn(x.y, a.y), __fsub_rn(x.z, a.z), __fsub_rn(x.w, a.w));
        case TOP_MUL: return KC_LANES(__fmul_rn(a.x, x.x), __fmul_rn(a.y, x.y), __fmul_rn(a.z, x.z), __fmul_rn(a.w, x.w));
        case TOP_DIV: return KC_LANES(__fdiv_rn(a.x, x.x), __fdiv_rn(a.y, x.y), __fdiv_rn(a.z, x.z), __fdiv_rn(a.w, x.w));
        case TOP_RDIV: return KC_LANES(__fdiv_rn(x.x, a.x), __fdiv_rn(x.y, a.y), __fdiv_rn(x.z, a.z), __fdiv_rn(x.w, a.w));
        case TOP_POW: return KC_LANES(kc_pow<EXACT>(a.x, x.x), kc_pow<EXACT>(a.y, x.y), kc_pow<EXACT>(a.z, x.z), kc_pow<EXACT>(a.w, x.w));
        default: return KC_LANES(kc_pow<EXACT>(x.x, a.x), kc_pow<EXACT>(x.y, a.y), kc_pow<EXACT>(x.z, a.z), kc_pow<EXACT>(x.w, a.w));
    }
}

// Interpret the tape for one float4 of pixels.  S[] holds the preloaded sources.
// `lanes` < 4 only for the single ragged tail group (scalar stores there).
template <bool EXACT>
__device__ __forceinline__ void run_tape(const KcTapeArgs& A, const float4 (&S)[KC_MAX_SRC], size_t pix, int lanes) {
    float4 T[KC_MAX_TMP];
#pragma unroll
    for (int j = 0; j < KC_MAX_TMP; ++j) T[j] = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    const uint32_t n_instr = A.n_instr;
    for (uint32_t pc = 0; pc < n_instr; ++pc) {
        const uint32_t in = A.instr[pc];
        const uint32_t op = in & 0xffu;
        const uint32_t arg = (in >> 8) & 0xffu;
        if (op <= TOP_RPOW) {
            float4 x;
            switch (arg) {  // warp-uniform; constant indices keep S/T in registers
                case 0: x = S[0]; break;
                case 1: x = S[1]; break;
                case 2: x = S[2]; break;
                case 3: x = S[3]; break;
                case 4: x = S[4]; break;
                case 5: x = S[5]; break;
                case 6: x = S[6]; break;
                case 7: x = S[7]; break;
                case 8: x = T[0]; break;
                case 9: x = T[1]; break;
                case 10: x = T[2]; break;
                case 11: x = T[3]; break;
                case 12: x = T[4]; break;
                case 13: x = T[5]; break;
                default: { const float v = A.imm[pc]; x = make_float4(v, v, v, v); } break;
            }
            acc = (op == TOP_LD) ? x : tape_binary<EXACT>(op, acc, x);
        } else if (op == TOP_ST_TMP) {
            switch (arg) {
                case 0: T[0] = acc; break;
                case 1: T[1] = acc; break;
                case 2: T[2] = acc; break;
                case 3: T[3] = acc; break;
                case 4: T[4] = acc; break;
                default: T[5] = acc; break;
            }
        } else if (op == TOP_ST_OUT) {
            float* o = A.out[arg] + pix;
            if (lanes == 4) {
                st_stream(o, acc);
            } else {
                o[0] = acc.x;
                if (lanes > 1) o[1] = acc.y;
                if (lanes > 2) o[2] = acc.z;
            }
        } else {
            uint32_t px[4];
            if (op == TOP_PACK_RGBA) {
                const float r[4] = {T[0].x, T[0].y, T[0].z, T[0].w};
                const float g[4] = {T[1].x, T[1].y, T[1].z, T[1].w};
                const float b[4] = {T[2].x, T[2].y, T[2].z, T[2].w};
                const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    uint32_t R, G, B;
                    if (arg) { R = kc_to_u8_srgb<EXACT>(r[l]); G = kc_to_u8_srgb<EXACT>(g[l]); B = kc_to_u8_srgb<EXACT>(b[l]); }
                    else { R = kc_to_u8(r[l]); G = kc_to_u8(g[l]); B = kc_to_u8(b[l]); }
                    px[l] = R | (G << 8) | (B << 16) | (kc_to_u8(a[l]) << 24);
                }
            } else {  // TOP_PACK_GRAY: [v, v, v, 255]
                const float v[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
                for (int l = 0; l < 4; ++l) {
                    const uint32_t u = arg ? kc_to_u8_srgb<EXACT>(v[l]) : kc_to_u8(v[l]);
                    px[l] = u | (u << 8) | (u << 16) | 0xff000000u;
                }
            }
            uint32_t* o = A.out_rgba8 + pix;
            if (lanes == 4) {
                __stcs(reinterpret_cast<uint4*>(o), make_uint4(px[0], px[1], px[2], px[3]));
            } else {
                o[0] = px[0];
                if (lanes > 1) o[1] = px[1];
                if (lanes > 2) o[2] = px[2];
            }
        }
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(256) kc_tape_kernel(const __grid_constant__ KcTapeArgs A) {
    const size_t nvec = (size_t)(A.n >> 2);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        float4 S[KC_MAX_SRC];
#pragma unroll
        for (int k = 0; k < KC_MAX_SRC; ++k)
            S[k] = (k < (int)A.n_src) ? ld_stream(A.src[k] + 4 * v) : make_float4(0.f, 0.f, 0.f, 0.f);
        run_tape<EXACT>(A, S, 4 * v, 4);
    }
    // ragged tail (n % 4 pixels): one thread, scalar loads
    const int tail = (int)(A.n & 3ull);
    if (tail && blockIdx.x == 0 && threadIdx.x == 0) {
        float4 S[KC_MAX_SRC];
#pragma unroll
        for (int k = 0; k < KC_MAX_SRC; ++k) {
            S[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (k < (int)A.n_src) {
                const float* p = A.src[k] + 4 * nvec;
                S[k].x = p[0];
                if (tail > 1) S[k].y = p[1];
                if (tail > 2) S[k].z = p[2];
            }
        }
        run_tape<EXACT>(A, S, 4 * nvec, tail);
    }
}

__global__ void __launch_bounds__(256) kc_fill_kernel(float* __restrict__ dst, size_t n, float v) {
    const size_t nvec = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const float4 v4 = make_float4(v, v, v, v);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride)
        __stcs(reinterpret_cast<float4*>(dst) + i, v4);
    if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[4 * nvec + threadIdx.x] = v;
}

// deconstruct_image, src/shared.rs:27-33: plane_c[i] = samples[i*C + c] as f32 / 255.
// One thread converts 4 consecutive pixels: it reads 4*C bytes and writes one
// float4 per channel plane.
template <int C>
__global__ void __launch_bounds__(256) kc_from_u8_kernel(const uint8_t* __restrict__ s, size_t n,
                                                          float* __restrict__ p0, float* __restrict__ p1,
                                                          float* __restrict__ p2, float* __restrict__ p3) {
    const size_t nvec = n >> 2;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    float* planes[4] = {p0, p1, p2, p3};
    for (size_t v = (size_t)blockIdx.x * blockDim.x + threadIdx.x; v < nvec; v += stride) {
        // 4*C bytes starting at a 4-byte aligned offset: C 32-bit loads
        uint32_t wds[C];
        const uint32_t* src = reinterpret_cast<const uint32_t*>(s + 4 * C * v);
#pragma unroll
        for (int k = 0; k < C; ++k) wds[k] = __ldcs(src + k);
        float vals[4][C];
#pragma unroll
        for (int b = 0; b < 4 * C; ++b) {
            const uint32_t byte = (wds[b >> 2] >> (8 * (b & 3))) & 0xffu;
            vals[b / C][b % C] = __fdiv_rn((float)byte, 255.0f);
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
            st_stream(planes[c] + 4 * v, make_float4(vals[0][c], vals[1][c], vals[2][c], vals[3][c]));
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

}  // namespace

int32_t kck_launch_tape(kc_context* ctx, const KcTapeArgs& args) {
    if (args.n == 0) return KC_OK;
    const int block = 256;
    // persistent-style grid: a whole number of CTAs per SM, grid-stride loop inside
    const int grid = grid_for(ctx, (size_t)((args.n + 3) >> 2), block, 8);
    KcTimed timed(ctx, KC_KERNEL_TAPE);
    if (ctx->opts.math_mode == KC_MATH_EXACT)
        kc_tape_kernel<true><<<grid, block, 0, ctx->stream>>>(args);
    else
        kc_tape_kernel<false><<<grid, block, 0, ctx->stream>>>(args);
    KC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    ctx->run_kernels++;
    return KC_OK;
}

int32_t kck_fill(kc_context* ctx, float* dst, size_t n, float v) {
    if (n == 0) return KC_OK;
    const int grid = grid_for(ctx, (n + 3) >> 2, 256, 8);
    KcTimed timed(ctx, KC_KERNEL_FILL);
    kc_fill_kernel<<<grid, 256, 0, ctx->stream>>>(dst, n, v);
    KC_CUDA(cudaGetLastError());
    ctx->kernel_launches++;
    ctx->run_kernels++;
    return KC_OK;
}

int32_t kck_from_u8(kc_context* ctx, const uint8_t* d_samples, uint32_t channels, size_t n,
                    float* const planes[4]) {
    if (n == 0) return KC_OK;
    const int grid = grid_for(ctx, (n + 3) >> 2, 256, 8);
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
