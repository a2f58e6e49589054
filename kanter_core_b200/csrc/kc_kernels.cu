// kc_kernels.cu — the elementwise device kernels.
//
//  * kc_tape_kernel: ONE kernel for every chain/tree of per-pixel node work:
//    Mix add/subtract/multiply/divide/pow (src/node/mix.rs:136-302), the
//    Rgba->Gray average of SlotImage::as_type (src/slot_image.rs:242-253),
//    constant fills (SlotImage::from_value, :28-64) and the f32->RGBA8 export of
//    SlotImage::to_u8 / to_u8_srgb (:142-207).  It interprets a short op tape
//    (kc_internal.h) over float4s of pixels with sources, temporaries and the
//    accumulator all in registers, so a fused group reads each source plane once
//    and writes each result plane once; intermediates never touch HBM.
//    blockIdx.y selects one of up to four independent segments (e.g. the R, G
//    and B expression chains of an Rgba graph), each thread interprets its
//    segment's tape once for V float4s at a time.
//  * kc_from_u8_kernel: deconstruct_image's u8/255 de-interleave (src/shared.rs:16-56).
//  * kc_fill_kernel: materialise a constant plane.
//
// HBM-bound streaming work: 16-byte coalesced accesses, streaming cache hints
// (every byte is touched once), all source loads of a pixel group issued before
// the first use, grid.x = SM count x resident CTAs (grid-stride inside).
#include "kc_internal.h"

namespace {

// ---------------------------------------------------------------------------
// pow.
// EXACT: Rust's f32::powf is the platform libm's powf; on Linux that is glibc's
// (2.28+: sysdeps/ieee754/flt-32/e_powf.c with powf_log2_data.c and
// exp2f_data.c, the ARM optimized-routines algorithm).  It is NOT correctly
// rounded (its fp64 log2/exp2 polynomials carry ~2^-33 relative error, so about
// one result in a thousand is the "other" neighbouring float), so bit-identical
// results need the same algorithm: this is a restatement of it in fp64 with the
// same tables, polynomial coefficients, evaluation order and special cases.
// ---------------------------------------------------------------------------
__constant__ double kPowLogTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
// bits(2^(i/32)) - (i << 47)
__constant__ unsigned long long kExp2Tab[32] = {
    0x3ff0000000000000ull, 0x3fefd9b0d3158574ull, 0x3fefb5586cf9890full, 0x3fef9301d0125b51ull,
    0x3fef72b83c7d517bull, 0x3fef54873168b9aaull, 0x3fef387a6e756238ull, 0x3fef1e9df51fdee1ull,
    0x3fef06fe0a31b715ull, 0x3feef1a7373aa9cbull, 0x3feedea64c123422ull, 0x3feece086061892dull,
    0x3feebfdad5362a27ull, 0x3feeb42b569d4f82ull, 0x3feeab07dd485429ull, 0x3feea47eb03a5585ull,
    0x3feea09e667f3bcdull, 0x3fee9f75e8ec5f74ull, 0x3feea11473eb0187ull, 0x3feea589994cce13ull,
    0x3feeace5422aa0dbull, 0x3feeb737b0cdc5e5ull, 0x3feec49182a3f090ull, 0x3feed503b23e255dull,
    0x3feee89f995ad3adull, 0x3feeff76f2fb5e47ull, 0x3fef199bdd85529cull, 0x3fef3720dcef9069ull,
    0x3fef5818dcfba487ull, 0x3fef7c97337b9b5full, 0x3fefa4afa2a490daull, 0x3fefd0765b6e4540ull};

__device__ __forceinline__ int pw_checkint(uint32_t iy) {  // 0: not an integer, 1: odd, 2: even
    const int e = (iy >> 23) & 0xff;
    if (e < 0x7f) return 0;
    if (e > 0x7f + 23) return 2;
    if (iy & ((1u << (0x7f + 23 - e)) - 1)) return 0;
    if (iy & (1u << (0x7f + 23 - e))) return 1;
    return 2;
}
__device__ __forceinline__ bool pw_zeroinfnan(uint32_t i) { return 2 * i - 1 >= 2u * 0x7f800000u - 1; }

__device__ __noinline__ float kc_pow_special(float x, float y, uint32_t& ix, uint32_t& sign_bias, bool& done) {
    // the rare half of glibc's powf: x < 2^-126, inf, nan, negative; y zero, inf, nan
    const uint32_t iy = __float_as_uint(y);
    done = true;
    if (pw_zeroinfnan(iy)) {
        if (2 * iy == 0) return 1.0f;
        if (ix == 0x3f800000u) return 1.0f;
        if (2 * ix > 2u * 0x7f800000u || 2 * iy > 2u * 0x7f800000u) return __fadd_rn(x, y);
        if (2 * ix == 2 * 0x3f800000u) return 1.0f;
        if ((2 * ix < 2 * 0x3f800000u) == !(iy & 0x80000000u)) return 0.0f;
        return __fmul_rn(y, y);
    }
    if (pw_zeroinfnan(ix)) {
        float x2 = __fmul_rn(x, x);
        if ((ix & 0x80000000u) && pw_checkint(iy) == 1) x2 = -x2;
        return (iy & 0x80000000u) ? __fdiv_rn(1.0f, x2) : x2;
    }
    if (ix & 0x80000000u) {
        const int yint = pw_checkint(iy);
        if (yint == 0) return __int_as_float(0x7fc00000);  // invalid: NaN
        if (yint == 1) sign_bias = 1u << 16;
        ix &= 0x7fffffffu;
    }
    if (ix < 0x00800000u) {  // subnormal x: normalise
        ix = __float_as_uint(__fmul_rn(x, 0x1p23f));
        ix &= 0x7fffffffu;
        ix -= 23u << 23;
    }
    done = false;
    return 0.0f;
}

__device__ __forceinline__ float kc_pow_exact(float x, float y) {
    uint32_t sign_bias = 0;
    uint32_t ix = __float_as_uint(x);
    const uint32_t iy = __float_as_uint(y);
    if (ix - 0x00800000u >= 0x7f800000u - 0x00800000u || pw_zeroinfnan(iy)) {
        bool done;
        const float r = kc_pow_special(x, y, ix, sign_bias, done);
        if (done) return r;
    }
    // log2_inline
    const uint32_t tmp = ix - 0x3f330000u;
    const int i = (tmp >> 19) & 15;
    const uint32_t top = tmp & 0xff800000u;
    const uint32_t iz = ix - top;
    const int k = (int)top >> 23;
    const double invc = kPowLogTab[i][0], logc = kPowLogTab[i][1];
    const double z = (double)__uint_as_float(iz);
    const double r = fma(z, invc, -1.0);
    const double y0 = __dadd_rn(logc, (double)k);
    const double r2 = __dmul_rn(r, r);
    double yy = fma(0x1.27616c9496e0bp-2, r, -0x1.71969a075c67ap-2);
    const double p = fma(0x1.ec70a6ca7baddp-2, r, -0x1.7154748bef6c8p-1);
    const double r4 = __dmul_rn(r2, r2);
    double q = fma(0x1.71547652ab82bp+0, r, y0);
    q = fma(p, r2, q);
    yy = fma(yy, r4, q);
    const double ylogx = __dmul_rn((double)y, yy);
    const unsigned long long yb = (unsigned long long)__double_as_longlong(ylogx);
    if (((yb >> 47) & 0xffff) >= (0x405f800000000000ull >> 47)) {  // |y*log2(x)| >= 126
        if (ylogx > 0x1.fffffffd1d571p+6) return sign_bias ? __int_as_float(0xff800000) : __int_as_float(0x7f800000);
        if (ylogx <= -150.0) return sign_bias ? -0.0f : 0.0f;
        if (ylogx < -149.0) return sign_bias ? -0x1p-149f : 0x1p-149f;
    }
    // exp2_inline
    double kd = __dadd_rn(ylogx, 0x1.8p+47);
    const unsigned long long ki = (unsigned long long)__double_as_longlong(kd);
    kd = __dsub_rn(kd, 0x1.8p+47);
    const double rr = __dsub_rn(ylogx, kd);
    unsigned long long t = kExp2Tab[ki & 31];
    t += (ki + sign_bias) << 47;
    const double s = __longlong_as_double((long long)t);
    const double zz = fma(0x1.c6af84b912394p-5, rr, 0x1.ebfce50fac4f3p-3);
    const double rr2 = __dmul_rn(rr, rr);
    double y2 = fma(0x1.62e42ff0c52d6p-1, rr, 1.0);
    y2 = fma(zz, rr2, y2);
    y2 = __dmul_rn(y2, s);
    return __double2float_rn(y2);
}

// FAST: x^y = 2^(y*log2 x) on the special-function unit for positive normal x
// and |y| <= 16, with the exponent product split so its rounding error does not
// scale with |log2 x| (x = m*2^e, m in [sqrt(.5), sqrt(2)); t = y*e + y*log2 m).
// ~4e-7 relative error; anything else takes the exact routine.
__device__ __forceinline__ float kc_pow_fast(float x, float y) {
    const uint32_t ix = __float_as_uint(x);
    if (ix - 0x00800000u < 0x7f000000u && fabsf(y) <= 16.0f) {
        const int e = (int)(ix - 0x3f3504f3u) >> 23;
        const float m = __uint_as_float(ix - ((uint32_t)e << 23));
        const float ef = (float)e;
        const float lm = __log2f(m);
        const float p1 = y * ef;
        const float r1 = fmaf(y, ef, -p1);  // exact residual of the product
        const float nf = rintf(p1);
        if (fabsf(nf) < 100.0f) {
            const float f = (p1 - nf) + fmaf(y, lm, r1);
            const float s = __uint_as_float((uint32_t)((int)nf + 127) << 23);
            return exp2f(f) * s;
        }
    }
    return kc_pow_exact(x, y);
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

#define KC_LANES(fn, a, x) make_float4(fn((a).x, (x).x), fn((a).y, (x).y), fn((a).z, (x).z), fn((a).w, (x).w))
#define KC_LANES_R(fn, a, x) make_float4(fn((x).x, (a).x), fn((x).y, (a).y), fn((x).z, (a).z), fn((x).w, (a).w))

template <bool EXACT>
__device__ __forceinline__ float4 tape_binary(uint32_t op, float4 a, float4 x) {
    switch (op) {
        case TOP_ADD: return KC_LANES(__fadd_rn, a, x);
        case TOP_SUB: return KC_LANES(__fsub_rn, a, x);
        case TOP_RSUB: return KC_LANES_R(__fsub_rn, a, x);
        case TOP_MUL: return KC_LANES(__fmul_rn, a, x);
        case TOP_DIV: return KC_LANES(__fdiv_rn, a, x);
        case TOP_RDIV: return KC_LANES_R(__fdiv_rn, a, x);
        case TOP_POW: return KC_LANES(kc_pow<EXACT>, a, x);
        default: return KC_LANES_R(kc_pow<EXACT>, a, x);
    }
}

// vector index `idx` of a plane of n pixels: full float4s below nfull, one ragged
// group of `tail` pixels at nfull.
__device__ __forceinline__ float4 load_group(const float* p, size_t idx, size_t nfull, int tail) {
    if (idx < nfull) return __ldcs(reinterpret_cast<const float4*>(p) + idx);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (idx == nfull && tail) {
        const float* q = p + 4 * nfull;
        v.x = q[0];
        if (tail > 1) v.y = q[1];
        if (tail > 2) v.z = q[2];
    }
    return v;
}
__device__ __forceinline__ void store_group(float* p, size_t idx, float4 v, size_t nfull, int tail) {
    if (idx < nfull) {
        __stcs(reinterpret_cast<float4*>(p) + idx, v);
    } else if (idx == nfull && tail) {
        float* q = p + 4 * nfull;
        q[0] = v.x;
        if (tail > 1) q[1] = v.y;
        if (tail > 2) q[2] = v.z;
    }
}
__device__ __forceinline__ void store_group_u32(uint32_t* p, size_t idx, uint4 v, size_t nfull, int tail) {
    if (idx < nfull) {
        __stcs(reinterpret_cast<uint4*>(p) + idx, v);
    } else if (idx == nfull && tail) {
        uint32_t* q = p + 4 * nfull;
        q[0] = v.x;
        if (tail > 1) q[1] = v.y;
        if (tail > 2) q[2] = v.z;
    }
}

constexpr int TAPE_BLOCK = 256;

// NS sources, NT temporaries, V float4s per thread and tape pass.
template <bool EXACT, int NS, int NT, int V, int MINB>
__global__ void __launch_bounds__(TAPE_BLOCK, MINB) kc_tape_kernel(const __grid_constant__ KcTapeArgs A) {
    const KcSegment& G = A.seg[blockIdx.y];
    const size_t nfull = (size_t)(A.n >> 2);
    const int tail = (int)(A.n & 3ull);
    const size_t ngroups = nfull + (tail ? 1 : 0);
    const uint32_t pc0 = G.tape_begin, pc1 = G.tape_end;
    const uint32_t n_src = G.n_src;
    for (size_t base = (size_t)blockIdx.x * (V * TAPE_BLOCK); base < ngroups; base += (size_t)gridDim.x * (V * TAPE_BLOCK)) {
        size_t idx[V];
#pragma unroll
        for (int j = 0; j < V; ++j) idx[j] = base + (size_t)j * TAPE_BLOCK + threadIdx.x;
        // every source load of this pixel group is issued before the first use
        float4 S[NS][V];
#pragma unroll
        for (int k = 0; k < NS; ++k)
#pragma unroll
            for (int j = 0; j < V; ++j)
                S[k][j] = (k < (int)n_src) ? load_group(G.src[k], idx[j], nfull, tail) : make_float4(0.f, 0.f, 0.f, 0.f);
        float4 T[NT][V];
        float4 acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) {
            acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int t = 0; t < NT; ++t) T[t][j] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (uint32_t pc = pc0; pc < pc1; ++pc) {
            const uint32_t in = A.instr[pc];
            const uint32_t op = in & 0xffu;
            const uint32_t arg = (in >> 8) & 0xffu;
            if (op <= TOP_RPOW) {
                float4 x[V];
                // warp-uniform switch; constant indices keep S/T in registers
#define KC_FETCH(arr, k) { _Pragma("unroll") for (int j = 0; j < V; ++j) x[j] = arr[k][j]; }
                switch (arg) {
                    case 0: KC_FETCH(S, 0) break;
                    case 1: if (NS > 1) KC_FETCH(S, NS > 1 ? 1 : 0) break;
                    case 2: if (NS > 2) KC_FETCH(S, NS > 2 ? 2 : 0) break;
                    case 3: if (NS > 3) KC_FETCH(S, NS > 3 ? 3 : 0) break;
                    case 4: if (NS > 4) KC_FETCH(S, NS > 4 ? 4 : 0) break;
                    case 5: if (NS > 5) KC_FETCH(S, NS > 5 ? 5 : 0) break;
                    case 6: if (NS > 6) KC_FETCH(S, NS > 6 ? 6 : 0) break;
                    case 7: if (NS > 7) KC_FETCH(S, NS > 7 ? 7 : 0) break;
                    case 8: KC_FETCH(T, 0) break;
                    case 9: if (NT > 1) KC_FETCH(T, NT > 1 ? 1 : 0) break;
                    case 10: if (NT > 2) KC_FETCH(T, NT > 2 ? 2 : 0) break;
                    case 11: if (NT > 3) KC_FETCH(T, NT > 3 ? 3 : 0) break;
                    case 12: if (NT > 4) KC_FETCH(T, NT > 4 ? 4 : 0) break;
                    case 13: if (NT > 5) KC_FETCH(T, NT > 5 ? 5 : 0) break;
                    default: {
                        const float v = A.imm[pc];
#pragma unroll
                        for (int j = 0; j < V; ++j) x[j] = make_float4(v, v, v, v);
                    } break;
                }
#undef KC_FETCH
                if (op == TOP_LD) {
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[j] = x[j];
                } else {
#pragma unroll
                    for (int j = 0; j < V; ++j) acc[j] = tape_binary<EXACT>(op, acc[j], x[j]);
                }
            } else if (op == TOP_ST_TMP) {
#define KC_PUT(k) { _Pragma("unroll") for (int j = 0; j < V; ++j) T[k][j] = acc[j]; }
                switch (arg) {
                    case 0: KC_PUT(0) break;
                    case 1: if (NT > 1) KC_PUT(NT > 1 ? 1 : 0) break;
                    case 2: if (NT > 2) KC_PUT(NT > 2 ? 2 : 0) break;
                    case 3: if (NT > 3) KC_PUT(NT > 3 ? 3 : 0) break;
                    case 4: if (NT > 4) KC_PUT(NT > 4 ? 4 : 0) break;
                    default: if (NT > 5) KC_PUT(NT > 5 ? 5 : 0) break;
                }
#undef KC_PUT
            } else if (op == TOP_ST_OUT) {
                float* o = G.out[arg];
#pragma unroll
                for (int j = 0; j < V; ++j) store_group(o, idx[j], acc[j], nfull, tail);
            } else if (NT >= 3 || op == TOP_PACK_GRAY) {
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    uint32_t px[4];
                    if (op == TOP_PACK_RGBA) {
                        const float4 tr = T[0][j], tg = T[NT > 1 ? 1 : 0][j], tb = T[NT > 2 ? 2 : 0][j];
                        const float r[4] = {tr.x, tr.y, tr.z, tr.w};
                        const float g[4] = {tg.x, tg.y, tg.z, tg.w};
                        const float b[4] = {tb.x, tb.y, tb.z, tb.w};
                        const float a[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            uint32_t R, Gc, B;
                            if (arg) { R = kc_to_u8_srgb<EXACT>(r[l]); Gc = kc_to_u8_srgb<EXACT>(g[l]); B = kc_to_u8_srgb<EXACT>(b[l]); }
                            else { R = kc_to_u8(r[l]); Gc = kc_to_u8(g[l]); B = kc_to_u8(b[l]); }
                            px[l] = R | (Gc << 8) | (B << 16) | (kc_to_u8(a[l]) << 24);
                        }
                    } else {  // TOP_PACK_GRAY: [v, v, v, 255]
                        const float v[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
#pragma unroll
                        for (int l = 0; l < 4; ++l) {
                            const uint32_t u = arg ? kc_to_u8_srgb<EXACT>(v[l]) : kc_to_u8(v[l]);
                            px[l] = u | (u << 8) | (u << 16) | 0xff000000u;
                        }
                    }
                    store_group_u32(G.out_rgba8, idx[j], make_uint4(px[0], px[1], px[2], px[3]), nfull, tail);
                }
            }
        }
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
            __stcs(reinterpret_cast<float4*>(planes[c]) + v, make_float4(vals[0][c], vals[1][c], vals[2][c], vals[3][c]));
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

template <bool EXACT>
void launch_tape_variant(const KcTapeArgs& a, dim3 grid, cudaStream_t st) {
    switch (a.variant) {
        case 0: kc_tape_kernel<EXACT, 2, 2, 2, 2><<<grid, TAPE_BLOCK, 0, st>>>(a); break;
        case 1: kc_tape_kernel<EXACT, 4, 3, 2, 2><<<grid, TAPE_BLOCK, 0, st>>>(a); break;
        default: kc_tape_kernel<EXACT, KC_MAX_SRC, KC_MAX_TMP, 1, 2><<<grid, TAPE_BLOCK, 0, st>>>(a); break;
    }
}

}  // namespace

int32_t kck_launch_tape(kc_context* ctx, const KcTapeArgs& args) {
    if (args.n == 0 || args.n_seg == 0) return KC_OK;
    const int v = args.variant <= 1 ? 2 : 1;
    const size_t groups = (size_t)((args.n + 3) >> 2);
    // a whole number of CTAs per SM across all segments; grid-stride loop inside
    const int per_seg_cap = std::max(1, (ctx->sm_count * 8) / (int)args.n_seg);
    size_t want = (groups + (size_t)v * TAPE_BLOCK - 1) / ((size_t)v * TAPE_BLOCK);
    dim3 grid((unsigned)std::min<size_t>(std::max<size_t>(want, 1), (size_t)per_seg_cap), args.n_seg);
    KcTimed timed(ctx, KC_KERNEL_TAPE);
    if (ctx->opts.math_mode == KC_MATH_EXACT) launch_tape_variant<true>(args, grid, ctx->stream);
    else launch_tape_variant<false>(args, grid, ctx->stream);
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
