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
#include <cstdlib>

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
// (global memory + __ldg, not __constant__: the index differs per lane and the constant
// cache would serialise the 16-way divergent lookups; L1 serves them in a few wavefronts)
__device__ const double kPowLogTab[16][2] = {
    {0x1.661ec79f8f3bep+0, -0x1.efec65b963019p-2}, {0x1.571ed4aaf883dp+0, -0x1.b0b6832d4fca4p-2},
    {0x1.49539f0f010b0p+0, -0x1.7418b0a1fb77bp-2}, {0x1.3c995b0b80385p+0, -0x1.39de91a6dcf7bp-2},
    {0x1.30d190c8864a5p+0, -0x1.01d9bf3f2b631p-2}, {0x1.25e227b0b8ea0p+0, -0x1.97c1d1b3b7af0p-3},
    {0x1.1bb4a4a1a343fp+0, -0x1.2f9e393af3c9fp-3}, {0x1.12358f08ae5bap+0, -0x1.960cbbf788d5cp-4},
    {0x1.0953f419900a7p+0, -0x1.a6f9db6475fcep-5}, {0x1.0000000000000p+0, 0x0.0p+0},
    {0x1.e608cfd9a47acp-1, 0x1.338ca9f24f53dp-4},  {0x1.ca4b31f026aa0p-1, 0x1.476a9543891bap-3},
    {0x1.b2036576afce6p-1, 0x1.e840b4ac4e4d2p-3},  {0x1.9c2d163a1aa2dp-1, 0x1.40645f0c6651cp-2},
    {0x1.886e6037841edp-1, 0x1.88e9c2c1b9ff8p-2},  {0x1.767dcf5534862p-1, 0x1.ce0a44eb17bccp-2}};
// bits(2^(i/32)) - (i << 47)
__device__ const unsigned long long kExp2Tab[32] = {
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

__device__ __forceinline__ float kc_pow_exact(float x, float y);
// out-of-line copy for the rare slow path of FAST mode (keeps the hot loop small)
__device__ __noinline__ float kc_pow_exact_call(float x, float y) { return kc_pow_exact(x, y); }

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
    const double2 tc = __ldg(reinterpret_cast<const double2*>(&kPowLogTab[i][0]));
    const double invc = tc.x, logc = tc.y;
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
    unsigned long long t = __ldg(&kExp2Tab[ki & 31]);
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
// Branch-free.  `rc`, `ay`, `ap` return the three quantities whose range decides
// whether the result is valid (checked once per float4 by the caller):
//   rc = bits(x) - bits(2^-126)  must be < 0x7f000000 (x positive, normal, finite)
//   ay = |y| <= 16,  ap = |y*e| < 100.
// +0 ^ positive (black pixels) is answered here: 0.  ~4e-7 relative error.
__device__ __forceinline__ float kc_pow_fast_core(float x, float y, uint32_t& rc, float& ay, float& ap) {
    const uint32_t ix = __float_as_uint(x);
    const uint32_t top = (ix - 0x3f3504f3u) & 0xff800000u;
    const float m = __uint_as_float(ix - top);
    const float ef = (float)((int)top >> 23);
    float lm;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lm) : "f"(m));
    const float p1 = y * ef;
    const float r1 = fmaf(y, ef, -p1);          // exact residual of the product
    const float t = p1 + 12582912.0f;           // 1.5 * 2^23: the integer nearest p1 sits in the low mantissa bits
    const float nf = t - 12582912.0f;
    const float f = (p1 - nf) + fmaf(y, lm, r1);
    float e2;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e2) : "f"(f));
    const float s = __int_as_float((__float_as_int(t) << 23) + 0x3f800000);  // 2^n
    const bool zero_pos = (ix == 0u) & (y > 0.0f);
    rc = zero_pos ? 0u : ix - 0x00800000u;
    ay = fabsf(y);
    ap = zero_pos ? 0.0f : fabsf(p1);
    return zero_pos ? 0.0f : e2 * s;
}

template <bool EXACT>
__device__ __forceinline__ float kc_pow(float a, float b) {
    if (EXACT) return kc_pow_exact(a, b);
    uint32_t rc; float ay, ap;
    const float r = kc_pow_fast_core(a, b, rc, ay, ap);
    const bool ok = (rc < 0x7f000000u) & (ay <= 16.0f) & (ap < 100.0f);
    return ok ? r : kc_pow_exact_call(a, b);
}

// SlotImage::f32_to_u8, src/slot_image.rs:142-145:
//   ((v.clamp(0,1) * 255.).min(255.)) as u8
// Rust's clamp keeps NaN, min(NaN,255) = 255, `as u8` truncates and saturates.
__device__ __forceinline__ uint32_t kc_to_u8(float v) {
    // clamp that keeps NaN (min.NaN / max.NaN), then the ordinary min drops it: NaN -> 255
    float c;
    asm("min.NaN.f32 %0, %1, 0f3F800000;" : "=f"(c) : "f"(v));
    asm("max.NaN.f32 %0, %1, 0f00000000;" : "=f"(c) : "f"(c));
    return __float2uint_rz(fminf(__fmul_rn(c, 255.0f), 255.0f));
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

// four lanes at once: one slow-path branch per float4 instead of one per lane
template <bool EXACT>
__device__ __forceinline__ float4 kc_pow4(float4 a, float4 b) {
    if (EXACT) return make_float4(kc_pow_exact(a.x, b.x), kc_pow_exact(a.y, b.y), kc_pow_exact(a.z, b.z), kc_pow_exact(a.w, b.w));
    uint32_t c0, c1, c2, c3;
    float y0, y1, y2, y3, p0, p1, p2, p3;
    float4 r = make_float4(kc_pow_fast_core(a.x, b.x, c0, y0, p0), kc_pow_fast_core(a.y, b.y, c1, y1, p1),
                           kc_pow_fast_core(a.z, b.z, c2, y2, p2), kc_pow_fast_core(a.w, b.w, c3, y3, p3));
    // one validity test for the four lanes (NaNs fail the float comparisons)
    const uint32_t cm = max(max(c0, c1), max(c2, c3));
    const float ym = fmaxf(fmaxf(y0, y1), fmaxf(y2, y3));
    const float pm = fmaxf(fmaxf(p0, p1), fmaxf(p2, p3));
    const bool nan_in = (y0 != y0) | (y1 != y1) | (y2 != y2) | (y3 != y3);  // fmaxf drops NaNs; |y*e| is NaN only if y is
    if (!((cm < 0x7f000000u) & (ym <= 16.0f) & (pm < 100.0f)) | nan_in) {
        if (!((c0 < 0x7f000000u) & (y0 <= 16.0f) & (p0 < 100.0f))) r.x = kc_pow_exact_call(a.x, b.x);
        if (!((c1 < 0x7f000000u) & (y1 <= 16.0f) & (p1 < 100.0f))) r.y = kc_pow_exact_call(a.y, b.y);
        if (!((c2 < 0x7f000000u) & (y2 <= 16.0f) & (p2 < 100.0f))) r.z = kc_pow_exact_call(a.z, b.z);
        if (!((c3 < 0x7f000000u) & (y3 <= 16.0f) & (p3 < 100.0f))) r.w = kc_pow_exact_call(a.w, b.w);
    }
    return r;
}

#define KC_LANES(fn, P_, Q_) make_float4(fn((P_).x, (Q_).x), fn((P_).y, (Q_).y), fn((P_).z, (Q_).z), fn((P_).w, (Q_).w))
#define KC_LANES_R(fn, P_, Q_) make_float4(fn((Q_).x, (P_).x), fn((Q_).y, (P_).y), fn((Q_).z, (P_).z), fn((Q_).w, (P_).w))

// ---------------------------------------------------------------------------
// The tile VM.
//
// Persistent CTAs walk (segment, tile) work items.  A tile is 1024*V pixels of
// every source plane of the segment, brought into shared memory by TMA bulk
// copies (cp.async.bulk, completion on an mbarrier) `stages`-1 tiles ahead of
// the arithmetic, so the bytes in flight per SM are set by the pipeline depth
// and not by how many registers the arithmetic needs.  The tape is then
// interpreted once per tile: each thread keeps the accumulator for its V
// float4s in registers, operands come from the shared-memory tile (sources) or
// from shared-memory temporaries, results go straight to global memory with
// 16-byte streaming stores.  The dispatch cost of an instruction is paid once
// per 4*V pixels per thread.
// ---------------------------------------------------------------------------
constexpr int TVM_THREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra WAIT_DONE;\n"
        "bra WAIT_LOOP;\n"
        "WAIT_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// global -> shared bulk copy (TMA, 1-D), bytes a multiple of 16, both sides 16-byte aligned;
// evict-first in L2: every source byte is read exactly once
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
                 : "memory");
}

// resolved instruction word (built by launch_tile_vm from the planner's op | arg << 8):
//   bits 0-3 op, bit 4 operand is a temporary, bit 5 operand is the immediate,
//   bits 8.. : operand byte offset inside the stage (sources) or the temporaries' block,
//              the output index (ST_OUT) or the sRGB flag (PACK_*)
constexpr uint32_t KC_R_TMP = 1u << 4;
constexpr uint32_t KC_R_IMM = 1u << 5;

__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts128(uint32_t addr, const float4& v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
#define KC_INPLACE4(INS, A_, X_)                          \
    do {                                                  \
        asm(INS : "+f"((A_).x) : "f"((X_).x));            \
        asm(INS : "+f"((A_).y) : "f"((X_).y));            \
        asm(INS : "+f"((A_).z) : "f"((X_).z));            \
        asm(INS : "+f"((A_).w) : "f"((X_).w));            \
    } while (0)

template <bool EXACT, int V, int MINB>
__global__ void __launch_bounds__(TVM_THREADS, MINB)
    kc_tile_vm_kernel(const __grid_constant__ KcTapeArgs A, int stages, int ns_max, uint32_t tiles_per_plane, uint32_t total_work) {
    constexpr int TILE_PX = 1024 * V;
    constexpr uint32_t TILE_B = TILE_PX * 4;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint64_t* mbar = reinterpret_cast<uint64_t*>(smem_raw);                 // [stages]
    float* stage_base = reinterpret_cast<float*>(smem_raw + 128);             // [stages][ns_max][TILE_PX]
    float* tmp_base = stage_base + (size_t)stages * ns_max * TILE_PX;         // [temporaries][TILE_PX]
    const int tid = threadIdx.x;
    const unsigned long long n = A.n;

    uint64_t policy = 0;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) mbar_init(&mbar[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    }
    __syncthreads();

    auto issue = [&](uint32_t w, int stage) {  // thread 0: start the loads of work item w
        const uint32_t seg = w / tiles_per_plane, tile = w - seg * tiles_per_plane;
        const unsigned long long px0 = (unsigned long long)tile * TILE_PX;
        const KcSegment& G = A.seg[seg];
        if (G.n_src == 0 || n - px0 < (unsigned long long)TILE_PX) return;  // ragged last tile: loaded cooperatively
        mbar_expect_tx(&mbar[stage], G.n_src * TILE_B);
        for (uint32_t k = 0; k < G.n_src; ++k)
            tma_load_1d(stage_base + ((size_t)stage * ns_max + k) * TILE_PX, G.src[k] + px0, TILE_B, &mbar[stage], policy);
    };

    const uint32_t w0 = blockIdx.x, wstride = gridDim.x;
    if (tid == 0)
        for (int s = 0; s < stages - 1; ++s) {
            const uint64_t w = (uint64_t)w0 + (uint64_t)s * wstride;
            if (w < total_work) issue((uint32_t)w, s);
        }
    uint32_t phase_bits = 0;  // bit s: parity of stage s's next completion
    uint32_t k_it = 0;
    for (uint64_t w = w0; w < total_work; w += wstride, ++k_it) {
        const int stage = (int)(k_it % (uint32_t)stages);
        if (tid == 0) {
            const uint64_t wn = w + (uint64_t)(stages - 1) * wstride;
            if (wn < total_work) issue((uint32_t)wn, (int)((k_it + stages - 1) % (uint32_t)stages));
        }
        const uint32_t seg = (uint32_t)w / tiles_per_plane, tile = (uint32_t)w - seg * tiles_per_plane;
        const unsigned long long px0 = (unsigned long long)tile * TILE_PX;
        const unsigned long long rem = n - px0;
        const bool full = rem >= (unsigned long long)TILE_PX;
        const KcSegment& G = A.seg[seg];
        float* sbuf = stage_base + (size_t)stage * ns_max * TILE_PX;
        if (G.n_src) {
            if (full) {
                mbar_wait(&mbar[stage], (phase_bits >> stage) & 1u);
                phase_bits ^= 1u << stage;
            } else {
                for (uint32_t k = 0; k < G.n_src; ++k)
                    for (int i = tid; i < TILE_PX; i += TVM_THREADS)
                        sbuf[(size_t)k * TILE_PX + i] = ((unsigned long long)i < rem) ? G.src[k][px0 + i] : 0.0f;
                __syncthreads();
            }
        }
        // ---- interpret the segment's tape over this tile ---------------------------------
        // The launch code resolved every operand to a byte offset (KC_R_* below), so decoding an
        // instruction is: constant-bank load, mask, one add, LDS.128.  The cheap ops update the
        // accumulator in place (inline PTX with "+f" operands): the compiler then keeps ONE copy
        // of acc across the switch arms instead of shuffling it through phi moves.
        float4 acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t src_addr = smem_u32(sbuf) + (uint32_t)tid * 16u;
        const uint32_t tmp_addr = smem_u32(tmp_base) + (uint32_t)tid * 16u;
        const uint32_t pc1 = G.tape_end;
        for (uint32_t pc = G.tape_begin; pc < pc1; ++pc) {
            const uint32_t in = A.instr[pc];
            const uint32_t op = in & 15u;
            const uint32_t arg = in >> 8;                 // operand byte offset / output index / sRGB flag
            if (op <= TOP_RPOW) {
                float4 x[V];
                if (in & KC_R_IMM) {
                    const float v = A.imm[pc];
#pragma unroll
                    for (int j = 0; j < V; ++j) x[j] = make_float4(v, v, v, v);
                } else {
                    const uint32_t addr = ((in & KC_R_TMP) ? tmp_addr : src_addr) + arg;
#pragma unroll
                    for (int j = 0; j < V; ++j) x[j] = lds128(addr + (uint32_t)j * (TVM_THREADS * 16u));
                }
                switch (op) {
                    case TOP_LD:
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = x[j];
                        break;
                    case TOP_ADD:
#pragma unroll
                        for (int j = 0; j < V; ++j) KC_INPLACE4("add.rn.f32 %0, %0, %1;", acc[j], x[j]);
                        break;
                    case TOP_SUB:
#pragma unroll
                        for (int j = 0; j < V; ++j) KC_INPLACE4("sub.rn.f32 %0, %0, %1;", acc[j], x[j]);
                        break;
                    case TOP_RSUB:
#pragma unroll
                        for (int j = 0; j < V; ++j) KC_INPLACE4("sub.rn.f32 %0, %1, %0;", acc[j], x[j]);
                        break;
                    case TOP_MUL:
#pragma unroll
                        for (int j = 0; j < V; ++j) KC_INPLACE4("mul.rn.f32 %0, %0, %1;", acc[j], x[j]);
                        break;
                    case TOP_DIV:
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = KC_LANES(__fdiv_rn, acc[j], x[j]);
                        break;
                    case TOP_RDIV:
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = KC_LANES_R(__fdiv_rn, acc[j], x[j]);
                        break;
                    case TOP_POW:
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = kc_pow4<EXACT>(acc[j], x[j]);
                        break;
                    default:  // TOP_RPOW
#pragma unroll
                        for (int j = 0; j < V; ++j) acc[j] = kc_pow4<EXACT>(x[j], acc[j]);
                        break;
                }
            } else if (op == TOP_ST_TMP) {
#pragma unroll
                for (int j = 0; j < V; ++j) sts128(tmp_addr + arg + (uint32_t)j * (TVM_THREADS * 16u), acc[j]);
            } else if (op == TOP_ST_OUT) {
                float* o = G.out[arg] + px0;
                if (full) {
#pragma unroll
                    for (int j = 0; j < V; ++j) __stcs(reinterpret_cast<float4*>(o) + j * TVM_THREADS + tid, acc[j]);
                } else {
#pragma unroll
                    for (int j = 0; j < V; ++j) {
                        const unsigned long long p = 4ull * (unsigned long long)(j * TVM_THREADS + tid);
                        if (p + 3 < rem) __stcs(reinterpret_cast<float4*>(o + p), acc[j]);
                        else if (p < rem) {
                            o[p] = acc[j].x;
                            if (p + 1 < rem) o[p + 1] = acc[j].y;
                            if (p + 2 < rem) o[p + 2] = acc[j].z;
                        }
                    }
                }
            } else {  // RGBA8 export
                const float4* t0 = reinterpret_cast<const float4*>(tmp_base) + tid;
                const float4* t1 = reinterpret_cast<const float4*>(tmp_base + TILE_PX) + tid;
                const float4* t2 = reinterpret_cast<const float4*>(tmp_base + 2 * TILE_PX) + tid;
                uint32_t* o = G.out_rgba8 + px0;
#pragma unroll
                for (int j = 0; j < V; ++j) {
                    uint32_t px[4];
                    if (op == TOP_PACK_RGBA) {
                        const float4 tr = t0[j * TVM_THREADS], tg = t1[j * TVM_THREADS], tb = t2[j * TVM_THREADS];
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
                    const unsigned long long p = 4ull * (unsigned long long)(j * TVM_THREADS + tid);
                    if (p + 3 < rem) __stcs(reinterpret_cast<uint4*>(o + p), make_uint4(px[0], px[1], px[2], px[3]));
                    else if (p < rem) {
                        o[p] = px[0];
                        if (p + 1 < rem) o[p + 1] = px[1];
                        if (p + 2 < rem) o[p + 2] = px[2];
                    }
                }
            }
        }
        __syncthreads();  // the stage (and the temporaries) may be overwritten from here on
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

// byte / 255.0f, correctly rounded, without the generic div.rn expansion: q = b*y, q += (b - 255 q)*y
// with y = RN(1/255).  All 256 inputs are checked against the oracle's `as f32 / 255.` by
// tests/test_gpu_ops.py::test_u8_roundtrip_every_value.
__device__ __forceinline__ float kc_u8_over_255(uint32_t byte) {
    const float a = (float)byte, y = 0x1.010102p-8f;
    const float q = __fmul_rn(a, y);
    return __fmaf_rn(__fmaf_rn(-255.0f, q, a), y, q);
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
            vals[b / C][b % C] = kc_u8_over_255(byte);
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
TvmConfig pick_config(int ns_max, int nt_max, unsigned long long n) {
    const int force_v = g_kc_tuning.tile_v, force_stages = g_kc_tuning.stages, force_ctas = g_kc_tuning.ctas;
    static const TvmConfig order[] = {{4, 3, 0}, {4, 2, 0}, {2, 3, 0}, {2, 2, 0}, {1, 3, 0}, {1, 2, 0}, {4, 1, 0}, {2, 1, 0}, {1, 1, 0}};
    for (int pass = 0; pass < 2; ++pass) {  // pass 0 honours the tuning overrides, pass 1 ignores them
        for (const TvmConfig& c : order) {
            if (pass == 0 && ((force_v && c.v != force_v) || (force_ctas && c.ctas != force_ctas))) continue;
            // do not use a tile (much) bigger than the plane
            if (c.v > 1 && n <= (unsigned long long)512 * c.v) continue;
            const size_t budget = (size_t)(227 * 1024) / (size_t)c.ctas - 1024 - 128;
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
extern "C" int32_t kc_debug_last_tile_config(int32_t* v, int32_t* ctas, int32_t* stages) {
    if (v) *v = g_kc_last_tile_config[0];
    if (ctas) *ctas = g_kc_last_tile_config[1];
    if (stages) *stages = g_kc_last_tile_config[2];
    return KC_OK;
}

int32_t kck_launch_tape(kc_context* ctx, const KcTapeArgs& args) {
    if (args.n == 0 || args.n_seg == 0) return KC_OK;
    int ns_max = 0;
    for (uint32_t s = 0; s < args.n_seg; ++s) ns_max = std::max<int>(ns_max, (int)args.seg[s].n_src);
    const int nt_max = (int)args.variant;  // temporaries the tapes touch (set by the planner)
    const TvmConfig c = pick_config(ns_max, nt_max, args.n);
    g_kc_last_tile_config[0] = c.v; g_kc_last_tile_config[1] = c.ctas; g_kc_last_tile_config[2] = c.stages;
    KcTimed timed(ctx, KC_KERNEL_TAPE);
    int32_t rc;
    const bool exact = ctx->opts.math_mode == KC_MATH_EXACT;
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
    const int grid = grid_for(ctx, (n + 3) >> 2, 256, 5);   // 5 CTAs of 256 threads are resident per SM: one wave
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
