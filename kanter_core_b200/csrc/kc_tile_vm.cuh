// kc_tile_vm.cuh -- device code of the fused elementwise kernel: arithmetic helpers (glibc-exact
// and fast pow, RGBA8 conversion), the TMA/mbarrier tile pipeline, and the tape interpreter.
// Included by kc_kernels.cu (nvcc) and, as text, by the NVRTC-specialised kernels of kc_jit.cu.
// Nothing in here may depend on host headers.
#pragma once
#include "kc_tape.h"

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
    if (!EXACT) {
        // FAST (+-1 LSB of the exact bytes): branch-free, reciprocal multiplies, and the curve's argument
        // ((c + .055) / 1.055 in (0.09, 1]) needs none of the general pow's range handling: 2^(2.4 * lg2 t)
        const float t = (c + 0.055f) * (1.0f / 1.055f);
        float lg, p;
        asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(lg) : "f"(t));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(p) : "f"(2.4f * lg));
        l = c <= 0.04045f ? c * (1.0f / 12.92f) : p;    // NaN compares false and comes out of lg2/ex2 as NaN
    } else if (c <= 0.0f) l = c;
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

// The kernel body, parameterised by the PROGRAM that turns one tile into results: the tape
// interpreter below (ahead-of-time kernels) or straight-line code generated for one tape
// (kc_jit.cu).  PROG::tile<EXACT, V>(ctx) is called once per (segment, tile) with the sources in
// shared memory.
template <int V>
struct KcTileCtx {
    const KcTapeArgs& A;
    const KcSegment& G;
    uint32_t seg;
    uint32_t src_addr, tmp_addr;   // shared-memory byte addresses of this thread's first float4 of source 0 / temporary 0
    float* tmp_base;
    int tid;
    unsigned long long px0, rem;   // first pixel of the tile, pixels from there to the end of the plane
    bool full;                     // rem >= tile size
};

// ---- the tape interpreter -----------------------------------------------------------------
struct KcInterp {
    template <bool EXACT, int V>
    static __device__ __forceinline__ void tile(const KcTileCtx<V>& c) {
        constexpr int TILE_PX = 1024 * V;
        const KcTapeArgs& A = c.A;
        const KcSegment& G = c.G;
        const int tid = c.tid;
        const unsigned long long px0 = c.px0, rem = c.rem;
        const bool full = c.full;
        float* tmp_base = c.tmp_base;
        // ---- interpret the segment's tape over this tile ---------------------------------
        // The launch code resolved every operand to a byte offset (KC_R_* below), so decoding an
        // instruction is: constant-bank load, mask, one add, LDS.128.  The cheap ops update the
        // accumulator in place (inline PTX with "+f" operands): the compiler then keeps ONE copy
        // of acc across the switch arms instead of shuffling it through phi moves.
        float4 acc[V];
#pragma unroll
        for (int j = 0; j < V; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        const uint32_t src_addr = c.src_addr, tmp_addr = c.tmp_addr;
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
    }
};

// ---- building blocks of the programs kc_jit.cu generates (straight-line code per tape) ---------
template <bool EXACT, int OP>
__device__ __forceinline__ void kc_apply(float4& acc, const float4& x) {
    if (OP == TOP_LD) acc = x;
    else if (OP == TOP_ADD) acc = KC_LANES(__fadd_rn, acc, x);
    else if (OP == TOP_SUB) acc = KC_LANES(__fsub_rn, acc, x);
    else if (OP == TOP_RSUB) acc = KC_LANES_R(__fsub_rn, acc, x);
    else if (OP == TOP_MUL) acc = KC_LANES(__fmul_rn, acc, x);
    else if (OP == TOP_DIV) acc = KC_LANES(__fdiv_rn, acc, x);
    else if (OP == TOP_RDIV) acc = KC_LANES_R(__fdiv_rn, acc, x);
    else if (OP == TOP_POW) acc = kc_pow4<EXACT>(acc, x);
    else acc = kc_pow4<EXACT>(x, acc);   // TOP_RPOW
}
template <int V>
__device__ __forceinline__ void kc_store_out(float* o, int tid, unsigned long long rem, bool full, const float4 (&acc)[V]) {
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
}
__device__ __forceinline__ void kc_store_px4(uint32_t* o, unsigned long long p, unsigned long long rem, const uint32_t (&px)[4]) {
    if (p + 3 < rem) __stcs(reinterpret_cast<uint4*>(o + p), make_uint4(px[0], px[1], px[2], px[3]));
    else if (p < rem) {
        o[p] = px[0];
        if (p + 1 < rem) o[p + 1] = px[1];
        if (p + 2 < rem) o[p + 2] = px[2];
    }
}
template <bool EXACT, int V>
__device__ __forceinline__ void kc_pack_rgba(uint32_t* o, bool srgb, const float4 (&t0)[V], const float4 (&t1)[V], const float4 (&t2)[V],
                                             const float4 (&acc)[V], int tid, unsigned long long rem) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float r[4] = {t0[j].x, t0[j].y, t0[j].z, t0[j].w};
        const float g[4] = {t1[j].x, t1[j].y, t1[j].z, t1[j].w};
        const float b[4] = {t2[j].x, t2[j].y, t2[j].z, t2[j].w};
        const float a[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
        uint32_t px[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            uint32_t R, Gc, B;
            if (srgb) { R = kc_to_u8_srgb<EXACT>(r[l]); Gc = kc_to_u8_srgb<EXACT>(g[l]); B = kc_to_u8_srgb<EXACT>(b[l]); }
            else { R = kc_to_u8(r[l]); Gc = kc_to_u8(g[l]); B = kc_to_u8(b[l]); }
            px[l] = R | (Gc << 8) | (B << 16) | (kc_to_u8(a[l]) << 24);
        }
        kc_store_px4(o, 4ull * (unsigned long long)(j * TVM_THREADS + tid), rem, px);
    }
}
template <bool EXACT, int V>
__device__ __forceinline__ void kc_pack_gray(uint32_t* o, bool srgb, const float4 (&acc)[V], int tid, unsigned long long rem) {
#pragma unroll
    for (int j = 0; j < V; ++j) {
        const float v[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
        uint32_t px[4];
#pragma unroll
        for (int l = 0; l < 4; ++l) {
            const uint32_t u = srgb ? kc_to_u8_srgb<EXACT>(v[l]) : kc_to_u8(v[l]);
            px[l] = u | (u << 8) | (u << 16) | 0xff000000u;
        }
        kc_store_px4(o, 4ull * (unsigned long long)(j * TVM_THREADS + tid), rem, px);
    }
}

template <bool EXACT, int V, class PROG>
__device__ __forceinline__ void kc_tile_vm_body(const KcTapeArgs& A, int stages, int ns_max, uint32_t tiles_per_plane, uint32_t total_work) {
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
        {
            const KcTileCtx<V> tc{A, G, seg, smem_u32(sbuf) + (uint32_t)tid * 16u, smem_u32(tmp_base) + (uint32_t)tid * 16u,
                                  tmp_base, tid, px0, rem, full};
            PROG::template tile<EXACT, V>(tc);
        }
        __syncthreads();  // the stage (and the temporaries) may be overwritten from here on
    }
}

