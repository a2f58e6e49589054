// kc_resize.cu — the implicit resize pre-pass of every node.
//
// Replaces resize_buffers' per-plane call into the third-party image crate,
// `image::imageops::resize(plane, w, h, filter)` (src/shared.rs:155-201; image
// 0.24.0 per Cargo.lock:237-240, imageops/sample.rs): a separable resampler
// that runs the VERTICAL pass first into an unclamped f32 intermediate and the
// HORIZONTAL pass second, clamping to [0,1]; per output index the tap window
// is [floor(c - s), ceil(c + s)) clipped to the source, the kernel is evaluated
// at (i - (c - 0.5)) / sratio and the taps are renormalised by their f32 sum.
//
// The per-axis tap tables are computed on the host in f32 with glibc
// sinf/expf (what Rust's f32::sin/exp call), so they are bit-identical to the
// reference's; the device kernels accumulate taps left to right with separate
// multiply and add roundings (EXACT) or FMA (FAST).
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "kc_internal.h"

// ---------------------------------------------------------------------------
// host: filter kernels and tap tables
// ---------------------------------------------------------------------------
namespace {

const float kPi = 3.14159265358979323846f;  // f32::consts::PI

float h_sinc(float t) {
    const float a = t * kPi;
    return t == 0.0f ? 1.0f : sinf(a) / a;
}

float h_filter(int filter, float x) {
    switch (filter) {
        case KC_FILTER_NEAREST:  // box_kernel
            return 1.0f;
        case KC_FILTER_TRIANGLE:
            return fabsf(x) < 1.0f ? 1.0f - fabsf(x) : 0.0f;
        case KC_FILTER_CATMULL_ROM: {  // bc_cubic_spline(x, b = 0, c = 0.5)
            const float b = 0.0f, c = 0.5f;
            const float a = fabsf(x);
            float k;
            if (a < 1.0f)
                k = (12.0f - 9.0f * b - 6.0f * c) * (a * a * a) + (-18.0f + 12.0f * b + 6.0f * c) * (a * a) + (6.0f - 2.0f * b);
            else if (a < 2.0f)
                k = (-b - 6.0f * c) * (a * a * a) + (6.0f * b + 30.0f * c) * (a * a) + (-12.0f * b - 48.0f * c) * a + (8.0f * b + 24.0f * c);
            else
                k = 0.0f;
            return k / 6.0f;
        }
        case KC_FILTER_GAUSSIAN: {  // gaussian(x, r = 0.5)
            const float r = 0.5f;
            return (1.0f / (sqrtf(2.0f * kPi) * r)) * expf(-(x * x) / (2.0f * (r * r)));
        }
        default:  // lanczos(x, 3)
            return fabsf(x) < 3.0f ? h_sinc(x) * h_sinc(x / 3.0f) : 0.0f;
    }
}

float h_support(int filter) {
    switch (filter) {
        case KC_FILTER_NEAREST: return 0.0f;
        case KC_FILTER_TRIANGLE: return 1.0f;
        case KC_FILTER_CATMULL_ROM: return 2.0f;
        default: return 3.0f;  // Gaussian, Lanczos3
    }
}

}  // namespace

void kc_resize_axis_host(uint32_t src_len, uint32_t dst_len, int filter, std::vector<uint32_t>& left,
                         std::vector<uint32_t>& count, std::vector<float>& weights, uint32_t& max_taps) {
    const float ratio = (float)src_len / (float)dst_len;
    const float sratio = ratio < 1.0f ? 1.0f : ratio;
    const float src_support = h_support(filter) * sratio;
    left.resize(dst_len);
    count.resize(dst_len);
    std::vector<std::vector<float>> ws(dst_len);
    max_taps = 1;
    for (uint32_t o = 0; o < dst_len; ++o) {
        float c = ((float)o + 0.5f) * ratio;
        int64_t l = (int64_t)floorf(c - src_support);
        l = std::min<int64_t>(std::max<int64_t>(l, 0), (int64_t)src_len - 1);
        int64_t r = (int64_t)ceilf(c + src_support);
        r = std::min<int64_t>(std::max<int64_t>(r, l + 1), (int64_t)src_len);
        c -= 0.5f;
        float sum = 0.0f;
        std::vector<float>& w = ws[o];
        for (int64_t i = l; i < r; ++i) {
            const float v = h_filter(filter, ((float)i - c) / sratio);
            w.push_back(v);
            sum += v;
        }
        for (float& v : w) v /= sum;
        left[o] = (uint32_t)l;
        count[o] = (uint32_t)w.size();
        max_taps = std::max<uint32_t>(max_taps, (uint32_t)w.size());
    }
    weights.assign((size_t)dst_len * max_taps, 0.0f);
    for (uint32_t o = 0; o < dst_len; ++o)
        for (size_t i = 0; i < ws[o].size(); ++i) weights[(size_t)o * max_taps + i] = ws[o][i];
}

namespace {

int32_t get_axis(kc_context* ctx, uint32_t src_len, uint32_t dst_len, int filter, std::shared_ptr<KcAxisTable>& out) {
    auto key = std::make_tuple(src_len, dst_len, filter);
    auto it = ctx->axis_tables.find(key);
    if (it != ctx->axis_tables.end()) {
        out = it->second;
        return KC_OK;
    }
    if (ctx->capturing) KC_FAIL(KC_ERR_GENERIC, "a tap table would have to be uploaded during a stream capture");
    auto t = std::make_shared<KcAxisTable>();
    t->src_len = src_len;
    t->dst_len = dst_len;
    t->filter = filter;
    kc_resize_axis_host(src_len, dst_len, filter, t->h_left, t->h_count, t->h_weights, t->max_taps);
    // tap-major copy for the device: weights[tap][o], so that consecutive
    // output indices read consecutive addresses
    std::vector<float> wt((size_t)t->max_taps * dst_len);
    for (uint32_t o = 0; o < dst_len; ++o)
        for (uint32_t k = 0; k < t->max_taps; ++k) wt[(size_t)k * dst_len + o] = t->h_weights[(size_t)o * t->max_taps + k];
    KC_CUDA(cudaMalloc((void**)&t->d_left, sizeof(uint32_t) * dst_len));
    KC_CUDA(cudaMalloc((void**)&t->d_count, sizeof(uint32_t) * dst_len));
    KC_CUDA(cudaMalloc((void**)&t->d_weights, sizeof(float) * wt.size()));
    // pageable sources: these copies complete before the vectors go out of scope
    KC_CUDA(cudaMemcpyAsync(t->d_left, t->h_left.data(), sizeof(uint32_t) * dst_len, cudaMemcpyHostToDevice, ctx->stream));
    KC_CUDA(cudaMemcpyAsync(t->d_count, t->h_count.data(), sizeof(uint32_t) * dst_len, cudaMemcpyHostToDevice, ctx->stream));
    KC_CUDA(cudaMemcpyAsync(t->d_weights, wt.data(), sizeof(float) * wt.size(), cudaMemcpyHostToDevice, ctx->stream));
    {
        std::vector<uint32_t> vt((size_t)(2 + t->max_taps) * dst_len);
        memcpy(vt.data(), t->h_left.data(), sizeof(uint32_t) * dst_len);
        memcpy(vt.data() + dst_len, t->h_count.data(), sizeof(uint32_t) * dst_len);
        memcpy(vt.data() + 2 * (size_t)dst_len, wt.data(), sizeof(float) * wt.size());
        KC_CUDA(cudaMalloc((void**)&t->d_vtab, sizeof(uint32_t) * vt.size()));
        KC_CUDA(cudaMemcpyAsync(t->d_vtab, vt.data(), sizeof(uint32_t) * vt.size(), cudaMemcpyHostToDevice, ctx->stream));
        KC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    KC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->axis_tables[key] = t;
    out = t;
    return KC_OK;
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library does not link libcuda)
PFN_cuTensorMapEncodeTiled tensor_map_encoder() {
    static PFN_cuTensorMapEncodeTiled fn = []() -> PFN_cuTensorMapEncodeTiled {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (PFN_cuTensorMapEncodeTiled)p;
    }();
    return fn;
}

// a row-major [rows][cols] array of 32-bit words with a [box_rows][box_cols] box; out-of-bounds elements read as zero
bool make_tensor_map_2d(CUtensorMap* m, const void* base, uint64_t cols, uint64_t rows, uint32_t box_cols, uint32_t box_rows) {
    PFN_cuTensorMapEncodeTiled enc = tensor_map_encoder();
    if (!enc) return false;
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {cols * 4};
    const cuuint32_t box[2] = {box_cols, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    return enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
               CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <bool EXACT>
__device__ __forceinline__ float tap(float acc, float s, float w) {
    return EXACT ? __fadd_rn(acc, __fmul_rn(s, w)) : fmaf(s, w, acc);
}

// the same for two sums at once (one FFMA2).  EXACT: the product and the sum round separately, as in the reference -- the
// sum as acc * one + product with `one` a kernel argument: ptxas contracts __fadd2_rn(acc, __fmul2_rn(s, w)) into one FFMA2
// (seen as 1-ulp differences against the oracle), which it cannot do when the multiplier is not a compile-time 1
template <bool EXACT>
__device__ __forceinline__ float2 tap2(float2 acc, float2 s, float2 w, float one) {
    return EXACT ? __ffma2_rn(acc, make_float2(one, one), __fmul2_rn(s, w)) : __ffma2_rn(s, w, acc);
}

// vertical_sample: tmp[oy][x] = sum_i src[left[oy]+i][x] * wv[i][oy]   (no clamp)
// One thread per (x, oy); consecutive threads walk x, so every tap row is a
// coalesced read and the store is coalesced.
template <bool EXACT>
__global__ void __launch_bounds__(256) kc_resize_v_kernel(const float* __restrict__ src, uint32_t sw, float* __restrict__ tmp,
                                                          uint32_t dh, const uint32_t* __restrict__ left,
                                                          const uint32_t* __restrict__ count, const float* __restrict__ wv) {
    const uint32_t x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= sw) return;
    for (uint32_t oy = blockIdx.y; oy < dh; oy += gridDim.y) {
        const uint32_t l = left[oy], n = count[oy];
        float acc = 0.0f;
        for (uint32_t i = 0; i < n; ++i) acc = tap<EXACT>(acc, __ldg(src + (size_t)(l + i) * sw + x), wv[(size_t)i * dh + oy]);
        tmp[(size_t)oy * sw + x] = acc;
    }
}

// horizontal_sample: dst[y][ox] = clamp(sum_j tmp[y][left[ox]+j] * wh[j][ox], 0, 1)
template <bool EXACT>
__global__ void __launch_bounds__(256) kc_resize_h_kernel(const float* __restrict__ tmp, uint32_t sw, float* __restrict__ dst,
                                                          uint32_t dw, uint32_t dh, const uint32_t* __restrict__ left,
                                                          const uint32_t* __restrict__ count, const float* __restrict__ wh, float clo, float chi) {
    const uint32_t ox = blockIdx.x * blockDim.x + threadIdx.x;
    if (ox >= dw) return;
    const uint32_t l = left[ox], n = count[ox];
    for (uint32_t y = blockIdx.y; y < dh; y += gridDim.y) {
        const float* row = tmp + (size_t)y * sw + l;
        float acc = 0.0f;
        for (uint32_t j = 0; j < n; ++j) acc = tap<EXACT>(acc, __ldg(row + j), wh[(size_t)j * dw + ox]);
        // image::math::utils::clamp keeps NaN
        acc = acc < clo ? clo : (acc > chi ? chi : acc);
        __stcs(dst + (size_t)y * dw + ox, acc);
    }
}

// ---------------------------------------------------------------------------
// Fused V∘H strip kernel for resizes whose tap windows are short (<= FS_MAXT taps
// per axis: every upsampling, and mild downsampling).
//
// A CTA owns a strip of FS_TW output columns and marches down the image in
// groups of FS_G output rows.  What depends only on the column -- the window
// and the horizontal weights of each thread's four output columns -- is loaded
// once and stays in registers for the whole march.  Per group:
//   (1) the source rows the group needs and the group's vertical taps arrive in
//       shared memory by cp.async, issued one group ahead (double buffered), so
//       their latency hides behind the previous group's arithmetic;
//   (2) vertical pass (the reference's unclamped f32 `tmp`, never in HBM): one
//       item = four adjacent source columns (LDS.128) x one output row, two packed
//       FFMA2 per tap (FMUL2 + FFMA2 in EXACT), result stored column-major;
//   (3) horizontal pass: each thread accumulates 4 columns x 16 rows in
//       registers, reading the intermediate four rows at a time (LDS.128) and
//       feeding row pairs to FFMA2 with the tap weight as the broadcast scalar;
//       when the four columns share one window (always, for integer upsampling
//       ratios >= 4) every loaded value is used by all four columns;
//   (4) clamp to [0,1] keeping NaN (image::math::utils::clamp), float4 streaming
//       stores: one full 2 KiB row segment per CTA per row.
// Two __syncthreads per 8192 output pixels.  Same operation order as the
// two-pass kernels: vertical taps top to bottom, horizontal taps left to right
// starting from +0, clamp last.
// ---------------------------------------------------------------------------
constexpr int FS_DEFAULT_THREADS = 128;       // threads per CTA unless KC_RESIZE_THREADS says otherwise
constexpr int FS_CPT = 4;                    // consecutive output columns per thread (one float4 store per row)
constexpr int FS_G = 16;                     // output rows per group
constexpr int FS_MAXT = 8;                   // taps per axis this kernel supports
constexpr int FS_NE = 6;                     // patch elements per thread kept as a precomputed list
constexpr int FS_TP = FS_G + 4;              // pitch of the column-major intermediate (16-byte aligned rows)

__device__ __forceinline__ void cp_async4(void* smem, const void* gmem) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// One tap on a pair of pixels.  FAST: FFMA2.  EXACT: the reference rounds the product and
// the sum separately.  ptxas (12.9, -O1 and up) contracts mul.rn.f32x2 + add.rn.f32x2 -- and
// even fma(s,w,-0) + fma(acc,1,p) -- into ONE FFMA2 regardless of --fmad=false, so the add is
// written as fma(acc, one, p) with `one` == 1.0f arriving as a kernel argument the optimiser
// cannot see through: rn(acc*1 + p) == rn(acc + p) bit for bit, signed zeros included.
template <bool EXACT>
__device__ __forceinline__ float2 tap2(float2 acc, float2 s, float w, float one) {
    const float2 ww = make_float2(w, w);
    if (EXACT) return __ffma2_rn(acc, make_float2(one, one), __fmul2_rn(s, ww));
    return __ffma2_rn(s, ww, acc);
}

// clamp to [lo, hi] = [0, 1]; NaN stays NaN (the reference's clamp is two comparisons).  The bounds are kernel
// arguments: (-inf, +inf) when the caller switched the clamp off (kc_options.resize_unclamped), which makes both no-ops.
__device__ __forceinline__ float clamp_keep_nan(float a, float lo, float hi) {
    float r;
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(hi));
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(r) : "f"(r), "f"(lo));
    return r;
}

struct FsGroupBuf {          // one prefetch stage, in shared memory
    float wv[FS_MAXT][FS_G];  // vertical weights, tap-major
    uint32_t vl[FS_G];        // first source row of each output row (absolute)
    uint32_t vc[FS_G];        // tap count of each output row (0: row past the image)
};

// N vertical taps of four adjacent source columns, top to bottom
template <bool EXACT, int N>
__device__ __forceinline__ void fs_vtaps(const float* sp, uint32_t pitch, const float* wc, float one, float2& a0, float2& a1) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(sp + k * pitch);
        const float wk = wc[k * FS_G];
        a0 = tap2<EXACT>(a0, make_float2(v.x, v.y), wk, one);
        a1 = tap2<EXACT>(a1, make_float2(v.z, v.w), wk, one);
    }
}

template <bool EXACT, int THREADS>
__global__ void __launch_bounds__(THREADS, 512 / THREADS) kc_resize_strip_kernel(
    const float* __restrict__ src, uint32_t sw, uint32_t sh, float* __restrict__ dst, uint32_t dw, uint32_t dh,
    const uint32_t* __restrict__ vleft, const uint32_t* __restrict__ vcount, const float* __restrict__ vw, uint32_t vtaps,
    const uint32_t* __restrict__ hleft, const uint32_t* __restrict__ hcount, const float* __restrict__ hw,
    uint32_t pcols, uint32_t prows, float one, uint32_t row0, uint32_t nrows, float clo, float chi) {
    // rows [row0, row0 + nrows) of the dw x dh result are produced; dst holds just those rows
    extern __shared__ __align__(16) float fsm[];
    float* Tm = fsm;                                              // [pcols][FS_TP] vertical-pass result, column-major
    float* Sbuf = Tm + (size_t)pcols * FS_TP;                     // [2][prows][pcols] source rows of a group
    float4* Wh = reinterpret_cast<float4*>(Sbuf + 2 * (size_t)prows * pcols);         // [FS_MAXT][THREADS]: tap j of a thread's 4 columns
    FsGroupBuf* gb = reinterpret_cast<FsGroupBuf*>(Wh + FS_MAXT * THREADS);         // [2]
    const int tid = threadIdx.x;
    const uint32_t ox0 = blockIdx.x * (THREADS * FS_CPT);
    const uint32_t oxl = min(ox0 + (THREADS * FS_CPT), dw) - 1;                // last valid column of the strip
    const uint32_t ngroups = (nrows + FS_G - 1) / FS_G;

    // ---- per-column state, loaded once --------------------------------------------------
    // (the horizontal weights live in shared memory, one float4 per tap per thread, read back
    // right before use: 32 fewer live registers buys a fourth resident CTA per SM)
    uint32_t left[FS_CPT], cnt[FS_CPT];
    const uint32_t cx0 = __ldg(hleft + ox0), cx1 = __ldg(hleft + oxl) + __ldg(hcount + oxl);
    const uint32_t ncx = cx1 - cx0;                               // source columns the strip reads
#pragma unroll
    for (int c = 0; c < FS_CPT; ++c) {
        const uint32_t ox = min(ox0 + FS_CPT * tid + c, oxl);
        left[c] = __ldg(hleft + ox) - cx0;
        cnt[c] = __ldg(hcount + ox);
#pragma unroll
        for (int j = 0; j < FS_MAXT; ++j)
            reinterpret_cast<float*>(Wh + j * THREADS + tid)[c] = __ldg(hw + (size_t)min((uint32_t)j, cnt[c] - 1) * dw + ox);
    }
    const bool shared_window = left[0] == left[1] && left[0] == left[2] && left[0] == left[3] &&
                               cnt[0] == cnt[1] && cnt[0] == cnt[2] && cnt[0] == cnt[3];
    const uint32_t oxt = ox0 + FS_CPT * tid;
    const bool col_live = oxt <= oxl;
    const bool vec = ((dw & 3u) == 0) && (oxt + 3 <= oxl);

    // The source patch of a group is always `prows` rows x `ncx` columns (first row pulled up
    // where the window would run past the image), so the patch elements a thread copies are
    // the same for every group: (row, column) of its first FS_NE elements are worked out once.
    const uint32_t npatch = prows * ncx;
    uint32_t soff[FS_NE], goff[FS_NE];    // shared-memory / source offsets (in floats); soff == ~0u: no element
#pragma unroll
    for (int i = 0; i < FS_NE; ++i) {
        const uint32_t e = tid + THREADS * i;
        const uint32_t r = e / ncx, c = e - r * ncx;
        soff[i] = e < npatch ? r * pcols + c : 0xffffffffu;
        goff[i] = r * sw + c;             // a patch spans < 2^32 source floats
    }
    // first source row of the patch of group g
    auto group_row0 = [&](uint32_t g) { return min(__ldg(vleft + row0 + g * FS_G), sh - prows); };
    // cp.async everything group g needs into stage b
    auto prefetch = [&](uint32_t g, int b, uint32_t ry0) {
        const uint32_t ly0 = g * FS_G;                            // first row of the group inside the window
        FsGroupBuf& G = gb[b];
        for (int i = tid; i < FS_MAXT * FS_G; i += THREADS) {
            const int k = i / FS_G, r = i % FS_G;
            const bool live = ly0 + r < nrows;
            const uint32_t oy = row0 + ly0 + r;                   // row of the full result: indexes the tap tables
            if (live && (uint32_t)k < vtaps) cp_async4(&G.wv[k][r], vw + (size_t)k * dh + oy);
            if (k == 0) {
                if (live) cp_async4(&G.vl[r], vleft + oy);
                else G.vl[r] = ry0;
            } else if (k == 1) {
                if (live) cp_async4(&G.vc[r], vcount + oy);
                else G.vc[r] = 0u;                                // rows past the image compute nothing
            }
        }
        float* S = Sbuf + (size_t)b * prows * pcols;
        const float* base = src + (size_t)ry0 * sw + cx0;
#pragma unroll
        for (int i = 0; i < FS_NE; ++i) {
            if (soff[i] != 0xffffffffu) cp_async4(S + soff[i], base + goff[i]);
        }
        for (uint32_t e = tid + THREADS * FS_NE; e < npatch; e += THREADS) {   // wide windows only
            const uint32_t r = e / ncx, c = e - r * ncx;
            cp_async4(S + r * pcols + c, base + (size_t)r * sw + c);
        }
    };

    uint32_t g = blockIdx.y;
    if (g >= ngroups) return;
    uint32_t ry0 = group_row0(g);
    prefetch(g, 0, ry0);
    cp_async_commit();
    int b = 0;
    for (; g < ngroups; g += gridDim.y) {
        const uint32_t gn = g + gridDim.y;
        uint32_t nry0 = 0;
        if (gn < ngroups) nry0 = group_row0(gn);                  // consumed after the vertical pass
        cp_async_wait_all();
        __syncthreads();                                          // stage b landed; Tm free again
        const FsGroupBuf& G = gb[b];
        const float* S = Sbuf + (size_t)b * prows * pcols;
        // ---- vertical pass: Tm[c][r] = sum_k S[vl[r]-ry0+k][c] * wv[k][r], two columns per item ----
        const uint32_t nquad = (ncx + 3) >> 2;
        for (uint32_t i = tid; i < nquad * FS_G; i += THREADS) {
            const uint32_t cq = i / FS_G, r = i % FS_G;
            const uint32_t n = G.vc[r];
            const float* sp = S + (size_t)(G.vl[r] - ry0) * pcols + 4 * cq;
            float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
            const float* wc = &G.wv[0][r];
            switch (n) {                                         // straight-line code per tap count
                case 1: fs_vtaps<EXACT, 1>(sp, pcols, wc, one, a0, a1); break;
                case 2: fs_vtaps<EXACT, 2>(sp, pcols, wc, one, a0, a1); break;
                case 3: fs_vtaps<EXACT, 3>(sp, pcols, wc, one, a0, a1); break;
                case 4: fs_vtaps<EXACT, 4>(sp, pcols, wc, one, a0, a1); break;
                case 5: fs_vtaps<EXACT, 5>(sp, pcols, wc, one, a0, a1); break;
                case 6: fs_vtaps<EXACT, 6>(sp, pcols, wc, one, a0, a1); break;
                case 7: fs_vtaps<EXACT, 7>(sp, pcols, wc, one, a0, a1); break;
                case 8: fs_vtaps<EXACT, 8>(sp, pcols, wc, one, a0, a1); break;
                default: break;                                  // 0: row past the image
            }
            float* t = Tm + (size_t)(4 * cq) * FS_TP + r;
            t[0] = a0.x;
            t[FS_TP] = a0.y;
            t[2 * FS_TP] = a1.x;
            t[3 * FS_TP] = a1.y;
        }
        if (gn < ngroups) prefetch(gn, b ^ 1, nry0);
        cp_async_commit();
        __syncthreads();                                          // Tm complete
        // ---- horizontal pass -----------------------------------------------------------------
        if (col_live) {
            float2 acc[FS_CPT][FS_G / 2];
#pragma unroll
            for (int c = 0; c < FS_CPT; ++c)
#pragma unroll
                for (int q = 0; q < FS_G / 2; ++q) acc[c][q] = make_float2(0.0f, 0.0f);
            if (shared_window) {
                const float4* t = reinterpret_cast<const float4*>(Tm + (size_t)left[0] * FS_TP);
#pragma unroll
                for (int j = 0; j < FS_MAXT; ++j) {                // (a switch over straight-line variants per tap count was slower: spills)
                    if ((uint32_t)j < cnt[0]) {
                        const float4 w4 = Wh[j * THREADS + tid];
                        const float w[FS_CPT] = {w4.x, w4.y, w4.z, w4.w};
                        float4 v[FS_G / 4];
#pragma unroll
                        for (int q = 0; q < FS_G / 4; ++q) v[q] = t[j * (FS_TP / 4) + q];
#pragma unroll
                        for (int c = 0; c < FS_CPT; ++c)
#pragma unroll
                            for (int q = 0; q < FS_G / 4; ++q) {
                                acc[c][2 * q] = tap2<EXACT>(acc[c][2 * q], make_float2(v[q].x, v[q].y), w[c], one);
                                acc[c][2 * q + 1] = tap2<EXACT>(acc[c][2 * q + 1], make_float2(v[q].z, v[q].w), w[c], one);
                            }
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < FS_CPT; ++c) {
                    const float4* t = reinterpret_cast<const float4*>(Tm + (size_t)left[c] * FS_TP);
#pragma unroll
                    for (int j = 0; j < FS_MAXT; ++j) {
                        if ((uint32_t)j < cnt[c]) {
                            const float wj = reinterpret_cast<const float*>(Wh + j * THREADS + tid)[c];
#pragma unroll
                            for (int q = 0; q < FS_G / 4; ++q) {
                                const float4 v = t[j * (FS_TP / 4) + q];
                                acc[c][2 * q] = tap2<EXACT>(acc[c][2 * q], make_float2(v.x, v.y), wj, one);
                                acc[c][2 * q + 1] = tap2<EXACT>(acc[c][2 * q + 1], make_float2(v.z, v.w), wj, one);
                            }
                        }
                    }
                }
            }
            const uint32_t oy0 = g * FS_G;
            const uint32_t nrow = min((uint32_t)FS_G, nrows - oy0);
            float* out = dst + (size_t)oy0 * dw + oxt;
            if (vec && nrow == (uint32_t)FS_G) {
                float4* o4 = reinterpret_cast<float4*>(out);
                const uint32_t dw4 = dw >> 2;
#pragma unroll
                for (int r = 0; r < FS_G; ++r) {
                    float v[FS_CPT];
#pragma unroll
                    for (int c = 0; c < FS_CPT; ++c) v[c] = clamp_keep_nan((r & 1) ? acc[c][r >> 1].y : acc[c][r >> 1].x, clo, chi);
                    __stcs(o4, make_float4(v[0], v[1], v[2], v[3]));
                    o4 += dw4;
                }
            } else {
#pragma unroll
                for (int r = 0; r < FS_G; ++r) {
                    if ((uint32_t)r < nrow) {
#pragma unroll
                        for (int c = 0; c < FS_CPT; ++c)
                            if (oxt + c <= oxl) out[(size_t)r * dw + c] = clamp_keep_nan((r & 1) ? acc[c][r >> 1].y : acc[c][r >> 1].x, clo, chi);
                    }
                }
            }
        }
        ry0 = nry0;
        b ^= 1;
    }
}


template <bool EXACT, int N, int WP>
__device__ __forceinline__ void ft_vtaps(const float* sp, uint32_t pitch, const float* wc, float one, float2& a0, float2& a1) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(sp + k * pitch);
        const float wk = wc[k * WP];
        a0 = tap2<EXACT>(a0, make_float2(v.x, v.y), wk, one);
        a1 = tap2<EXACT>(a1, make_float2(v.z, v.w), wk, one);
    }
}

// ---------------------------------------------------------------------------
// The same fused V∘H march with TMA on both sides (the default for every upsample of a plane whose
// widths are multiples of four): what changes is who moves the bytes.
//   in : per group ONE 2-D tensor-map load of the source patch (cp.async.bulk.tensor, UTMALDG: prows x pcols,
//        out-of-bounds columns arrive as zeros) and ONE of the group's slice of the vertical table
//        ([left | count | tap 0 .. tap n-1] x G rows), both completing on an mbarrier, issued one whole group
//        ahead by thread 0 -- no per-thread cp.async, no address arithmetic in the other 127 threads;
//   out: each warp leaves its clamped RC x 128 tile in shared memory (conflict-free STS.128) and its lane 0 hands it to
//        the TMA as a 2-D tensor store (UTMASTG) that clips at the right and bottom edges itself: no ragged-tail code,
//        no warp holding 64 accumulators while 16 STG.128 drain, and -- the tiles being double-buffered per warp --
//        no block-wide wait for a store: ONE __syncthreads per group (the intermediate is double-buffered too).
// The horizontal pass works on RC rows at a time (RC x 4 accumulators), the tap weights come from L1 (__ldg of the
// tap-major table: the same 16 KiB every group): ~half the registers of the kernel above, twice the resident warps.
// Arithmetic, operation order and clamp are those of kc_resize_strip_kernel, bit for bit.
// ---------------------------------------------------------------------------
constexpr int FT_THREADS = 128;
constexpr int FT_TW = FT_THREADS * FS_CPT;      // 512 output columns per CTA
constexpr int FT_WARP_COLS = 32 * FS_CPT;       // 128: the columns one warp computes (and, with WARP_STORE, stages and stores itself)
constexpr int FT_HALF = 256;                    // columns per block-wide tensor store (a box side is at most 256 elements)

__device__ __forceinline__ uint32_t ft_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void ft_mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(ft_smem(bar)), "r"(count));
}
__device__ __forceinline__ void ft_mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(ft_smem(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void ft_mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "FT_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra FT_DONE;\n"
        "bra FT_WAIT;\n"
        "FT_DONE:\n"
        "}\n" ::"r"(ft_smem(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void ft_tma_load_2d(void* dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint64_t* bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(ft_smem(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(ft_smem(bar))
                 : "memory");
}
__device__ __forceinline__ void ft_tma_store_2d(const CUtensorMap* map, uint32_t c0, uint32_t c1, const void* src, uint64_t policy) {
    // evict-first in L2, like the st.global.cs of the other kernels: the result is a pure stream nobody re-reads soon
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%1, %2}], [%3], %4;" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(ft_smem(src)), "l"(policy)
                 : "memory");
}

// N vertical taps of four adjacent source columns for FOUR consecutive output rows that share one window: each source
// float4 is loaded once for the four rows, their weights arrive as one float4 per tap (rows are adjacent in the table)
template <bool EXACT, int N, int WP>
__device__ __forceinline__ void ft_vtaps4(const float* sp, uint32_t pitch, const float* wc, float one, float2 (&a0)[4], float2 (&a1)[4]) {
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(sp + k * pitch);
        const float4 w4 = *reinterpret_cast<const float4*>(wc + k * WP);
        const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            a0[j] = tap2<EXACT>(a0[j], make_float2(v.x, v.y), w[j], one);
            a1[j] = tap2<EXACT>(a1[j], make_float2(v.z, v.w), w[j], one);
        }
    }
}

// clamp to [lo, hi]; `sat`: the bounds are [0, 1] and a is known to be a number, so one saturating add does it
__device__ __forceinline__ float ft_clamp(float a, float lo, float hi, bool sat) {
    if (sat) {
        float r;
        asm("add.rn.sat.f32 %0, %1, 0f00000000;" : "=f"(r) : "f"(a));
        return r;
    }
    return clamp_keep_nan(a, lo, hi);
}

// the planes of one launch (blockIdx.z picks one): an RGBA image resizes in ONE launch, with the same tables for every plane
constexpr int FT_MAX_PLANES = 4;
struct FtPlaneMaps {
    CUtensorMap src[FT_MAX_PLANES];
    CUtensorMap dst[FT_MAX_PLANES];
};

struct FtLayout {   // byte offsets into dynamic shared memory (host and device agree through this)
    uint32_t tm, tm_buf, s, s_stage, vt, vt_stage, o, o_buf, total;
    __host__ __device__ FtLayout(uint32_t pcols, uint32_t prows, uint32_t vrows, int G, int RC, bool warp_store) {
        auto up = [](uint32_t x) { return (x + 127u) & ~127u; };
        uint32_t p = 128;                                   // two mbarriers live in the first 16 bytes
        tm_buf = up(pcols * (uint32_t)(G + 4) * 4u);
        tm = p; p += 2 * tm_buf;                            // the intermediate of this group and of the previous one
        s_stage = up(prows * pcols * 4u);
        s = p; p += 2 * s_stage;
        vt_stage = up(vrows * (uint32_t)(G + 4) * 4u);      // G + 4 columns: the slice starts at the 16-byte boundary below its first row
        vt = p; p += 2 * vt_stage;
        if (warp_store) {
            o_buf = (uint32_t)RC * FT_WARP_COLS * 4u;       // one warp's RC x 128 result tile
            o = p; p += (FT_THREADS / 32) * 2u * o_buf;     // two per warp
        } else {
            o_buf = (uint32_t)G * FT_HALF * 4u;             // half of the block's G x 512 result tile: one tensor store
            o = p; p += 2u * o_buf;
        }
        total = p;
    }
};

template <bool EXACT, int G, int RC, int MINB, bool WARP_STORE>
__global__ void __launch_bounds__(FT_THREADS, MINB) kc_resize_tma_kernel(
    const __grid_constant__ FtPlaneMaps pm, const __grid_constant__ CUtensorMap tm_vtab,
    uint32_t sh, uint32_t dw, const uint32_t* __restrict__ vleft, uint32_t vtaps,
    const uint32_t* __restrict__ hleft, const uint32_t* __restrict__ hcount, const float* __restrict__ hw,
    uint32_t pcols, uint32_t prows, float one, uint32_t row0, uint32_t nrows, float clo, float chi) {
    static_assert(G % RC == 0 && RC % 4 == 0, "row chunks are whole float4s");
    constexpr int TP = G + 4;                                       // pitch of the column-major intermediate
    constexpr int VP = G + 4;                                       // pitch of the vertical-table slice (see issue())
    extern __shared__ __align__(128) unsigned char ftm[];
    const FtLayout L(pcols, prows, vtaps + 2, G, RC, WARP_STORE);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(ftm);
    // nonfinite[i % 3]: some thread saw a NaN or an infinity in the intermediate of the i-th group this block processes.  A group
    // without any takes the one-instruction clamp below.  Three flags: the one of group i + 2 is cleared between the barriers
    // of groups i and i + 1, when nobody can be writing it.
    volatile uint32_t* nonfinite = reinterpret_cast<volatile uint32_t*>(ftm + 16);
    const bool unit_clamp = clo == 0.0f && chi == 1.0f;              // image-0.24's clamp; (-inf, +inf) when the caller switched it off
    uint64_t policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const uint32_t ox0 = blockIdx.x * FT_TW;
    const uint32_t oxl = min(ox0 + FT_TW, dw) - 1;
    const uint32_t ngroups = (nrows + G - 1) / G;
    uint32_t g = blockIdx.y;
    if (g >= ngroups) return;

    if (tid == 0) {
        ft_mbar_init(&mbar[0], 1);
        ft_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        nonfinite[0] = nonfinite[1] = nonfinite[2] = 0u;
    }
    // ---- per-column state, loaded once ----
    // A tensor-map box must START on a 16-byte boundary of global memory (an unaligned innermost coordinate is an
    // illegal-instruction fault): the patch begins at the multiple of four columns at or below the strip's first source column
    const uint32_t cx0 = __ldg(hleft + ox0) & ~3u, cx1 = __ldg(hleft + oxl) + __ldg(hcount + oxl);
    const uint32_t ncx = cx1 - cx0;
    const uint32_t oxt = ox0 + FS_CPT * tid;
    const bool col_live = oxt <= oxl;                               // dw % 4 == 0: a live thread owns four real columns
    const bool warp_live = ox0 + (uint32_t)warp * FT_WARP_COLS <= oxl;
    uint32_t left[FS_CPT], cnt[FS_CPT];
#pragma unroll
    for (int c = 0; c < FS_CPT; ++c) {
        const uint32_t ox = min(oxt + c, oxl);
        left[c] = __ldg(hleft + ox) - cx0;
        cnt[c] = __ldg(hcount + ox);
    }
    const bool shared_window = left[0] == left[1] && left[0] == left[2] && left[0] == left[3] &&
                               cnt[0] == cnt[1] && cnt[0] == cnt[2] && cnt[0] == cnt[3];
    const uint32_t cmax = max(max(cnt[0], cnt[1]), max(cnt[2], cnt[3]));
    const float4* hw4 = reinterpret_cast<const float4*>(hw + (col_live ? oxt : ox0));   // tap j of my four columns: hw4[j * dw/4]
    const uint32_t dw4 = dw >> 2;
    const uint32_t stage_bytes = prows * pcols * 4u + (vtaps + 2) * (uint32_t)VP * 4u;
    float* Ow = reinterpret_cast<float*>(ftm + L.o + (uint32_t)warp * 2u * L.o_buf);   // this warp's two result tiles

    auto group_row0 = [&](uint32_t gg) { return min(__ldg(vleft + row0 + gg * G), sh - prows); };
    auto issue = [&](uint32_t gg, int b, uint32_t ry) {               // thread 0 only
        ft_mbar_expect_tx(&mbar[b], stage_bytes);
        ft_tma_load_2d(ftm + L.s + b * L.s_stage, &pm.src[blockIdx.z], cx0, ry, &mbar[b]);
        ft_tma_load_2d(ftm + L.vt + b * L.vt_stage, &tm_vtab, (row0 + gg * G) & ~3u, 0u, &mbar[b]);   // same rule: aligned start, G + 4 wide
    };
    __syncthreads();                                                  // the barriers are initialised
    uint32_t ry0 = group_row0(g);
    if (tid == 0) issue(g, 0, ry0);
    uint32_t phase = 0;                                               // bit b: parity stage b completes with next
    uint32_t chunk = 0;                                               // result tiles this warp has handed to the TMA
    uint32_t it = 0;                                                  // groups this block has processed, mod 3
    int b = 0;
    for (; g < ngroups; g += gridDim.y) {
        const uint32_t gn = g + gridDim.y;
        uint32_t nry0 = 0;
        if (gn < ngroups) {
            nry0 = group_row0(gn);
            if (tid == 0) issue(gn, b ^ 1, nry0);                     // stage b^1 was drained before the barrier of the previous group
        }
        ft_mbar_wait(&mbar[b], (phase >> b) & 1u);
        phase ^= 1u << b;
        const float* S = reinterpret_cast<const float*>(ftm + L.s + b * L.s_stage);
        const uint32_t* VT = reinterpret_cast<const uint32_t*>(ftm + L.vt + b * L.vt_stage) + ((row0 + g * G) & 3u);   // [2 + vtaps][VP], this group's first row
        const float* WV = reinterpret_cast<const float*>(VT + 2 * VP);
        float* Tm = reinterpret_cast<float*>(ftm + L.tm + b * L.tm_buf);   // the other one may still be read by a slower warp's horizontal pass
        const uint32_t live_rows = min((uint32_t)G, nrows - g * G);   // rows past the strip compute nothing (their table rows may be real)
        // ---- vertical pass: Tm[c][r] = sum_k S[vl[r]-ry0+k][c] * wv[k][r] ----
        const uint32_t nquad = (ncx + 3) >> 2;
        auto vrow = [&](uint32_t cq, uint32_t r) {                    // one output row of four source columns
            const uint32_t n = r < live_rows ? VT[VP + r] : 0u;
            const float* sp = S + (size_t)(VT[r] - ry0) * pcols + 4 * cq;
            float2 a0 = make_float2(0.0f, 0.0f), a1 = make_float2(0.0f, 0.0f);
            const float* wc = WV + r;
            switch (n) {
                case 1: ft_vtaps<EXACT, 1, VP>(sp, pcols, wc, one, a0, a1); break;
                case 2: ft_vtaps<EXACT, 2, VP>(sp, pcols, wc, one, a0, a1); break;
                case 3: ft_vtaps<EXACT, 3, VP>(sp, pcols, wc, one, a0, a1); break;
                case 4: ft_vtaps<EXACT, 4, VP>(sp, pcols, wc, one, a0, a1); break;
                case 5: ft_vtaps<EXACT, 5, VP>(sp, pcols, wc, one, a0, a1); break;
                case 6: ft_vtaps<EXACT, 6, VP>(sp, pcols, wc, one, a0, a1); break;
                case 7: ft_vtaps<EXACT, 7, VP>(sp, pcols, wc, one, a0, a1); break;
                case 8: ft_vtaps<EXACT, 8, VP>(sp, pcols, wc, one, a0, a1); break;
                default: break;
            }
            float* t = Tm + (size_t)(4 * cq) * TP + r;
            t[0] = a0.x;
            t[TP] = a0.y;
            t[2 * TP] = a1.x;
            t[3 * TP] = a1.y;
            // NaN-propagating maximum of the four magnitudes: anything but a finite number raises the group's flag
            float m0, m1;
            asm("max.NaN.f32 %0, %1, %2;" : "=f"(m0) : "f"(fabsf(a0.x)), "f"(fabsf(a0.y)));
            asm("max.NaN.f32 %0, %1, %2;" : "=f"(m1) : "f"(fabsf(a1.x)), "f"(fabsf(a1.y)));
            asm("max.NaN.f32 %0, %1, %2;" : "=f"(m0) : "f"(m0), "f"(m1));
            if (!(m0 <= 3.402823466e+38f)) nonfinite[it] = 1u;
        };
        if (((row0 + g * G) & 3u) == 0) {
            // items of four output rows: when they share one window (always, for integer ratios >= 4) each source value is
            // loaded once for the four and a tap's four weights are one LDS.128
            for (uint32_t i = tid; i < nquad * (G / 4); i += FT_THREADS) {
                const uint32_t cq = i / (G / 4), r = 4 * (i % (G / 4));
                const uint4 l4 = *reinterpret_cast<const uint4*>(VT + r), c4 = *reinterpret_cast<const uint4*>(VT + VP + r);
                const bool same = l4.x == l4.y && l4.x == l4.z && l4.x == l4.w && c4.x == c4.y && c4.x == c4.z && c4.x == c4.w && r + 3 < live_rows;
                if (!same) {
#pragma unroll 1
                    for (uint32_t j = 0; j < 4; ++j) vrow(cq, r + j);
                    continue;
                }
                const float* sp = S + (size_t)(l4.x - ry0) * pcols + 4 * cq;
                const float* wc = WV + r;
                float2 a0[4], a1[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) a0[j] = a1[j] = make_float2(0.0f, 0.0f);
                switch (c4.x) {
                    case 1: ft_vtaps4<EXACT, 1, VP>(sp, pcols, wc, one, a0, a1); break;
                    case 2: ft_vtaps4<EXACT, 2, VP>(sp, pcols, wc, one, a0, a1); break;
                    case 3: ft_vtaps4<EXACT, 3, VP>(sp, pcols, wc, one, a0, a1); break;
                    case 4: ft_vtaps4<EXACT, 4, VP>(sp, pcols, wc, one, a0, a1); break;
                    case 5: ft_vtaps4<EXACT, 5, VP>(sp, pcols, wc, one, a0, a1); break;
                    case 6: ft_vtaps4<EXACT, 6, VP>(sp, pcols, wc, one, a0, a1); break;
                    case 7: ft_vtaps4<EXACT, 7, VP>(sp, pcols, wc, one, a0, a1); break;
                    case 8: ft_vtaps4<EXACT, 8, VP>(sp, pcols, wc, one, a0, a1); break;
                    default: break;
                }
                float4* t = reinterpret_cast<float4*>(Tm + (size_t)(4 * cq) * TP + r);      // column-major: four rows of a column are one float4
                t[0] = make_float4(a0[0].x, a0[1].x, a0[2].x, a0[3].x);
                t[TP / 4] = make_float4(a0[0].y, a0[1].y, a0[2].y, a0[3].y);
                t[2 * (TP / 4)] = make_float4(a1[0].x, a1[1].x, a1[2].x, a1[3].x);
                t[3 * (TP / 4)] = make_float4(a1[0].y, a1[1].y, a1[2].y, a1[3].y);
                float m = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(fabsf(a0[j].x)));
                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(fabsf(a0[j].y)));
                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(fabsf(a1[j].x)));
                    asm("max.NaN.f32 %0, %0, %1;" : "+f"(m) : "f"(fabsf(a1[j].y)));
                }
                if (!(m <= 3.402823466e+38f)) nonfinite[it] = 1u;
            }
        } else {   // a strip that starts off the four-row grid of the table: row by row
            for (uint32_t i = tid; i < nquad * G; i += FT_THREADS) vrow(i / G, i % G);
        }
        // block-wide stores: the two tensor stores of the previous group must have READ the tile before it is overwritten
        if (!WARP_STORE && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();                                              // Tm complete, stage b drained (the ONE barrier per group with WARP_STORE)
        // Every intermediate finite => every horizontal sum is a number (an overflow gives +-inf, never inf - inf), and
        // clamp(x, 0, 1) of a number is ONE saturating add of +0 (FADD.SAT) instead of min.NaN + max.NaN: rn(x + 0) == x,
        // -0 cannot occur (the sums start from +0).  A group that holds a NaN or an infinity keeps the two-instruction clamp.
        const bool sat = unit_clamp && nonfinite[it] == 0u;
        if (tid == 0) nonfinite[it == 0 ? 2 : it - 1] = 0u;           // == (it + 2) % 3: next used two barriers from here
        // ---- horizontal pass, RC rows at a time ----
        if (warp_live) {
#pragma unroll 1
            for (int rc = 0; rc < G; rc += RC) {
                float2 acc[FS_CPT][RC / 2];
#pragma unroll
                for (int c = 0; c < FS_CPT; ++c)
#pragma unroll
                    for (int q = 0; q < RC / 2; ++q) acc[c][q] = make_float2(0.0f, 0.0f);
                if (shared_window) {
                    const float4* t = reinterpret_cast<const float4*>(Tm + (size_t)left[0] * TP + rc);
#pragma unroll
                    for (int j = 0; j < FS_MAXT; ++j) {
                        if ((uint32_t)j < cnt[0]) {
                            const float4 w4 = __ldg(hw4 + (size_t)j * dw4);
                            const float w[FS_CPT] = {w4.x, w4.y, w4.z, w4.w};
                            float4 v[RC / 4];
#pragma unroll
                            for (int q = 0; q < RC / 4; ++q) v[q] = t[j * (TP / 4) + q];
#pragma unroll
                            for (int c = 0; c < FS_CPT; ++c)
#pragma unroll
                                for (int q = 0; q < RC / 4; ++q) {
                                    acc[c][2 * q] = tap2<EXACT>(acc[c][2 * q], make_float2(v[q].x, v[q].y), w[c], one);
                                    acc[c][2 * q + 1] = tap2<EXACT>(acc[c][2 * q + 1], make_float2(v[q].z, v[q].w), w[c], one);
                                }
                        }
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < FS_MAXT; ++j) {
                        if ((uint32_t)j < cmax) {
                            const float4 w4 = __ldg(hw4 + (size_t)j * dw4);   // taps past a column's count are zero padding in the table, and skipped below
                            const float w[FS_CPT] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
                            for (int c = 0; c < FS_CPT; ++c) {
                                if ((uint32_t)j < cnt[c]) {
                                    const float4* t = reinterpret_cast<const float4*>(Tm + (size_t)(left[c] + j) * TP + rc);
#pragma unroll
                                    for (int q = 0; q < RC / 4; ++q) {
                                        const float4 v = t[q];
                                        acc[c][2 * q] = tap2<EXACT>(acc[c][2 * q], make_float2(v.x, v.y), w[c], one);
                                        acc[c][2 * q + 1] = tap2<EXACT>(acc[c][2 * q + 1], make_float2(v.z, v.w), w[c], one);
                                    }
                                }
                            }
                        }
                    }
                }
                if (WARP_STORE) {
                    // each warp stages and stores its own RC x 128 tile; the store issued two tiles ago must have READ this buffer
                    if (lane == 0 && chunk >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                    __syncwarp();
                    float* O = Ow + (chunk & 1u) * (size_t)(RC * FT_WARP_COLS) + FS_CPT * lane;
#pragma unroll
                    for (int r = 0; r < RC; ++r) {
                        float v[FS_CPT];
#pragma unroll
                        for (int c = 0; c < FS_CPT; ++c) v[c] = ft_clamp((r & 1) ? acc[c][r >> 1].y : acc[c][r >> 1].x, clo, chi, sat);
                        *reinterpret_cast<float4*>(O + (size_t)r * FT_WARP_COLS) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile must be visible to the TMA (async proxy)
                    __syncwarp();
                    if (lane == 0) {                                      // rows and columns past the result are clipped by the TMA
                        ft_tma_store_2d(&pm.dst[blockIdx.z], ox0 + (uint32_t)warp * FT_WARP_COLS, g * G + rc, Ow + (chunk & 1u) * (size_t)(RC * FT_WARP_COLS), policy);
                        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                    }
                    ++chunk;
                } else {
                    // the block's G x 512 tile as two G x 256 halves: thread t owns columns 4t .. 4t+3
                    float* O = reinterpret_cast<float*>(ftm + L.o) + (tid >= FT_HALF / FS_CPT ? (size_t)G * FT_HALF : 0) + FS_CPT * (tid & (FT_HALF / FS_CPT - 1));
#pragma unroll
                    for (int r = 0; r < RC; ++r) {
                        float v[FS_CPT];
#pragma unroll
                        for (int c = 0; c < FS_CPT; ++c) v[c] = ft_clamp((r & 1) ? acc[c][r >> 1].y : acc[c][r >> 1].x, clo, chi, sat);
                        *reinterpret_cast<float4*>(O + (size_t)(rc + r) * FT_HALF) = make_float4(v[0], v[1], v[2], v[3]);
                    }
                }
            }
        }
        if (!WARP_STORE) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tile must be visible to the TMA (async proxy)
            __syncthreads();                                              // tile complete
            if (tid == 0) {
                const float* O = reinterpret_cast<const float*>(ftm + L.o);
                ft_tma_store_2d(&pm.dst[blockIdx.z], ox0, g * G, O, policy);          // rows and columns past the result are clipped by the TMA
                if (ox0 + FT_HALF < dw) ft_tma_store_2d(&pm.dst[blockIdx.z], ox0 + FT_HALF, g * G, O + (size_t)G * FT_HALF, policy);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        ry0 = nry0;
        b ^= 1;
        it = it == 2 ? 0 : it + 1;
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // shared memory outlives the reads of the last stores
}

// horizontal_sample for LONG windows.  One CTA = 256 adjacent outputs x HT_ROWS rows: the stretch of
// the intermediate those outputs read is staged in shared memory with coalesced loads (index i
// lives at i + i/32, so the stride-R reads of neighbouring outputs spread over the banks), and a
// thread walks its window ONCE for all the rows -- each tap weight is fetched once per HT_ROWS
// outputs and the rows are independent accumulation chains.  Same tap order and clamp as
// kc_resize_h_kernel.

// Four rows of a 256-output tile per CTA.  The stretch of the intermediate those outputs read is staged in shared
// memory TRANSPOSED -- one 16-byte slot per source column holding its four rows -- so a tap costs one LDS.128 and
// four FMAs (the row-major layout needed four LDS and the index arithmetic four times over: 14.6 M warp
// instructions for 8192x1024 -> 1024x1024, issue slots 59 % busy).  A spare slot after every eight columns
// spreads the stride-R reads of adjacent outputs over the banks.
// L2 residency control for the long-window path.  The 256 MiB source is a pure stream (evict-first) EXCEPT the rows two
// neighbouring blocks both need (the window overlap at a chunk boundary, ~15 % of the source): the lower block reads them
// first, at its start, and marks them evict-last so that they are still in L2 when the upper block gets to them at its end
// -- without that every overlap row came out of DRAM twice (dram__bytes_read 1.15 x the source, profiles/ncu_full_resize_down_v_r01).
// The 32 MiB intermediate is written evict-last as well and read back evict-first by the horizontal pass: it never leaves L2.
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}
__device__ __forceinline__ float4 ld_nc_hint4(const float4* a, uint64_t policy) {
    float4 v;
    asm volatile("ld.global.nc.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(a), "l"(policy));
    return v;
}
__device__ __forceinline__ float ld_nc_hint(const float* a, uint64_t policy) {
    float v;
    asm volatile("ld.global.nc.L2::cache_hint.f32 %0, [%1], %2;" : "=f"(v) : "l"(a), "l"(policy));
    return v;
}
__device__ __forceinline__ void st_hint4(float4* a, const float4& v, uint64_t policy) {
    asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;" ::"l"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w), "l"(policy) : "memory");
}

constexpr int HT_ROWS = 4;

__device__ __forceinline__ void vt_bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ft_smem(dst)), "l"(src), "r"(bytes),
                 "r"(ft_smem(bar))
                 : "memory");
}

// T threads, 2 T output columns per block: a thread owns two ADJACENT outputs.  Their windows overlap (by five sixths at a
// ratio of 8 with Lanczos3), and a tap value fetched from shared memory once serves both; the block's tap weights arrive in
// shared memory by one bulk copy (the host lays them out block by block) and the tile is staged with float4 loads.  Each
// output still sums its own taps left to right.  Measured, 8192 -> 1024 columns x 1024 rows (round 2): 0.019-0.021 ms for
// every block shape tried, against 0.023 ms for one output per thread with weights from L2 -- the remainder is the read of
// the intermediate, which ncu shows coming from DRAM (42 MB read, L2 hit rate 19 %) although the march stores it
// evict-last; neither the load hint nor the shape of the blocks changes that.
__device__ __forceinline__ uint32_t ht_slot2(uint32_t i) { return i + (i >> 4); }   // lanes 2 x ratio columns apart: 17 float4 at a ratio of 8

template <bool EXACT, int T>
__global__ void __launch_bounds__(T * 4) kc_resize_h_tile_kernel(const float* __restrict__ tmp, uint32_t sw, float* __restrict__ dst, uint32_t dw,
                                                             uint32_t dh, const uint32_t* __restrict__ left, const uint32_t* __restrict__ count,
                                                             const float* __restrict__ weo, uint32_t max_taps, uint32_t tile_f4, float clo, float chi,
                                                             float one) {
    // shared memory: the block's tap weights [max_taps][T even outputs | T odd outputs] (one bulk copy: the host lays the
    // table out block by block), then the tile [ht_slot2(ncol)] x (4 rows)
    extern __shared__ __align__(128) unsigned char hts[];
    uint64_t* wbar = reinterpret_cast<uint64_t*>(hts);
    float* ws = reinterpret_cast<float*>(hts + 16);
    const uint32_t wbytes = max_taps * (2 * T) * (uint32_t)sizeof(float);
    // blockDim.y groups of four rows share the weights (what limits the resident warps is shared memory); a tile each
    float4* htile4 = reinterpret_cast<float4*>(hts + 16 + wbytes) + (size_t)threadIdx.y * tile_f4;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
        ft_mbar_init(wbar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        ft_mbar_expect_tx(wbar, wbytes);
        vt_bulk_load_1d(ws, weo + (size_t)blockIdx.x * max_taps * (2 * T), wbytes, wbar);
    }
    const uint32_t ox0 = blockIdx.x * (2 * T), oxl = min(ox0 + 2 * T, dw) - 1;
    const uint32_t grp = blockIdx.y * blockDim.y + threadIdx.y;
    const bool live = grp * HT_ROWS < dh;                           // a group past the last row stages row 0 and stores nothing
    const uint32_t y0 = live ? grp * HT_ROWS : 0u, nrow = live ? min((uint32_t)HT_ROWS, dh - y0) : 1u;
    // the block's window starts at a multiple of four columns where rows are 16-byte aligned: the staging then moves float4s
    const bool vec = (sw & 3u) == 0 && (reinterpret_cast<uintptr_t>(tmp) & 15u) == 0;
    const uint32_t c0 = vec ? __ldg(left + ox0) & ~3u : __ldg(left + ox0), ncol = __ldg(left + oxl) + __ldg(count + oxl) - c0;
    const float* r0 = tmp + (size_t)y0 * sw + c0;
    const float* r1 = tmp + (size_t)(y0 + min(1u, nrow - 1)) * sw + c0;   // rows past nrow: duplicates, never stored
    const float* r2 = tmp + (size_t)(y0 + min(2u, nrow - 1)) * sw + c0;
    const float* r3 = tmp + (size_t)(y0 + min(3u, nrow - 1)) * sw + c0;
    // staging.  The march stored the intermediate evict-last (ncu: most of it comes back from DRAM all the same); this is its
    // last use, so the reads are evict-first.  Four columns x four rows per load group, four groups in flight per thread
    const uint64_t pol = l2_policy_evict_first();
    if (vec) {
        const uint32_t ncol4 = (ncol + 3u) >> 2;                   // sw % 4 == 0: the last float4 ends inside the row
        const float4 *q0 = reinterpret_cast<const float4*>(r0), *q1 = reinterpret_cast<const float4*>(r1), *q2 = reinterpret_cast<const float4*>(r2),
                     *q3 = reinterpret_cast<const float4*>(r3);
        constexpr int NG = 4;                                        // load groups in flight per thread (16 float4 loads)
        for (uint32_t i0 = threadIdx.x; i0 < ncol4; i0 += NG * T) {
            float4 v[NG][4];
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint32_t i = min(i0 + (uint32_t)(g * T), ncol4 - 1);
                v[g][0] = ld_nc_hint4(q0 + i, pol);
                v[g][1] = ld_nc_hint4(q1 + i, pol);
                v[g][2] = ld_nc_hint4(q2 + i, pol);
                v[g][3] = ld_nc_hint4(q3 + i, pol);
            }
#pragma unroll
            for (int g = 0; g < NG; ++g) {
                const uint32_t i = i0 + (uint32_t)(g * T);
                if (i < ncol4) {
                    htile4[ht_slot2(4 * i + 0)] = make_float4(v[g][0].x, v[g][1].x, v[g][2].x, v[g][3].x);
                    htile4[ht_slot2(4 * i + 1)] = make_float4(v[g][0].y, v[g][1].y, v[g][2].y, v[g][3].y);
                    htile4[ht_slot2(4 * i + 2)] = make_float4(v[g][0].z, v[g][1].z, v[g][2].z, v[g][3].z);
                    htile4[ht_slot2(4 * i + 3)] = make_float4(v[g][0].w, v[g][1].w, v[g][2].w, v[g][3].w);
                }
            }
        }
    } else {
        for (uint32_t i0 = threadIdx.x; i0 < ncol; i0 += 2 * T) {
            const uint32_t ia = i0, ib = min(i0 + (uint32_t)T, ncol - 1);
            const float4 va = make_float4(ld_nc_hint(r0 + ia, pol), ld_nc_hint(r1 + ia, pol), ld_nc_hint(r2 + ia, pol), ld_nc_hint(r3 + ia, pol));
            const float4 vb = make_float4(ld_nc_hint(r0 + ib, pol), ld_nc_hint(r1 + ib, pol), ld_nc_hint(r2 + ib, pol), ld_nc_hint(r3 + ib, pol));
            htile4[ht_slot2(ia)] = va;
            if (i0 + (uint32_t)T < ncol) htile4[ht_slot2(i0 + (uint32_t)T)] = vb;
        }
    }
    __syncthreads();
    ft_mbar_wait(wbar, 0);
    const uint32_t xa = ox0 + 2 * threadIdx.x, xb = xa + 1;
    if (xa > oxl || !live) return;
    const bool hb = xb <= oxl;
    // windows [ca, ea) and [cb, eb) in tile columns; the four rows of an output as two FFMA2 operand pairs, the weight table
    // holding every weight as a (w, w) pair
    uint32_t ca = __ldg(left + xa) - c0, cb = hb ? __ldg(left + xb) - c0 : 0u;
    const uint32_t ea = ca + __ldg(count + xa), eb = hb ? cb + __ldg(count + xb) : 0u;
    const float* wa = ws + threadIdx.x;                            // next weight of a: wa[0], then wa += 2 T
    const float* wb = ws + T + threadIdx.x;
    constexpr int WS = 2 * T;
    float2 a01 = make_float2(0.0f, 0.0f), a23 = a01, b01 = a01, b23 = a01;
    // taps of a alone, left of b's window
    for (const uint32_t e = min(ea, hb ? cb : ea); ca < e; ++ca, wa += WS) {
        const float w = wa[0];
        const float4 sv = htile4[ht_slot2(ca)];
        a01 = tap2<EXACT>(a01, make_float2(sv.x, sv.y), make_float2(w, w), one);
        a23 = tap2<EXACT>(a23, make_float2(sv.z, sv.w), make_float2(w, w), one);
    }
    // the columns both windows hold: one fetch, four FFMA2
    if (hb && ca == cb) {
        const uint32_t e = min(ea, eb);
        constexpr int WB = 4;
        for (; ca + WB <= e; ca += WB, wa += WB * WS, wb += WB * WS) {
#pragma unroll
            for (int k = 0; k < WB; ++k) {
                const float u = wa[k * WS], v = wb[k * WS];
                const float4 sv = htile4[ht_slot2(ca + k)];
                const float2 lo = make_float2(sv.x, sv.y), hi = make_float2(sv.z, sv.w);
                a01 = tap2<EXACT>(a01, lo, make_float2(u, u), one);
                a23 = tap2<EXACT>(a23, hi, make_float2(u, u), one);
                b01 = tap2<EXACT>(b01, lo, make_float2(v, v), one);
                b23 = tap2<EXACT>(b23, hi, make_float2(v, v), one);
            }
        }
        for (; ca < e; ++ca, wa += WS, wb += WS) {
            const float u = wa[0], v = wb[0];
            const float4 sv = htile4[ht_slot2(ca)];
            const float2 lo = make_float2(sv.x, sv.y), hi = make_float2(sv.z, sv.w);
            a01 = tap2<EXACT>(a01, lo, make_float2(u, u), one);
            a23 = tap2<EXACT>(a23, hi, make_float2(u, u), one);
            b01 = tap2<EXACT>(b01, lo, make_float2(v, v), one);
            b23 = tap2<EXACT>(b23, hi, make_float2(v, v), one);
        }
        cb = ca;
    }
    // what is left of either window
    for (; ca < ea; ++ca, wa += WS) {
        const float w = wa[0];
        const float4 sv = htile4[ht_slot2(ca)];
        a01 = tap2<EXACT>(a01, make_float2(sv.x, sv.y), make_float2(w, w), one);
        a23 = tap2<EXACT>(a23, make_float2(sv.z, sv.w), make_float2(w, w), one);
    }
    for (; cb < eb; ++cb, wb += WS) {
        const float w = wb[0];
        const float4 sv = htile4[ht_slot2(cb)];
        b01 = tap2<EXACT>(b01, make_float2(sv.x, sv.y), make_float2(w, w), one);
        b23 = tap2<EXACT>(b23, make_float2(sv.z, sv.w), make_float2(w, w), one);
    }
    const float a[HT_ROWS] = {a01.x, a01.y, a23.x, a23.y}, b[HT_ROWS] = {b01.x, b01.y, b23.x, b23.y};
#pragma unroll
    for (int r = 0; r < HT_ROWS; ++r)
        if ((uint32_t)r < nrow) {
            float* o = dst + (size_t)(y0 + r) * dw + xa;
            const float va = a[r] < clo ? clo : (a[r] > chi ? chi : a[r]);   // image::math::utils::clamp keeps NaN
            const float vb = b[r] < clo ? clo : (b[r] > chi ? chi : b[r]);
            o[0] = va;
            if (hb) o[1] = vb;
        }
}

// ---------------------------------------------------------------------------
// Vertical pass for LONG windows (downsampling: 6 R + 1 taps for a ratio R with Lanczos3), as
// a march over the SOURCE rows.  The two-pass kernel above reads every source row once per
// output row whose window holds it (~6x, all of it L2 traffic); here a thread owns four
// columns, streams the source rows of its CTA's output range exactly once and keeps the up to
// eight output rows in flight in registers: output o lives in ring slot o mod 8 (windows of o
// and o + 8 never overlap), so for a source row the host-built table gives, per slot, the tap
// weight (NaN: the row is in no window of that slot) and the output that is complete after it.
// Each output still sums its taps top to bottom starting from +0: the reference's order.
// ---------------------------------------------------------------------------
constexpr int VM_THREADS = 128;
constexpr int VM_SLOTS = 8;
constexpr int VM_ROWS_PER_CTA = 32;   // output rows per CTA (its source range overlaps the neighbours' by one window)

template <bool EXACT>
__device__ __forceinline__ void vm_tap4(float4& a, const float4& v, float w) {
    if (EXACT) {
        a.x = __fadd_rn(a.x, __fmul_rn(v.x, w)); a.y = __fadd_rn(a.y, __fmul_rn(v.y, w));
        a.z = __fadd_rn(a.z, __fmul_rn(v.z, w)); a.w = __fadd_rn(a.w, __fmul_rn(v.w, w));
    } else {                                               // two packed FFMA2 instead of four FFMA
        const float2 ww = make_float2(w, w);
        const float2 lo = __ffma2_rn(make_float2(v.x, v.y), ww, make_float2(a.x, a.y));
        const float2 hi = __ffma2_rn(make_float2(v.z, v.w), ww, make_float2(a.z, a.w));
        a = make_float4(lo.x, lo.y, hi.x, hi.y);
    }
}

template <bool EXACT>
__global__ void __launch_bounds__(VM_THREADS) kc_resize_v_march_kernel(const float4* __restrict__ src, uint32_t sw4, float4* __restrict__ tmp,
                                                                       uint32_t dh, const uint32_t* __restrict__ vleft,
                                                                       const uint32_t* __restrict__ vcount, const float4* __restrict__ mw,
                                                                       const int4* __restrict__ mo, uint32_t rows_per_cta) {
    extern __shared__ __align__(16) float4 vm_sm[];        // the tables of this CTA's source rows: [rows][2] weights, [rows][2] retire ids
    const uint32_t x4 = blockIdx.x * VM_THREADS + threadIdx.x;
    const uint32_t oyA = blockIdx.y * rows_per_cta, oyB = min(oyA + rows_per_cta, dh);
    const uint32_t r0 = __ldg(vleft + oyA), r1 = __ldg(vleft + oyB - 1) + __ldg(vcount + oyB - 1);
    const uint32_t nr = r1 - r0;
    // rows [r0, r_shared) are also the LAST rows of the block above: keep them in L2 for it
    const uint32_t r_shared = oyA == 0 ? r0 : min(r1, __ldg(vleft + oyA - 1) + __ldg(vcount + oyA - 1));
    const uint32_t n_shared = r_shared > r0 ? r_shared - r0 : 0u;
    const uint64_t pol_stream = l2_policy_evict_first(), pol_keep = l2_policy_evict_last();
    float4* sw_ = vm_sm;
    int4* so_ = reinterpret_cast<int4*>(vm_sm + 2 * (size_t)nr);
    for (uint32_t i = threadIdx.x; i < 2 * nr; i += VM_THREADS) {
        sw_[i] = __ldg(mw + 2 * (size_t)r0 + i);
        so_[i] = __ldg(mo + 2 * (size_t)r0 + i);
    }
    __syncthreads();
    if (x4 >= sw4) return;
    float4 acc[VM_SLOTS];
#pragma unroll
    for (int s = 0; s < VM_SLOTS; ++s) acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4* col = src + x4;
    constexpr int U = 8;                                   // source rows whose loads are in flight together
    for (uint32_t rb = 0; rb < nr; rb += U) {
        float4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t rr = min(rb + u, nr - 1);
            v[u] = ld_nc_hint4(col + (size_t)(r0 + rr) * sw4, rr < n_shared ? pol_keep : pol_stream);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t r = rb + u;                     // relative to r0
            if (r >= nr) break;                            // uniform
            const float4 wa = sw_[2 * r], wb = sw_[2 * r + 1];
            const float w[VM_SLOTS] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
            for (int s = 0; s < VM_SLOTS; ++s)
                if (w[s] == w[s]) vm_tap4<EXACT>(acc[s], v[u], w[s]);
            const int4 oa = so_[2 * r], ob = so_[2 * r + 1];
            const int o[VM_SLOTS] = {oa.x, oa.y, oa.z, oa.w, ob.x, ob.y, ob.z, ob.w};
            // "no output completes here" is every entry -1: the AND of all eight keeps its sign bit
            if ((oa.x & oa.y & oa.z & oa.w & ob.x & ob.y & ob.z & ob.w) < 0) continue;
#pragma unroll
            for (int s = 0; s < VM_SLOTS; ++s)
                if (o[s] >= 0) {
                    if ((uint32_t)o[s] >= oyA && (uint32_t)o[s] < oyB) st_hint4(tmp + (size_t)o[s] * sw4 + x4, acc[s], pol_keep);
                    acc[s] = make_float4(0.f, 0.f, 0.f, 0.f);
                }
        }
    }
}


// ---------------------------------------------------------------------------
// The same march fed by the TMA.  The kernel above lives on the loads each warp has in flight between two stretches of
// arithmetic (issue eight rows, wait, compute them, repeat): at 443 threads per SM that is latency-bound at 3.7 TB/s.  Here a
// PRODUCER warp keeps VT_STAGES stages of VT_R source rows x 512 columns in flight as 2-D tensor-map loads (two 256-column
// boxes per stage, completing on an mbarrier, with the L2 hints of the kernel above: rows the block above also needs are
// evict-last, the rest evict-first) plus the stage's slice of the marching tables as two 1-D bulk copies; four COMPUTE warps
// (a thread owns four columns) consume the stages out of shared memory and hand them back through an "empty" mbarrier.  Bytes
// in flight are set by the stage ring, not by what the compute warps happen to be doing.  Arithmetic and order are those of
// kc_resize_v_march_kernel.
// ---------------------------------------------------------------------------
constexpr int VT_R = 8;                       // source rows per stage
constexpr int VT_COLS = 512;                  // source columns per block (128 compute threads x 4)
constexpr int VT_PIX_BYTES = VT_R * VT_COLS * 4;
constexpr int VT_INFO = 12;                   // ints per source row in the retire table: eight output ids, the flag word, padding
constexpr int VT_STAGE_BYTES = VT_PIX_BYTES + VT_R * 64 + VT_R * VT_INFO * 4;  // pixels + weights (each one twice: an FFMA2 operand pair) + retire info

__device__ __forceinline__ void vt_tma_load_2d_hint(void* dst, const CUtensorMap* map, uint32_t c0, uint32_t c1, uint64_t* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(
                     ft_smem(dst)),
                 "l"(map), "r"(c0), "r"(c1), "r"(ft_smem(bar)), "l"(policy)
                 : "memory");
}
__device__ __forceinline__ void vt_mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(ft_smem(bar)) : "memory");
}

// one source row into one open output: four columns as two FFMA2 operand pairs (EXACT: the product rounds on its own)
template <bool EXACT>
__device__ __forceinline__ void vt_tap(float2& alo, float2& ahi, const float2& vlo, const float2& vhi, const float2& ww, float one) {
    if (EXACT) {
        alo = __ffma2_rn(alo, make_float2(one, one), __fmul2_rn(vlo, ww));
        ahi = __ffma2_rn(ahi, make_float2(one, one), __fmul2_rn(vhi, ww));
    } else {
        alo = __ffma2_rn(vlo, ww, alo);
        ahi = __ffma2_rn(vhi, ww, ahi);
    }
}

// CTAS: resident blocks per SM the kernel is compiled for (the register budget), VT_STAGES: depth of its ring.  The pass is bound
// by the latency of a warp's row (its time is proportional to the source rows per block whether one or three blocks share an
// SM), so what pays is MORE blocks with fewer rows each: four per SM in FAST (94 registers); EXACT needs 107 and stays at
// three.  The ring's depth does not matter (two stages are as fast as four).
template <bool EXACT, int CTAS, int VT_STAGES>
__global__ void __launch_bounds__(160, CTAS) kc_resize_v_tma_kernel(const __grid_constant__ CUtensorMap tm_src, uint32_t sw4, uint32_t sh, float4* __restrict__ tmp,
                                                              uint32_t dh, const uint32_t* __restrict__ vleft, const uint32_t* __restrict__ vcount,
                                                              const float2* __restrict__ mw2, const int* __restrict__ mi, uint32_t rows_per_cta, float one) {
    extern __shared__ __align__(128) unsigned char vts[];
    uint64_t* full = reinterpret_cast<uint64_t*>(vts);               // [VT_STAGES]
    uint64_t* empty = full + VT_STAGES;                              // [VT_STAGES]
    unsigned char* stage0 = vts + 128;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t oyA = blockIdx.y * rows_per_cta, oyB = min(oyA + rows_per_cta, dh);
    const uint32_t r0 = __ldg(vleft + oyA), r1 = __ldg(vleft + oyB - 1) + __ldg(vcount + oyB - 1);
    const uint32_t nr = r1 - r0;
    const uint32_t r_shared = oyA == 0 ? r0 : min(r1, __ldg(vleft + oyA - 1) + __ldg(vcount + oyA - 1));
    const uint32_t n_shared = r_shared > r0 ? r_shared - r0 : 0u;    // rows the block above needs as well
    const uint32_t nstage = (nr + VT_R - 1) / VT_R;
    const uint32_t cx0 = blockIdx.x * VT_COLS;
    if (tid == 0) {
        for (int i = 0; i < VT_STAGES; ++i) {
            ft_mbar_init(&full[i], 1);
            ft_mbar_init(&empty[i], 4);                              // one arrival per compute warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // one stage's loads: two tensor boxes of pixels, the stage's slice of the two marching tables
    auto issue = [&](uint32_t k) {
        const int st = k % VT_STAGES;
        unsigned char* S = stage0 + (size_t)st * VT_STAGE_BYTES;
        const uint32_t row = r0 + k * VT_R;                          // rows past the image arrive as zeros and are never used
        const uint32_t trows = min((uint32_t)VT_R, sh - row);        // table rows that exist
        ft_mbar_expect_tx(&full[st], VT_PIX_BYTES + trows * 64 + trows * VT_INFO * 4);
        const uint64_t pol = k * VT_R < n_shared ? l2_policy_evict_last() : l2_policy_evict_first();
        vt_tma_load_2d_hint(S, &tm_src, cx0, row, &full[st], pol);
        vt_tma_load_2d_hint(S + VT_R * 256 * 4, &tm_src, cx0 + 256, row, &full[st], pol);
        vt_bulk_load_1d(S + VT_PIX_BYTES, mw2 + (size_t)row * VM_SLOTS, trows * 64, &full[st]);
        vt_bulk_load_1d(S + VT_PIX_BYTES + VT_R * 64, mi + (size_t)row * VT_INFO, trows * VT_INFO * 4, &full[st]);
    };
    if (warp == 4) {
        // ---- producer warp (folding its work into thread 0 of a compute warp, for 128-thread blocks and five per SM, was
        // slower: 0.070 against 0.062 ms -- that warp waits for the others before every refill) ----
        if (lane == 0)
            for (uint32_t k = 0; k < nstage; ++k) {
                if (k >= (uint32_t)VT_STAGES) ft_mbar_wait(&empty[k % VT_STAGES], ((k / VT_STAGES) - 1) & 1u);
                issue(k);
            }
        return;
    }
    // ---- compute: thread t owns columns cx0 + 4t .. 4t+3 ----
    const uint32_t x4 = blockIdx.x * (VT_COLS / 4) + tid;
    const bool live = x4 < sw4;
    const uint64_t pol_keep = l2_policy_evict_last();
    // the four columns' running sums of the eight open outputs, as the two operand pairs of an FFMA2
    float2 alo[VM_SLOTS], ahi[VM_SLOTS];
#pragma unroll
    for (int s = 0; s < VM_SLOTS; ++s) alo[s] = ahi[s] = make_float2(0.f, 0.f);
    const uint32_t half = tid >> 6, c4 = tid & 63;                   // which 256-column box, which float4 in its rows
    for (uint32_t k = 0; k < nstage; ++k) {
        const int st = k % VT_STAGES;
        ft_mbar_wait(&full[st], (k / VT_STAGES) & 1u);
        const unsigned char* S = stage0 + (size_t)st * VT_STAGE_BYTES;
        const float4* px = reinterpret_cast<const float4*>(S + (size_t)half * VT_R * 256 * 4) + c4;
        const float4* wt = reinterpret_cast<const float4*>(S + VT_PIX_BYTES);               // per row: (w0,w0,w1,w1) (w2,w2,w3,w3) ...
        const int* ot = reinterpret_cast<const int*>(S + VT_PIX_BYTES + VT_R * 64);         // per row: eight output ids, flags
        const uint32_t ucount = min((uint32_t)VT_R, nr - k * VT_R);
        // the operands of a row are fetched one row ahead of their use, so the shared-memory latency sits behind the
        // sixteen FFMA2 of the row before (rows past ucount are fetched from the stage too, and not used)
        // A slot the row is no tap of has weight 0 and a running sum of +0 (it was cleared when its last output left):
        // +0 + 0*v stays +0 for every finite v, so the products need no test per slot.  A non-finite v would turn that +0
        // into NaN; those four columns take the tested path for the row.  The test of a row (0*v summed over the four
        // columns is 0) is made one row ahead as well: the branch never waits for its predicate.
        auto all_finite = [](const float4& p) {
            const float2 c = __ffma2_rn(make_float2(p.x, p.y), make_float2(0.f, 0.f), __fmul2_rn(make_float2(p.z, p.w), make_float2(0.f, 0.f)));
            return c.x + c.y == 0.f;
        };
        float4 v_n = px[0], w_n[VM_SLOTS / 2];
        uint32_t flags_n = (uint32_t)ot[8];
#pragma unroll
        for (int q = 0; q < VM_SLOTS / 2; ++q) w_n[q] = wt[q];
        bool fin_n = all_finite(v_n);
#pragma unroll
        for (int u = 0; u < VT_R; ++u) {
            if ((uint32_t)u >= ucount) break;                        // uniform
            const float4 v = v_n;
            const uint32_t flags = flags_n;                          // bits 0-7: slots completing on this row; 8-15: slots it is a tap of
            const bool fin = fin_n;
            float4 w[VM_SLOTS / 2];                                  // the weights of slots 2q and 2q+1, each already a pair
#pragma unroll
            for (int q = 0; q < VM_SLOTS / 2; ++q) w[q] = w_n[q];
            if (u + 1 < VT_R) {
                v_n = px[(u + 1) * 64];
                flags_n = (uint32_t)ot[VT_INFO * (u + 1) + 8];
#pragma unroll
                for (int q = 0; q < VM_SLOTS / 2; ++q) w_n[q] = wt[4 * (u + 1) + q];
            }
            const float2 vlo = make_float2(v.x, v.y), vhi = make_float2(v.z, v.w);
            if (fin) {
#pragma unroll
                for (int q = 0; q < VM_SLOTS / 2; ++q) {
                    vt_tap<EXACT>(alo[2 * q], ahi[2 * q], vlo, vhi, make_float2(w[q].x, w[q].y), one);
                    vt_tap<EXACT>(alo[2 * q + 1], ahi[2 * q + 1], vlo, vhi, make_float2(w[q].z, w[q].w), one);
                }
            } else {
#pragma unroll
                for (int q = 0; q < VM_SLOTS / 2; ++q) {
                    if (flags & (0x100u << (2 * q))) vt_tap<EXACT>(alo[2 * q], ahi[2 * q], vlo, vhi, make_float2(w[q].x, w[q].y), one);
                    if (flags & (0x200u << (2 * q))) vt_tap<EXACT>(alo[2 * q + 1], ahi[2 * q + 1], vlo, vhi, make_float2(w[q].z, w[q].w), one);
                }
            }
            if (u + 1 < VT_R) fin_n = all_finite(v_n);
            if ((flags & 0xffu) == 0u) continue;                     // nothing completes on this row
#pragma unroll
            for (int s = 0; s < VM_SLOTS; ++s)
                if (flags & (1u << s)) {
                    const uint32_t o = (uint32_t)ot[VT_INFO * u + s];
                    if (live && o >= oyA && o < oyB)
                        st_hint4(tmp + (size_t)o * sw4 + x4, make_float4(alo[s].x, alo[s].y, ahi[s].x, ahi[s].y), pol_keep);
                    alo[s] = ahi[s] = make_float2(0.f, 0.f);
                }
        }
        __syncwarp();
        if (lane == 0) vt_mbar_arrive(&empty[st]);                   // this warp is done with the stage
    }
}

// host: the weights of an axis as the long-window horizontal pass wants them: block by block of `t2` outputs, tap-major inside
// a block, the even outputs of the block before the odd ones -- [block][tap][t2/2 even | t2/2 odd]; one bulk copy per block
int32_t build_block_weights(kc_context* ctx, KcAxisTable& t, uint32_t t2, const float** out) {
    const int slot = t2 == 256 ? 0 : t2 == 128 ? 1 : 2;
    if (t.d_weights_eo[slot]) { *out = t.d_weights_eo[slot]; return KC_OK; }
    if (ctx->capturing) KC_FAIL(KC_ERR_GENERIC, "a weight table would have to be uploaded during a stream capture");
    const uint32_t nblk = (t.dst_len + t2 - 1) / t2, half = t2 / 2;
    std::vector<float> w((size_t)nblk * t.max_taps * t2, 0.0f);
    for (uint32_t o = 0; o < t.dst_len; ++o) {
        const uint32_t b = o / t2, i = o % t2, pos = (i & 1u) * half + (i >> 1);
        for (uint32_t k = 0; k < t.max_taps; ++k) w[((size_t)b * t.max_taps + k) * t2 + pos] = t.h_weights[(size_t)o * t.max_taps + k];
    }
    KC_CUDA(cudaMalloc((void**)&t.d_weights_eo[slot], w.size() * sizeof(float)));
    KC_CUDA(cudaMemcpyAsync(t.d_weights_eo[slot], w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    KC_CUDA(cudaStreamSynchronize(ctx->stream));
    *out = t.d_weights_eo[slot];
    return KC_OK;
}

// host: the marching tables of an axis; false when some source index sits in more than one
// window of a ring slot, or a real weight is NaN (the sentinel) -- the caller then falls back
int32_t build_march_tables(kc_context* ctx, KcAxisTable& t) {
    if (t.march_state != 0) return KC_OK;
    if (ctx->capturing) KC_FAIL(KC_ERR_GENERIC, "a marching table would have to be uploaded during a stream capture");
    t.march_state = -1;
    const uint32_t S = t.src_len, D = t.dst_len;
    std::vector<float> w((size_t)S * VM_SLOTS, nanf(""));
    std::vector<int32_t> o((size_t)S * VM_SLOTS, -1);
    for (uint32_t oy = 0; oy < D; ++oy) {
        const int s = (int)(oy % VM_SLOTS);
        const uint32_t l = t.h_left[oy], n = t.h_count[oy];
        for (uint32_t k = 0; k < n; ++k) {
            const float wk = t.h_weights[(size_t)oy * t.max_taps + k];
            float& cell = w[(size_t)(l + k) * VM_SLOTS + s];
            if (wk != wk || cell == cell) return KC_OK;        // NaN weight, or the slot is taken: cannot march
            cell = wk;
        }
        if (o[(size_t)(l + n - 1) * VM_SLOTS + s] >= 0) return KC_OK;
        o[(size_t)(l + n - 1) * VM_SLOTS + s] = (int32_t)oy;
    }
    {   // the TMA-fed march's tables.  Weights: every entry doubled (an aligned (w, w) pair is what an FFMA2 takes as its
        // broadcast operand) and 0 where the row is no tap of the slot.  Info: the eight output ids and a flag word per row
        // (bits 0-7 the slots completing there, bits 8-15 the slots the row is a tap of).
        std::vector<float> w2(w.size() * 2);
        std::vector<int32_t> info((size_t)S * VT_INFO, 0);
        for (size_t r = 0; r < S; ++r) {
            uint32_t flags = 0;
            for (int sl = 0; sl < VM_SLOTS; ++sl) {
                const size_t i = r * VM_SLOTS + sl;
                const bool tap = w[i] == w[i];
                w2[2 * i] = w2[2 * i + 1] = tap ? w[i] : 0.0f;
                if (tap) flags |= 0x100u << sl;
                if (o[i] >= 0) flags |= 1u << sl;
                info[r * VT_INFO + sl] = o[i];
            }
            info[r * VT_INFO + 8] = (int32_t)flags;
        }
        KC_CUDA(cudaMalloc((void**)&t.d_march_w2, w2.size() * sizeof(float)));
        KC_CUDA(cudaMalloc((void**)&t.d_march_info, info.size() * sizeof(int32_t)));
        KC_CUDA(cudaMemcpyAsync(t.d_march_w2, w2.data(), w2.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
        KC_CUDA(cudaMemcpyAsync(t.d_march_info, info.data(), info.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
        KC_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    KC_CUDA(cudaMalloc((void**)&t.d_march_w, w.size() * sizeof(float)));
    KC_CUDA(cudaMalloc((void**)&t.d_march_o, o.size() * sizeof(int32_t)));
    KC_CUDA(cudaMemcpyAsync(t.d_march_w, w.data(), w.size() * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
    KC_CUDA(cudaMemcpyAsync(t.d_march_o, o.data(), o.size() * sizeof(int32_t), cudaMemcpyHostToDevice, ctx->stream));
    KC_CUDA(cudaStreamSynchronize(ctx->stream));           // the vectors go out of scope
    t.march_state = 1;
    return KC_OK;
}

// largest source window any tile of FT_T output elements touches along one axis
uint32_t max_window(const KcAxisTable& t, uint32_t tile) {
    uint32_t mx = 0;
    for (uint32_t o0 = 0; o0 < t.dst_len; o0 += tile) {
        const uint32_t ol = std::min(o0 + tile, t.dst_len) - 1;
        mx = std::max(mx, t.h_left[ol] + t.h_count[ol] - t.h_left[o0]);
    }
    return mx;
}

// the same for tiles that may start at any output index (row windows of a strip)
uint32_t max_window_sliding(const KcAxisTable& t, uint32_t tile) {
    uint32_t mx = 0;
    for (uint32_t o0 = 0; o0 < t.dst_len; ++o0) {
        const uint32_t ol = std::min(o0 + tile, t.dst_len) - 1;
        mx = std::max(mx, t.h_left[ol] + t.h_count[ol] - t.h_left[o0]);
    }
    return mx;
}

}  // namespace

int32_t kck_resize_plane(kc_context* ctx, const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t dw,
                         uint32_t dh, int filter) {
    return kck_resize_plane_rows(ctx, src, sw, sh, dst, dw, dh, filter, 0, dh);
}

// rows [row0, row0 + nrows) of the dw x dh resize of src, written to dst (dw x nrows).  A GPU that
// owns a horizontal strip of the result calls this with its rows: it needs the source rows
// [left(row0), right(row0 + nrows - 1)) only -- for an upsample, simply the whole (small) source.
// Up to four planes of the same geometry in ONE launch of the tensor-map kernel (grid.z = plane).  *done = false: the
// configuration is not one that kernel takes (the caller resizes plane by plane).
int32_t kck_resize_planes_rows_batched(kc_context* ctx, const float* const* srcs, float* const* dsts, int n, uint32_t sw, uint32_t sh,
                                       uint32_t dw, uint32_t dh, int filter, uint32_t row0, uint32_t nrows, bool* done) {
    *done = false;
    if (n < 1 || n > FT_MAX_PLANES || dw == 0 || dh == 0 || nrows == 0 || sw == 0 || sh == 0 || row0 > dh || nrows > dh - row0) return KC_OK;
    std::shared_ptr<KcAxisTable> tv, th;
    KC_TRY(get_axis(ctx, sh, dh, filter, tv));
    KC_TRY(get_axis(ctx, sw, dw, filter, th));
    const bool exact_mode = ctx->opts.math_mode == KC_MATH_EXACT;
    const float clo = ctx->opts.resize_unclamped ? -INFINITY : 0.0f, chi = ctx->opts.resize_unclamped ? INFINITY : 1.0f;
    static const bool no_fused = getenv("KC_RESIZE_TWO_PASS") != nullptr;
    const bool no_tma = g_kc_tuning.resize_tma < 0;
    uintptr_t align = 0;
    for (int i = 0; i < n; ++i) align |= (uintptr_t)srcs[i] | (uintptr_t)dsts[i];
    if (no_fused || no_tma || tv->max_taps > (uint32_t)FS_MAXT || th->max_taps > (uint32_t)FS_MAXT || (sw & 3u) != 0 || (dw & 3u) != 0 ||
        dw < (uint32_t)FT_TW || nrows < 32 || (align & 15u) != 0 || !tensor_map_encoder())
        return KC_OK;
    // rows per group / rows per accumulator chunk / CTAs per SM: tuning knobs (kc_debug_set_tuning, scripts/resize_sweep.py);
    // defaults from profiles/resize_sweep_r02_{fast,exact}.json: every warp stores its own tiles; FAST 16-row groups in
    // 4-row chunks, EXACT (twice the floating-point instructions per tap) 32-row groups in 8-row chunks
    const int G = g_kc_tuning.resize_g == 8 ? 8 : g_kc_tuning.resize_g == 32 ? 32 : g_kc_tuning.resize_g == 16 ? 16 : (exact_mode ? 32 : 16);
    const int minb = g_kc_tuning.resize_minb == 8 ? 8 : g_kc_tuning.resize_minb == 6 ? 6 : g_kc_tuning.resize_minb == 4 ? 4 : (G >= 16 ? 4 : 6);
    const int RC = g_kc_tuning.resize_rc == 8 ? 8 : (g_kc_tuning.resize_rc == 16 && G == 16) ? 16 : g_kc_tuning.resize_rc == 4 ? 4 : (exact_mode ? 8 : 4);
    uint32_t pcols = 0;                                 // widest patch of any strip, from the 16-byte boundary below its first column
    for (uint32_t o0 = 0; o0 < dw; o0 += (uint32_t)FT_TW) {
        const uint32_t ol = std::min(o0 + (uint32_t)FT_TW, dw) - 1;
        pcols = std::max(pcols, th->h_left[ol] + th->h_count[ol] - (th->h_left[o0] & ~3u));
    }
    pcols = (pcols + 3u) & ~3u;
    const uint32_t prows = max_window_sliding(*tv, (uint32_t)G);
    const uint32_t vrows = tv->max_taps + 2;
    const bool warp_store = g_kc_tuning.resize_store >= 0;     // default: each warp stores its RC x 128 tiles; -1: the block stores G x 256 halves
    const FtLayout L(pcols, prows, vrows, G, RC, warp_store);
    if (!(pcols <= 256 && prows <= 256 && pcols <= sw && prows <= sh && L.total <= 200 * 1024)) return KC_OK;
    FtPlaneMaps pm;
    CUtensorMap m_vt;
    memset(&pm, 0, sizeof pm);
    if (!make_tensor_map_2d(&m_vt, tv->d_vtab, dh, vrows, (uint32_t)G + 4u, vrows)) return KC_OK;
    for (int i = 0; i < n; ++i)
        if (!make_tensor_map_2d(&pm.src[i], srcs[i], sw, sh, pcols, prows) ||
            !make_tensor_map_2d(&pm.dst[i], dsts[i], dw, nrows, warp_store ? (uint32_t)FT_WARP_COLS : (uint32_t)FT_HALF, warp_store ? (uint32_t)RC : (uint32_t)G))
            return KC_OK;
    const void* fn = nullptr;
#define KC_FT2(E, W) (G == 8 ? (RC == 4 ? (minb == 8 ? (const void*)kc_resize_tma_kernel<E, 8, 4, 8, W> : (const void*)kc_resize_tma_kernel<E, 8, 4, 6, W>)   \
                                        : (minb == 8 ? (const void*)kc_resize_tma_kernel<E, 8, 8, 8, W> : (const void*)kc_resize_tma_kernel<E, 8, 8, 6, W>))  \
                      : G == 16 ? (RC == 4 ? (minb >= 6 ? (const void*)kc_resize_tma_kernel<E, 16, 4, 6, W> : (const void*)kc_resize_tma_kernel<E, 16, 4, 4, W>)  \
                                        : RC == 8 ? (const void*)kc_resize_tma_kernel<E, 16, 8, 4, W> : (const void*)kc_resize_tma_kernel<E, 16, 16, 4, W>)    \
                                : (RC == 4 ? (const void*)kc_resize_tma_kernel<E, 32, 4, 4, W> : (const void*)kc_resize_tma_kernel<E, 32, 8, 4, W>))
#define KC_FT(E) (warp_store ? KC_FT2(E, true) : KC_FT2(E, false))
    fn = exact_mode ? KC_FT(true) : KC_FT(false);
#undef KC_FT2
#undef KC_FT
    static std::map<std::tuple<int, const void*, size_t>, int> occ;
    static std::mutex occ_mu;
    int per_sm = 1;
    {
        std::lock_guard<std::mutex> lk(occ_mu);
        auto it = occ.find(std::make_tuple(ctx->device, fn, (size_t)L.total));
        if (it == occ.end()) {
            KC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
            int nb = 1;
            cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, fn, FT_THREADS, L.total);
            it = occ.emplace(std::make_tuple(ctx->device, fn, (size_t)L.total), std::max(nb, 1)).first;
        }
        per_sm = it->second;
    }
    // one resident wave over all the planes: strips x row-march lanes x planes ~= SMs x resident CTAs
    const uint32_t strips = (dw + FT_TW - 1) / FT_TW;
    const uint32_t ngroups = (nrows + G - 1) / G;
    const uint32_t lanes = std::max<uint32_t>(1u, std::min<uint32_t>(ngroups, (uint32_t)(ctx->sm_count * per_sm) / std::max(strips * (uint32_t)n, 1u)));
    dim3 grid(strips, std::min<uint32_t>(lanes, 65535u), (unsigned)n);
    if (ctx->capture_log) {
        KcFootprint f;
        for (int i = 0; i < n; ++i) {
            f.reads.push_back({srcs[i], (size_t)sw * sh * 4});
            f.writes.push_back({dsts[i], (size_t)dw * nrows * 4});
        }
        ctx->capture_log->push_back(std::move(f));
    }
    KcTimed timed(ctx, KC_KERNEL_RESIZE_H);
    const float one = 1.0f;
    void* args[] = {(void*)&pm, (void*)&m_vt, (void*)&sh, (void*)&dw, (void*)&tv->d_left, (void*)&tv->max_taps,
                    (void*)&th->d_left, (void*)&th->d_count, (void*)&th->d_weights, (void*)&pcols, (void*)&prows, (void*)&one,
                    (void*)&row0, (void*)&nrows, (void*)&clo, (void*)&chi};
    KC_CUDA(cudaLaunchKernel(fn, grid, dim3(FT_THREADS), args, L.total, ctx->stream));
    ctx->kernel_launches++;
    ctx->run_kernels++;
    *done = true;
    return KC_OK;
}

int32_t kck_resize_plane_rows(kc_context* ctx, const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t dw,
                              uint32_t dh, int filter, uint32_t row0, uint32_t nrows) {
    if (dw == 0 || dh == 0 || nrows == 0) return KC_OK;
    if (row0 > dh || nrows > dh - row0) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "rows [%u, %u) are outside the %u-row result", row0, row0 + nrows, dh);
    if (sw == 0 || sh == 0) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "cannot resize an empty plane");
    std::shared_ptr<KcAxisTable> tv, th;
    KC_TRY(get_axis(ctx, sh, dh, filter, tv));
    KC_TRY(get_axis(ctx, sw, dw, filter, th));
    const bool exact_mode = ctx->opts.math_mode == KC_MATH_EXACT;
    // the [0,1] clamp of image-0.24's horizontal pass (unpinned by the reference's goldens: switchable)
    const float clo = ctx->opts.resize_unclamped ? -INFINITY : 0.0f, chi = ctx->opts.resize_unclamped ? INFINITY : 1.0f;
    static const bool no_fused = getenv("KC_RESIZE_TWO_PASS") != nullptr;
    // ---- tensor-map loads of the source patch and the vertical table, tensor-map stores of the result ----
    {
        bool done = false;
        KC_TRY(kck_resize_planes_rows_batched(ctx, &src, &dst, 1, sw, sh, dw, dh, filter, row0, nrows, &done));
        if (done) return KC_OK;
    }
    if (!no_fused && tv->max_taps <= (uint32_t)FS_MAXT && th->max_taps <= (uint32_t)FS_MAXT) {
        // threads per CTA: 4 output columns each.  Narrow CTAs (one or two warps) march
        // independently, so no warp waits at a barrier for another's vertical pass.
        const int env_threads = g_kc_tuning.resize_threads;
        const int threads = env_threads == 32 || env_threads == 64 || env_threads == 128 ? env_threads : FS_DEFAULT_THREADS;
        const uint32_t tw = (uint32_t)threads * FS_CPT;
        // row pitch a multiple of 4 floats: the vertical pass reads column quads (LDS.128); the
        // quad of the last columns may run up to 3 columns past the window
        const uint32_t pcols = (max_window(*th, tw) + 3u) & ~3u;
        const uint32_t prows = max_window_sliding(*tv, FS_G);   // groups start at row0 + 16 k: any alignment
        const size_t smem = sizeof(float) * ((size_t)pcols * FS_TP + 2 * (size_t)prows * pcols) + sizeof(float4) * FS_MAXT * threads + 2 * sizeof(FsGroupBuf);
        if (smem <= 200 * 1024) {
            const void* fn = nullptr;
#define KC_PICK(T) (exact_mode ? (const void*)kc_resize_strip_kernel<true, T> : (const void*)kc_resize_strip_kernel<false, T>)
            fn = threads == 32 ? KC_PICK(32) : threads == 64 ? KC_PICK(64) : KC_PICK(128);
#undef KC_PICK
            // resident CTAs per SM for this (kernel, shared-memory size): asked once
            static std::map<std::tuple<int, const void*, size_t>, int> occupancy;   // per device: the attribute is, too
            static std::mutex occupancy_mu;
            int per_sm = 1;
            {
                std::lock_guard<std::mutex> lk(occupancy_mu);
                auto it = occupancy.find(std::make_tuple(ctx->device, fn, smem));
                if (it == occupancy.end()) {
                    KC_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                    int n = 1;
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, fn, threads, smem);
                    it = occupancy.emplace(std::make_tuple(ctx->device, fn, smem), std::max(n, 1)).first;
                }
                per_sm = it->second;
            }
            // one resident wave: strips x row-march lanes ~= SMs x resident CTAs
            const uint32_t strips = (dw + tw - 1) / tw;
            const uint32_t ngroups = (nrows + FS_G - 1) / FS_G;
            const uint32_t lanes = std::max<uint32_t>(1u, std::min<uint32_t>(ngroups, (uint32_t)(ctx->sm_count * per_sm) / std::max(strips, 1u)));
            {
                dim3 grid(strips, std::min<uint32_t>(lanes, 65535u));
                kc_log_launch(ctx, {{src, (size_t)sw * sh * 4}}, {{dst, (size_t)dw * nrows * 4}});
                KcTimed timed(ctx, KC_KERNEL_RESIZE_H);
                const float one = 1.0f;
                void* args[] = {(void*)&src, (void*)&sw, (void*)&sh, (void*)&dst, (void*)&dw, (void*)&dh,
                                (void*)&tv->d_left, (void*)&tv->d_count, (void*)&tv->d_weights, (void*)&tv->max_taps,
                                (void*)&th->d_left, (void*)&th->d_count, (void*)&th->d_weights,
                                (void*)&pcols, (void*)&prows, (void*)&one, (void*)&row0, (void*)&nrows, (void*)&clo, (void*)&chi};
                KC_CUDA(cudaLaunchKernel(fn, grid, dim3(threads), args, smem, ctx->stream));
                ctx->kernel_launches++;
                ctx->run_kernels++;
                return KC_OK;
            }
        }
    }
    if (row0 != 0 || nrows != dh) {
        // long windows (downsampling): resize the whole plane with the two-pass kernels, keep the rows
        float* full = nullptr;
        const size_t full_bytes = ((sizeof(float) * (size_t)dw * dh + 15) / 16) * 16;
        KC_TRY(kc_dev_alloc(ctx, full_bytes, (void**)&full));
        int32_t rc = kck_resize_plane_rows(ctx, src, sw, sh, full, dw, dh, filter, 0, dh);
        if (rc == KC_OK) {
            cudaError_t e = cudaMemcpyAsync(dst, full + (size_t)row0 * dw, sizeof(float) * (size_t)dw * nrows, cudaMemcpyDeviceToDevice, ctx->stream);
            if (e != cudaSuccess) { kc_set_error("row copy failed: %s", cudaGetErrorString(e)); rc = KC_ERR_CUDA; }
        }
        kc_dev_free(ctx, full, full_bytes);
        return rc;
    }
    float* tmp = nullptr;
    const size_t tmp_bytes = ((sizeof(float) * (size_t)sw * dh + 15) / 16) * 16;
    KC_TRY(kc_dev_alloc(ctx, tmp_bytes, (void**)&tmp));
    const bool exact = ctx->opts.math_mode == KC_MATH_EXACT;
    bool marched = false;
    kc_log_launch(ctx, {{src, (size_t)sw * sh * 4}}, {{tmp, (size_t)sw * dh * 4}});      // the vertical pass below, whichever kernel runs it
    static const bool no_march = getenv("KC_RESIZE_NO_MARCH") != nullptr;
    if (!no_march && (sw & 3u) == 0 && tv->max_taps > (uint32_t)FS_MAXT) {
        KC_TRY(build_march_tables(ctx, *tv));
        if (tv->march_state == 1) {
            // output rows per CTA: as many as keep the tables of its source range within 64 KiB of shared memory
            static const int env_rows = getenv("KC_VM_ROWS") ? atoi(getenv("KC_VM_ROWS")) : 0;
            uint32_t rows = env_rows > 0 ? (uint32_t)env_rows : (uint32_t)VM_ROWS_PER_CTA;
            auto range = [&](uint32_t rpc) {
                uint32_t mx = 0;
                for (uint32_t a = 0; a < dh; a += rpc) {
                    const uint32_t bb = std::min(a + rpc, dh) - 1;
                    mx = std::max(mx, tv->h_left[bb] + tv->h_count[bb] - tv->h_left[a]);
                }
                return mx;
            };
            // the TMA-fed march (default where its shape fits: widths that are multiples of four, a tensor-map encoder at hand)
            static const bool no_vtma = getenv("KC_RESIZE_NO_VTMA") != nullptr;
            if (!no_vtma && tensor_map_encoder() && (((uintptr_t)src | (uintptr_t)tmp) & 15u) == 0 && sw >= (uint32_t)VT_COLS / 2) {
                CUtensorMap m_src;
                // output rows per block: one whole wave of blocks at the kernel's residency (FAST four per SM, EXACT three;
                // two stages either way), never fewer than 16 rows.  Measured 8192^2 -> 1024^2 Lanczos3, ms FAST / EXACT:
                // 3 per SM x 4 stages 0.066 / 0.079, 3 x 2 0.064 / 0.077, 4 x 3 0.062 / 0.079, 4 x 2 0.062 / 0.080.  KC_VT_SHAPE=<ctas><stages> picks another compiled shape
                static const int env_shape = getenv("KC_VT_SHAPE") ? atoi(getenv("KC_VT_SHAPE")) : 0;
                const int shape = env_shape ? env_shape : exact ? 32 : 42;
                const int ctas = shape / 10, stages = shape % 10;
                const uint32_t strips = (sw + VT_COLS - 1) / VT_COLS;
                const uint32_t want_gy = std::max<uint32_t>(1u, (uint32_t)(ctx->sm_count * ctas) / strips);
                static const uint32_t min_rpc = getenv("KC_VT_MIN_ROWS") ? (uint32_t)atoi(getenv("KC_VT_MIN_ROWS")) : 4u;
                uint32_t rpc = std::max<uint32_t>(std::max(1u, min_rpc), (dh + want_gy - 1) / want_gy);
                if (env_rows > 0) rpc = (uint32_t)env_rows;
                const uint32_t gy2 = (dh + rpc - 1) / rpc;
                const size_t smem2 = 128 + (size_t)stages * VT_STAGE_BYTES;
                const void* fn = nullptr;
#define KC_VT(C, S) (exact ? (const void*)kc_resize_v_tma_kernel<true, C, S> : (const void*)kc_resize_v_tma_kernel<false, C, S>)
                switch (shape) {
                    case 34: fn = KC_VT(3, 4); break;
                    case 32: fn = KC_VT(3, 2); break;
                    case 43: fn = KC_VT(4, 3); break;
                    case 42: fn = KC_VT(4, 2); break;
                    default: break;
                }
#undef KC_VT
                if (fn && gy2 <= 65535u && make_tensor_map_2d(&m_src, src, sw, sh, 256u, (uint32_t)VT_R)) {
                    KC_TRY(kc_ensure_smem_attr(ctx, fn, 100 * 1024));
                    const uint32_t sw4 = sw >> 2;
                    float4* tmp4 = (float4*)tmp;
                    const float one = 1.0f;
                    void* args[] = {(void*)&m_src, (void*)&sw4, (void*)&sh, (void*)&tmp4, (void*)&dh, (void*)&tv->d_left, (void*)&tv->d_count,
                                    (void*)&tv->d_march_w2, (void*)&tv->d_march_info, (void*)&rpc, (void*)&one};
                    KcTimed timed(ctx, KC_KERNEL_RESIZE_V);
                    KC_CUDA(cudaLaunchKernel(fn, dim3(strips, gy2), dim3(160), args, smem2, ctx->stream));
                    marched = true;
                }
            }
            while (!marched && rows > 1 && (size_t)range(rows) * 64 > 64 * 1024) rows >>= 1;
            const size_t smem = (size_t)range(rows) * 64;
            const uint32_t gy = (dh + rows - 1) / rows;
            if (!marched && smem <= 64 * 1024 && gy <= 65535u) {
                KC_TRY(kc_ensure_smem_attr(ctx, exact ? (const void*)kc_resize_v_march_kernel<true> : (const void*)kc_resize_v_march_kernel<false>, 64 * 1024));
                dim3 grid(((sw >> 2) + VM_THREADS - 1) / VM_THREADS, gy);
                KcTimed timed(ctx, KC_KERNEL_RESIZE_V);
                if (exact) kc_resize_v_march_kernel<true><<<grid, VM_THREADS, smem, ctx->stream>>>((const float4*)src, sw >> 2, (float4*)tmp, dh, tv->d_left, tv->d_count, (const float4*)tv->d_march_w, (const int4*)tv->d_march_o, rows);
                else kc_resize_v_march_kernel<false><<<grid, VM_THREADS, smem, ctx->stream>>>((const float4*)src, sw >> 2, (float4*)tmp, dh, tv->d_left, tv->d_count, (const float4*)tv->d_march_w, (const int4*)tv->d_march_o, rows);
                marched = true;
            }
        }
    }
    if (!marched) {
        dim3 grid((sw + 255) / 256, std::min<uint32_t>(dh, 65535u));
        KcTimed timed(ctx, KC_KERNEL_RESIZE_V);
        if (exact) kc_resize_v_kernel<true><<<grid, 256, 0, ctx->stream>>>(src, sw, tmp, dh, tv->d_left, tv->d_count, tv->d_weights);
        else kc_resize_v_kernel<false><<<grid, 256, 0, ctx->stream>>>(src, sw, tmp, dh, tv->d_left, tv->d_count, tv->d_weights);
    }
    kc_log_launch(ctx, {{tmp, (size_t)sw * dh * 4}}, {{dst, (size_t)dw * dh * 4}});      // the horizontal pass
    // outputs per block (two per thread) and groups of four rows per block (they share the block's copy of the weights; what
    // limits the resident warps is shared memory).  Measured (scripts/probes/downsample_split.py, four shapes from 4096 -> 512
    // to 8192 -> 2048 columns): 128 outputs x 4 groups is the fastest or tied everywhere it fits (0.0191 / 0.0352 ms against
    // 0.0209 / 0.0374 for 64 x 4 and 0.0230 / 0.0414 for 64 x 1), so: the first shape of that order that fits 96 KB and still
    // gives every SM a block; failing that, the one with the most blocks.
    static const int env_ht = getenv("KC_RESIZE_HT") ? atoi(getenv("KC_RESIZE_HT")) : 0;
    static const int env_hrw = getenv("KC_RESIZE_HRW") ? atoi(getenv("KC_RESIZE_HRW")) : 0;
    const uint32_t hgroups = (dh + HT_ROWS - 1) / HT_ROWS;
    uint32_t ht = 0, hrw = 0, htile_f4 = 0;
    size_t hsmem = 0;
    {
        static const uint32_t order[][2] = {{128, 4}, {128, 2}, {64, 4}, {64, 2}, {256, 1}, {128, 1}, {64, 1}};
        uint64_t most = 0;
        for (const auto& c : order) {
            const uint32_t t = c[0], rw = c[1];
            if ((env_ht > 0 && (uint32_t)env_ht != t) || (env_hrw > 0 && (uint32_t)env_hrw != rw)) continue;
            const uint32_t win = max_window(*th, t) + 8;                   // the staging may start up to 3 columns early and end up to 3 late
            const uint32_t tile = win + (win >> 4) + 2;
            const size_t sm = 16 + sizeof(float) * (size_t)th->max_taps * t + sizeof(float4) * (size_t)tile * rw;
            if (sm > 96 * 1024) continue;
            const uint64_t blocks = (uint64_t)((dw + t - 1) / t) * ((hgroups + rw - 1) / rw);
            if (blocks > most) { most = blocks; ht = t; hrw = rw; hsmem = sm; htile_f4 = tile; }
            if (blocks >= (uint64_t)ctx->sm_count) break;
        }
    }
    const uint32_t hgy = hrw ? (hgroups + hrw - 1) / hrw : 0;
    if (!no_march && th->max_taps > (uint32_t)FS_MAXT && ht != 0 && hgy <= 65535u) {
        const void* fn = ht == 256 ? (exact ? (const void*)kc_resize_h_tile_kernel<true, 128> : (const void*)kc_resize_h_tile_kernel<false, 128>)
                       : ht == 128 ? (exact ? (const void*)kc_resize_h_tile_kernel<true, 64> : (const void*)kc_resize_h_tile_kernel<false, 64>)
                                   : (exact ? (const void*)kc_resize_h_tile_kernel<true, 32> : (const void*)kc_resize_h_tile_kernel<false, 32>);
        KC_TRY(kc_ensure_smem_attr(ctx, fn, 96 * 1024));
        const float* weo = nullptr;
        KC_TRY(build_block_weights(ctx, *th, ht, &weo));
        const float* tmp_c = tmp;
        const float one = 1.0f;
        void* args[] = {(void*)&tmp_c, (void*)&sw, (void*)&dst, (void*)&dw, (void*)&dh, (void*)&th->d_left, (void*)&th->d_count, (void*)&weo,
                        (void*)&th->max_taps, (void*)&htile_f4, (void*)&clo, (void*)&chi, (void*)&one};
        KcTimed timed(ctx, KC_KERNEL_RESIZE_H);
        KC_CUDA(cudaLaunchKernel(fn, dim3((dw + ht - 1) / ht, hgy), dim3(ht / 2, hrw), args, hsmem, ctx->stream));
    } else {
        dim3 grid((dw + 255) / 256, std::min<uint32_t>(dh, 65535u));
        KcTimed timed(ctx, KC_KERNEL_RESIZE_H);
        if (exact) kc_resize_h_kernel<true><<<grid, 256, 0, ctx->stream>>>(tmp, sw, dst, dw, dh, th->d_left, th->d_count, th->d_weights, clo, chi);
        else kc_resize_h_kernel<false><<<grid, 256, 0, ctx->stream>>>(tmp, sw, dst, dw, dh, th->d_left, th->d_count, th->d_weights, clo, chi);
    }
    cudaError_t e = cudaGetLastError();
    kc_dev_free(ctx, tmp, tmp_bytes);
    KC_CUDA(e);
    ctx->kernel_launches += 2;
    ctx->run_kernels += 2;
    return KC_OK;
}
