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
    KC_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->axis_tables[key] = t;
    out = t;
    return KC_OK;
}

template <bool EXACT>
__device__ __forceinline__ float tap(float acc, float s, float w) {
    return EXACT ? __fadd_rn(acc, __fmul_rn(s, w)) : fmaf(s, w, acc);
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
                                                          const uint32_t* __restrict__ count, const float* __restrict__ wh) {
    const uint32_t ox = blockIdx.x * blockDim.x + threadIdx.x;
    if (ox >= dw) return;
    const uint32_t l = left[ox], n = count[ox];
    for (uint32_t y = blockIdx.y; y < dh; y += gridDim.y) {
        const float* row = tmp + (size_t)y * sw + l;
        float acc = 0.0f;
        for (uint32_t j = 0; j < n; ++j) acc = tap<EXACT>(acc, __ldg(row + j), wh[(size_t)j * dw + ox]);
        // image::math::utils::clamp keeps NaN
        acc = acc < 0.0f ? 0.0f : (acc > 1.0f ? 1.0f : acc);
        __stcs(dst + (size_t)y * dw + ox, acc);
    }
}

// ---------------------------------------------------------------------------
// Fused V∘H tile kernel for resizes whose tap windows are short (<= FT_MAXT taps
// per axis: every upsampling, and mild downsampling).  One CTA produces a
// FT_TW x FT_TH tile of the output: it stages the source patch the tile depends on
// in shared memory, runs the vertical pass for the tile's rows into a second
// shared-memory buffer (the unclamped f32 intermediate of the reference, never
// written to HBM), then each thread owns one output column: its horizontal
// weights sit in registers and it walks down the tile's rows, so a row of the
// tile is one coalesced 128-byte store per warp.  Same operation order as the
// two-pass kernels: vertical taps top to bottom, then horizontal taps left to
// right, clamp to [0,1] last.
// ---------------------------------------------------------------------------
constexpr int FT_THREADS = 128;
constexpr int FT_CPT = 4;                    // consecutive output columns per thread (one float4 store per row)
constexpr int FT_TW = FT_THREADS * FT_CPT;   // output columns per CTA
constexpr int FT_TH = 16;                    // output rows per CTA
constexpr int FT_MAXT = 8;                   // taps per axis this kernel supports
constexpr int FT_TP = FT_TH + 4;             // pitch of the column-major intermediate: 16-byte aligned, conflict-free

template <bool EXACT>
__global__ void __launch_bounds__(FT_THREADS) kc_resize_fused_kernel(const float* __restrict__ src, uint32_t sw, uint32_t sh,
                                                                     float* __restrict__ dst, uint32_t dw, uint32_t dh,
                                                                     const uint32_t* __restrict__ vleft, const uint32_t* __restrict__ vcount,
                                                                     const float* __restrict__ vw, const uint32_t* __restrict__ hleft,
                                                                     const uint32_t* __restrict__ hcount, const float* __restrict__ hw,
                                                                     uint32_t pcols, uint32_t prows) {
    extern __shared__ __align__(16) float fsm[];
    float* Tm = fsm;                                  // [pcols][FT_TP]   vertical-pass result, column-major
    float* S = Tm + (size_t)pcols * FT_TP;            // [prows][pcols]   source patch
    float* Wv = S + (size_t)prows * pcols;            // [FT_TH][FT_MAXT] vertical weights of the tile's rows
    __shared__ uint32_t vl[FT_TH], vc[FT_TH];
    const int tid = threadIdx.x;
    const uint32_t ox0 = blockIdx.x * FT_TW, oy0 = blockIdx.y * FT_TH;
    const uint32_t oxl = min(ox0 + FT_TW, dw) - 1, oyl = min(oy0 + FT_TH, dh) - 1;  // last valid column / row
    const uint32_t nrow = oyl - oy0 + 1;
    // this thread's four columns: windows and horizontal weights are requested up front,
    // so their latency overlaps the patch load and the vertical pass
    uint32_t left[FT_CPT], cnt[FT_CPT];
    float w[FT_CPT][FT_MAXT];
#pragma unroll
    for (int c = 0; c < FT_CPT; ++c) {
        const uint32_t ox = min(ox0 + FT_CPT * tid + c, oxl);
        left[c] = __ldg(hleft + ox);
        cnt[c] = __ldg(hcount + ox);
#pragma unroll
        for (int j = 0; j < FT_MAXT; ++j) w[c][j] = __ldg(hw + (size_t)min((uint32_t)j, cnt[c] - 1) * dw + ox);
    }
    // source window of the tile (left[] and left[]+count[] are non-decreasing in o)
    const uint32_t cx0 = __ldg(hleft + ox0), cx1 = __ldg(hleft + oxl) + __ldg(hcount + oxl);
    const uint32_t ry0 = __ldg(vleft + oy0), ry1 = __ldg(vleft + oyl) + __ldg(vcount + oyl);
    const uint32_t ncx = cx1 - cx0, nry = ry1 - ry0;
    if (tid < FT_TH) {
        const bool live = (uint32_t)tid < nrow;
        vl[tid] = live ? vleft[oy0 + tid] - ry0 : 0u;
        vc[tid] = live ? vcount[oy0 + tid] : 0u;   // rows past the image compute nothing
    }
    for (int i = tid; i < (int)nrow * FT_MAXT; i += FT_THREADS) {
        const int r = i / FT_MAXT, k = i - r * FT_MAXT;
        Wv[i] = vw[(size_t)k * dh + oy0 + r];  // tap-major table; entries past count are never used
    }
    for (uint32_t i = tid; i < nry * ncx; i += FT_THREADS) {
        const uint32_t r = i / ncx, c = i - r * ncx;
        S[r * pcols + c] = __ldg(src + (size_t)(ry0 + r) * sw + cx0 + c);
    }
    __syncthreads();
    // vertical pass: Tm[c][r] = sum_i S[vl[r]+i][c] * Wv[r][i]   (all FT_TH rows, dead ones give 0)
    for (uint32_t i = tid; i < (uint32_t)FT_TH * ncx; i += FT_THREADS) {
        const uint32_t c = i / FT_TH, r = i - c * FT_TH;
        const uint32_t l = vl[r], n = vc[r];
        float acc = 0.0f;
        for (uint32_t k = 0; k < n; ++k) acc = tap<EXACT>(acc, S[(l + k) * pcols + c], Wv[r * FT_MAXT + k]);
        Tm[c * FT_TP + r] = acc;
    }
    __syncthreads();
    // horizontal pass: four adjacent output columns per thread, all rows of the tile accumulate
    // in registers; taps outermost (the per-column tap count is tested once per tap, not per
    // pixel), the intermediate is read four rows at a time (LDS.128)
    if (ox0 + FT_CPT * tid > oxl) return;
    float acc[FT_CPT][FT_TH];
#pragma unroll
    for (int c = 0; c < FT_CPT; ++c)
#pragma unroll
        for (int r = 0; r < FT_TH; ++r) acc[c][r] = 0.0f;
#pragma unroll
    for (int c = 0; c < FT_CPT; ++c) {
        const uint32_t l = left[c] - cx0;
#pragma unroll
        for (int j = 0; j < FT_MAXT; ++j) {
            if ((uint32_t)j < cnt[c]) {
                const float4* t = reinterpret_cast<const float4*>(Tm + (size_t)(l + j) * FT_TP);
#pragma unroll
                for (int q = 0; q < FT_TH / 4; ++q) {
                    const float4 v = t[q];
                    acc[c][4 * q + 0] = tap<EXACT>(acc[c][4 * q + 0], v.x, w[c][j]);
                    acc[c][4 * q + 1] = tap<EXACT>(acc[c][4 * q + 1], v.y, w[c][j]);
                    acc[c][4 * q + 2] = tap<EXACT>(acc[c][4 * q + 2], v.z, w[c][j]);
                    acc[c][4 * q + 3] = tap<EXACT>(acc[c][4 * q + 3], v.w, w[c][j]);
                }
            }
        }
    }
    const uint32_t oxt = ox0 + FT_CPT * tid;
    float* out = dst + (size_t)oy0 * dw + oxt;
    const bool vec = ((dw & 3u) == 0) && (oxt + 3 <= oxl);
#pragma unroll
    for (int r = 0; r < FT_TH; ++r) {
        if ((uint32_t)r < nrow) {
            float v[FT_CPT];
#pragma unroll
            for (int c = 0; c < FT_CPT; ++c) {
                const float a = acc[c][r];
                v[c] = a < 0.0f ? 0.0f : (a > 1.0f ? 1.0f : a);  // image::math::utils::clamp keeps NaN
            }
            if (vec) {
                __stcs(reinterpret_cast<float4*>(out + (size_t)r * dw), make_float4(v[0], v[1], v[2], v[3]));
            } else {
#pragma unroll
                for (int c = 0; c < FT_CPT; ++c)
                    if (oxt + c <= oxl) out[(size_t)r * dw + c] = v[c];
            }
        }
    }
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

}  // namespace

int32_t kck_resize_plane(kc_context* ctx, const float* src, uint32_t sw, uint32_t sh, float* dst, uint32_t dw,
                         uint32_t dh, int filter) {
    if (dw == 0 || dh == 0) return KC_OK;
    if (sw == 0 || sh == 0) KC_FAIL(KC_ERR_INVALID_ARGUMENT, "cannot resize an empty plane");
    std::shared_ptr<KcAxisTable> tv, th;
    KC_TRY(get_axis(ctx, sh, dh, filter, tv));
    KC_TRY(get_axis(ctx, sw, dw, filter, th));
    const bool exact_mode = ctx->opts.math_mode == KC_MATH_EXACT;
    static const bool no_fused = getenv("KC_RESIZE_TWO_PASS") != nullptr;
    if (!no_fused && tv->max_taps <= (uint32_t)FT_MAXT && th->max_taps <= (uint32_t)FT_MAXT) {
        const uint32_t pcols = max_window(*th, FT_TW) | 1u;  // odd row pitch: no systematic bank conflicts
        const uint32_t prows = max_window(*tv, FT_TH);
        const size_t smem = sizeof(float) * ((size_t)prows * pcols + (size_t)pcols * FT_TP + (size_t)FT_TH * FT_MAXT) + 16;
        if (smem <= 200 * 1024) {
            static bool attr_set = false;
            if (!attr_set) {
                KC_CUDA(cudaFuncSetAttribute(kc_resize_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                KC_CUDA(cudaFuncSetAttribute(kc_resize_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                attr_set = true;
            }
            dim3 grid((dw + FT_TW - 1) / FT_TW, (dh + FT_TH - 1) / FT_TH);
            if (grid.y <= 65535u) {
                KcTimed timed(ctx, KC_KERNEL_RESIZE_H);
                if (exact_mode)
                    kc_resize_fused_kernel<true><<<grid, FT_THREADS, smem, ctx->stream>>>(src, sw, sh, dst, dw, dh, tv->d_left, tv->d_count, tv->d_weights,
                                                                                      th->d_left, th->d_count, th->d_weights, pcols, prows);
                else
                    kc_resize_fused_kernel<false><<<grid, FT_THREADS, smem, ctx->stream>>>(src, sw, sh, dst, dw, dh, tv->d_left, tv->d_count, tv->d_weights,
                                                                                       th->d_left, th->d_count, th->d_weights, pcols, prows);
                KC_CUDA(cudaGetLastError());
                ctx->kernel_launches++;
                ctx->run_kernels++;
                return KC_OK;
            }
        }
    }
    float* tmp = nullptr;
    const size_t tmp_bytes = ((sizeof(float) * (size_t)sw * dh + 15) / 16) * 16;
    KC_TRY(kc_dev_alloc(ctx, tmp_bytes, (void**)&tmp));
    const bool exact = ctx->opts.math_mode == KC_MATH_EXACT;
    {
        dim3 grid((sw + 255) / 256, std::min<uint32_t>(dh, 65535u));
        KcTimed timed(ctx, KC_KERNEL_RESIZE_V);
        if (exact) kc_resize_v_kernel<true><<<grid, 256, 0, ctx->stream>>>(src, sw, tmp, dh, tv->d_left, tv->d_count, tv->d_weights);
        else kc_resize_v_kernel<false><<<grid, 256, 0, ctx->stream>>>(src, sw, tmp, dh, tv->d_left, tv->d_count, tv->d_weights);
    }
    {
        dim3 grid((dw + 255) / 256, std::min<uint32_t>(dh, 65535u));
        KcTimed timed(ctx, KC_KERNEL_RESIZE_H);
        if (exact) kc_resize_h_kernel<true><<<grid, 256, 0, ctx->stream>>>(tmp, sw, dst, dw, dh, th->d_left, th->d_count, th->d_weights);
        else kc_resize_h_kernel<false><<<grid, 256, 0, ctx->stream>>>(tmp, sw, dst, dw, dh, th->d_left, th->d_count, th->d_weights);
    }
    cudaError_t e = cudaGetLastError();
    kc_dev_free(ctx, tmp, tmp_bytes);
    KC_CUDA(e);
    ctx->kernel_launches += 2;
    ctx->run_kernels += 2;
    return KC_OK;
}
