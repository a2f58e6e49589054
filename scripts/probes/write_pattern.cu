// Write-only HBM bandwidth as a function of the store pattern (probe, not product code):
// a 8192x8192 f32 plane (256 MiB) is filled by CTAs that each own a strip of SEG columns and
// march down groups of 16 rows, like kc_resize_strip_kernel, versus a linear grid-stride fill.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o write_pattern write_pattern.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void strip_fill(float4* dst, int w4, int h, int seg4, int rows_per_group, float v) {
    // blockDim.x threads cover seg4 float4 columns (seg4 == blockDim.x)
    const int strip = blockIdx.x, lane = blockIdx.y, lanes = gridDim.y;
    const int ngroups = h / rows_per_group;
    const int x = strip * seg4 + threadIdx.x;
    for (int g = lane; g < ngroups; g += lanes) {
        float4* o = dst + (size_t)g * rows_per_group * w4 + x;
#pragma unroll 16
        for (int r = 0; r < rows_per_group; ++r) { __stcs(o, make_float4(v, v, v, v)); o += w4; }
    }
}
__global__ void linear_fill(float4* dst, size_t n4, float v) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x)
        __stcs(dst + i, make_float4(v, v, v, v));
}
// each CTA writes whole tiles of `tile4` consecutive float4 (linear chunks), persistent
__global__ void tile_fill(float4* dst, size_t n4, int tile4, float v) {
    const size_t ntiles = n4 / tile4;
    for (size_t t = blockIdx.x; t < ntiles; t += gridDim.x)
        for (int i = threadIdx.x; i < tile4; i += blockDim.x) __stcs(dst + t * tile4 + i, make_float4(v, v, v, v));
}
template <class F> float timeit(F f) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; ++i) f();
    float best = 1e9f;
    for (int i = 0; i < 10; ++i) { cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b); float ms; cudaEventElapsedTime(&ms, a, b); if (ms < best) best = ms; }
    return best;
}
int main() {
    const int W = 8192, H = 8192; const size_t n4 = (size_t)W * H / 4; const double bytes = (double)W * H * 4;
    float4* d; cudaMalloc(&d, bytes);
    float ms = timeit([&] { linear_fill<<<148 * 8, 256>>>(d, n4, 1.f); });
    printf("linear grid-stride 148x8x256            %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    for (int tile4 : {1024, 4096}) {
        ms = timeit([&] { tile_fill<<<148 * 4, 256>>>(d, n4, tile4, 1.f); });
        printf("linear tiles of %5d B, 592 CTAs        %7.1f us  %6.0f GB/s\n", tile4 * 16, ms * 1e3, bytes / ms / 1e6);
    }
    for (int threads : {32, 128, 256, 512}) {
        for (int rpg : {16, 64}) {
            const int strips = W / 4 / threads; const int per_sm = 2048 / threads > 16 ? 16 : 2048 / threads;
            int lanes = 148 * per_sm / strips; if (lanes < 1) lanes = 1; if (lanes > H / rpg) lanes = H / rpg;
            ms = timeit([&] { strip_fill<<<dim3(strips, lanes), threads>>>(d, W / 4, H, threads, rpg, 1.f); });
            printf("strips of %5d B x %2d rows, grid %3dx%3d   %7.1f us  %6.0f GB/s\n", threads * 16, rpg, strips, lanes, ms * 1e3, bytes / ms / 1e6);
        }
    }
    // the resize kernel's actual shape: 128 threads, 4 resident CTAs/SM
    ms = timeit([&] { strip_fill<<<dim3(16, 37), 128>>>(d, W / 4, H, 128, 16, 1.f); });
    printf("strips of  2048 B x 16 rows, grid  16x 37   %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    ms = timeit([&] { strip_fill<<<dim3(16, 74), 128>>>(d, W / 4, H, 128, 16, 1.f); });
    printf("strips of  2048 B x 16 rows, grid  16x 74   %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    ms = timeit([&] { strip_fill<<<dim3(16, 148), 128>>>(d, W / 4, H, 128, 16, 1.f); });
    printf("strips of  2048 B x 16 rows, grid  16x148   %7.1f us  %6.0f GB/s\n", ms * 1e3, bytes / ms / 1e6);
    cudaFree(d);
    return 0;
}
