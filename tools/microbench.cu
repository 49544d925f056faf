// Design microbenchmarks for the backward scatter (not product code): what does B200 give for
//  (a) scalar global RED.F32, coalesced          (b) red.global.add.v4.f32
//  (c) bulk smem->global reduce-add (cp.reduce.async.bulk)   (d) shared float atomicAdd (CAS loop)
//  (e) shared LDS+FADD+STS read-modify-write     (f) 12 scalar REDs per pixel at 12-byte lane stride (the naive dU scatter)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

__global__ void k_red_scalar(float* dst, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
            atomicAdd(dst + i, 1.0f);
}
__global__ void k_red_v4(float* dst, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" :: "l"(dst + 4 * i), "f"(1.0f), "f"(1.0f), "f"(1.0f), "f"(1.0f) : "memory");
}
__global__ void k_store(float* dst, size_t n, int reps) {
    for (int r = 0; r < reps; ++r)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n / 4; i += (size_t)gridDim.x * blockDim.x)
            reinterpret_cast<float4*>(dst)[i] = make_float4(1, 1, 1, 1);
}
// each CTA owns tiles of ROWS x ROWB bytes in smem and bulk-reduces them row by row into a 2-D region of dst
template <int ROWS, int ROWB>
__global__ void k_bulk_reduce(float* dst, size_t pitch_f, int tiles_x, int tiles_y, int reps) {
    extern __shared__ __align__(128) float sm[];
    for (int i = threadIdx.x; i < ROWS * ROWB / 4; i += blockDim.x) sm[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    for (int r = 0; r < reps; ++r)
        for (int t = blockIdx.x; t < tiles_x * tiles_y; t += gridDim.x) {
            const int ty = t / tiles_x, tx = t % tiles_x;
            if (threadIdx.x < ROWS) {
                float* g = dst + ((size_t)ty * ROWS + threadIdx.x) * pitch_f + (size_t)tx * (ROWB / 4);
                const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + threadIdx.x * (ROWB / 4));
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" :: "l"(g), "r"(s), "n"(ROWB) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            if (threadIdx.x < ROWS) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            __syncthreads();
        }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__global__ void k_smem_atomic(float* out, int iters, int stride) {
    __shared__ float s[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = 0;
    __syncthreads();
    int a = (threadIdx.x * stride) & 8191;
    for (int i = 0; i < iters; ++i) { atomicAdd(&s[a], 1.0f); a = (a + 32 * stride) & 8191; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}
__global__ void k_smem_rmw(float* out, int iters, int stride) {
    __shared__ float s[8192];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) s[i] = 0;
    __syncthreads();
    int a = (threadIdx.x * stride) & 8191;
    for (int i = 0; i < iters; ++i) { s[a] += 1.0f; __syncwarp(); a = (a + 32 * stride) & 8191; }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[0];
}
// naive dU scatter: thread = pixel, 4 taps x 3 channels scalar REDs into an NHWC image (identity-ish warp)
__global__ void k_scatter12(float* dst, int Himg, int Wimg, int N) {
    const size_t P = (size_t)N * Himg * Wimg;
    for (size_t p = blockIdx.x * (size_t)blockDim.x + threadIdx.x; p < P; p += (size_t)gridDim.x * blockDim.x) {
        const int c = p % Wimg, r = (p / Wimg) % Himg; const size_t n = p / ((size_t)Wimg * Himg);
        const int x0 = min(c, Wimg - 2), y0 = min(r, Himg - 2);
        float* b = dst + n * (size_t)Himg * Wimg * 3;
#pragma unroll
        for (int dy = 0; dy < 2; ++dy)
#pragma unroll
            for (int dx = 0; dx < 2; ++dx)
#pragma unroll
                for (int ch = 0; ch < 3; ++ch) atomicAdd(b + ((size_t)(y0 + dy) * Wimg + x0 + dx) * 3 + ch, 0.25f);
    }
}

int main() {
    const size_t n = (size_t)32 * 288 * 512 * 3;      // floats in dU at config #2 (56.6 MB)
    float* d; CK(cudaMalloc(&d, n * 4 + 4096)); CK(cudaMemset(d, 0, n * 4));
    float* o; CK(cudaMalloc(&o, 4096 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int grid = 148 * 8;
#define TIME(name, bytes, ...) do { __VA_ARGS__; CK(cudaDeviceSynchronize()); cudaEventRecord(e0); for (int it = 0; it < 5; ++it) { __VA_ARGS__; } \
        cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); ms /= 5; \
        printf("%-44s %8.1f us  %8.1f GB/s (payload)\n", name, ms * 1e3, (bytes) / (ms * 1e-3) / 1e9); } while (0)
    TIME("store float4 (56.6MB)", n * 4.0, (k_store<<<grid, 256>>>(d, n, 1)));
    TIME("RED.F32 scalar coalesced (56.6MB)", n * 4.0, (k_red_scalar<<<grid, 256>>>(d, n, 1)));
    TIME("red.v4.f32 coalesced (56.6MB)", n * 4.0, (k_red_v4<<<grid, 256>>>(d, n, 1)));
    {
        constexpr int ROWS = 32, ROWB = 1152;        // 32 rows x 96 px x 3 ch
        const size_t pitch_f = 512 * 3; const int tiles_x = 512 * 3 * 4 / ROWB, tiles_y = 32 * 288 / ROWS;
        TIME("bulk reduce-add 32x1152B tiles (56.6MB)", n * 4.0, (k_bulk_reduce<ROWS, ROWB><<<148 * 4, 128, ROWS * ROWB>>>(d, pitch_f, tiles_x, tiles_y, 1)));
    }
    {
        constexpr int ROWS = 16, ROWB = 384;         // small tiles: 16 rows x 32 px x 3 ch
        const size_t pitch_f = 512 * 3; const int tiles_x = 512 * 3 * 4 / ROWB, tiles_y = 32 * 288 / ROWS;
        TIME("bulk reduce-add 16x384B tiles (56.6MB)", n * 4.0, (k_bulk_reduce<ROWS, ROWB><<<148 * 8, 128, ROWS * ROWB>>>(d, pitch_f, tiles_x, tiles_y, 1)));
    }
    TIME("naive scatter 12 RED/px (4.7 Mpx)", 32.0 * 288 * 512 * 12 * 4, (k_scatter12<<<grid, 256>>>(d, 288, 512, 32)));
    const int iters = 4096;
    for (int stride : {1, 3, 12}) {
        char nm[64];
        snprintf(nm, 64, "smem atomicAdd f32 CAS, lane stride %d", stride);
        TIME(nm, 148.0 * 8 * 256 * iters * 4, (k_smem_atomic<<<148 * 8, 256>>>(o, iters, stride)));
        snprintf(nm, 64, "smem LDS+FADD+STS rmw, lane stride %d", stride);
        TIME(nm, 148.0 * 8 * 256 * iters * 4, (k_smem_rmw<<<148 * 8, 256>>>(o, iters, stride)));
    }
    printf("done\n");
    return 0;
}
