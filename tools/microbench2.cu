// shared-memory atomic flavours at the occupancy of the backward kernel (2 CTAs x 256 threads per SM), 12 independent
// updates per "pixel" at a 3-float lane stride: (a) atomicAdd(float) = CAS spin loop, (b) atomicAdd(int) native,
// (c) atomicAdd(int) result unused, (d) plain LDS+FADD+STS.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)
template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters)
{
    extern __shared__ float s[];
    const int n = 9216;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s[i] = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int* si = reinterpret_cast<int*>(s);
    float acc = 0;
    for (int it = 0; it < iters; ++it) {
        const int base = ((w * 37 + it * 5) % 24) * 252 + lane * 3 + (it & 7) * 3;     // row in a 36 x 84 x 3 box
#pragma unroll
        for (int t = 0; t < 4; ++t)
#pragma unroll
            for (int ch = 0; ch < 3; ++ch) {
                const int a = base + (t & 1) * 3 + (t >> 1) * 252 + ch;
                if (MODE == 0) atomicAdd(&s[a], 1.0f);
                if (MODE == 1) acc += (float)atomicAdd(&si[a], it * 3 + lane + ch);
                if (MODE == 2) atomicAdd(&si[a], it * 3 + lane + ch);
                if (MODE == 3) s[a] += 1.0f;
            }
        if (MODE == 3) __syncwarp();
    }
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = s[7] + acc;
}
int main()
{
    float* o; CK(cudaMalloc(&o, 4096 * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 2000; float ms;
    const char* names[4] = {"atomicAdd float (CAS spin)", "atomicAdd int, result used", "atomicAdd int, result unused (RED-like)", "plain LDS+FADD+STS"};
    for (int ctas_per_sm : {2, 4, 8}) {
        const int grid = 148 * ctas_per_sm; const size_t smem = (ctas_per_sm == 2 ? 100 : (ctas_per_sm == 4 ? 50 : 37)) * 1024;
#define RUN(M) CK(cudaFuncSetAttribute(k<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024)); k<M><<<grid, 256, smem>>>(o, iters); CK(cudaDeviceSynchronize()); \
        cudaEventRecord(e0); k<M><<<grid, 256, smem>>>(o, iters); cudaEventRecord(e1); CK(cudaDeviceSynchronize()); cudaEventElapsedTime(&ms, e0, e1); \
        printf("%d CTAs/SM  %-42s %8.1f us  -> %6.2f cycles per warp-instruction per SM (at 1.92 GHz)\n", ctas_per_sm, names[M], ms * 1e3, ms * 1e-3 * 1.92e9 / ((double)ctas_per_sm * 8 * iters * 12));
        RUN(0) RUN(1) RUN(2) RUN(3)
    }
    return 0;
}
