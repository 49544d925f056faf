// TMA box-load throughput on the warp kernels' access pattern (not product code): persistent CTAs, one producer thread
// issuing cp.async.bulk.tensor.3d box loads S stages deep, one consumer thread releasing the stages.  No compute.
// usage: tma_bw  bi br  tw_f th  S ctas_per_sm  xoff  [N]
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda.h>
#include <cuda_runtime.h>
#include "../deep-online-video-stabilization_b200/csrc/mgw_tma.cuh"
using namespace mgw;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(2); } } while (0)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static CUtensorMap make(void* base, int inner, int rows, int N, int bi, int br)
{
    void* p = nullptr; cudaDriverEntryPointQueryResult q;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q));
    cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)N};
    cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * 4 * rows};
    cuuint32_t box[3] = {(cuuint32_t)bi, (cuuint32_t)br, 1};
    cuuint32_t es[3] = {1, 1, 1};
    CUtensorMap m;
    CUresult r = ((EncodeTiledFn)p)(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); exit(3); }
    return m;
}

__global__ void k(const __grid_constant__ CUtensorMap m, int bi, int br, int twf, int th, int S, int ntx, int nty, int total, int xoff, int yoff, float* sink)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int stage = ((bi * br * 4 + 127) / 128) * 128;
    uint64_t* full = (uint64_t*)(smem + (size_t)S * stage);
    uint64_t* empty = full + S;
    if (threadIdx.x == 0) { for (int s = 0; s < S; ++s) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, 1); } tma::fence_barrier_init(); }
    __syncthreads();
    if (threadIdx.x == 32) {
        for (int it = 0, t = blockIdx.x; t < total; ++it, t += gridDim.x) {
            const int s = it % S;
            tma::mbar_wait(empty + s, ((it / S) & 1) ^ 1);
            const int tx = t % ntx, q = t / ntx, ty = q % nty, n = q / nty;
            tma::mbar_expect_tx(full + s, bi * br * 4);
            int x = tx * twf + xoff, y = ty * th + yoff;
            tma::load_3d(smem + (size_t)s * stage, &m, full + s, x < 0 ? 0 : x, y < 0 ? 0 : y, n);
        }
    } else if (threadIdx.x == 0) {
        float acc = 0;
        for (int it = 0, t = blockIdx.x; t < total; ++it, t += gridDim.x) {
            const int s = it % S;
            tma::mbar_wait(full + s, (it / S) & 1);
            acc += ((float*)(smem + (size_t)s * stage))[5];
            tma::mbar_arrive(empty + s);
        }
        if (acc == 123.456f) *sink = acc;
    }
}

int main(int argc, char** argv)
{
    const int bi = atoi(argv[1]), br = atoi(argv[2]), twf = atoi(argv[3]), th = atoi(argv[4]), S = atoi(argv[5]), cps = atoi(argv[6]);
    const int xoff = atoi(argv[7]);
    const int N = argc > 8 ? atoi(argv[8]) : 96;
    const int H = 288, inner = 512 * 3;
    float* d; CK(cudaMalloc(&d, (size_t)N * H * inner * 4)); CK(cudaMemset(d, 0, (size_t)N * H * inner * 4));
    float* flush; CK(cudaMalloc(&flush, 256u << 20));
    float* sink; CK(cudaMalloc(&sink, 4));
    CUtensorMap m = make(d, inner, H, N, bi, br);
    const int ntx = inner / twf, nty = H / th, total = N * ntx * nty;
    const int stage = ((bi * br * 4 + 127) / 128) * 128;
    const size_t smem = (size_t)S * stage + 2 * S * 8 + 64;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e9;
    for (int rep = 0; rep < 5; ++rep) {
        CK(cudaMemset(flush, rep, 256u << 20));
        cudaEventRecord(e0);
        k<<<148 * cps, 64, smem>>>(m, bi, br, twf, th, S, ntx, nty, total, xoff, -((br - th) / 2), sink);
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double uniq = (double)N * H * inner * 4, boxb = (double)total * bi * br * 4;
    printf("box %3dx%2d tile %3dx%2d S=%d ctas/SM=%d xoff=%d smem/CTA=%zuKB: %7.1f us  unique %6.0f GB/s  box fill %6.0f GB/s\n", bi, br, twf, th, S, cps, xoff,
           smem >> 10, best * 1e3, uniq / best * 1e-6, boxb / best * 1e-6);
    return 0;
}
