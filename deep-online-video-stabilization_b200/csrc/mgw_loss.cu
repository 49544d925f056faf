// Loss epilogues fused onto the warped output (north_star part 4):
//   img_loss      s_net_bundle_nobm.py:347-352
//   feature_loss  s_net_bundle_nobm.py:335-343 with warp_pts :215-230
//   temp_loss     train_bundle_nobm.py:115-125 (two interpolate() passes + masked MSE in ONE kernel)
// Forward kernels produce the per-sample partial sums the losses are built from; the final O(N) scalar
// arithmetic (s0/(s1+1e-8), /batch) is left to the caller so the multi-GPU path can divide by the GLOBAL batch.
#include <cstdlib>

#include "mgw_internal.h"

namespace mgw {

__device__ __forceinline__ float block_sum(float v, float* sh)
{
    v = warp_sum(v);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) sh[w] = v;
    __syncthreads();
    v = (threadIdx.x < (blockDim.x >> 5)) ? sh[threadIdx.x] : 0.0f;
    if (w == 0) v = warp_sum(v);
    return v;                       // valid in thread 0
}

// ---------------------------------------------------------------- img_loss
__global__ void __launch_bounds__(256)
img_loss_fwd_kernel(const float* __restrict__ out, const float* __restrict__ y, const float* __restrict__ black,
                    int HW, int C, float* __restrict__ sums)
{
    __shared__ float sh[8];
    const int n = blockIdx.y;
    float se = 0.0f, sm = 0.0f;
    for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < HW; q += gridDim.x * blockDim.x) {
        const size_t p = (size_t)n * HW + q;
        const float nb = 1.0f - __ldg(black + p);
        sm += nb;
        for (int ch = 0; ch < C; ++ch) {
            const float e = (__ldg(out + p * C + ch) - __ldg(y + p * C + ch)) * nb;
            se = fmaf(e, e, se);
        }
    }
    se = block_sum(se, sh);
    sm = block_sum(sm, sh);
    if (threadIdx.x == 0) { atomicAdd(sums + 2 * n, se); atomicAdd(sums + 2 * n + 1, sm); }
}

__global__ void __launch_bounds__(256)
img_loss_bwd_kernel(const float* __restrict__ out, const float* __restrict__ y, const float* __restrict__ black,
                    const float* __restrict__ sums, float upstream, const float* __restrict__ up_dev, const float* __restrict__ add,
                    int N, int HW, int C, float* __restrict__ d_out)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= (long long)N * HW) return;
    const int n = (int)(p / HW);
    if (up_dev) upstream *= __ldg(up_dev);
    const float k = upstream * 2.0f / ((__ldg(sums + 2 * n + 1) + 1e-8f) * (float)N);
    const float nb = 1.0f - __ldg(black + p);
    for (int ch = 0; ch < C; ++ch)
        d_out[p * C + ch] = k * (__ldg(out + p * C + ch) - __ldg(y + p * C + ch)) * nb * nb + (add ? __ldg(add + p * C + ch) : 0.0f);
}

// ---------------------------------------------------------------- feature_loss
// warp_pts: px = round_half_even(clip((sx+1)/2*W, 0, W-1))  (s_net_bundle_nobm.py:216-221)
__device__ __forceinline__ int match_pixel(float s, int size)
{
    float v = __fmul_rn(__fdiv_rn(__fadd_rn(s, 1.0f), 2.0f), (float)size);
    v = fminf(fmaxf(v, 0.0f), (float)(size - 1));
    return __float2int_rn(v);
}

__global__ void __launch_bounds__(256)
feature_loss_fwd_kernel(const float* __restrict__ matches, const float* __restrict__ mask, const float* __restrict__ img,
                        int M, int H, int W, float* __restrict__ warpped, float* __restrict__ per_sample)
{
    __shared__ float sh[8];
    const int n = blockIdx.x;
    float acc = 0.0f, cnt = 0.0f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const float4 mt = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)n * M + m);
        const int px = match_pixel(mt.x, W), py = match_pixel(mt.y, H);
        const float2 g = __ldg(reinterpret_cast<const float2*>(img) + ((size_t)n * H + py) * W + px);
        if (warpped) reinterpret_cast<float2*>(warpped)[(size_t)n * M + m] = g;
        const float mk = __ldg(mask + (size_t)n * M + m);
        acc = fmaf(fabsf(g.x - mt.z) + fabsf(g.y - mt.w), mk, acc);
        cnt += mk;
    }
    acc = block_sum(acc, sh);
    cnt = block_sum(cnt, sh);
    if (threadIdx.x == 0) per_sample[n] = acc / fmaxf(cnt, 1.0f);
}

__global__ void __launch_bounds__(256)
feature_loss_bwd_kernel(const float* __restrict__ matches, const float* __restrict__ mask, const float* __restrict__ img,
                        float upstream, const float* __restrict__ up_dev, int N, int M, int H, int W, float* __restrict__ d_img)
{
    __shared__ float sh[8];
    __shared__ float s_cnt;
    const int n = blockIdx.x;
    float cnt = 0.0f;
    for (int m = threadIdx.x; m < M; m += blockDim.x) cnt += __ldg(mask + (size_t)n * M + m);
    cnt = block_sum(cnt, sh);
    if (threadIdx.x == 0) s_cnt = fmaxf(cnt, 1.0f);
    __syncthreads();
    if (up_dev) upstream *= __ldg(up_dev);
    const float k = upstream / (s_cnt * (float)N);
    for (int m = threadIdx.x; m < M; m += blockDim.x) {
        const float mk = __ldg(mask + (size_t)n * M + m);
        if (mk == 0.0f) continue;
        const float4 mt = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)n * M + m);
        const int px = match_pixel(mt.x, W), py = match_pixel(mt.y, H);
        const size_t q = ((size_t)n * H + py) * W + px;
        const float2 g = __ldg(reinterpret_cast<const float2*>(img) + q);
        const float dx = g.x - mt.z, dy = g.y - mt.w;                  // d|t| = sign(t), sign(0) = 0
        const float sx = (dx > 0.0f) ? 1.0f : ((dx < 0.0f) ? -1.0f : 0.0f);
        const float sy = (dy > 0.0f) ? 1.0f : ((dy < 0.0f) ? -1.0f : 0.0f);
        atomicAdd(d_img + 2 * q, k * mk * sx);
        atomicAdd(d_img + 2 * q + 1, k * mk * sy);
    }
}

// ---------------------------------------------------------------- temp_loss
// CC = compile-time channel count (1, 3, 4) or 0 = run-time C.  Per-sample base pointers and 32-bit offsets inside a sample
// (H*W*C < 2^31 is checked by the C ABI): the 64-bit index arithmetic of the first version cost 74 registers (3 blocks per SM).
template <bool BWD, int CC>
__global__ void __launch_bounds__(256, BWD ? 4 : 5)
temp_loss_kernel(const float* __restrict__ out1, const float* __restrict__ black1, const float* __restrict__ out2,
                 const float* __restrict__ black2, const float* __restrict__ flow, const float* __restrict__ sums_in,
                 float upstream, const float* __restrict__ up_dev, int N, int H, int W, int Crt, float* __restrict__ sums,
                 float* __restrict__ d_out1,
                 float* __restrict__ d_out2)
{
    __shared__ float sh[8];
    const int C = CC ? CC : Crt;
    const int n = blockIdx.y, HW = H * W;
    float se = 0.0f, sm = 0.0f;
    float k = 0.0f;
    if (BWD && up_dev) upstream *= __ldg(up_dev);
    if (BWD) k = upstream * 2.0f / ((__ldg(sums_in + 2 * n + 1) + 1e-8f) * (float)N);
    const size_t sbase = (size_t)n * HW;
    const float* b2 = black2 + sbase;
    const float* o2 = out2 + sbase * C;
    const float* o1 = out1 + sbase * C;
    const float* b1 = black1 + sbase;
    const float2* fl = reinterpret_cast<const float2*>(flow) + sbase;
    float* g1p = BWD ? d_out1 + sbase * C : nullptr;
    float* d2 = BWD ? d_out2 + sbase * C : nullptr;
    // each iteration is two dependent round trips to memory (the flow sample, then the taps it points at): the next
    // iteration's flow sample is fetched one iteration ahead, and the grid is sized for ~8 iterations per thread
    const int stride = gridDim.x * blockDim.x;
    int q = blockIdx.x * blockDim.x + threadIdx.x;
    float2 f_next = (q < HW) ? __ldg(fl + q) : make_float2(0.0f, 0.0f);
    for (; q < HW; q += stride) {
        const float2 f = f_next;
        if (q + stride < HW) f_next = __ldg(fl + q + stride);
        const Taps t = make_taps(f.x, f.y, H, W);                                   // train_bundle_nobm.py:117-118
        const int ia = t.y0 * W + t.x0, ib = t.y1 * W + t.x0, ic = t.y0 * W + t.x1, id = t.y1 * W + t.x1;
        const float nb2 = blend(t, 1.0f - __ldg(b2 + ia), 1.0f - __ldg(b2 + ib), 1.0f - __ldg(b2 + ic), 1.0f - __ldg(b2 + id));
        const float m = (1.0f - __ldg(b1 + q)) * nb2;                               // :121
        sm += m;
        const float wa = t.ax * t.ay, wb = t.ax * t.by, wc = t.bx * t.ay, wd = t.bx * t.by;
        const bool scatter = BWD && taps_scatter(t);
#pragma unroll
        for (int ch = 0; ch < (CC ? CC : 1); ++ch) {
            for (int c2 = ch; c2 < C; c2 += (CC ? C : 1)) {                         // run-time C: the inner loop walks the channels
                const float v2 = blend(t, __ldg(o2 + ia * C + c2), __ldg(o2 + ib * C + c2), __ldg(o2 + ic * C + c2), __ldg(o2 + id * C + c2));
                const float e = (__ldg(o1 + q * C + c2) - v2) * m;                  // :120,:122
                if (!BWD) {
                    se = fmaf(e, e, se);
                } else {
                    const float g1 = k * e * m;
                    g1p[q * C + c2] = g1;
                    if (scatter) {
                        atomicAdd(d2 + ia * C + c2, -g1 * wa);
                        atomicAdd(d2 + ib * C + c2, -g1 * wb);
                        atomicAdd(d2 + ic * C + c2, -g1 * wc);
                        atomicAdd(d2 + id * C + c2, -g1 * wd);
                    }
                }
            }
        }
    }
    if (!BWD) {
        se = block_sum(se, sh);
        sm = block_sum(sm, sh);
        if (threadIdx.x == 0) { atomicAdd(sums + 2 * n, se); atomicAdd(sums + 2 * n + 1, sm); }
    }
}

template <bool BWD>
static void launch_temp_loss(int C, dim3 grid, cudaStream_t st, const float* out1, const float* black1, const float* out2,
                             const float* black2, const float* flow, const float* sums_in, float upstream, const float* up_dev, int N,
                             int H, int W, float* sums, float* d_out1, float* d_out2)
{
#define MGW_TL(CC) temp_loss_kernel<BWD, CC><<<grid, 256, 0, st>>>(out1, black1, out2, black2, flow, sums_in, upstream, up_dev, N, H, W, C, \
                                                                   sums, d_out1, d_out2)
    switch (C) {
    case 1: MGW_TL(1); break;
    case 3: MGW_TL(3); break;
    case 4: MGW_TL(4); break;
    default: MGW_TL(0); break;
    }
#undef MGW_TL
}

// ---------------------------------------------------------------- launchers
// blocks per sample of the grid-stride reductions: about 8 resident blocks per SM over the whole batch (cudaDevAttrMultiProcessorCount SMs), so that a
// block amortises its two block-wide reductions and its two atomics over many pixels (one pixel per thread made these
// kernels latency-bound: 58 us for 132 MB)
static int sm_count()
{
    static int n[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!n[dev & 63]) cudaDeviceGetAttribute(&n[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    return n[dev & 63] > 0 ? n[dev & 63] : 148;
}

static unsigned blocks_for(int HW, int N, int resident_per_sm = 8)
{
    const unsigned full = (unsigned)((HW + 255) / 256);
    unsigned per = (unsigned)((sm_count() * resident_per_sm + N - 1) / N);
    if (per < 1) per = 1;
    return per < full ? per : full;
}

#ifndef MGW_LOSS_ITERS
#define MGW_LOSS_ITERS 4
#endif
// the same, but never more than `iters` grid-stride iterations per thread (kernels whose iterations are dependent memory round trips)
static unsigned blocks_iters(int HW, int N, int resident_per_sm, int iters)
{
    const unsigned a = blocks_for(HW, N, resident_per_sm), b = (unsigned)((HW + 256 * iters - 1) / (256 * iters));
    return a > b ? a : b;
}

int launch_img_loss_fwd(const float* out, const float* y, const float* black, int N, int H, int W, int C, float* sums, cudaStream_t st)
{
    cudaMemsetAsync(sums, 0, sizeof(float) * 2 * N, st);
    img_loss_fwd_kernel<<<dim3(blocks_iters(H * W, N, 8, MGW_LOSS_ITERS), N), 256, 0, st>>>(out, y, black, H * W, C, sums);
    return check_launch("img_loss_fwd");
}

int launch_img_loss_bwd(const float* out, const float* y, const float* black, const float* sums, float upstream,
                        const float* up_dev, int N, int H, int W, int C, float* d_out, cudaStream_t st, const float* add)
{
    const unsigned grid = (unsigned)(((long long)N * H * W + 255) / 256);
    img_loss_bwd_kernel<<<grid, 256, 0, st>>>(out, y, black, sums, upstream, up_dev, add, N, H * W, C, d_out);
    return check_launch("img_loss_bwd");
}

int launch_feature_loss_fwd(const float* matches, const float* mask, const float* img, int N, int M, int H, int W,
                            float* warpped, float* per_sample, cudaStream_t st)
{
    feature_loss_fwd_kernel<<<N, 256, 0, st>>>(matches, mask, img, M, H, W, warpped, per_sample);
    return check_launch("feature_loss_fwd");
}

int launch_feature_loss_bwd(const float* matches, const float* mask, const float* img, float upstream, const float* up_dev, int N, int M,
                            int H, int W, float* d_img, cudaStream_t st)
{
    feature_loss_bwd_kernel<<<N, 256, 0, st>>>(matches, mask, img, upstream, up_dev, N, M, H, W, d_img);
    return check_launch("feature_loss_bwd");
}

int launch_temp_loss_fwd(const float* out1, const float* black1, const float* out2, const float* black2,
                         const float* flow, int N, int H, int W, int C, float* sums, cudaStream_t st)
{
    cudaMemsetAsync(sums, 0, sizeof(float) * 2 * N, st);
    // (the tiled kernel of mgw_loss_tile.cu also has a forward, MGW_LOSS_TILE=fwd; measured at config #2 size it is no faster
    // than this per-pixel kernel -- 68 vs 66 us at C = 3: the forward has no scatter to win on)
    if (impl_mode() != 1 && getenv("MGW_LOSS_TILE") && getenv("MGW_LOSS_TILE")[0] == 'f' && temp_loss_tile_supported(out2, black2, out2, N, H, W, C))
        return launch_temp_loss_tile_fwd(out1, black1, out2, black2, flow, N, H, W, C, sums, st);
    launch_temp_loss<false>(C, dim3(blocks_iters(H * W, N, 5, MGW_LOSS_ITERS), N), st, out1, black1, out2, black2, flow, nullptr, 0.0f, nullptr, N, H, W,
                            sums, nullptr, nullptr);
    return check_launch("temp_loss_fwd");
}

int launch_temp_loss_bwd(const float* out1, const float* black1, const float* out2, const float* black2,
                         const float* flow, const float* sums, float upstream, const float* up_dev, int N, int H, int W, int C,
                         float* d_out1, float* d_out2, cudaStream_t st)
{
    // tiles pay where the scatter dominates: 12 atomics per pixel at C = 3 (220 -> 105 us at config #2 size incl. the zero-fill);
    // at C = 1 the per-pixel kernel's 4 global atomics are cheaper than staging the boxes (60 vs 70 us)
    if (impl_mode() != 1 && C >= 3 && temp_loss_tile_supported(out2, black2, d_out2, N, H, W, C)) {
        const int rc = launch_fill_zero(d_out2, sizeof(float) * (size_t)N * H * W * C, true, st);      // W % 4 == 0: a multiple of 16 bytes
        if (rc != MGW_OK) return rc;
        return launch_temp_loss_tile_bwd(out1, black1, out2, black2, flow, sums, upstream, up_dev, N, H, W, C, d_out1, d_out2, st);
    }
    cudaMemsetAsync(d_out2, 0, sizeof(float) * (size_t)N * H * W * C, st);
    launch_temp_loss<true>(C, dim3(blocks_iters(H * W, N, 4, MGW_LOSS_ITERS), N), st, out1, black1, out2, black2, flow, sums, upstream, up_dev, N, H, W,
                           nullptr, d_out1, d_out2);
    return check_launch("temp_loss_bwd");
}

}  // namespace mgw
