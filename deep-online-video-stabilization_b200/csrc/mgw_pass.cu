// One training pass (s_net_bundle_nobm.py:266-381) as a handful of launches: the pieces around the warp kernels.
//
//   feature_acc_fwd : feature_loss / warp_pts (:215-230, :335-343) over a (match chunks x samples) grid -> per-sample
//                     (sum of masked |residual|, sum of mask) accumulated into facc[N,2]
//   feature_dh      : the backward of feature_loss taken straight to dH.  d(loss)/d(flow map) is non-zero at <= M pixels per sample
//                     and the flow map has no other consumer: instead of scattering it into a dense [N,H,W,2] tensor that the warp
//                     backward then streams through (zero-fill + scatter + 8 B/px of reads), each match adds its dH terms (the
//                     same terms the warp backward forms from d_img: SURVEY.md 8a-bwd) to its cell -> one extra partial per cell
//                     for K4
//   objective_fwd   : the scalar epilogue (:308-317, :347-359): IMG, FEAT from the per-sample sums, the four vertex sums, the
//                     configured multipliers -> total loss + the weighted parts of the reference's `ret`, one tiny launch
#include "mgw_internal.h"

namespace mgw {

namespace {

__device__ __forceinline__ float block_sum_n(float v, float* sh)
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

// warp_pts: px = round_half_even(clip((sx+1)/2*W, 0, W-1))  (s_net_bundle_nobm.py:216-221)
__device__ __forceinline__ int match_px(float s, int size)
{
    float v = __fmul_rn(__fdiv_rn(__fadd_rn(s, 1.0f), 2.0f), (float)size);
    v = fminf(fmaxf(v, 0.0f), (float)(size - 1));
    return __float2int_rn(v);
}

constexpr int kFeatThreads = 256;

__global__ void __launch_bounds__(kFeatThreads)
feature_acc_fwd_kernel(const float* __restrict__ matches, const float* __restrict__ mask, const float* __restrict__ img, int M, int H,
                       int W, float* __restrict__ warpped, float* __restrict__ facc)
{
    __shared__ float sh[kFeatThreads / 32];
    const int n = blockIdx.y, m = blockIdx.x * kFeatThreads + threadIdx.x;
    float acc = 0.0f, cnt = 0.0f;
    if (m < M) {
        const float4 mt = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)n * M + m);
        const int px = match_px(mt.x, W), py = match_px(mt.y, H);
        const float2 g = __ldg(reinterpret_cast<const float2*>(img) + ((size_t)n * H + py) * W + px);
        if (warpped) reinterpret_cast<float2*>(warpped)[(size_t)n * M + m] = g;
        const float mk = __ldg(mask + (size_t)n * M + m);
        acc = (fabsf(g.x - mt.z) + fabsf(g.y - mt.w)) * mk;
        cnt = mk;
    }
    acc = block_sum_n(acc, sh);
    cnt = block_sum_n(cnt, sh);
    if (threadIdx.x == 0) { atomicAdd(facc + 2 * n, acc); atomicAdd(facc + 2 * n + 1, cnt); }
}

// one thread per match over a (match chunks x samples) grid; the eight dH terms of a match go to its cell's slot of extra_part
// (zeroed by the launcher) by reduce-adds -- <= M * 8 of them per sample on gh*gw*8 addresses.  facc[n][1] = the forward's mask
// count of the sample.  (A deterministic form -- one block per (cell, sample) walking all M matches -- is a chain of 24 dependent
// L2 round trips: 28 us against 4.)
__global__ void __launch_bounds__(kFeatThreads)
feature_dh_kernel(const float* __restrict__ matches, const float* __restrict__ mask, const float* __restrict__ img,
                  const float* __restrict__ Hs, const float* __restrict__ facc, float upstream, const float* __restrict__ up_dev, int N,
                  int M, int H, int W, int gh, int gw, float* __restrict__ extra_part)
{
    const int n = blockIdx.y, m = blockIdx.x * kFeatThreads + threadIdx.x;
    if (m >= M) return;
    const float mk = __ldg(mask + (size_t)n * M + m);
    if (mk == 0.0f) return;
    const float4 mt = __ldg(reinterpret_cast<const float4*>(matches) + (size_t)n * M + m);
    const int px = match_px(mt.x, W), py = match_px(mt.y, H);
    const int cell = cell_of(py, H / gh, gh) * gw + cell_of(px, W / gw, gw);
    const float2 g = __ldg(reinterpret_cast<const float2*>(img) + ((size_t)n * H + py) * W + px);
    float Hc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Hc[k] = __ldg(Hs + ((size_t)n * gh * gw + cell) * 9 + k);
    if (up_dev) upstream *= __ldg(up_dev);
    const float kk = upstream / (fmaxf(__ldg(facc + 2 * n + 1), 1.0f) * (float)N);
    const float dx = g.x - mt.z, dy = g.y - mt.w;                  // d|t| = sign(t), sign(0) = 0
    const float gxn = kk * mk * ((dx > 0.0f) ? 1.0f : ((dx < 0.0f) ? -1.0f : 0.0f));
    const float gyn = kk * mk * ((dy > 0.0f) ? 1.0f : ((dy < 0.0f) ? -1.0f : 0.0f));
    if (gxn == 0.0f && gyn == 0.0f) return;
    const float xt = lin_at(px, lin_step(W)), yt = lin_at(py, lin_step(H));
    const Proj pr = project(Hc, xt, yt);
    // dH terms of one pixel: xn = xs/zs, yn = ys/zs
    const float rz = __frcp_rn(pr.zs);
    const float dxs = gxn * rz, dys = gyn * rz, dzs = -(gxn * pr.xn + gyn * pr.yn) * rz;
    float* dst = extra_part + ((size_t)n * gh * gw + cell) * 8;
    atomicAdd(dst + 0, dxs * xt); atomicAdd(dst + 1, dxs * yt); atomicAdd(dst + 2, dxs);
    atomicAdd(dst + 3, dys * xt); atomicAdd(dst + 4, dys * yt); atomicAdd(dst + 5, dys);
    atomicAdd(dst + 6, dzs * xt); atomicAdd(dst + 7, dzs * yt);
}

constexpr int kObjThreads = 128;

__global__ void __launch_bounds__(kObjThreads)
objective_fwd_kernel(const float* __restrict__ img_sums, const float* __restrict__ facc, const float* __restrict__ vsums,
                     const float* __restrict__ regu_dev, int N, const PassCoef c, float* __restrict__ result)
{
    __shared__ float sh[kObjThreads / 32];
    float a = 0.0f, b = 0.0f;
    for (int n = threadIdx.x; n < N; n += kObjThreads) {
        a += __ldg(img_sums + 2 * n) / (__ldg(img_sums + 2 * n + 1) + 1e-8f);          // :347-351
        b += __ldg(facc + 2 * n) / fmaxf(__ldg(facc + 2 * n + 1), 1.0f);               // :338-342
    }
    a = block_sum_n(a, sh);
    b = block_sum_n(b, sh);
    if (threadIdx.x == 0) {
        const float IMG = a * c.inv_batch, FEAT = b * c.inv_batch, REGU = regu_dev ? __ldg(regu_dev) : 0.0f;
        const float id = c.v[0] * __ldg(vsums), black = c.v[1] * __ldg(vsums + 1), dist = c.v[2] * __ldg(vsums + 2);
        const float cons = c.v[3] * __ldg(vsums + 3), im = c.img * IMG, ft = c.feat * FEAT, rg = c.regu * REGU;
        result[0] = id + c.gate * (black + dist + cons + im + ft + rg);
        // the reference's ret[...] parts (s_net_bundle_nobm.py:361-375), in the order of losses.PASS_PARTS
        result[1] = id * c.theta_share; result[2] = id * c.grid_theta_share; result[3] = black; result[4] = dist;
        result[5] = cons; result[6] = ft; result[7] = im; result[8] = rg;
    }
}

// scale * sum_n s0_n / (s1_n + 1e-8)   (clamp = false: img_loss :347-351, temp_loss train:121-125)
// scale * sum_n s0_n / max(s1_n, 1)      (clamp = true : feature_loss :338-342)
__global__ void __launch_bounds__(kObjThreads)
ratio_sum_kernel(const float* __restrict__ sums, int N, bool clamp, float scale, float* __restrict__ out)
{
    __shared__ float sh[kObjThreads / 32];
    float a = 0.0f;
    for (int n = threadIdx.x; n < N; n += kObjThreads) {
        const float s0 = __ldg(sums + 2 * n), s1 = __ldg(sums + 2 * n + 1);
        a += s0 / (clamp ? fmaxf(s1, 1.0f) : s1 + 1e-8f);
    }
    a = block_sum_n(a, sh);
    if (threadIdx.x == 0) out[0] = a * scale;
}

}  // namespace

int launch_ratio_sum(const float* sums, int N, bool clamp, float scale, float* out, cudaStream_t st)
{
    ratio_sum_kernel<<<1, kObjThreads, 0, st>>>(sums, N, clamp, scale, out);
    return check_launch("ratio_sum");
}

int launch_feature_acc_fwd(const float* matches, const float* mask, const float* img, int N, int M, int H, int W, float* warpped,
                           float* facc, cudaStream_t st)
{
    if (M <= 0) return MGW_OK;
    feature_acc_fwd_kernel<<<dim3((M + kFeatThreads - 1) / kFeatThreads, N), kFeatThreads, 0, st>>>(matches, mask, img, M, H, W, warpped, facc);
    return check_launch("feature_acc_fwd");
}

int launch_feature_dh(const float* matches, const float* mask, const float* img, const float* Hs, const float* facc, float upstream,
                      const float* up_dev, int N, int M, int H, int W, int gh, int gw, float* extra_part, cudaStream_t st)
{
    if (cudaMemsetAsync(extra_part, 0, sizeof(float) * (size_t)N * gh * gw * 8, st) != cudaSuccess)
        return set_error(MGW_ERR_CUDA, "memset dH partial: %s", cudaGetErrorString(cudaGetLastError()));
    feature_dh_kernel<<<dim3((M + kFeatThreads - 1) / kFeatThreads, N), kFeatThreads, 0, st>>>(matches, mask, img, Hs, facc, upstream, up_dev,
                                                                                               N, M, H, W, gh, gw, extra_part);
    return check_launch("feature_dh");
}

int launch_objective_fwd(const float* img_sums, const float* facc, const float* vsums, const float* regu_dev, int N, const PassCoef& c,
                         float* result, cudaStream_t st)
{
    objective_fwd_kernel<<<1, kObjThreads, 0, st>>>(img_sums, facc, vsums, regu_dev, N, c, result);
    return check_launch("objective_fwd");
}

}  // namespace mgw
