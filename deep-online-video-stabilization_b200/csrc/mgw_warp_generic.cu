// K2/K3, generic variant: one thread per output pixel, taps gathered straight from global memory (L1/L2),
// dU scattered with global fp32 RED atomics.  Handles ANY shape (odd widths, any C, any mesh, out_size !=
// input size) and is the fallback when the TMA-staged kernels (mgw_warp_tma.cu) cannot be used.
//
//   forward : _transform3 body + _interpolate, spatial_transformer3.py:227-301 / :62-123
//             (and spatial_transformer.py:143-193 with normalize=true, gh=gw=1)
//   backward: closed form of SURVEY.md 8a-bwd
#include "mgw_internal.h"

namespace mgw {

constexpr int kMaxC = 16;

struct PixelCtx {
    int n, r, c, cell;
    float xt, yt;
    float Hc[9];
};

__device__ __forceinline__ bool pixel_ctx(const WarpShape& s, const float* __restrict__ Hs, bool normalize, PixelCtx& px)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long P = (long long)s.N * s.OH * s.OW;
    if (p >= P) return false;
    px.c = (int)(p % s.OW);
    const long long q = p / s.OW;
    px.r = (int)(q % s.OH);
    px.n = (int)(q / s.OH);
    const int ci = cell_of(px.r, s.OH / s.gh, s.gh), cj = cell_of(px.c, s.OW / s.gw, s.gw);   // :227-243
    px.cell = ci * s.gw + cj;
    const float* h = Hs + ((size_t)px.n * s.gh * s.gw + px.cell) * 9;
#pragma unroll
    for (int k = 0; k < 9; ++k) px.Hc[k] = __ldg(h + k);
    if (normalize) {                                                         // spatial_transformer.py:151-153
        const float d = px.Hc[8];
#pragma unroll
        for (int k = 0; k < 9; ++k) px.Hc[k] = __fdiv_rn(px.Hc[k], d);
    }
    px.xt = lin_at(px.c, lin_step(s.OW));
    px.yt = lin_at(px.r, lin_step(s.OH));
    return true;
}

template <int CT>
__global__ void __launch_bounds__(256)
warp_fwd_generic_kernel(const float* __restrict__ U, const float* __restrict__ Hs, WarpShape s, bool normalize,
                        float* __restrict__ out, float* __restrict__ black, float* __restrict__ img,
                        int32_t* __restrict__ cell_idx)
{
    PixelCtx px;
    if (!pixel_ctx(s, Hs, normalize, px)) return;
    const Proj pr = project(px.Hc, px.xt, px.yt);
    const size_t p = ((size_t)px.n * s.OH + px.r) * s.OW + px.c;
    if (img) reinterpret_cast<float2*>(img)[p] = make_float2(pr.xn, pr.yn);      // x_map,y_map :271-272,:278
    if (black) black[p] = black_of(pr.xn, pr.yn);                                // :284-286
    if (cell_idx) cell_idx[p] = px.cell;
    if (!out) return;
    const int C = CT > 0 ? CT : s.C;
    const Taps t = make_taps(pr.xn, pr.yn, s.H, s.W);
    const float* Un = U + (size_t)px.n * s.H * s.W * C;
    const float* pa = Un + ((size_t)t.y0 * s.W + t.x0) * C;
    const float* pb = Un + ((size_t)t.y1 * s.W + t.x0) * C;
    const float* pc = Un + ((size_t)t.y0 * s.W + t.x1) * C;
    const float* pd = Un + ((size_t)t.y1 * s.W + t.x1) * C;
    float* o = out + p * C;
#pragma unroll
    for (int ch = 0; ch < (CT > 0 ? CT : kMaxC); ++ch) {
        if (CT == 0 && ch >= C) break;
        o[ch] = blend(t, __ldg(pa + ch), __ldg(pb + ch), __ldg(pc + ch), __ldg(pd + ch));
    }
}

template <int CT>
__global__ void __launch_bounds__(256)
warp_bwd_generic_kernel(const float* __restrict__ U, const float* __restrict__ Hs, const float* __restrict__ d_out,
                        const float* __restrict__ d_img, WarpShape s, bool normalize, float* __restrict__ dU,
                        float* __restrict__ dHs)
{
    PixelCtx px;
    const bool live = pixel_ctx(s, Hs, normalize, px);
    float dh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dh[k] = 0.0f;
    int key = -1;
    if (live) {
        const int C = CT > 0 ? CT : s.C;
        const Proj pr = project(px.Hc, px.xt, px.yt);
        const size_t p = ((size_t)px.n * s.OH + px.r) * s.OW + px.c;
        const Taps t = make_taps(pr.xn, pr.yn, s.H, s.W);
        const size_t ia = ((size_t)t.y0 * s.W + t.x0) * C, ib = ((size_t)t.y1 * s.W + t.x0) * C;
        const size_t ic = ((size_t)t.y0 * s.W + t.x1) * C, id = ((size_t)t.y1 * s.W + t.x1) * C;
        const float* Un = U + (size_t)px.n * s.H * s.W * C;
        float* dUn = (dU && taps_scatter(t)) ? dU + (size_t)px.n * s.H * s.W * C : nullptr;
        const float wa = t.ax * t.ay, wb = t.ax * t.by, wc = t.bx * t.ay, wd = t.bx * t.by;
        float gx = 0.0f, gy = 0.0f;
#pragma unroll
        for (int ch = 0; ch < (CT > 0 ? CT : kMaxC); ++ch) {
            if (CT == 0 && ch >= C) break;
            const float g = __ldg(d_out + p * C + ch);
            const float Ia = __ldg(Un + ia + ch), Ib = __ldg(Un + ib + ch), Ic = __ldg(Un + ic + ch), Id = __ldg(Un + id + ch);
            gx = fmaf(g, fmaf(Ic - Ia, t.ay, (Id - Ib) * t.by), gx);
            gy = fmaf(g, fmaf(Ib - Ia, t.ax, (Id - Ic) * t.bx), gy);
            if (dUn) {
                atomicAdd(dUn + ia + ch, wa * g);
                atomicAdd(dUn + ib + ch, wb * g);
                atomicAdd(dUn + ic + ch, wc * g);
                atomicAdd(dUn + id + ch, wd * g);
            }
        }
        float gxn = gx * (0.5f * (float)s.W), gyn = gy * (0.5f * (float)s.H);
        if (d_img) {
            const float2 di = __ldg(reinterpret_cast<const float2*>(d_img) + p);
            gxn += di.x; gyn += di.y;
        }
        const float rz = 1.0f / pr.zs;
        const float dxs = gxn * rz, dys = gyn * rz;
        const float dzs = -(gxn * pr.xn + gyn * pr.yn) * rz;
        dh[0] = dxs * px.xt; dh[1] = dxs * px.yt; dh[2] = dxs;
        dh[3] = dys * px.xt; dh[4] = dys * px.yt; dh[5] = dys;
        dh[6] = dzs * px.xt; dh[7] = dzs * px.yt;
        key = px.n * s.gh * s.gw + px.cell;
    }
    // reduce the 8 dH terms: one atomic per warp when the warp sits in a single (sample, cell)
    const int key0 = __shfl_sync(0xffffffffu, key, 0);
    if (__all_sync(0xffffffffu, key == key0)) {
        if (key0 < 0) return;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const float v = warp_sum(dh[k]);
            if ((threadIdx.x & 31) == 0) atomicAdd(dHs + (size_t)key0 * 9 + k, v);
        }
    } else if (live) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(dHs + (size_t)key * 9 + k, dh[k]);
    }
}

// d theta of theta/theta[8] (spatial_transformer.py:151-153): dth_k = dHn_k/th8 (k<8), dth_8 = -sum_k dHn_k th_k/th8^2
__global__ void homography_finish_bwd_kernel(const float* __restrict__ theta, const float* __restrict__ dHn, int N,
                                             float* __restrict__ dtheta)
{
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= N) return;
    const double t8 = theta[n * 9 + 8];
    double acc = 0;
    for (int k = 0; k < 8; ++k) {
        const double g = dHn[n * 9 + k];
        dtheta[n * 9 + k] = (float)(g / t8);
        acc += g * (double)theta[n * 9 + k];
    }
    dtheta[n * 9 + 8] = (float)(-acc / (t8 * t8));
}

template <typename F>
static int dispatch_c(int C, F&& f)
{
    switch (C) {
        case 1: return f(std::integral_constant<int, 1>());
        case 2: return f(std::integral_constant<int, 2>());
        case 3: return f(std::integral_constant<int, 3>());
        case 4: return f(std::integral_constant<int, 4>());
        default: return f(std::integral_constant<int, 0>());
    }
}

int launch_warp_fwd_generic(const float* U, const float* Hs, const WarpShape& s, bool normalize, float* out,
                            float* black, float* img, int32_t* cell_idx, cudaStream_t st)
{
    if (s.C > kMaxC) return set_error(MGW_ERR_UNSUPPORTED, "warp_fwd: C=%d > %d", s.C, kMaxC);
    const long long P = (long long)s.N * s.OH * s.OW;
    const unsigned grid = (unsigned)((P + 255) / 256);
    return dispatch_c(s.C, [&](auto ct) {
        warp_fwd_generic_kernel<decltype(ct)::value><<<grid, 256, 0, st>>>(U, Hs, s, normalize, out, black, img, cell_idx);
        return check_launch("warp_fwd_generic");
    });
}

int launch_warp_bwd_generic(const float* U, const float* Hs, const float* d_out, const float* d_img,
                            const WarpShape& s, bool normalize, float* dU, float* dHs, cudaStream_t st)
{
    if (s.C > kMaxC) return set_error(MGW_ERR_UNSUPPORTED, "warp_bwd: C=%d > %d", s.C, kMaxC);
    const long long P = (long long)s.N * s.OH * s.OW;
    const unsigned grid = (unsigned)((P + 255) / 256);
    return dispatch_c(s.C, [&](auto ct) {
        warp_bwd_generic_kernel<decltype(ct)::value><<<grid, 256, 0, st>>>(U, Hs, d_out, d_img, s, normalize, dU, dHs);
        return check_launch("warp_bwd_generic");
    });
}

int launch_homography_finish_bwd(const float* theta, const float* dHn, int N, float* dtheta, cudaStream_t st)
{
    homography_finish_bwd_kernel<<<(N + 127) / 128, 128, 0, st>>>(theta, dHn, N, dtheta);
    return check_launch("homography_finish_bwd");
}

}  // namespace mgw
