// interpolate(im, x, y, out_size): bilinear gather at caller-supplied normalised coordinates
// (spatial_transformer.py:200-281; identical core to spatial_transformer3.py:62-123), forward and backward.
// One thread per output pixel; coordinates are read (not computed), the flow field is arbitrary so there is no
// source-tile locality to stage: taps go through L1/L2, d_im is scattered with fp32 RED atomics.
#include "mgw_internal.h"

namespace mgw {

constexpr int kMaxC = 16;

template <int CT>
__global__ void __launch_bounds__(256)
interp_fwd_kernel(const float* __restrict__ im, const float* __restrict__ x, const float* __restrict__ y, int N,
                  int IH, int IW, int Crt, int OH, int OW, float* __restrict__ out)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)OH * OW;
    if (p >= per * N) return;
    const int C = CT > 0 ? CT : Crt;
    const int n = (int)(p / per);
    const Taps t = make_taps(__ldg(x + p), __ldg(y + p), IH, IW);
    const float* imn = im + (size_t)n * IH * IW * C;
    const float* pa = imn + ((size_t)t.y0 * IW + t.x0) * C;
    const float* pb = imn + ((size_t)t.y1 * IW + t.x0) * C;
    const float* pc = imn + ((size_t)t.y0 * IW + t.x1) * C;
    const float* pd = imn + ((size_t)t.y1 * IW + t.x1) * C;
#pragma unroll
    for (int ch = 0; ch < (CT > 0 ? CT : kMaxC); ++ch) {
        if (CT == 0 && ch >= C) break;
        out[p * C + ch] = blend(t, __ldg(pa + ch), __ldg(pb + ch), __ldg(pc + ch), __ldg(pd + ch));
    }
}

template <int CT>
__global__ void __launch_bounds__(256)
interp_bwd_kernel(const float* __restrict__ im, const float* __restrict__ x, const float* __restrict__ y,
                  const float* __restrict__ d_out, int N, int IH, int IW, int Crt, int OH, int OW,
                  float* __restrict__ d_im, float* __restrict__ dx, float* __restrict__ dy)
{
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per = (long long)OH * OW;
    if (p >= per * N) return;
    const int C = CT > 0 ? CT : Crt;
    const int n = (int)(p / per);
    const Taps t = make_taps(__ldg(x + p), __ldg(y + p), IH, IW);
    const size_t ia = ((size_t)t.y0 * IW + t.x0) * C, ib = ((size_t)t.y1 * IW + t.x0) * C;
    const size_t ic = ((size_t)t.y0 * IW + t.x1) * C, id = ((size_t)t.y1 * IW + t.x1) * C;
    const float* imn = im + (size_t)n * IH * IW * C;
    float* dn = (d_im && taps_scatter(t)) ? d_im + (size_t)n * IH * IW * C : nullptr;
    const float wa = t.ax * t.ay, wb = t.ax * t.by, wc = t.bx * t.ay, wd = t.bx * t.by;
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int ch = 0; ch < (CT > 0 ? CT : kMaxC); ++ch) {
        if (CT == 0 && ch >= C) break;
        const float g = __ldg(d_out + p * C + ch);
        if (dx || dy) {
            const float Ia = __ldg(imn + ia + ch), Ib = __ldg(imn + ib + ch), Ic = __ldg(imn + ic + ch), Id = __ldg(imn + id + ch);
            gx = fmaf(g, fmaf(Ic - Ia, t.ay, (Id - Ib) * t.by), gx);
            gy = fmaf(g, fmaf(Ib - Ia, t.ax, (Id - Ic) * t.bx), gy);
        }
        if (dn) {
            atomicAdd(dn + ia + ch, wa * g);
            atomicAdd(dn + ib + ch, wb * g);
            atomicAdd(dn + ic + ch, wc * g);
            atomicAdd(dn + id + ch, wd * g);
        }
    }
    if (dx) dx[p] = gx * (0.5f * (float)IW);        // x = (xn+1)*W/2  (:228)
    if (dy) dy[p] = gy * (0.5f * (float)IH);
}

int launch_interp_fwd(const float* im, const float* x, const float* y, int N, int IH, int IW, int C, int OH, int OW,
                      float* out, cudaStream_t st)
{
    if (C > kMaxC) return set_error(MGW_ERR_UNSUPPORTED, "interp_fwd: C=%d > %d", C, kMaxC);
    const unsigned grid = (unsigned)(((long long)N * OH * OW + 255) / 256);
    switch (C) {
        case 1: interp_fwd_kernel<1><<<grid, 256, 0, st>>>(im, x, y, N, IH, IW, C, OH, OW, out); break;
        case 3: interp_fwd_kernel<3><<<grid, 256, 0, st>>>(im, x, y, N, IH, IW, C, OH, OW, out); break;
        default: interp_fwd_kernel<0><<<grid, 256, 0, st>>>(im, x, y, N, IH, IW, C, OH, OW, out); break;
    }
    return check_launch("interp_fwd");
}

int launch_interp_bwd(const float* im, const float* x, const float* y, const float* d_out, int N, int IH, int IW,
                      int C, int OH, int OW, float* d_im, float* dx, float* dy, cudaStream_t st)
{
    if (C > kMaxC) return set_error(MGW_ERR_UNSUPPORTED, "interp_bwd: C=%d > %d", C, kMaxC);
    const unsigned grid = (unsigned)(((long long)N * OH * OW + 255) / 256);
    switch (C) {
        case 1: interp_bwd_kernel<1><<<grid, 256, 0, st>>>(im, x, y, d_out, N, IH, IW, C, OH, OW, d_im, dx, dy); break;
        case 3: interp_bwd_kernel<3><<<grid, 256, 0, st>>>(im, x, y, d_out, N, IH, IW, C, OH, OW, d_im, dx, dy); break;
        default: interp_bwd_kernel<0><<<grid, 256, 0, st>>>(im, x, y, d_out, N, IH, IW, C, OH, OW, d_im, dx, dy); break;
    }
    return check_launch("interp_bwd");
}

}  // namespace mgw
