// Deploy-side colour-frame warp: warpRevBundle2 of deploy_bundle.py:136-146.
//   x_map, y_map (the operator's outputs, [N,H,W,2] interleaved)  --cv2.resize /4-->  small maps  --cv2.resize x4-->
//   smoothed maps  --((m+1)/2*size)-->  pixel coordinates  --cv2.remap(INTER_LINEAR, constant border 0)-->  uint8 frame
// Two launches: K5a writes the /4 maps (a 74 KB scratch at 288x512, L2 resident), K5b upsamples them per output pixel,
// converts to OpenCV's 1/32-pixel fixed point and blends the uint8 taps with OpenCV's 15-bit integer weights.  The
// arithmetic restates OpenCV's published algorithm (resize.cpp resizeGeneric_/HResizeLinear/VResizeLinear, plain C++
// path; imgwarp.cpp remapBilinear) operation by operation: deploy_ref.py (test infrastructure) is the numpy statement of the same thing
// and is pinned bit-exact against cv2 (tests/golden/deploy_remap.npz).  HBM-bound byte work: no tensor cores.
#include "mgw_internal.h"

namespace mgw {

namespace {

struct Lin { int i0, i1; float a0, a1; };

// resize.cpp: inv_scale = dsize / (double)ssize, scale = 1. / inv_scale, fx = (float)((d + 0.5) * scale - 0.5)
__device__ __forceinline__ float src_pos(int d, int n_out, int n_in)
{
    const double scale = 1.0 / ((double)n_out / (double)n_in);
    return (float)(((double)d + 0.5) * scale - 0.5);
}

// horizontal taps: a tap outside the row is folded onto the border with weight 0
__device__ __forceinline__ Lin hcoef(int d, int n_out, int n_in)
{
    float fx = src_pos(d, n_out, n_in);
    int sx = (int)floorf(fx);
    fx = __fsub_rn(fx, (float)sx);
    if (sx < 0) { sx = 0; fx = 0.0f; }
    if (sx >= n_in - 1) { sx = n_in - 1; fx = 0.0f; }
    Lin l;
    l.i0 = sx; l.i1 = min(sx + 1, n_in - 1); l.a0 = __fsub_rn(1.0f, fx); l.a1 = fx;
    return l;
}

// vertical taps: row indices clamped, weights kept
__device__ __forceinline__ Lin vcoef(int d, int n_out, int n_in)
{
    float fy = src_pos(d, n_out, n_in);
    const int sy = (int)floorf(fy);
    fy = __fsub_rn(fy, (float)sy);
    Lin l;
    l.i0 = min(max(sy, 0), n_in - 1); l.i1 = min(max(sy + 1, 0), n_in - 1); l.a0 = __fsub_rn(1.0f, fy); l.a1 = fy;
    return l;
}

// one bilinear resize sample of an interleaved (x,y) map: rows first horizontally, then vertically, fp32 mul / add
__device__ __forceinline__ float2 resize_sample(const float2* __restrict__ s, int sw, const Lin& h, const Lin& v)
{
    const float2 p00 = __ldg(s + (size_t)v.i0 * sw + h.i0), p01 = __ldg(s + (size_t)v.i0 * sw + h.i1);
    const float2 p10 = __ldg(s + (size_t)v.i1 * sw + h.i0), p11 = __ldg(s + (size_t)v.i1 * sw + h.i1);
    const float r0x = __fadd_rn(__fmul_rn(p00.x, h.a0), __fmul_rn(p01.x, h.a1));
    const float r1x = __fadd_rn(__fmul_rn(p10.x, h.a0), __fmul_rn(p11.x, h.a1));
    const float r0y = __fadd_rn(__fmul_rn(p00.y, h.a0), __fmul_rn(p01.y, h.a1));
    const float r1y = __fadd_rn(__fmul_rn(p10.y, h.a0), __fmul_rn(p11.y, h.a1));
    return make_float2(__fadd_rn(__fmul_rn(r0x, v.a0), __fmul_rn(r1x, v.a1)), __fadd_rn(__fmul_rn(r0y, v.a0), __fmul_rn(r1y, v.a1)));
}

__global__ void __launch_bounds__(256)
maps_down_kernel(const float2* __restrict__ xy, int N, int H, int W, int h4, int w4, float2* __restrict__ small)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * h4 * w4) return;
    const int c = t % w4, r = (t / w4) % h4, n = t / (w4 * h4);
    const Lin h = hcoef(c, w4, W), v = vcoef(r, h4, H);
    small[t] = resize_sample(xy + (size_t)n * H * W, W, h, v);
}

// cvRound(v) as x86 does it (cvtss2si: round half to even; out of range / NaN -> INT_MIN)
__device__ __forceinline__ int cv_round(float v)
{
    return (fabsf(v) < 2147483648.0f) ? __float2int_rn(v) : (int)0x80000000;
}

// remapBilinear on 1/32-pixel fixed-point source coordinates: taps clamped to int16, 15-bit integer weights summing to 2^15,
// constant border 0 (imgwarp.cpp); shared by cv2.remap (float maps) and cv2.warpPerspective (its own fixed-point maps)
template <int C>
__device__ __forceinline__ void bilinear_fixed_u8(const uint8_t* __restrict__ base, int H, int W, int sx, int sy, uint8_t* __restrict__ o)
{
    const int ix = min(max(sx >> 5, -32768), 32767), iy = min(max(sy >> 5, -32768), 32767);
    const int fx = sx & 31, fy = sy & 31;
    const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
    const bool x0 = (unsigned)ix < (unsigned)W, x1 = (unsigned)(ix + 1) < (unsigned)W;
    const bool y0 = (unsigned)iy < (unsigned)H, y1 = (unsigned)(iy + 1) < (unsigned)H;
    const uint8_t* p00 = base + ((long long)iy * W + ix) * C;
    int acc[C];
#pragma unroll
    for (int ch = 0; ch < C; ++ch) acc[ch] = 1 << 14;
    if (y0 && x0) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] += w00 * (int)__ldg(p00 + ch);
    }
    if (y0 && x1) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] += w01 * (int)__ldg(p00 + C + ch);
    }
    if (y1 && x0) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] += w10 * (int)__ldg(p00 + (long long)W * C + ch);
    }
    if (y1 && x1) {
#pragma unroll
        for (int ch = 0; ch < C; ++ch) acc[ch] += w11 * (int)__ldg(p00 + (long long)W * C + C + ch);
    }
#pragma unroll
    for (int ch = 0; ch < C; ++ch) o[ch] = (uint8_t)min(max(acc[ch] >> 15, 0), 255);
}

// warpRevBundle (deploy_bundle.py:148-173): output cell (i, j) = that region of cv2.warpPerspective(img, Hs_cvt[i][j],
// WARP_INVERSE_MAP | INTER_LINEAR).  OpenCV walks the destination in blocks of bw columns and evaluates, in double,
//   X0 = M0*bx + M1*y + M2 (left to right), W = W0 + M6*x1, W = W ? 32/W : 0, fX = clamp((X0 + M0*x1)*W), X = cvRound(fX)
// with bx the block's first column and x1 = x - bx; restated operation by operation (no contraction: __dmul_rn / __dadd_rn).
__device__ __forceinline__ int persp_fixed(double a, double b, double c, double bx, double y, double x1, double Wq)
{
    const double X0 = __dadd_rn(__dadd_rn(__dmul_rn(a, bx), __dmul_rn(b, y)), c);
    double f = __dmul_rn(__dadd_rn(X0, __dmul_rn(a, x1)), Wq);
    f = (f < 2147483647.0) ? f : 2147483647.0;                    // std::min((double)INT_MAX, f)
    f = (-2147483648.0 < f) ? f : -2147483648.0;                  // std::max((double)INT_MIN, f)
    return __double2int_rn(f);                                    // cvRound: round half to even
}

template <int C>
__global__ void __launch_bounds__(256)
warp_rev_bundle_kernel(const uint8_t* __restrict__ img, const double* __restrict__ Hc, int N, int H, int W, int gh, int gw, int bw,
                       uint8_t* __restrict__ dst)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * H * W) return;
    const int x = t % W, y = (t / W) % H, n = t / (W * H);
    const int ci = min(y / (H / gh), gh - 1), cj = min(x / (W / gw), gw - 1);       // the last cell takes the remainder (:162-165)
    const double* M = Hc + ((size_t)(n * gh + ci) * gw + cj) * 9;
    const int bxi = (x / bw) * bw;
    const double bx = (double)bxi, x1 = (double)(x - bxi), yd = (double)y;
    const double W0 = __dadd_rn(__dadd_rn(__dmul_rn(__ldg(M + 6), bx), __dmul_rn(__ldg(M + 7), yd)), __ldg(M + 8));
    double Wq = __dadd_rn(W0, __dmul_rn(__ldg(M + 6), x1));
    Wq = (Wq != 0.0) ? __ddiv_rn(32.0, Wq) : 0.0;
    const int sx = persp_fixed(__ldg(M + 0), __ldg(M + 1), __ldg(M + 2), bx, yd, x1, Wq);
    const int sy = persp_fixed(__ldg(M + 3), __ldg(M + 4), __ldg(M + 5), bx, yd, x1, Wq);
    bilinear_fixed_u8<C>(img + (size_t)n * H * W * C, H, W, sx, sy, dst + (size_t)t * C);
}

template <int C>
__global__ void __launch_bounds__(256)
remap_up_kernel(const uint8_t* __restrict__ img, const float2* __restrict__ small, int N, int H, int W, int h4, int w4,
                uint8_t* __restrict__ dst)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * H * W) return;
    const int c = t % W, r = (t / W) % H, n = t / (W * H);
    const Lin h = hcoef(c, W, w4), v = vcoef(r, H, h4);
    const float2 m = resize_sample(small + (size_t)n * h4 * w4, w4, h, v);
    // (map + 1) / 2 * size in fp32 (deploy_bundle.py:142-143)
    const float xp = __fmul_rn(__fmul_rn(__fadd_rn(m.x, 1.0f), 0.5f), (float)W);
    const float yp = __fmul_rn(__fmul_rn(__fadd_rn(m.y, 1.0f), 0.5f), (float)H);
    const int sx = cv_round(__fmul_rn(xp, 32.0f)), sy = cv_round(__fmul_rn(yp, 32.0f));
    bilinear_fixed_u8<C>(img + (size_t)n * H * W * C, H, W, sx, sy, dst + (size_t)t * C);
}

}  // namespace

size_t remap_bundle_workspace_bytes(int N, int H, int W) { return sizeof(float) * 2 * (size_t)N * (H / 4) * (W / 4); }

int launch_remap_bundle_u8(const uint8_t* img, const float* xy, int N, int H, int W, int C, uint8_t* dst, void* workspace,
                           cudaStream_t st)
{
    const int h4 = H / 4, w4 = W / 4;       // int(height / rate), int(width / rate)
    float2* small = reinterpret_cast<float2*>(workspace);
    const int t1 = N * h4 * w4, t2 = N * H * W;
    maps_down_kernel<<<(t1 + 255) / 256, 256, 0, st>>>(reinterpret_cast<const float2*>(xy), N, H, W, h4, w4, small);
    int rc = check_launch("maps_down");
    if (rc != MGW_OK) return rc;
    const int grid = (t2 + 255) / 256;
    switch (C) {
    case 1: remap_up_kernel<1><<<grid, 256, 0, st>>>(img, small, N, H, W, h4, w4, dst); break;
    case 3: remap_up_kernel<3><<<grid, 256, 0, st>>>(img, small, N, H, W, h4, w4, dst); break;
    case 4: remap_up_kernel<4><<<grid, 256, 0, st>>>(img, small, N, H, W, h4, w4, dst); break;
    default: return set_error(MGW_ERR_UNSUPPORTED, "remap_bundle_u8: C must be 1, 3 or 4 (got %d)", C);
    }
    return check_launch("remap_up");
}

// ---------------------------------------------------------------------------------------------------------------------
// cvt_img2train (config.py:6-21): BGR uint8 frame -> cv2 BGR2GRAY -> Pillow resize(BILINEAR) [-> crop] -> v*(1/255) - 0.5.
// The host binding computes Pillow's 22-bit fixed-point coefficient tables (double arithmetic, Resample.c) once per shape
// and slices them for the crop; the two passes below are Pillow's: horizontal first into a uint8 image, then vertical.
namespace {

__global__ void __launch_bounds__(256)
gray_hpass_kernel(const uint8_t* __restrict__ bgr, int H, int W, const int32_t* __restrict__ kx, const int32_t* __restrict__ x0,
                  const int32_t* __restrict__ xn, int ks, int out_w, uint8_t* __restrict__ tmp)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= H * out_w) return;
    const int xx = t % out_w, row = t / out_w;
    const uint8_t* p = bgr + ((size_t)row * W + __ldg(x0 + xx)) * 3;
    const int n = __ldg(xn + xx);
    int acc = 1 << 21;
    for (int k = 0; k < n; ++k, p += 3) {
        // RGB2Gray<uchar>: (B*3735 + G*19235 + R*9798 + 2^14) >> 15
        const int g = ((int)__ldg(p) * 3735 + (int)__ldg(p + 1) * 19235 + (int)__ldg(p + 2) * 9798 + (1 << 14)) >> 15;
        acc += g * __ldg(kx + (size_t)xx * ks + k);
    }
    tmp[t] = (uint8_t)min(max(acc >> 22, 0), 255);
}

__global__ void __launch_bounds__(256)
vpass_norm_kernel(const uint8_t* __restrict__ tmp, int out_w, const int32_t* __restrict__ ky, const int32_t* __restrict__ y0,
                  const int32_t* __restrict__ yn, int ks, int out_h, float* __restrict__ out)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= out_h * out_w) return;
    const int xx = t % out_w, yy = t / out_w;
    const uint8_t* p = tmp + (size_t)__ldg(y0 + yy) * out_w + xx;
    const int n = __ldg(yn + yy);
    int acc = 1 << 21;
    for (int k = 0; k < n; ++k, p += out_w) acc += (int)__ldg(p) * __ldg(ky + (size_t)yy * ks + k);
    const int v = min(max(acc >> 22, 0), 255);
    // img * (1. / 255) - 0.5 in double (numpy), then the float32 cast of the network's placeholder
    out[t] = (float)__dadd_rn(__dmul_rn((double)v, 1.0 / 255), -0.5);
}

}  // namespace

int launch_cvt_img2train_u8(const uint8_t* bgr, int H, int W, const int32_t* kx, const int32_t* x0, const int32_t* xn, int ksx,
                            const int32_t* ky, const int32_t* y0, const int32_t* yn, int ksy, int out_h, int out_w, uint8_t* tmp,
                            float* out, cudaStream_t st)
{
    gray_hpass_kernel<<<(H * out_w + 255) / 256, 256, 0, st>>>(bgr, H, W, kx, x0, xn, ksx, out_w, tmp);
    if (int rc = check_launch("gray_hpass")) return rc;
    vpass_norm_kernel<<<(out_h * out_w + 255) / 256, 256, 0, st>>>(tmp, out_w, ky, y0, yn, ksy, out_h, out);
    return check_launch("vpass_norm");
}

// ---------------------------------------------------------------------------------------------------------------------
// cv2.resize(frame, (width, height)) of the uint8 colour frame (deploy_bundle.py:301; INTER_LINEAR): OpenCV's 8-bit path --
// 11-bit fixed-point weights (tables from the host binding: x0, x1, a0, a1 per output column / row), int32 horizontal sums,
// ((b0*(S0>>4))>>16) + ((b1*(S1>>4))>>16) + 2) >> 2 vertically; an exact 2x2 decimation is OpenCV's INTER_AREA shortcut.
namespace {

template <int C>
__global__ void __launch_bounds__(256)
resize_u8_kernel(const uint8_t* __restrict__ img, int H, int W, const int4* __restrict__ xtab, const int4* __restrict__ ytab,
                 int out_h, int out_w, int area2, uint8_t* __restrict__ dst)
{
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= out_h * out_w) return;
    const int x = t % out_w, y = t / out_w;
    uint8_t* o = dst + (size_t)t * C;
    if (area2) {
        const uint8_t* p = img + ((size_t)(2 * y) * W + 2 * x) * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch)
            o[ch] = (uint8_t)(((int)__ldg(p + ch) + (int)__ldg(p + C + ch) + (int)__ldg(p + (size_t)W * C + ch) +
                               (int)__ldg(p + (size_t)W * C + C + ch) + 2) >> 2);
        return;
    }
    const int4 xt = __ldg(xtab + x), yt = __ldg(ytab + y);            // {i0, i1, a0, a1}
    const uint8_t* r0 = img + (size_t)yt.x * W * C;
    const uint8_t* r1 = img + (size_t)yt.y * W * C;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        const int h0 = (int)__ldg(r0 + xt.x * C + ch) * xt.z + (int)__ldg(r0 + xt.y * C + ch) * xt.w;
        const int h1 = (int)__ldg(r1 + xt.x * C + ch) * xt.z + (int)__ldg(r1 + xt.y * C + ch) * xt.w;
        const int v = (((yt.z * (h0 >> 4)) >> 16) + ((yt.w * (h1 >> 4)) >> 16) + 2) >> 2;
        o[ch] = (uint8_t)min(max(v, 0), 255);
    }
}

}  // namespace

int launch_resize_linear_u8(const uint8_t* img, int H, int W, int C, const int32_t* xtab, const int32_t* ytab, int out_h, int out_w,
                            uint8_t* dst, cudaStream_t st)
{
    const int area2 = (W == 2 * out_w && H == 2 * out_h) ? 1 : 0;
    if (!area2 && (!xtab || !ytab)) return set_error(MGW_ERR_INVALID, "resize_linear_u8: coefficient tables are required");
    const int grid = (out_h * out_w + 255) / 256;
    const int4* xt = reinterpret_cast<const int4*>(xtab);
    const int4* yt = reinterpret_cast<const int4*>(ytab);
    switch (C) {
    case 1: resize_u8_kernel<1><<<grid, 256, 0, st>>>(img, H, W, xt, yt, out_h, out_w, area2, dst); break;
    case 3: resize_u8_kernel<3><<<grid, 256, 0, st>>>(img, H, W, xt, yt, out_h, out_w, area2, dst); break;
    case 4: resize_u8_kernel<4><<<grid, 256, 0, st>>>(img, H, W, xt, yt, out_h, out_w, area2, dst); break;
    default: return set_error(MGW_ERR_UNSUPPORTED, "resize_linear_u8: C must be 1, 3 or 4 (got %d)", C);
    }
    return check_launch("resize_u8");
}

int launch_warp_rev_bundle_u8(const uint8_t* img, const double* Hs_cvt, int N, int H, int W, int C, int gh, int gw, uint8_t* dst,
                              cudaStream_t st)
{
    // WarpPerspectiveInvoker's block width: bh0 = min(16, H); bw0 = min(1024 / bh0, W)
    const int bh0 = H < 16 ? H : 16;
    const int bw = (1024 / bh0) < W ? (1024 / bh0) : W;
    const int t = N * H * W, grid = (t + 255) / 256;
    switch (C) {
    case 1: warp_rev_bundle_kernel<1><<<grid, 256, 0, st>>>(img, Hs_cvt, N, H, W, gh, gw, bw, dst); break;
    case 3: warp_rev_bundle_kernel<3><<<grid, 256, 0, st>>>(img, Hs_cvt, N, H, W, gh, gw, bw, dst); break;
    case 4: warp_rev_bundle_kernel<4><<<grid, 256, 0, st>>>(img, Hs_cvt, N, H, W, gh, gw, bw, dst); break;
    default: return set_error(MGW_ERR_UNSUPPORTED, "warp_rev_bundle_u8: C must be 1, 3 or 4 (got %d)", C);
    }
    return check_launch("warp_rev_bundle");
}

// ---------------------------------------------------------------- frame transport
// config.py:19: img * (1. / 255) - 0.5 in double (numpy: uint8 array times a Python float), then the float32 cast of the network's
// placeholder.  Frames cross PCIe as uint8 (a quarter of the fp32 bytes) and are widened here.
__global__ void u8_to_train_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, size_t n)
{
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n && ((reinterpret_cast<uintptr_t>(src) | (reinterpret_cast<uintptr_t>(dst) >> 2)) & 3) == 0) {
        const uchar4 v = *reinterpret_cast<const uchar4*>(src + i4);
        float4 o;
        o.x = (float)__dadd_rn(__dmul_rn((double)v.x, 1.0 / 255), -0.5); o.y = (float)__dadd_rn(__dmul_rn((double)v.y, 1.0 / 255), -0.5);
        o.z = (float)__dadd_rn(__dmul_rn((double)v.z, 1.0 / 255), -0.5); o.w = (float)__dadd_rn(__dmul_rn((double)v.w, 1.0 / 255), -0.5);
        *reinterpret_cast<float4*>(dst + i4) = o;
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; ++i) dst[i] = (float)__dadd_rn(__dmul_rn((double)src[i], 1.0 / 255), -0.5);
    }
}

// cvt_train2img, deploy_bundle.py:75: ((x + 0.5) * 255).astype(np.uint8) on a float32 array: two fp32 operations, truncation
// towards zero; values outside [0, 256) wrap like the x86 conversion numpy compiles to (int32 truncation, low byte; NaN -> 0).
__device__ __forceinline__ uint8_t train_to_u8(float x)
{
    const float v = __fmul_rn(__fadd_rn(x, 0.5f), 255.0f);
    const int t = (fabsf(v) < 2147483648.0f) ? __float2int_rz(v) : (int)0x80000000;
    return (uint8_t)(t & 0xff);
}

__global__ void train_to_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, size_t n)
{
    const size_t i4 = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    if (i4 >= n) return;
    if (i4 + 4 <= n && (((reinterpret_cast<uintptr_t>(src) >> 2) | reinterpret_cast<uintptr_t>(dst)) & 3) == 0) {
        const float4 v = *reinterpret_cast<const float4*>(src + i4);
        *reinterpret_cast<uchar4*>(dst + i4) = make_uchar4(train_to_u8(v.x), train_to_u8(v.y), train_to_u8(v.z), train_to_u8(v.w));
    } else {
        for (size_t i = i4; i < n && i < i4 + 4; ++i) dst[i] = train_to_u8(src[i]);
    }
}

// 16 bytes per thread and iteration; keep = the lines are written with an L2 evict_last policy, so that a buffer which is
// accumulated into next (dU of the backward: reductions at L2) is still resident when its kernel starts
// Small blocks with a handful of registers on purpose: the fill is meant to run NEXT TO the persistent warp kernels, which
// leave only ~3 K registers and a few hundred thread slots per SM free (a 256-thread block does not fit and would wait for
// them to finish).
template <bool KEEP>
__global__ void __launch_bounds__(64) fill_zero_kernel(uint4* __restrict__ p, size_t n16)
{
    // programmatic dependent launch on both sides: the blocks may be scheduled under the tail of the previous kernel (the forward
    // warp) and the next one (the backward warp) may set itself up while the fill runs
    griddep_wait();
    griddep_launch_dependents();
    uint64_t pol = 0;
    if (KEEP) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) {
        if (KEEP) asm volatile("st.global.L2::cache_hint.v4.b32 [%0], {%1, %1, %1, %1}, %2;" ::"l"(p + i), "r"(0), "l"(pol) : "memory");
        else p[i] = make_uint4(0, 0, 0, 0);
    }
}

int launch_u8_to_train(const uint8_t* src, float* dst, size_t n, cudaStream_t st)
{
    const size_t blocks = ((n + 3) / 4 + 255) / 256;
    if (blocks > 0x7fffffffULL) return set_error(MGW_ERR_INVALID, "u8_to_train: too many elements");
    u8_to_train_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n);
    return check_launch("u8_to_train");
}

int launch_train_to_u8(const float* src, uint8_t* dst, size_t n, cudaStream_t st)
{
    const size_t blocks = ((n + 3) / 4 + 255) / 256;
    if (blocks > 0x7fffffffULL) return set_error(MGW_ERR_INVALID, "train_to_u8: too many elements");
    train_to_u8_kernel<<<(unsigned)blocks, 256, 0, st>>>(src, dst, n);
    return check_launch("train_to_u8");
}

int launch_fill_zero(void* p, size_t bytes, bool keep_in_l2, cudaStream_t st)
{
    const size_t n16 = bytes / 16;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // 24 blocks of 64 threads per SM: measured next to the tile backward that now starts under the fill (step at config #2 with
    // 4 / 8 / 16 / 24 / 32 blocks per SM: 131.7 / 118.1 / 109.6 / 108.3 / 108.7 us)
    size_t per_sm = 24;
    if (const char* e = getenv("MGW_FILL_BLOCKS_PER_SM")) { const int v = atoi(e); if (v > 0) per_sm = (size_t)v; }      // tuning aid
    const size_t want = (n16 + 63) / 64, cap = (size_t)sms * per_sm;
    const unsigned grid = (unsigned)(want < cap ? want : cap);
    const cudaError_t e = keep_in_l2 ? launch_ex(fill_zero_kernel<true>, dim3(grid), dim3(64), 0, st, pdl_enabled(), reinterpret_cast<uint4*>(p), n16)
                                     : launch_ex(fill_zero_kernel<false>, dim3(grid), dim3(64), 0, st, pdl_enabled(), reinterpret_cast<uint4*>(p), n16);
    if (e != cudaSuccess) { count_launches(1); return set_error(MGW_ERR_CUDA, "fill_zero: %s", cudaGetErrorString(e)); }
    return check_launch("fill_zero");
}

}  // namespace mgw
