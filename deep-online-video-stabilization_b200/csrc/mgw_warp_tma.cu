// K2/K3, TMA-staged tile variant (the fast path of the multi-grid warp on sm_100a).
//
// One CTA = one output tile of TH x TW pixels lying inside ONE mesh cell (so one homography), 256 threads.
//   forward : the source bounding box of the tile (projective image of its 4 corners, clipped to the image) is
//             fetched by ONE TMA load (cp.async.bulk.tensor.3d) into shared memory while the threads compute the
//             per-pixel projective map; the bilinear gather then reads shared memory with conflict-free 12-byte
//             lane strides; out / x_map,y_map / black_pix are staged in shared memory and leave by TMA stores.
//   backward: the same source box of U plus the d_out / d_img tiles arrive by TMA; dU is pre-accumulated in a
//             shared-memory box (CAS float atomics, ~4 cycles per conflict-free warp instruction on B200) and leaves
//             by ONE TMA reduce-add (cp.reduce.async.bulk.tensor, L2 atomics at line granularity); the 8 dH terms
//             are reduced warp-shuffle -> shared -> one deterministic partial per tile (no global atomics).
// Every tap is checked against the staged box; taps outside it (folded cells, extreme magnification, far
// out-of-range pixels) fall back to global loads / global atomics, so results never depend on the box heuristic.
// Arithmetic is mgw_device.cuh's, bit-identical to the generic kernels and to the C oracle.
#include <cuda.h>

#include "mgw_internal.h"
#include "mgw_tma.cuh"

namespace mgw {

#define TRY_RC(expr) do { const int rc_ = (expr); if (rc_ != MGW_OK) return rc_; } while (0)

constexpr int kThreads = 256;
constexpr int kMaxRun = 6;          // rows per thread (vertical run)

struct TileCfg {
    int N, H, W, gh, gw;
    int cell_h, cell_w;             // floor(H/gh), floor(W/gw): spatial_transformer3.py:227-228
    int TH, TW, K;                  // tile rows / cols, rows per thread
    int SBH, SBW;                   // staged source box (pixels)
    int nty, ntx;                   // tiles per image
    int parts_y, parts_x;           // max tiles per cell (backward partial layout)
    int xalign;                     // box / tile start columns must be multiples of this (16-byte TMA start address)
};

struct TilePos {
    int n, ci, cj;                  // sample, cell
    int r0, c0;                     // first row / col of the tile
    int vr0, vc0;                   // first row / col this tile OWNS (edge tiles are shifted inward and overlap)
    int py, px;                     // tile index inside the cell
};

__device__ __forceinline__ void decode_axis(int t, int ncell, int cell_px, int total, int T, int& cell, int& start, int& vstart, int& part)
{
    for (cell = 0; cell < ncell; ++cell) {
        const int s = cell * cell_px;
        const int len = (cell == ncell - 1) ? total - s : cell_px;      // the last cell absorbs the remainder (:240-243)
        const int nt = (len + T - 1) / T;
        if (t < nt) {
            vstart = s + t * T;
            start = min(vstart, s + len - T);
            part = t;
            return;
        }
        t -= nt;
    }
}

__device__ __forceinline__ TilePos decode_tile(const TileCfg& c, int b)
{
    TilePos p;
    const int per = c.nty * c.ntx;
    p.n = b / per;
    const int rem = b - p.n * per;
    decode_axis(rem / c.ntx, c.gh, c.cell_h, c.H, c.TH, p.ci, p.r0, p.vr0, p.py);
    decode_axis(rem % c.ntx, c.gw, c.cell_w, c.W, c.TW, p.cj, p.c0, p.vc0, p.px);
    return p;
}

// Source box of a tile: bbox of the 4 projected corners (+1 px margin for rounding, +1 for the x1/y1 taps),
// clipped to the image like the taps are.  A projective map without a pole inside the tile (z of one sign at the 4
// corners) sends the rectangle into the convex hull of its corner images, so this box holds every tap; otherwise
// any box will do (per-tap fallback).
__device__ __forceinline__ void source_box(const float (&Hc)[9], const TileCfg& c, const TilePos& p, float stepx, float stepy,
                                           int& bx0, int& by0)
{
    float xmin = 3.0e38f, xmax = -3.0e38f, ymin = 3.0e38f, ymax = -3.0e38f;
    int sgn = 0;
    bool ok = true;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int rr = p.r0 + ((k & 2) ? c.TH - 1 : 0), cc = p.c0 + ((k & 1) ? c.TW - 1 : 0);
        const Proj q = project(Hc, lin_at(cc, stepx), lin_at(rr, stepy));
        const float x = (q.xn + 1.0f) * (float)c.W * 0.5f, y = (q.yn + 1.0f) * (float)c.H * 0.5f;
        ok = ok && (fabsf(x) < 1.0e8f) && (fabsf(y) < 1.0e8f);
        sgn += (q.zs > 0.0f) ? 1 : -1;
        xmin = fminf(xmin, x); xmax = fmaxf(xmax, x);
        ymin = fminf(ymin, y); ymax = fmaxf(ymax, y);
    }
    ok = ok && (sgn == 4 || sgn == -4);
    if (!ok) { bx0 = 0; by0 = 0; return; }
    const int ix0 = clipi((int)floorf(xmin) - 1, 0, c.W - 1), ix1 = clipi((int)floorf(xmax) + 2, 0, c.W - 1);
    const int iy0 = clipi((int)floorf(ymin) - 1, 0, c.H - 1), iy1 = clipi((int)floorf(ymax) + 2, 0, c.H - 1);
    const int needw = ix1 - ix0 + 1, needh = iy1 - iy0 + 1;
    bx0 = needw <= c.SBW ? ix0 : ix0 + (needw - c.SBW) / 2;
    by0 = needh <= c.SBH ? iy0 : iy0 + (needh - c.SBH) / 2;
    // TMA needs the box to start on a 16-byte boundary of global memory (measured on B200: a start that is not a
    // multiple of 4 floats raises "illegal instruction"): round the first column down to a multiple of xalign pixels
    bx0 -= bx0 % c.xalign;
}

// Adds val[i] to base[idx[i]] for NV shared-memory words with all NV compare-and-swaps in flight at once.  The
// compiler's atomicAdd(float*) on shared memory is a load/add/CAS loop PER CALL (ATOMS.CAST.SPIN), which serialises
// 4*C dependent ~300-cycle round trips per pixel; batching them leaves one round trip per pixel.  A CAS that loses
// (another lane, or this thread's own clipped duplicate tap, hit the same word) is retried with the value it saw.
template <int NV>
__device__ __forceinline__ void smem_add_batch(float* __restrict__ base, const int (&idx)[NV], const float (&val)[NV])
{
    int old[NV], got[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) old[i] = __float_as_int(base[idx[i]]);
#pragma unroll
    for (int i = 0; i < NV; ++i)
        got[i] = atomicCAS(reinterpret_cast<int*>(base + idx[i]), old[i], __float_as_int(__int_as_float(old[i]) + val[i]));
    unsigned pending = 0;
#pragma unroll
    for (int i = 0; i < NV; ++i) pending |= (got[i] != old[i]) ? (1u << i) : 0u;
    while (pending) {
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (pending & (1u << i)) {
                old[i] = got[i];
                got[i] = atomicCAS(reinterpret_cast<int*>(base + idx[i]), old[i], __float_as_int(__int_as_float(old[i]) + val[i]));
                if (got[i] == old[i]) pending &= ~(1u << i);
            }
        }
    }
}

__host__ __device__ constexpr int up32(int v) { return (v + 31) / 32 * 32; }     // 128-byte chunks of floats

// ------------------------------------------------------------------------------------------------ forward
template <int C>
__global__ void __launch_bounds__(kThreads)
warp_fwd_tma_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapOut,
                    const __grid_constant__ CUtensorMap mapImg, const __grid_constant__ CUtensorMap mapBlack,
                    const float* __restrict__ U, const float* __restrict__ Hs, const TileCfg cfg, const int want_out,
                    const int want_img, const int want_black)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_src = reinterpret_cast<float*>(smem_raw);
    float* s_out = s_src + up32(cfg.SBH * cfg.SBW * C);
    float* s_img = s_out + up32(cfg.TH * cfg.TW * C);
    float* s_blk = s_img + up32(cfg.TH * cfg.TW * 2);
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_blk + up32(cfg.TH * cfg.TW));

    const int tid = threadIdx.x;
    const TilePos tp = decode_tile(cfg, blockIdx.x);
    if (tid == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_barrier_init();
    }
    float Hc[9];
    {
        const float* h = Hs + ((size_t)(tp.n * cfg.gh + tp.ci) * cfg.gw + tp.cj) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) Hc[k] = __ldg(h + k);
    }
    const float stepx = lin_step(cfg.W), stepy = lin_step(cfg.H);
    int bx0, by0;
    source_box(Hc, cfg, tp, stepx, stepy, bx0, by0);
    __syncthreads();
    if (tid == 0 && want_out) {
        tma::mbar_expect_tx(bar, (uint32_t)(cfg.SBH * cfg.SBW * C * sizeof(float)));
        tma::load_3d(s_src, &mapU, bar, bx0 * C, by0, tp.n);
    }

    const int tx = tid % cfg.TW, tyg = tid / cfg.TW;
    const float xt = lin_at(tp.c0 + tx, stepx);
    float xn[kMaxRun], yn[kMaxRun];
#pragma unroll
    for (int k = 0; k < kMaxRun; ++k) {
        const int lr = tyg * cfg.K + k;
        xn[k] = 0.0f; yn[k] = 0.0f;
        if (k < cfg.K && lr < cfg.TH) {
            const Proj q = project(Hc, xt, lin_at(tp.r0 + lr, stepy));
            xn[k] = q.xn; yn[k] = q.yn;
            const int o = lr * cfg.TW + tx;
            reinterpret_cast<float2*>(s_img)[o] = make_float2(q.xn, q.yn);
            s_blk[o] = black_of(q.xn, q.yn);
        }
    }
    if (want_out) {
        tma::mbar_wait(bar, 0);
        const float* Un = U + (size_t)tp.n * cfg.H * cfg.W * C;
#pragma unroll
        for (int k = 0; k < kMaxRun; ++k) {
            const int lr = tyg * cfg.K + k;
            if (k < cfg.K && lr < cfg.TH) {
                const Taps t = make_taps(xn[k], yn[k], cfg.H, cfg.W);
                const int sx0 = t.x0 - bx0, sx1 = t.x1 - bx0, sy0 = t.y0 - by0, sy1 = t.y1 - by0;
                float* o = s_out + (lr * cfg.TW + tx) * C;
                if (sx0 >= 0 && sx1 < cfg.SBW && sy0 >= 0 && sy1 < cfg.SBH) {
                    const float* pa = s_src + (sy0 * cfg.SBW + sx0) * C;
                    const float* pb = s_src + (sy1 * cfg.SBW + sx0) * C;
                    const float* pc = s_src + (sy0 * cfg.SBW + sx1) * C;
                    const float* pd = s_src + (sy1 * cfg.SBW + sx1) * C;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) o[ch] = blend(t, pa[ch], pb[ch], pc[ch], pd[ch]);
                } else {
                    const float* pa = Un + ((size_t)t.y0 * cfg.W + t.x0) * C;
                    const float* pb = Un + ((size_t)t.y1 * cfg.W + t.x0) * C;
                    const float* pc = Un + ((size_t)t.y0 * cfg.W + t.x1) * C;
                    const float* pd = Un + ((size_t)t.y1 * cfg.W + t.x1) * C;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) o[ch] = blend(t, __ldg(pa + ch), __ldg(pb + ch), __ldg(pc + ch), __ldg(pd + ch));
                }
            }
        }
    }
    tma::fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        if (want_out) tma::store_3d(&mapOut, s_out, tp.c0 * C, tp.r0, tp.n);
        if (want_img) tma::store_3d(&mapImg, s_img, tp.c0 * 2, tp.r0, tp.n);
        if (want_black) tma::store_3d(&mapBlack, s_blk, tp.c0, tp.r0, tp.n);
        tma::commit_group();
        tma::wait_group_read0();
    }
}

// ------------------------------------------------------------------------------------------------ backward
template <int C>
__global__ void __launch_bounds__(kThreads)
warp_bwd_tma_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapDout,
                    const __grid_constant__ CUtensorMap mapDimg, const __grid_constant__ CUtensorMap mapDU,
                    const float* __restrict__ U, const float* __restrict__ Hs, const TileCfg cfg, float* __restrict__ dU,
                    const int has_dimg, float* __restrict__ parts)
{
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    const int box_f = up32(cfg.SBH * cfg.SBW * C);
    float* s_src = reinterpret_cast<float*>(smem_raw);
    float* s_acc = s_src + box_f;
    float* s_dout = s_acc + (dU ? box_f : 0);
    float* s_dimg = s_dout + up32(cfg.TH * cfg.TW * C);
    float* s_red = s_dimg + (has_dimg ? up32(cfg.TH * cfg.TW * 2) : 0);      // [8 warps][8]
    uint64_t* bar = reinterpret_cast<uint64_t*>(s_red + 64);

    const int tid = threadIdx.x;
    const TilePos tp = decode_tile(cfg, blockIdx.x);
    if (tid == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_barrier_init();
    }
    float Hc[9];
    {
        const float* h = Hs + ((size_t)(tp.n * cfg.gh + tp.ci) * cfg.gw + tp.cj) * 9;
#pragma unroll
        for (int k = 0; k < 9; ++k) Hc[k] = __ldg(h + k);
    }
    const float stepx = lin_step(cfg.W), stepy = lin_step(cfg.H);
    int bx0, by0;
    source_box(Hc, cfg, tp, stepx, stepy, bx0, by0);
    __syncthreads();
    if (tid == 0) {
        const uint32_t bytes = (uint32_t)((cfg.SBH * cfg.SBW * C + cfg.TH * cfg.TW * C + (has_dimg ? cfg.TH * cfg.TW * 2 : 0)) * sizeof(float));
        tma::mbar_expect_tx(bar, bytes);
        tma::load_3d(s_src, &mapU, bar, bx0 * C, by0, tp.n);
        tma::load_3d(s_dout, &mapDout, bar, tp.c0 * C, tp.r0, tp.n);
        if (has_dimg) tma::load_3d(s_dimg, &mapDimg, bar, tp.c0 * 2, tp.r0, tp.n);
    }
    if (dU) {
        float4* a4 = reinterpret_cast<float4*>(s_acc);
        for (int i = tid; i < box_f / 4; i += kThreads) a4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    __syncthreads();

    const int tx = tid % cfg.TW, tyg = tid / cfg.TW;
    const int col = tp.c0 + tx;
    const float xt = lin_at(col, stepx);
    const float* Un = U + (size_t)tp.n * cfg.H * cfg.W * C;
    float* dUn = dU ? dU + (size_t)tp.n * cfg.H * cfg.W * C : nullptr;
    float dh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dh[k] = 0.0f;
    const float halfW = 0.5f * (float)cfg.W, halfH = 0.5f * (float)cfg.H;

    tma::mbar_wait(bar, 0);
#pragma unroll
    for (int k = 0; k < kMaxRun; ++k) {
        const int lr = tyg * cfg.K + k;
        const int row = tp.r0 + lr;
        if (k < cfg.K && lr < cfg.TH && row >= tp.vr0 && col >= tp.vc0) {
            const float yt = lin_at(row, stepy);
            const Proj q = project(Hc, xt, yt);
            const Taps t = make_taps(q.xn, q.yn, cfg.H, cfg.W);
            const int sx0 = t.x0 - bx0, sx1 = t.x1 - bx0, sy0 = t.y0 - by0, sy1 = t.y1 - by0;
            const bool inbox = sx0 >= 0 && sx1 < cfg.SBW && sy0 >= 0 && sy1 < cfg.SBH;
            const int o = lr * cfg.TW + tx;
            const float wa = t.ax * t.ay, wb = t.ax * t.by, wc = t.bx * t.ay, wd = t.bx * t.by;
            float gx = 0.0f, gy = 0.0f;
            if (inbox) {
                const int ia = (sy0 * cfg.SBW + sx0) * C, ib = (sy1 * cfg.SBW + sx0) * C;
                const int ic = (sy0 * cfg.SBW + sx1) * C, id = (sy1 * cfg.SBW + sx1) * C;
                int aidx[4 * C];
                float aval[4 * C];
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float g = s_dout[o * C + ch];
                    const float Ia = s_src[ia + ch], Ib = s_src[ib + ch], Ic = s_src[ic + ch], Id = s_src[id + ch];
                    gx = fmaf(g, fmaf(Ic - Ia, t.ay, (Id - Ib) * t.by), gx);
                    gy = fmaf(g, fmaf(Ib - Ia, t.ax, (Id - Ic) * t.bx), gy);
                    aidx[4 * ch + 0] = ia + ch; aval[4 * ch + 0] = wa * g;
                    aidx[4 * ch + 1] = ib + ch; aval[4 * ch + 1] = wb * g;
                    aidx[4 * ch + 2] = ic + ch; aval[4 * ch + 2] = wc * g;
                    aidx[4 * ch + 3] = id + ch; aval[4 * ch + 3] = wd * g;
                }
                if (dU) smem_add_batch<4 * C>(s_acc, aidx, aval);
            } else {
                const size_t ia = ((size_t)t.y0 * cfg.W + t.x0) * C, ib = ((size_t)t.y1 * cfg.W + t.x0) * C;
                const size_t ic = ((size_t)t.y0 * cfg.W + t.x1) * C, id = ((size_t)t.y1 * cfg.W + t.x1) * C;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float g = s_dout[o * C + ch];
                    const float Ia = __ldg(Un + ia + ch), Ib = __ldg(Un + ib + ch), Ic = __ldg(Un + ic + ch), Id = __ldg(Un + id + ch);
                    gx = fmaf(g, fmaf(Ic - Ia, t.ay, (Id - Ib) * t.by), gx);
                    gy = fmaf(g, fmaf(Ib - Ia, t.ax, (Id - Ic) * t.bx), gy);
                    if (dU) {
                        atomicAdd(dUn + ia + ch, wa * g);
                        atomicAdd(dUn + ib + ch, wb * g);
                        atomicAdd(dUn + ic + ch, wc * g);
                        atomicAdd(dUn + id + ch, wd * g);
                    }
                }
            }
            float gxn = gx * halfW, gyn = gy * halfH;
            if (has_dimg) {
                const float2 di = reinterpret_cast<const float2*>(s_dimg)[o];
                gxn += di.x; gyn += di.y;
            }
            const float rz = 1.0f / q.zs;
            const float dxs = gxn * rz, dys = gyn * rz;
            const float dzs = -(gxn * q.xn + gyn * q.yn) * rz;
            dh[0] = fmaf(dxs, xt, dh[0]); dh[1] = fmaf(dxs, yt, dh[1]); dh[2] += dxs;
            dh[3] = fmaf(dys, xt, dh[3]); dh[4] = fmaf(dys, yt, dh[4]); dh[5] += dys;
            dh[6] = fmaf(dzs, xt, dh[6]); dh[7] = fmaf(dzs, yt, dh[7]);
        }
    }
    // dH: warp shuffle -> shared -> one partial per tile
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const float v = warp_sum(dh[k]);
        if ((tid & 31) == 0) s_red[(tid >> 5) * 8 + k] = v;
    }
    if (dU) tma::fence_proxy_async();
    __syncthreads();
    if (tid < 8) {
        float v = 0.0f;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) v += s_red[w * 8 + tid];
        const size_t cell = (size_t)(tp.n * cfg.gh + tp.ci) * cfg.gw + tp.cj;
        const int part = tp.py * cfg.parts_x + tp.px;
        parts[(cell * (cfg.parts_y * cfg.parts_x) + part) * 8 + tid] = v;
    }
    if (tid == 0 && dU) {
        tma::reduce_add_3d(&mapDU, s_acc, bx0 * C, by0, tp.n);
        tma::commit_group();
        tma::wait_group_read0();
    }
}

// ------------------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 3-D view [N][rows][inner floats] of an NHWC tensor, box = [1][box_rows][box_inner]
static int make_map(CUtensorMap* m, const void* base, int inner, int rows, int N, int box_inner, int box_rows)
{
    EncodeTiledFn enc = encode_fn();
    if (!enc) return set_error(MGW_ERR_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[3] = {(cuuint64_t)inner, (cuuint64_t)rows, (cuuint64_t)N};
    const cuuint64_t strides[2] = {(cuuint64_t)inner * 4, (cuuint64_t)inner * 4 * rows};
    const cuuint32_t box[3] = {(cuuint32_t)box_inner, (cuuint32_t)box_rows, 1};
    const cuuint32_t es[3] = {1, 1, 1};
    const CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<void*>(base), dims, strides, box, es,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(MGW_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) inner=%d rows=%d box=%dx%d", (int)r, inner, rows, box_inner, box_rows);
    return MGW_OK;
}

static int tiles_along(int ncell, int cell_px, int total, int T, int* max_per_cell)
{
    int n = 0, mx = 0;
    for (int c = 0; c < ncell; ++c) {
        const int s = c * cell_px, len = (c == ncell - 1) ? total - s : cell_px;
        const int nt = (len + T - 1) / T;
        n += nt;
        mx = nt > mx ? nt : mx;
    }
    *max_per_cell = mx;
    return n;
}

// Picks the tiling for a shape; false if the TMA path cannot serve it (the generic kernels then do).
static bool plan(const WarpShape& s, TileCfg* out)
{
    if (s.OH != s.H || s.OW != s.W) return false;
    if (s.C < 1 || s.C > 4) return false;
    if (((size_t)s.W * s.C * 4) % 16 != 0 || ((size_t)s.W * 4) % 16 != 0) return false;      // TMA global strides
    TileCfg c;
    c.N = s.N; c.H = s.H; c.W = s.W; c.gh = s.gh; c.gw = s.gw;
    c.cell_h = s.H / s.gh; c.cell_w = s.W / s.gw;
    // tiles start at cell boundaries (or cell end - TW) and every TMA start address must be 16-byte aligned for
    // out (C floats/px), black (1) and x/y maps (2): columns of cell boundaries must be multiples of 4
    if (s.W % 4 != 0 || c.cell_w % 4 != 0) return false;
    c.xalign = (s.C % 4 == 0) ? 1 : ((s.C % 2 == 0) ? 2 : 4);
    const int max_inner = 256;                                                               // TMA box dim limit
    c.TW = 0;
    for (int tw : {64, 32}) {
        if (tw > c.cell_w) continue;
        int sbw = ((tw * 13 + 9) / 10 + 4 + (c.xalign - 1) + 3) / 4 * 4;
        const int cap = (max_inner / s.C) / 4 * 4;
        if (sbw > cap) sbw = cap;
        if (sbw < tw + 4) continue;
        if (tw * s.C > max_inner) continue;
        c.TW = tw; c.SBW = sbw;
        break;
    }
    if (c.TW == 0) return false;
    const int groups = kThreads / c.TW;
    const int th_max = groups * kMaxRun < 24 ? groups * kMaxRun : 24;
    if (c.cell_h < 8) return false;
    int best = 0; double best_eff = 0;
    for (int th = 8; th <= th_max && th <= c.cell_h; ++th) {
        const int nt = (c.cell_h + th - 1) / th;
        const double eff = (double)c.cell_h / (nt * th) * (th >= 12 ? 1.0 : 0.9);
        if (eff >= best_eff) { best_eff = eff; best = th; }
    }
    c.TH = best;
    c.K = (c.TH + groups - 1) / groups;
    c.SBH = (c.TH * 13 + 9) / 10 + 4;
    c.nty = tiles_along(s.gh, c.cell_h, s.H, c.TH, &c.parts_y);
    c.ntx = tiles_along(s.gw, c.cell_w, s.W, c.TW, &c.parts_x);
    if ((long long)c.N * c.nty * c.ntx > 0x7fffffffLL) return false;
    *out = c;
    return true;
}

static size_t fwd_smem(const TileCfg& c, int C)
{
    return (size_t)(up32(c.SBH * c.SBW * C) + up32(c.TH * c.TW * C) + up32(c.TH * c.TW * 2) + up32(c.TH * c.TW)) * 4 + 128;
}

static size_t bwd_smem(const TileCfg& c, int C, bool has_dU, bool has_dimg)
{
    return (size_t)(up32(c.SBH * c.SBW * C) * (has_dU ? 2 : 1) + up32(c.TH * c.TW * C) + (has_dimg ? up32(c.TH * c.TW * 2) : 0) + 64) * 4 + 128;
}

bool tma_fwd_supported(const WarpShape& s)
{
    TileCfg c;
    return plan(s, &c) && fwd_smem(c, s.C) <= 200 * 1024;
}

bool tma_bwd_supported(const WarpShape& s)
{
    TileCfg c;
    return plan(s, &c) && bwd_smem(c, s.C, true, true) <= 200 * 1024;
}

size_t tma_bwd_workspace_bytes(const WarpShape& s)
{
    TileCfg c;
    if (!plan(s, &c)) return 0;
    return (size_t)s.N * s.gh * s.gw * c.parts_y * c.parts_x * 8 * sizeof(float);
}

template <int C>
static int launch_fwd_c(const float* U, const float* Hs, const TileCfg& c, float* out, float* black, float* img, cudaStream_t st)
{
    CUtensorMap mU, mOut, mImg, mBlk;
    TRY_RC(make_map(&mU, U, c.W * C, c.H, c.N, c.SBW * C, c.SBH));
    TRY_RC(make_map(&mOut, out ? out : U, c.W * C, c.H, c.N, c.TW * C, c.TH));
    if (img) TRY_RC(make_map(&mImg, img, c.W * 2, c.H, c.N, c.TW * 2, c.TH)); else mImg = mOut;
    if (black) TRY_RC(make_map(&mBlk, black, c.W, c.H, c.N, c.TW, c.TH)); else mBlk = mOut;
    const size_t smem = fwd_smem(c, C);
    static bool attr_set[64] = {};
    int devid = 0;
    cudaGetDevice(&devid);
    if (!attr_set[devid & 63]) {
        if (cudaFuncSetAttribute(warp_fwd_tma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return set_error(MGW_ERR_CUDA, "cudaFuncSetAttribute(warp_fwd_tma): %s", cudaGetErrorString(cudaGetLastError()));
        attr_set[devid & 63] = true;
    }
    const unsigned grid = (unsigned)(c.N * c.nty * c.ntx);
    warp_fwd_tma_kernel<C><<<grid, kThreads, smem, st>>>(mU, mOut, mImg, mBlk, U, Hs, c, out != nullptr, img != nullptr, black != nullptr);
    return check_launch("warp_fwd_tma");
}

int launch_warp_fwd_tma(const float* U, const float* Hs, const WarpShape& s, float* out, float* black, float* img, cudaStream_t st)
{
    TileCfg c;
    if (!plan(s, &c)) return set_error(MGW_ERR_UNSUPPORTED, "warp_fwd_tma: unsupported shape");
    switch (s.C) {
        case 1: return launch_fwd_c<1>(U, Hs, c, out, black, img, st);
        case 2: return launch_fwd_c<2>(U, Hs, c, out, black, img, st);
        case 3: return launch_fwd_c<3>(U, Hs, c, out, black, img, st);
        default: return launch_fwd_c<4>(U, Hs, c, out, black, img, st);
    }
}

template <int C>
static int launch_bwd_c(const float* U, const float* Hs, const float* d_out, const float* d_img, const TileCfg& c, float* dU,
                        float* parts, cudaStream_t st)
{
    CUtensorMap mU, mDout, mDimg, mDU;
    TRY_RC(make_map(&mU, U, c.W * C, c.H, c.N, c.SBW * C, c.SBH));
    TRY_RC(make_map(&mDout, d_out, c.W * C, c.H, c.N, c.TW * C, c.TH));
    if (d_img) TRY_RC(make_map(&mDimg, d_img, c.W * 2, c.H, c.N, c.TW * 2, c.TH)); else mDimg = mDout;
    if (dU) TRY_RC(make_map(&mDU, dU, c.W * C, c.H, c.N, c.SBW * C, c.SBH)); else mDU = mU;
    const size_t smem = bwd_smem(c, C, dU != nullptr, d_img != nullptr);
    static bool attr_set[64] = {};
    int devid = 0;
    cudaGetDevice(&devid);
    if (!attr_set[devid & 63]) {
        if (cudaFuncSetAttribute(warp_bwd_tma_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return set_error(MGW_ERR_CUDA, "cudaFuncSetAttribute(warp_bwd_tma): %s", cudaGetErrorString(cudaGetLastError()));
        attr_set[devid & 63] = true;
    }
    const unsigned grid = (unsigned)(c.N * c.nty * c.ntx);
    warp_bwd_tma_kernel<C><<<grid, kThreads, smem, st>>>(mU, mDout, mDimg, mDU, U, Hs, c, dU, d_img != nullptr, parts);
    return check_launch("warp_bwd_tma");
}

int launch_warp_bwd_tma(const float* U, const float* Hs, const float* d_out, const float* d_img, const WarpShape& s, float* dU,
                        float* parts, int* nparts, cudaStream_t st)
{
    TileCfg c;
    if (!plan(s, &c)) return set_error(MGW_ERR_UNSUPPORTED, "warp_bwd_tma: unsupported shape");
    *nparts = c.parts_y * c.parts_x;
    // cells with fewer tiles than parts_y*parts_x leave slots untouched: zero them (a few hundred KB at most)
    if (cudaMemsetAsync(parts, 0, tma_bwd_workspace_bytes(s), st) != cudaSuccess)
        return set_error(MGW_ERR_CUDA, "memset parts: %s", cudaGetErrorString(cudaGetLastError()));
    switch (s.C) {
        case 1: return launch_bwd_c<1>(U, Hs, d_out, d_img, c, dU, parts, st);
        case 2: return launch_bwd_c<2>(U, Hs, d_out, d_img, c, dU, parts, st);
        case 3: return launch_bwd_c<3>(U, Hs, d_out, d_img, c, dU, parts, st);
        default: return launch_bwd_c<4>(U, Hs, d_out, d_img, c, dU, parts, st);
    }
}

}  // namespace mgw
