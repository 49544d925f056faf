// Shared pieces of the persistent, warp-specialised TMA pipelines (mgw_warp_pipe.cu: forward; mgw_warp_bwd_pipe.cu: backward):
// tile records prepared by the producer warp, staged-box geometry, per-pixel tap state, tile planning.
#pragma once
#include "mgw_tile.cuh"

namespace mgw {
namespace pipe {


struct PipeCfg {
    TileCfg t;
    int nty, ntx;                   // tiles per image along y / x
    int total;                      // N * nty * ntx
    // backward: a CTA walks CHUNKS of chunk_L vertically adjacent tiles of one cell (one dH partial per chunk); 1 = tile by tile
    int chunk_L, chunk_rows, nchunks;      // tiles per chunk, chunks per image column (nty / chunk_L), N * chunk_rows * ntx
};

// BW x BH = staged source box in pixels.  The benchmark meshes (sigma = 0.05) stretch and shear a 32 x 24 tile into a
// tap bounding box of 39 x 28 px at the median and 57 x 39 px at the 99th percentile (tools/box_stats.py): the box is
// sized for ~p99 and the rest takes the per-pixel fallback.
template <int C, int TW, int TH, int BW, int BH>
struct PGeo {
    static constexpr int kXalign = (C % 4 == 0) ? 1 : ((C % 2 == 0) ? 2 : 4);
    // row pitch of the staged box in floats (= TMA box inner dimension, <= 256): a multiple of 32 so that a tap's bank
    // depends only on its column (see mgw_warp_tma.cu)
    static constexpr int kWantF = (BW * C + 31) / 32 * 32;
    static constexpr int kRowF = kWantF < 256 ? kWantF : 256;
    static constexpr int SBW = kRowF / C;
    static constexpr int SBH = BH;
    static constexpr int kBoxF = SBH * kRowF;
    static constexpr int kOutF = TH * TW * C;
    static_assert(SBW >= TW + 4, "source box too narrow for this tile width / channel count");
    static_assert(TW % 32 == 0, "a warp covers 32 consecutive columns");
};

// what the producer publishes per stage
struct __align__(16) PInfo {
    float Hc[9];
    int n, r0, c0;                  // image, first row / column of the tile
    int vr0, vc0;                   // first row / column the tile OWNS (edge tiles are shifted inward)
    int bx0, by0;                   // first column / row of the staged source box
    int complete;                   // every (clipped) tap of every pixel lies inside the staged box
    int cell, part;                 // backward: flat cell index, index of the tile inside its cell (dH partial slot)
    int nrow, nq;                   // backward: rows / 16-byte groups per row of the box that can hold a tap (what the drain visits)
    int pad[11];
};
static_assert(sizeof(PInfo) == 128, "PInfo is one 128-byte record");

// The producer prepares kRoundTiles tiles at a time, 4 lanes per tile (one projected corner each), into a ring of records
// that the consumers read directly: the per-tile serial chain (decode, homography fetch, projection, box) is paid once per
// round instead of once per tile, so a single producer warp keeps up with the consumers.
constexpr int kRoundTiles = 8;
constexpr int kInfoRing = 16;

// lanes 4j..4j+3 handle tile t (all four decode it; lane&3 picks the corner); lane 4j writes the record.
// Source box = bbox of the projected corners (+1 px for rounding, +1 for the x1/y1 taps), clipped to the image like the
// taps are.
template <class G, int TW, int TH, int C>
__device__ __forceinline__ void make_record(const PipeCfg& cfg, const float* __restrict__ Hs, int t, bool valid, float stepx,
                                            float stepy, PInfo* rec, int lane, bool writer)
{
    const int H = cfg.t.H, W = cfg.t.W;
    const int tx = t % cfg.ntx, qq = t / cfg.ntx, ty = qq % cfg.nty;
    const int n = qq / cfg.nty;
    const int r0 = cfg.t.rows.start[ty], c0 = cfg.t.cols.start[tx];
    const int cell = (n * cfg.t.gh + cfg.t.rows.cell[ty]) * cfg.t.gw + cfg.t.cols.cell[tx];
    float Hc[9];
#pragma unroll
    for (int k = 0; k < 9; ++k) Hc[k] = __ldg(Hs + (size_t)cell * 9 + k);
    const int k4 = lane & 3;
    const int rr = r0 + ((k4 & 2) ? TH - 1 : 0), cc = c0 + ((k4 & 1) ? TW - 1 : 0);
    const float xtc = lin_at(cc, stepx), ytc = lin_at(rr, stepy);
    const Proj q = project(Hc, xtc, ytc);
    const float x = (q.xn + 1.0f) * (float)W * 0.5f, y = (q.yn + 1.0f) * (float)H * 0.5f;
    bool ok = (fabsf(x) < 1.0e8f) && (fabsf(y) < 1.0e8f);
    // xs, ys, zs are affine over the tile: bounded by their corner values (z of one sign => |z| >= its corner minimum).
    // Inside (2^-50, 2^50) the consumers' shared-reciprocal division needs no range test on the denominator.
    ok = ok && fabsf(q.zs) > 8.9e-16f && fabsf(q.zs) < 1.1e15f && fabsf(hrow(Hc[0], Hc[1], Hc[2], xtc, ytc)) < 1.1e15f &&
         fabsf(hrow(Hc[3], Hc[4], Hc[5], xtc, ytc)) < 1.1e15f;
    int sgn = (q.zs > 0.0f) ? 1 : -1;
    float xmin = x, xmax = x, ymin = y, ymax = y;
#pragma unroll
    for (int o = 1; o < 4; o <<= 1) {
        xmin = fminf(xmin, __shfl_xor_sync(0xffffffffu, xmin, o));
        xmax = fmaxf(xmax, __shfl_xor_sync(0xffffffffu, xmax, o));
        ymin = fminf(ymin, __shfl_xor_sync(0xffffffffu, ymin, o));
        ymax = fmaxf(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
        sgn += __shfl_xor_sync(0xffffffffu, sgn, o);
        ok = ok && (__shfl_xor_sync(0xffffffffu, (int)ok, o) != 0);
    }
    if (!writer || !valid) return;
    ok = ok && (sgn == 4 || sgn == -4);
    int bx0 = 0, by0 = 0, complete = 0, ncol = G::SBW, nrow = G::SBH;
    if (ok) {
        const int ux0 = (int)floorf(xmin) - 1, ux1 = (int)floorf(xmax) + 2;      // unclipped tap range, 1 px of slack
        const int uy0 = (int)floorf(ymin) - 1, uy1 = (int)floorf(ymax) + 2;
        const int ix0 = clipi(ux0, 0, W - 1), ix1 = clipi(ux1, 0, W - 1);      // the taps are clipped like this too
        const int iy0 = clipi(uy0, 0, H - 1), iy1 = clipi(uy1, 0, H - 1);
        const int needw = ix1 - ix0 + 1, needh = iy1 - iy0 + 1;
        bx0 = needw <= G::SBW ? ix0 : ix0 + (needw - G::SBW) / 2;              // too large: centre the box, the rest falls back
        by0 = needh <= G::SBH ? iy0 : iy0 + (needh - G::SBH) / 2;
        // TMA needs the box to start on a 16-byte boundary of global memory: round the first column down
        bx0 -= bx0 % G::kXalign;
        complete = (ix0 >= bx0 && ix1 - bx0 < G::SBW && iy0 >= by0 && iy1 - by0 < G::SBH) ? 1 : 0;
        if (complete) { ncol = ix1 - bx0 + 1; nrow = iy1 - by0 + 1; }
    }
#pragma unroll
    for (int k = 0; k < 9; ++k) rec->Hc[k] = Hc[k];
    rec->n = n; rec->r0 = r0; rec->c0 = c0; rec->vr0 = cfg.t.rows.vstart[ty]; rec->vc0 = cfg.t.cols.vstart[tx];
    rec->bx0 = bx0; rec->by0 = by0; rec->complete = complete;
    rec->cell = cell;
    rec->part = cfg.chunk_L > 1 ? cfg.t.cols.part[tx] : cfg.t.rows.part[ty] * cfg.t.parts_x + cfg.t.cols.part[tx];
    // the part of the box inside the image ((W - bx0) * C is a multiple of 4 floats: W % 4 == 0 and bx0 % kXalign == 0)
    rec->nrow = min(nrow, H - by0);
    rec->nq = min((min(ncol * C, G::kRowF) + 3) / 4, (W - bx0) * C / 4);
}

// records of tiles it .. it+kRoundTiles-1 of this CTA (tile index t, stride gridDim.x)
template <class G, int TW, int TH, int C>
__device__ __forceinline__ void prepare_round(const PipeCfg& cfg, const float* __restrict__ Hs, int it, int t, float stepx, float stepy,
                                              PInfo* info, int lane)
{
    const int j = lane >> 2;
    const long long tj = (long long)t + (long long)j * gridDim.x;
    const bool valid = tj < cfg.total;
    make_record<G, TW, TH, C>(cfg, Hs, valid ? (int)tj : cfg.total - 1, valid, stepx, stepy, info + ((it + j) % kInfoRing), lane, (lane & 3) == 0);
    __syncwarp();
}

// Per-pixel state of a COMPLETE tile between the phases of the consumers' loops.  The loops run phase by phase over the
// thread's K pixels with no branch inside a phase, so that the compiler interleaves the K independent chains.
//   taps, common case (no tap clipped): x0f = floor(x), x1f = x0f + 1, so bx = x - x0f is exact and
//   ax = x1f - x = RN(1 - bx): the reference's weights (spatial_transformer3.py:114-121) bit for bit.
//   Border pixels (a rare, out-of-line fix-up) redo it with the clipped integers.
struct PixTaps {
    int off;                        // BYTE offset of tap (y0,x0) in the box
    int dx, dy;                     // byte offsets to the x1 / y1 taps (0 when clipping collapsed the pair)
    float ax, bx, ay, by;
};

template <int C, int kRowF>
__device__ __forceinline__ bool pix_taps(float xn, float yn, int IH, int IW, int offbase, PixTaps& t)
{
    const float x = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.0f), (float)IW), 0.5f);   // ((xn+1)*W)/2  (:81)
    const float y = __fmul_rn(__fmul_rn(__fadd_rn(yn, 1.0f), (float)IH), 0.5f);
    const float fx = floorf(x), fy = floorf(y);
    const int x0 = __float2int_rz(fx), y0 = __float2int_rz(fy);
    t.bx = __fsub_rn(x, fx); t.ax = __fsub_rn(1.0f, t.bx);
    t.by = __fsub_rn(y, fy); t.ay = __fsub_rn(1.0f, t.by);
    t.dx = C * 4; t.dy = kRowF * 4;
    t.off = (y0 * kRowF + x0 * C + offbase) * 4;
    return !((unsigned)x0 < (unsigned)(IW - 1) && (unsigned)y0 < (unsigned)(IH - 1));
}

template <int C, int kRowF>
__device__ __forceinline__ void pix_taps_clipped(float xn, float yn, int IH, int IW, int offbase, PixTaps& t)
{
    const float x = __fmul_rn(__fmul_rn(__fadd_rn(xn, 1.0f), (float)IW), 0.5f);
    const float y = __fmul_rn(__fmul_rn(__fadd_rn(yn, 1.0f), (float)IH), 0.5f);
    const int x0 = __float2int_rz(floorf(x)), y0 = __float2int_rz(floorf(y));
    const int x0c = clipi(x0, 0, IW - 1), x1c = clipi(x0 + 1, 0, IW - 1);
    const int y0c = clipi(y0, 0, IH - 1), y1c = clipi(y0 + 1, 0, IH - 1);
    t.ax = __fsub_rn((float)x1c, x); t.bx = __fsub_rn(x, (float)x0c);
    t.ay = __fsub_rn((float)y1c, y); t.by = __fsub_rn(y, (float)y0c);
    t.dx = (x1c - x0c) * C * 4; t.dy = (y1c - y0c) * kRowF * 4;
    t.off = (y0c * kRowF + x0c * C + offbase) * 4;
}

// xs/zs, ys/zs in a COMPLETE tile: the denominator and the numerators' upper bound were range-checked by the producer;
// what is left per pixel is a numerator that is zero / tiny (the projected axis crosses the tile), flagged in `bad`
__device__ __forceinline__ float div2_tile(float xs, float ys, float zs, float& xn, float& yn, bool& bad)
{
    float r0;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(zs));
    const float e = __fmaf_rn(-zs, r0, 1.0f);
    const float r = __fmaf_rn(r0, e, r0);
    const float qx = __fmul_rn(xs, r), qy = __fmul_rn(ys, r);
    xn = __fmaf_rn(r, __fmaf_rn(-zs, qx, xs), qx);
    yn = __fmaf_rn(r, __fmaf_rn(-zs, qy, ys), qy);
    bad = bad || !(fminf(fabsf(xs), fabsf(ys)) > 8.6736174e-19f);      // 2^-60
    return r;
}

__device__ __forceinline__ float blend4(float ax, float bx, float ay, float by, float Ia, float Ib, float Ic, float Id)
{
    const float wa = __fmul_rn(ax, ay), wb = __fmul_rn(ax, by), wc = __fmul_rn(bx, ay), wd = __fmul_rn(bx, by);
    float s = __fmul_rn(wa, Ia);
    s = __fadd_rn(s, __fmul_rn(wb, Ib));
    s = __fadd_rn(s, __fmul_rn(wc, Ic));
    s = __fadd_rn(s, __fmul_rn(wd, Id));
    return s;
}

// ---- host side
struct PipePlan {
    PipeCfg cfg;
    int TW, TH;
};

// coverage efficiency of a tile shape on this mesh (tiles never straddle cells; edge tiles are shifted inward), 0 = no fit
static double tile_eff(const WarpShape& s, int TW, int TH)
{
    if (s.OH != s.H || s.OW != s.W) return 0;
    if (s.H > 65535 || s.W > 65535 || s.gh > 255 || s.gw > 255 || s.N > 65535) return 0;
    const int cell_h = s.H / s.gh, cell_w = s.W / s.gw;
    // tiles start at cell boundaries (or cell end - TW) and every TMA start address must be 16-byte aligned
    if (s.W % 4 != 0 || cell_w % 4 != 0) return 0;
    if (TW > cell_w || TH > cell_h) return 0;
    const double ey = (double)cell_h / (((cell_h + TH - 1) / TH) * TH), ex = (double)cell_w / (((cell_w + TW - 1) / TW) * TW);
    return ey * ex;
}

static bool plan(const WarpShape& s, int TW, int TH, PipePlan* out)
{
    if (tile_eff(s, TW, TH) <= 0) return false;
    const int cell_h = s.H / s.gh, cell_w = s.W / s.gw;
    PipePlan p;
    p.TW = TW; p.TH = TH;
    TileCfg& c = p.cfg.t;
    c.N = s.N; c.H = s.H; c.W = s.W; c.gh = s.gh; c.gw = s.gw;
    tile_steps(&c);
    p.cfg.nty = fill_axis(&c.rows, s.gh, cell_h, s.H, TH, &c.parts_y);
    p.cfg.ntx = fill_axis(&c.cols, s.gw, cell_w, s.W, TW, &c.parts_x);
    if (p.cfg.nty < 0 || p.cfg.ntx < 0 || c.parts_y > 255 || c.parts_x > 255) return false;
    const long long total = (long long)s.N * p.cfg.nty * p.cfg.ntx;
    if (total >= (1LL << 31)) return false;
    p.cfg.total = (int)total;
    p.cfg.chunk_L = 1; p.cfg.chunk_rows = p.cfg.nty; p.cfg.nchunks = p.cfg.total;
    *out = p;
    return true;
}

static int sm_count()
{
    static int n[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!n[dev & 63]) cudaDeviceGetAttribute(&n[dev & 63], cudaDevAttrMultiProcessorCount, dev);
    return n[dev & 63] > 0 ? n[dev & 63] : 148;
}

// tuning aid: MGW_PIPE_GRID=<CTAs> overrides the persistent grid size
static int grid_for(int total, int ctas_per_sm)
{
    int g = ctas_per_sm * sm_count();
    if (const char* e = getenv("MGW_PIPE_GRID")) { const int v = atoi(e); if (v > 0) g = v; }
    return g < total ? g : total;
}


}  // namespace pipe
}  // namespace mgw
