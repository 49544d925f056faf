// K2 as a PERSISTENT, WARP-SPECIALISED PIPELINE (the fast path of the multi-grid warp forward on sm_100a).
//
// grid = 3 CTAs per SM, each looping over output tiles (TH x TW pixels inside ONE mesh cell => one homography).
//   producer warp : runs S-1 tiles ahead.  Prepares 8 tiles per round with 4 lanes per tile (decode, homography fetch,
//                   projection of the 4 tile corners, source bounding box) into a ring of records the consumers read, and
//                   issues the TMA load of each tile's box of U into stage s (cp.async.bulk.tensor, SASS UTMALDG) against
//                   the stage's `full` mbarrier (its expect_tx arrive is the release that publishes the record).
//   consumer warps: wait on full[s] (acquire) and work phase by phase over their K pixels with no branch inside a phase:
//                   projective map (two correctly rounded divisions sharing one reciprocal) -> x/y maps and mask stored
//                   directly (8 / 4 B per lane, coalesced) -> taps -> bilinear gather from the staged box.  They hand the
//                   stage back through `empty[s]` (one arrive per warp); the output tile is staged in one of two buffers
//                   and written by a TMA store (UTMASTG) that overlaps the next tile.
// A tile is COMPLETE when its projected corners are finite with z of one sign and the (clipped) tap range fits the
// staged box: a projective map without a pole inside the tile sends it into the convex hull of its corner images, so
// every tap of every pixel is then inside the box.  Complete tiles run one path whose common case (no tap clipped) has no
// clipping arithmetic; border pixels take a short side branch.  Everything else (folded cells, poles, extreme magnification,
// NaN) runs an out-of-line general per-pixel routine that tests every tap against the box and reads global memory for what
// is outside, so results never depend on the box heuristic.  Arithmetic is mgw_device.cuh's: bit-identical to the generic kernels and the C oracle on the forward.
#include "mgw_pipe.cuh"

#ifdef MGW_PROBE
__device__ unsigned long long g_probe_f[16];
extern "C" __attribute__((visibility("default"))) int mgw_debug_probe_fwd(unsigned long long* out, int reset)
{
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_probe_f, sizeof(g_probe_f));
    if (reset) { unsigned long long z[16] = {}; cudaMemcpyToSymbol(g_probe_f, z, sizeof(z)); }
    return 0;
}
#define FPROBE(i) do { if (threadIdx.x == MGW_PROBE_TID) { const long long t_ = clock64(); atomicAdd(&g_probe_f[i], (unsigned long long)(t_ - tprev)); tprev = t_; } } while (0)
#else
#define FPROBE(i) do {} while (0)
#endif
#ifndef MGW_PROBE_TID
#define MGW_PROBE_TID 0
#endif

namespace mgw {

using namespace pipe;

namespace {

// General per-pixel routine of the tiles that are not complete (box too small for the tile's tap range, folded cells,
// poles, NaN): out of line, every tap tested against the box, global memory for what is outside.
template <class G, int C>
__device__ __noinline__ void pixel_general_fwd(const float* __restrict__ Un, const float* src, int bx0, int by0, int IH, int IW,
                                               float xn, float yn, float* o)
{
    const Taps t = make_taps(xn, yn, IH, IW);
    const int sx0 = t.x0 - bx0, sx1 = t.x1 - bx0, sy0 = t.y0 - by0, sy1 = t.y1 - by0;
    if (sx0 >= 0 && sx1 < G::SBW && sy0 >= 0 && sy1 < G::SBH) {
        const float* pa = src + sy0 * G::kRowF + sx0 * C;
        const float* pb = src + sy1 * G::kRowF + sx0 * C;
        const float* pc = src + sy0 * G::kRowF + sx1 * C;
        const float* pd = src + sy1 * G::kRowF + sx1 * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) o[ch] = blend(t, pa[ch], pb[ch], pc[ch], pd[ch]);
    } else {
        const float* pa = Un + ((size_t)t.y0 * IW + t.x0) * C;
        const float* pb = Un + ((size_t)t.y1 * IW + t.x0) * C;
        const float* pc = Un + ((size_t)t.y0 * IW + t.x1) * C;
        const float* pd = Un + ((size_t)t.y1 * IW + t.x1) * C;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) o[ch] = blend(t, __ldg(pa + ch), __ldg(pb + ch), __ldg(pc + ch), __ldg(pd + ch));
    }
}

// ------------------------------------------------------------------------------------------------ forward
template <int C, int TW, int K, int NC, int S, int BW, int BH>
struct FwdLayout {
    static constexpr int TH = (NC / TW) * K;
    using G = PGeo<C, TW, TH, BW, BH>;
    static constexpr size_t kOut = (size_t)S * G::kBoxF * 4;
    static constexpr size_t kBar = kOut + 2 * (size_t)G::kOutF * 4;
    static constexpr size_t kInfo = kBar + 128;
    static constexpr size_t kTotal = kInfo + (size_t)kInfoRing * sizeof(PInfo);
    static_assert((2 * S + 2) * 8 <= 128, "barriers fit their slot");
    // round r+1 is written (over round r-1's slots) once the consumers have released tile 8r+3-S: they are past round r-1
    static_assert(kInfoRing == 2 * kRoundTiles && S <= kRoundTiles / 2, "a record must outlive its tile");
};

#ifndef MGW_PRODUCER_HINT
#define MGW_PRODUCER_HINT 2000      // ns the producer may sleep between two looks at `empty` (0 = plain try_wait loop)
#endif
#ifndef MGW_PIPE_FWD_MINB
#define MGW_PIPE_FWD_MINB 3
#endif
template <int C, int TW, int K, int NC, int S, int BW, int BH>
__global__ void __launch_bounds__(NC + 32, MGW_PIPE_FWD_MINB)
warp_fwd_pipe_kernel(const __grid_constant__ CUtensorMap mapU, const __grid_constant__ CUtensorMap mapOut,
                     const float* __restrict__ U, const float* __restrict__ Hs, const __grid_constant__ PipeCfg cfg,
                     float* __restrict__ img, float* __restrict__ black)
{
    using L = FwdLayout<C, TW, K, NC, S, BW, BH>;
    using G = typename L::G;
    constexpr int TH = L::TH, NCW = NC / 32;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_src = reinterpret_cast<float*>(smem_raw);
    float* s_out = reinterpret_cast<float*>(smem_raw + L::kOut);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L::kBar);
    uint64_t* empty = full + S;
    uint64_t* recbar = empty + S;                    // records of round r are published through recbar[r & 1]
    PInfo* info = reinterpret_cast<PInfo*>(smem_raw + L::kInfo);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = cfg.t.H, W = cfg.t.W;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) { tma::mbar_init(full + s, 1); tma::mbar_init(empty + s, NCW); }
        tma::mbar_init(recbar, 1); tma::mbar_init(recbar + 1, 1);
        tma::fence_barrier_init();
        tma::prefetch_map(&mapU); tma::prefetch_map(&mapOut);
    }
    __syncthreads();
    // programmatic dependent launch: everything above overlapped the previous kernel of the stream (K1 when called through
    // mgw_mesh_warp_fwd); nothing below may touch global memory before that kernel has completed
    griddep_wait();
    griddep_launch_dependents();          // a programmatically launched successor (the backward pipeline) may set itself up under our tail
    const float stepx = lin_step(W), stepy = lin_step(H);

    if (warp == NCW) {
        // ------------------------------------------------------------ producer
        // Records run half a round ahead of the loads (round r+1 is prepared at tile 8r+4) and are published through their
        // own barrier, so that the consumers can run the phases that do not need the source box while it is still in flight.
        for (int it = 0, t = blockIdx.x; t < cfg.total; ++it, t += gridDim.x) {
            if (it == 0) {
                prepare_round<G, TW, TH, C>(cfg, Hs, 0, t, stepx, stepy, info, lane);
                if (lane == 0) tma::mbar_arrive(recbar);
            }
            if (it % kRoundTiles == kRoundTiles / 2) {
                const long long tn = (long long)t + (long long)(kRoundTiles / 2) * gridDim.x;
                if (tn < cfg.total) {
                    const int itn = it + kRoundTiles / 2;
                    prepare_round<G, TW, TH, C>(cfg, Hs, itn, (int)tn, stepx, stepy, info, lane);
                    if (lane == 0) tma::mbar_arrive(recbar + ((itn / kRoundTiles) & 1));
                }
            }
            if (lane == 0) {
                const int s = it % S;
                const PInfo* in = info + (it % kInfoRing);
                #if MGW_PRODUCER_HINT > 0
                tma::mbar_wait_hint(empty + s, ((it / S) & 1) ^ 1, MGW_PRODUCER_HINT);
#else
                tma::mbar_wait(empty + s, ((it / S) & 1) ^ 1);
#endif
                tma::mbar_expect_tx(full + s, (uint32_t)(G::kBoxF * 4));
                tma::load_3d(s_src + (size_t)s * G::kBoxF, &mapU, full + s, in->bx0 * C, in->by0, in->n);
            }
        }
        return;
    }

    // ---------------------------------------------------------------- consumers
    const int tx = tid % TW, g = tid / TW;
    const unsigned char* sbase = smem_raw;
#ifdef MGW_PROBE
    long long tprev = clock64();
#endif
    for (int it = 0, t = blockIdx.x; t < cfg.total; ++it, t += gridDim.x) {
        const int s = it % S;
        if (it % kRoundTiles == 0) tma::mbar_wait(recbar + ((it / kRoundTiles) & 1), (it / (2 * kRoundTiles)) & 1);
        FPROBE(0);     // wait for the records
#ifdef MGW_PROBE
        if (tid == MGW_PROBE_TID) atomicAdd(&g_probe_f[15], 1ull);
#endif
        const PInfo* in = info + (it % kInfoRing);
        const int n = in->n, r0 = in->r0, c0 = in->c0, complete = in->complete, bx0 = in->bx0, by0 = in->by0;
        const int offbase = -(by0 * G::kRowF + bx0 * C) + s * G::kBoxF;
        float Hc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Hc[k] = in->Hc[k];
        float* so = s_out + (it & 1) * G::kOutF + (g * K * TW + tx) * C;
        const float xt = lin_at(c0 + tx, stepx);
        const float hx0 = __fmul_rn(Hc[0], xt), hx3 = __fmul_rn(Hc[3], xt), hx6 = __fmul_rn(Hc[6], xt);   // first term of hrow()
        const int pix0 = (n * H + r0 + g * K) * W + c0 + tx;
        float2* pimg = reinterpret_cast<float2*>(img) + pix0;
        float* pblk = black + pix0;
        if (complete) {
            // phase 1: projective map of the K pixels
            float xn[K], yn[K];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float yt = lin_at(r0 + g * K + k, stepy);
                const float xs = __fadd_rn(__fmaf_rn(Hc[1], yt, hx0), Hc[2]);
                const float ys = __fadd_rn(__fmaf_rn(Hc[4], yt, hx3), Hc[5]);
                float zs = __fadd_rn(__fmaf_rn(Hc[7], yt, hx6), Hc[8]);
                zs = __fadd_rn(zs, (zs >= 0.0f) ? 1e-8f : -1e-8f);
                div2_tile(xs, ys, zs, xn[k], yn[k], bad);
            }
            if (bad) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const Proj q = project(Hc, xt, lin_at(r0 + g * K + k, stepy));      // IEEE divisions
                    xn[k] = q.xn; yn[k] = q.yn;
                }
            }
            FPROBE(1);     // phase 1 (record read + projective map)
            // phase 2: x_map,y_map and black_pix: 8 / 4 contiguous bytes per lane, plain coalesced stores
#pragma unroll
            for (int k = 0; k < K; ++k) {
                pimg[k * W] = make_float2(xn[k], yn[k]);
                pblk[k * W] = black_of(xn[k], yn[k]);
            }
            FPROBE(2);     // phase 2 (map / mask stores)
            // phase 3: taps
            PixTaps tp[K];
            bool clip[K], anyclip = false;
#pragma unroll
            for (int k = 0; k < K; ++k) { clip[k] = pix_taps<C, G::kRowF>(xn[k], yn[k], H, W, offbase, tp[k]); anyclip = anyclip || clip[k]; }
            if (anyclip) {
#pragma unroll
                for (int k = 0; k < K; ++k)
                    if (clip[k]) pix_taps_clipped<C, G::kRowF>(xn[k], yn[k], H, W, offbase, tp[k]);
            }
            FPROBE(3);     // phase 3 (taps)
            // phase 4: bilinear gather from the staged box (the only phase that needs it)
            tma::mbar_wait(full + s, (it / S) & 1);
            FPROBE(4);     // wait for the box
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const unsigned char* pa = sbase + tp[k].off;
                const unsigned char* pb = pa + tp[k].dy;
                const unsigned char* pc = pa + tp[k].dx;
                const unsigned char* pd = pb + tp[k].dx;
#pragma unroll
                for (int ch = 0; ch < C; ++ch)
                    so[k * TW * C + ch] = blend4(tp[k].ax, tp[k].bx, tp[k].ay, tp[k].by, reinterpret_cast<const float*>(pa)[ch],
                                                 reinterpret_cast<const float*>(pb)[ch], reinterpret_cast<const float*>(pc)[ch],
                                                 reinterpret_cast<const float*>(pd)[ch]);
            }
        } else {
            const float* Un = U + (size_t)n * H * W * C;
            const float* src = s_src + (size_t)s * G::kBoxF;
            tma::mbar_wait(full + s, (it / S) & 1);
#pragma unroll 1
            for (int k = 0; k < K; ++k) {
                const Proj q = project(Hc, xt, lin_at(r0 + g * K + k, stepy));
                pimg[k * W] = make_float2(q.xn, q.yn);
                pblk[k * W] = black_of(q.xn, q.yn);
                pixel_general_fwd<G, C>(Un, src, bx0, by0, H, W, q.xn, q.yn, so + k * TW * C);
            }
        }
        FPROBE(5);     // phase 4 (gather) / general path
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(empty + s);               // the source box may be refilled
        tma::fence_proxy_async();                                 // my part of the output tile -> visible to the TMA store
        if (tid == 0) tma::wait_group_read0();                    // the previous tile's store has left ITS buffer (the next one's)
        FPROBE(6);     // release, proxy fence, wait for the previous tile's store
        tma::named_bar_sync<1, NC>();
        FPROBE(7);     // consumer barrier
        if (tid == 0) {
            tma::store_3d(&mapOut, s_out + (it & 1) * G::kOutF, c0 * C, r0, n);
            tma::commit_group();
        }
        FPROBE(8);     // store issue
    }
    if (tid == 0) tma::wait_group_read0();
}

template <int C, int TW, int K, int NC, int S, int BW, int BH>
static int launch_fwd_v(const float* U, const float* Hs, const PipePlan& p, float* out, float* black, float* img, cudaStream_t st)
{
    using L = FwdLayout<C, TW, K, NC, S, BW, BH>;
    using G = typename L::G;
    const TileCfg& c = p.cfg.t;
    CUtensorMap mU, mOut;
    TRY_RC(make_map(&mU, U, c.W * C, c.H, c.N, G::kRowF, G::SBH));
    TRY_RC(make_map(&mOut, out, c.W * C, c.H, c.N, TW * C, L::TH));
    static bool attr[64] = {};
    TRY_RC(allow_smem(warp_fwd_pipe_kernel<C, TW, K, NC, S, BW, BH>, attr, "warp_fwd_pipe"));
    const cudaError_t e = launch_ex(warp_fwd_pipe_kernel<C, TW, K, NC, S, BW, BH>, dim3(grid_for(p.cfg.total, MGW_PIPE_FWD_MINB)), dim3(NC + 32),
                                    L::kTotal, st, pdl_enabled(), mU, mOut, U, Hs, p.cfg, img, black);
    if (e != cudaSuccess) { count_launches(1); return set_error(MGW_ERR_CUDA, "warp_fwd_pipe: %s", cudaGetErrorString(e)); }
    return check_launch("warp_fwd_pipe");
}

// compiled variants: (TW, K, consumer threads, stages, box width px, box height px); TH = K * threads / TW.
// "big" serves cells of at least 24 / 18 rows (the benchmark's 72 x 128 cells: 3 / 4 tiles per cell height), "small" the rest.
// (nvcc splits -D values at commas, so the tuning overrides are per field: -DMGW_PF_K=2 -DMGW_PF_NC=384 ...)
#ifndef MGW_PF_K
#define MGW_PF_K 3
#endif
#ifndef MGW_PF_NC
#define MGW_PF_NC 256
#endif
#ifndef MGW_PF_S
#define MGW_PF_S 2
#endif
#ifndef MGW_PF_BH
#define MGW_PF_BH 36
#endif
#define MGW_PIPE_FWD_BIG 32, MGW_PF_K, MGW_PF_NC, MGW_PF_S, 64, MGW_PF_BH
#ifndef MGW_PIPE_FWD_SMALL
#define MGW_PIPE_FWD_SMALL 32, 1, 256, 3, 64, 20
#endif

template <int TW, int K, int NC, int S, int BW, int BH> struct Variant { static constexpr int tw = TW, th = (NC / TW) * K; };
using VFB = Variant<MGW_PIPE_FWD_BIG>;
using VFS = Variant<MGW_PIPE_FWD_SMALL>;

template <class VB, class VS>
static bool plan2(const WarpShape& s, PipePlan* p, bool* big)
{
    if (s.C != 1 && s.C != 3 && s.C != 4) return false;
    const double eb = tile_eff(s, VB::tw, VB::th), es = tile_eff(s, VS::tw, VS::th) * 0.85;      // larger tiles amortise the halo
    if (eb <= 0 && es <= 0) return false;
    *big = eb >= es;
    return *big ? plan(s, VB::tw, VB::th, p) : plan(s, VS::tw, VS::th, p);
}

}  // namespace

bool pipe_fwd_supported(const WarpShape& s) { PipePlan p; bool b; return plan2<VFB, VFS>(s, &p, &b); }

int launch_warp_fwd_pipe(const float* U, const float* Hs, const WarpShape& s, float* out, float* black, float* img, cudaStream_t st)
{
    PipePlan p; bool big;
    if (!out || !plan2<VFB, VFS>(s, &p, &big)) return set_error(MGW_ERR_UNSUPPORTED, "warp_fwd_pipe: unsupported shape");
    if (big) {
        if (s.C == 1) return launch_fwd_v<1, MGW_PIPE_FWD_BIG>(U, Hs, p, out, black, img, st);
        if (s.C == 3) return launch_fwd_v<3, MGW_PIPE_FWD_BIG>(U, Hs, p, out, black, img, st);
        if (s.C == 4) return launch_fwd_v<4, MGW_PIPE_FWD_BIG>(U, Hs, p, out, black, img, st);
    } else {
        if (s.C == 1) return launch_fwd_v<1, MGW_PIPE_FWD_SMALL>(U, Hs, p, out, black, img, st);
        if (s.C == 3) return launch_fwd_v<3, MGW_PIPE_FWD_SMALL>(U, Hs, p, out, black, img, st);
        if (s.C == 4) return launch_fwd_v<4, MGW_PIPE_FWD_SMALL>(U, Hs, p, out, black, img, st);
    }
    return set_error(MGW_ERR_UNSUPPORTED, "warp_fwd_pipe: no variant");
}

}  // namespace mgw
