// K3 as a PERSISTENT TMA PIPELINE: backward of the multi-grid warp (dU scatter + dH reduction) on sm_100a.
// Reference: autodiff of spatial_transformer3.py:108-122 (gather backward), :248-260 (projective map); SURVEY.md 8a-bwd.
//
// grid = 2 CTAs per SM x 8 warps (128 registers per thread: 4 warps per SM sub-partition), each CTA walking CHUNKS of
// vertically adjacent output tiles (TH x TW pixels inside ONE mesh cell => one homography).  There is no producer warp (a 9th
// warp would cost every thread 32 registers): the per-tile bookkeeping is spread over the 8 warps.
//   - tile records (decode, homography, projected corners, source box) are prepared half a round (8 tiles) ahead, one record
//     per warp, and published through an mbarrier (mgw_pipe.cuh);
//   - the source box of U for tile i+2 is requested by ONE thread right after the CTA barrier of tile i, into the stage tile i
//     has just finished with (cp.async.bulk.tensor, SASS UTMALDG, completion on the stage's `full` mbarrier);
//   - the upstream gradients (d_out, d_img: 12 + 8 contiguous bytes per lane) of tile i+1 are loaded into REGISTERS while
//     tile i is processed: no global-load latency is exposed and nothing but U goes through shared memory;
//   - projective map and taps run before the wait on `full` (they do not need the box);
//   - dU is pre-accumulated in a shared-memory box in FIXED POINT with native integer shared atomics (ATOMS.ADD): the quantum
//     is 2^-21 of the tile's max|d_out| (published per warp one tile ahead, so the scale costs no barrier of its own).  A pixel
//     adds at most one full-size term to a word and a tile has TH*TW <= 1024 pixels, so a word can never overflow -- no
//     magnification heuristic;
//   - two accumulators alternate: after ONE CTA barrier per tile the part of the box that can hold taps is converted back to
//     fp32, re-zeroed and sent to dU with coalesced 16-byte reductions (red.global.add.v4.f32) while the atomics of the next
//     tile already go to the other accumulator;
//   - the 8 dH terms stay in registers over a chunk, then warp shuffle -> shared -> one deterministic partial per chunk.
// A tile that is not COMPLETE (box too small for its tap range, folded cells, poles, NaN) or whose d_out holds Inf/NaN runs an
// out-of-line per-pixel routine that tests every tap against the box and uses global fp32 atomics for what is outside, so
// results never depend on the box heuristic.  Pixels with a clipped tap contribute only their d_img term (their dU terms
// cancel exactly and gx = gy = 0 up to the reference's rounding residue; see mgw_device.cuh, taps_scatter).
#include "mgw_pipe.cuh"

namespace mgw {

using namespace pipe;

namespace {

// float -> fixed point without the conversion unit: adding 1.5*2^23 leaves round-to-nearest-even(v) in the low mantissa bits
constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23
constexpr int kMagicBits = 0x4B400000;
constexpr int kFixedBits = 21;                 // |w * g * scale| <= 2^21 per term, <= 1024 terms per word: |sum| <= 2^31
__device__ __forceinline__ int fixed_of(float w, float gs) { return __float_as_int(__fmaf_rn(w, gs, kMagic)) - kMagicBits; }

// dH terms of one pixel with the reciprocal of zs at hand
__device__ __forceinline__ void accumulate_dh_r(float (&dh)[8], float gxn, float gyn, float xn, float yn, float rz, float xt, float yt)
{
    const float dxs = gxn * rz, dys = gyn * rz;
    const float dzs = -(gxn * xn + gyn * yn) * rz;
    dh[0] = fmaf(dxs, xt, dh[0]); dh[1] = fmaf(dxs, yt, dh[1]); dh[2] += dxs;
    dh[3] = fmaf(dys, xt, dh[3]); dh[4] = fmaf(dys, yt, dh[4]); dh[5] += dys;
    dh[6] = fmaf(dzs, xt, dh[6]); dh[7] = fmaf(dzs, yt, dh[7]);
}

// upstream gradient of the fused img_loss (s_net_bundle_nobm.py:347-352): d_out = kn[n] * (out - y) * (1-black)^2
struct LossSrc {
    const float* out;
    const float* y;
    const float* black;
    const float* sums;          // [N,2] from the fused forward
    float kscale;               // upstream * 2 / batch
    const float* kscale_dev;    // nullable device factor on kscale
};

// General per-pixel routine (out of line): every tap tested against the box; fixed-point shared atomics for taps inside it
// when the tile has a scale, global fp32 atomics otherwise.  Returns (gx, gy).
template <class G, int C>
__device__ __noinline__ float2 pixel_general_bwd(const float* __restrict__ Un, float* __restrict__ dUn, const float* src, int* acc,
                                                 int bx0, int by0, int IH, int IW, float xn, float yn, float g0, float g1, float g2,
                                                 float g3, int fixed, float scale)
{
    const Taps t = make_taps(xn, yn, IH, IW);
    const int sx0 = t.x0 - bx0, sx1 = t.x1 - bx0, sy0 = t.y0 - by0, sy1 = t.y1 - by0;
    const bool inbox = sx0 >= 0 && sx1 < G::SBW && sy0 >= 0 && sy1 < G::SBH;
    const float wa = t.ax * t.ay, wb = t.ax * t.by, wc = t.bx * t.ay, wd = t.bx * t.by;
    const bool scatter = taps_scatter(t);
    const size_t ga = ((size_t)t.y0 * IW + t.x0) * C, gb = ((size_t)t.y1 * IW + t.x0) * C;
    const size_t gc = ((size_t)t.y0 * IW + t.x1) * C, gd = ((size_t)t.y1 * IW + t.x1) * C;
    const int ia = sy0 * G::kRowF + sx0 * C, ib = sy1 * G::kRowF + sx0 * C;
    const int ic = sy0 * G::kRowF + sx1 * C, id = sy1 * G::kRowF + sx1 * C;
    const float gg[4] = {g0, g1, g2, g3};
    float gx = 0.0f, gy = 0.0f;
#pragma unroll
    for (int ch = 0; ch < C; ++ch) {
        const float gch = gg[ch];
        float Ia, Ib, Ic, Id;
        if (inbox) { Ia = src[ia + ch]; Ib = src[ib + ch]; Ic = src[ic + ch]; Id = src[id + ch]; }
        else { Ia = __ldg(Un + ga + ch); Ib = __ldg(Un + gb + ch); Ic = __ldg(Un + gc + ch); Id = __ldg(Un + gd + ch); }
        gx = fmaf(gch, fmaf(Ic - Ia, t.ay, (Id - Ib) * t.by), gx);
        gy = fmaf(gch, fmaf(Ib - Ia, t.ax, (Id - Ic) * t.bx), gy);
        if (scatter) {
            if (inbox && fixed) {          // unclipped taps have weights in [0,1], so |w*g| <= max|d_out|
                const float gs = gch * scale;
                atomicAdd(acc + ia + ch, fixed_of(wa, gs));
                atomicAdd(acc + ib + ch, fixed_of(wb, gs));
                atomicAdd(acc + ic + ch, fixed_of(wc, gs));
                atomicAdd(acc + id + ch, fixed_of(wd, gs));
            } else {
                atomicAdd(dUn + ga + ch, wa * gch);
                atomicAdd(dUn + gb + ch, wb * gch);
                atomicAdd(dUn + gc + ch, wc * gch);
                atomicAdd(dUn + gd + ch, wd * gch);
            }
        }
    }
    return make_float2(gx, gy);
}

template <int C, int TW, int K, int NT, int S, int BW, int BH>
struct BwdLayout {
    static constexpr int TH = (NT / TW) * K, NW = NT / 32;
    using G = PGeo<C, TW, TH, BW, BH>;
    static constexpr size_t kAcc = (size_t)S * G::kBoxF * 4;                 // two fixed-point accumulators after the stages
    static constexpr size_t kBar = kAcc + 2 * (size_t)G::kBoxF * 4;          // full[S], recbar[2]
    static constexpr size_t kRed = kBar + 128;                               // [NW][8] dH partials
    static constexpr size_t kMax = kRed + NW * 8 * 4;                        // [2][NW] per-warp max|d_out| (bit patterns)
    static constexpr size_t kInfo = (kMax + 2 * NW * 4 + 127) / 128 * 128;
    static constexpr size_t kTotal = kInfo + (size_t)kInfoRing * sizeof(PInfo);
    static_assert((S + 2) * 8 <= 128, "barriers fit their slot");
    static_assert(kInfoRing == 2 * kRoundTiles && NW == kRoundTiles, "one record per warp and round");
    static_assert(S == 2, "the stage a tile releases is refilled with tile + 2");
    static_assert(TH * TW <= 1024, "fixed-point headroom: at most 1024 full-size terms per word");
};

// flat tile index (n, ty, tx) of the CTA's j-th tile (j < the CTA's tile count).  Chunks are dealt round-robin to the CTAs.
// Integer divisions by run-time values: only the warp that prepares a record pays them, once per tile.
__device__ __forceinline__ int tile_of(const PipeCfg& cfg, int j)
{
    const int c = (int)blockIdx.x + (j / cfg.chunk_L) * (int)gridDim.x, tx = c % cfg.ntx, q = c / cfg.ntx;
    const int ty = (q % cfg.chunk_rows) * cfg.chunk_L + j % cfg.chunk_L, n = q / cfg.chunk_rows;
    return (n * cfg.nty + ty) * cfg.ntx + tx;
}

#ifndef MGW_PIPE_BWD_MINB
#define MGW_PIPE_BWD_MINB 2
#endif

// records of round r (the CTA's tiles 8r .. 8r+7): warp w prepares tile 8r + w (4 lanes project the corners)
template <class G, int TW, int TH, int C>
__device__ __forceinline__ void prepare_round_by_warps(const PipeCfg& cfg, const float* __restrict__ Hs, int r, int nj, PInfo* info, uint64_t* recbar)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int j = r * kRoundTiles + warp;
    const bool valid = j < nj;
    make_record<G, TW, TH, C>(cfg, Hs, valid ? tile_of(cfg, j) : 0, valid, lin_step(cfg.t.W), lin_step(cfg.t.H), info + (j % kInfoRing), lane, lane == 0);
    __syncwarp();
    if (lane == 0) tma::mbar_arrive(recbar + (r & 1));
}

template <int C, int TW, int K, int NT, int S, int BW, int BH, bool LOSS>
__global__ void __launch_bounds__(NT, MGW_PIPE_BWD_MINB)
warp_bwd_pipe_kernel(const __grid_constant__ CUtensorMap mapU, const float* __restrict__ U, const float* __restrict__ Hs,
                     const float* __restrict__ d_out, const float* __restrict__ d_img, const __grid_constant__ PipeCfg cfg,
                     float* __restrict__ dU, float* __restrict__ parts, const LossSrc loss)
{
    using L = BwdLayout<C, TW, K, NT, S, BW, BH>;
    using G = typename L::G;
    constexpr int TH = L::TH, NW = L::NW;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_src = reinterpret_cast<float*>(smem_raw);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem_raw + L::kBar);
    uint64_t* recbar = full + S;
    float* s_red = reinterpret_cast<float*>(smem_raw + L::kRed);
    unsigned* s_max = reinterpret_cast<unsigned*>(smem_raw + L::kMax);
    PInfo* info = reinterpret_cast<PInfo*>(smem_raw + L::kInfo);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int H = cfg.t.H, W = cfg.t.W;
    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < S; ++s) tma::mbar_init(full + s, 1);
        tma::mbar_init(recbar, NW); tma::mbar_init(recbar + 1, NW);
        tma::fence_barrier_init();
        tma::prefetch_map(&mapU);
    }
    // both accumulators start out zero and every drain leaves what it visited zero again
    {
        int4* a4 = reinterpret_cast<int4*>(smem_raw + L::kAcc);
        for (int i = tid; i < 2 * G::kBoxF / 4; i += NT) a4[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
    // programmatic dependent launch: the set-up above overlapped the tail of the previous kernel of the stream; nothing below
    // touches global memory before that kernel has completed.  K4 may be scheduled from now on: its factorisation overlaps us.
    griddep_wait();
    griddep_launch_dependents();
    const float stepx = lin_step(W), stepy = lin_step(H);
    const int tx = tid % TW, g = tid / TW;
    // tiles of this CTA: its chunks (dealt round-robin) times the tiles per chunk
    const int nj = (int)blockIdx.x < cfg.nchunks ? ((cfg.nchunks - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * cfg.chunk_L : 0;

    // ---- bookkeeping spread over the warps
    auto wait_round = [&](int r) { tma::mbar_wait(recbar + (r & 1), (r >> 1) & 1); };
    // source box of the CTA's j-th tile -> stage j % S (one thread)
    auto request_box = [&](int j) {
        if (j >= nj) return;
        if (j % kRoundTiles == 0) wait_round(j / kRoundTiles);
        const PInfo* in = info + (j % kInfoRing);
        const int s = j % S;
        tma::mbar_expect_tx(full + s, (uint32_t)(G::kBoxF * 4));
        tma::load_3d(s_src + (size_t)s * G::kBoxF, &mapU, full + s, in->bx0 * C, in->by0, in->n);
    };

    // ---- upstream gradients of this thread's K pixels: `gcur` of the tile being processed, raw loads of the next one
    float gcur[K][C], gicur[K][2];
    float na[K][C], nb[LOSS ? K : 1][LOSS ? C : 1], nbk[LOSS ? K : 1];
    // issue the loads of a tile's upstream gradients (pixels the tile does not own are never read: they count as zero)
    // (pixel indices fit 32 bits -- N*H*W < 2^31 is validated by the C ABI -- so every address is one IMAD.WIDE off a base)
    auto load_next = [&](const PInfo* in) {
        const int row0 = in->r0 + g * K, col = in->c0 + tx;
        const int kfirst = (col >= in->vc0) ? in->vr0 - row0 : K;      // first owned pixel of the thread (<= 0: all of them)
        const int p0 = (in->n * H + row0) * W + col;
        const float* gsrc = LOSS ? loss.out : d_out;
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const bool own = k >= kfirst;
            const int p = p0 + k * W;
            const float* po = gsrc + (size_t)(unsigned)p * C;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) na[k][ch] = own ? __ldg(po + ch) : 0.0f;
            if constexpr (LOSS) {
                nbk[k] = own ? 1.0f - __ldg(loss.black + (unsigned)p) : 0.0f;
                const float* py = loss.y + (size_t)(unsigned)p * C;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) nb[k][ch] = own ? __ldg(py + ch) : 0.0f;
            }
            // d_img of the next tile only travels to L2 now (a warp's row is 256 contiguous bytes: two lines) and is loaded at
            // the top of its own tile: 6 registers less across the gather / scatter phase, where ptxas would spill them
            if (d_img != nullptr && (lane & 15) == 0)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const float2*>(d_img) + (unsigned)p));
        }
        if constexpr (LOSS) {      // the per-sample factor rides in nbk: kn * (1-black)^2
            const float kn = (loss.kscale_dev ? loss.kscale * __ldg(loss.kscale_dev) : loss.kscale) / (__ldg(loss.sums + 2 * in->n + 1) + 1e-8f);
#pragma unroll
            for (int k = 0; k < K; ++k) nbk[k] = kn * nbk[k] * nbk[k];
        }
    };
    // raw loads -> gradients in `gcur`; publishes this warp's max|d_out| (bit patterns: Inf/NaN win) for the tile in slot `b`
    auto adopt_next = [&](int b) {
        unsigned m = 0u;
#pragma unroll
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int ch = 0; ch < C; ++ch) {
                if constexpr (LOSS) gcur[k][ch] = nbk[k] * (na[k][ch] - nb[k][ch]); else gcur[k][ch] = na[k][ch];
                m = max(m, (unsigned)__float_as_int(gcur[k][ch]) & 0x7fffffffu);
            }
        }
        m = __reduce_max_sync(0xffffffffu, m);
        if (lane == 0) s_max[b * NW + warp] = m;
    };

    if (nj > 0) {
        prepare_round_by_warps<G, TW, TH, C>(cfg, Hs, 0, nj, info, recbar);
        if (nj > kRoundTiles) prepare_round_by_warps<G, TW, TH, C>(cfg, Hs, 1, nj, info, recbar);
        if (tid == 0) { request_box(0); request_box(1); }
        wait_round(0);
        load_next(info);
        adopt_next(0);
    }
    __syncthreads();

    float dh[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) dh[k] = 0.0f;
    int pj = 0;                                                    // j % chunk_L, tracked without dividing
    for (int j = 0; j < nj; ++j) {
        const int s = j % S, b = j & 1;
        const PInfo* in = info + (j % kInfoRing);
        // records of the round after the next one are due in 5 tiles: every warp prepares its one now
        if (j % kRoundTiles == 3 && j >= kRoundTiles && (j / kRoundTiles + 1) * kRoundTiles < nj) prepare_round_by_warps<G, TW, TH, C>(cfg, Hs, j / kRoundTiles + 1, nj, info, recbar);
        const bool has_next = j + 1 < nj;
        if (has_next) {
            if ((j + 1) % kRoundTiles == 0) wait_round((j + 1) / kRoundTiles);
            load_next(info + ((j + 1) % kInfoRing));
        }
        const int r0 = in->r0, c0 = in->c0, complete = in->complete, bx0 = in->bx0, by0 = in->by0;
        {       // d_img of this tile (L2-resident since the previous tile; consumed at the end of each pixel)
            const int row0 = r0 + g * K, col = c0 + tx;
            const int kfirst = (col >= in->vc0) ? in->vr0 - row0 : K;
            const int p0 = (in->n * H + row0) * W + col;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float2 di = make_float2(0.0f, 0.0f);
                if (d_img != nullptr && k >= kfirst) di = __ldg(reinterpret_cast<const float2*>(d_img) + (unsigned)(p0 + k * W));
                gicur[k][0] = di.x; gicur[k][1] = di.y;
            }
        }
        float Hc[9];
#pragma unroll
        for (int k = 0; k < 9; ++k) Hc[k] = in->Hc[k];
        // fixed-point scale of the tile from the per-warp maxima published one tile ago
        unsigned mb = 0u;
#pragma unroll
        for (int w = 0; w < NW; ++w) mb = max(mb, s_max[b * NW + w]);
        const int e = (int)(mb >> 23) - 127;                                 // floor(log2 max|d_out|); 128 for Inf/NaN
        const bool allzero = mb == 0u;
        const bool fixed = allzero || (e > -100 && e < 100);
        const float scale = allzero ? 0.0f : __int_as_float((kFixedBits - 1 - e + 127) << 23);      // |g| * scale < 2^kFixedBits

        const int col = c0 + tx, row0 = r0 + g * K;
        const float xt = lin_at(col, stepx);
        const float hx0 = __fmul_rn(Hc[0], xt), hx3 = __fmul_rn(Hc[3], xt), hx6 = __fmul_rn(Hc[6], xt);   // first term of hrow()
        const float halfW = 0.5f * (float)W, halfH = 0.5f * (float)H;
        unsigned char* abase = smem_raw + L::kAcc + (size_t)b * G::kBoxF * 4;

        if (complete && fixed) {
            // phase 1: projective map of the K pixels
            float xn[K], yn[K], rz[K];
            bool bad = false;
#pragma unroll
            for (int k = 0; k < K; ++k) {
                const float yt = lin_at(row0 + k, stepy);
                const float xs = __fadd_rn(__fmaf_rn(Hc[1], yt, hx0), Hc[2]);
                const float ys = __fadd_rn(__fmaf_rn(Hc[4], yt, hx3), Hc[5]);
                float zs = __fadd_rn(__fmaf_rn(Hc[7], yt, hx6), Hc[8]);
                zs = __fadd_rn(zs, (zs >= 0.0f) ? 1e-8f : -1e-8f);
                rz[k] = div2_tile(xs, ys, zs, xn[k], yn[k], bad);
            }
            if (bad) {
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const Proj q = project(Hc, xt, lin_at(row0 + k, stepy));      // IEEE divisions
                    xn[k] = q.xn; yn[k] = q.yn; rz[k] = __frcp_rn(q.zs);
                }
            }
            // phase 2 (needs the staged box): taps in their interior form (a pixel with a clipped tap skips the gather / scatter
            // altogether), gather, image gradient, fixed-point scatter
            tma::mbar_wait(full + s, (j / S) & 1);
            const unsigned char* sbase = smem_raw + (size_t)s * G::kBoxF * 4;
            const int offbase = -(by0 * G::kRowF + bx0 * C);
#pragma unroll
            for (int k = 0; k < K; ++k) {
                float gx = 0.0f, gy = 0.0f;
                PixTaps tp;
                if (!pix_taps<C, G::kRowF>(xn[k], yn[k], H, W, offbase, tp)) {
                    const float* pa = reinterpret_cast<const float*>(sbase + tp.off);
                    int* qa = reinterpret_cast<int*>(abase + tp.off);
                    const float ax = tp.ax, bx = tp.bx, ay = tp.ay, by = tp.by;
                    const float wa = ax * ay, wb = ax * by, wc = bx * ay, wd = bx * by;
                    // gx = sum_c g_c [(Ic-Ia) ay + (Id-Ib) by], gy = sum_c g_c [(Ib-Ia) ax + (Id-Ic) bx], channel sums per tap first.
                    // (Packed FFMA2 / FMUL2 for these 24 multiply-adds was measured: 14 instructions fewer per pixel, but the
                    // kernel ran 6 % SLOWER -- 95.2 -> 101.4 us -- so everything stays scalar.)
                    float sa = 0.0f, sb = 0.0f, sc = 0.0f, sd = 0.0f;
#pragma unroll
                    for (int ch = 0; ch < C; ++ch) {
                        const float gch = gcur[k][ch];
                        sa = fmaf(gch, pa[ch], sa); sb = fmaf(gch, pa[G::kRowF + ch], sb);
                        sc = fmaf(gch, pa[C + ch], sc); sd = fmaf(gch, pa[G::kRowF + C + ch], sd);
                        const float gs = gch * scale;
                        atomicAdd(qa + ch, fixed_of(wa, gs));
                        atomicAdd(qa + G::kRowF + ch, fixed_of(wb, gs));
                        atomicAdd(qa + C + ch, fixed_of(wc, gs));
                        atomicAdd(qa + G::kRowF + C + ch, fixed_of(wd, gs));
                    }
                    gx = fmaf(sc - sa, ay, (sd - sb) * by); gy = fmaf(sb - sa, ax, (sd - sc) * bx);
                }
                accumulate_dh_r(dh, fmaf(gx, halfW, gicur[k][0]), fmaf(gy, halfH, gicur[k][1]), xn[k], yn[k], rz[k], xt,
                                lin_at(row0 + k, stepy));
            }
        } else {
            const float* Un = U + (size_t)in->n * H * W * C;
            float* dUn = dU + (size_t)in->n * H * W * C;
            const float* src = s_src + (size_t)s * G::kBoxF;
            const bool ownc = col >= in->vc0;
            const int vr0 = in->vr0;
            tma::mbar_wait(full + s, (j / S) & 1);
#pragma unroll
            for (int k = 0; k < K; ++k) {                     // unrolled: gcur / gicur must stay in registers
                if (!(ownc && row0 + k >= vr0)) continue;
                const float yt = lin_at(row0 + k, stepy);
                const Proj q = project(Hc, xt, yt);
                float gk[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
                for (int ch = 0; ch < C; ++ch) gk[ch] = gcur[k][ch];
                const float2 gxy = pixel_general_bwd<G, C>(Un, dUn, src, reinterpret_cast<int*>(abase), bx0, by0, H, W, q.xn, q.yn,
                                                           gk[0], gk[1], gk[2], gk[3], (fixed && !allzero) ? 1 : 0, scale);
                accumulate_dh_r(dh, fmaf(gxy.x, halfW, gicur[k][0]), fmaf(gxy.y, halfH, gicur[k][1]), q.xn, q.yn, __frcp_rn(q.zs), xt, yt);
            }
        }
        // dH: one partial per CHUNK.  Halving butterfly (9 shuffles for the 8 sums) -> shared
        const bool chunk_end = pj == cfg.chunk_L - 1;
        pj = chunk_end ? 0 : pj + 1;
        if (chunk_end) {
            float v4[4], v2[2], v1;
            const bool hi16 = lane & 16, hi8 = lane & 8, hi4 = lane & 4;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float send = hi16 ? dh[i] : dh[i + 4], keep = hi16 ? dh[i + 4] : dh[i];
                v4[i] = keep + __shfl_xor_sync(0xffffffffu, send, 16);
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                const float send = hi8 ? v4[i] : v4[i + 2], keep = hi8 ? v4[i + 2] : v4[i];
                v2[i] = keep + __shfl_xor_sync(0xffffffffu, send, 8);
            }
            {
                const float send = hi4 ? v2[0] : v2[1], keep = hi4 ? v2[1] : v2[0];
                v1 = keep + __shfl_xor_sync(0xffffffffu, send, 4);
            }
            v1 += __shfl_xor_sync(0xffffffffu, v1, 2);
            v1 += __shfl_xor_sync(0xffffffffu, v1, 1);
            // lane l now holds term  4*bit4(l) + 2*bit3(l) + bit2(l)
            if ((lane & 3) == 0) s_red[warp * 8 + ((lane >> 4) & 1) * 4 + ((lane >> 3) & 1) * 2 + ((lane >> 2) & 1)] = v1;
#pragma unroll
            for (int k = 0; k < 8; ++k) dh[k] = 0.0f;
        }
        if (has_next) adopt_next(b ^ 1);                          // the next tile's gradients have had a whole tile to arrive
        __syncthreads();                 // every atomic of this tile has been issued, its source stage is free; s_red / s_max are complete
        if (tid == 0) request_box(j + 2);                         // refill the stage this tile has just released
        if (chunk_end && tid < 8) {
            float v = 0.0f;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += s_red[w * 8 + tid];
            const int nparts = cfg.chunk_L > 1 ? cfg.t.parts_x : cfg.t.parts_y * cfg.t.parts_x;
            parts[((size_t)in->cell * nparts + in->part) * 8 + tid] = v;
        }
        // drain: fixed point -> fp32, re-zero, coalesced 16-byte reductions into dU.  Only the part of the box that can hold taps
        // is visited (nrow x nq 16-byte groups; the whole in-image box for tiles on the general path), all-zero groups are
        // skipped.  (The record is still there: its slot is rewritten three tiles later at the earliest.)
        if (fixed && !allzero) {
            const float inv_scale = __int_as_float((e - (kFixedBits - 1) + 127) << 23);
            int4* a4 = reinterpret_cast<int4*>(abase);
            float* dbox = dU + (size_t)(unsigned)((in->n * H + in->by0) * W + in->bx0) * C;      // one 64-bit base per tile ...
            const int nq = in->nq, items = in->nrow * nq;
            const unsigned pitch = (unsigned)(W * C);                                             // ... 32-bit offsets inside the box
            const float rnq = __frcp_rn((float)nq);
            for (int idx = tid; idx < items; idx += NT) {
                const int r = __float2int_rz(__fmul_rn((float)idx + 0.5f, rnq)), q = idx - r * nq;      // exact: idx < 2^11
                int4* cell = a4 + r * (G::kRowF / 4) + q;
                const int4 v = *cell;
                if ((v.x | v.y | v.z | v.w) != 0) {
                    *cell = make_int4(0, 0, 0, 0);
                    tma::red_add_v4(dbox + ((unsigned)r * pitch + 4u * (unsigned)q), (float)v.x * inv_scale, (float)v.y * inv_scale,
                                    (float)v.z * inv_scale, (float)v.w * inv_scale);
                }
            }
        }
    }
}

template <int C, int TW, int K, int NT, int S, int BW, int BH>
static int launch_v(const float* U, const float* Hs, const float* d_out, const float* d_img, const PipePlan& p, float* dU, float* parts,
                    const LossSrc* loss, cudaStream_t st)
{
    using L = BwdLayout<C, TW, K, NT, S, BW, BH>;
    using G = typename L::G;
    const TileCfg& c = p.cfg.t;
    CUtensorMap mU;
    TRY_RC(make_map(&mU, U, c.W * C, c.H, c.N, G::kRowF, G::SBH));
    cudaError_t e;
    const float* no_d_out = nullptr;
    const int grid = grid_for(p.cfg.nchunks, MGW_PIPE_BWD_MINB);
    if (loss) {
        static bool attr[64] = {};
        TRY_RC(allow_smem(warp_bwd_pipe_kernel<C, TW, K, NT, S, BW, BH, true>, attr, "warp_bwd_pipe(loss)"));
        e = launch_ex(warp_bwd_pipe_kernel<C, TW, K, NT, S, BW, BH, true>, dim3(grid), dim3(NT), L::kTotal, st, pdl_enabled(), mU, U, Hs,
                      no_d_out, d_img, p.cfg, dU, parts, *loss);
    } else {
        static bool attr[64] = {};
        TRY_RC(allow_smem(warp_bwd_pipe_kernel<C, TW, K, NT, S, BW, BH, false>, attr, "warp_bwd_pipe"));
        e = launch_ex(warp_bwd_pipe_kernel<C, TW, K, NT, S, BW, BH, false>, dim3(grid), dim3(NT), L::kTotal, st, pdl_enabled(), mU, U, Hs,
                      d_out, d_img, p.cfg, dU, parts, LossSrc{});
    }
    if (e != cudaSuccess) { count_launches(1); return set_error(MGW_ERR_CUDA, "warp_bwd_pipe: %s", cudaGetErrorString(e)); }
    return check_launch("warp_bwd_pipe");
}

// compiled variant: (TW, K, threads, stages, box width px, box height px); TH = K * threads / TW
#ifndef MGW_PB_K
#define MGW_PB_K 3
#endif
#ifndef MGW_PB_BH
#define MGW_PB_BH 36
#endif
#define MGW_PIPE_BWD 32, MGW_PB_K, 256, 2, 64, MGW_PB_BH
constexpr int kTW = 32, kTH = (256 / 32) * MGW_PB_K;

static bool plan_bwd(const WarpShape& s, PipePlan* p)
{
    if (s.C != 1 && s.C != 3 && s.C != 4) return false;
    if (!(tile_eff(s, kTW, kTH) > 0 && plan(s, kTW, kTH, p))) return false;
    // chunk = the tiles of one tile column inside one cell, when every cell row has the same number of them (uniform mesh);
    // otherwise the CTAs walk single tiles
    PipeCfg& c = p->cfg;
    const int cell_h = s.H / s.gh;
    const bool uniform = (s.H % s.gh == 0) && c.nty == s.gh * c.t.parts_y && c.t.parts_y == (cell_h + kTH - 1) / kTH;
    if (uniform && c.t.parts_y > 1 && getenv("MGW_PB_NOCHUNK") == nullptr) {
        c.chunk_L = c.t.parts_y; c.chunk_rows = s.gh; c.nchunks = s.N * s.gh * c.ntx;
    }
    return true;
}

static int nparts_of(const PipeCfg& c) { return c.chunk_L > 1 ? c.t.parts_x : c.t.parts_y * c.t.parts_x; }

}  // namespace

bool pipe_bwd_supported(const WarpShape& s) { PipePlan p; return plan_bwd(s, &p); }

// dH partials, sized for the tile-by-tile layout (the chunked one needs less): independent of the schedule switch
static size_t parts_bytes(const WarpShape& s, const PipeCfg& c)
{
    return ((size_t)s.N * s.gh * s.gw * c.t.parts_y * c.t.parts_x * 8 * sizeof(float) + 255) / 256 * 256;
}

size_t pipe_bwd_workspace_bytes(const WarpShape& s)
{
    PipePlan p;
    if (!plan_bwd(s, &p)) return 0;
    return parts_bytes(s, p.cfg);
}

int launch_warp_bwd_pipe(const float* U, const float* Hs, const float* d_out, const float* d_img, const WarpShape& s, float* dU,
                         float* parts, int* nparts, const FusedImgLoss* fl, cudaStream_t st)
{
    PipePlan p;
    if (!dU || !plan_bwd(s, &p)) return set_error(MGW_ERR_UNSUPPORTED, "warp_bwd_pipe: unsupported shape");
    LossSrc ls{};
    const LossSrc* loss = nullptr;
    if (fl) { ls.out = fl->out; ls.y = fl->y; ls.black = fl->black; ls.sums = fl->sums; ls.kscale = fl->kscale; ls.kscale_dev = fl->kscale_dev; loss = &ls; }
    *nparts = nparts_of(p.cfg);
    // cells with fewer tiles (tile columns) than slots leave slots untouched: zero them.  Every cell of a uniform mesh has the
    // same tiling, i.e. every slot is written: no zero-fill
    const int cell_h = s.H / s.gh, cell_w = s.W / s.gw;
    const bool all_slots_written = (s.H % s.gh == 0) && (s.W % s.gw == 0) && (p.cfg.nty == s.gh * ((cell_h + kTH - 1) / kTH)) &&
                                   (p.cfg.ntx == s.gw * ((cell_w + kTW - 1) / kTW));
    if (!all_slots_written && cudaMemsetAsync(parts, 0, parts_bytes(s, p.cfg), st) != cudaSuccess)
        return set_error(MGW_ERR_CUDA, "memset parts: %s", cudaGetErrorString(cudaGetLastError()));
    if (s.C == 1) return launch_v<1, MGW_PIPE_BWD>(U, Hs, d_out, d_img, p, dU, parts, loss, st);
    if (s.C == 3) return launch_v<3, MGW_PIPE_BWD>(U, Hs, d_out, d_img, p, dU, parts, loss, st);
    return launch_v<4, MGW_PIPE_BWD>(U, Hs, d_out, d_img, p, dU, parts, loss, st);
}

}  // namespace mgw
