// temp_loss (train_bundle_nobm.py:115-125) on TMA-staged tiles: the two interpolate() passes through a flow field, the masked
// MSE and -- backward -- the scatter of d(out2), for flow fields that are locally compact (optical flow: a 32 x 24 tile of
// output pixels samples a source box of at most 64 x 36 pixels).
//
// One CTA = one 32 x 24 tile of output pixels, 256 threads, thread (tx, g) owns column tx and 3 consecutive rows.
//   1. every thread reads its flow samples and forms its taps (spatial_transformer.py:200-281 = the _interpolate core, weights
//      from the CLIPPED integers); the tile's tap bounding box is reduced warp -> shared -> CTA;
//   2. if the box fits, ONE thread requests the boxes of out2 and (1 - black2's source) black2 by TMA (cp.async.bulk.tensor,
//      SASS UTMALDG); all taps then come from shared memory;
//   3. forward: per-sample sums (sum e^2, sum m) by block reduction + two atomics per tile;
//      backward: d_out1 is stored directly; d_out2 is pre-accumulated in shared memory in FIXED POINT with native integer
//      shared atomics (a pixel adds at most one term to a word and a tile has 768 pixels: 2^-21 of the tile's max|gradient|
//      per term cannot overflow an int32), then converted and sent to global memory with coalesced 16-byte reductions;
//   4. tiles whose box does not fit (or whose gradients hold Inf/NaN) run the per-pixel global path of mgw_loss.cu's kernel.
// Arithmetic of the values (taps, blend, e, m) is the same sequence of fp32 operations as the per-pixel kernel.
#include "mgw_tile.cuh"

namespace mgw {

namespace {

constexpr float kMagic = 12582912.0f;          // 1.5 * 2^23
constexpr int kMagicBits = 0x4B400000;
constexpr int kFixedBits = 21;
__device__ __forceinline__ int fixed_of(float w, float gs) { return __float_as_int(__fmaf_rn(w, gs, kMagic)) - kMagicBits; }

template <int C>
struct TGeo {
    static constexpr int TW = 32, K = 3, NT = 256, TH = (NT / TW) * K;       // 32 x 24
    static constexpr int BW = 64, BH = 36;
    static constexpr int kRowF = (BW * C + 31) / 32 * 32 < 256 ? (BW * C + 31) / 32 * 32 : 256;     // multiple of 32 floats (banks)
    static constexpr int SBW = kRowF / C, SBH = BH;
    static constexpr int kBoxF = SBH * kRowF;
    static constexpr int kMaskRowF = 64;                                    // black2 box: 64 floats per row
    static constexpr int kMaskF = SBH * kMaskRowF;
    static constexpr size_t kAcc = (size_t)kBoxF * 4;
    static constexpr size_t kMask = 2 * (size_t)kBoxF * 4;
    static constexpr size_t kBar = kMask + (size_t)kMaskF * 4;
    static constexpr size_t kRed = kBar + 64;                                // [8][4] ints / floats of the block reductions
    static constexpr size_t kTotal = kRed + 256;
    static_assert(SBW >= 64 || C == 4, "box width");
};

template <bool BWD, int C>
__global__ void __launch_bounds__(256, 3)
temp_loss_tile_kernel(const __grid_constant__ CUtensorMap mapO2, const __grid_constant__ CUtensorMap mapB2,
                      const float* __restrict__ out1, const float* __restrict__ black1, const float* __restrict__ out2,
                      const float* __restrict__ black2, const float* __restrict__ flow, const float* __restrict__ sums_in,
                      float upstream, const float* __restrict__ up_dev, int N, int H, int W, float* __restrict__ sums,
                      float* __restrict__ d_out1, float* __restrict__ d_out2)
{
    using G = TGeo<C>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    float* s_src = reinterpret_cast<float*>(smem_raw);
    int* s_acc = reinterpret_cast<int*>(smem_raw + G::kAcc);
    float* s_mask = reinterpret_cast<float*>(smem_raw + G::kMask);
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw + G::kBar);
    int* s_red = reinterpret_cast<int*>(smem_raw + G::kRed);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.z, r0 = blockIdx.y * G::TH, c0 = blockIdx.x * G::TW;
    const int tx = tid % G::TW, g = tid / G::TW;
    const int col = c0 + tx, row0 = r0 + g * G::K;
    if (tid == 0) {
        tma::mbar_init(bar, 1);
        tma::fence_barrier_init();
    }
    float k = 0.0f;
    if (BWD) {
        if (up_dev) upstream *= __ldg(up_dev);
        k = upstream * 2.0f / ((__ldg(sums_in + 2 * n + 1) + 1e-8f) * (float)N);
    }
    const int HW = H * W;
    const size_t sbase = (size_t)n * HW;
    // 1. flow samples and taps of the thread's pixels; the tile's tap bounding box
    Taps t[G::K];
    bool valid[G::K];
    int xmin = 0x7fffffff, xmax = -1, ymin = 0x7fffffff, ymax = -1;
#pragma unroll
    for (int kk = 0; kk < G::K; ++kk) {
        valid[kk] = (row0 + kk < H) && (col < W);
        float2 f = make_float2(0.0f, 0.0f);
        if (valid[kk]) f = __ldg(reinterpret_cast<const float2*>(flow) + sbase + (size_t)(row0 + kk) * W + col);
        t[kk] = make_taps(f.x, f.y, H, W);                                    // train_bundle_nobm.py:117-118
        if (valid[kk]) {
            xmin = min(xmin, t[kk].x0); xmax = max(xmax, t[kk].x1);
            ymin = min(ymin, t[kk].y0); ymax = max(ymax, t[kk].y1);
        }
    }
    xmin = __reduce_min_sync(0xffffffffu, xmin); xmax = __reduce_max_sync(0xffffffffu, xmax);
    ymin = __reduce_min_sync(0xffffffffu, ymin); ymax = __reduce_max_sync(0xffffffffu, ymax);
    if (lane == 0) { s_red[warp * 4] = xmin; s_red[warp * 4 + 1] = xmax; s_red[warp * 4 + 2] = ymin; s_red[warp * 4 + 3] = ymax; }
    if (BWD) {          // the accumulator is zeroed under the latency of the loads above
        int4* a4 = reinterpret_cast<int4*>(s_acc);
        for (int i = tid; i < G::kBoxF / 4; i += G::NT) a4[i] = make_int4(0, 0, 0, 0);
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        xmin = min(xmin, s_red[w * 4]); xmax = max(xmax, s_red[w * 4 + 1]);
        ymin = min(ymin, s_red[w * 4 + 2]); ymax = max(ymax, s_red[w * 4 + 3]);
    }
    // (clipped taps lie inside the image; a tile without a valid pixel does not exist: the grid covers the image)
    constexpr int kXalign = (C % 4 == 0) ? 1 : 4;          // the box must start on a 16-byte boundary in out2 AND in black2
    const int bx0 = xmin - xmin % 4, by0 = ymin;
    const bool fits = (xmax - bx0 < G::SBW) && (xmax - bx0 < G::kMaskRowF) && (ymax - by0 < G::SBH);
    (void)kXalign;
    // 2. stage the boxes
    if (fits && tid == 0) {
        tma::mbar_expect_tx(bar, (uint32_t)((G::kBoxF + G::kMaskF) * 4));
        tma::load_3d(s_src, &mapO2, bar, bx0 * C, by0, n);
        tma::load_3d(s_mask, &mapB2, bar, bx0, by0, n);
    }
    const float* b2 = black2 + sbase;
    const float* o2 = out2 + sbase * C;
    const float* o1 = out1 + sbase * C;
    const float* b1 = black1 + sbase;
    float b1v[G::K], o1v[G::K][C];
#pragma unroll
    for (int kk = 0; kk < G::K; ++kk) {
        const int q = (row0 + kk) * W + col;
        b1v[kk] = valid[kk] ? __ldg(b1 + q) : 1.0f;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) o1v[kk][ch] = valid[kk] ? __ldg(o1 + (size_t)q * C + ch) : 0.0f;
    }
    if (fits) tma::mbar_wait(bar, 0);
    // 3. values
    float se = 0.0f, sm = 0.0f;
    float g1[BWD ? G::K : 1][BWD ? C : 1];
#pragma unroll
    for (int kk = 0; kk < G::K; ++kk) {
        const Taps& tp = t[kk];
        const int q = (row0 + kk) * W + col;
        float nb2 = 0.0f, v2[C];
#pragma unroll
        for (int ch = 0; ch < C; ++ch) v2[ch] = 0.0f;
        if (!valid[kk]) {
            // a pixel past the image edge of a partial tile: no taps (its dummy taps are not covered by the box)
        } else if (fits) {
            const int ma = (tp.y0 - by0) * G::kMaskRowF + (tp.x0 - bx0), mb = (tp.y1 - by0) * G::kMaskRowF + (tp.x0 - bx0);
            const int mc = (tp.y0 - by0) * G::kMaskRowF + (tp.x1 - bx0), md = (tp.y1 - by0) * G::kMaskRowF + (tp.x1 - bx0);
            nb2 = blend(tp, 1.0f - s_mask[ma], 1.0f - s_mask[mb], 1.0f - s_mask[mc], 1.0f - s_mask[md]);
            const int ia = (tp.y0 - by0) * G::kRowF + (tp.x0 - bx0) * C, ib = (tp.y1 - by0) * G::kRowF + (tp.x0 - bx0) * C;
            const int ic = (tp.y0 - by0) * G::kRowF + (tp.x1 - bx0) * C, id = (tp.y1 - by0) * G::kRowF + (tp.x1 - bx0) * C;
#pragma unroll
            for (int ch = 0; ch < C; ++ch) v2[ch] = blend(tp, s_src[ia + ch], s_src[ib + ch], s_src[ic + ch], s_src[id + ch]);
        } else {
            const int ia = tp.y0 * W + tp.x0, ib = tp.y1 * W + tp.x0, ic = tp.y0 * W + tp.x1, id = tp.y1 * W + tp.x1;
            nb2 = blend(tp, 1.0f - __ldg(b2 + ia), 1.0f - __ldg(b2 + ib), 1.0f - __ldg(b2 + ic), 1.0f - __ldg(b2 + id));
#pragma unroll
            for (int ch = 0; ch < C; ++ch)
                v2[ch] = blend(tp, __ldg(o2 + (size_t)ia * C + ch), __ldg(o2 + (size_t)ib * C + ch), __ldg(o2 + (size_t)ic * C + ch),
                               __ldg(o2 + (size_t)id * C + ch));
        }
        const float m = (1.0f - b1v[kk]) * nb2;                                 // :121
        if (valid[kk]) sm += m;
#pragma unroll
        for (int ch = 0; ch < C; ++ch) {
            const float e = (o1v[kk][ch] - v2[ch]) * m;                          // :120,:122
            if constexpr (!BWD) {
                if (valid[kk]) se = fmaf(e, e, se);
            } else {
                const float gv = k * e * m;
                g1[kk][ch] = valid[kk] ? gv : 0.0f;
                if (valid[kk]) d_out1[(sbase + q) * C + ch] = gv;
            }
        }
    }
    if constexpr (!BWD) {
        se = warp_sum(se); sm = warp_sum(sm);
        float* s_f = reinterpret_cast<float*>(s_red);
        __syncthreads();
        if (lane == 0) { s_f[warp * 2] = se; s_f[warp * 2 + 1] = sm; }
        __syncthreads();
        if (tid == 0) {
            float a = 0.0f, b = 0.0f;
#pragma unroll
            for (int w = 0; w < 8; ++w) { a += s_f[w * 2]; b += s_f[w * 2 + 1]; }
            atomicAdd(sums + 2 * n, a);
            atomicAdd(sums + 2 * n + 1, b);
        }
        return;
    }
    if constexpr (BWD) {
        // 4. scatter of d(out2) = -g1 * w: per-tile fixed-point scale from max|g1|
        unsigned mx = 0u;
#pragma unroll
        for (int kk = 0; kk < G::K; ++kk)
#pragma unroll
            for (int ch = 0; ch < C; ++ch) mx = max(mx, (unsigned)__float_as_int(g1[kk][ch]) & 0x7fffffffu);
        mx = __reduce_max_sync(0xffffffffu, mx);
        __syncthreads();                                  // the bounding-box words of s_red have been read by everybody
        if (lane == 0) s_red[warp] = (int)mx;
        __syncthreads();
        unsigned mb = 0u;
#pragma unroll
        for (int w = 0; w < 8; ++w) mb = max(mb, (unsigned)s_red[w]);
        if (mb == 0u) return;                             // every gradient of the tile is zero
        const int e2 = (int)(mb >> 23) - 127;
        const bool fixed = fits && e2 > -100 && e2 < 100;
        const float scale = __int_as_float((kFixedBits - 1 - e2 + 127) << 23);
        float* d2 = d_out2 + sbase * C;
#pragma unroll
        for (int kk = 0; kk < G::K; ++kk) {
            const Taps& tp = t[kk];
            if (!valid[kk] || !taps_scatter(tp)) continue;      // a clipped pair's two terms cancel exactly (mgw_device.cuh)
            const float wa = tp.ax * tp.ay, wb = tp.ax * tp.by, wc = tp.bx * tp.ay, wd = tp.bx * tp.by;
            if (fixed) {
                const int ia = (tp.y0 - by0) * G::kRowF + (tp.x0 - bx0) * C;
                int* qa = s_acc + ia;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float gs = -g1[kk][ch] * scale;
                    atomicAdd(qa + ch, fixed_of(wa, gs));
                    atomicAdd(qa + G::kRowF + ch, fixed_of(wb, gs));
                    atomicAdd(qa + C + ch, fixed_of(wc, gs));
                    atomicAdd(qa + G::kRowF + C + ch, fixed_of(wd, gs));
                }
            } else {
                const int ia = tp.y0 * W + tp.x0, ib = tp.y1 * W + tp.x0, ic = tp.y0 * W + tp.x1, id = tp.y1 * W + tp.x1;
#pragma unroll
                for (int ch = 0; ch < C; ++ch) {
                    const float gv = g1[kk][ch];
                    atomicAdd(d2 + (size_t)ia * C + ch, -gv * wa);
                    atomicAdd(d2 + (size_t)ib * C + ch, -gv * wb);
                    atomicAdd(d2 + (size_t)ic * C + ch, -gv * wc);
                    atomicAdd(d2 + (size_t)id * C + ch, -gv * wd);
                }
            }
        }
        if (!fixed) return;
        __syncthreads();
        // drain the rows / 16-byte groups of the box that can hold a tap
        const float inv_scale = __int_as_float((e2 - (kFixedBits - 1) + 127) << 23);
        const int nrow = ymax - by0 + 1;
        const int nq = min(((xmax - bx0 + 1) * C + 3) / 4, (W - bx0) * C / 4);
        const int4* a4 = reinterpret_cast<const int4*>(s_acc);
        float* dbox = d2 + (size_t)(unsigned)(by0 * W + bx0) * C;
        const unsigned pitch = (unsigned)(W * C);
        const float rnq = __frcp_rn((float)nq);
        for (int idx = tid; idx < nrow * nq; idx += G::NT) {
            const int r = __float2int_rz(__fmul_rn((float)idx + 0.5f, rnq)), q = idx - r * nq;
            const int4 v = a4[r * (G::kRowF / 4) + q];
            if ((v.x | v.y | v.z | v.w) != 0)
                tma::red_add_v4(dbox + ((unsigned)r * pitch + 4u * (unsigned)q), (float)v.x * inv_scale, (float)v.y * inv_scale,
                                (float)v.z * inv_scale, (float)v.w * inv_scale);
        }
    }
}

template <bool BWD, int C>
static int launch_c(const float* out1, const float* black1, const float* out2, const float* black2, const float* flow, const float* sums_in,
                    float upstream, const float* up_dev, int N, int H, int W, float* sums, float* d_out1, float* d_out2, cudaStream_t st)
{
    using G = TGeo<C>;
    CUtensorMap mO2, mB2;
    TRY_RC(make_map(&mO2, out2, W * C, H, N, G::kRowF, G::SBH));
    TRY_RC(make_map(&mB2, black2, W, H, N, G::kMaskRowF, G::SBH));
    static bool attr[64] = {};
    TRY_RC(allow_smem(temp_loss_tile_kernel<BWD, C>, attr, "temp_loss_tile"));
    const dim3 grid((W + G::TW - 1) / G::TW, (H + G::TH - 1) / G::TH, N);
    temp_loss_tile_kernel<BWD, C><<<grid, G::NT, G::kTotal, st>>>(mO2, mB2, out1, black1, out2, black2, flow, sums_in, upstream, up_dev, N, H, W,
                                                                  sums, d_out1, d_out2);
    return check_launch(BWD ? "temp_loss_tile_bwd" : "temp_loss_tile_fwd");
}

}  // namespace

// the tile kernels need TMA-addressable images: rows of out2 / black2 that are multiples of 16 bytes, 16-byte aligned bases
bool temp_loss_tile_supported(const float* out2, const float* black2, const float* d_out2, int N, int H, int W, int C)
{
    if (C != 1 && C != 3 && C != 4) return false;
    if (W % 4 != 0 || H < 2 || W < 8 || N > 65535 || (H + 23) / 24 > 65535) return false;
    if (((uintptr_t)out2 | (uintptr_t)black2 | (uintptr_t)d_out2) % 16 != 0) return false;
    if (const char* e = getenv("MGW_LOSS_TILE")) if (e[0] == '0') return false;      // tuning aid
    return true;
}

int launch_temp_loss_tile_fwd(const float* out1, const float* black1, const float* out2, const float* black2, const float* flow, int N,
                              int H, int W, int C, float* sums, cudaStream_t st)
{
    if (C == 1) return launch_c<false, 1>(out1, black1, out2, black2, flow, nullptr, 0.0f, nullptr, N, H, W, sums, nullptr, nullptr, st);
    if (C == 3) return launch_c<false, 3>(out1, black1, out2, black2, flow, nullptr, 0.0f, nullptr, N, H, W, sums, nullptr, nullptr, st);
    return launch_c<false, 4>(out1, black1, out2, black2, flow, nullptr, 0.0f, nullptr, N, H, W, sums, nullptr, nullptr, st);
}

int launch_temp_loss_tile_bwd(const float* out1, const float* black1, const float* out2, const float* black2, const float* flow,
                              const float* sums, float upstream, const float* up_dev, int N, int H, int W, int C, float* d_out1,
                              float* d_out2, cudaStream_t st)
{
    if (C == 1) return launch_c<true, 1>(out1, black1, out2, black2, flow, sums, upstream, up_dev, N, H, W, nullptr, d_out1, d_out2, st);
    if (C == 3) return launch_c<true, 3>(out1, black1, out2, black2, flow, sums, upstream, up_dev, N, H, W, nullptr, d_out1, d_out2, st);
    return launch_c<true, 4>(out1, black1, out2, black2, flow, sums, upstream, up_dev, N, H, W, nullptr, d_out1, d_out2, st);
}

}  // namespace mgw
