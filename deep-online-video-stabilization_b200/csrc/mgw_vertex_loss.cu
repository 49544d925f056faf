// Vertex regularisers of the training step (s_net_bundle_nobm.py): the four loss terms that depend only on the mesh head
//   id loss           :246-247   mean(|theta|) * id_mul
//   black_pos loss    :139-148, :313-317   mean(black_err^2) on pts1, black_err = overshoot of a vertex beyond +-1/do_crop_rate
//   distortion loss   :150-184   eight rotated-edge residuals per cell on pts1, mean / 8
//   consistency loss  :186-210   second differences of the vertex lattice pts2 (every triple enters twice, once per end)
// One CTA computes all four SUMS in a fixed order (deterministic; the work is a few thousand elements); the caller divides by
// the element counts -- analytic, and the GLOBAL ones under data parallelism.  The backward kernel takes the four
// d(total)/d(sum) factors and writes d_theta (id term), d_pts1 (black_pos + distortion) and d_pts2 (consistency); the
// chain back to the head goes through mgw_vertices_bwd.
#include "mgw_internal.h"

namespace mgw {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float block_sum256(float v, float* sh)
{
    v = warp_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = v;
    __syncthreads();
    v = (threadIdx.x < kThreads / 32) ? sh[threadIdx.x] : 0.0f;
    if (threadIdx.x < 32) v = warp_sum(v);
    return v;                                   // valid in thread 0
}

// calc_distortion_loss(p0, p1, p2, clock, hw) for one cell (:150-166): e = R (p1 - p0) - (p2 - p1), R = [0,-k;k,0] (clock: its
// transpose), loss |e|^2 per component.  A, B, C = which corners play p0, p1, p2.
template <int A, int B, int C, int CLOCK, int HW>
__device__ __forceinline__ float edge_fwd(const float (&x)[4], const float (&y)[4], float k0, float k1)
{
    const float k = HW ? k1 : k0;
    const float vx = x[B] - x[A], vy = y[B] - y[A];
    const float rx = CLOCK ? k * vy : -k * vy, ry = CLOCK ? -k * vx : k * vx;
    const float ex = fabsf(rx - (x[C] - x[B])), ey = fabsf(ry - (y[C] - y[B]));
    return ex * ex + ey * ey;
}

template <int A, int B, int C, int CLOCK, int HW>
__device__ __forceinline__ void edge_bwd(const float (&x)[4], const float (&y)[4], float (&gx)[4], float (&gy)[4], float k0, float k1,
                                         float f2)
{
    const float k = HW ? k1 : k0;
    const float vx = x[B] - x[A], vy = y[B] - y[A];
    const float rx = CLOCK ? k * vy : -k * vy, ry = CLOCK ? -k * vx : k * vx;
    // d(|e|^2) = 2 |e| sign(e) = 2 e;  f2 = 2 * d(total)/d(sum)
    const float ux = f2 * (rx - (x[C] - x[B])), uy = f2 * (ry - (y[C] - y[B]));
    // R^T u : R = [0,-k;k,0] -> (k uy, -k ux);  clock -> (-k uy, k ux)
    const float tx = CLOCK ? -k * uy : k * uy, ty = CLOCK ? k * ux : -k * ux;
    gx[A] -= tx; gy[A] -= ty;
    gx[B] += tx + ux; gy[B] += ty + uy;
    gx[C] -= ux; gy[C] -= uy;
}

// the eight calls of get_distortion_loss (:174-181), corners p0..p3 = tl, tr, bl, br
#define MGW_EDGES(F, ...)                                                                                                  \
    F<0, 1, 3, 0, 0>(__VA_ARGS__); F<1, 3, 2, 0, 1>(__VA_ARGS__); F<3, 2, 0, 0, 0>(__VA_ARGS__); F<2, 0, 1, 0, 1>(__VA_ARGS__); \
    F<1, 0, 2, 1, 0>(__VA_ARGS__); F<0, 2, 3, 1, 1>(__VA_ARGS__); F<2, 3, 1, 1, 0>(__VA_ARGS__); F<3, 1, 0, 1, 1>(__VA_ARGS__);

__device__ __forceinline__ float overshoot(float p, float one, float* sign)
{
    // where(p > one, p - one, 0) + where(-one > p, -one - p, 0)
    const float hi = p > one ? p - one : 0.0f, lo = -one > p ? -one - p : 0.0f;
    if (sign) *sign = p > one ? 1.0f : (-one > p ? -1.0f : 0.0f);
    return hi + lo;
}

// the two spellings of a lattice second difference the reference sums: from the far end  2 m - e - o   (e = the vertex the
// term is listed under, o = the opposite end)
__device__ __forceinline__ float second_diff(float m, float e, float o) { return fabsf(2.0f * m - e - o); }

__global__ void __launch_bounds__(kThreads)
vertex_losses_fwd_kernel(const float* __restrict__ theta, const float* __restrict__ pts1, const float* __restrict__ pts2, int N, int gh,
                         int gw, float one, float k0, float k1, float* __restrict__ sums, float* __restrict__ black_err)
{
    __shared__ float sh[kThreads / 32];
    const int V = (gh + 1) * (gw + 1);
    float s_id = 0.0f, s_black = 0.0f, s_dist = 0.0f, s_cons = 0.0f;
    // one block: fixed summation order, sums overwritten; several blocks (sums zeroed by the caller): one reduce-add per block and term
    const int t0 = blockIdx.x * kThreads + threadIdx.x, nthr = gridDim.x * kThreads;
    if (theta)
        for (int q = t0; q < N * V * 2; q += nthr) s_id += fabsf(__ldg(theta + q));
    if (pts1)
        for (int q = t0; q < N * gh * gw; q += nthr) {
            float x[4], y[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { x[k] = __ldg(pts1 + (size_t)q * 8 + k); y[k] = __ldg(pts1 + (size_t)q * 8 + 4 + k); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float ex = overshoot(x[k], one, nullptr), ey = overshoot(y[k], one, nullptr);
                if (black_err) { black_err[(size_t)q * 8 + k] = ex; black_err[(size_t)q * 8 + 4 + k] = ey; }
                s_black += ex * ex + ey * ey;
            }
            float cell = 0.0f;
            MGW_EDGES(cell += edge_fwd, x, y, k0, k1)
            s_dist += cell;
        }
    if (pts2)
        for (int q = t0; q < N * V * 2; q += nthr) {
            const int comp = q & 1, v = (q >> 1) % V, n = (q >> 1) / V, i = v / (gw + 1), j = v % (gw + 1);
            const float* P = pts2 + (size_t)n * V * 2 + comp;
            auto at = [&](int ii, int jj) { return __ldg(P + (size_t)(ii * (gw + 1) + jj) * 2); };
            const float p = at(i, j);
            float s = 0.0f, d;
            if (i > 1) { d = second_diff(at(i - 1, j), p, at(i - 2, j)); s += d * d; }
            if (j > 1) { d = second_diff(at(i, j - 1), p, at(i, j - 2)); s += d * d; }
            if (i < gh - 1) { d = second_diff(at(i + 1, j), p, at(i + 2, j)); s += d * d; }
            if (j < gw - 1) { d = second_diff(at(i, j + 1), p, at(i, j + 2)); s += d * d; }
            s_cons += s;
        }
    s_id = block_sum256(s_id, sh);
    s_black = block_sum256(s_black, sh);
    s_dist = block_sum256(s_dist, sh);
    s_cons = block_sum256(s_cons, sh);
    if (threadIdx.x == 0) {
        if (gridDim.x == 1) { sums[0] = s_id; sums[1] = s_black; sums[2] = s_dist; sums[3] = s_cons; }
        else { atomicAdd(sums, s_id); atomicAdd(sums + 1, s_black); atomicAdd(sums + 2, s_dist); atomicAdd(sums + 3, s_cons); }
    }
}

// f[k] (device) = d(total) / d(sums[k]); or (f == nullptr) coef[k] * (g_dev ? *g_dev : 1).  ACC: d_pts2 += instead of =
struct Coef4 { float v[4]; };
template <bool ACC>
__global__ void __launch_bounds__(kThreads)
vertex_losses_bwd_kernel(const float* __restrict__ theta, const float* __restrict__ pts1, const float* __restrict__ pts2, int N, int gh,
                         int gw, float one, float k0, float k1, const float* f, const Coef4 coef, const float* __restrict__ g_dev,
                         float* __restrict__ d_theta, float* __restrict__ d_pts1, float* __restrict__ d_pts2)
{
    const int V = (gh + 1) * (gw + 1);
    const int tid = blockIdx.x * kThreads + threadIdx.x, nthr = gridDim.x * kThreads;
    const float gg = g_dev ? __ldg(g_dev) : 1.0f;
    const float f_id = f ? __ldg(f) : coef.v[0] * gg, f_black = f ? __ldg(f + 1) : coef.v[1] * gg;
    const float f_dist = f ? __ldg(f + 2) : coef.v[2] * gg, f_cons = f ? __ldg(f + 3) : coef.v[3] * gg;
    if (theta && d_theta)
        for (int q = tid; q < N * V * 2; q += nthr) {
            const float t = __ldg(theta + q);
            d_theta[q] = f_id * (t > 0.0f ? 1.0f : (t < 0.0f ? -1.0f : 0.0f));              // d|t| = sign(t), 0 at 0
        }
    if (pts1 && d_pts1)
        for (int q = tid; q < N * gh * gw; q += nthr) {
            float x[4], y[4], gx[4], gy[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) { x[k] = __ldg(pts1 + (size_t)q * 8 + k); y[k] = __ldg(pts1 + (size_t)q * 8 + 4 + k); }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float sx, sy;
                const float ex = overshoot(x[k], one, &sx), ey = overshoot(y[k], one, &sy);
                gx[k] = f_black * 2.0f * ex * sx;
                gy[k] = f_black * 2.0f * ey * sy;
            }
            MGW_EDGES(edge_bwd, x, y, gx, gy, k0, k1, f_dist * 2.0f)
#pragma unroll
            for (int k = 0; k < 4; ++k) { d_pts1[(size_t)q * 8 + k] = gx[k]; d_pts1[(size_t)q * 8 + 4 + k] = gy[k]; }
        }
    if (pts2 && d_pts2)
        for (int q = tid; q < N * V * 2; q += nthr) {
            const int comp = q & 1, v = (q >> 1) % V, n = (q >> 1) / V, i = v / (gw + 1), j = v % (gw + 1);
            const float* P = pts2 + (size_t)n * V * 2 + comp;
            auto at = [&](int ii, int jj) { return __ldg(P + (size_t)(ii * (gw + 1) + jj) * 2); };
            // signed residuals of the triple centred at `m` along one axis, both spellings: (2m - hi - lo) + (2m - lo - hi)
            auto pair = [&](float lo, float m, float hi) { return (2.0f * m - hi - lo) + (2.0f * m - lo - hi); };
            float g = 0.0f;
            // vertical triples (rows r-1, r, r+1), r in [1, gh-1]; horizontal ones likewise
            if (i >= 1 && i <= gh - 1) g += 2.0f * pair(at(i - 1, j), at(i, j), at(i + 1, j));      // this vertex is the middle
            if (i >= 2) g -= pair(at(i - 2, j), at(i - 1, j), at(i, j));                            // ... the lower end
            if (i + 2 <= gh) g -= pair(at(i, j), at(i + 1, j), at(i + 2, j));                       // ... the upper end
            if (j >= 1 && j <= gw - 1) g += 2.0f * pair(at(i, j - 1), at(i, j), at(i, j + 1));
            if (j >= 2) g -= pair(at(i, j - 2), at(i, j - 1), at(i, j));
            if (j + 2 <= gw) g -= pair(at(i, j), at(i, j + 1), at(i, j + 2));
            if (ACC) d_pts2[q] += f_cons * 2.0f * g; else d_pts2[q] = f_cons * 2.0f * g;
        }
}

}  // namespace

int launch_vertex_losses_fwd(const float* theta, const float* pts1, const float* pts2, int N, int gh, int gw, float do_crop_rate,
                             float* sums, float* black_err, cudaStream_t st, bool sums_zeroed)
{
    // h = 2.0 / grid_h, w = 2.0 / grid_w; k = h / w (hw == 0) or w / h, in Python doubles then a float32 constant (:151-163)
    const double h = 2.0 / gh, w = 2.0 / gw;
    const int items = N * (gh + 1) * (gw + 1) * 2;
    const int blocks = sums_zeroed ? std::max(1, std::min(32, items / kThreads)) : 1;
    vertex_losses_fwd_kernel<<<blocks, kThreads, 0, st>>>(theta, pts1, pts2, N, gh, gw, 1.0f / do_crop_rate, (float)(h / w), (float)(w / h), sums,
                                                     black_err);
    return check_launch("vertex_losses_fwd");
}

int launch_vertex_losses_bwd(const float* theta, const float* pts1, const float* pts2, int N, int gh, int gw, float do_crop_rate,
                             const float* f, float* d_theta, float* d_pts1, float* d_pts2,
                             cudaStream_t st)
{
    const double h = 2.0 / gh, w = 2.0 / gw;
    const int items = N * (gh + 1) * (gw + 1) * 2;
    const int blocks = std::min(148, (items + kThreads - 1) / kThreads);
    vertex_losses_bwd_kernel<false><<<blocks, kThreads, 0, st>>>(theta, pts1, pts2, N, gh, gw, 1.0f / do_crop_rate, (float)(h / w),
                                                                  (float)(w / h), f, Coef4{}, nullptr, d_theta, d_pts1, d_pts2);
    return check_launch("vertex_losses_bwd");
}

int launch_vertex_losses_bwd_coef(const float* pts1, const float* pts2, int N, int gh, int gw, float do_crop_rate, const float coef[4],
                                  const float* g_dev, float* d_pts1, float* d_pts2_acc, cudaStream_t st)
{
    const double h = 2.0 / gh, w = 2.0 / gw;
    const int items = N * (gh + 1) * (gw + 1) * 2;
    const int blocks = std::min(148, (items + kThreads - 1) / kThreads);
    const Coef4 c{{coef[0], coef[1], coef[2], coef[3]}};
    vertex_losses_bwd_kernel<true><<<blocks, kThreads, 0, st>>>(nullptr, pts1, pts2, N, gh, gw, 1.0f / do_crop_rate, (float)(h / w),
                                                                 (float)(w / h), nullptr, c, g_dev, nullptr, d_pts1, d_pts2_acc);
    return check_launch("vertex_losses_bwd");
}

}  // namespace mgw
