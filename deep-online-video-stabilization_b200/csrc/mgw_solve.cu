// K0/K1/K4: mesh-vertex builder, batched per-cell 8x8 DLT homography solve, and its adjoint.
//
//   vertices_*  : get_4_pts, s_net_bundle_nobm.py:29-71
//   solve_h_fwd : get_Hs/get_H/pinv, spatial_transformer3.py:144-198  (h = inverse(A + 1e-4 I) . b)
//   solve_h_bwd : closed-form adjoint of the above, fp64 inside
//
// Work is tiny (N*gh*gw cells, 512 for config #2) and latency-bound: one thread per cell, the 8x8 system in
// registers (no shuffles, no shared memory in the solve itself).  The operation order is

// exactly oracle/mgw_oracle.c's ORC_SOLVE, so Hs is bit-identical to the C oracle's.
#include "mgw_internal.h"

namespace mgw {

// ---------------------------------------------------------------- K0: vertices
__global__ void vertices_fwd_kernel(const float* __restrict__ head, int N, int gh, int gw, float do_crop_rate,
                                    float* __restrict__ pts2, float* __restrict__ pts1)
{
    const int nv = (gh + 1) * (gw + 1);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * nv * 2) return;
    const int k = t & 1, tot = (t >> 1) % nv, n = (t >> 1) / nv;
    const int i = tot / (gw + 1), j = tot % (gw + 1);
    const double h = 2.0 / gh, w = 2.0 / gw;
    const float base = k == 0 ? (float)(j * w - 1) : (float)(i * h - 1);        // :44-46
    const float lim = __fdiv_rn(1.0f, do_crop_rate);                            // :37
    float p = __fadd_rn(base, head[t]);                                         // :47,:55
    p = fminf(fmaxf(p, -lim), lim);                                             // :58
    pts2[t] = p;
    if (pts1) {                                                                 // :63-68
        // vertex (i,j) is TL of cell (i,j), TR of (i,j-1), BL of (i-1,j), BR of (i-1,j-1)
        const int ci[4] = { i, i, i - 1, i - 1 }, cj[4] = { j, j - 1, j, j - 1 };
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (ci[c] >= 0 && ci[c] < gh && cj[c] >= 0 && cj[c] < gw)
                pts1[(((size_t)n * gh + ci[c]) * gw + cj[c]) * 8 + k * 4 + c] = p;
    }
}

__global__ void vertices_bwd_kernel(const float* __restrict__ head, const float* __restrict__ d_pts2,
                                    const float* __restrict__ d_pts1, int N, int gh, int gw, float do_crop_rate,
                                    float* __restrict__ d_head, float id_coef, const float* __restrict__ id_dev)
{
    const int nv = (gh + 1) * (gw + 1);
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= N * nv * 2) return;
    const int k = t & 1, tot = (t >> 1) % nv, n = (t >> 1) / nv;
    const int i = tot / (gw + 1), j = tot % (gw + 1);
    const double h = 2.0 / gh, w = 2.0 / gw;
    const float base = k == 0 ? (float)(j * w - 1) : (float)(i * h - 1);
    const float lim = __fdiv_rn(1.0f, do_crop_rate);
    const float p = __fadd_rn(base, head[t]);
    float g = d_pts2 ? d_pts2[t] : 0.0f;
    if (d_pts1) {
        const int ci[4] = { i, i, i - 1, i - 1 }, cj[4] = { j, j - 1, j, j - 1 };
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (ci[c] >= 0 && ci[c] < gh && cj[c] >= 0 && cj[c] < gw)
                g += d_pts1[(((size_t)n * gh + ci[c]) * gw + cj[c]) * 8 + k * 4 + c];
    }
    // tf.minimum(tf.maximum(p,-lim),lim): the gradient passes where the clamp is inactive.  At an exact tie
    // TF's Maximum/Minimum gradients route to the first argument (x >= y / x <= y), i.e. they pass.
    float d = (p >= -lim && p <= lim) ? g : 0.0f;
    if (id_coef != 0.0f) {                               // id loss on the head itself: d|t| = sign(t), 0 at 0
        const float t0 = head[t], f = id_dev ? id_coef * __ldg(id_dev) : id_coef;
        d += f * (t0 > 0.0f ? 1.0f : (t0 < 0.0f ? -1.0f : 0.0f));
    }
    d_head[t] = d;
}

// ---------------------------------------------------------------- K1: per-cell DLT solve
// cell corners in output space (ori) and the vertex ids of its 4 corners (TL,TR,BL,BR)
__device__ __forceinline__ void cell_corner(int i, int j, int gh, int gw, int k, float& x, float& y, int& vid)
{
    const double h = 2.0 / gh, w = 2.0 / gw, hh = i * h - 1, ww = j * w - 1;      // spatial_transformer3.py:182-188
    x = (float)((k & 1) ? ww + w : ww);
    y = (float)((k & 2) ? hh + h : hh);
    vid = (i + (k >> 1)) * (gw + 1) + j + (k & 1);
}

template <typename T> __device__ __forceinline__ T tfma(T a, T b, T c);
template <> __device__ __forceinline__ float tfma<float>(float a, float b, float c) { return __fmaf_rn(a, b, c); }
template <> __device__ __forceinline__ double tfma<double>(double a, double b, double c) { return fma(a, b, c); }
template <typename T> __device__ __forceinline__ T tabs(T a) { return a < 0 ? -a : a; }

// 8x8 LU with partial pivoting held entirely in one thread's registers (every index is a compile-time constant after
// unrolling; the data-dependent row swap is a predicated exchange).  Mirrors ORC_SOLVE of oracle/mgw_oracle.c: first
// maximum wins the pivot search, the column is scaled by the reciprocal of the pivot, updates are fused multiply-adds.
// One thread per cell beats a shuffle-cooperative 8-lane version by ~3x here: the work is a latency chain, and a
// register move costs 4 cycles where a shuffle costs ~25.
template <typename T>
__device__ __forceinline__ void lu8(T (&M)[8][8], int (&piv)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        int p = k;
        T best = tabs(M[k][k]);
#pragma unroll
        for (int r = k + 1; r < 8; ++r) {
            const T a = tabs(M[r][k]);
            if (a > best) { best = a; p = r; }
        }
        piv[k] = p;
#pragma unroll
        for (int r = k + 1; r < 8; ++r) {               // branch-free exchange keeps the matrix in registers
            const bool sw = (p == r);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                const T a = M[k][c], b = M[r][c];
                M[k][c] = sw ? b : a;
                M[r][c] = sw ? a : b;
            }
        }
        const T rp = (T)1 / M[k][k];
#pragma unroll
        for (int r = k + 1; r < 8; ++r) {
            const T l = M[r][k] * rp;
            M[r][k] = l;
#pragma unroll
            for (int c = k + 1; c < 8; ++c) M[r][c] = tfma<T>(-l, M[k][c], M[r][c]);
        }
    }
}

// Solve LU x = P rhs in place (LAPACK getrs): row interchanges, unit-lower forward substitution, back substitution.
template <typename T>
__device__ __forceinline__ void getrs8(const T (&LU)[8][8], const int (&piv)[8], T (&col)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int r = k + 1; r < 8; ++r) {
            const bool sw = (piv[k] == r);
            const T a = col[k], b = col[r];
            col[k] = sw ? b : a;
            col[r] = sw ? a : b;
        }
    }
#pragma unroll
    for (int r = 1; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < r; ++k) col[r] = tfma<T>(-LU[r][k], col[k], col[r]);
#pragma unroll
    for (int r = 7; r >= 0; --r) {
#pragma unroll
        for (int k = r + 1; k < 8; ++k) col[r] = tfma<T>(-LU[r][k], col[k], col[r]);
        col[r] = col[r] / LU[r][r];
    }
}

// DLT row g of A + 1e-4 I (spatial_transformer3.py:145,160-167) and its right-hand side b[g] (:169)
template <typename T>
__device__ __forceinline__ void dlt_row(const float* __restrict__ theta_n, int i, int j, int gh, int gw, int g,
                                        T (&row)[8], T& b)
{
    float xf, yf; int vid;
    cell_corner(i, j, gh, gw, g & 3, xf, yf, vid);
    const T x = xf, y = yf, u = theta_n[vid * 2], v = theta_n[vid * 2 + 1];
    const T t = (g < 4) ? u : v;
    const T nx = -x * t, ny = -y * t;
#pragma unroll
    for (int c = 0; c < 8; ++c) row[c] = 0;
    if (g < 4) { row[0] = x; row[1] = y; row[2] = 1; } else { row[3] = x; row[4] = y; row[5] = 1; }
    row[6] = nx; row[7] = ny;
#pragma unroll
    for (int c = 0; c < 8; ++c) if (c == g) row[c] = row[c] + (T)1e-4f;
    b = t;
}

// K1 runs the same LU cooperatively on 8 lanes per cell (lane = matrix row for the factorisation, lane = inverse column
// for the substitutions; measured 6.5 us against 10 us for the one-thread-per-cell form: its chain is the longest,
// 8 right-hand sides).  In-group (8 lanes) LU with partial pivoting of the row-distributed matrix `row` (lane g holds row g).
// Mirrors ORC_SOLVE: first-max pivot, reciprocal scaling, fma updates.  piv[] is group-uniform.
template <typename T>
__device__ __forceinline__ void group_lu(T (&row)[8], int (&piv)[8], int g)
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        T best = (g >= k) ? tabs(row[k]) : (T)-1;
        int p = g;
#pragma unroll
        for (int o = 4; o > 0; o >>= 1) {
            const T ob = __shfl_xor_sync(0xffffffffu, best, o, 8);
            const int op = __shfl_xor_sync(0xffffffffu, p, o, 8);
            if (ob > best || (ob == best && op < p)) { best = ob; p = op; }
        }
        piv[k] = p;
        const int src = (g == k) ? p : ((g == p) ? k : g);
        T pr[8];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            row[c] = __shfl_sync(0xffffffffu, row[c], src, 8);
            pr[c] = __shfl_sync(0xffffffffu, row[c], k, 8);
        }
        const T rp = (T)1 / pr[k];
        if (g > k) {
            const T l = row[k] * rp;
            row[k] = l;
#pragma unroll
            for (int c = k + 1; c < 8; ++c) row[c] = tfma<T>(-l, pr[c], row[c]);
        }
    }
}

// Solve LU x = P e (getrs): lane g owns one right-hand side `col`; LU is read from shared memory.
template <typename T>
__device__ __forceinline__ void group_getrs(const T* __restrict__ LU, const int (&piv)[8], T (&col)[8])
{
#pragma unroll
    for (int k = 0; k < 8; ++k) {
#pragma unroll
        for (int r = k + 1; r < 8; ++r)
            if (piv[k] == r) { const T t = col[k]; col[k] = col[r]; col[r] = t; }
    }
#pragma unroll
    for (int r = 1; r < 8; ++r)
#pragma unroll
        for (int k = 0; k < r; ++k) col[r] = tfma<T>(-LU[r * 8 + k], col[k], col[r]);
#pragma unroll
    for (int r = 7; r >= 0; --r) {
#pragma unroll
        for (int k = r + 1; k < 8; ++k) col[r] = tfma<T>(-LU[r * 8 + k], col[k], col[r]);
        col[r] = col[r] / LU[r * 8 + r];
    }
}

constexpr int kCellsPerBlock = 16;      // 128 threads

template <bool REDUNDANT_LU>
__global__ void __launch_bounds__(kCellsPerBlock * 8)
solve_h_fwd_kernel(const float* __restrict__ theta, int N, int gh, int gw, float* __restrict__ Hs)
{
    __shared__ float sLU[kCellsPerBlock][64];
    __shared__ float sINV[kCellsPerBlock][64];
    griddep_launch_dependents();                       // the warp kernel that follows may set itself up while this one runs
    const int slot = threadIdx.x >> 3, g = threadIdx.x & 7;
    const int ncell = N * gh * gw;
    int cell = blockIdx.x * kCellsPerBlock + slot;
    const bool live = cell < ncell;
    if (!live) cell = ncell - 1;                       // keep the whole warp converged for the shuffles
    const int n = cell / (gh * gw), ij = cell % (gh * gw), i = ij / gw, j = ij % gw;
    const float* theta_n = theta + (size_t)n * (gh + 1) * (gw + 1) * 2;

    float col[8], bk[8];
    if (REDUNDANT_LU) {
        // every lane of the group factorises the whole system in its own registers (no shuffles on the pivot chain), then
        // lane g solves for column g of the inverse
        float M[8][8]; int piv[8];
#pragma unroll
        for (int r = 0; r < 8; ++r) dlt_row<float>(theta_n, i, j, gh, gw, r, M[r], bk[r]);
        lu8<float>(M, piv);
#pragma unroll
        for (int r = 0; r < 8; ++r) col[r] = (r == g) ? 1.0f : 0.0f;
        getrs8<float>(M, piv, col);
    } else {
        float row[8], b; int piv[8];
        dlt_row<float>(theta_n, i, j, gh, gw, g, row, b);
        group_lu<float>(row, piv, g);
#pragma unroll
        for (int c = 0; c < 8; ++c) sLU[slot][g * 8 + c] = row[c];
        __syncwarp();
#pragma unroll
        for (int r = 0; r < 8; ++r) col[r] = (r == g) ? 1.0f : 0.0f;
        group_getrs<float>(sLU[slot], piv, col);
#pragma unroll
        for (int k = 0; k < 8; ++k) bk[k] = __shfl_sync(0xffffffffu, b, k, 8);
    }
#pragma unroll
    for (int r = 0; r < 8; ++r) sINV[slot][r * 8 + g] = col[r];
    __syncwarp();
    // matmul(pinv(A), b): FMA chain over k (spatial_transformer3.py:173)
    float acc = __fmul_rn(sINV[slot][g * 8], bk[0]);
#pragma unroll
    for (int k = 1; k < 8; ++k) acc = __fmaf_rn(sINV[slot][g * 8 + k], bk[k], acc);
    if (live) {
        Hs[(size_t)cell * 9 + g] = acc;
        if (g == 0) Hs[(size_t)cell * 9 + 8] = 1.0f;
    }
}

// ---------------------------------------------------------------- K4: adjoint of the solve
// dHs_part [N*gh*gw, nparts, part_stride] tile partials (nparts may be 1) -> dtheta [N,gh+1,gw+1,2].
// One thread per cell, fp64 (eight for the sum of the partials): g = sum of the partials, lambda = (A+1e-4 I)^-T g by a fresh pivoted LU of the transposed
// system, d u_k = lambda_k * s_k, d v_k = lambda_{4+k} * s_k with s_k = 1 + h6 x_k + h7 y_k; then
// each vertex gathers its (up to) four cells in a fixed order -> deterministic, no atomics.  One block per sample.
__global__ void solve_h_bwd_kernel(const float* __restrict__ theta, const float* __restrict__ Hs,
                                   const float* __restrict__ dHs_part, int nparts, int part_stride,
                                   int N, int gh, int gw, float* __restrict__ dtheta, const float* __restrict__ extra_part)
{
    extern __shared__ double sDuv[];                   // [gh*gw][8]  (du0..3, dv0..3), then [gh*gw][8] summed partials
    const int ncell_s = gh * gw;
    double* sRhs = sDuv + (size_t)ncell_s * 8;
    const int n = blockIdx.x;
    const float* theta_n = theta + (size_t)n * (gh + 1) * (gw + 1) * 2;
    for (int base = 0; base < ncell_s; base += blockDim.x) {
        const int ij = base + threadIdx.x;
        const bool live = ij < ncell_s;
        const int i = ij / gw, j = ij % gw;
        const size_t cell = (size_t)n * ncell_s + ij;
        double Mt[8][8], rhs[8], h6 = 0, h7 = 0;
        int piv[8];
        if (live) {
            h6 = Hs[cell * 9 + 6]; h7 = Hs[cell * 9 + 7];      // written two kernels back (K1): complete under any launch form
#pragma unroll
            for (int r = 0; r < 8; ++r) {
                double row[8], bdummy;
                dlt_row<double>(theta_n, i, j, gh, gw, r, row, bdummy);
#pragma unroll
                for (int c = 0; c < 8; ++c) Mt[c][r] = row[c];
            }
            lu8<double>(Mt, piv);
        }
        // the factorisation needs theta only: under a programmatic dependent launch it runs while the backward warp kernel is
        // still busy; the tile partials are read after that kernel has completed
        if (base == 0) griddep_wait();
        // the partials of this round's cells, summed by all threads of the block: one (cell, component) pair per thread, every
        // load independent of the others (a one-thread-per-cell loop pays one L2 round trip per partial)
        const int cells_here = min((int)blockDim.x, ncell_s - base);
        for (int e = threadIdx.x; e < cells_here * 8; e += blockDim.x) {
            const float* src = dHs_part + (((size_t)n * ncell_s + base + (e >> 3)) * nparts) * part_stride + (e & 7);
            double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
            int p = 0;
            for (; p + 4 <= nparts; p += 4) {
                const float a = __ldcg(src + (size_t)p * part_stride), b = __ldcg(src + (size_t)(p + 1) * part_stride);
                const float c = __ldcg(src + (size_t)(p + 2) * part_stride), d = __ldcg(src + (size_t)(p + 3) * part_stride);
                s0 += (double)a; s1 += (double)b; s2 += (double)c; s3 += (double)d;
            }
            for (; p < nparts; ++p) s0 += (double)__ldcg(src + (size_t)p * part_stride);
            if (extra_part) s1 += (double)__ldcg(extra_part + ((size_t)n * ncell_s + base + (e >> 3)) * 8 + (e & 7));
            sRhs[(size_t)(base + (e >> 3)) * 8 + (e & 7)] = (s0 + s1) + (s2 + s3);
        }
        __syncthreads();
        if (live) {
#pragma unroll
            for (int r = 0; r < 8; ++r) rhs[r] = sRhs[(size_t)ij * 8 + r];
            getrs8<double>(Mt, piv, rhs);                  // rhs -> lambda
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                float xf, yf; int vid;
                cell_corner(i, j, gh, gw, k, xf, yf, vid);
                const double s = 1.0 + h6 * (double)xf + h7 * (double)yf;
                sDuv[(size_t)ij * 8 + k] = rhs[k] * s;
                sDuv[(size_t)ij * 8 + 4 + k] = rhs[4 + k] * s;
            }
        }
    }
    __syncthreads();
    const int nv2 = (gh + 1) * (gw + 1) * 2;
    for (int t = threadIdx.x; t < nv2; t += blockDim.x) {
        const int k = t & 1, tot = t >> 1, i = tot / (gw + 1), j = tot % (gw + 1);
        const int ci[4] = { i, i, i - 1, i - 1 }, cj[4] = { j, j - 1, j, j - 1 };
        double acc = 0;
#pragma unroll
        for (int c = 0; c < 4; ++c)
            if (ci[c] >= 0 && ci[c] < gh && cj[c] >= 0 && cj[c] < gw)
                acc += sDuv[(size_t)(ci[c] * gw + cj[c]) * 8 + k * 4 + c];
        dtheta[(size_t)n * nv2 + t] = (float)acc;
    }
}

// ---------------------------------------------------------------- launchers
int launch_vertices_fwd(const float* head, int N, int gh, int gw, float do_crop_rate, float* pts2, float* pts1,
                        cudaStream_t st)
{
    const int tot = N * (gh + 1) * (gw + 1) * 2;
    vertices_fwd_kernel<<<(tot + 127) / 128, 128, 0, st>>>(head, N, gh, gw, do_crop_rate, pts2, pts1);
    return check_launch("vertices_fwd");
}

int launch_vertices_bwd(const float* head, const float* d_pts2, const float* d_pts1, int N, int gh, int gw,
                        float do_crop_rate, float* d_head, cudaStream_t st, float id_coef, const float* id_dev)
{
    const int tot = N * (gh + 1) * (gw + 1) * 2;
    vertices_bwd_kernel<<<(tot + 127) / 128, 128, 0, st>>>(head, d_pts2, d_pts1, N, gh, gw, do_crop_rate, d_head, id_coef, id_dev);
    return check_launch("vertices_bwd");
}

int launch_solve_h_fwd(const float* theta, int N, int gh, int gw, float* Hs, cudaStream_t st)
{
    const int ncell = N * gh * gw;
    static const bool group = [] { const char* v = getenv("MGW_K1"); return v && v[0] == 'g'; }();
    const int blocks = (ncell + kCellsPerBlock - 1) / kCellsPerBlock;
    if (group) solve_h_fwd_kernel<false><<<blocks, kCellsPerBlock * 8, 0, st>>>(theta, N, gh, gw, Hs);
    else solve_h_fwd_kernel<true><<<blocks, kCellsPerBlock * 8, 0, st>>>(theta, N, gh, gw, Hs);
    return check_launch("solve_h_fwd");
}

int launch_solve_h_bwd(const float* theta, const float* Hs, const float* dHs_part, int nparts, int part_stride,
                       int N, int gh, int gw, float* dtheta, cudaStream_t st, bool after_own_warp_kernel, const float* extra_part)
{
    const int ncell_s = gh * gw;
    // 8 threads per cell for the sum of the partials (one thread per cell for the solve itself)
    int threads = ((ncell_s * 8 > (gh + 1) * (gw + 1) * 2 ? ncell_s * 8 : (gh + 1) * (gw + 1) * 2) + 31) / 32 * 32;
    if (threads > 128) threads = 128;
    const size_t smem = (size_t)ncell_s * 16 * sizeof(double);
    if (smem > 48 * 1024) return set_error(MGW_ERR_UNSUPPORTED, "solve_h_bwd: grid too large (gh*gw > 384)");
    // after_own_warp_kernel: the previous kernel of the stream is this library's backward warp kernel, which read Hs itself and
    // releases its dependents at once: launch programmatically, the factorisation overlaps it
    const cudaError_t e = launch_ex(solve_h_bwd_kernel, dim3(N), dim3(threads), smem, st, after_own_warp_kernel && pdl_enabled(),
                                    theta, Hs, dHs_part, nparts, part_stride, N, gh, gw, dtheta, extra_part);
    if (e != cudaSuccess) { count_launches(1); return set_error(MGW_ERR_CUDA, "solve_h_bwd: %s", cudaGetErrorString(e)); }
    return check_launch("solve_h_bwd");
}

}  // namespace mgw
