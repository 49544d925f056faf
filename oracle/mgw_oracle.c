/* TEST INFRASTRUCTURE -- NOT PRODUCT CODE.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg may load this.  The product path (deep-online-video-stabilization_b200/) never does.
 *
 * Plain-C, scalar, single-threaded restatement of the reference's multi-grid warp FORWARD path with the
 * fp32 rounding order made explicit (compile with -ffp-contract=off; every fused multiply-add below is a
 * deliberate fmaf).  It is the bit-exact checker for the CUDA kernels' per-pixel stage.
 *
 * Pinned against tests/golden/*.npz, which are outputs of the unmodified reference sources executed on
 * the torch-CPU tensorflow shim (oracle/make_golden.py): given the reference's Hs, x_map/y_map/black_pix/
 * output_img/cell index reproduce the golden vectors BIT FOR BIT (tests/test_oracle_golden.py).  The 8x8
 * solve is LAPACK-shaped but not LAPACK, so Hs itself is pinned to tolerance (see orc_solve_h).
 *
 * Citations are file:line under /root/reference.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

/* ---- a0: get_4_pts, s_net_bundle_nobm.py:29-71 -------------------------------------------------------
 * head [N, 2*(gh+1)*(gw+1)] -> pts2 [N,gh+1,gw+1,2] (x,y), pts1 [N,gh,gw,8] = [x_tl,x_tr,x_bl,x_br,y_tl,..]
 * vertex = clamp((ww,hh) + head[2*tot:2*tot+2], +-1/do_crop_rate)  (:44-58); ww,hh are python doubles cast
 * to fp32 by tf.constant (:46). */
void orc_vertices(const float *head, int N, int gh, int gw, float do_crop_rate, float *pts2, float *pts1)
{
    const double h = 2.0 / gh, w = 2.0 / gw;
    const float lim = 1.0f / do_crop_rate;              /* tf.ones(..)/do_crop_rate (:37) */
    const int nv = (gh + 1) * (gw + 1);
    for (int n = 0; n < N; ++n) {
        for (int i = 0; i <= gh; ++i)
            for (int j = 0; j <= gw; ++j) {
                const int tot = i * (gw + 1) + j;
                const float base[2] = { (float)(j * w - 1), (float)(i * h - 1) };
                for (int k = 0; k < 2; ++k) {
                    float p = base[k] + head[(size_t)n * 2 * nv + 2 * tot + k];
                    p = fminf(fmaxf(p, -1.0f * lim), lim);                       /* :58 */
                    pts2[((size_t)n * nv + tot) * 2 + k] = p;
                }
            }
        if (pts1)
            for (int i = 0; i < gh; ++i)
                for (int j = 0; j < gw; ++j) {
                    const int v[4] = { i * (gw + 1) + j, i * (gw + 1) + j + 1, (i + 1) * (gw + 1) + j, (i + 1) * (gw + 1) + j + 1 };
                    float *g = pts1 + (((size_t)n * gh + i) * gw + j) * 8;       /* :65-66: [2,4] -> 8 */
                    for (int k = 0; k < 2; ++k)
                        for (int c = 0; c < 4; ++c)
                            g[k * 4 + c] = pts2[((size_t)n * nv + v[c]) * 2 + k];
                }
    }
}

/* ---- a1/a2: get_Hs / get_H / pinv, spatial_transformer3.py:144-198 ------------------------------------
 * Per cell: A (8x8 DLT, :160-167), b (:169), h = inverse(A + 1e-4*I) . b (:145,:173), H = [h,1].
 * The inverse is formed explicitly (LU with partial pivoting, then solve for the identity columns, the
 * route torch.linalg.inv / LAPACK getrf+getrs take) and THEN multiplied by b, to mimic the reference's
 * less accurate inverse-then-matmul rather than a direct solve. */
#define ORC_SOLVE(NAME, T, FMA)                                                                         \
    static void NAME(const float *ori, const float *tar, T *hout)                                       \
    {                                                                                                   \
        T M[8][8], inv[8][8], b[8];                                                                     \
        int piv[8];                                                                                     \
        for (int k = 0; k < 4; ++k) {                                                                   \
            const T x = ori[2 * k], y = ori[2 * k + 1], u = tar[2 * k], v = tar[2 * k + 1];             \
            const T r0[8] = { x, y, 1, 0, 0, 0, -x * u, -y * u };                                       \
            const T r1[8] = { 0, 0, 0, x, y, 1, -x * v, -y * v };                                       \
            for (int c = 0; c < 8; ++c) { M[k][c] = r0[c]; M[4 + k][c] = r1[c]; }                       \
            b[k] = u; b[4 + k] = v;                                                                     \
        }                                                                                               \
        for (int d = 0; d < 8; ++d) M[d][d] = M[d][d] + (T)1e-4f;       /* tf.eye(8)*1e-4 is fp32 */    \
        for (int k = 0; k < 8; ++k) {                      /* LU, partial pivoting, unit-lower L */     \
            int p = k; T best = M[k][k] < 0 ? -M[k][k] : M[k][k];                                       \
            for (int r = k + 1; r < 8; ++r) { T a = M[r][k] < 0 ? -M[r][k] : M[r][k]; if (a > best) { best = a; p = r; } } \
            piv[k] = p;                                                                                 \
            if (p != k) for (int c = 0; c < 8; ++c) { T t = M[k][c]; M[k][c] = M[p][c]; M[p][c] = t; }  \
            const T rp = (T)1 / M[k][k];                   /* getf2 scales by the reciprocal */        \
            for (int r = k + 1; r < 8; ++r) {                                                           \
                const T l = M[r][k] * rp; M[r][k] = l;                                                  \
                for (int c = k + 1; c < 8; ++c) M[r][c] = FMA(-l, M[k][c], M[r][c]);                    \
            }                                                                                           \
        }                                                                                               \
        for (int c = 0; c < 8; ++c) {                      /* getrs on identity column c */             \
            T col[8];                                                                                   \
            for (int r = 0; r < 8; ++r) col[r] = (r == c) ? (T)1 : (T)0;                                \
            for (int k = 0; k < 8; ++k) if (piv[k] != k) { T t = col[k]; col[k] = col[piv[k]]; col[piv[k]] = t; } \
            for (int r = 1; r < 8; ++r) for (int k = 0; k < r; ++k) col[r] = FMA(-M[r][k], col[k], col[r]); \
            for (int r = 7; r >= 0; --r) {                                                              \
                for (int k = r + 1; k < 8; ++k) col[r] = FMA(-M[r][k], col[k], col[r]);                 \
                col[r] = col[r] / M[r][r];                                                              \
            }                                                                                           \
            for (int r = 0; r < 8; ++r) inv[r][c] = col[r];                                             \
        }                                                                                               \
        for (int r = 0; r < 8; ++r) {                      /* matmul(pinv(A), b): FMA chain over k */   \
            T acc = inv[r][0] * b[0];                                                                   \
            for (int k = 1; k < 8; ++k) acc = FMA(inv[r][k], b[k], acc);                                \
            hout[r] = acc;                                                                              \
        }                                                                                               \
        hout[8] = 1;                                                                                    \
    }
ORC_SOLVE(solve_cell_f32, float, fmaf)
ORC_SOLVE(solve_cell_f64, double, fma)

static void cell_ori(int i, int j, int gh, int gw, float *ori)
{   /* :182-189: python doubles, cast to fp32 by tf.constant(dtype=tf.float32) */
    const double h = 2.0 / gh, w = 2.0 / gw, hh = i * h - 1, ww = j * w - 1;
    const double o[8] = { ww, hh, ww + w, hh, ww, hh + h, ww + w, hh + h };
    for (int k = 0; k < 8; ++k) ori[k] = (float)o[k];
}

static void cell_tar(const float *theta, int n, int i, int j, int gh, int gw, float *tar)
{   /* :191-193: vertices TL, TR, BL, BR as (x,y) pairs */
    const int nv = (gh + 1) * (gw + 1);
    const int v[4] = { i * (gw + 1) + j, i * (gw + 1) + j + 1, (i + 1) * (gw + 1) + j, (i + 1) * (gw + 1) + j + 1 };
    for (int c = 0; c < 4; ++c) {
        tar[2 * c] = theta[((size_t)n * nv + v[c]) * 2];
        tar[2 * c + 1] = theta[((size_t)n * nv + v[c]) * 2 + 1];
    }
}

void orc_solve_h(const float *theta, int N, int gh, int gw, float *Hs)
{
    float ori[8], tar[8];
    for (int n = 0; n < N; ++n)
        for (int i = 0; i < gh; ++i)
            for (int j = 0; j < gw; ++j) {
                cell_ori(i, j, gh, gw, ori);
                cell_tar(theta, n, i, j, gh, gw, tar);
                solve_cell_f32(ori, tar, Hs + (((size_t)n * gh + i) * gw + j) * 9);
            }
}

void orc_solve_h_f64(const float *theta, int N, int gh, int gw, double *Hs)
{
    float ori[8], tar[8];
    for (int n = 0; n < N; ++n)
        for (int i = 0; i < gh; ++i)
            for (int j = 0; j < gw; ++j) {
                cell_ori(i, j, gh, gw, ori);
                cell_tar(theta, n, i, j, gh, gw, tar);
                solve_cell_f64(ori, tar, Hs + (((size_t)n * gh + i) * gw + j) * 9);
            }
}

/* ---- shared per-pixel pieces -------------------------------------------------------------------------- */

/* tf.linspace(-1,1,num)[i] = start + step*i in fp32, step=(stop-start)/(num-1) (TF 1.x LinSpaceOp;
 * spatial_transformer3.py:205-206).  mul then add, no fma. */
static inline float lin(int i, int num)
{
    const float step = (1.0f - (-1.0f)) / (float)(num - 1);
    const float t = step * (float)i;
    return -1.0f + t;
}

/* tf.cast(tf.floor(x),'int32') on x86: out-of-range / NaN -> INT32_MIN ("integer indefinite"). */
static inline int32_t floor_to_i32(float x)
{
    const float f = floorf(x);
    if (!(fabsf(f) < 2147483648.0f)) return INT32_MIN;
    return (int32_t)f;
}

static inline int32_t clipi(int32_t v, int32_t lo, int32_t hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* _interpolate, spatial_transformer3.py:62-123, one output pixel, C channels. */
static inline void interp_px(const float *im, int IH, int IW, int C, float xn, float yn, float *out)
{
    const float x = ((xn + 1.0f) * (float)IW) / 2.0f;                                  /* :81 */
    const float y = ((yn + 1.0f) * (float)IH) / 2.0f;                                  /* :82 */
    int32_t x0 = floor_to_i32(x), y0 = floor_to_i32(y);                                /* :85,:87 */
    int32_t x1 = (int32_t)((uint32_t)x0 + 1u), y1 = (int32_t)((uint32_t)y0 + 1u);      /* :86,:88 */
    x0 = clipi(x0, 0, IW - 1); x1 = clipi(x1, 0, IW - 1);                              /* :90-93 */
    y0 = clipi(y0, 0, IH - 1); y1 = clipi(y1, 0, IH - 1);
    const float x0f = (float)x0, x1f = (float)x1, y0f = (float)y0, y1f = (float)y1;    /* :114-117 */
    const float wa = (x1f - x) * (y1f - y), wb = (x1f - x) * (y - y0f);                /* :118-121 */
    const float wc = (x - x0f) * (y1f - y), wd = (x - x0f) * (y - y0f);
    const float *Ia = im + ((size_t)y0 * IW + x0) * C, *Ib = im + ((size_t)y1 * IW + x0) * C;   /* :99-111 */
    const float *Ic = im + ((size_t)y0 * IW + x1) * C, *Id = im + ((size_t)y1 * IW + x1) * C;
    for (int c = 0; c < C; ++c) {
        float s = wa * Ia[c];                                                          /* add_n, left to right :122 */
        float t = wb * Ib[c]; s = s + t;
        t = wc * Ic[c]; s = s + t;
        t = wd * Id[c]; s = s + t;
        out[c] = s;
    }
}

/* T_g = matmul(H, grid) row (spatial_transformer3.py:248): K=3 GEMM accumulation as every FMA-based
 * GEMM does it -- acc=h0*x; acc=fma(h1,y,acc); acc=fma(h2,1,acc) -- which is what the reference
 * produces on this container's BLAS bit for bit (tests/test_oracle_golden.py). */
static inline float hrow(float h0, float h1, float h2, float x, float y)
{
    float acc = h0 * x;
    acc = fmaf(h1, y, acc);
    return acc + h2;
}

/* projective map of one pixel: :248-260 */
static inline void project(const float *Hc, float xt, float yt, float *xn, float *yn)
{
    const float xs = hrow(Hc[0], Hc[1], Hc[2], xt, yt);
    const float ys = hrow(Hc[3], Hc[4], Hc[5], xt, yt);
    float zs = hrow(Hc[6], Hc[7], Hc[8], xt, yt);
    const float sign = (zs >= 0.0f) ? 1.0f : -1.0f;                                    /* :257 */
    zs = zs + sign * 1e-8f;                                                            /* :258 */
    *xn = xs / zs;                                                                     /* :259-260 */
    *yn = ys / zs;
}

static inline float black_of(float xn, float yn)
{   /* :284-286 strict compares on normalised coords; NaN -> 0 */
    return ((-1.0f > xn) || (xn > 1.0f) || (-1.0f > yn) || (yn > 1.0f)) ? 1.0f : 0.0f;
}

/* ---- a3-a5: _transform3 given Hs, spatial_transformer3.py:218-301 ------------------------------------- */
void orc_warp(const float *U, const float *Hs, int N, int H, int W, int C, int gh, int gw,
              float *out, float *black, float *img, int32_t *cell_idx)
{
    const int ghp = H / gh, gwp = W / gw;                                              /* :227-228 */
    for (int n = 0; n < N; ++n)
        for (int r = 0; r < H; ++r) {
            int ci = r / ghp; if (ci > gh - 1) ci = gh - 1;                            /* :236-243 last cell absorbs */
            const float yt = lin(r, H);
            for (int c = 0; c < W; ++c) {
                int cj = c / gwp; if (cj > gw - 1) cj = gw - 1;
                const float xt = lin(c, W);
                const float *Hc = Hs + (((size_t)n * gh + ci) * gw + cj) * 9;
                float xn, yn;
                project(Hc, xt, yt, &xn, &yn);
                const size_t p = ((size_t)n * H + r) * W + c;
                if (img) { img[2 * p] = xn; img[2 * p + 1] = yn; }                     /* :271-272,:278 */
                if (black) black[p] = black_of(xn, yn);
                if (cell_idx) cell_idx[p] = ci * gw + cj;
                if (out) interp_px(U + (size_t)n * H * W * C, H, W, C, xn, yn, out + p * C);
            }
        }
}

/* ---- a6: interpolate(im, x, y, out_size), spatial_transformer.py:200-281 ------------------------------ */
void orc_interp(const float *im, const float *x, const float *y, int N, int IH, int IW, int C, int OH, int OW, float *out)
{
    for (int n = 0; n < N; ++n)
        for (size_t q = 0; q < (size_t)OH * OW; ++q) {
            const size_t p = (size_t)n * OH * OW + q;
            interp_px(im + (size_t)n * IH * IW * C, IH, IW, C, x[p], y[p], out + p * C);
        }
}

/* ---- a7: spatial_transformer.transformer(U, theta[N,9], out_size), spatial_transformer.py:143-193 ----
 * theta /= theta[2,2] (:151-153), one global meshgrid over the OUTPUT size (:161), same sign-eps/divide/
 * mask/_interpolate.  black is reshaped with the INPUT dims (:184), so out_size must equal the input size. */
void orc_homography_warp(const float *U, const float *theta, int N, int H, int W, int C, int OH, int OW,
                         float *out, float *black, float *img)
{
    for (int n = 0; n < N; ++n) {
        float Hn[9];
        for (int k = 0; k < 9; ++k) Hn[k] = theta[n * 9 + k] / theta[n * 9 + 8];
        for (int r = 0; r < OH; ++r) {
            const float yt = lin(r, OH);
            for (int c = 0; c < OW; ++c) {
                const float xt = lin(c, OW);
                float xn, yn;
                project(Hn, xt, yt, &xn, &yn);
                const size_t p = ((size_t)n * OH + r) * OW + c;
                if (img) { img[2 * p] = xn; img[2 * p + 1] = yn; }
                if (black) black[p] = black_of(xn, yn);
                if (out) interp_px(U + (size_t)n * H * W * C, H, W, C, xn, yn, out + p * C);
            }
        }
    }
}

int orc_version(void) { return 1; }
