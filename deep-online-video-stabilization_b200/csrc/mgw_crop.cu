// Deploy-side crop (deploy_bundle.py:240,291 accumulation of the black masks over the whole video; :344-365 search for the
// largest rectangle without a black pixel whose top-left corner lies on a `step` = 10 pixel lattice in the top-left
// quadrant).  The reference walks every (corner, height, width) triple in Python (O(H*W*H) summed-area lookups); here
//   runs[r][c] = number of consecutive never-black pixels starting at (r, c) going right      (one warp per row),
//   for a corner (i, j): width(hh) = min_{r in i..hh} runs[r][j], area(hh) = (hh - i + 1) * width(hh)  (one thread per corner),
// and the winner is the first (i, j, hh) in the reference's loop order that reaches the maximum area (its update is a
// strict `s > max_s`), kept with one 64-bit atomicMax on (area, ~order).  Integer work: exact.
#include "mgw_internal.h"

namespace mgw {

namespace {

__global__ void __launch_bounds__(256)
black_accumulate_kernel(const float* __restrict__ black, int32_t* __restrict__ all_black, int n)
{
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) all_black[p] += (int32_t)rintf(__ldg(black + p));            // np.round(black).astype(np.int64)   (:291)
}

// one warp per row, 32 columns at a time from the right edge
__global__ void __launch_bounds__(128)
crop_runs_kernel(const int32_t* __restrict__ all_black, int H, int W, int32_t* __restrict__ runs, unsigned long long* __restrict__ best)
{
    const int lane = threadIdx.x & 31;
    const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (blockIdx.x == 0 && threadIdx.x == 0) *best = 0ull;
    if (r >= H) return;
    int carry = 0;                                                          // run length at the first column of the chunk to the right
    for (int c0 = ((W - 1) / 32) * 32; c0 >= 0; c0 -= 32) {
        const int c = c0 + lane;
        const bool blk = c >= W || __ldg(all_black + (size_t)r * W + c) > 0;   // beyond the edge ends a run like a black pixel
        const unsigned m = __ballot_sync(0xffffffffu, blk) >> lane;         // bit k = pixel (c + k) blocks
        const int run = m ? __ffs(m) - 1 : (32 - lane) + carry;
        if (c < W) runs[(size_t)r * W + c] = run;
        carry = __shfl_sync(0xffffffffu, run, 0);
    }
}

// one thread per lattice corner, in the reference's loop order (i outer, j inner)
__global__ void __launch_bounds__(128)
crop_corner_kernel(const int32_t* __restrict__ runs, int H, int W, int step, int ni, int nj, unsigned long long* __restrict__ best)
{
    const int a = blockIdx.x * blockDim.x + threadIdx.x;
    if (a >= ni * nj) return;
    const int i = (a / nj) * step, j = (a % nj) * step;
    int w = runs[(size_t)i * W + j];
    if (w == 0) return;                                                     // all_black[i][j] > 0: continue   (:354)
    long long best_area = 0;
    int best_hh = i;
    for (int hh = i; hh < H && w > 0; ++hh) {
        w = min(w, __ldg(runs + (size_t)hh * W + j));
        const long long s = (long long)(hh - i + 1) * w;
        if (s > best_area) { best_area = s; best_hh = hh; }
    }
    if (best_area == 0) return;
    const unsigned order = (unsigned)a * (unsigned)H + (unsigned)best_hh;   // earlier in the reference's loops = smaller
    atomicMax(best, ((unsigned long long)best_area << 32) | (unsigned long long)(~order));
}

__global__ void crop_decode_kernel(const unsigned long long* __restrict__ best, const int32_t* __restrict__ runs, int H, int W, int step,
                                   int nj, int32_t* __restrict__ rect)
{
    const unsigned long long k = *best;
    if (k == 0ull) { rect[0] = rect[1] = rect[2] = rect[3] = -1; return; }  // the reference's `ans` stays []
    const unsigned order = ~(unsigned)(k & 0xffffffffull);
    const long long area = (long long)(k >> 32);
    const int a = (int)(order / (unsigned)H), hh = (int)(order % (unsigned)H);
    const int i = (a / nj) * step, j = (a % nj) * step;
    rect[0] = i; rect[1] = j; rect[2] = hh; rect[3] = j + (int)(area / (hh - i + 1)) - 1;    // ans = [i, j, hh, ww]   (:365)
}

}  // namespace

int launch_black_accumulate(const float* black, int32_t* all_black, int n, cudaStream_t st)
{
    black_accumulate_kernel<<<(n + 255) / 256, 256, 0, st>>>(black, all_black, n);
    return check_launch("black_accumulate");
}

size_t crop_rect_workspace_bytes(int H, int W) { return (size_t)H * W * sizeof(int32_t) + 16; }

int launch_crop_rect(const int32_t* all_black, int H, int W, int step, void* workspace, int32_t* rect, cudaStream_t st)
{
    // range(0, int(math.floor(height * 0.5)), 10)   (:351, :353)
    const int ni = (H / 2 + step - 1) / step, nj = (W / 2 + step - 1) / step;
    if (ni <= 0 || nj <= 0) return set_error(MGW_ERR_INVALID, "crop_rect: %dx%d has no corner lattice", H, W);
    if ((unsigned long long)ni * nj * H >= (1ull << 32)) return set_error(MGW_ERR_INVALID, "crop_rect: %dx%d too large", H, W);
    unsigned long long* best = reinterpret_cast<unsigned long long*>(workspace);
    int32_t* runs = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(workspace) + 16);
    crop_runs_kernel<<<(H + 3) / 4, 128, 0, st>>>(all_black, H, W, runs, best);
    if (int rc = check_launch("crop_runs")) return rc;
    crop_corner_kernel<<<(ni * nj + 127) / 128, 128, 0, st>>>(runs, H, W, step, ni, nj, best);
    if (int rc = check_launch("crop_corners")) return rc;
    crop_decode_kernel<<<1, 1, 0, st>>>(best, runs, H, W, step, nj, rect);
    return check_launch("crop_decode");
}

}  // namespace mgw
