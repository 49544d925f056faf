// Tile bookkeeping shared by the TMA-staged kernel families (mgw_warp_tma.cu: one tile per CTA;
// mgw_warp_pipe.cu: persistent, warp-specialised pipeline): axis tables, tensor-map encoding, smem opt-in.
#pragma once
#include <cuda.h>
#include <cstdio>
#include <cstdlib>

#include "mgw_internal.h"
#include "mgw_tma.cuh"

namespace mgw {

#define TRY_RC(expr) do { const int rc_ = (expr); if (rc_ != MGW_OK) return rc_; } while (0)

constexpr int kMaxTilesPerAxis = 128;

// One tile row (or column) of the image: which cell it lies in, where it starts, and from where it OWNS pixels
// (edge tiles are shifted inward so that every tile is full; the overlap is owned by the earlier tile).
struct AxisTab {
    unsigned short start[kMaxTilesPerAxis];
    unsigned short vstart[kMaxTilesPerAxis];
    unsigned char cell[kMaxTilesPerAxis];
    unsigned char part[kMaxTilesPerAxis];
};

struct TileCfg {
    int N, H, W, gh, gw;
    int parts_y, parts_x;           // max tiles per cell (backward partial layout)
    float stepx, stepy;             // tf.linspace step 2/(W-1), 2/(H-1): the host's IEEE single division == __fdiv_rn (set by tile_steps())
    AxisTab rows, cols;             // grid = (cols, rows, N): blockIdx is the tile, no index arithmetic on the device
};

// ------------------------------------------------------------------------------------------------ host side
// lin_step() of mgw_device.cuh on the host: one correctly rounded fp32 division, bit-identical to the device's __fdiv_rn
static inline void tile_steps(TileCfg* c)
{
    volatile float two = 2.0f, w1 = (float)(c->W - 1), h1 = (float)(c->H - 1);      // volatile: no constant folding at another precision
    c->stepx = two / w1; c->stepy = two / h1;
}

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

// fills the axis table; returns the tile count (or -1 if it does not fit) and the max tiles per cell
static int fill_axis(AxisTab* tab, int ncell, int cell_px, int total, int T, int* max_per_cell)
{
    int n = 0, mx = 0;
    for (int c = 0; c < ncell; ++c) {
        const int s = c * cell_px, len = (c == ncell - 1) ? total - s : cell_px;      // the last cell absorbs the remainder
        const int nt = (len + T - 1) / T;
        for (int t = 0; t < nt; ++t) {
            if (n >= kMaxTilesPerAxis) return -1;
            const int vstart = s + t * T;                          // first row/col the tile OWNS
            const int start = vstart < s + len - T ? vstart : s + len - T;      // edge tiles are shifted inward
            tab->start[n] = (unsigned short)start; tab->vstart[n] = (unsigned short)vstart;
            tab->cell[n] = (unsigned char)c; tab->part[n] = (unsigned char)t;
            ++n;
        }
        mx = nt > mx ? nt : mx;
    }
    *max_per_cell = mx;
    return n;
}

template <typename KernelT>
static int allow_smem(KernelT kernel, bool* done_for_device, const char* what)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (!done_for_device[dev & 63]) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
            return set_error(MGW_ERR_CUDA, "cudaFuncSetAttribute(%s): %s", what, cudaGetErrorString(cudaGetLastError()));
        done_for_device[dev & 63] = true;
    }
    return MGW_OK;
}

}  // namespace mgw
