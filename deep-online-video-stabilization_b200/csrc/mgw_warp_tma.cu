// placeholder: TMA-staged tile kernels (filled in next)
#include "mgw_internal.h"
namespace mgw {
bool tma_fwd_supported(const WarpShape&) { return false; }
int launch_warp_fwd_tma(const float*, const float*, const WarpShape&, float*, float*, float*, cudaStream_t) { return set_error(MGW_ERR_UNSUPPORTED, "tma fwd not built"); }
bool tma_bwd_supported(const WarpShape&) { return false; }
size_t tma_bwd_workspace_bytes(const WarpShape&) { return 0; }
int launch_warp_bwd_tma(const float*, const float*, const float*, const float*, const WarpShape&, float*, float*, int*, cudaStream_t) { return set_error(MGW_ERR_UNSUPPORTED, "tma bwd not built"); }
}
