// Instantiates the TIES merge kernels for ONE source dtype; built 3 times with -DMC_TIES_DT=0..2
// (8 source counts x 3 merge functions each) so the template fan-out compiles in parallel.
#include "mc_ties_kernels.cuh"

namespace mc {
#if MC_TIES_DT == 0
TiesKernels pick_ties_bf16(int n, int f) { return ties_pick_func<__nv_bfloat16>(n, f); }
ties_metrics_fn_t pick_ties_metrics_bf16(int n) { return ties_pick_metrics<__nv_bfloat16>(n); }
#elif MC_TIES_DT == 1
TiesKernels pick_ties_f16(int n, int f) { return ties_pick_func<__half>(n, f); }
ties_metrics_fn_t pick_ties_metrics_f16(int n) { return ties_pick_metrics<__half>(n); }
#elif MC_TIES_DT == 2
TiesKernels pick_ties_f32(int n, int f) { return ties_pick_func<float>(n, f); }
ties_metrics_fn_t pick_ties_metrics_f32(int n) { return ties_pick_metrics<float>(n); }
#else
#error "MC_TIES_DT must be 0..2"
#endif
}  // namespace mc
