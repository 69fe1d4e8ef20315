// Instantiates the merge kernels for ONE (src, dst) dtype pair; built 7 times with -DMC_PAIR=0..6
// so the template fan-out (8 source counts x 6 variants) compiles in parallel.
#include "mc_merge_kernels.cuh"

namespace mc {
#if MC_PAIR == 0
merge_fn_t pick_merge_bf16_bf16(int n, int v) { return pick_nsrc<__nv_bfloat16, __nv_bfloat16>(n, v); }
#elif MC_PAIR == 1
merge_fn_t pick_merge_f16_f16(int n, int v) { return pick_nsrc<__half, __half>(n, v); }
#elif MC_PAIR == 2
merge_fn_t pick_merge_f32_f32(int n, int v) { return pick_nsrc<float, float>(n, v); }
#elif MC_PAIR == 3
merge_fn_t pick_merge_bf16_f32(int n, int v) { return pick_nsrc<__nv_bfloat16, float>(n, v); }
#elif MC_PAIR == 4
merge_fn_t pick_merge_f16_f32(int n, int v) { return pick_nsrc<__half, float>(n, v); }
#elif MC_PAIR == 5
merge_fn_t pick_merge_f32_bf16(int n, int v) { return pick_nsrc<float, __nv_bfloat16>(n, v); }
#elif MC_PAIR == 6
merge_fn_t pick_merge_f32_f16(int n, int v) { return pick_nsrc<float, __half>(n, v); }
#else
#error "MC_PAIR must be 0..6"
#endif
}  // namespace mc
