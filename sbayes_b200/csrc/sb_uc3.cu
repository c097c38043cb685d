// sb_uc3.cu -- the step kernel instantiated for problems with EXACTLY three mixture components (cluster effect, universal,
// one more confounder: every BASELINE configuration): sb_core.cuh with SB_UC = 3 and SB_EXACT_C, so that the component count
// is a compile-time constant and the guards of the unrolled component loops fold away.  Its own translation unit: it compiles in
// parallel with sb_api.cu (at most 4 components) and sb_uc8.cu (at most 8).
#include <cstring>
#define SB_UC 3
#define SB_NS uc3
#define SB_EXACT_C 1
#include "sb_core.cuh"
#undef SB_EXACT_C
#undef SB_UC
#undef SB_NS

const void *sb_uc3_run_kernel_ptr() {
#ifdef SB_HOSTSIM
    return nullptr;
#else
    return (const void *)uc3::sb_run_kernel;
#endif
}
void sb_uc3_run_kernel_host(const SbLaunch &L) {
#ifdef SB_HOSTSIM
    uc3::sb_run_kernel(L);
#else
    (void)L;
#endif
}
long sb_uc3_tiled_calls(int kind) {
#ifdef SB_HOSTSIM
    return uc3::g_tiled_calls[kind];
#else
    (void)kind; return 0;
#endif
}
