// sb_uc8.cu -- the step kernel instantiated for up to 8 mixture components (sb_core.cuh with SB_UC = 8), as
// its own translation unit so that it compiles in parallel with sb_api.cu (which holds the 4-component
// instantiation every BASELINE configuration uses).
#include <cstring>
#define SB_UC 8
#define SB_NS uc8
#include "sb_core.cuh"
#undef SB_UC
#undef SB_NS

const void *sb_uc8_run_kernel_ptr() {
#ifdef SB_HOSTSIM
    return nullptr;
#else
    return (const void *)uc8::sb_run_kernel;
#endif
}
void sb_uc8_run_kernel_host(const SbLaunch &L) {
#ifdef SB_HOSTSIM
    uc8::sb_run_kernel(L);
#else
    (void)L;
#endif
}
long sb_uc8_tiled_calls(int kind) {
#ifdef SB_HOSTSIM
    return uc8::g_tiled_calls[kind];
#else
    (void)kind; return 0;
#endif
}
