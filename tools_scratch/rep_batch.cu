// reproduction experiment for batch_items on sm_100a
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
struct Prod { double m; int e; };
struct Marg { Prod p0, p1; double geo; int on; };
__device__ __forceinline__ Prod prod_one() { Prod p; p.m = 1.0; p.e = 0; return p; }
__device__ __forceinline__ Prod wred(Prod p) {
    for (int off = 16; off > 0; off >>= 1) { p.m *= __shfl_sync(0xffffffffu, p.m, (threadIdx.x & 31) ^ off); p.e += __shfl_sync(0xffffffffu, p.e, (threadIdx.x & 31) ^ off); }
    return p;
}
template <class Acc, class WarpFn, class LaneFn>
__device__ __forceinline__ void batch_items(int lo, int hi, WarpFn warp_fn, LaneFn lane_fn) {
    const int gw = threadIdx.x >> 5, nw = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (int i0 = lo + gw; i0 < hi; i0 += nw * 32) {
        Acc mine; bool have = false;
        for (int i = 0; i < 32; i++) {
            int it = i0 + nw * i;
            if (it >= hi) break;
            Acc a = warp_fn(it);
#ifdef VARIANT_SELECT
            mine.on = lane == i ? a.on : mine.on; mine.geo = lane == i ? a.geo : mine.geo;
            mine.p0.m = lane == i ? a.p0.m : mine.p0.m; mine.p0.e = lane == i ? a.p0.e : mine.p0.e;
            mine.p1.m = lane == i ? a.p1.m : mine.p1.m; mine.p1.e = lane == i ? a.p1.e : mine.p1.e;
            have = have || lane == i;
#else
            if (lane == i) { mine = a; have = true; }
#endif
        }
        if (have) lane_fn(i0 + nw * lane, mine);
    }
}
template <class Sel, class Out>
__device__ __forceinline__ void pass(const double *tab, int n, bool flat, Sel sel, Out out) {
    batch_items<Marg>(0, n,
        [&](int o) {
            Marg a; a.on = sel(o) ? 1 : 0; a.geo = 0.0; a.p0 = prod_one(); a.p1 = prod_one();
            if (a.on && !flat) { a.p0.m = tab[(o * 7 + (threadIdx.x & 31)) & 1023]; a.p0 = wred(a.p0); a.p1 = wred(a.p0); }
            return a;
        },
        [&](int o, const Marg &a) { out(o, a.on != 0, a.on ? a.p0.m : 0.0); });
    __syncthreads();
}
__global__ void k(unsigned char *flag, const double *tab, int n, int flat, int *out, double *cnt) {
    for (int o = threadIdx.x; o < n; o += blockDim.x) flag[o] = (o * 2654435761u >> 7) & 1;
    __syncthreads();
    double ex[2] = {0, 0};
    pass(tab, n, flat != 0, [&](int o) { return flag[o] != 0; }, [&](int o, bool on, double v) { out[o] = on; if (on) ex[0] += 1.0; });
    atomicAdd(cnt, ex[0]);
}
int main() {
    int n = 40; unsigned char *flag; double *tab, *cnt; int *out;
    cudaMalloc(&flag, 4096); cudaMalloc(&tab, 8192); cudaMalloc(&out, 4096 * 4); cudaMalloc(&cnt, 8);
    cudaMemset(tab, 0x3f, 8192); 
    for (int flat = 0; flat < 2; flat++) for (int nn : {40, 200, 1000}) {
        n = nn; cudaMemset(cnt, 0, 8);
        k<<<1, 512>>>(flag, tab, n, flat, out, cnt);
        cudaDeviceSynchronize();
        int h[4096]; unsigned char f[4096]; double c;
        cudaMemcpy(h, out, n * 4, cudaMemcpyDeviceToHost); cudaMemcpy(f, flag, n, cudaMemcpyDeviceToHost); cudaMemcpy(&c, cnt, 8, cudaMemcpyDeviceToHost);
        int bad = 0, tot = 0; for (int i = 0; i < n; i++) { bad += h[i] != f[i]; tot += f[i]; }
        printf("flat=%d n=%d mismatches=%d cnt=%g expected=%d err=%s\n", flat, n, bad, c, tot, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
