// Does data written by one kernel stay in L2 for the next one?  Kernel W writes X MB (plain / .cs / evict_last /
// evict_first stores), kernel R reads it back (plain / evict_first / evict_last loads); R's time and effective
// bandwidth tell whether it came from L2 or DRAM.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_retention l2_retention.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define EL 0x14F0000000000000ull
#define EF 0x12F0000000000000ull
template <int MODE>
__global__ void w(double2* p, size_t n)
{
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 v = make_double2((double)i, 1.0);
        if (MODE == 0) p[i] = v;
        if (MODE == 1) __stcs(p + i, v);
        if (MODE == 2) asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p + i), "d"(v.x), "d"(v.y), "l"(EL) : "memory");
        if (MODE == 3) asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p + i), "d"(v.x), "d"(v.y), "l"(EF) : "memory");
    }
}
template <int MODE>
__global__ void r(const double2* p, size_t n, double* out)
{
    double s = 0;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        double2 v;
        if (MODE == 0) v = __ldg(p + i);
        if (MODE == 1) asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p + i), "l"(EF));
        if (MODE == 2) asm volatile("ld.global.nc.L2::cache_hint.v2.f64 {%0, %1}, [%2], %3;" : "=d"(v.x), "=d"(v.y) : "l"(p + i), "l"(EL));
        s += v.x + v.y;
    }
    if (s == 12345.678) *out = s;
}
int main()
{
    const size_t maxb = 512ull << 20;
    double2* p;
    double* o;
    cudaMalloc(&p, maxb);
    cudaMalloc(&o, 8);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const int mbs[] = {16, 32, 48, 64, 80, 96, 112, 128, 192, 384};
    const char* wn[] = {"st", "st.cs", "st.evict_last", "st.evict_first"};
    const char* rn[] = {"ldg", "ld.evict_first", "ld.evict_last"};
    for (int wm = 0; wm < 4; wm++)
        for (int rm = 0; rm < 3; rm++) {
            if (rm == 1 && wm == 1) continue;
            printf("%-15s -> %-15s :", wn[wm], rn[rm]);
            for (int mb : mbs) {
                size_t n = ((size_t)mb << 20) / 16;
                float best = 1e9;
                for (int it = 0; it < 5; it++) {
                    if (wm == 0) w<0><<<148 * 8, 256>>>(p, n);
                    if (wm == 1) w<1><<<148 * 8, 256>>>(p, n);
                    if (wm == 2) w<2><<<148 * 8, 256>>>(p, n);
                    if (wm == 3) w<3><<<148 * 8, 256>>>(p, n);
                    cudaEventRecord(e0);
                    if (rm == 0) r<0><<<148 * 8, 256>>>(p, n, o);
                    if (rm == 1) r<1><<<148 * 8, 256>>>(p, n, o);
                    if (rm == 2) r<2><<<148 * 8, 256>>>(p, n, o);
                    cudaEventRecord(e1);
                    cudaEventSynchronize(e1);
                    float ms;
                    cudaEventElapsedTime(&ms, e0, e1);
                    if (ms < best) best = ms;
                }
                printf(" %dMB %.0fGB/s", mb, mb * 1.048576 / best);
            }
            printf("\n");
        }
    // ping-pong like the tracker: two buffers of X MB; frame f reads A (written at f-1) and writes B
    printf("ping-pong (read buf a, write buf b, swap), evict_last both / plain+cs:\n");
    for (int mode = 0; mode < 2; mode++)
        for (int mb : {16, 32, 48, 64, 96, 128}) {
            size_t n = ((size_t)mb << 20) / 16;
            double2 *a = p, *b = p + (256ull << 20) / 16;
            float tot = 0;
            for (int it = 0; it < 12; it++) {
                cudaEventRecord(e0);
                if (mode == 0) {
                    r<2><<<148 * 8, 256>>>(a, n, o);
                    w<2><<<148 * 8, 256>>>(b, n);
                } else {
                    r<0><<<148 * 8, 256>>>(a, n, o);
                    w<1><<<148 * 8, 256>>>(b, n);
                }
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (it >= 4) tot += ms;
                double2* t = a; a = b; b = t;
            }
            printf("  mode %d  %d MB per buffer: %.1f us per frame (%.0f GB/s r+w)\n", mode, mb, tot / 8 * 1e3, 2 * mb * 1.048576 / (tot / 8));
        }
    return 0;
}
