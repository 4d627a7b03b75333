// Microbenchmark: does a coarser per-slot chunk in the tile layout cut the DRAM read traffic of the
// gather-by-parent (only ~62 % of the parents are selected) enough to pay for its padding?
// Mimics k_slot_update's memory traffic (gather 90 doubles of the parent, write 90 doubles of the child)
// without the arithmetic.  CH = doubles per (slot, row) chunk: 2 = current layout, 4, 8.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>
#include <cuda_runtime.h>

template <int CH>
__global__ void __launch_bounds__(128, 2) k_gather(const double* __restrict__ in, double* __restrict__ out,
                                                   const int* __restrict__ parent, long long total, int N)
{
    constexpr int NCH = (90 + CH - 1) / CH;
    const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= total) return;
    const long long t = s / N;
    const long long sp = t * N + parent[s];
    const double* src = in + ((sp >> 5) * (long long)(NCH * 32) + (sp & 31)) * CH;
    double v[NCH * CH];
#pragma unroll
    for (int c = 0; c < NCH; c++) {
#pragma unroll
        for (int h = 0; h < CH / 2; h++) {
            const double2 q = __ldg(reinterpret_cast<const double2*>(src + (long long)c * 32 * CH) + h);
            v[c * CH + 2 * h] = q.x;
            v[c * CH + 2 * h + 1] = q.y;
        }
    }
    // a little dependent arithmetic so nothing is optimised away
    double acc = 0.0;
#pragma unroll
    for (int e = 0; e < NCH * CH; e++) acc += v[e];
    double* dst = out + ((s >> 5) * (long long)(NCH * 32) + (s & 31)) * CH;
#pragma unroll
    for (int c = 0; c < NCH; c++) {
#pragma unroll
        for (int h = 0; h < CH / 2; h++) {
            double2 q;
            q.x = v[c * CH + 2 * h] + acc * 1e-300;
            q.y = v[c * CH + 2 * h + 1];
            __stcs(reinterpret_cast<double2*>(dst + (long long)c * 32 * CH) + h, q);
        }
    }
}

template <int CH>
float run(const int* d_parent, long long total, int N, int reps)
{
    constexpr int NCH = (90 + CH - 1) / CH;
    const size_t bytes = (size_t)((total + 31) / 32) * NCH * 32 * CH * sizeof(double);
    double *a, *b;
    cudaMalloc(&a, bytes);
    cudaMalloc(&b, bytes);
    cudaMemset(a, 0, bytes);
    cudaMemset(b, 0, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    const unsigned grid = (unsigned)((total + 127) / 128);
    for (int i = 0; i < 3; i++) k_gather<CH><<<grid, 128>>>(a, b, d_parent, total, N);
    cudaEventRecord(e0);
    for (int i = 0; i < reps; i++) {
        k_gather<CH><<<grid, 128>>>(i & 1 ? b : a, i & 1 ? a : b, d_parent, total, N);
    }
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    cudaFree(a);
    cudaFree(b);
    return ms / reps;
}

int main()
{
    const int T = 4096, N = 500;
    const long long total = (long long)T * N;
    std::vector<int> parent(total), ident(total);
    std::mt19937_64 rng(1);
    std::lognormal_distribution<double> ln(0.0, 1.0); // ESS ~ N/2.7 like the steady-state filter
    double distinct = 0;
    for (int t = 0; t < T; t++) {
        std::vector<double> w(N);
        double sum = 0;
        for (auto& x : w) sum += (x = ln(rng));
        double u = std::uniform_real_distribution<double>(0, 1)(rng), c = 0;
        int k = 0;
        c = w[0] / sum;
        int last = -1;
        for (int i = 0; i < N; i++) {
            const double thr = (u + i) / N;
            while (c < thr && k < N - 1) c += w[++k] / sum;
            parent[(size_t)t * N + i] = k;
            ident[(size_t)t * N + i] = i;
            if (k != last) distinct += 1, last = k;
        }
    }
    printf("distinct parents: %.1f%%\n", 100.0 * distinct / total);
    int *d_par, *d_id;
    cudaMalloc(&d_par, total * 4);
    cudaMalloc(&d_id, total * 4);
    cudaMemcpy(d_par, parent.data(), total * 4, cudaMemcpyHostToDevice);
    cudaMemcpy(d_id, ident.data(), total * 4, cudaMemcpyHostToDevice);
    size_t gran = 0;
    cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
    printf("L2 fetch granularity limit: %zu\n", gran);
    for (int pass = 0; pass < 2; pass++) {
        if (pass == 1) {
            cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, 32);
            cudaDeviceGetLimit(&gran, cudaLimitMaxL2FetchGranularity);
            printf("-- after requesting 32 B: %zu\n", gran);
        }
        printf("CH=2 (16 B/slot/row, 720 B/slot): gather %.1f us  identity %.1f us\n", 1e3 * run<2>(d_par, total, N, 20), 1e3 * run<2>(d_id, total, N, 20));
        printf("CH=4 (32 B, 736 B/slot)         : gather %.1f us  identity %.1f us\n", 1e3 * run<4>(d_par, total, N, 20), 1e3 * run<4>(d_id, total, N, 20));
        printf("CH=8 (64 B, 768 B/slot)         : gather %.1f us  identity %.1f us\n", 1e3 * run<8>(d_par, total, N, 20), 1e3 * run<8>(d_id, total, N, 20));
    }
    return 0;
}
