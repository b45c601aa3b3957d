// Microbenchmark: the memory access pattern of one cfg2 step WITHOUT the arithmetic, as a plain high-occupancy kernel.
// Per (env, EV) slot: read action f32, soc f64, hours_left f32, soh f64, one soc_deg ring row element f64, one 32-byte
// schedule record gathered at a random time index per env; write soc, hours_left, the next ring row element; per env
// write a D-float observation row (per-EV terms from the slot threads, header gathered from a table row).
// What it answers: which fraction of the measured HBM peak this PATTERN reaches when nothing else is in the way.
// build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o stream_pattern stream_pattern.cu ; run: ./stream_pattern
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <vector>

constexpr int N = 50, D = 388, H = 38, R = 16, T = 35040;
struct __align__(16) Rec { double sr; float tl; int flags; float a0, a1, a2, a3; };

template <int kVariant>
__global__ void __launch_bounds__(256) pattern(int E, const float* __restrict__ act, double* soc, float* hl, const double* __restrict__ soh,
                                              double* hist, const Rec* __restrict__ rec, const float* __restrict__ hdr, const int* __restrict__ tix,
                                              float* obs, int k) {
    const int B = 256 / N;                                   // envs per CTA tile, like the product kernel
    const int ntiles = (E + B - 1) / B;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int j = threadIdx.x;
        const int b = j / N, n = j - b * N, e = tile * B + b;
        if (j < B * N && e < E) {
            const size_t i = (size_t)e * N + n;
            const int t = tix[e];
            const float a = __ldcs(act + i);
            double s = __ldcs(soc + i);
            float h = __ldcs(hl + i);
            const double so = __ldcs(soh + i);
            const size_t EN_ = (size_t)E * N;
            // kVariant & 16: ring rows as planes [R][E][N] indexed by the GLOBAL step (contiguous like soc) instead of [E][R][N]
            const size_t h_rd = (kVariant & 16) ? (size_t)(k & (R - 1)) * EN_ + i : ((size_t)e * R + ((k + e) & (R - 1))) * N + n;
            const size_t h_wr = (kVariant & 16) ? (size_t)((k + 1) & (R - 1)) * EN_ + i : ((size_t)e * R + ((k + e + 1) & (R - 1))) * N + n;
            const double sd = (kVariant & 4) ? 0.0 : __ldcs(hist + h_rd);
            int4 r0 = make_int4(0, 0, 0, 0); float4 r1 = make_float4(0, 0, 0, 0);
            if (!(kVariant & 1)) {
                if (kVariant & 8) r0 = __ldg(reinterpret_cast<const int4*>(rec) + (size_t)t * N + n);      // 16-byte records
                else {
                    r0 = __ldg(reinterpret_cast<const int4*>(rec + (size_t)t * N + n));
                    r1 = __ldg(reinterpret_cast<const float4*>(rec + (size_t)t * N + n) + 1);
                }
            }
            s = s + (double)a * 1e-3 * so; h = h + __int_as_float(r0.z) * 1e-6f;
            __stcs(soc + i, s);
            __stcs(hl + i, h);
            if (!(kVariant & 4)) __stcs(hist + h_wr, sd + __hiloint2double(r0.y, r0.x) * 1e-9);
            else if (sd + __hiloint2double(r0.y, r0.x) == 12345.678) __stcs(soc + i, 0.0);
            float* o = obs + (size_t)e * D;
            if (kVariant & 2) { if (r1.x == 12345.f) __stcs(o, r1.y + r1.z + r1.w + (float)r0.w); continue; }
            __stcs(o + n, (float)s); __stcs(o + N + n, h);
            float* ax = o + 2 * N + 28 + n;
            __stcs(ax, (float)r0.w); __stcs(ax + N, r1.x); __stcs(ax + 2 * N, r1.y); __stcs(ax + 3 * N, r1.z); __stcs(ax + 4 * N, r1.w);
            if (n < H && !(kVariant & 1)) __stcs(o + (n < 28 ? 2 * N + n : 2 * N + 28 + 5 * N + (n - 28)), __ldg(hdr + (size_t)t * 40 + n));
        }
    }
}

int main(int argc, char** argv) {
    const int E = argc > 1 ? atoi(argv[1]) : 65536;
    const size_t EN = (size_t)E * N;
    float *act, *hl, *hdr, *obs; double *soc, *soh, *hist; Rec* rec; int* tix;
    cudaMalloc(&act, EN * 4 * 8); cudaMalloc(&hl, EN * 4); cudaMalloc(&soc, EN * 8); cudaMalloc(&soh, EN * 8);
    cudaMalloc(&hist, EN * R * 8); cudaMalloc(&rec, (size_t)T * N * sizeof(Rec)); cudaMalloc(&hdr, (size_t)T * 40 * 4);
    cudaMalloc(&obs, (size_t)E * D * 4); cudaMalloc(&tix, E * 4);
    cudaMemset(act, 0, EN * 4 * 8); cudaMemset(hl, 0, EN * 4); cudaMemset(soc, 0, EN * 8); cudaMemset(soh, 0, EN * 8);
    cudaMemset(hist, 0, EN * R * 8); cudaMemset(rec, 0, (size_t)T * N * sizeof(Rec)); cudaMemset(hdr, 0, (size_t)T * 40 * 4);
    std::vector<int> t(E);
    srand(1);
    for (int e = 0; e < E; e++) t[e] = rand() % (T - 200);
    cudaMemcpy(tix, t.data(), E * 4, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double alg = (52.0 + (4.0 * D + 13.0) / N) * EN;
    int sms = 148;
    auto run = [&](auto kern, const char* name, double bytes) {
        const int ntiles = (E + 4) / 5;
        for (int w = 0; w < 5; w++) kern<<<ntiles, 256>>>(E, act + (w % 8) * EN, soc, hl, soh, hist, rec, hdr, tix, obs, w);
        cudaEventRecord(e0);
        const int K = 50;
        for (int k = 0; k < K; k++) kern<<<ntiles, 256>>>(E, act + (k % 8) * EN, soc, hl, soh, hist, rec, hdr, tix, obs, k);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("  %-58s %.1f us per launch, %.0f GB/s of its own %.0f MB\n", name, ms * 1e3 / K, bytes / (ms * 1e-3 / K) / 1e9, bytes / 1e6);
    };
    printf("one CTA per tile, variants of the pattern (E=%d):\n", E);
    run(pattern<0>, "full pattern", alg);
    run(pattern<1>, "without the schedule-record / header gathers", alg);
    run(pattern<2>, "without the observation writes", alg - 4.0 * D * E);
    run(pattern<4>, "without the soc_deg ring row (read + write)", alg - 16.0 * EN);
    run(pattern<8>, "16-byte schedule records (no auxiliary terms)", alg);
    run(pattern<16>, "ring rows as [R][E][N] planes indexed by the global step", alg);
    run(pattern<24>, "both", alg);
    run(pattern<7>, "state streams only (action, soc, hours_left, soh)", 36.0 * EN);
    for (int ctas_per_sm : {2, 4, 8, 0}) {
        const int ntiles = (E + 4) / 5;
        const int grid = ctas_per_sm ? sms * ctas_per_sm : ntiles;
        for (int w = 0; w < 5; w++) pattern<0><<<grid, 256>>>(E, act + (w % 8) * EN, soc, hl, soh, hist, rec, hdr, tix, obs, w);
        cudaEventRecord(e0);
        const int K = 50;
        for (int k = 0; k < K; k++) pattern<0><<<grid, 256>>>(E, act + (k % 8) * EN, soc, hl, soh, hist, rec, hdr, tix, obs, k);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("E=%d grid=%d (%s): %.1f us per launch, %.0f GB/s algorithmic (B_alg 83.3 B/EV-step) = %.1f %% of 6549 GB/s; err=%s\n", E, grid,
               ctas_per_sm ? "persistent" : "one CTA per tile", ms * 1e3 / K, alg / (ms * 1e-3 / K) / 1e9, 100 * alg / (ms * 1e-3 / K) / 1e9 / 6549.4,
               cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
