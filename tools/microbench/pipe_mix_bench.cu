// pipe_mix_bench.cu -- does anything issue "for free" next to the carry-chained wide multiply-adds that make up fe_mul?
// DESIGN.md "Roofline": the hot kernels keep the fmaheavy pipe ~88 % busy at 45 % issue-slot utilisation.  r1 showed that
// plain IMAD and ALU instructions both slow the wide chain down, r2 that DFMA shares the multiplier.  This bench measures
// the remaining candidates: FP32 FFMA (which can issue on the fmalite half of the FMA pipe), FP32 FMUL/FADD, and the
// half-precision HFMA2, each mixed K-to-2 with the wide multiply-add chain.  Output: one JSON object on stdout; rates are
// lane-operations per clock per SM (per_s / (SMs x probed SM clock)).
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>

#define U 8

__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

__global__ void k_clock_probe(double *mhz) {
    long long g0 = gtime_ns(), c0 = clock64();
    while (gtime_ns() - g0 < 1000000) { }
    long long g1 = gtime_ns(), c1 = clock64();
    *mhz = (double)(c1 - c0) * 1e3 / (double)(g1 - g0);
}

// WIDE: 1 = two carry-chained wide multiply-adds per (i, r);  FK: FFMA per (i, r);  HK: HFMA2 per (i, r);  AK: FADD per (i, r)
template <int WIDE, int FK, int HK, int AK>
__global__ void __launch_bounds__(256) k_mix(int iters, uint32_t a, uint32_t b, float fa, float fb, uint32_t *sink) {
    uint32_t x[U];
    uint64_t w[U];
    float f[U][4];
    __half2 h[U][2];
    for (int i = 0; i < U; i++) {
        x[i] = threadIdx.x * 7u + i;
        w[i] = ((uint64_t)threadIdx.x << 32) | (i + 1);
        for (int k = 0; k < 4; k++) f[i][k] = 1.0f + i + k + threadIdx.x;
        for (int k = 0; k < 2; k++) h[i][k] = __floats2half2_rn(1.0f + i, 0.5f + k);
    }
    const __half2 ha = __floats2half2_rn(fa, fa), hb = __floats2half2_rn(fb, fb);
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < U; i++) {
#pragma unroll
                for (int k = 0; k < FK; k++) f[i][k % 4] = fmaf(f[i][k % 4], fa, fb);
#pragma unroll
                for (int k = 0; k < AK; k++) f[i][k % 4] = f[i][k % 4] + fb;
#pragma unroll
                for (int k = 0; k < HK; k++) h[i][k % 2] = __hfma2(h[i][k % 2], ha, hb);
                if (WIDE) {
                    uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\t"
                                 "madc.lo.cc.u32 %0, %2, %4, %0;\n\tmadc.hi.u32 %1, %2, %4, %1;"
                                 : "+r"(lo), "+r"(hi) : "r"(x[i]), "r"(a), "r"(b));
                    w[i] = ((uint64_t)hi << 32) | lo;
                }
            }
        }
    }
    uint32_t acc = 0;
    for (int i = 0; i < U; i++) {
        acc ^= x[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32);
        for (int k = 0; k < 4; k++) acc ^= __float_as_uint(f[i][k]);
        for (int k = 0; k < 2; k++) acc ^= *reinterpret_cast<uint32_t *>(&h[i][k]);
    }
    if (acc == 0x12345678u) sink[0] = acc;
}

struct result { std::string name; double wide_per_s, other_per_s, ms; };

int main(int argc, char **argv) {
    int iters = 4096;
    if (argc > 1) iters = atoi(argv[1]);
    cudaSetDevice(0);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int sms = prop.multiProcessorCount;
    uint32_t *sink;
    cudaMalloc(&sink, 64);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<result> results;
    for (int w = 0; w < 40; w++) k_mix<1, 0, 0, 0><<<sms * 4, 256>>>(1 << 12, 0x9e3779b1u, 12345u, 1.0000001f, 0.5f, sink);
    cudaDeviceSynchronize();

    const int blocks = sms * 4, threads = 256;   // 32 warps per SM
    const double groups = (double)iters * 4 * U * threads * blocks;   // (i, r) groups executed per launch
    auto run = [&](const char *name, auto launch, int wide, int other) {
        launch();
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        results.push_back({name, groups * 2 * wide / (ms * 1e-3), groups * other / (ms * 1e-3), ms});
    };
#define MIX(W, FK, HK, AK, nm) run(nm, [&] { k_mix<W, FK, HK, AK><<<blocks, threads>>>(iters, 0x9e3779b1u, 12345u, 1.0000001f, 0.5f, sink); }, W, FK + HK + AK)
    MIX(1, 0, 0, 0, "wide_chain_alone");
    MIX(0, 4, 0, 0, "ffma_alone");
    MIX(0, 0, 2, 0, "hfma2_alone");
    MIX(0, 0, 0, 4, "fadd_alone");
    MIX(1, 1, 0, 0, "wide_chain_plus_1ffma_per_2");
    MIX(1, 2, 0, 0, "wide_chain_plus_2ffma_per_2");
    MIX(1, 4, 0, 0, "wide_chain_plus_4ffma_per_2");
    MIX(1, 8, 0, 0, "wide_chain_plus_8ffma_per_2");
    MIX(1, 0, 2, 0, "wide_chain_plus_2hfma2_per_2");
    MIX(1, 0, 4, 0, "wide_chain_plus_4hfma2_per_2");
    MIX(1, 0, 0, 2, "wide_chain_plus_2fadd_per_2");
    MIX(1, 0, 0, 4, "wide_chain_plus_4fadd_per_2");

    double *d_mhz, h_mhz = 0;
    cudaMalloc(&d_mhz, sizeof(double));
    cudaStream_t s2;
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    k_mix<1, 0, 0, 0><<<sms * 2, 256>>>(1 << 14, 0x9e3779b1u, 12345u, 1.0000001f, 0.5f, sink);
    k_clock_probe<<<1, 1, 0, s2>>>(d_mhz);
    cudaDeviceSynchronize();
    cudaMemcpy(&h_mhz, d_mhz, sizeof(double), cudaMemcpyDeviceToHost);

    const double denom = sms * h_mhz * 1e6;
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz_probed\": %.1f, \"iters\": %d, \"unit\": \"lane-ops per clock per SM\", \"tests\": {",
           prop.name, sms, h_mhz, iters);
    for (size_t i = 0; i < results.size(); i++)
        printf("%s\"%s\": {\"wide_mads\": %.2f, \"other\": %.2f, \"ms\": %.3f}", i ? ", " : "", results[i].name.c_str(),
               results[i].wide_per_s / denom, results[i].other_per_s / denom, results[i].ms);
    printf("}}\n");
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(ce)); return 1; }
    return 0;
}
