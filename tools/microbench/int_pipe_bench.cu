// int_pipe_bench.cu -- measures the INT32 multiply-add issue rates of the device this library runs on.
// The hot kernels of this library are bound by the integer pipes (DESIGN.md "Roofline"); MEASURED_PEAKS.json only
// holds HBM and bf16 numbers, so the denominator of `roofline.frac` is measured here, live, by bench.py.
//
// Each test runs `U` independent dependency chains per thread for `iters` iterations, one CTA of 256 threads per
// SM x `ctas_per_sm`, and reports lane-operations per clock per SM (from clock64 inside the kernel) and per second
// (from CUDA events).  Output: one JSON object on stdout.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <vector>
#include <string>
#include "../../elastic_elgamal_b200/csrc/fe.cuh"

#define U 8

__device__ __forceinline__ long long gtime_ns() { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }

// SM clock actually in effect: clock64 ticks per globaltimer nanosecond over a ~1 ms spin (thread 0 of block 0)
__global__ void k_clock_probe(double *mhz) {
    long long g0 = gtime_ns(), c0 = clock64();
    while (gtime_ns() - g0 < 1000000) { }
    long long g1 = gtime_ns(), c1 = clock64();
    *mhz = (double)(c1 - c0) * 1e3 / (double)(g1 - g0);
}

template <int KIND>
__global__ void __launch_bounds__(256) k_pipe(int iters, uint32_t a, uint32_t b, uint32_t *sink, long long *cycles) {
    uint32_t x[U];
    uint64_t w[U];
    double d[U];
    for (int i = 0; i < U; i++) { x[i] = threadIdx.x * 7u + i; w[i] = ((uint64_t)threadIdx.x << 32) | (i + 1); d[i] = 1.0 + i + threadIdx.x; }
    double da = 1.0000001, db = 0.5;
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
#pragma unroll
            for (int i = 0; i < U; i++) {
                if (KIND == 0) x[i] = x[i] * a + b;                                   // IMAD
                if (KIND == 1) asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[i]) : "r"(a), "r"(b));   // IMAD.HI
                if (KIND == 2) w[i] = (uint64_t)(uint32_t)w[i] * a + w[i];            // IMAD.WIDE
                if (KIND == 3) x[i] = x[i] + a + (x[i] >> 3);                         // IADD3 / SHF mix on the ALU pipe
                if (KIND == 4) { x[i] = x[i] * a + b; w[i] += x[(i + 1) % U] ^ w[i]; }          // IMAD + ALU
                if (KIND == 5) d[i] = fma(d[i], da, db);                              // DFMA
                if (KIND == 6) { w[i] = (uint64_t)(uint32_t)w[i] * a + w[i]; d[i] = fma(d[i], da, db); }  // IMAD.WIDE + DFMA
                if (KIND == 8 || KIND == 10) x[i] = x[i] * a + b;                      // + 1 IMAD per 2 wide mads
                if (KIND == 9 || KIND == 10) x[i] = (x[i] ^ a) + (x[i] >> 5);          // + 3 ALU ops (LOP3, SHF, IADD3) per 2 wide mads
                if (KIND == 11) { x[i] = (x[i] ^ a) + (x[i] >> 5); x[i] = (x[i] ^ b) + (x[i] >> 7); }   // + 6 ALU ops
                if (KIND >= 7) {                                                       // carry-chained wide mads (as in fe_mul)
                    uint32_t lo = (uint32_t)w[i], hi = (uint32_t)(w[i] >> 32);
                    asm volatile("mad.lo.cc.u32 %0, %2, %3, %0;\n\tmadc.hi.cc.u32 %1, %2, %3, %1;\n\t"
                                 "madc.lo.cc.u32 %0, %2, %4, %0;\n\tmadc.hi.u32 %1, %2, %4, %1;"
                                 : "+r"(lo), "+r"(hi) : "r"(x[i]), "r"(a), "r"(b));
                    w[i] = ((uint64_t)hi << 32) | lo;
                }
            }
        }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < U; i++) acc ^= x[i] ^ (uint32_t)w[i] ^ (uint32_t)(w[i] >> 32) ^ (uint32_t)d[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

static __device__ __noinline__ eg::fe fe_mul_sf_call(const eg::fe a, const eg::fe b) {
    eg::fe r;
#if defined(__CUDA_ARCH__)
    eg::fe_mul_ptx_sf(r, a, b);
#else
    eg::fe_mul_portable(r, a, b);
#endif
    return r;
}
static __device__ __noinline__ eg::fe fe_mul_k_call(const eg::fe a, const eg::fe b) {
    eg::fe r;
#if defined(__CUDA_ARCH__)
    eg::fe_mul_ptx_k(r, a, b);
#else
    eg::fe_mul_portable(r, a, b);
#endif
    return r;
}
struct fe_pair { eg::fe x, y; };
// two independent multiplications / squarings in one call: does the extra instruction-level parallelism pay?
static __device__ __noinline__ fe_pair fe_mul2_call(const eg::fe a0, const eg::fe b0, const eg::fe a1, const eg::fe b1) {
    fe_pair r = {};
#if defined(__CUDA_ARCH__)
    eg::fe_mul_ptx(r.x, a0, b0);
    eg::fe_mul_ptx(r.y, a1, b1);
#endif
    return r;
}
static __device__ __noinline__ fe_pair fe_sq2_call(const eg::fe a0, const eg::fe a1) {
    fe_pair r = {};
#if defined(__CUDA_ARCH__)
    eg::fe_sq_ptx(r.x, a0);
    eg::fe_sq_ptx(r.y, a1);
#endif
    return r;
}
static __device__ __noinline__ eg::fe fe_sq_sf_call(const eg::fe a) {
    eg::fe r;
#if defined(__CUDA_ARCH__)
    eg::fe_sq_ptx_sf(r, a);
#else
    eg::fe_sq_portable(r, a);
#endif
    return r;
}

// dependent chain of field multiplications / squarings per thread (the unit of the library's roofline)
template <int KIND>
__global__ void __launch_bounds__(128) k_field(int iters, const uint32_t *seed, uint32_t *sink, long long *cycles) {
    eg::fe x, y;
    for (int i = 0; i < 8; i++) { x.v[i] = seed[i] + threadIdx.x * 977u + blockIdx.x; y.v[i] = seed[8 + i] ^ (threadIdx.x << 3); }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
        if (KIND == 0) { eg::fe_mul(x, x, y); eg::fe_mul(y, y, x); }
        if (KIND == 1) { eg::fe_sq(x, x); eg::fe_sq(y, y); }
        if (KIND == 2) { eg::fe_mul_portable(x, x, y); eg::fe_mul_portable(y, y, x); }
        if (KIND == 3) { eg::fe_sq_portable(x, x); eg::fe_sq_portable(y, y); }
        if (KIND == 4) { eg::fe_add(x, x, y); eg::fe_sub(y, y, x); }
        if (KIND == 5) { x = fe_mul_sf_call(x, y); y = fe_mul_sf_call(y, x); }
        if (KIND == 6) { x = fe_sq_sf_call(x); y = fe_sq_sf_call(y); }
        if (KIND == 7) { x = fe_mul_k_call(x, y); y = fe_mul_k_call(y, x); }
        if (KIND == 8) { fe_pair q = fe_mul2_call(x, y, y, x); x = q.x; y = q.y; }
        if (KIND == 9) { fe_pair q = fe_sq2_call(x, y); x = q.x; y = q.y; }
    }
    long long t1 = clock64();
    uint32_t acc = 0;
    for (int i = 0; i < 8; i++) acc ^= x.v[i] ^ y.v[i];
    if (acc == 0x12345678u) sink[0] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

struct result { std::string name; double ops_per_clk_sm, ops_per_s, ms; };

int main(int argc, char **argv) {
    int dev = 0, iters = 4096;
    if (argc > 1) iters = atoi(argv[1]);
    cudaSetDevice(dev);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, dev);
    const int sms = prop.multiProcessorCount;
    uint32_t *sink, *seed;
    long long *cycles;
    cudaMalloc(&sink, 64); cudaMalloc(&seed, 64); cudaMalloc(&cycles, sizeof(long long) * sms * 16);
    uint32_t hseed[16];
    for (int i = 0; i < 16; i++) hseed[i] = 0x9e3779b9u * (i + 1);
    cudaMemcpy(seed, hseed, 64, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    std::vector<result> results;

    // spin for ~0.4 s first so that the SM clock has ramped up from idle before anything is measured
    for (int w = 0; w < 40; w++) k_pipe<0><<<sms * 4, 256>>>(1 << 15, 0x9e3779b1u, 12345u, sink, cycles);
    cudaDeviceSynchronize();

    auto run = [&](const char *name, auto launch, int blocks, int threads, double lane_ops_per_thread) {
        launch(blocks, threads);   // warm-up
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        launch(blocks, threads);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        std::vector<long long> h(blocks);
        cudaMemcpy(h.data(), cycles, sizeof(long long) * blocks, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (long long c : h) avg += (double)c;
        avg /= blocks;
        const double ctas_per_sm = (double)blocks / sms;
        double ops_sm = lane_ops_per_thread * threads * ctas_per_sm;      // all CTAs of an SM run concurrently
        results.push_back({name, ops_sm / avg, lane_ops_per_thread * threads * (double)blocks / (ms * 1e-3), ms});
    };
    const double pipe_ops = (double)iters * 4 * U;
    const int cps = 4;   // 4 CTAs x 256 threads = 32 warps per SM
#define PIPE(kind, nm, mult) run(nm, [&](int bl, int th) { k_pipe<kind><<<bl, th>>>(iters, 0x9e3779b1u, 12345u, sink, cycles); }, sms * cps, 256, pipe_ops * mult)
    PIPE(0, "imad", 1);
    PIPE(1, "imad_hi", 1);
    PIPE(2, "imad_wide", 1);
    PIPE(3, "alu_iadd3_shf", 1);
    PIPE(4, "imad_plus_alu(imad count)", 1);
    PIPE(5, "dfma", 1);
    PIPE(6, "imad_wide_plus_dfma(pairs)", 1);
    PIPE(7, "mad_wide_carry_chain(wide mads)", 2);
    PIPE(8, "wide_chain_plus_imad_2to1(wide mads)", 2);
    PIPE(9, "wide_chain_plus_3alu_per_2(wide mads)", 2);
    PIPE(10, "wide_chain_plus_imad_plus_3alu(wide mads)", 2);
    PIPE(11, "wide_chain_plus_6alu_per_2(wide mads)", 2);
    const int fiters = iters / 4;
#define FIELD(kind, nm, cps_) run(nm, [&](int bl, int th) { k_field<kind><<<bl, th>>>(fiters, seed, sink, cycles); }, sms * cps_, 128, (double)fiters * 2)
    FIELD(0, "fe_mul_ptx_w16", 4);
    FIELD(1, "fe_sq_ptx_w16", 4);
    FIELD(2, "fe_mul_portable_w16", 4);
    FIELD(3, "fe_sq_portable_w16", 4);
    FIELD(4, "fe_addsub_w16", 4);
    FIELD(5, "fe_mul_shiftfold_w16", 4);
    FIELD(6, "fe_sq_shiftfold_w16", 4);
    FIELD(8, "fe_mul2_pair_w8", 2);
    FIELD(8, "fe_mul2_pair_w12", 3);
    FIELD(8, "fe_mul2_pair_w16", 4);
    FIELD(9, "fe_sq2_pair_w8", 2);
    FIELD(9, "fe_sq2_pair_w12", 3);
    FIELD(9, "fe_sq2_pair_w16", 4);
    FIELD(7, "fe_mul_karatsuba_w16", 4);
    FIELD(7, "fe_mul_karatsuba_w20", 5);
    FIELD(7, "fe_mul_karatsuba_w32", 8);
    FIELD(5, "fe_mul_shiftfold_w20", 5);
    FIELD(6, "fe_sq_shiftfold_w20", 5);
    FIELD(0, "fe_mul_ptx_w20", 5);
    FIELD(1, "fe_sq_ptx_w20", 5);
    FIELD(0, "fe_mul_ptx_w8", 2);
    FIELD(1, "fe_sq_ptx_w8", 2);
    FIELD(0, "fe_mul_ptx_w32", 8);
    FIELD(1, "fe_sq_ptx_w32", 8);

    // SM clock in effect while the integer pipe is loaded: probe kernel on a second stream next to a long IMAD launch
    double *d_mhz, h_mhz = 0;
    cudaMalloc(&d_mhz, sizeof(double));
    cudaStream_t s2;
    cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking);
    k_pipe<0><<<sms * 2, 256>>>(1 << 17, 0x9e3779b1u, 12345u, sink, cycles);
    k_clock_probe<<<1, 1, 0, s2>>>(d_mhz);
    cudaDeviceSynchronize();
    cudaMemcpy(&h_mhz, d_mhz, sizeof(double), cudaMemcpyDeviceToHost);

    int clock_khz = 0;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, dev);
    // `per_s` (CUDA events around the launch) is the rate the roofline uses; `per_clk_per_sm` = per_s / (SMs x the probed
    // SM clock); `per_clk_per_sm_clock64` is the in-kernel clock64 estimate (kept for comparison only: it assumes that
    // all CTAs of an SM are co-resident for the whole launch).
    printf("{\"device\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f, \"sm_clock_mhz_probed\": %.1f, \"iters\": %d, \"tests\": {",
           prop.name, sms, clock_khz / 1000.0, h_mhz, iters);
    for (size_t i = 0; i < results.size(); i++)
        printf("%s\"%s\": {\"per_s\": %.4e, \"per_clk_per_sm\": %.2f, \"per_clk_per_sm_clock64\": %.2f, \"ms\": %.3f}", i ? ", " : "",
               results[i].name.c_str(), results[i].ops_per_s, h_mhz > 0 ? results[i].ops_per_s / (sms * h_mhz * 1e6) : 0.0,
               results[i].ops_per_clk_sm, results[i].ms);
    printf("}}\n");
    cudaError_t ce = cudaGetLastError();
    if (ce != cudaSuccess) { fprintf(stderr, "CUDA error: %s\n", cudaGetErrorString(ce)); return 1; }
    return 0;
}
