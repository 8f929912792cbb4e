// fp64_field_bench.cu -- can the FP64 pipe carry field multiplications next to the INT32 multiplier pipe?
//
// VERDICT r1 "weak" #2: k_ring sits at the fmaheavy ceiling of its integer field layer (88 % pipe-busy, 46 % issue slots) and
// the FP64 pipe (DFMA 62 lanes/clk/SM on B200) is idle.  This microbenchmark measures, in field multiplications per clock
// per SM, (a) the production integer fe_mul / fe_sq (fe.cuh, as non-inlined calls, like the hot kernels), (b) a GF(2^255-19)
// multiplication on the FP64 pipe: 12 limbs of 21.25 bits carried as doubles WITH their weights (limb k is a multiple of
// 2^ceil(21.25 k)), 144 DFMA + 12 DMUL for the x19 wrap + a two-chain carry of 3 DADD per limb -- every partial sum stays
// below 2^53 so the arithmetic is exact -- and (c) both at once, J of every 4 warps of each SM sub-partition on the FP64
// path.  Each warp runs dependent chains until a clock deadline and reports how many multiplications it finished, so the
// rates are steady-state co-running rates.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o fp64_field_bench fp64_field_bench.cu && ./fp64_field_bench
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda_runtime.h>

#include "../../elastic_elgamal_b200/csrc/fe.cuh"

using namespace eg;

// ---- FP64 field element: 12 doubles, limb k a multiple of 2^W[k] with |limb| <~ 2^(W[k+1]-1)
struct fd { double v[12]; };

__device__ __forceinline__ constexpr int fd_w(int k) { return (85 * k + 3) / 4; }     // ceil(21.25 k): 0,22,43,64,85,107,...,234,255

__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(1023 + e) << 52); }

// carry limb k into limb k+1 (k = 11 wraps into limb 0 with 19 * 2^-255)
template <int K>
__device__ __forceinline__ void fd_carry(double c[12]) {
    const double M = 3.0 * pow2(51 + fd_w(K + 1));
    const double t = (c[K] + M) - M;                     // c[K] rounded to a multiple of 2^W[K+1]
    c[K] -= t;
    if (K == 11) c[0] = fma(t, 19.0 * pow2(-255), c[0]);
    else c[(K + 1) % 12] += t;
}

__device__ __forceinline__ void fd_reduce(double c[12]) {
    // two interleaved chains (limbs 0..5 and 6..11), then one more step at the two seams
    fd_carry<0>(c); fd_carry<6>(c);
    fd_carry<1>(c); fd_carry<7>(c);
    fd_carry<2>(c); fd_carry<8>(c);
    fd_carry<3>(c); fd_carry<9>(c);
    fd_carry<4>(c); fd_carry<10>(c);
    fd_carry<5>(c); fd_carry<11>(c);
    fd_carry<6>(c); fd_carry<0>(c);
}

__device__ __noinline__ fd fd_mul_v(fd a, fd b) {
    fd r;
    double b19[12], c[12];
    const double k19 = 19.0 * pow2(-255);
#pragma unroll
    for (int j = 0; j < 12; j++) b19[j] = b.v[j] * k19;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        double s = 0.0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            const int j = k - i;
            if (j >= 0) s = fma(a.v[i], b.v[j], s);
            else s = fma(a.v[i], b19[j + 12], s);
        }
        c[k] = s;
    }
    fd_reduce(c);
#pragma unroll
    for (int k = 0; k < 12; k++) r.v[k] = c[k];
    return r;
}
__device__ __forceinline__ void fd_mul(fd &r, const fd &a, const fd &b) { r = fd_mul_v(a, b); }

__device__ __noinline__ fd fd_sq_v(fd a) {
    fd r;
    double a2[12], a19[12], c[12];
    const double k19 = 19.0 * pow2(-255);
#pragma unroll
    for (int j = 0; j < 12; j++) { a2[j] = a.v[j] + a.v[j]; a19[j] = a.v[j] * k19; }
#pragma unroll
    for (int k = 0; k < 12; k++) {
        double s = 0.0;
        // pairs (i, j), i < j, i + j = k (mod 12): 2 a_i a_j ; squares a_i^2 when 2 i = k (mod 12)
#pragma unroll
        for (int i = 0; i < 12; i++) {
#pragma unroll
            for (int j = i; j < 12; j++) {
                if ((i + j) % 12 != k) continue;
                const bool wrap = i + j >= 12;
                const double x = (i == j) ? a.v[i] : a2[i];
                s = fma(x, wrap ? a19[j] : a.v[j], s);
            }
        }
        c[k] = s;
    }
    fd_reduce(c);
#pragma unroll
    for (int k = 0; k < 12; k++) r.v[k] = c[k];
    return r;
}
__device__ __forceinline__ void fd_sq(fd &r, const fd &a) { r = fd_sq_v(a); }

// exact conversion from / to little-endian bits (test only; slow path)
__device__ void fd_from_u32(fd &r, const uint32_t w[8]) {
    for (int k = 0; k < 12; k++) {
        const int lo = fd_w(k), hi = fd_w(k + 1);
        uint64_t v = 0;
        for (int bit = lo; bit < hi && bit < 256; bit++) v |= (uint64_t)((w[bit >> 5] >> (bit & 31)) & 1u) << (bit - lo);
        r.v[k] = (double)v * pow2(lo);
    }
}

// value mod p as 8 words (via 64-bit integer limbs; test only)
__device__ void fd_to_u32(uint32_t w[8], const fd &a) {
    // limbs may be negative: accumulate into a signed 320-bit two's-complement number, then reduce mod p with fe code
    long long limb[12];
    for (int k = 0; k < 12; k++) limb[k] = (long long)(a.v[k] * pow2(-fd_w(k)));
    // add a multiple of p large enough to make everything positive: work with fe arithmetic instead
    fe acc = fe_zero();
    for (int k = 11; k >= 0; k--) {
        // acc = acc * 2^(W[k+1]-W[k]) + limb[k]
        const int sh = fd_w(k + 1) - fd_w(k);
        fe m = fe_zero(); m.v[0] = 1u << sh;
        fe t; fe_mul(t, acc, m);
        fe l = fe_zero();
        const long long v = limb[k];
        const unsigned long long mag = v < 0 ? (unsigned long long)(-v) : (unsigned long long)v;
        l.v[0] = (uint32_t)mag; l.v[1] = (uint32_t)(mag >> 32);
        if (v < 0) fe_sub(acc, t, l); else fe_add(acc, t, l);
    }
    fe_towords(w, acc);
}

// ---- correctness: FP64 product == integer product on random operands
__global__ void k_check(unsigned long long seed, unsigned long long *bad, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t wa[8], wb[8];
    unsigned long long s = seed + 0x9E3779B97F4A7C15ull * (unsigned long long)(i + 1);
    for (int k = 0; k < 8; k++) { s = s * 6364136223846793005ull + 1442695040888963407ull; wa[k] = (uint32_t)(s >> 32); s = s * 6364136223846793005ull + 1442695040888963407ull; wb[k] = (uint32_t)(s >> 32); }
    wa[7] &= 0x7fffffffu; wb[7] &= 0x7fffffffu;
    if (i % 7 == 0) for (int k = 0; k < 8; k++) wa[k] = 0xffffffffu >> (k == 7);
    if (i % 11 == 0) for (int k = 0; k < 8; k++) wb[k] = 0xffffffffu >> (k == 7);
    fe a, b, p, q;
    for (int k = 0; k < 8; k++) { a.v[k] = wa[k]; b.v[k] = wb[k]; }
    fe_mul(p, a, b);
    fe_sq(q, a);
    for (int r = 0; r < 3; r++) { fe_mul(p, p, b); fe_sq(q, q); }
    fd fa, fb, fp, fq;
    fd_from_u32(fa, wa); fd_from_u32(fb, wb);
    fd_mul(fp, fa, fb);
    fd_sq(fq, fa);
    for (int r = 0; r < 3; r++) { fd_mul(fp, fp, fb); fd_sq(fq, fq); }
    uint32_t w1[8], w2[8];
    fe_towords(w1, p); fd_to_u32(w2, fp);
    bool ok = true;
    for (int k = 0; k < 8; k++) ok = ok && w1[k] == w2[k];
    fe_towords(w1, q); fd_to_u32(w2, fq);
    for (int k = 0; k < 8; k++) ok = ok && w1[k] == w2[k];
    if (!ok) atomicAdd(bad, 1ull);
}

// ---- throughput: J of every 4 warps per SM sub-partition on the FP64 path, the rest on the integer path
// mode 0: multiplications, 1: squarings
__global__ void __launch_bounds__(512, 1) k_mix(int J, int mode, long long cycles, unsigned long long *counts /* [2] */, uint32_t *sink) {
    const int warp = threadIdx.x >> 5;
    const bool fp = ((warp >> 2) & 3) < J;          // warps w, w+4, w+8, w+12 share a sub-partition
    const long long t0 = clock64();
    unsigned long long done = 0;
    if (fp) {
        fd x, y;
        for (int k = 0; k < 12; k++) { x.v[k] = (double)(threadIdx.x + 3 + k) * pow2(fd_w(k)); y.v[k] = (double)(blockIdx.x + 5 + 2 * k) * pow2(fd_w(k)); }
        while (clock64() - t0 < cycles) {
#pragma unroll 1
            for (int r = 0; r < 8; r++) {
                if (mode == 0) { fd_mul(x, x, y); fd_mul(y, y, x); }
                else { fd_sq(x, x); fd_sq(y, y); }
            }
            done += 16;
        }
        double acc = 0;
        for (int k = 0; k < 12; k++) acc += x.v[k] + y.v[k];
        if (acc == 1.2345) sink[0] = 1;
    } else {
        fe x, y;
        for (int k = 0; k < 8; k++) { x.v[k] = threadIdx.x * 2654435761u + k; y.v[k] = blockIdx.x * 40503u + 7 * k + 1; }
        while (clock64() - t0 < cycles) {
#pragma unroll 1
            for (int r = 0; r < 8; r++) {
                if (mode == 0) { fe_mul(x, x, y); fe_mul(y, y, x); }
                else { fe_sq(x, x); fe_sq(y, y); }
            }
            done += 16;
        }
        uint32_t acc = 0;
        for (int k = 0; k < 8; k++) acc ^= x.v[k] ^ y.v[k];
        if (acc == 0x12345678u) sink[1] = 1;
    }
    if ((threadIdx.x & 31) == 0) atomicAdd(&counts[fp ? 1 : 0], done * 32ull);
}

int main(int argc, char **argv) {
    long long cycles = argc > 1 ? atoll(argv[1]) : 40000000ll;      // ~20 ms at 1.9 GHz
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    unsigned long long *d_counts, *d_bad;
    uint32_t *d_sink;
    cudaMalloc(&d_counts, 16); cudaMalloc(&d_bad, 8); cudaMalloc(&d_sink, 8);
    cudaMemset(d_bad, 0, 8);
    k_check<<<64, 128>>>(12345ull, d_bad, 64 * 128);
    unsigned long long bad = 0;
    cudaMemcpy(&bad, d_bad, 8, cudaMemcpyDeviceToHost);
    cudaError_t ce = cudaDeviceSynchronize();
    printf("{\"check\": {\"operands\": %d, \"mismatches\": %llu, \"cuda\": \"%s\"}, \"sms\": %d, \"cycles\": %lld, \"runs\": [", 64 * 128, bad,
           cudaGetErrorString(ce), sms, cycles);
    bool first = true;
    for (int mode = 0; mode < 2; mode++)
        for (int J = 0; J <= 4; J++) {
            cudaMemset(d_counts, 0, 16);
            k_mix<<<sms, 512>>>(J, mode, cycles / 20, d_counts, d_sink);      // warm-up
            cudaMemset(d_counts, 0, 16);
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            k_mix<<<sms, 512>>>(J, mode, cycles, d_counts, d_sink);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms = 0;
            cudaEventElapsedTime(&ms, e0, e1);
            unsigned long long c[2];
            cudaMemcpy(c, d_counts, 16, cudaMemcpyDeviceToHost);
            const double per_clk_sm_int = (double)c[0] / (double)cycles / sms, per_clk_sm_fp = (double)c[1] / (double)cycles / sms;
            printf("%s{\"op\": \"%s\", \"fp64_warps_of_4\": %d, \"int_per_clk_per_sm\": %.4f, \"fp64_per_clk_per_sm\": %.4f, \"total_per_clk_per_sm\": %.4f, \"ms\": %.2f}",
                   first ? "" : ", ", mode == 0 ? "mul" : "sq", J, per_clk_sm_int, per_clk_sm_fp, per_clk_sm_int + per_clk_sm_fp, ms);
            first = false;
        }
    printf("]}\n");
    return bad != 0;
}
