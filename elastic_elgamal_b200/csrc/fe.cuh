// fe.cuh -- GF(2^255-19) on 8 saturated 32-bit limbs, for sm_100a.
//
// Representation: value = sum v[i] * 2^(32 i), any representative in [0, 2^256) ("loose"); canonical
// form is produced only by fe_tobytes / comparisons.  2^256 = 38 (mod p) folds the high half of a
// product back with one more row of multiply-adds.  The INT32 multiply-add pipe is the roofline of this
// library: fe_mul / fe_sq are the two functions every cycle of the hot kernels goes through.
// The portable formulation (also compiled for the host by tests/hostsim, a test-only harness) and the
// device-tuned one live side by side and are checked against each other and against the oracle.
//
// This layer replaces curve25519-dalek's FieldElement, which the reference reaches through
// src/group/ristretto.rs:88-95 (compress / decompress) and every point operation.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define EG_HD __host__ __device__ __forceinline__
#define EG_HD_NOINLINE __host__ __device__ __noinline__
#else
#define EG_HD inline
#define EG_HD_NOINLINE
#endif

namespace eg {

struct fe { uint32_t v[8]; };

EG_HD fe fe_zero() { fe r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
EG_HD fe fe_one() { fe r = fe_zero(); r.v[0] = 1; return r; }
EG_HD fe fe_make(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t a4, uint32_t a5, uint32_t a6, uint32_t a7) {
    fe r; r.v[0] = a0; r.v[1] = a1; r.v[2] = a2; r.v[3] = a3; r.v[4] = a4; r.v[5] = a5; r.v[6] = a6; r.v[7] = a7; return r;
}

// curve constants (tools/gen_constants.py prints them from their defining equations)
EG_HD fe fe_const_d()  { return fe_make(0x135978a3u, 0x75eb4dcau, 0x4141d8abu, 0x00700a4du, 0x7779e898u, 0x8cc74079u, 0x2b6ffe73u, 0x52036ceeu); }
EG_HD fe fe_const_2d() { return fe_make(0x26b2f159u, 0xebd69b94u, 0x8283b156u, 0x00e0149au, 0xeef3d130u, 0x198e80f2u, 0x56dffce7u, 0x2406d9dcu); }
EG_HD fe fe_const_sqrtm1() { return fe_make(0x4a0ea0b0u, 0xc4ee1b27u, 0xad2fe478u, 0x2f431806u, 0x3dfbd7a7u, 0x2b4d0099u, 0x4fc1df0bu, 0x2b832480u); }
EG_HD fe fe_const_invsqrt_a_minus_d() { return fe_make(0x805d40eau, 0x99c8fdaau, 0x5a4172beu, 0x9d2f1617u, 0xfe01d840u, 0x16c27b91u, 0xcfaffca2u, 0x786c8905u); }

// ------------------------------------------------------------------ add / sub

EG_HD void fe_add(fe &r, const fe &a, const fe &b) {
#if defined(__CUDA_ARCH__)
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, c;
    asm("add.cc.u32 %0, %9, %17;\n\t"
        "addc.cc.u32 %1, %10, %18;\n\t"
        "addc.cc.u32 %2, %11, %19;\n\t"
        "addc.cc.u32 %3, %12, %20;\n\t"
        "addc.cc.u32 %4, %13, %21;\n\t"
        "addc.cc.u32 %5, %14, %22;\n\t"
        "addc.cc.u32 %6, %15, %23;\n\t"
        "addc.cc.u32 %7, %16, %24;\n\t"
        "addc.u32 %8, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // fold the carry: r += 38*c; if that overflows the wrapped value is tiny, so a final +38 stays in limb 0
    uint32_t k = c * 38u;
    asm("add.cc.u32 %0, %0, %9;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.cc.u32 %2, %2, 0;\n\t"
        "addc.cc.u32 %3, %3, 0;\n\t"
        "addc.cc.u32 %4, %4, 0;\n\t"
        "addc.cc.u32 %5, %5, 0;\n\t"
        "addc.cc.u32 %6, %6, 0;\n\t"
        "addc.cc.u32 %7, %7, 0;\n\t"
        "addc.u32 %8, 0, 0;"
        : "+r"(t0), "+r"(t1), "+r"(t2), "+r"(t3), "+r"(t4), "+r"(t5), "+r"(t6), "+r"(t7), "=r"(c)
        : "r"(k));
    t0 += c * 38u;
    r.v[0] = t0; r.v[1] = t1; r.v[2] = t2; r.v[3] = t3; r.v[4] = t4; r.v[5] = t5; r.v[6] = t6; r.v[7] = t7;
#else
    uint64_t c = 0;
    uint32_t t[8];
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    c *= 38;
    for (int i = 0; i < 8; i++) { c += t[i]; t[i] = (uint32_t)c; c >>= 32; }
    t[0] += (uint32_t)c * 38u;
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
#endif
}

EG_HD void fe_sub(fe &r, const fe &a, const fe &b) {
#if defined(__CUDA_ARCH__)
    uint32_t t0, t1, t2, t3, t4, t5, t6, t7, c;
    asm("sub.cc.u32 %0, %9, %17;\n\t"
        "subc.cc.u32 %1, %10, %18;\n\t"
        "subc.cc.u32 %2, %11, %19;\n\t"
        "subc.cc.u32 %3, %12, %20;\n\t"
        "subc.cc.u32 %4, %13, %21;\n\t"
        "subc.cc.u32 %5, %14, %22;\n\t"
        "subc.cc.u32 %6, %15, %23;\n\t"
        "subc.cc.u32 %7, %16, %24;\n\t"
        "subc.u32 %8, 0, 0;"
        : "=r"(t0), "=r"(t1), "=r"(t2), "=r"(t3), "=r"(t4), "=r"(t5), "=r"(t6), "=r"(t7), "=r"(c)
        : "r"(a.v[0]), "r"(a.v[1]), "r"(a.v[2]), "r"(a.v[3]), "r"(a.v[4]), "r"(a.v[5]), "r"(a.v[6]), "r"(a.v[7]),
          "r"(b.v[0]), "r"(b.v[1]), "r"(b.v[2]), "r"(b.v[3]), "r"(b.v[4]), "r"(b.v[5]), "r"(b.v[6]), "r"(b.v[7]));
    // c = 0xffffffff on borrow: a-b+2^256 = a-b+38 (mod p), so take 38 back; at most once more
    uint32_t k = c & 38u;
    asm("sub.cc.u32 %0, %0, %9;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.cc.u32 %2, %2, 0;\n\t"
        "subc.cc.u32 %3, %3, 0;\n\t"
        "subc.cc.u32 %4, %4, 0;\n\t"
        "subc.cc.u32 %5, %5, 0;\n\t"
        "subc.cc.u32 %6, %6, 0;\n\t"
        "subc.cc.u32 %7, %7, 0;\n\t"
        "subc.u32 %8, 0, 0;"
        : "+r"(t0), "+r"(t1), "+r"(t2), "+r"(t3), "+r"(t4), "+r"(t5), "+r"(t6), "+r"(t7), "=r"(c)
        : "r"(k));
    t0 -= c & 38u;
    r.v[0] = t0; r.v[1] = t1; r.v[2] = t2; r.v[3] = t3; r.v[4] = t4; r.v[5] = t5; r.v[6] = t6; r.v[7] = t7;
#else
    int64_t c = 0;
    uint32_t t[8];
    for (int i = 0; i < 8; i++) { c += (int64_t)a.v[i] - (int64_t)b.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    c = c ? -38 : 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)t[i]; t[i] = (uint32_t)c; c >>= 32; }
    if (c) t[0] -= 38u;
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
#endif
}

EG_HD void fe_neg(fe &r, const fe &a) { fe z = fe_zero(); fe_sub(r, z, a); }

// ------------------------------------------------------------------ mul / sq

// 512 -> 256 bits: lo + 38*hi, then the small carry limb, then (rarely) one last +38
EG_HD void fe_fold512(fe &r, const uint32_t t[16]) {
    uint64_t c = 0;
    uint32_t s[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c += (uint64_t)t[i] + (uint64_t)t[i + 8] * 38u; s[i] = (uint32_t)c; c >>= 32; }
    c *= 38;
#pragma unroll
    for (int i = 0; i < 8; i++) { c += s[i]; s[i] = (uint32_t)c; c >>= 32; }
    s[0] += (uint32_t)c * 38u;
#pragma unroll
    for (int i = 0; i < 8; i++) r.v[i] = s[i];
}

EG_HD void fe_mul_portable(fe &r, const fe &a, const fe &b) {
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = 0; j < 8; j++) {
            c += (uint64_t)a.v[i] * b.v[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
    fe_fold512(r, t);
}

EG_HD void fe_sq_portable(fe &r, const fe &a) {
    // off-diagonal products once, doubled, plus the diagonal
    uint32_t t[16];
#pragma unroll
    for (int i = 0; i < 16; i++) t[i] = 0;
#pragma unroll
    for (int i = 0; i < 7; i++) {
        uint64_t c = 0;
#pragma unroll
        for (int j = i + 1; j < 8; j++) {
            c += (uint64_t)a.v[i] * a.v[j] + t[i + j];
            t[i + j] = (uint32_t)c;
            c >>= 32;
        }
        t[i + 8] = (uint32_t)c;
    }
    uint32_t top = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) { uint32_t n = t[i] >> 31; t[i] = (t[i] << 1) | top; top = n; }
    uint64_t c = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) {
        uint64_t d = (uint64_t)a.v[i] * a.v[i];
        c += (uint64_t)t[2 * i] + (uint32_t)d;
        t[2 * i] = (uint32_t)c; c >>= 32;
        c += (uint64_t)t[2 * i + 1] + (d >> 32);
        t[2 * i + 1] = (uint32_t)c; c >>= 32;
    }
    fe_fold512(r, t);
}

}  // namespace eg

#if defined(__CUDA_ARCH__)
#include "fe_ptx.cuh"
#endif

namespace eg {

// On the device fe_mul / fe_sq are real (non-inlined) functions taking and returning their operands BY VALUE: the
// CUDA ABI keeps such small aggregates in registers (no local-memory traffic), and every point formula becomes a
// short sequence of calls.  This keeps the hot loops of the multi-scalar kernels inside the instruction cache:
// with everything inlined k_commit was 229 KB of SASS and ncu showed `no_instruction` as its top stall reason
// (profiles/r1_k_commit_inlined.txt).
#if defined(__CUDA_ARCH__) && !defined(EG_INLINE_FE)
#if defined(EG_PORTABLE_FE)
static __device__ __noinline__ fe fe_mul_call(const fe a, const fe b) { fe r; fe_mul_portable(r, a, b); return r; }
static __device__ __noinline__ fe fe_sq_call(const fe a) { fe r; fe_sq_portable(r, a); return r; }
#else
#if defined(EG_FE_SHIFT_FOLD)
// A/B variant: the 2^256 = 38 fold built from funnel shifts and add chains (ALU pipe) instead of 8 wide multiply-adds
#define EG_FE_MUL_PTX fe_mul_ptx_sf
#define EG_FE_SQ_PTX fe_sq_ptx_sf
#elif defined(EG_FE_KARATSUBA)
// A/B variant: one Karatsuba level (three 4 x 4 limb products): 56 instead of 72 wide multiply-adds, ~75 more ALU instructions
#define EG_FE_MUL_PTX fe_mul_ptx_k
#define EG_FE_SQ_PTX fe_sq_ptx
#else
#define EG_FE_MUL_PTX fe_mul_ptx
#define EG_FE_SQ_PTX fe_sq_ptx
#endif
static __device__ __noinline__ fe fe_mul_call(const fe a, const fe b) { fe r; EG_FE_MUL_PTX(r, a, b); return r; }
static __device__ __noinline__ fe fe_sq_call(const fe a) { fe r; EG_FE_SQ_PTX(r, a); return r; }
#endif
#endif

EG_HD void fe_mul(fe &r, const fe &a, const fe &b) {
#if defined(__CUDA_ARCH__) && !defined(EG_INLINE_FE)
    r = fe_mul_call(a, b);
#elif defined(__CUDA_ARCH__) && !defined(EG_PORTABLE_FE)
    fe_mul_ptx(r, a, b);
#else
    fe_mul_portable(r, a, b);
#endif
}

EG_HD void fe_sq(fe &r, const fe &a) {
#if defined(__CUDA_ARCH__) && !defined(EG_INLINE_FE)
    r = fe_sq_call(a);
#elif defined(__CUDA_ARCH__) && !defined(EG_PORTABLE_FE)
    fe_sq_ptx(r, a);
#else
    fe_sq_portable(r, a);
#endif
}

// n >= 1 squarings; kept as a rolled loop so the exponentiation chains stay small in the I-cache
EG_HD void fe_sqn(fe &r, const fe &a, int n) {
    fe t = a;
#pragma unroll 1
    for (int i = 0; i < n; i++) fe_sq(t, t);
    r = t;
}

// ------------------------------------------------------------------ canonical form, predicates

// fully reduce into [0, p)
EG_HD void fe_canon(fe &r, const fe &a) {
    // t = (a mod 2^255) + 19*(a >> 255) < 2^255 + 19
    uint32_t top = a.v[7] >> 31;
    uint32_t t[8];
    uint64_t c = (uint64_t)top * 19u;
    for (int i = 0; i < 8; i++) { c += (i == 7) ? (a.v[7] & 0x7fffffffu) : a.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    // u = t + 19; if u >= 2^255 then t >= p and the answer is u - 2^255
    uint32_t u[8];
    c = 19;
    for (int i = 0; i < 8; i++) { c += t[i]; u[i] = (uint32_t)c; c >>= 32; }
    uint32_t ge = u[7] >> 31;
    u[7] &= 0x7fffffffu;
    for (int i = 0; i < 8; i++) r.v[i] = ge ? u[i] : t[i];
}

EG_HD void fe_tobytes(uint8_t s[32], const fe &a) {
    fe t; fe_canon(t, a);
    for (int i = 0; i < 8; i++) {
        s[4 * i] = (uint8_t)t.v[i]; s[4 * i + 1] = (uint8_t)(t.v[i] >> 8);
        s[4 * i + 2] = (uint8_t)(t.v[i] >> 16); s[4 * i + 3] = (uint8_t)(t.v[i] >> 24);
    }
}

// canonical little-endian words (for coalesced 32-bit stores)
EG_HD void fe_towords(uint32_t w[8], const fe &a) {
    fe t; fe_canon(t, a);
    for (int i = 0; i < 8; i++) w[i] = t.v[i];
}

// loads 32 bytes, returns false when the encoding is not canonical (value >= p or bit 255 set)
EG_HD bool fe_frombytes_canonical(fe &r, const uint8_t s[32]) {
    for (int i = 0; i < 8; i++)
        r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    // canonical iff r < p = 2^255 - 19
    if (r.v[7] >> 31) return false;
    bool all_ones = r.v[7] == 0x7fffffffu;
    for (int i = 1; i < 7; i++) all_ones = all_ones && (r.v[i] == 0xffffffffu);
    if (all_ones && r.v[0] >= 0xffffffedu) return false;
    return true;
}

EG_HD bool fe_fromwords_canonical(fe &r, const uint32_t w[8]) {
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    if (r.v[7] >> 31) return false;
    bool all_ones = r.v[7] == 0x7fffffffu;
    for (int i = 1; i < 7; i++) all_ones = all_ones && (r.v[i] == 0xffffffffu);
    if (all_ones && r.v[0] >= 0xffffffedu) return false;
    return true;
}

EG_HD bool fe_iszero(const fe &a) {
    fe t; fe_canon(t, a);
    uint32_t o = 0;
    for (int i = 0; i < 8; i++) o |= t.v[i];
    return o == 0;
}

EG_HD bool fe_isneg(const fe &a) { fe t; fe_canon(t, a); return (t.v[0] & 1u) != 0; }

EG_HD bool fe_eq(const fe &a, const fe &b) { fe d; fe_sub(d, a, b); return fe_iszero(d); }

EG_HD void fe_select(fe &r, const fe &a, const fe &b, bool take_b) {
    for (int i = 0; i < 8; i++) r.v[i] = take_b ? b.v[i] : a.v[i];
}

EG_HD void fe_cneg(fe &r, const fe &a, bool negate) {
    fe n; fe_neg(n, a);
    fe_select(r, a, n, negate);
}

EG_HD void fe_abs(fe &r, const fe &a) { fe_cneg(r, a, fe_isneg(a)); }

// ------------------------------------------------------------------ exponentiations

// returns z^(2^250-1) and z^11
EG_HD void fe_pow22501(fe &t250, fe &z11, const fe &z) {
    fe z2, z9, t, z2_5_0, z2_10_0, z2_20_0, z2_50_0, z2_100_0;
    fe_sq(z2, z);
    fe_sqn(t, z2, 2);
    fe_mul(z9, t, z);
    fe_mul(z11, z9, z2);
    fe_sq(t, z11);
    fe_mul(z2_5_0, t, z9);
    fe_sqn(t, z2_5_0, 5);
    fe_mul(z2_10_0, t, z2_5_0);
    fe_sqn(t, z2_10_0, 10);
    fe_mul(z2_20_0, t, z2_10_0);
    fe_sqn(t, z2_20_0, 20);
    fe_mul(t, t, z2_20_0);
    fe_sqn(t, t, 10);
    fe_mul(z2_50_0, t, z2_10_0);
    fe_sqn(t, z2_50_0, 50);
    fe_mul(z2_100_0, t, z2_50_0);
    fe_sqn(t, z2_100_0, 100);
    fe_mul(t, t, z2_100_0);
    fe_sqn(t, t, 50);
    fe_mul(t250, t, z2_50_0);
}

EG_HD void fe_invert(fe &r, const fe &z) {
    fe t250, z11, t;
    fe_pow22501(t250, z11, z);
    fe_sqn(t, t250, 5);
    fe_mul(r, t, z11);
}

EG_HD void fe_pow22523(fe &r, const fe &z) {
    fe t250, z11, t;
    fe_pow22501(t250, z11, z);
    fe_sqn(t, t250, 2);
    fe_mul(r, t, z);
}

// RFC 9496 4.2 SQRT_RATIO_M1: r = sqrt(u/v) (non-negative) when u/v is square, else sqrt(i*u/v)
EG_HD bool fe_sqrt_ratio_i(fe &r, const fe &u, const fe &v) {
    fe v3, v7, t, x, check, neg_u, neg_u_i;
    const fe sqrtm1 = fe_const_sqrtm1();
    fe_sq(t, v); fe_mul(v3, t, v);
    fe_sq(t, v3); fe_mul(v7, t, v);
    fe_mul(t, u, v7);
    fe_pow22523(t, t);
    fe_mul(x, u, v3);
    fe_mul(x, x, t);
    fe_sq(t, x);
    fe_mul(check, v, t);
    fe_neg(neg_u, u);
    fe_mul(neg_u_i, neg_u, sqrtm1);
    bool correct_sign = fe_eq(check, u);
    bool flipped_sign = fe_eq(check, neg_u);
    bool flipped_sign_i = fe_eq(check, neg_u_i);
    fe xi; fe_mul(xi, x, sqrtm1);
    fe_select(x, x, xi, flipped_sign || flipped_sign_i);
    fe_abs(r, x);
    return correct_sign || flipped_sign;
}

// 1/sqrt(v) specialisation (u = 1) used by both ristretto maps: saves two multiplications
EG_HD bool fe_invsqrt(fe &r, const fe &v) {
    return fe_sqrt_ratio_i(r, fe_one(), v);
}

}  // namespace eg
