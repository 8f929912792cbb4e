// sc.cuh -- scalars mod l = 2^252 + 27742317777372353535851937790883648493 on 8 x 32-bit limbs.
//
// Replaces curve25519-dalek's Scalar as used by src/group/ristretto.rs:23-70: canonical parsing
// (deserialize_scalar :59-62), 64-byte wide reduction (generate_scalar :28-32, scalar_from_random_bytes
// :34-38), negation / add / mul (ScalarOps bounds, group/mod.rs:68-82).  Montgomery form (R = 2^256) is
// internal to sc_mul / sc_from_wide.  Scalars are <1 % of the work; this code favours clarity.
#pragma once
#include <stdint.h>
#include "fe.cuh"

namespace eg {

struct sc { uint32_t v[8]; };   // canonical: value < l

EG_HD uint32_t sc_L(int i) {
    const uint32_t L[8] = {0x5cf5d3edu, 0x5812631au, 0xa2f79cd6u, 0x14def9deu, 0, 0, 0, 0x10000000u};
    return L[i];
}
EG_HD uint32_t sc_R1(int i) {   // 2^256 mod l
    const uint32_t R1[8] = {0x8d98951du, 0xd6ec3174u, 0x737dcf70u, 0xc6ef5bf4u, 0xfffffffeu, 0xffffffffu, 0xffffffffu, 0x0fffffffu};
    return R1[i];
}
EG_HD uint32_t sc_RR(int i) {   // 2^512 mod l
    const uint32_t RR[8] = {0x449c0f01u, 0xa40611e3u, 0x68859347u, 0xd00e1ba7u, 0x17f5be65u, 0xceec73d2u, 0x7c309a3du, 0x0399411bu};
    return RR[i];
}
#define EG_SC_LFACTOR 0x12547e1bu   // -l^{-1} mod 2^32

EG_HD sc sc_zero() { sc r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
EG_HD sc sc_from_u64(uint64_t x) { sc r = sc_zero(); r.v[0] = (uint32_t)x; r.v[1] = (uint32_t)(x >> 32); return r; }

EG_HD bool sc_geq_l(const uint32_t a[8]) {
    for (int i = 7; i >= 0; i--) {
        uint32_t l = sc_L(i);
        if (a[i] > l) return true;
        if (a[i] < l) return false;
    }
    return true;
}

EG_HD void sc_sub_l(uint32_t a[8]) {
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)a[i] - (int64_t)sc_L(i); a[i] = (uint32_t)c; c >>= 32; }
}

EG_HD bool sc_is_canonical_words(const uint32_t w[8]) { return !sc_geq_l(w); }

EG_HD bool sc_from_words(sc &r, const uint32_t w[8]) {
    for (int i = 0; i < 8; i++) r.v[i] = w[i];
    return !sc_geq_l(r.v);
}

EG_HD bool sc_frombytes(sc &r, const uint8_t s[32]) {
    for (int i = 0; i < 8; i++)
        r.v[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    return !sc_geq_l(r.v);
}

EG_HD void sc_tobytes(uint8_t s[32], const sc &a) {
    for (int i = 0; i < 8; i++) {
        s[4 * i] = (uint8_t)a.v[i]; s[4 * i + 1] = (uint8_t)(a.v[i] >> 8);
        s[4 * i + 2] = (uint8_t)(a.v[i] >> 16); s[4 * i + 3] = (uint8_t)(a.v[i] >> 24);
    }
}

EG_HD bool sc_iszero(const sc &a) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i]; return o == 0; }
EG_HD bool sc_eq(const sc &a, const sc &b) { uint32_t o = 0; for (int i = 0; i < 8; i++) o |= a.v[i] ^ b.v[i]; return o == 0; }

// t (< 2 l) -> t mod l without a data-dependent branch: subtract l, keep the difference unless it borrowed.  The operands
// of sc_add / sc_neg / sc_sub / sc_mul are secrets on the proving side (randomness, nonces), so these stay branch-free.
EG_HD void sc_csub_l(uint32_t t[8]) {
    uint32_t d[8];
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)t[i] - (int64_t)sc_L(i); d[i] = (uint32_t)c; c >>= 32; }
    const uint32_t keep = (uint32_t)c;                      // all ones when t < l (the subtraction borrowed)
    for (int i = 0; i < 8; i++) t[i] = (t[i] & keep) | (d[i] & ~keep);
}

EG_HD void sc_add(sc &r, const sc &a, const sc &b) {
    uint64_t c = 0;
    uint32_t t[8];
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + b.v[i]; t[i] = (uint32_t)c; c >>= 32; }
    sc_csub_l(t);                       // a, b < l < 2^253: no carry out
    for (int i = 0; i < 8; i++) r.v[i] = t[i];
}

EG_HD void sc_neg(sc &r, const sc &a) {
    uint32_t nz = 0;
    for (int i = 0; i < 8; i++) nz |= a.v[i];
    const uint32_t m = (uint32_t)0 - (uint32_t)((nz | ((uint32_t)0 - nz)) >> 31);     // all ones unless a == 0
    int64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (int64_t)sc_L(i) - (int64_t)a.v[i]; r.v[i] = (uint32_t)c & m; c >>= 32; }
}

EG_HD void sc_sub(sc &r, const sc &a, const sc &b) { sc n; sc_neg(n, b); sc_add(r, a, n); }

// Montgomery product a*b/2^256 mod l (CIOS); requires a*b < l*2^256; result < l
EG_HD void sc_montmul(uint32_t r[8], const uint32_t a[8], const uint32_t b[8]) {
    uint32_t t[10];
    for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll 1
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)a[j] * b[i] + t[j]; t[j] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[8] = (uint32_t)c; t[9] = (uint32_t)(c >> 32);
        uint32_t m = t[0] * EG_SC_LFACTOR;
        c = ((uint64_t)m * sc_L(0) + t[0]) >> 32;
        for (int j = 1; j < 8; j++) { c += (uint64_t)m * sc_L(j) + t[j]; t[j - 1] = (uint32_t)c; c >>= 32; }
        c += t[8]; t[7] = (uint32_t)c; t[8] = t[9] + (uint32_t)(c >> 32); t[9] = 0;
    }
    {   // result < 2 l: one conditional subtraction over 9 limbs, branch-free (keep the difference unless it borrowed)
        uint32_t d[8];
        int64_t c = 0;
        for (int i = 0; i < 8; i++) { c += (int64_t)t[i] - (int64_t)sc_L(i); d[i] = (uint32_t)c; c >>= 32; }
        c += (int64_t)t[8];
        const uint32_t keep = (uint32_t)(c >> 32);          // all ones when t < l
        for (int i = 0; i < 8; i++) r[i] = (t[i] & keep) | (d[i] & ~keep);
    }
}

EG_HD void sc_mul(sc &r, const sc &a, const sc &b) {
    uint32_t t[8], rr[8];
    for (int i = 0; i < 8; i++) rr[i] = sc_RR(i);
    sc_montmul(t, a.v, b.v);
    sc_montmul(r.v, t, rr);
}

// r = a*b + c
EG_HD void sc_muladd(sc &r, const sc &a, const sc &b, const sc &c) { sc t; sc_mul(t, a, b); sc_add(r, t, c); }

// 64 little-endian bytes (as 16 words) -> mod l : lo + hi * 2^256
EG_HD void sc_from_wide_words(sc &r, const uint32_t w[16]) {
    uint32_t r1[8], rr[8], lo[8], hi[8];
    for (int i = 0; i < 8; i++) { r1[i] = sc_R1(i); rr[i] = sc_RR(i); }
    sc x, y;
    sc_montmul(lo, w, r1);          // lo * R / R
    sc_montmul(hi, w + 8, rr);      // hi * R^2 / R
    for (int i = 0; i < 8; i++) { x.v[i] = lo[i]; y.v[i] = hi[i]; }
    sc_add(r, x, y);
}

EG_HD void sc_from_wide_bytes(sc &r, const uint8_t s[64]) {
    uint32_t w[16];
    for (int i = 0; i < 16; i++)
        w[i] = (uint32_t)s[4 * i] | ((uint32_t)s[4 * i + 1] << 8) | ((uint32_t)s[4 * i + 2] << 16) | ((uint32_t)s[4 * i + 3] << 24);
    sc_from_wide_words(r, w);
}

// a^(l-2); variable time (inputs are public Lagrange denominators, sharing/mod.rs:139-170)
EG_HD void sc_invert(sc &r, const sc &a) {
    sc acc = sc_from_u64(1), base = a;
#pragma unroll 1
    for (int i = 0; i < 253; i++) {
        uint32_t e = sc_L(i >> 5) - ((i >> 5) == 0 ? 2u : 0u);   // l - 2: low word 0x5cf5d3ed - 2, no borrow
        if ((e >> (i & 31)) & 1u) sc_mul(acc, acc, base);
        sc_mul(base, base, base);
    }
    r = acc;
}

}  // namespace eg
