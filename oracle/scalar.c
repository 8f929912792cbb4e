/*
 * scalar.c -- integers mod l = 2^252 + 27742317777372353535851937790883648493 for the CPU oracle
 * (test infrastructure, see eg_oracle.h).  Replaces curve25519-dalek's `Scalar` as used by
 * src/group/ristretto.rs:23-70 (generate_scalar, scalar_from_random_bytes, invert_scalars,
 * serialize_scalar, deserialize_scalar).  Montgomery form (R = 2^256) is internal to mul.
 */
#include "eg_oracle.h"
#include <string.h>

typedef unsigned __int128 u128;

static const uint64_t L[4] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0x0ULL, 0x1000000000000000ULL};
static const uint64_t R1[4] = {0xd6ec31748d98951dULL, 0xc6ef5bf4737dcf70ULL, 0xfffffffffffffffeULL, 0x0fffffffffffffffULL};
static const uint64_t RR[4] = {0xa40611e3449c0f01ULL, 0xd00e1ba768859347ULL, 0xceec73d217f5be65ULL, 0x0399411b7c309a3dULL};
static const uint64_t LFACTOR = 0xd2b51da312547e1bULL;   /* -l^{-1} mod 2^64 */

static int geq_l(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > L[i]) return 1;
        if (a[i] < L[i]) return 0;
    }
    return 1;
}

static void sub_l(uint64_t a[4]) {
    uint64_t borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)a[i] - L[i] - borrow;
        a[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
}

/* Montgomery product a*b/R mod l, inputs a*b < l*R, output < l */
static void mont_mul(uint64_t r[4], const uint64_t a[4], const uint64_t b[4]) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        u128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (u128)t[j] + (u128)a[j] * b[i];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * LFACTOR;
        c = (u128)t[0] + (u128)m * L[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (u128)t[j] + (u128)m * L[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
        t[5] = 0;
    }
    /* t < 2l */
    uint64_t out[4] = {t[0], t[1], t[2], t[3]};
    if (t[4] || geq_l(out)) sub_l(out);
    memcpy(r, out, 32);
}

static void load_le(uint64_t out[4], const uint8_t b[32]) {
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0;
        for (int j = 0; j < 8; j++) w |= (uint64_t)b[8 * i + j] << (8 * j);
        out[i] = w;
    }
}

int eo_sc_from_canonical(eo_sc *s, const uint8_t b[32]) {
    load_le(s->v, b);
    return !geq_l(s->v);
}

void eo_sc_from_u64(eo_sc *s, uint64_t x) {
    s->v[0] = x; s->v[1] = s->v[2] = s->v[3] = 0;
}

void eo_sc_tobytes(uint8_t b[32], const eo_sc *s) {
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) b[8 * i + j] = (uint8_t)(s->v[i] >> (8 * j));
}

void eo_sc_add(eo_sc *r, const eo_sc *a, const eo_sc *b) {
    uint64_t t[4], carry = 0;
    for (int i = 0; i < 4; i++) {
        u128 c = (u128)a->v[i] + b->v[i] + carry;
        t[i] = (uint64_t)c;
        carry = (uint64_t)(c >> 64);
    }
    /* a,b < l < 2^253 so no carry out */
    if (geq_l(t)) sub_l(t);
    memcpy(r->v, t, 32);
}

void eo_sc_neg(eo_sc *r, const eo_sc *a) {
    if ((a->v[0] | a->v[1] | a->v[2] | a->v[3]) == 0) { memset(r->v, 0, 32); return; }
    uint64_t t[4], borrow = 0;
    for (int i = 0; i < 4; i++) {
        u128 d = (u128)L[i] - a->v[i] - borrow;
        t[i] = (uint64_t)d;
        borrow = (uint64_t)(d >> 64) & 1;
    }
    memcpy(r->v, t, 32);
}

void eo_sc_sub(eo_sc *r, const eo_sc *a, const eo_sc *b) {
    eo_sc nb;
    eo_sc_neg(&nb, b);
    eo_sc_add(r, a, &nb);
}

void eo_sc_mul(eo_sc *r, const eo_sc *a, const eo_sc *b) {
    uint64_t t[4];
    mont_mul(t, a->v, b->v);     /* ab/R */
    mont_mul(r->v, t, RR);       /* ab */
}

/* 512-bit little-endian -> mod l:  lo + hi*R = lo*R/R + hi*R^2/R */
void eo_sc_from_wide(eo_sc *s, const uint8_t b[64]) {
    uint64_t lo[4], hi[4], a[4], c[4];
    load_le(lo, b);
    load_le(hi, b + 32);
    mont_mul(a, lo, R1);
    mont_mul(c, hi, RR);
    eo_sc x, y;
    memcpy(x.v, a, 32);
    memcpy(y.v, c, 32);
    eo_sc_add(s, &x, &y);
}

int eo_sc_eq(const eo_sc *a, const eo_sc *b) {
    return memcmp(a->v, b->v, 32) == 0;
}

/* a^(l-2) by square-and-multiply; variable time is fine (ristretto.rs:41-47 is not secret here) */
void eo_sc_invert(eo_sc *r, const eo_sc *a) {
    uint64_t e[4] = {L[0] - 2, L[1], L[2], L[3]};
    eo_sc acc, base = *a;
    eo_sc_from_u64(&acc, 1);
    for (int i = 0; i < 253; i++) {
        if ((e[i >> 6] >> (i & 63)) & 1) eo_sc_mul(&acc, &acc, &base);
        eo_sc_mul(&base, &base, &base);
    }
    *r = acc;
}

void eo_scalar_reduce_wide_bytes(uint8_t out[32], const uint8_t in[64]) {
    eo_sc s;
    eo_sc_from_wide(&s, in);
    eo_sc_tobytes(out, &s);
}

int eo_scalar_is_canonical(const uint8_t in[32]) {
    eo_sc s;
    return eo_sc_from_canonical(&s, in);
}

void eo_scalar_muladd_bytes(uint8_t out[32], const uint8_t a[32], const uint8_t b[32], const uint8_t c[32]) {
    eo_sc x, y, z;
    eo_sc_from_canonical(&x, a);
    eo_sc_from_canonical(&y, b);
    eo_sc_from_canonical(&z, c);
    eo_sc_mul(&x, &x, &y);
    eo_sc_add(&x, &x, &z);
    eo_sc_tobytes(out, &x);
}

void eo_scalar_invert_bytes(uint8_t out[32], const uint8_t a[32]) {
    eo_sc x;
    eo_sc_from_canonical(&x, a);
    eo_sc_invert(&x, &x);
    eo_sc_tobytes(out, &x);
}
