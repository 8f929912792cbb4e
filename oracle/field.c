/*
 * field.c -- GF(2^255-19) for the CPU oracle (test infrastructure, see eg_oracle.h).
 * Radix-2^51 limbs with unsigned __int128 products.  Restated from RFC 8032 5.1 / RFC 9496 4.1-4.2;
 * replaces curve25519-dalek's `FieldElement` (not vendored in /root/reference; used through
 * src/group/ristretto.rs:88-95 compress/decompress).
 */
#include "eg_oracle.h"
#include <string.h>

typedef unsigned __int128 u128;
#define MASK51 ((1ULL << 51) - 1)

static uint64_t load64(const uint8_t *p) {
    uint64_t r = 0;
    for (int i = 0; i < 8; i++) r |= (uint64_t)p[i] << (8 * i);
    return r;
}

void eo_fe_frombytes(eo_fe *h, const uint8_t s[32]) {
    h->v[0] = load64(s) & MASK51;
    h->v[1] = (load64(s + 6) >> 3) & MASK51;
    h->v[2] = (load64(s + 12) >> 6) & MASK51;
    h->v[3] = (load64(s + 19) >> 1) & MASK51;
    h->v[4] = (load64(s + 24) >> 12) & MASK51;   /* drops bit 255 */
}

static void fe_carry(eo_fe *h) {
    uint64_t c;
    c = h->v[0] >> 51; h->v[0] &= MASK51; h->v[1] += c;
    c = h->v[1] >> 51; h->v[1] &= MASK51; h->v[2] += c;
    c = h->v[2] >> 51; h->v[2] &= MASK51; h->v[3] += c;
    c = h->v[3] >> 51; h->v[3] &= MASK51; h->v[4] += c;
    c = h->v[4] >> 51; h->v[4] &= MASK51; h->v[0] += c * 19;
    c = h->v[0] >> 51; h->v[0] &= MASK51; h->v[1] += c;
}

void eo_fe_tobytes(uint8_t s[32], const eo_fe *f) {
    eo_fe t = *f;
    fe_carry(&t);
    fe_carry(&t);
    /* now t < 2^255 + small; compute t mod p by trial-adding 19 */
    uint64_t q = (t.v[0] + 19) >> 51;
    q = (t.v[1] + q) >> 51;
    q = (t.v[2] + q) >> 51;
    q = (t.v[3] + q) >> 51;
    q = (t.v[4] + q) >> 51;   /* q = 1 iff t >= p */
    t.v[0] += 19 * q;
    uint64_t c;
    c = t.v[0] >> 51; t.v[0] &= MASK51; t.v[1] += c;
    c = t.v[1] >> 51; t.v[1] &= MASK51; t.v[2] += c;
    c = t.v[2] >> 51; t.v[2] &= MASK51; t.v[3] += c;
    c = t.v[3] >> 51; t.v[3] &= MASK51; t.v[4] += c;
    t.v[4] &= MASK51;
    uint64_t w0 = t.v[0] | (t.v[1] << 51);
    uint64_t w1 = (t.v[1] >> 13) | (t.v[2] << 38);
    uint64_t w2 = (t.v[2] >> 26) | (t.v[3] << 25);
    uint64_t w3 = (t.v[3] >> 39) | (t.v[4] << 12);
    uint64_t w[4] = {w0, w1, w2, w3};
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 8; j++) s[8 * i + j] = (uint8_t)(w[i] >> (8 * j));
}

void eo_fe_add(eo_fe *h, const eo_fe *f, const eo_fe *g) {
    for (int i = 0; i < 5; i++) h->v[i] = f->v[i] + g->v[i];
    fe_carry(h);
}

void eo_fe_sub(eo_fe *h, const eo_fe *f, const eo_fe *g) {
    /* add 4p to keep limbs positive (inputs are carried: limbs < 2^52) */
    h->v[0] = f->v[0] + 0x1FFFFFFFFFFFB4ULL - g->v[0];
    h->v[1] = f->v[1] + 0x1FFFFFFFFFFFFCULL - g->v[1];
    h->v[2] = f->v[2] + 0x1FFFFFFFFFFFFCULL - g->v[2];
    h->v[3] = f->v[3] + 0x1FFFFFFFFFFFFCULL - g->v[3];
    h->v[4] = f->v[4] + 0x1FFFFFFFFFFFFCULL - g->v[4];
    fe_carry(h);
}

void eo_fe_neg(eo_fe *h, const eo_fe *f) {
    eo_fe z = {{0, 0, 0, 0, 0}};
    eo_fe_sub(h, &z, f);
}

void eo_fe_mul(eo_fe *h, const eo_fe *f, const eo_fe *g) {
    const uint64_t f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
    const uint64_t g0 = g->v[0], g1 = g->v[1], g2 = g->v[2], g3 = g->v[3], g4 = g->v[4];
    const uint64_t g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4;
    u128 r0 = (u128)f0 * g0 + (u128)f1 * g4_19 + (u128)f2 * g3_19 + (u128)f3 * g2_19 + (u128)f4 * g1_19;
    u128 r1 = (u128)f0 * g1 + (u128)f1 * g0 + (u128)f2 * g4_19 + (u128)f3 * g3_19 + (u128)f4 * g2_19;
    u128 r2 = (u128)f0 * g2 + (u128)f1 * g1 + (u128)f2 * g0 + (u128)f3 * g4_19 + (u128)f4 * g3_19;
    u128 r3 = (u128)f0 * g3 + (u128)f1 * g2 + (u128)f2 * g1 + (u128)f3 * g0 + (u128)f4 * g4_19;
    u128 r4 = (u128)f0 * g4 + (u128)f1 * g3 + (u128)f2 * g2 + (u128)f3 * g1 + (u128)f4 * g0;
    uint64_t c;
    r1 += (uint64_t)(r0 >> 51); h->v[0] = (uint64_t)r0 & MASK51;
    r2 += (uint64_t)(r1 >> 51); h->v[1] = (uint64_t)r1 & MASK51;
    r3 += (uint64_t)(r2 >> 51); h->v[2] = (uint64_t)r2 & MASK51;
    r4 += (uint64_t)(r3 >> 51); h->v[3] = (uint64_t)r3 & MASK51;
    c = (uint64_t)(r4 >> 51);   h->v[4] = (uint64_t)r4 & MASK51;
    h->v[0] += c * 19;
    c = h->v[0] >> 51; h->v[0] &= MASK51; h->v[1] += c;
}

void eo_fe_sq(eo_fe *h, const eo_fe *f) {
    const uint64_t f0 = f->v[0], f1 = f->v[1], f2 = f->v[2], f3 = f->v[3], f4 = f->v[4];
    const uint64_t f0_2 = 2 * f0, f1_2 = 2 * f1;
    const uint64_t f3_19 = 19 * f3, f4_19 = 19 * f4;
    u128 r0 = (u128)f0 * f0 + (u128)(2 * f1) * f4_19 + (u128)(2 * f2) * f3_19;
    u128 r1 = (u128)f0_2 * f1 + (u128)(2 * f2) * f4_19 + (u128)f3 * f3_19;
    u128 r2 = (u128)f0_2 * f2 + (u128)f1 * f1 + (u128)(2 * f3) * f4_19;
    u128 r3 = (u128)f0_2 * f3 + (u128)f1_2 * f2 + (u128)f4 * f4_19;
    u128 r4 = (u128)f0_2 * f4 + (u128)f1_2 * f3 + (u128)f2 * f2;
    uint64_t c;
    r1 += (uint64_t)(r0 >> 51); h->v[0] = (uint64_t)r0 & MASK51;
    r2 += (uint64_t)(r1 >> 51); h->v[1] = (uint64_t)r1 & MASK51;
    r3 += (uint64_t)(r2 >> 51); h->v[2] = (uint64_t)r2 & MASK51;
    r4 += (uint64_t)(r3 >> 51); h->v[3] = (uint64_t)r3 & MASK51;
    c = (uint64_t)(r4 >> 51);   h->v[4] = (uint64_t)r4 & MASK51;
    h->v[0] += c * 19;
    c = h->v[0] >> 51; h->v[0] &= MASK51; h->v[1] += c;
}

static void fe_sqn(eo_fe *h, const eo_fe *f, int n) {
    eo_fe_sq(h, f);
    for (int i = 1; i < n; i++) eo_fe_sq(h, h);
}

/* z^(2^250 - 1) and z^11, shared by invert and pow22523 */
static void fe_pow22501(eo_fe *t250, eo_fe *z11, const eo_fe *z) {
    eo_fe z2, z9, t, z2_5_0, z2_10_0, z2_20_0, z2_50_0, z2_100_0;
    eo_fe_sq(&z2, z);
    fe_sqn(&t, &z2, 2);
    eo_fe_mul(&z9, &t, z);
    eo_fe_mul(z11, &z9, &z2);
    eo_fe_sq(&t, z11);
    eo_fe_mul(&z2_5_0, &t, &z9);
    fe_sqn(&t, &z2_5_0, 5);
    eo_fe_mul(&z2_10_0, &t, &z2_5_0);
    fe_sqn(&t, &z2_10_0, 10);
    eo_fe_mul(&z2_20_0, &t, &z2_10_0);
    fe_sqn(&t, &z2_20_0, 20);
    eo_fe_mul(&t, &t, &z2_20_0);
    fe_sqn(&t, &t, 10);
    eo_fe_mul(&z2_50_0, &t, &z2_10_0);
    fe_sqn(&t, &z2_50_0, 50);
    eo_fe_mul(&z2_100_0, &t, &z2_50_0);
    fe_sqn(&t, &z2_100_0, 100);
    eo_fe_mul(&t, &t, &z2_100_0);
    fe_sqn(&t, &t, 50);
    eo_fe_mul(t250, &t, &z2_50_0);
}

void eo_fe_invert(eo_fe *h, const eo_fe *f) {
    eo_fe t250, z11, t;
    fe_pow22501(&t250, &z11, f);
    fe_sqn(&t, &t250, 5);
    eo_fe_mul(h, &t, &z11);       /* z^(2^255 - 21) */
}

static void fe_pow22523(eo_fe *h, const eo_fe *f) {
    eo_fe t250, z11, t;
    fe_pow22501(&t250, &z11, f);
    fe_sqn(&t, &t250, 2);
    eo_fe_mul(h, &t, f);          /* z^(2^252 - 3) = z^((p-5)/8) */
}

static int fe_iszero(const eo_fe *f) {
    uint8_t s[32];
    eo_fe_tobytes(s, f);
    uint8_t r = 0;
    for (int i = 0; i < 32; i++) r |= s[i];
    return r == 0;
}

static int fe_isnegative(const eo_fe *f) {
    uint8_t s[32];
    eo_fe_tobytes(s, f);
    return s[0] & 1;
}

static int fe_eq(const eo_fe *a, const eo_fe *b) {
    uint8_t x[32], y[32];
    eo_fe_tobytes(x, a);
    eo_fe_tobytes(y, b);
    return memcmp(x, y, 32) == 0;
}

/* exported for group.c */
int eo_fe_iszero(const eo_fe *f) { return fe_iszero(f); }
int eo_fe_isnegative(const eo_fe *f) { return fe_isnegative(f); }
int eo_fe_eq(const eo_fe *a, const eo_fe *b) { return fe_eq(a, b); }

static eo_fe SQRT_M1;
static int sqrt_m1_ready = 0;

const eo_fe *eo_fe_sqrt_m1(void) {
    if (!sqrt_m1_ready) {
        /* sqrt(-1) = 2^((p-1)/4) = 2^(2^253 - 5): (2^(2^252-3))^2 * 2 */
        eo_fe two = {{2, 0, 0, 0, 0}}, t;
        fe_pow22523(&t, &two);
        eo_fe_sq(&t, &t);
        eo_fe_mul(&SQRT_M1, &t, &two);
        sqrt_m1_ready = 1;
    }
    return &SQRT_M1;
}

/* RFC 9496 4.2 SQRT_RATIO_M1 */
int eo_fe_sqrt_ratio_i(eo_fe *out, const eo_fe *u, const eo_fe *v) {
    const eo_fe *i = eo_fe_sqrt_m1();
    eo_fe v3, v7, r, check, t, neg_u, neg_u_i;
    eo_fe_sq(&t, v);
    eo_fe_mul(&v3, &t, v);
    eo_fe_sq(&t, &v3);
    eo_fe_mul(&v7, &t, v);
    eo_fe_mul(&t, u, &v7);
    fe_pow22523(&t, &t);
    eo_fe_mul(&r, u, &v3);
    eo_fe_mul(&r, &r, &t);
    eo_fe_sq(&t, &r);
    eo_fe_mul(&check, v, &t);
    eo_fe_neg(&neg_u, u);
    eo_fe_mul(&neg_u_i, &neg_u, i);
    int correct_sign = fe_eq(&check, u);
    int flipped_sign = fe_eq(&check, &neg_u);
    int flipped_sign_i = fe_eq(&check, &neg_u_i);
    if (flipped_sign || flipped_sign_i) eo_fe_mul(&r, &r, i);
    if (fe_isnegative(&r)) eo_fe_neg(&r, &r);
    *out = r;
    return correct_sign || flipped_sign;
}

void eo_fe_mul_bytes(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) {
    eo_fe x, y;
    eo_fe_frombytes(&x, a);
    eo_fe_frombytes(&y, b);
    eo_fe_mul(&x, &x, &y);
    eo_fe_tobytes(out, &x);
}

void eo_fe_invert_bytes(uint8_t out[32], const uint8_t a[32]) {
    eo_fe x;
    eo_fe_frombytes(&x, a);
    eo_fe_invert(&x, &x);
    eo_fe_tobytes(out, &x);
}
