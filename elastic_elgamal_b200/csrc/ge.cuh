// ge.cuh -- ristretto255 group elements: extended twisted-Edwards arithmetic (a = -1), the RFC 9496
// encoding, and the multi-scalar chain used by every verification equation.
//
// Replaces curve25519-dalek's RistrettoPoint / CompressedRistretto / RISTRETTO_BASEPOINT_TABLE /
// VartimeMultiscalarMul as reached through src/group/ristretto.rs:72-146:
//   serialize_element :88 -> ge_encode, deserialize_element :93 -> ge_decode,
//   vartime_double_mul_generator :131 and vartime_multi_mul :139 -> ge_msm_chain (Straus with shared
//   doublings; fixed 4-bit signed windows for per-item bases, 8-bit signed windows over a 128-entry
//   affine table for the fixed bases G and K).  Results are group elements, so the choice of algorithm
//   is invisible in the canonical encodings the verifier hashes.
#pragma once
#include "fe.cuh"
#include "sc.cuh"

namespace eg {

struct ge_ext { fe X, Y, Z, T; };             // x = X/Z, y = Y/Z, xy = T/Z
struct ge_cached { fe YpX, YmX, Z, T2d; };    // (Y+X, Y-X, Z, 2dT)
struct ge_niels { fe ypx, ymx, xy2d; };       // affine: (y+x, y-x, 2dxy)
struct ge_p1p1 { fe E, F, G, H; };            // completed: X = EF, Y = GH, Z = FG, T = EH

EG_HD ge_ext ge_identity() { ge_ext r; r.X = fe_zero(); r.Y = fe_one(); r.Z = fe_one(); r.T = fe_zero(); return r; }

EG_HD ge_ext ge_generator() {
    ge_ext r;
    r.X = fe_make(0x8f25d51au, 0xc9562d60u, 0x9525a7b2u, 0x692cc760u, 0xfdd6dc5cu, 0xc0a4e231u, 0xcd6e53feu, 0x216936d3u);
    r.Y = fe_make(0x66666658u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u, 0x66666666u);
    r.Z = fe_one();
    r.T = fe_make(0xa5b7dda3u, 0x6dde8ab3u, 0x775152f5u, 0x20f09f80u, 0x64abe37du, 0x66ea4e8eu, 0xd78b7665u, 0x67875f0fu);
    return r;
}

EG_HD void ge_p1p1_to_ext(ge_ext &r, const ge_p1p1 &p) {
    fe_mul(r.X, p.E, p.F); fe_mul(r.Y, p.G, p.H); fe_mul(r.Z, p.F, p.G); fe_mul(r.T, p.E, p.H);
}

// projective only (T left stale): enough when a doubling follows
EG_HD void ge_p1p1_to_proj(ge_ext &r, const ge_p1p1 &p) {
    fe_mul(r.X, p.E, p.F); fe_mul(r.Y, p.G, p.H); fe_mul(r.Z, p.F, p.G);
}

// dbl-2008-hwcd; reads X, Y, Z only
EG_HD void ge_dbl_p1p1(ge_p1p1 &r, const ge_ext &p) {
    fe a, b, c, t;
    fe_sq(a, p.X); fe_sq(b, p.Y);
    fe_sq(c, p.Z); fe_add(c, c, c);
    fe_add(t, p.X, p.Y); fe_sq(t, t);
    fe_add(r.H, a, b);              // A + B
    fe_sub(r.E, t, r.H);            // E = (X+Y)^2 - A - B
    fe_sub(r.G, b, a);              // G = B - A
    fe_sub(r.F, r.G, c);            // F = G - C
    fe_neg(r.H, r.H);               // H = -(A + B)
}

EG_HD void ge_to_cached(ge_cached &c, const ge_ext &p) {
    fe_add(c.YpX, p.Y, p.X); fe_sub(c.YmX, p.Y, p.X); c.Z = p.Z; fe_mul(c.T2d, p.T, fe_const_2d());
}

// add-2008-hwcd-3 with a cached second operand; neg subtracts it
EG_HD void ge_add_cached_p1p1(ge_p1p1 &r, const ge_ext &p, const ge_cached &q, bool neg) {
    fe a, b, c, d, pm, pp;
    fe_sub(a, p.Y, p.X); fe_add(b, p.Y, p.X);
    fe_select(pm, q.YmX, q.YpX, neg); fe_select(pp, q.YpX, q.YmX, neg);
    fe_mul(a, a, pm); fe_mul(b, b, pp);
    fe_mul(c, p.T, q.T2d);
    fe_mul(d, p.Z, q.Z); fe_add(d, d, d);
    fe_sub(r.E, b, a); fe_add(r.H, b, a);
    fe nf, ng;
    fe_sub(nf, d, c); fe_add(ng, d, c);
    fe_select(r.F, nf, ng, neg); fe_select(r.G, ng, nf, neg);     // negating q flips the sign of C
}

// mixed addition with an affine Niels operand (Z2 = 1)
EG_HD void ge_add_niels_p1p1(ge_p1p1 &r, const ge_ext &p, const ge_niels &q, bool neg) {
    fe a, b, c, d, pm, pp;
    fe_sub(a, p.Y, p.X); fe_add(b, p.Y, p.X);
    fe_select(pm, q.ymx, q.ypx, neg); fe_select(pp, q.ypx, q.ymx, neg);
    fe_mul(a, a, pm); fe_mul(b, b, pp);
    fe_mul(c, p.T, q.xy2d);
    fe_add(d, p.Z, p.Z);
    fe_sub(r.E, b, a); fe_add(r.H, b, a);
    fe nf, ng;
    fe_sub(nf, d, c); fe_add(ng, d, c);
    fe_select(r.F, nf, ng, neg); fe_select(r.G, ng, nf, neg);
}

EG_HD void ge_add(ge_ext &r, const ge_ext &p, const ge_ext &q) {
    ge_cached c; ge_p1p1 t;
    ge_to_cached(c, q);
    ge_add_cached_p1p1(t, p, c, false);
    ge_p1p1_to_ext(r, t);
}

EG_HD void ge_sub(ge_ext &r, const ge_ext &p, const ge_ext &q) {
    ge_cached c; ge_p1p1 t;
    ge_to_cached(c, q);
    ge_add_cached_p1p1(t, p, c, true);
    ge_p1p1_to_ext(r, t);
}

EG_HD void ge_dbl(ge_ext &r, const ge_ext &p) { ge_p1p1 t; ge_dbl_p1p1(t, p); ge_p1p1_to_ext(r, t); }

EG_HD void ge_neg(ge_ext &r, const ge_ext &p) { fe_neg(r.X, p.X); r.Y = p.Y; r.Z = p.Z; fe_neg(r.T, p.T); }

// ristretto equality with the identity class: X == 0 or Y == 0
EG_HD bool ge_is_identity(const ge_ext &p) { return fe_iszero(p.X) || fe_iszero(p.Y); }

// ------------------------------------------------------------------ RFC 9496 4.3.1 / 4.3.2

// s: canonical little-endian words of the 32-byte encoding.  Returns false (and the identity) when the
// encoding is rejected: non-canonical, negative, non-square, negative t, or y = 0.
EG_HD bool ge_decode(ge_ext &p, const uint32_t w[8]) {
    fe s, ss, u1, u2, u2s, v, t, inv, denx, deny, x, y;
    bool ok = fe_fromwords_canonical(s, w);
    ok = ok && ((w[0] & 1u) == 0);
    if (!ok) s = fe_zero();
    fe_sq(ss, s);
    fe_sub(u1, fe_one(), ss);
    fe_add(u2, fe_one(), ss);
    fe_sq(u2s, u2);
    fe_sq(t, u1); fe_mul(t, t, fe_const_d()); fe_neg(t, t);
    fe_sub(v, t, u2s);
    fe_mul(t, v, u2s);
    bool was_square = fe_invsqrt(inv, t);
    fe_mul(denx, inv, u2);
    fe_mul(deny, inv, denx); fe_mul(deny, deny, v);
    fe_add(t, s, s); fe_mul(x, t, denx);
    fe_abs(x, x);
    fe_mul(y, u1, deny);
    fe_mul(t, x, y);
    ok = ok && was_square && !fe_isneg(t) && !fe_iszero(y);
    if (!ok) { p = ge_identity(); return false; }
    p.X = x; p.Y = y; p.Z = fe_one(); p.T = t;
    return true;
}

EG_HD void ge_encode(uint32_t w[8], const ge_ext &p) {
    fe u1, u2, t, inv, den1, den2, zinv, ix, iy, ench, x, y, deninv, s;
    const fe sqrtm1 = fe_const_sqrtm1();
    fe_add(u1, p.Z, p.Y); fe_sub(t, p.Z, p.Y); fe_mul(u1, u1, t);
    fe_mul(u2, p.X, p.Y);
    fe_sq(t, u2); fe_mul(t, t, u1);
    fe_invsqrt(inv, t);
    fe_mul(den1, inv, u1);
    fe_mul(den2, inv, u2);
    fe_mul(zinv, den1, den2); fe_mul(zinv, zinv, p.T);
    fe_mul(ix, p.X, sqrtm1);
    fe_mul(iy, p.Y, sqrtm1);
    fe_mul(ench, den1, fe_const_invsqrt_a_minus_d());
    fe_mul(t, p.T, zinv);
    bool rotate = fe_isneg(t);
    fe_select(x, p.X, iy, rotate);
    fe_select(y, p.Y, ix, rotate);
    fe_select(deninv, den2, ench, rotate);
    fe_mul(t, x, zinv);
    fe_cneg(y, y, fe_isneg(t));
    fe_sub(t, p.Z, y);
    fe_mul(s, deninv, t);
    fe_abs(s, s);
    fe_towords(w, s);
}

// ------------------------------------------------------------------ fixed-base tables

#define EG_FIXED_TABLE_ENTRIES 128           // [1..128] F, affine Niels, 96 B each = 12 KB per base
#define EG_FIXED_TABLE_WORDS (EG_FIXED_TABLE_ENTRIES * 24)

EG_HD void ge_niels_load(ge_niels &n, const uint32_t *tbl, int idx) {
    const uint32_t *e = tbl + idx * 24;
    for (int i = 0; i < 8; i++) { n.ypx.v[i] = e[i]; n.ymx.v[i] = e[8 + i]; n.xy2d.v[i] = e[16 + i]; }
}

// one table entry: (k F) normalised to affine Niels form
EG_HD void ge_niels_from_ext(uint32_t out[24], const ge_ext &p) {
    fe zi, x, y, t;
    fe_invert(zi, p.Z);
    fe_mul(x, p.X, zi); fe_mul(y, p.Y, zi);
    fe_add(t, y, x); fe_towords(out, t);
    fe_sub(t, y, x); fe_towords(out + 8, t);
    fe_mul(t, x, y); fe_mul(t, t, fe_const_2d()); fe_towords(out + 16, t);
}

// ------------------------------------------------------------------ scalar recoding

// 4-bit signed windows: a + 0x888...8 so that digit_i = nibble_i - 8 in [-8, 7]  (a < 2^253)
EG_HD void sc_recode4(uint32_t out[8], const sc &a) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + 0x88888888u; out[i] = (uint32_t)c; c >>= 32; }
}
// 8-bit signed windows: a + 0x8080...80, digit_i = byte_i - 128 in [-128, 127]
EG_HD void sc_recode8(uint32_t out[8], const sc &a) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) { c += (uint64_t)a.v[i] + 0x80808080u; out[i] = (uint32_t)c; c >>= 32; }
}
EG_HD int sc_digit4(const uint32_t r[8], int i) { return (int)((r[i >> 3] >> ((i & 7) * 4)) & 15u) - 8; }
EG_HD int sc_digit8(const uint32_t r[8], int i) { return (int)((r[i >> 2] >> ((i & 3) * 8)) & 255u) - 128; }

// ------------------------------------------------------------------ the multi-scalar chain

// acc = sum_{v<NV} a_v P_v + sum_{f<NF} b_f F_f
//   P_v : per-item points (extended), windows of 4 bits over a per-thread table of [1..8]P_v
//   F_f : fixed bases with 128-entry affine tables (shared or global memory), windows of 8 bits
// One shared doubling chain of 252 doublings (Straus).  NV, NF are compile-time.
template <int NV, int NF>
EG_HD void ge_msm_chain(ge_ext &out, const ge_ext *P, const sc *a, const uint32_t *const *ftab, const sc *b) {
    ge_cached tbl[NV > 0 ? NV : 1][8];
    uint32_t ra[NV > 0 ? NV : 1][8];
    uint32_t rb[NF > 0 ? NF : 1][8];
#pragma unroll 1
    for (int v = 0; v < NV; v++) {
        ge_ext cur = P[v];
        ge_cached c1;
        ge_to_cached(c1, cur);
        tbl[v][0] = c1;
#pragma unroll 1
        for (int k = 1; k < 8; k++) {
            ge_p1p1 t;
            ge_add_cached_p1p1(t, cur, c1, false);
            ge_p1p1_to_ext(cur, t);
            ge_to_cached(tbl[v][k], cur);
        }
        sc_recode4(ra[v], a[v]);
    }
    for (int f = 0; f < NF; f++) sc_recode8(rb[f], b[f]);

    ge_ext acc = ge_identity();
    ge_p1p1 t;
#pragma unroll 1
    for (int i = 63; i >= 0; i--) {
        // acc = 16 acc ; the last doubling also produces T for the additions below
#pragma unroll 1
        for (int k = 0; k < 3; k++) { ge_dbl_p1p1(t, acc); ge_p1p1_to_proj(acc, t); }
        ge_dbl_p1p1(t, acc); ge_p1p1_to_ext(acc, t);
#pragma unroll 1
        for (int v = 0; v < NV; v++) {
            int d = sc_digit4(ra[v], i);
            if (d != 0) {
                int m = d < 0 ? -d : d;
                ge_add_cached_p1p1(t, acc, tbl[v][m - 1], d < 0);
                ge_p1p1_to_ext(acc, t);
            }
        }
        if ((i & 1) == 0) {
#pragma unroll 1
            for (int f = 0; f < NF; f++) {
                int d = sc_digit8(rb[f], i >> 1);
                if (d != 0) {
                    int m = d < 0 ? -d : d;
                    ge_niels n;
                    ge_niels_load(n, ftab[f], m - 1);
                    ge_add_niels_p1p1(t, acc, n, d < 0);
                    ge_p1p1_to_ext(acc, t);
                }
            }
        }
    }
    out = acc;
}


// Same chain with run-time term counts (nv <= MAXV, nf <= 2), for the equations that are not of the [a]P + [b]F shape:
// share verification (two per-item bases), SumOfSquaresProof (G, K and one per-item base; (n+2)-term sums) and
// Lagrange recombination.  Replaces the general vartime_multi_mul (ristretto.rs:139-146).
template <int MAXV>
EG_HD void ge_msm_chain_rt(ge_ext &out, int nv, const ge_ext *P, const sc *a, int nf, const uint32_t *const *ftab, const sc *b) {
    ge_cached tbl[MAXV][8];
    uint32_t ra[MAXV][8];
    uint32_t rb[2][8];
#pragma unroll 1
    for (int v = 0; v < nv; v++) {
        ge_ext cur = P[v];
        ge_cached c1;
        ge_to_cached(c1, cur);
        tbl[v][0] = c1;
#pragma unroll 1
        for (int k = 1; k < 8; k++) {
            ge_p1p1 t;
            ge_add_cached_p1p1(t, cur, c1, false);
            ge_p1p1_to_ext(cur, t);
            ge_to_cached(tbl[v][k], cur);
        }
        sc_recode4(ra[v], a[v]);
    }
#pragma unroll 1
    for (int f = 0; f < nf; f++) sc_recode8(rb[f], b[f]);
    ge_ext acc = ge_identity();
    ge_p1p1 t;
#pragma unroll 1
    for (int i = 63; i >= 0; i--) {
#pragma unroll 1
        for (int k = 0; k < 3; k++) { ge_dbl_p1p1(t, acc); ge_p1p1_to_proj(acc, t); }
        ge_dbl_p1p1(t, acc); ge_p1p1_to_ext(acc, t);
#pragma unroll 1
        for (int v = 0; v < nv; v++) {
            int d = sc_digit4(ra[v], i);
            if (d != 0) {
                int m = d < 0 ? -d : d;
                ge_add_cached_p1p1(t, acc, tbl[v][m - 1], d < 0);
                ge_p1p1_to_ext(acc, t);
            }
        }
        if ((i & 1) == 0) {
#pragma unroll 1
            for (int f = 0; f < nf; f++) {
                int d = sc_digit8(rb[f], i >> 1);
                if (d != 0) {
                    int m = d < 0 ? -d : d;
                    ge_niels n;
                    ge_niels_load(n, ftab[f], m - 1);
                    ge_add_niels_p1p1(t, acc, n, d < 0);
                    ge_p1p1_to_ext(acc, t);
                }
            }
        }
    }
    out = acc;
}

}  // namespace eg
