// kernels.cuh -- batch kernels of the verification hot path (sm_100a).
//
// Data layout in HBM (DESIGN.md "Layout"): per-batch scratch is *planar* -- word w of object q of item i
// lives at base[(q * W + w) * n + i] -- so that a warp of 32 consecutive items reads/writes 128 contiguous
// bytes per word.  Inputs stay in the reference's own array-of-structures byte layouts (to_bytes forms).
//
// The kernel bodies are __host__ __device__ functions of (item, slot) so that tests/hostsim can run the very
// same logic on the CPU (test harness only; the product library launches only the __global__ wrappers).
#pragma once
#include "fe.cuh"
#include "ge.cuh"
#include "merlin.cuh"
#include "sc.cuh"

namespace eg {

#define EG_MAX_SLOTS 40
#define EG_MAX_RINGS 64

// ------------------------------------------------------------------ planar accessors

EG_HD void planar_store_words(uint32_t *base, size_t n, size_t q, int W, size_t i, const uint32_t *w) {
    for (int k = 0; k < W; k++) base[(q * W + k) * n + i] = w[k];
}
EG_HD void planar_load_words(uint32_t *w, const uint32_t *base, size_t n, size_t q, int W, size_t i) {
    for (int k = 0; k < W; k++) w[k] = base[(q * W + k) * n + i];
}
EG_HD void planar_store_point(uint32_t *base, size_t n, size_t q, size_t i, const ge_ext &p) {
    for (int k = 0; k < 8; k++) {
        base[(q * 32 + k) * n + i] = p.X.v[k];
        base[(q * 32 + 8 + k) * n + i] = p.Y.v[k];
        base[(q * 32 + 16 + k) * n + i] = p.Z.v[k];
        base[(q * 32 + 24 + k) * n + i] = p.T.v[k];
    }
}
EG_HD void planar_load_point(ge_ext &p, const uint32_t *base, size_t n, size_t q, size_t i) {
    for (int k = 0; k < 8; k++) {
        p.X.v[k] = base[(q * 32 + k) * n + i];
        p.Y.v[k] = base[(q * 32 + 8 + k) * n + i];
        p.Z.v[k] = base[(q * 32 + 16 + k) * n + i];
        p.T.v[k] = base[(q * 32 + 24 + k) * n + i];
    }
}

// 32 bytes at an arbitrary (4-byte aligned or not) address -> little-endian words
EG_HD void load32_bytes(uint32_t w[8], const uint8_t *p) {
    if ((((uintptr_t)p) & 3u) == 0) {
        const uint32_t *q = (const uint32_t *)p;
        for (int k = 0; k < 8; k++) w[k] = q[k];
    } else {
        for (int k = 0; k < 8; k++)
            w[k] = (uint32_t)p[4 * k] | ((uint32_t)p[4 * k + 1] << 8) | ((uint32_t)p[4 * k + 2] << 16) | ((uint32_t)p[4 * k + 3] << 24);
    }
}
EG_HD void store32_bytes(uint8_t *p, const uint32_t w[8]) {
    if ((((uintptr_t)p) & 3u) == 0) {
        uint32_t *q = (uint32_t *)p;
        for (int k = 0; k < 8; k++) q[k] = w[k];
    } else {
        for (int k = 0; k < 8; k++) { p[4 * k] = (uint8_t)w[k]; p[4 * k + 1] = (uint8_t)(w[k] >> 8); p[4 * k + 2] = (uint8_t)(w[k] >> 16); p[4 * k + 3] = (uint8_t)(w[k] >> 24); }
    }
}

// ------------------------------------------------------------------ decode / validate

// Input buffers of a batch call (array-of-structures, reference byte layouts)
struct in_bufs {
    const uint8_t *buf[4];
    uint32_t stride[4];         // bytes per item in buf[k]
};

struct decode_slot {
    uint8_t buf;                // which input buffer
    uint8_t want_enc;           // also store the 8 encoding words (for transcript "enc" messages)
    uint16_t enc_index;         // planar index in enc (8 words each)
    uint32_t offset;            // byte offset inside the item
    uint32_t p_index;           // planar point index
};

struct decode_params {
    in_bufs in;
    size_t n;
    int n_slots;
    decode_slot slots[EG_MAX_SLOTS];
    uint32_t *pts;              // planar points
    uint32_t *enc;              // planar encodings
    uint32_t *flags;            // per item: bit 0 = malformed
};

EG_HD void decode_body(const decode_params &P, size_t item, int slot) {
    const decode_slot &s = P.slots[slot];
    uint32_t w[8];
    load32_bytes(w, P.in.buf[s.buf] + item * P.in.stride[s.buf] + s.offset);
    ge_ext p;
    bool ok = ge_decode(p, w);
    planar_store_point(P.pts, P.n, s.p_index, item, p);
    if (s.want_enc) planar_store_words(P.enc, P.n, s.enc_index, 8, item, w);
    if (!ok) {
#if defined(__CUDA_ARCH__)
        atomicOr(&P.flags[item], 1u);
#else
        P.flags[item] |= 1u;
#endif
    }
}

struct scalar_slot { uint8_t buf; uint32_t offset; uint32_t count; };   // `count` consecutive scalars

struct scalars_params {
    in_bufs in;
    size_t n;
    int n_slots;
    scalar_slot slots[8];
    uint32_t *flags;
};

EG_HD void scalars_body(const scalars_params &P, size_t item) {
    bool ok = true;
    for (int s = 0; s < P.n_slots; s++) {
        const uint8_t *base = P.in.buf[P.slots[s].buf] + item * P.in.stride[P.slots[s].buf] + P.slots[s].offset;
        for (uint32_t k = 0; k < P.slots[s].count; k++) {
            uint32_t w[8];
            load32_bytes(w, base + 32 * k);
            ok = ok && sc_is_canonical_words(w);
        }
    }
    if (!ok) {
#if defined(__CUDA_ARCH__)
        atomicOr(&P.flags[item], 1u);
#else
        P.flags[item] |= 1u;
#endif
    }
}

// ------------------------------------------------------------------ commitments (the hot kernel)

// One slot = one verification equation side:  C = [-e] (P - X) + [s] F
//   reference: ring.rs:342-350, log_equality.rs:160-164 (vartime_double_mul_generator / vartime_multi_mul)
struct commit_slot {
    uint32_t p_index;           // planar index of P
    int32_t adm_index;          // admissible value X (cached form) to subtract, -1 for none (x = identity)
    uint8_t base;               // 0: G, 1: K
    uint8_t e_planar;           // challenge from the planar `chal` buffer (1) or from an input buffer (0)
    uint8_t e_buf, s_buf;       // input buffers of the challenge / response scalars
    uint32_t e_offset;          // byte offset in item (e_planar = 0) or planar scalar index (e_planar = 1)
    uint32_t s_offset;          // byte offset of the response scalar in its item
    uint32_t out_index;         // planar index of the 8-word commitment encoding
};

struct commit_params {
    in_bufs in;
    size_t n;
    int n_slots;
    commit_slot slots[EG_MAX_SLOTS];
    const uint32_t *pts;
    const uint32_t *chal;       // planar scalars (8 words)
    uint32_t *commit;           // planar encodings (8 words)
    const uint32_t *adm;        // admissible values: cached points, 32 words each (YpX, YmX, Z, T2d)
    const uint32_t *table_g;    // 128-entry affine tables (24 words per entry)
    const uint32_t *table_k;
};

EG_HD void commit_body(const commit_params &P, size_t item, int slot, const uint32_t *tab_g, const uint32_t *tab_k) {
    const commit_slot &s = P.slots[slot];
    ge_ext pt;
    planar_load_point(pt, P.pts, P.n, s.p_index, item);
    if (s.adm_index >= 0) {
        ge_cached x;
        const uint32_t *a = P.adm + (size_t)s.adm_index * 32;
        for (int k = 0; k < 8; k++) { x.YpX.v[k] = a[k]; x.YmX.v[k] = a[8 + k]; x.Z.v[k] = a[16 + k]; x.T2d.v[k] = a[24 + k]; }
        ge_p1p1 t;
        ge_add_cached_p1p1(t, pt, x, true);
        ge_p1p1_to_ext(pt, t);
    }
    uint32_t w[8];
    sc e, r, ne;
    if (s.e_planar) planar_load_words(w, P.chal, P.n, s.e_offset, 8, item);
    else load32_bytes(w, P.in.buf[s.e_buf] + item * P.in.stride[s.e_buf] + s.e_offset);
    bool ok = sc_from_words(e, w);
    load32_bytes(w, P.in.buf[s.s_buf] + item * P.in.stride[s.s_buf] + s.s_offset);
    ok = sc_from_words(r, w) && ok;
    if (!ok) { e = sc_zero(); r = sc_zero(); }      // malformed items are already flagged; keep the math defined
    sc_neg(ne, e);
    ge_ext acc;
    const uint32_t *ft[1] = {s.base ? tab_k : tab_g};
    ge_msm_chain<1, 1>(acc, &pt, &ne, ft, &r);
    ge_encode(w, acc);
    planar_store_words(P.commit, P.n, s.out_index, 8, item, w);
}

// ------------------------------------------------------------------ ring transcripts

// One slot = one ring in one hash stage: e_{j+1} = H(ring transcript, j, R_G(j), R_K(j))   (ring.rs:325-360)
struct ring_hash_slot {
    uint32_t ring_index;        // value appended as "i"
    uint32_t eq_index;          // value appended as "j"
    uint32_t enc_index;         // planar index of enc(R); enc(B) is enc_index + 1
    uint32_t commit_index;      // planar index of R_G; R_K is commit_index + 1
    uint32_t chal_index;        // planar index of the output challenge
};

struct ring_hash_params {
    size_t n;
    int n_slots;
    ring_hash_slot slots[EG_MAX_SLOTS];
    transcript prefix;          // state after RingProof::initialize_transcript (ring.rs:290-293)
    const uint32_t *enc;
    const uint32_t *commit;
    uint32_t *chal;
};

EG_HD void ring_hash_body(const ring_hash_params &P, size_t item, int slot) {
    const ring_hash_slot &s = P.slots[slot];
    transcript t = P.prefix;
    uint32_t w[16];
    merlin_append_message(t, EG_LBL("dom-sep"), (const uint8_t *)"ring_enc", 8);
    planar_load_words(w, P.enc, P.n, s.enc_index, 8, item);
    planar_load_words(w + 8, P.enc, P.n, s.enc_index + 1, 8, item);
    merlin_append_words(t, EG_LBL("enc"), w, 16);
    merlin_append_u64(t, EG_LBL("i"), s.ring_index);
    merlin_append_u64(t, EG_LBL("j"), s.eq_index);
    planar_load_words(w, P.commit, P.n, s.commit_index, 8, item);
    merlin_append_words(t, EG_LBL("R_G"), w, 8);
    planar_load_words(w, P.commit, P.n, s.commit_index + 1, 8, item);
    merlin_append_words(t, EG_LBL("R_K"), w, 8);
    sc c;
    merlin_challenge_scalar(t, EG_LBL("c"), c);
    planar_store_words(P.chal, P.n, s.chal_index, 8, item, c.v);
}

// Outer transcript: absorb every ring's terminal commitments, compare with the common challenge (ring.rs:364-373)
struct ring_final_params {
    in_bufs in;
    size_t n;
    uint32_t n_rings;
    uint32_t commit_index0;     // ring r's terminal (R_G, R_K) at commit_index0 + 2r, +1
    uint8_t proof_buf;          // input buffer holding the ring proof; common challenge at offset `cc_offset`
    uint32_t cc_offset;
    transcript prefix;          // state after initialize_transcript
    const uint32_t *commit;
    uint32_t *result;           // per item: 1 = challenge matches
};

EG_HD void ring_final_body(const ring_final_params &P, size_t item) {
    transcript t = P.prefix;
    uint32_t w[8];
#pragma unroll 1
    for (uint32_t r = 0; r < P.n_rings; r++) {
        planar_load_words(w, P.commit, P.n, P.commit_index0 + 2 * r, 8, item);
        merlin_append_words(t, EG_LBL("R_G"), w, 8);
        planar_load_words(w, P.commit, P.n, P.commit_index0 + 2 * r + 1, 8, item);
        merlin_append_words(t, EG_LBL("R_K"), w, 8);
    }
    sc c, cc;
    merlin_challenge_scalar(t, EG_LBL("c"), c);
    load32_bytes(w, P.in.buf[P.proof_buf] + item * P.in.stride[P.proof_buf] + P.cc_offset);
    bool ok = sc_from_words(cc, w);
    P.result[item] = (ok && sc_eq(c, cc)) ? 1u : 0u;
}

// LogEqualityProof transcript tail (log_equality.rs:169-179): prefix holds start_proof + "K"
struct logeq_final_params {
    in_bufs in;
    size_t n;
    uint32_t pow_enc_index;     // planar enc of the two powers ([r]G, [r]K) at pow_enc_index, +1
    uint32_t commit_index;      // planar commitments ([x]G, [x]K) at commit_index, +1
    uint8_t proof_buf;
    uint32_t c_offset;          // byte offset of the challenge in the proof item
    uint8_t prefix_per_item;    // 1: the transcript prefix is finished per item (share proofs: "K" = enc(R))
    uint32_t key_enc_index;     // planar enc index of the per-item log base (prefix_per_item = 1)
    transcript prefix;
    const uint32_t *enc;
    const uint32_t *commit;
    uint32_t *result;
};

EG_HD void logeq_final_body(const logeq_final_params &P, size_t item) {
    transcript t = P.prefix;
    uint32_t w[8];
    if (P.prefix_per_item) {
        merlin_append_message(t, EG_LBL("dom-sep"), (const uint8_t *)"log_eq", 6);
        planar_load_words(w, P.enc, P.n, P.key_enc_index, 8, item);
        merlin_append_words(t, EG_LBL("K"), w, 8);
    }
    planar_load_words(w, P.enc, P.n, P.pow_enc_index, 8, item);
    merlin_append_words(t, EG_LBL("[r]G"), w, 8);
    planar_load_words(w, P.enc, P.n, P.pow_enc_index + 1, 8, item);
    merlin_append_words(t, EG_LBL("[r]K"), w, 8);
    planar_load_words(w, P.commit, P.n, P.commit_index, 8, item);
    merlin_append_words(t, EG_LBL("[x]G"), w, 8);
    planar_load_words(w, P.commit, P.n, P.commit_index + 1, 8, item);
    merlin_append_words(t, EG_LBL("[x]K"), w, 8);
    sc c, cc;
    merlin_challenge_scalar(t, EG_LBL("c"), c);
    load32_bytes(w, P.in.buf[P.proof_buf] + item * P.in.stride[P.proof_buf] + P.c_offset);
    bool ok = sc_from_words(cc, w);
    P.result[item] = (ok && sc_eq(c, cc)) ? 1u : 0u;
}

// ------------------------------------------------------------------ derived ciphertexts

// EncryptedChoice: sum of the option ciphertexts (choice.rs:363) and the sum-proof powers (choice.rs:83-86):
// side 0: sum R ; side 1: sum B - G.  Stores the point and its encoding (both are hashed / multiplied later).
struct choice_sum_params {
    size_t n;
    uint32_t options;
    uint32_t out_p_index;       // points at out_p_index + side
    uint32_t out_enc_index;     // encodings at out_enc_index + side
    uint32_t *pts;
    uint32_t *enc;
};

EG_HD void choice_sum_body(const choice_sum_params &P, size_t item, int side) {
    ge_ext acc, q;
    planar_load_point(acc, P.pts, P.n, side, item);
#pragma unroll 1
    for (uint32_t k = 1; k < P.options; k++) {
        planar_load_point(q, P.pts, P.n, 2 * k + side, item);
        ge_add(acc, acc, q);
    }
    if (side == 1) ge_sub(acc, acc, ge_generator());
    planar_store_point(P.pts, P.n, P.out_p_index + side, item, acc);
    uint32_t w[8];
    ge_encode(w, acc);
    planar_store_words(P.enc, P.n, P.out_enc_index + side, 8, item, w);
}

// RangeProof: last ring's ciphertext = ct - sum(partial) (range.rs:564-572); point + encoding, per side
struct range_last_params {
    size_t n;
    uint32_t n_partial;         // n_rings - 1 partial ciphertexts at point indexes 2k + side (k < n_partial)
    uint32_t ct_p_index;        // main ciphertext at ct_p_index + side
    uint32_t out_p_index;       // last ring ciphertext at out_p_index + side
    uint32_t out_enc_index;
    uint32_t *pts;
    uint32_t *enc;
};

EG_HD void range_last_body(const range_last_params &P, size_t item, int side) {
    ge_ext acc, q;
    planar_load_point(acc, P.pts, P.n, P.ct_p_index + side, item);
#pragma unroll 1
    for (uint32_t k = 0; k < P.n_partial; k++) {
        planar_load_point(q, P.pts, P.n, 2 * k + side, item);
        ge_sub(acc, acc, q);
    }
    planar_store_point(P.pts, P.n, P.out_p_index + side, item, acc);
    uint32_t w[8];
    ge_encode(w, acc);
    planar_store_words(P.enc, P.n, P.out_enc_index + side, 8, item, w);
}

// ------------------------------------------------------------------ verdicts

// verdict precedence: malformed -> first failing check in the reference's order
struct verdict_params {
    size_t n;
    const uint32_t *flags;
    int n_checks;
    const uint32_t *check[8];   // per item 1 = passed
    size_t check_stride[8];     // item stride (1, or k when the check array holds k results per item)
    size_t check_offset[8];
    uint8_t code[8];            // verdict when check k is the first to fail
    uint8_t *verdicts;
};

EG_HD void verdict_body(const verdict_params &P, size_t item) {
    uint8_t v = 0;
    if (P.flags[item] & 1u) v = 1;
    else {
        for (int k = 0; k < P.n_checks; k++)
            if (!P.check[k][item * P.check_stride[k] + P.check_offset[k]]) { v = P.code[k]; break; }
    }
    P.verdicts[item] = v;
}

}  // namespace eg

// ------------------------------------------------------------------ setup bodies (once per context / receiver / range)

namespace eg {

// table[k] = (k+1) F in affine Niels form, F given as an encoding; status: 0 ok, 1 undecodable, 2 identity
EG_HD void build_table_body(int tidx, const uint32_t *enc_words, int use_generator, uint32_t *table, uint32_t *status) {
    ge_ext F;
    bool ok = true;
    if (use_generator) F = ge_generator();
    else {
        uint32_t w[8];
        for (int k = 0; k < 8; k++) w[k] = enc_words[k];
        ok = ge_decode(F, w);
    }
    if (tidx == 0) *status = !ok ? 1u : (ge_is_identity(F) ? 2u : 0u);
    if (!ok) return;
    int m = tidx + 1;               // 1..128
    ge_ext acc = ge_identity();
#pragma unroll 1
    for (int bit = 7; bit >= 0; bit--) {
        ge_dbl(acc, acc);
        if ((m >> bit) & 1) ge_add(acc, acc, F);
    }
    ge_niels_from_ext(table + tidx * 24, acc);
}

// adm[tid] = cached form of [values[tid]] G   (PreparedRange::new range.rs:341-355)
EG_HD void admissible_body(int tid, const uint64_t *values, uint32_t *adm) {
    uint64_t v = values[tid];
    ge_ext G = ge_generator(), acc = ge_identity();
#pragma unroll 1
    for (int bit = 63; bit >= 0; bit--) {
        ge_dbl(acc, acc);
        if ((v >> bit) & 1) ge_add(acc, acc, G);
    }
    ge_cached c;
    ge_to_cached(c, acc);
    uint32_t *o = adm + (size_t)tid * 32;
    for (int k = 0; k < 8; k++) { o[k] = c.YpX.v[k]; o[8 + k] = c.YmX.v[k]; o[16 + k] = c.Z.v[k]; o[24 + k] = c.T2d.v[k]; }
}

EG_HD void point_to_words32(uint32_t *o, const ge_ext &p) {
    for (int k = 0; k < 8; k++) { o[k] = p.X.v[k]; o[8 + k] = p.Y.v[k]; o[16 + k] = p.Z.v[k]; o[24 + k] = p.T.v[k]; }
}
EG_HD void point_from_words32(ge_ext &p, const uint32_t *o) {
    for (int k = 0; k < 8; k++) { p.X.v[k] = o[k]; p.Y.v[k] = o[8 + k]; p.Z.v[k] = o[16 + k]; p.T.v[k] = o[24 + k]; }
}

EG_HD void elements_validate_body(size_t i, const uint8_t *enc, uint8_t *ok) {
    uint32_t w[8];
    load32_bytes(w, enc + 32 * i);
    ge_ext p;
    ok[i] = ge_decode(p, w) ? 1 : 0;
}

EG_HD void scalars_validate_body(size_t i, const uint8_t *s, uint8_t *ok) {
    uint32_t w[8];
    load32_bytes(w, s + 32 * i);
    ok[i] = sc_is_canonical_words(w) ? 1 : 0;
}

EG_HD void scalars_from_wide_body(size_t i, const uint8_t *wide, uint8_t *out) {
    uint32_t w[16];
    load32_bytes(w, wide + 64 * i);
    load32_bytes(w + 8, wide + 64 * i + 32);
    sc r;
    sc_from_wide_words(r, w);
    store32_bytes(out + 32 * i, r.v);
}

// out = [a]A + [b]G (mode 0) or [b]G (mode 1): Group::vartime_double_mul_generator / mul_generator
EG_HD void double_mul_body(size_t i, const uint8_t *a, const uint8_t *A, const uint8_t *b, int mode, const uint32_t *tab_g,
                           uint8_t *out, uint8_t *okv) {
    uint32_t w[8];
    sc sa = sc_zero(), sb;
    ge_ext P = ge_identity();
    bool ok = true;
    if (mode == 0) {
        load32_bytes(w, a + 32 * i); ok = sc_from_words(sa, w) && ok;
        load32_bytes(w, A + 32 * i); ok = ge_decode(P, w) && ok;
    }
    load32_bytes(w, b + 32 * i); ok = sc_from_words(sb, w) && ok;
    ge_ext acc = ge_identity();
    if (ok) {
        const uint32_t *ft[1] = {tab_g};
        ge_msm_chain<1, 1>(acc, &P, &sa, ft, &sb);
    }
    ge_encode(w, acc);
    store32_bytes(out + 32 * i, w);
    if (okv) okv[i] = ok ? 1 : 0;
}

// out[c] = sum_p parts[p][c], ciphertext halves independently; tid over (c, half)
EG_HD void ciphertexts_sum_body(size_t tid, const uint8_t *parts, size_t n_parts, size_t n_cts, uint8_t *out, uint32_t *bad) {
    ge_ext acc = ge_identity(), q;
    uint32_t w[8];
    for (size_t p = 0; p < n_parts; p++) {
        load32_bytes(w, parts + (p * n_cts * 2 + tid) * 32);
        if (!ge_decode(q, w)) {
#if defined(__CUDA_ARCH__)
            atomicOr(bad, 1u);
#else
            *bad |= 1u;
#endif
        }
        ge_add(acc, acc, q);
    }
    ge_encode(w, acc);
    store32_bytes(out + 32 * tid, w);
}

}  // namespace eg
