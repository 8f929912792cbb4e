// merlin.cuh -- Keccak-f[1600], the STROBE-128 subset Merlin uses, and Merlin v1.0 transcripts, as
// __host__ __device__ code so that challenges are recomputed on the GPU with no host round trip and the
// per-election transcript prefixes can be prepared once on the host with the same code.
//
// Replaces merlin 3.0.0 (+ keccak 0.1.6) as used by src/proofs/mod.rs:29-57 (TranscriptForGroup) and
// src/group/mod.rs:37-62 (RandomBytesProvider).  Framing: append_message(l, m) = meta_ad(l) ||
// meta_ad(LE32(|m|), more) || ad(m); challenge_bytes(l, n) = meta_ad(l) || meta_ad(LE32(n), more) || prf(n).
#pragma once
#include <stdint.h>
#include "fe.cuh"
#include "sc.cuh"

namespace eg {

#define EG_STROBE_R 166

struct transcript {
    uint64_t st[25];     // Keccak state, little-endian lanes
    uint32_t pos, pos_begin;
};

EG_HD uint64_t rotl64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

static EG_HD_NOINLINE void keccak_f1600(uint64_t a[25]) {
    const uint64_t RC[24] = {
        0x0000000000000001ULL, 0x0000000000008082ULL, 0x800000000000808aULL, 0x8000000080008000ULL,
        0x000000000000808bULL, 0x0000000080000001ULL, 0x8000000080008081ULL, 0x8000000000008009ULL,
        0x000000000000008aULL, 0x0000000000000088ULL, 0x0000000080008009ULL, 0x000000008000000aULL,
        0x000000008000808bULL, 0x800000000000008bULL, 0x8000000000008089ULL, 0x8000000000008003ULL,
        0x8000000000008002ULL, 0x8000000000000080ULL, 0x000000000000800aULL, 0x800000008000000aULL,
        0x8000000080008081ULL, 0x8000000000008080ULL, 0x0000000080000001ULL, 0x8000000080008008ULL};
    uint64_t s00 = a[0], s01 = a[1], s02 = a[2], s03 = a[3], s04 = a[4];
    uint64_t s05 = a[5], s06 = a[6], s07 = a[7], s08 = a[8], s09 = a[9];
    uint64_t s10 = a[10], s11 = a[11], s12 = a[12], s13 = a[13], s14 = a[14];
    uint64_t s15 = a[15], s16 = a[16], s17 = a[17], s18 = a[18], s19 = a[19];
    uint64_t s20 = a[20], s21 = a[21], s22 = a[22], s23 = a[23], s24 = a[24];
#pragma unroll 1
    for (int round = 0; round < 24; round++) {
        // theta
        uint64_t c0 = s00 ^ s05 ^ s10 ^ s15 ^ s20, c1 = s01 ^ s06 ^ s11 ^ s16 ^ s21, c2 = s02 ^ s07 ^ s12 ^ s17 ^ s22,
                 c3 = s03 ^ s08 ^ s13 ^ s18 ^ s23, c4 = s04 ^ s09 ^ s14 ^ s19 ^ s24;
        uint64_t d0 = c4 ^ rotl64(c1, 1), d1 = c0 ^ rotl64(c2, 1), d2 = c1 ^ rotl64(c3, 1), d3 = c2 ^ rotl64(c4, 1),
                 d4 = c3 ^ rotl64(c0, 1);
        s00 ^= d0; s05 ^= d0; s10 ^= d0; s15 ^= d0; s20 ^= d0;
        s01 ^= d1; s06 ^= d1; s11 ^= d1; s16 ^= d1; s21 ^= d1;
        s02 ^= d2; s07 ^= d2; s12 ^= d2; s17 ^= d2; s22 ^= d2;
        s03 ^= d3; s08 ^= d3; s13 ^= d3; s18 ^= d3; s23 ^= d3;
        s04 ^= d4; s09 ^= d4; s14 ^= d4; s19 ^= d4; s24 ^= d4;
        // rho + pi: b[y][2x+3y] = rot(a[x][y])
        uint64_t b00 = s00,              b10 = rotl64(s01, 1),  b20 = rotl64(s02, 62), b05 = rotl64(s03, 28), b15 = rotl64(s04, 27);
        uint64_t b16 = rotl64(s05, 36),  b01 = rotl64(s06, 44), b11 = rotl64(s07, 6),  b21 = rotl64(s08, 55), b06 = rotl64(s09, 20);
        uint64_t b07 = rotl64(s10, 3),   b17 = rotl64(s11, 10), b02 = rotl64(s12, 43), b12 = rotl64(s13, 25), b22 = rotl64(s14, 39);
        uint64_t b23 = rotl64(s15, 41),  b08 = rotl64(s16, 45), b18 = rotl64(s17, 15), b03 = rotl64(s18, 21), b13 = rotl64(s19, 8);
        uint64_t b14 = rotl64(s20, 18),  b24 = rotl64(s21, 2),  b09 = rotl64(s22, 61), b19 = rotl64(s23, 56), b04 = rotl64(s24, 14);
        // chi
        s00 = b00 ^ (~b01 & b02); s01 = b01 ^ (~b02 & b03); s02 = b02 ^ (~b03 & b04); s03 = b03 ^ (~b04 & b00); s04 = b04 ^ (~b00 & b01);
        s05 = b05 ^ (~b06 & b07); s06 = b06 ^ (~b07 & b08); s07 = b07 ^ (~b08 & b09); s08 = b08 ^ (~b09 & b05); s09 = b09 ^ (~b05 & b06);
        s10 = b10 ^ (~b11 & b12); s11 = b11 ^ (~b12 & b13); s12 = b12 ^ (~b13 & b14); s13 = b13 ^ (~b14 & b10); s14 = b14 ^ (~b10 & b11);
        s15 = b15 ^ (~b16 & b17); s16 = b16 ^ (~b17 & b18); s17 = b17 ^ (~b18 & b19); s18 = b18 ^ (~b19 & b15); s19 = b19 ^ (~b15 & b16);
        s20 = b20 ^ (~b21 & b22); s21 = b21 ^ (~b22 & b23); s22 = b22 ^ (~b23 & b24); s23 = b23 ^ (~b24 & b20); s24 = b24 ^ (~b20 & b21);
        s00 ^= RC[round];
    }
    a[0] = s00; a[1] = s01; a[2] = s02; a[3] = s03; a[4] = s04; a[5] = s05; a[6] = s06; a[7] = s07; a[8] = s08; a[9] = s09;
    a[10] = s10; a[11] = s11; a[12] = s12; a[13] = s13; a[14] = s14; a[15] = s15; a[16] = s16; a[17] = s17; a[18] = s18;
    a[19] = s19; a[20] = s20; a[21] = s21; a[22] = s22; a[23] = s23; a[24] = s24;
}

EG_HD void st_xor_byte(transcript &t, uint32_t pos, uint8_t b) { t.st[pos >> 3] ^= (uint64_t)b << ((pos & 7) * 8); }
EG_HD uint8_t st_get_byte(const transcript &t, uint32_t pos) { return (uint8_t)(t.st[pos >> 3] >> ((pos & 7) * 8)); }
EG_HD void st_clear_byte(transcript &t, uint32_t pos) { t.st[pos >> 3] &= ~((uint64_t)0xff << ((pos & 7) * 8)); }

static EG_HD_NOINLINE void strobe_run_f(transcript &t) {
    st_xor_byte(t, t.pos, (uint8_t)t.pos_begin);
    st_xor_byte(t, t.pos + 1, 0x04);
    st_xor_byte(t, EG_STROBE_R + 1, 0x80);
    keccak_f1600(t.st);
    t.pos = 0; t.pos_begin = 0;
}

static EG_HD_NOINLINE void strobe_absorb(transcript &t, const uint8_t *d, uint32_t n) {
#pragma unroll 1
    for (uint32_t i = 0; i < n; i++) {
        st_xor_byte(t, t.pos, d[i]);
        if (++t.pos == EG_STROBE_R) strobe_run_f(t);
    }
}

// absorb little-endian words (n_bytes multiple of 4): the hot path appends 32-byte elements and u64s
static EG_HD_NOINLINE void strobe_absorb_words(transcript &t, const uint32_t *w, uint32_t n_words) {
#pragma unroll 1
    for (uint32_t i = 0; i < n_words; i++) {
        uint32_t x = w[i];
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            st_xor_byte(t, t.pos, (uint8_t)(x >> (8 * k)));
            if (++t.pos == EG_STROBE_R) strobe_run_f(t);
        }
    }
}

static EG_HD_NOINLINE void strobe_begin_op(transcript &t, uint8_t flags) {
    uint8_t hdr[2] = {(uint8_t)t.pos_begin, flags};
    t.pos_begin = t.pos + 1;
    strobe_absorb(t, hdr, 2);
    if ((flags & (4 | 32)) && t.pos != 0) strobe_run_f(t);     // C or K forces a permutation
}

#define EG_FLAG_META_AD 0x12
#define EG_FLAG_AD 0x02
#define EG_FLAG_PRF 0x07

EG_HD void merlin_header(transcript &t, const char *label, uint32_t label_len, uint32_t msg_len) {
    strobe_begin_op(t, EG_FLAG_META_AD);
    strobe_absorb(t, (const uint8_t *)label, label_len);
    uint8_t l[4] = {(uint8_t)msg_len, (uint8_t)(msg_len >> 8), (uint8_t)(msg_len >> 16), (uint8_t)(msg_len >> 24)};
    strobe_absorb(t, l, 4);          // meta_ad(.., more = true): same operation continues
}

EG_HD void merlin_append_message(transcript &t, const char *label, uint32_t label_len, const uint8_t *m, uint32_t n) {
    merlin_header(t, label, label_len, n);
    strobe_begin_op(t, EG_FLAG_AD);
    strobe_absorb(t, m, n);
}

EG_HD void merlin_append_words(transcript &t, const char *label, uint32_t label_len, const uint32_t *w, uint32_t n_words) {
    merlin_header(t, label, label_len, 4 * n_words);
    strobe_begin_op(t, EG_FLAG_AD);
    strobe_absorb_words(t, w, n_words);
}

EG_HD void merlin_append_u64(transcript &t, const char *label, uint32_t label_len, uint64_t x) {
    uint32_t w[2] = {(uint32_t)x, (uint32_t)(x >> 32)};
    merlin_append_words(t, label, label_len, w, 2);
}

// proofs/mod.rs:54-56 -> ristretto.rs:34-38: 64 challenge bytes, wide-reduced
static EG_HD_NOINLINE void merlin_challenge_scalar(transcript &t, const char *label, uint32_t label_len, sc &out) {
    merlin_header(t, label, label_len, 64);
    strobe_begin_op(t, EG_FLAG_PRF);
    uint32_t w[16];
#pragma unroll 1
    for (int i = 0; i < 16; i++) {
        uint32_t x = 0;
#pragma unroll 1
        for (int k = 0; k < 4; k++) {
            x |= (uint32_t)st_get_byte(t, t.pos) << (8 * k);
            st_clear_byte(t, t.pos);
            if (++t.pos == EG_STROBE_R) strobe_run_f(t);
        }
        w[i] = x;
    }
    sc_from_wide_words(out, w);
}

EG_HD void merlin_new(transcript &t, const char *label, uint32_t label_len) {
    for (int i = 0; i < 25; i++) t.st[i] = 0;
    const uint8_t init[18] = {1, EG_STROBE_R + 2, 1, 0, 1, 96, 'S', 'T', 'R', 'O', 'B', 'E', 'v', '1', '.', '0', '.', '2'};
    for (uint32_t i = 0; i < 18; i++) st_xor_byte(t, i, init[i]);
    keccak_f1600(t.st);
    t.pos = 0; t.pos_begin = 0;
    strobe_begin_op(t, EG_FLAG_META_AD);
    strobe_absorb(t, (const uint8_t *)"Merlin v1.0", 11);
    merlin_append_message(t, "dom-sep", 7, (const uint8_t *)label, label_len);
}

#define EG_LBL(s) (s), (uint32_t)(sizeof(s) - 1)

}  // namespace eg
