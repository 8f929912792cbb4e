// chacha.cuh -- the provers' randomness: caller-supplied 64-byte blocks, or ChaCha20 generated in the kernel.
//
// The reference draws every secret scalar with Ristretto::generate_scalar (src/group/ristretto.rs:28-32): 64 bytes from a
// CryptoRng, reduced mod l.  With rand_chacha's ChaCha20Rng (tests/snapshots.rs:31-34, benches/basics.rs:17) those 64 bytes
// are exactly one ChaCha20 block (RFC 8439 block function, 64-bit block counter in words 12-13, stream id 0 in words
// 14-15), so "draw number k of the stream" is "block k".  The seeded entry points (eg_*_batch_seeded) give item i of a batch
// its own stream of that generator -- key = the caller's 32-byte seed, block counter = counter_base + (i << 20) + k, the
// layout of SURVEY.md 8(d) -- and produce the blocks in the kernel, so that no randomness crosses PCIe (64 B per draw
// otherwise: 2560 B per range proof, 3392 B per quadratic-voting ballot).
#pragma once
#include <stdint.h>
#include "fe.cuh"

namespace eg {

EG_HD uint32_t chacha_rotl(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

#define EG_CHACHA_QR(a, b, c, d)                       \
    a += b; d ^= a; d = chacha_rotl(d, 16);            \
    c += d; b ^= c; b = chacha_rotl(b, 12);            \
    a += b; d ^= a; d = chacha_rotl(d, 8);             \
    c += d; b ^= c; b = chacha_rotl(b, 7);

// one 64-byte keystream block as 16 little-endian words
EG_HD void chacha20_block(uint32_t out[16], const uint32_t key[8], uint64_t counter) {
    uint32_t s[16], x[16];
    s[0] = 0x61707865u; s[1] = 0x3320646eu; s[2] = 0x79622d32u; s[3] = 0x6b206574u;
    for (int i = 0; i < 8; i++) s[4 + i] = key[i];
    s[12] = (uint32_t)counter; s[13] = (uint32_t)(counter >> 32); s[14] = 0; s[15] = 0;
    for (int i = 0; i < 16; i++) x[i] = s[i];
#pragma unroll 1
    for (int r = 0; r < 10; r++) {
        EG_CHACHA_QR(x[0], x[4], x[8], x[12])
        EG_CHACHA_QR(x[1], x[5], x[9], x[13])
        EG_CHACHA_QR(x[2], x[6], x[10], x[14])
        EG_CHACHA_QR(x[3], x[7], x[11], x[15])
        EG_CHACHA_QR(x[0], x[5], x[10], x[15])
        EG_CHACHA_QR(x[1], x[6], x[11], x[12])
        EG_CHACHA_QR(x[2], x[7], x[8], x[13])
        EG_CHACHA_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) out[i] = x[i] + s[i];
}

// Where a prover kernel takes its draws from.  seeded == 0: `wide` holds the blocks (layout owned by the kernel's params);
// seeded == 1: block k of (global) item i is chacha20_block(key, counter_base + ((item0 + i) << 20) + block0 + k).
struct rand_src {
    uint32_t key[8];
    uint64_t counter_base;
    uint64_t item0;             // global index of the chunk's first item
    uint32_t block0;            // first block of this proof inside the item's stream (a proof embedded in a larger object)
    uint8_t seeded;
    uint8_t ct;                 // constant-time fixed-base arithmetic for the secret scalars (eg_ctx_set_prover_mode)
};

#define EG_ITEM_STREAM_SHIFT 20

EG_HD uint64_t rand_counter(const rand_src &R, size_t item, uint32_t block) {
    return R.counter_base + ((R.item0 + (uint64_t)item) << EG_ITEM_STREAM_SHIFT) + R.block0 + block;
}

}  // namespace eg
