// eg_b200.cu -- the C ABI (include/eg_b200.h) and kernel launches of the B200 batch engine.
//
// Host side: context / scratch management, per-election transcript prefixes (computed once with the same
// __host__ __device__ Merlin code the kernels use), slot tables that describe which equation of which
// proof each thread evaluates, chunked execution.  All group / field / hash arithmetic of a batch runs in
// the kernels below; there is no CPU path (eg_ctx_create fails without a device).
#ifdef EG_HOSTSIM
#include "hostsim_cuda.h"      // test-only stand-in for the CUDA runtime (tests/hostsim): malloc/memcpy, no device
#else
#include <cuda_runtime.h>
#endif

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/eg_b200.h"
#include "kernels.cuh"

using namespace eg;

// =================================================================== kernels
//
// Every kernel is a thin __global__ wrapper around a body in kernels.cuh.  launch_*() is the only place a
// kernel is started; under EG_HOSTSIM (tests/hostsim, test harness only) the same bodies run in a host loop.

#ifndef EG_COMMIT_THREADS
#define EG_COMMIT_THREADS 128
#endif
#ifndef EG_COMMIT_MINBLOCKS
#define EG_COMMIT_MINBLOCKS 4
#endif
#define EG_TALLY_THREADS 128
#define EG_TALLY_BLOCKS 148       // per slot: one CTA per SM

#ifndef EG_HOSTSIM

__global__ void __launch_bounds__(256) k_decode(const decode_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n * (size_t)P.n_slots) return;
    decode_body(P, tid % P.n, (int)(tid / P.n));
}

__global__ void __launch_bounds__(256) k_scalars(const scalars_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n) return;
    scalars_body(P, tid);
}

// The hot kernel.  Both 12 KB fixed-base tables are staged in shared memory once per CTA.
__global__ void __launch_bounds__(EG_COMMIT_THREADS, EG_COMMIT_MINBLOCKS) k_commit(const commit_params P) {
    __shared__ __align__(16) uint32_t s_tab[2 * EG_FIXED_TABLE_WORDS];
    for (int k = threadIdx.x; k < EG_FIXED_TABLE_WORDS; k += blockDim.x) {
        s_tab[k] = P.table_g[k];
        s_tab[EG_FIXED_TABLE_WORDS + k] = P.table_k[k];
    }
    __syncthreads();
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n * (size_t)P.n_slots) return;
    commit_body(P, tid % P.n, (int)(tid / P.n), s_tab, s_tab + EG_FIXED_TABLE_WORDS);
}

// v2 ring engine: persistent grid (one CTA slot per resident CTA), each thread walks (item, ring) pairs with a fixed
// 8 KB scratch region for the window tables of its current ring.  Both 48 KB chunked fixed-base tables live in shared
// memory (dynamic, 96 KB per CTA).
// Two launch shapes, chosen per job by the ring sizes (measured on B200, profiles/r1_ab_launch_shapes.txt):
//   * rings of two equations (bool / choice): ONE CTA of 640 threads per SM (20 warps = 5 per scheduler, <= 102 registers):
//     2 x 256 threads 1.784 M ballots/s, 1 x 512 1.823 M, 1 x 640 1.866 M, 1 x 768 1.833 M; warp counts that do not divide
//     by the 4 schedulers (576, 704) lose 8-10 %.  One CTA per SM also leaves 132 KB instead of 36 KB of the unified L1
//     to the per-thread window tables.
//   * longer rings (range proofs, QV ballots): 2 CTAs of 256 threads (128 registers): 643 k range proofs/s vs 600 k with
//     1 x 640 -- the longer equation loop suffers from the extra spills of the 102-register build.
// EG_RING_THREADS / EG_RING_MINBLOCKS is the shape of the long-ring variant and of the prover kernels.
#ifndef EG_RING_THREADS
#define EG_RING_THREADS 256
#endif
#ifndef EG_RING_MINBLOCKS
#define EG_RING_MINBLOCKS 2
#endif
#ifndef EG_RING2_THREADS
#define EG_RING2_THREADS 640
#endif
#ifndef EG_RING2_MINBLOCKS
#define EG_RING2_MINBLOCKS 1
#endif
template <int THREADS, int MINBLOCKS>
__global__ void __launch_bounds__(THREADS, MINBLOCKS) k_ring(const ring_params P) {
    extern __shared__ __align__(16) uint32_t s_rtab[];
    for (int k = threadIdx.x; k < EG_FCHUNK_TABLE_WORDS; k += blockDim.x) {
        s_rtab[k] = P.table_g[k];
        s_rtab[EG_FCHUNK_TABLE_WORDS + k] = P.table_k[k];
    }
    __syncthreads();
    const size_t total = P.n * (size_t)P.n_rings, stride = (size_t)gridDim.x * blockDim.x;
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t *scratch = P.scratch + slot * (2 * EG_VTAB_WORDS);
    for (size_t tid = slot; tid < total; tid += stride)
        ring_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
}

// proving side (encrypt_bool / EncryptedChoice::new): same persistent shape and shared-memory tables as k_ring.
// PHASE 0: ciphertexts + ring construction, 1: common challenge + sum proof (one thread per item), 2: finalize
template <int PHASE>
__global__ void __launch_bounds__(EG_RING_THREADS, EG_RING_MINBLOCKS) k_prove(const prove_params P, uint32_t *scratch_base) {
    extern __shared__ __align__(16) uint32_t s_rtab[];
    for (int k = threadIdx.x; k < EG_FCHUNK_TABLE_WORDS; k += blockDim.x) {
        s_rtab[k] = P.table_g[k];
        s_rtab[EG_FCHUNK_TABLE_WORDS + k] = P.table_k[k];
    }
    __syncthreads();
    const size_t total = PHASE == 1 ? P.n : P.n * (size_t)P.options, stride = (size_t)gridDim.x * blockDim.x;
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t *scratch = scratch_base + slot * (2 * EG_VTAB_WORDS);
    for (size_t tid = slot; tid < total; tid += stride) {
        if (PHASE == 0) prove_ring1_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
        if (PHASE == 1) prove_common_body(P, tid, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
        if (PHASE == 2) prove_ring2_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
    }
}

// RangeProof::new: same persistent shape.  PHASE 0: ciphertexts + Ring::new (n_rings + 1 slots per item), 1: common challenge,
// 2: Ring::finalize
template <int PHASE>
__global__ void __launch_bounds__(EG_RING_THREADS, EG_RING_MINBLOCKS) k_rprove(const rprove_params P, uint32_t *scratch_base) {
    extern __shared__ __align__(16) uint32_t s_rtab[];
    if (PHASE != 1) {
        for (int k = threadIdx.x; k < EG_FCHUNK_TABLE_WORDS; k += blockDim.x) {
            s_rtab[k] = P.table_g[k];
            s_rtab[EG_FCHUNK_TABLE_WORDS + k] = P.table_k[k];
        }
        __syncthreads();
    }
    const size_t slots = PHASE == 0 ? P.n_rings + 1 : (PHASE == 1 ? 1 : P.n_rings);
    const size_t total = P.n * slots, stride = (size_t)gridDim.x * blockDim.x;
    const size_t slot = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t *scratch = scratch_base + slot * (2 * EG_VTAB_WORDS);
    for (size_t tid = slot; tid < total; tid += stride) {
        if (PHASE == 0) rprove_ring1_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
        if (PHASE == 1) rprove_common_body(P, tid);
        if (PHASE == 2) rprove_ring2_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
    }
}

// PublicKey::encrypt / encrypt_zero, one thread per item; persistent grid, chunked fixed-base tables in shared memory
__global__ void __launch_bounds__(EG_RING_THREADS, EG_RING_MINBLOCKS) k_encrypt(const encrypt_params P) {
    extern __shared__ __align__(16) uint32_t s_rtab[];
    for (int k = threadIdx.x; k < EG_FCHUNK_TABLE_WORDS; k += blockDim.x) {
        s_rtab[k] = P.table_g[k];
        s_rtab[EG_FCHUNK_TABLE_WORDS + k] = P.table_k[k];
    }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; tid < P.n; tid += stride)
        encrypt_body(P, tid, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
}

// SumOfSquaresProof::new, one thread per item; persistent grid, chunked fixed-base tables in shared memory
__global__ void __launch_bounds__(EG_RING_THREADS, EG_RING_MINBLOCKS) k_sumsq_prove(const sumsq_prove_params P) {
    extern __shared__ __align__(16) uint32_t s_rtab[];
    for (int k = threadIdx.x; k < EG_FCHUNK_TABLE_WORDS; k += blockDim.x) {
        s_rtab[k] = P.table_g[k];
        s_rtab[EG_FCHUNK_TABLE_WORDS + k] = P.table_k[k];
    }
    __syncthreads();
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x; tid < P.n; tid += stride)
        sumsq_prove_body(P, tid, s_rtab, s_rtab + EG_FCHUNK_TABLE_WORDS);
}

__global__ void __launch_bounds__(128) k_ring_hash(const ring_hash_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n * (size_t)P.n_slots) return;
    ring_hash_body(P, tid % P.n, (int)(tid / P.n));
}

__global__ void __launch_bounds__(128) k_ring_final(const ring_final_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n) return;
    ring_final_body(P, tid);
}

__global__ void __launch_bounds__(128) k_logeq_final(const logeq_final_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n) return;
    logeq_final_body(P, tid);
}

__global__ void __launch_bounds__(128) k_choice_sum(const choice_sum_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= 2 * P.n) return;
    choice_sum_body(P, tid % P.n, (int)(tid / P.n));
}

__global__ void __launch_bounds__(128) k_range_last(const range_last_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= 2 * P.n) return;
    range_last_body(P, tid % P.n, (int)(tid / P.n));
}

__global__ void __launch_bounds__(256) k_verdict(const verdict_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n) return;
    verdict_body(P, tid);
}

__global__ void k_build_table(const uint32_t *enc_words, int use_generator, uint32_t *table, uint32_t *status) {
    build_table_body(blockIdx.x * blockDim.x + threadIdx.x, enc_words, use_generator, table, status);
}

__global__ void k_admissible(const uint64_t *values, int count, uint32_t *adm) {
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < count) admissible_body(tid, values, adm);
}

// --- tally: masked point sums (warp-shuffle tree -> block -> grid) -------------------------------------

__device__ __forceinline__ void shfl_point_down(ge_ext &q, const ge_ext &p, int delta) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
        q.X.v[k] = __shfl_down_sync(0xffffffffu, p.X.v[k], delta);
        q.Y.v[k] = __shfl_down_sync(0xffffffffu, p.Y.v[k], delta);
        q.Z.v[k] = __shfl_down_sync(0xffffffffu, p.Z.v[k], delta);
        q.T.v[k] = __shfl_down_sync(0xffffffffu, p.T.v[k], delta);
    }
}

// partial[(slot * gridDim.x + block) * 32 ..] = sum over this block's items with verdict OK of pts[slot]
__global__ void __launch_bounds__(EG_TALLY_THREADS) k_tally_partial(const uint32_t *pts, size_t n, const uint8_t *verdicts,
                                                                    uint32_t *partial) {
    __shared__ uint32_t s_pt[EG_TALLY_THREADS / 32][32];
    const int slot = blockIdx.y;
    ge_ext acc = ge_identity(), q;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        if (verdicts[i] == 0) {
            planar_load_point(q, pts, n, slot, i);
            ge_add(acc, acc, q);
        }
    }
#pragma unroll 1
    for (int delta = 16; delta >= 1; delta >>= 1) {
        shfl_point_down(q, acc, delta);
        ge_add(acc, acc, q);          // lanes >= delta compute values that are never read
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) point_to_words32(s_pt[warp], acc);
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int wp = 1; wp < EG_TALLY_THREADS / 32; wp++) {
            point_from_words32(q, s_pt[wp]);
            ge_add(acc, acc, q);
        }
        point_to_words32(partial + ((size_t)slot * gridDim.x + blockIdx.x) * 32, acc);
    }
}

// running[slot] (+)= sum of `count` partial points; one warp per slot.  If `encode_out` is set the total is encoded.
__global__ void __launch_bounds__(32) k_tally_final(const uint32_t *partial, int count, uint32_t *running, int accumulate,
                                                    uint8_t *encode_out) {
    const int slot = blockIdx.x, lane = threadIdx.x;
    ge_ext acc = ge_identity(), q;
    for (int i = lane; i < count; i += 32) {
        point_from_words32(q, partial + ((size_t)slot * count + i) * 32);
        ge_add(acc, acc, q);
    }
#pragma unroll 1
    for (int delta = 16; delta >= 1; delta >>= 1) {
        shfl_point_down(q, acc, delta);
        ge_add(acc, acc, q);
    }
    if (lane == 0) {
        uint32_t *r = running + (size_t)slot * 32;
        if (accumulate) { point_from_words32(q, r); ge_add(acc, acc, q); }
        point_to_words32(r, acc);
        if (encode_out) {
            uint32_t w[8];
            ge_encode(w, acc);
            store32_bytes(encode_out + 32 * slot, w);
        }
    }
}

__global__ void __launch_bounds__(128) k_elements_validate(const uint8_t *enc, size_t n, uint8_t *ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) elements_validate_body(i, enc, ok);
}

__global__ void __launch_bounds__(256) k_scalars_validate(const uint8_t *s, size_t n, uint8_t *ok) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scalars_validate_body(i, s, ok);
}

__global__ void __launch_bounds__(256) k_keyset_verdict(const keyset_verdict_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < P.n) keyset_verdict_body(P, tid);
}

__global__ void __launch_bounds__(256) k_unpack(const unpack_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < P.n) unpack_body(P, tid);
}

template <int ENCODE>
__global__ void __launch_bounds__(256) k_b64url(const b64_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n * (size_t)P.groups) return;
    if (ENCODE) b64url_encode_body(P, tid); else b64url_decode_body(P, tid);
}

__global__ void __launch_bounds__(256) k_scalars_from_wide(const uint8_t *wide, size_t n, uint8_t *out) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) scalars_from_wide_body(i, wide, out);
}

__global__ void __launch_bounds__(EG_COMMIT_THREADS) k_double_mul(const uint8_t *a, const uint8_t *A, const uint8_t *b, size_t n,
                                                                  int mode, const uint32_t *table_g, uint8_t *out, uint8_t *okv) {
    __shared__ uint32_t s_tab[EG_FIXED_TABLE_WORDS];
    for (int k = threadIdx.x; k < EG_FIXED_TABLE_WORDS; k += blockDim.x) s_tab[k] = table_g[k];
    __syncthreads();
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) double_mul_body(i, a, A, b, mode, s_tab, out, okv);
}

__global__ void k_ciphertexts_sum(const uint8_t *parts, size_t n_parts, size_t n_cts, uint8_t *out, uint32_t *bad) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < 2 * n_cts) ciphertexts_sum_body(tid, parts, n_parts, n_cts, out, bad);
}

// general multi-scalar equations (share proofs, SumOfSquaresProof, Lagrange recombination)
__global__ void __launch_bounds__(128) k_msm(const msm_params P) {
    __shared__ __align__(16) uint32_t s_tab[3 * EG_FIXED_TABLE_WORDS];      // G, K and (when set) the Pedersen base H
    for (int k = threadIdx.x; k < EG_FIXED_TABLE_WORDS; k += blockDim.x) {
        s_tab[k] = P.table_g[k];
        s_tab[EG_FIXED_TABLE_WORDS + k] = P.table_k[k];
        if (P.table_h) s_tab[2 * EG_FIXED_TABLE_WORDS + k] = P.table_h[k];
    }
    __syncthreads();
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= P.n * (size_t)P.n_slots) return;
    msm_body(P, tid % P.n, (int)(tid / P.n), s_tab, s_tab + EG_FIXED_TABLE_WORDS, s_tab + 2 * EG_FIXED_TABLE_WORDS);
}

__global__ void __launch_bounds__(128) k_sigma_final(const sigma_final_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < P.n) sigma_final_body(P, tid);
}

__global__ void __launch_bounds__(128) k_sumsq_final(const sumsq_final_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < P.n) sumsq_final_body(P, tid);
}

__global__ void __launch_bounds__(128) k_share_final(const share_final_params *Pp) {
    const share_final_params &P = *Pp;
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < P.n * P.n_shares) share_final_body(P, tid % P.n, (int)(tid / P.n));
}

__global__ void __launch_bounds__(256) k_qv_verdict(const qv_verdict_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < P.n) qv_verdict_body(P, tid);
}

__global__ void __launch_bounds__(256) k_share_verdict(const share_verdict_params P, size_t total) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < total) share_verdict_body(P, tid);
}

__global__ void __launch_bounds__(128) k_dlog_build(const dlog_build_params P, size_t threads) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < threads) dlog_build_body(P, tid);
}

__global__ void __launch_bounds__(128) k_dlog_lookup(const dlog_lookup_params P) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid < P.n) dlog_lookup_body(P, tid);
}

// on-device self-test of the tuned field arithmetic against the portable formulation (eg_selftest_field)
__global__ void __launch_bounds__(128) k_selftest_field(size_t n, uint64_t seed, unsigned long long *mismatches) {
    size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= n) return;
    uint64_t x = seed + 0x9e3779b97f4a7c15ULL * (tid + 1);
    fe a, b;
    for (int i = 0; i < 8; i++) {
        x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ULL; x ^= x >> 27; x *= 0x94d049bb133111ebULL; x ^= x >> 31;
        a.v[i] = (uint32_t)x; b.v[i] = (uint32_t)(x >> 32);
        x += 0x9e3779b97f4a7c15ULL;
    }
    // edge patterns on the first threads: all ones, p, p-1, 2^256-38, small values
    const int e = (int)(tid % 64);
    if (tid < 4096) {
        if (e == 0) for (int i = 0; i < 8; i++) a.v[i] = 0xffffffffu;
        if (e == 1) for (int i = 0; i < 8; i++) b.v[i] = 0xffffffffu;
        if (e == 2) { for (int i = 0; i < 8; i++) a.v[i] = b.v[i] = 0xffffffffu; }
        if (e == 3) { for (int i = 1; i < 7; i++) a.v[i] = 0xffffffffu; a.v[0] = 0xffffffedu; a.v[7] = 0x7fffffffu; }
        if (e == 4) { for (int i = 1; i < 8; i++) a.v[i] = 0xffffffffu; a.v[0] = 0xffffffdau; }
        if (e == 5) { a = fe_zero(); }
        if (e == 6) { a = fe_one(); for (int i = 0; i < 8; i++) b.v[i] = 0xffffffffu; }
        if (e == 7) { for (int i = 0; i < 8; i++) a.v[i] = (i & 1) ? 0xffffffffu : 0u; }
    }
    fe m1, m2, s1, s2, d;
    fe_mul(m1, a, b); fe_mul_portable(m2, a, b);
    fe_sq(s1, a); fe_sq_portable(s2, a);
    unsigned bad = 0;
    fe_sub(d, m1, m2); if (!fe_iszero(d)) bad++;
    fe_sub(d, s1, s2); if (!fe_iszero(d)) bad++;
    fe_mul_portable(m2, a, a); fe_sub(d, s1, m2); if (!fe_iszero(d)) bad++;
    // (a + b) - b == a and a * 1 == a through the carry-chain add/sub
    fe t; fe_add(t, a, b); fe_sub(t, t, b); fe_sub(d, t, a); if (!fe_iszero(d)) bad++;
    if (bad) atomicAdd(mismatches, (unsigned long long)bad);
}

#endif  // !EG_HOSTSIM

// =================================================================== context

struct dev_buf {
    void *p = nullptr;
    size_t cap = 0;
};

struct eg_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool has_receiver = false, has_blinding_base = false;
    uint8_t key[32];
    uint8_t blinding_base[32];
    uint32_t *d_table_g = nullptr, *d_table_k = nullptr, *d_table_h = nullptr, *d_status = nullptr;
    std::string err;
    uint64_t launches = 0, commit_launches = 0, commit_tasks = 0;
    float timings[5] = {0, 0, 0, 0, 0};
    cudaEvent_t ev[8];
    std::vector<cudaEvent_t> commit_ev;     // pairs (start, stop) around every k_commit / k_ring launch of the current call
    std::vector<uint8_t> commit_ev_kind;    // per pair: 0 = k_commit, 1 = k_ring
    size_t commit_ev_used = 0;
    uint64_t call_commit_tasks = 0, call_commit_launches = 0;
    uint64_t kind_tasks[2] = {0, 0}, kind_launches[2] = {0, 0};   // per kind, current call
    float kind_ms[2] = {0, 0};
    // grow-only scratch
    cudaStream_t copy_stream = nullptr;     // host -> device prefetch of the next chunk
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr};
    dev_buf in2[3];
    dev_buf pts, enc, commit, chal, flags, res[3], in[4], verdicts, partial, running, adm, misc, slots, consts, res_big;
    size_t adm_used = 0;      // cached points in `adm` (32 words each); entries 0,1 = the [O, G] pair
    std::map<std::string, std::vector<uint64_t>> adm_cache_key;
    size_t chunk_items = 0;   // 0 = default
    int ring_mode = 2;        // 2: k_ring (one thread per ring, chunked tables); 1: k_commit / k_ring_hash launches per equation
    int ring_grid[2] = {0, 0};   // resident CTAs of the two k_ring shapes (queried once)
    int prove_grid[3] = {0, 0, 0};
    int rprove_grid[3] = {0, 0, 0};
    int sumsq_prove_grid = 0, encrypt_grid = 0;
    dev_buf ring_scratch;
};

struct eg_dlog_table {
    eg_ctx *ctx;
    uint64_t lo, hi;
    size_t cap;
    uint32_t *d_keys;     // cap * 8 words
    uint64_t *d_vals;     // cap
};

static eg_status fail(eg_ctx *ctx, eg_status st, const char *what, cudaError_t ce = cudaSuccess) {
    if (ctx) {
        ctx->err = what;
        if (ce != cudaSuccess) { ctx->err += ": "; ctx->err += cudaGetErrorString(ce); }
    }
    return st;
}

#define CU(call)                                                            \
    do {                                                                    \
        cudaError_t ce_ = (call);                                           \
        if (ce_ != cudaSuccess) return fail(ctx, EG_ERR_CUDA, #call, ce_);  \
    } while (0)

static eg_status ensure(eg_ctx *ctx, dev_buf &b, size_t bytes) {
    if (b.cap >= bytes) return EG_SUCCESS;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
    size_t want = bytes + bytes / 8;
    cudaError_t ce = cudaMalloc(&b.p, want);
    if (ce != cudaSuccess) { cudaGetLastError(); return fail(ctx, EG_ERR_OUT_OF_MEMORY, "cudaMalloc scratch", ce); }
    b.cap = want;
    return EG_SUCCESS;
}

#define TRY(expr)                              \
    do {                                       \
        eg_status st_ = (expr);                \
        if (st_ != EG_SUCCESS) return st_;     \
    } while (0)

static inline unsigned grid_for(size_t threads, unsigned block) { return (unsigned)((threads + block - 1) / block); }

// ------------------------------------------------------------------- launchers (the only kernel start sites)

#ifdef EG_HOSTSIM
#define EG_FOR_HOST(total, stmt) for (size_t tid = 0; tid < (size_t)(total); tid++) { stmt; }
#endif

static void launch_decode(eg_ctx *ctx, const decode_params &P) {
    size_t total = P.n * (size_t)P.n_slots;
#ifdef EG_HOSTSIM
    EG_FOR_HOST(total, decode_body(P, tid % P.n, (int)(tid / P.n)))
#else
    k_decode<<<grid_for(total, 256), 256, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_scalars(eg_ctx *ctx, const scalars_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, scalars_body(P, tid))
#else
    k_scalars<<<grid_for(P.n, 256), 256, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_commit(eg_ctx *ctx, const commit_params &P) {
    size_t total = P.n * (size_t)P.n_slots;
    if (ctx->commit_ev_used + 2 > ctx->commit_ev.size()) {
        cudaEvent_t a = nullptr, b = nullptr;
        cudaEventCreate(&a); cudaEventCreate(&b);
        ctx->commit_ev.push_back(a); ctx->commit_ev.push_back(b);
    }
    cudaEvent_t e_start = ctx->commit_ev[ctx->commit_ev_used], e_stop = ctx->commit_ev[ctx->commit_ev_used + 1];
    if (ctx->commit_ev_kind.size() < ctx->commit_ev.size() / 2) ctx->commit_ev_kind.resize(ctx->commit_ev.size() / 2);
    ctx->commit_ev_kind[ctx->commit_ev_used / 2] = 0;
    ctx->kind_tasks[0] += total; ctx->kind_launches[0]++;
    ctx->commit_ev_used += 2;
    cudaEventRecord(e_start, ctx->stream);
#ifdef EG_HOSTSIM
    EG_FOR_HOST(total, commit_body(P, tid % P.n, (int)(tid / P.n), P.table_g, P.table_k))
#else
    k_commit<<<grid_for(total, EG_COMMIT_THREADS), EG_COMMIT_THREADS, 0, ctx->stream>>>(P);
#endif
    cudaEventRecord(e_stop, ctx->stream);
    ctx->launches++;
    ctx->commit_launches++;
    ctx->commit_tasks += total;
    ctx->call_commit_tasks += total;
    ctx->call_commit_launches++;
}

// One launch evaluates every equation of every ring of the chunk.  Accounted like k_commit: tasks = equation sides.
static eg_status launch_ring(eg_ctx *ctx, ring_params &P) {
    size_t sides = 0;
    for (uint32_t r = 0; r < P.n_rings; r++) sides += 2 * (size_t)P.sizes[r];
    sides *= P.n;
    if (ctx->commit_ev_used + 2 > ctx->commit_ev.size()) {
        cudaEvent_t a = nullptr, b = nullptr;
        cudaEventCreate(&a); cudaEventCreate(&b);
        ctx->commit_ev.push_back(a); ctx->commit_ev.push_back(b);
    }
    cudaEvent_t e_start = ctx->commit_ev[ctx->commit_ev_used], e_stop = ctx->commit_ev[ctx->commit_ev_used + 1];
    if (ctx->commit_ev_kind.size() < ctx->commit_ev.size() / 2) ctx->commit_ev_kind.resize(ctx->commit_ev.size() / 2);
    ctx->commit_ev_kind[ctx->commit_ev_used / 2] = 1;
    ctx->kind_tasks[1] += sides; ctx->kind_launches[1]++;
    ctx->commit_ev_used += 2;
    const size_t total = P.n * (size_t)P.n_rings;
#ifdef EG_HOSTSIM
    TRY(ensure(ctx, ctx->ring_scratch, 2 * EG_VTAB_WORDS * 4));
    P.scratch = (uint32_t *)ctx->ring_scratch.p;
    cudaEventRecord(e_start, ctx->stream);
    EG_FOR_HOST(total, ring_body(P, tid % P.n, (uint32_t)(tid / P.n), P.scratch, P.table_g, P.table_k))
#else
    const size_t smem = 2 * EG_FCHUNK_TABLE_WORDS * 4;
    bool short_rings = true;
    for (uint32_t r = 0; r < P.n_rings; r++) short_rings = short_rings && P.sizes[r] <= 2;
    const int shape = short_rings ? 1 : 0;
    const int threads = shape ? EG_RING2_THREADS : EG_RING_THREADS;
    if (ctx->ring_grid[shape] == 0) {
        int per_sm = 0, sms = 0;
        if (shape) {
            CU(cudaFuncSetAttribute(k_ring<EG_RING2_THREADS, EG_RING2_MINBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ring<EG_RING2_THREADS, EG_RING2_MINBLOCKS>, threads, smem));
        } else {
            CU(cudaFuncSetAttribute(k_ring<EG_RING_THREADS, EG_RING_MINBLOCKS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_ring<EG_RING_THREADS, EG_RING_MINBLOCKS>, threads, smem));
        }
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (per_sm < 1) return fail(ctx, EG_ERR_CUDA, "k_ring does not fit on an SM");
        ctx->ring_grid[shape] = per_sm * sms;
    }
    const size_t resident = (size_t)ctx->ring_grid[shape];
    const unsigned grid = (unsigned)std::min<size_t>(resident, (total + threads - 1) / threads);
    TRY(ensure(ctx, ctx->ring_scratch, resident * threads * 2 * EG_VTAB_WORDS * 4));
    P.scratch = (uint32_t *)ctx->ring_scratch.p;
    cudaEventRecord(e_start, ctx->stream);
    if (shape) k_ring<EG_RING2_THREADS, EG_RING2_MINBLOCKS><<<grid, threads, smem, ctx->stream>>>(P);
    else k_ring<EG_RING_THREADS, EG_RING_MINBLOCKS><<<grid, threads, smem, ctx->stream>>>(P);
#endif
    cudaEventRecord(e_stop, ctx->stream);
    ctx->launches++;
    ctx->commit_launches++;
    ctx->commit_tasks += sides;
    ctx->call_commit_tasks += sides;
    ctx->call_commit_launches++;
    return EG_SUCCESS;
}

#ifndef EG_HOSTSIM
template <int PHASE>
static eg_status launch_prove_phase(eg_ctx *ctx, const prove_params &P, int &grid_cache) {
    const size_t smem = 2 * EG_FCHUNK_TABLE_WORDS * 4;
    if (grid_cache == 0) {
        CU(cudaFuncSetAttribute(k_prove<PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_prove<PHASE>, EG_RING_THREADS, smem));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (per_sm < 1) return fail(ctx, EG_ERR_CUDA, "k_prove does not fit on an SM");
        grid_cache = per_sm * sms;
    }
    const size_t total = PHASE == 1 ? P.n : P.n * (size_t)P.options;
    const unsigned grid = (unsigned)std::min<size_t>((size_t)grid_cache, (total + EG_RING_THREADS - 1) / EG_RING_THREADS);
    TRY(ensure(ctx, ctx->ring_scratch, (size_t)grid_cache * EG_RING_THREADS * 2 * EG_VTAB_WORDS * 4));
    k_prove<PHASE><<<grid, EG_RING_THREADS, smem, ctx->stream>>>(P, (uint32_t *)ctx->ring_scratch.p);
    ctx->launches++;
    return EG_SUCCESS;
}
#endif

static eg_status launch_prove(eg_ctx *ctx, const prove_params &P) {
#ifdef EG_HOSTSIM
    TRY(ensure(ctx, ctx->ring_scratch, 2 * EG_VTAB_WORDS * 4));
    uint32_t *scratch = (uint32_t *)ctx->ring_scratch.p;
    EG_FOR_HOST(P.n * (size_t)P.options, prove_ring1_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, P.table_g, P.table_k))
    EG_FOR_HOST(P.n, prove_common_body(P, tid, P.table_g, P.table_k))
    EG_FOR_HOST(P.n * (size_t)P.options, prove_ring2_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, P.table_g, P.table_k))
    ctx->launches += 3;
#else
    TRY(launch_prove_phase<0>(ctx, P, ctx->prove_grid[0]));
    TRY(launch_prove_phase<1>(ctx, P, ctx->prove_grid[1]));
    TRY(launch_prove_phase<2>(ctx, P, ctx->prove_grid[2]));
#endif
    return EG_SUCCESS;
}

#ifndef EG_HOSTSIM
template <int PHASE>
static eg_status launch_rprove_phase(eg_ctx *ctx, const rprove_params &P, int &grid_cache) {
    const size_t smem = PHASE == 1 ? 0 : 2 * EG_FCHUNK_TABLE_WORDS * 4;
    if (grid_cache == 0) {
        CU(cudaFuncSetAttribute(k_rprove<PHASE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_rprove<PHASE>, EG_RING_THREADS, smem));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (per_sm < 1) return fail(ctx, EG_ERR_CUDA, "k_rprove does not fit on an SM");
        grid_cache = per_sm * sms;
    }
    const size_t total = P.n * (size_t)(PHASE == 0 ? P.n_rings + 1 : (PHASE == 1 ? 1 : P.n_rings));
    const unsigned grid = (unsigned)std::min<size_t>((size_t)grid_cache, (total + EG_RING_THREADS - 1) / EG_RING_THREADS);
    TRY(ensure(ctx, ctx->ring_scratch, (size_t)grid_cache * EG_RING_THREADS * 2 * EG_VTAB_WORDS * 4));
    k_rprove<PHASE><<<grid, EG_RING_THREADS, smem, ctx->stream>>>(P, (uint32_t *)ctx->ring_scratch.p);
    ctx->launches++;
    return EG_SUCCESS;
}
#endif

static eg_status launch_rprove(eg_ctx *ctx, const rprove_params &P) {
#ifdef EG_HOSTSIM
    TRY(ensure(ctx, ctx->ring_scratch, 2 * EG_VTAB_WORDS * 4));
    uint32_t *scratch = (uint32_t *)ctx->ring_scratch.p;
    EG_FOR_HOST(P.n * (size_t)(P.n_rings + 1), rprove_ring1_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, P.table_g, P.table_k))
    EG_FOR_HOST(P.n, rprove_common_body(P, tid))
    EG_FOR_HOST(P.n * (size_t)P.n_rings, rprove_ring2_body(P, tid % P.n, (uint32_t)(tid / P.n), scratch, P.table_g, P.table_k))
    ctx->launches += 3;
#else
    TRY(launch_rprove_phase<0>(ctx, P, ctx->rprove_grid[0]));
    TRY(launch_rprove_phase<1>(ctx, P, ctx->rprove_grid[1]));
    TRY(launch_rprove_phase<2>(ctx, P, ctx->rprove_grid[2]));
#endif
    return EG_SUCCESS;
}

static eg_status launch_encrypt(eg_ctx *ctx, const encrypt_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, encrypt_body(P, tid, P.table_g, P.table_k))
#else
    const size_t smem = 2 * EG_FCHUNK_TABLE_WORDS * 4;
    if (ctx->encrypt_grid == 0) {
        CU(cudaFuncSetAttribute(k_encrypt, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_encrypt, EG_RING_THREADS, smem));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (per_sm < 1) return fail(ctx, EG_ERR_CUDA, "k_encrypt does not fit on an SM");
        ctx->encrypt_grid = per_sm * sms;
    }
    const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->encrypt_grid, (P.n + EG_RING_THREADS - 1) / EG_RING_THREADS);
    k_encrypt<<<grid, EG_RING_THREADS, smem, ctx->stream>>>(P);
#endif
    ctx->launches++;
    return EG_SUCCESS;
}

static eg_status launch_sumsq_prove(eg_ctx *ctx, const sumsq_prove_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, sumsq_prove_body(P, tid, P.table_g, P.table_k))
#else
    const size_t smem = 2 * EG_FCHUNK_TABLE_WORDS * 4;
    if (ctx->sumsq_prove_grid == 0) {
        CU(cudaFuncSetAttribute(k_sumsq_prove, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 0, sms = 0;
        CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sumsq_prove, EG_RING_THREADS, smem));
        CU(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
        if (per_sm < 1) return fail(ctx, EG_ERR_CUDA, "k_sumsq_prove does not fit on an SM");
        ctx->sumsq_prove_grid = per_sm * sms;
    }
    const unsigned grid = (unsigned)std::min<size_t>((size_t)ctx->sumsq_prove_grid, (P.n + EG_RING_THREADS - 1) / EG_RING_THREADS);
    k_sumsq_prove<<<grid, EG_RING_THREADS, smem, ctx->stream>>>(P);
#endif
    ctx->launches++;
    return EG_SUCCESS;
}

static void launch_ring_hash(eg_ctx *ctx, const ring_hash_params &P) {
    size_t total = P.n * (size_t)P.n_slots;
#ifdef EG_HOSTSIM
    EG_FOR_HOST(total, ring_hash_body(P, tid % P.n, (int)(tid / P.n)))
#else
    k_ring_hash<<<grid_for(total, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_ring_final(eg_ctx *ctx, const ring_final_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, ring_final_body(P, tid))
#else
    k_ring_final<<<grid_for(P.n, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_logeq_final(eg_ctx *ctx, const logeq_final_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, logeq_final_body(P, tid))
#else
    k_logeq_final<<<grid_for(P.n, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_choice_sum(eg_ctx *ctx, const choice_sum_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(2 * P.n, choice_sum_body(P, tid % P.n, (int)(tid / P.n)))
#else
    k_choice_sum<<<grid_for(2 * P.n, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_range_last(eg_ctx *ctx, const range_last_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(2 * P.n, range_last_body(P, tid % P.n, (int)(tid / P.n)))
#else
    k_range_last<<<grid_for(2 * P.n, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_verdict(eg_ctx *ctx, const verdict_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, verdict_body(P, tid))
#else
    k_verdict<<<grid_for(P.n, 256), 256, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_build_table(eg_ctx *ctx, const uint32_t *enc_words, int use_generator, uint32_t *table, uint32_t *status) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(EG_VCHUNKS * EG_FIXED_TABLE_ENTRIES, build_table_body((int)tid, enc_words, use_generator, table, status))
#else
    k_build_table<<<EG_VCHUNKS, EG_FIXED_TABLE_ENTRIES, 0, ctx->stream>>>(enc_words, use_generator, table, status);
#endif
    ctx->launches++;
}

static void launch_admissible(eg_ctx *ctx, const uint64_t *values, int count, uint32_t *adm) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(count, admissible_body((int)tid, values, adm))
#else
    k_admissible<<<grid_for(count, 64), 64, 0, ctx->stream>>>(values, count, adm);
#endif
    ctx->launches++;
}

// per-slot masked sums of planar points into `partial` (blocks per slot returned)
static int launch_tally_partial(eg_ctx *ctx, const uint32_t *pts, size_t n, int n_slots, const uint8_t *verdicts, uint32_t *partial) {
    const int blocks = (int)std::max<size_t>(1, std::min<size_t>(EG_TALLY_BLOCKS, (n + EG_TALLY_THREADS - 1) / EG_TALLY_THREADS));
#ifdef EG_HOSTSIM
    for (int slot = 0; slot < n_slots; slot++)
        for (int b = 0; b < blocks; b++) {
            ge_ext acc = ge_identity(), q;
            for (size_t i = (size_t)b; i < n; i += (size_t)blocks)
                if (verdicts[i] == 0) { planar_load_point(q, pts, n, slot, i); ge_add(acc, acc, q); }
            point_to_words32(partial + ((size_t)slot * blocks + b) * 32, acc);
        }
#else
    dim3 grid(blocks, n_slots);
    k_tally_partial<<<grid, EG_TALLY_THREADS, 0, ctx->stream>>>(pts, n, verdicts, partial);
#endif
    ctx->launches++;
    return blocks;
}

static void launch_tally_final(eg_ctx *ctx, const uint32_t *partial, int count, int n_slots, uint32_t *running, int accumulate,
                               uint8_t *encode_out) {
#ifdef EG_HOSTSIM
    for (int slot = 0; slot < n_slots; slot++) {
        ge_ext acc = ge_identity(), q;
        for (int i = 0; i < count; i++) { point_from_words32(q, partial + ((size_t)slot * count + i) * 32); ge_add(acc, acc, q); }
        uint32_t *r = running + (size_t)slot * 32;
        if (accumulate) { point_from_words32(q, r); ge_add(acc, acc, q); }
        point_to_words32(r, acc);
        if (encode_out) { uint32_t w[8]; ge_encode(w, acc); store32_bytes(encode_out + 32 * slot, w); }
    }
#else
    k_tally_final<<<n_slots, 32, 0, ctx->stream>>>(partial, count, running, accumulate, encode_out);
#endif
    ctx->launches++;
}

static void launch_elements_validate(eg_ctx *ctx, const uint8_t *enc, size_t n, uint8_t *ok) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(n, elements_validate_body(tid, enc, ok))
#else
    k_elements_validate<<<grid_for(n, 128), 128, 0, ctx->stream>>>(enc, n, ok);
#endif
    ctx->launches++;
}

static void launch_scalars_validate(eg_ctx *ctx, const uint8_t *sc_in, size_t n, uint8_t *ok) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(n, scalars_validate_body(tid, sc_in, ok))
#else
    k_scalars_validate<<<grid_for(n, 256), 256, 0, ctx->stream>>>(sc_in, n, ok);
#endif
    ctx->launches++;
}

static void launch_keyset_verdict(eg_ctx *ctx, const keyset_verdict_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, keyset_verdict_body(P, tid))
#else
    k_keyset_verdict<<<grid_for(P.n, 256), 256, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_unpack(eg_ctx *ctx, const unpack_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, unpack_body(P, tid))
#else
    k_unpack<<<grid_for(P.n, 256), 256, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_b64url(eg_ctx *ctx, const b64_params &P, bool encode) {
    const size_t total = P.n * (size_t)P.groups;
#ifdef EG_HOSTSIM
    if (encode) { EG_FOR_HOST(total, b64url_encode_body(P, tid)) } else { EG_FOR_HOST(total, b64url_decode_body(P, tid)) }
#else
    if (encode) k_b64url<1><<<grid_for(total, 256), 256, 0, ctx->stream>>>(P);
    else k_b64url<0><<<grid_for(total, 256), 256, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_scalars_from_wide(eg_ctx *ctx, const uint8_t *wide, size_t n, uint8_t *out) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(n, scalars_from_wide_body(tid, wide, out))
#else
    k_scalars_from_wide<<<grid_for(n, 256), 256, 0, ctx->stream>>>(wide, n, out);
#endif
    ctx->launches++;
}

static void launch_double_mul(eg_ctx *ctx, const uint8_t *a, const uint8_t *A, const uint8_t *b, size_t n, int mode, uint8_t *out, uint8_t *okv) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(n, double_mul_body(tid, a, A, b, mode, ctx->d_table_g, out, okv))
#else
    k_double_mul<<<grid_for(n, EG_COMMIT_THREADS), EG_COMMIT_THREADS, 0, ctx->stream>>>(a, A, b, n, mode, ctx->d_table_g, out, okv);
#endif
    ctx->launches++;
}

static void launch_ciphertexts_sum(eg_ctx *ctx, const uint8_t *parts, size_t n_parts, size_t n_cts, uint8_t *out, uint32_t *bad) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(2 * n_cts, ciphertexts_sum_body(tid, parts, n_parts, n_cts, out, bad))
#else
    k_ciphertexts_sum<<<grid_for(2 * n_cts, 64), 64, 0, ctx->stream>>>(parts, n_parts, n_cts, out, bad);
#endif
    ctx->launches++;
}

static void launch_msm(eg_ctx *ctx, const msm_params &P) {
    size_t total = P.n * (size_t)P.n_slots;
#ifdef EG_HOSTSIM
    EG_FOR_HOST(total, msm_body(P, tid % P.n, (int)(tid / P.n), P.table_g, P.table_k, P.table_h))
#else
    k_msm<<<grid_for(total, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_sigma_final(eg_ctx *ctx, const sigma_final_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, sigma_final_body(P, tid))
#else
    k_sigma_final<<<grid_for(P.n, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_sumsq_final(eg_ctx *ctx, const sumsq_final_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, sumsq_final_body(P, tid))
#else
    k_sumsq_final<<<grid_for(P.n, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

// `P` lives in host memory; the device copy `d_P` is what the kernel reads (the struct exceeds the 4 KB launch limit)
static void launch_share_final(eg_ctx *ctx, const share_final_params &P, const share_final_params *d_P) {
    size_t total = P.n * (size_t)P.n_shares;
#ifdef EG_HOSTSIM
    (void)d_P;
    EG_FOR_HOST(total, share_final_body(P, tid % P.n, (int)(tid / P.n)))
#else
    k_share_final<<<grid_for(total, 128), 128, 0, ctx->stream>>>(d_P);
#endif
    ctx->launches++;
}

static void launch_dlog_build(eg_ctx *ctx, const dlog_build_params &P, size_t threads) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(threads, dlog_build_body(P, tid))
#else
    k_dlog_build<<<grid_for(threads, 128), 128, 0, ctx->stream>>>(P, threads);
#endif
    ctx->launches++;
}

static void launch_dlog_lookup(eg_ctx *ctx, const dlog_lookup_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, dlog_lookup_body(P, tid))
#else
    k_dlog_lookup<<<grid_for(P.n, 128), 128, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

extern "C" const char *eg_version(void) { return "eg_b200 0.1.0 sm_100a"; }

extern "C" const char *eg_last_error(const eg_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

extern "C" uint64_t eg_kernel_launch_count(const eg_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" void *eg_ctx_stream(const eg_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

extern "C" eg_status eg_last_commit_stats(const eg_ctx *ctx, uint64_t *launches, uint64_t *tasks, float *ms) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    if (launches) *launches = ctx->call_commit_launches;
    if (tasks) *tasks = ctx->call_commit_tasks;
    if (ms) *ms = ctx->timings[2];
    return EG_SUCCESS;
}

// kind 0: k_commit launches of the last call, kind 1: k_ring launches (tasks = equation sides)
extern "C" eg_status eg_last_kernel_stats(const eg_ctx *ctx, int kind, uint64_t *launches, uint64_t *tasks, float *ms) {
    if (!ctx || kind < 0 || kind > 1) return EG_ERR_INVALID_ARG;
    if (launches) *launches = ctx->kind_launches[kind];
    if (tasks) *tasks = ctx->kind_tasks[kind];
    if (ms) *ms = ctx->kind_ms[kind];
    return EG_SUCCESS;
}

extern "C" eg_status eg_last_timings(const eg_ctx *ctx, float out_ms[5]) {
    if (!ctx || !out_ms) return EG_ERR_INVALID_ARG;
    for (int i = 0; i < 5; i++) out_ms[i] = ctx->timings[i];
    return EG_SUCCESS;
}


extern "C" eg_status eg_selftest_field(eg_ctx *ctx, size_t n, uint64_t seed, uint64_t *mismatches) {
    if (!ctx || !mismatches) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    *mismatches = 0;
#ifndef EG_HOSTSIM
    unsigned long long *d = (unsigned long long *)(ctx->d_status + 64);
    CU(cudaMemsetAsync(d, 0, 8, ctx->stream));
    k_selftest_field<<<grid_for(n, 128), 128, 0, ctx->stream>>>(n, seed, d);
    ctx->launches++;
    unsigned long long h = 0;
    CU(cudaMemcpyAsync(&h, d, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    *mismatches = h;
#else
    (void)n; (void)seed;
#endif
    return EG_SUCCESS;
}

extern "C" eg_status eg_ctx_set_chunk_items(eg_ctx *ctx, size_t items) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    ctx->chunk_items = items;
    return EG_SUCCESS;
}

extern "C" eg_status eg_ctx_set_ring_mode(eg_ctx *ctx, int mode) {
    if (!ctx || (mode != 1 && mode != 2)) return EG_ERR_INVALID_ARG;
    ctx->ring_mode = mode;
    return EG_SUCCESS;
}

extern "C" eg_status eg_ctx_create(int device_id, eg_ctx **out) {
    if (!out) return EG_ERR_INVALID_ARG;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0 || device_id < 0 || device_id >= count) {
        cudaGetLastError();
        return EG_ERR_NO_DEVICE;
    }
    eg_ctx *ctx = new (std::nothrow) eg_ctx();
    if (!ctx) return EG_ERR_OUT_OF_MEMORY;
    ctx->device = device_id;
    auto bail = [&](eg_status st) { eg_ctx_destroy(ctx); return st; };
    if (cudaSetDevice(device_id) != cudaSuccess) return bail(EG_ERR_NO_DEVICE);
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(EG_ERR_CUDA);
    for (auto &e : ctx->ev) if (cudaEventCreate(&e) != cudaSuccess) return bail(EG_ERR_CUDA);
    if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) return bail(EG_ERR_CUDA);
    for (auto &e : ctx->ev_h2d) if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return bail(EG_ERR_CUDA);
    if (cudaMalloc(&ctx->d_table_g, EG_FCHUNK_TABLE_WORDS * 4) != cudaSuccess) return bail(EG_ERR_OUT_OF_MEMORY);
    if (cudaMalloc(&ctx->d_table_k, EG_FCHUNK_TABLE_WORDS * 4) != cudaSuccess) return bail(EG_ERR_OUT_OF_MEMORY);
    if (cudaMalloc(&ctx->d_status, 1024) != cudaSuccess) return bail(EG_ERR_OUT_OF_MEMORY);
    launch_build_table(ctx, nullptr, 1, ctx->d_table_g, ctx->d_status);
    if (cudaStreamSynchronize(ctx->stream) != cudaSuccess || cudaGetLastError() != cudaSuccess) return bail(EG_ERR_CUDA);
    *out = ctx;
    return EG_SUCCESS;
}

extern "C" void eg_ctx_destroy(eg_ctx *ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    dev_buf *bufs[] = {&ctx->pts, &ctx->enc, &ctx->commit, &ctx->chal, &ctx->flags, &ctx->res[0], &ctx->res[1], &ctx->res[2],
                       &ctx->in[0], &ctx->in[1], &ctx->in[2], &ctx->in[3], &ctx->verdicts, &ctx->partial, &ctx->running,
                       &ctx->adm, &ctx->misc, &ctx->slots, &ctx->consts, &ctx->res_big, &ctx->ring_scratch,
                       &ctx->in2[0], &ctx->in2[1], &ctx->in2[2]};
    for (dev_buf *b : bufs) if (b->p) cudaFree(b->p);
    if (ctx->d_table_g) cudaFree(ctx->d_table_g);
    if (ctx->d_table_k) cudaFree(ctx->d_table_k);
    if (ctx->d_table_h) cudaFree(ctx->d_table_h);
    if (ctx->d_status) cudaFree(ctx->d_status);
    for (auto &e : ctx->ev) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->commit_ev) if (e) cudaEventDestroy(e);
    for (auto &e : ctx->ev_h2d) if (e) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" eg_status eg_ctx_set_receiver(eg_ctx *ctx, const uint8_t key[32]) {
    if (!ctx || !key) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    uint32_t *d_key = ctx->d_status + 8;
    CU(cudaMemcpyAsync(d_key, key, 32, cudaMemcpyHostToDevice, ctx->stream));
    launch_build_table(ctx, d_key, 0, ctx->d_table_k, ctx->d_status);
    uint32_t status = 0;
    CU(cudaMemcpyAsync(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    if (status == 1) { ctx->has_receiver = false; return fail(ctx, EG_ERR_INVALID_ELEMENT, "receiver key does not represent a group element"); }
    if (status == 2) { ctx->has_receiver = false; return fail(ctx, EG_ERR_IDENTITY_KEY, "receiver key is the group identity"); }
    memcpy(ctx->key, key, 32);
    ctx->has_receiver = true;
    return EG_SUCCESS;
}

// Pedersen blinding base H of CommitmentEquivalenceProof (commitment.rs:140-145: `commitment_blinding_base`), e.g. the
// Bulletproofs base of tests/snapshots.rs:253-257.  Gets the same chunked fixed-base table as G and K.
extern "C" eg_status eg_ctx_set_blinding_base(eg_ctx *ctx, const uint8_t base[32]) {
    if (!ctx || !base) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (!ctx->d_table_h) {
        cudaError_t ce = cudaMalloc(&ctx->d_table_h, EG_FCHUNK_TABLE_WORDS * 4);
        if (ce != cudaSuccess) { cudaGetLastError(); ctx->d_table_h = nullptr; return fail(ctx, EG_ERR_OUT_OF_MEMORY, "cudaMalloc H table", ce); }
    }
    uint32_t *d_key = ctx->d_status + 16;
    CU(cudaMemcpyAsync(d_key, base, 32, cudaMemcpyHostToDevice, ctx->stream));
    launch_build_table(ctx, d_key, 0, ctx->d_table_h, ctx->d_status);
    uint32_t status = 0;
    CU(cudaMemcpyAsync(&status, ctx->d_status, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    ctx->has_blinding_base = false;
    if (status == 1) return fail(ctx, EG_ERR_INVALID_ELEMENT, "blinding base does not represent a group element");
    if (status == 2) return fail(ctx, EG_ERR_IDENTITY_KEY, "blinding base is the group identity");
    memcpy(ctx->blinding_base, base, 32);
    ctx->has_blinding_base = true;
    return EG_SUCCESS;
}

// =================================================================== ring-proof engine

// Host description of one batch of RingProof::verify calls sharing a shape (ring.rs:302-374).
struct ring_job {
    uint32_t n_rings = 0;
    uint32_t sizes[EG_MAX_RINGS];
    uint32_t ct_p_index[EG_MAX_RINGS];      // ring r ciphertext: R at ct_p_index[r], B at +1
    uint32_t ct_enc_index[EG_MAX_RINGS];    // enc(R) at ct_enc_index[r], enc(B) at +1
    int32_t adm_index[EG_MAX_RINGS];        // admissible value j of ring r (j >= 1) at adm_index[r] + j; -1: [O, G] pair
    uint64_t adm_step[EG_MAX_RINGS];        // admissible value j of ring r = [j * adm_step[r]] G
    uint8_t proof_buf = 0;                  // input buffer of the ring proof (e0 | responses)
    uint32_t proof_offset = 0;              // byte offset of the proof inside the item
    uint32_t commit_index0 = 0;             // planar commitments: ring r -> commit_index0 + 2r (+1)
    uint32_t chal_index0 = 0;               // planar challenges: ring r -> chal_index0 + r
    transcript prefix;                      // after initialize_transcript (ring.rs:290-293)
    uint32_t *d_result = nullptr;
};

static void host_ring_initialize(transcript &t, const uint8_t key[32]) {      // ring.rs:290-293
    merlin_append_message(t, EG_LBL("dom-sep"), (const uint8_t *)"multi_ring_enc", 14);
    merlin_append_message(t, EG_LBL("K"), key, 32);
}

// Launch every stage of the ring engine.  `extra` are additional commit slots evaluated together with stage 0
// (e.g. the sum proof of an EncryptedChoice), so that the first launch is as wide as possible.
static eg_status run_ring_job(eg_ctx *ctx, const ring_job &job, const in_bufs &in, size_t n, const commit_slot *extra, int n_extra,
                              const uint32_t *d_adm) {
    uint32_t max_size = 0, starts[EG_MAX_RINGS], start = 0;
    for (uint32_t r = 0; r < job.n_rings; r++) { starts[r] = start; start += job.sizes[r]; max_size = std::max(max_size, job.sizes[r]); }
    if (ctx->ring_mode == 2) {
        if (n_extra > 0) {          // e.g. the sum proof of an EncryptedChoice: single-use points, plain chain
            commit_params cp;
            memset(&cp, 0, sizeof cp);
            cp.in = in; cp.n = n;
            cp.pts = (const uint32_t *)ctx->pts.p; cp.chal = (const uint32_t *)ctx->chal.p; cp.commit = (uint32_t *)ctx->commit.p;
            cp.adm = d_adm; cp.table_g = ctx->d_table_g; cp.table_k = ctx->d_table_k;
            for (int k = 0; k < n_extra; k++) cp.slots[k] = extra[k];
            cp.n_slots = n_extra;
            launch_commit(ctx, cp);
        }
        ring_params rp;
        memset(&rp, 0, sizeof rp);
        rp.in = in; rp.n = n; rp.n_rings = job.n_rings;
        for (uint32_t r = 0; r < job.n_rings; r++) {
            rp.sizes[r] = (uint16_t)job.sizes[r]; rp.starts[r] = (uint16_t)starts[r];
            rp.ct_p_index[r] = job.ct_p_index[r]; rp.ct_enc_index[r] = job.ct_enc_index[r]; rp.adm_step[r] = job.adm_step[r];
        }
        rp.proof_buf = job.proof_buf; rp.proof_offset = job.proof_offset; rp.commit_index0 = job.commit_index0;
        rp.prefix = job.prefix;
        rp.pts = (const uint32_t *)ctx->pts.p; rp.enc = (const uint32_t *)ctx->enc.p; rp.commit = (uint32_t *)ctx->commit.p;
        rp.table_g = ctx->d_table_g; rp.table_k = ctx->d_table_k;
        TRY(launch_ring(ctx, rp));
        max_size = 0;               // skip the per-equation launches below
    }
    for (uint32_t j = 0; j < max_size; j++) {
        // ---- commitments of equation j for every ring that has one (ring.rs:342-350)
        uint32_t r0 = 0;
        bool first_launch = true;
        while (r0 < job.n_rings || (first_launch && j == 0 && n_extra > 0)) {
            commit_params cp;
            memset(&cp, 0, sizeof cp);
            cp.in = in; cp.n = n;
            cp.pts = (const uint32_t *)ctx->pts.p; cp.chal = (const uint32_t *)ctx->chal.p; cp.commit = (uint32_t *)ctx->commit.p;
            cp.adm = d_adm; cp.table_g = ctx->d_table_g; cp.table_k = ctx->d_table_k;
            int ns = 0;
            if (first_launch && j == 0)
                for (int k = 0; k < n_extra; k++) cp.slots[ns++] = extra[k];
            first_launch = false;
            for (; r0 < job.n_rings && ns + 2 <= EG_MAX_SLOTS; r0++) {
                if (job.sizes[r0] <= j) continue;
                for (int side = 0; side < 2; side++) {
                    commit_slot s;
                    memset(&s, 0, sizeof s);
                    s.p_index = job.ct_p_index[r0] + side;
                    s.adm_index = -1;
                    if (side == 1 && j >= 1) s.adm_index = job.adm_index[r0] + (int32_t)j;
                    s.base = (uint8_t)side;
                    s.e_planar = j > 0;
                    s.e_buf = job.proof_buf; s.s_buf = job.proof_buf;
                    s.e_offset = j > 0 ? job.chal_index0 + r0 : job.proof_offset;
                    s.s_offset = job.proof_offset + 32 * (1 + starts[r0] + j);
                    s.out_index = job.commit_index0 + 2 * r0 + side;
                    cp.slots[ns++] = s;
                }
            }
            if (ns == 0) break;
            cp.n_slots = ns;
            launch_commit(ctx, cp);
        }
        // ---- next challenge for rings with a further equation (ring.rs:354-360)
        r0 = 0;
        while (r0 < job.n_rings) {
            ring_hash_params hp;
            memset(&hp, 0, sizeof hp);
            hp.n = n; hp.prefix = job.prefix;
            hp.enc = (const uint32_t *)ctx->enc.p; hp.commit = (const uint32_t *)ctx->commit.p; hp.chal = (uint32_t *)ctx->chal.p;
            int ns = 0;
            for (; r0 < job.n_rings && ns < EG_MAX_SLOTS; r0++) {
                if (job.sizes[r0] <= j + 1) continue;
                ring_hash_slot s;
                s.ring_index = r0; s.eq_index = j;
                s.enc_index = job.ct_enc_index[r0];
                s.commit_index = job.commit_index0 + 2 * r0;
                s.chal_index = job.chal_index0 + r0;
                hp.slots[ns++] = s;
            }
            if (ns == 0) break;
            hp.n_slots = ns;
            launch_ring_hash(ctx, hp);
        }
    }
    ring_final_params fp;
    memset(&fp, 0, sizeof fp);
    fp.in = in; fp.n = n; fp.n_rings = job.n_rings; fp.commit_index0 = job.commit_index0;
    fp.proof_buf = job.proof_buf; fp.cc_offset = job.proof_offset; fp.prefix = job.prefix;
    fp.commit = (const uint32_t *)ctx->commit.p; fp.result = job.d_result;
    launch_ring_final(ctx, fp);
    return EG_SUCCESS;
}

// The [O, G] admissible pair used by bool / choice rings: index 1 holds cached(G)
static eg_status ensure_bool_adm(eg_ctx *ctx) {
    if (ctx->adm.p && ctx->adm_cache_key.count("bool")) return EG_SUCCESS;
    TRY(ensure(ctx, ctx->adm, 4096 * 128));
    uint64_t vals[2] = {0, 1};
    uint64_t *d_vals = (uint64_t *)(ctx->d_status + 32);   // byte 128 of the 1 KB status block
    CU(cudaMemcpyAsync(d_vals, vals, sizeof vals, cudaMemcpyHostToDevice, ctx->stream));
    launch_admissible(ctx, d_vals, 2, (uint32_t *)ctx->adm.p);
    CU(cudaStreamSynchronize(ctx->stream));
    ctx->adm_cache_key.clear();
    ctx->adm_cache_key["bool"] = {0, 1};
    return EG_SUCCESS;
}

static bool valid_label(const char *label) { return label && strlen(label) > 0 && strlen(label) < 256; }

static size_t default_chunk(const eg_ctx *ctx) { return ctx->chunk_items ? ctx->chunk_items : ((size_t)1 << 18); }

static eg_status begin_call(eg_ctx *ctx) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    if (!ctx->has_receiver) return fail(ctx, EG_ERR_NO_RECEIVER, "eg_ctx_set_receiver has not been called");
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    ctx->commit_ev_used = 0;
    ctx->call_commit_tasks = 0;
    ctx->call_commit_launches = 0; ctx->kind_tasks[0] = ctx->kind_tasks[1] = 0; ctx->kind_launches[0] = ctx->kind_launches[1] = 0;
    return EG_SUCCESS;
}

static eg_status finish_call(eg_ctx *ctx) {
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    float commit_ms = 0;
    ctx->kind_ms[0] = ctx->kind_ms[1] = 0;
    for (size_t k = 0; k + 1 < ctx->commit_ev_used; k += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, ctx->commit_ev[k], ctx->commit_ev[k + 1]) == cudaSuccess) {
            commit_ms += ms;
            ctx->kind_ms[ctx->commit_ev_kind[k / 2] ? 1 : 0] += ms;
        }
    }
    ctx->timings[2] = commit_ms;
    return EG_SUCCESS;
}

// =================================================================== verify_bool

// one chunk, device pointers
static eg_status verify_bool_chunk(eg_ctx *ctx, size_t n, const uint8_t *d_cts, const uint8_t *d_proofs, uint8_t *d_verdicts) {
    TRY(ensure(ctx, ctx->pts, n * 2 * 128));
    TRY(ensure(ctx, ctx->enc, n * 2 * 32));
    TRY(ensure(ctx, ctx->commit, n * 2 * 32));
    TRY(ensure(ctx, ctx->chal, n * 1 * 32));
    TRY(ensure(ctx, ctx->flags, n * 4));
    TRY(ensure(ctx, ctx->res[0], n * 4));
    TRY(ensure_bool_adm(ctx));
    CU(cudaMemsetAsync(ctx->flags.p, 0, n * 4, ctx->stream));
    in_bufs in;
    memset(&in, 0, sizeof in);
    in.buf[0] = d_cts; in.stride[0] = 64;
    in.buf[1] = d_proofs; in.stride[1] = 96;

    decode_params dp;
    memset(&dp, 0, sizeof dp);
    dp.in = in; dp.n = n; dp.n_slots = 2;
    for (int k = 0; k < 2; k++) { dp.slots[k].buf = 0; dp.slots[k].want_enc = 1; dp.slots[k].enc_index = (uint16_t)k; dp.slots[k].offset = 32 * k; dp.slots[k].p_index = k; }
    dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = (uint32_t *)ctx->flags.p;
    launch_decode(ctx, dp);

    scalars_params sp;
    memset(&sp, 0, sizeof sp);
    sp.in = in; sp.n = n; sp.n_slots = 1; sp.slots[0].buf = 1; sp.slots[0].offset = 0; sp.slots[0].count = 3;
    sp.flags = (uint32_t *)ctx->flags.p;
    launch_scalars(ctx, sp);

    ring_job job;
    job.n_rings = 1; job.sizes[0] = 2; job.ct_p_index[0] = 0; job.ct_enc_index[0] = 0; job.adm_index[0] = 0; job.adm_step[0] = 1;
    job.proof_buf = 1; job.proof_offset = 0; job.commit_index0 = 0; job.chal_index0 = 0;
    merlin_new(job.prefix, EG_LBL("bool_encryption"));            // keys/impls.rs:111
    host_ring_initialize(job.prefix, ctx->key);
    job.d_result = (uint32_t *)ctx->res[0].p;
    TRY(run_ring_job(ctx, job, in, n, nullptr, 0, (const uint32_t *)ctx->adm.p));

    verdict_params vp;
    memset(&vp, 0, sizeof vp);
    vp.n = n; vp.flags = (const uint32_t *)ctx->flags.p; vp.n_checks = 1;
    vp.check[0] = (const uint32_t *)ctx->res[0].p; vp.check_stride[0] = 1; vp.code[0] = EG_V_CHALLENGE_MISMATCH;
    vp.verdicts = d_verdicts;
    launch_verdict(ctx, vp);
    return EG_SUCCESS;
}

extern "C" eg_status eg_verify_bool_batch_dev(eg_ctx *ctx, size_t n, const uint8_t *d_cts, const uint8_t *d_proofs, uint8_t *d_verdicts) {
    TRY(begin_call(ctx));
    if (n == 0) return EG_SUCCESS;
    if (!d_cts || !d_proofs || !d_verdicts) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t chunk = default_chunk(ctx);
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        TRY(verify_bool_chunk(ctx, m, d_cts + 64 * off, d_proofs + 96 * off, d_verdicts + off));
    }
    return finish_call(ctx);
}

extern "C" eg_status eg_verify_bool_batch(eg_ctx *ctx, size_t n, const uint8_t *cts, const uint8_t *proofs, uint8_t *verdicts) {
    TRY(begin_call(ctx));
    if (n == 0) return EG_SUCCESS;
    if (!cts || !proofs || !verdicts) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t chunk = default_chunk(ctx);
    TRY(ensure(ctx, ctx->in[0], std::min(chunk, n) * 64));
    TRY(ensure(ctx, ctx->in[1], std::min(chunk, n) * 96));
    TRY(ensure(ctx, ctx->verdicts, std::min(chunk, n)));
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, cts + 64 * off, m * 64, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, proofs + 96 * off, m * 96, cudaMemcpyHostToDevice, ctx->stream));
        TRY(verify_bool_chunk(ctx, m, (const uint8_t *)ctx->in[0].p, (const uint8_t *)ctx->in[1].p, (uint8_t *)ctx->verdicts.p));
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, m, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== verify_zero

static eg_status verify_zero_chunk(eg_ctx *ctx, size_t n, const uint8_t *d_cts, const uint8_t *d_proofs, uint8_t *d_verdicts) {
    TRY(ensure(ctx, ctx->pts, n * 2 * 128));
    TRY(ensure(ctx, ctx->enc, n * 2 * 32));
    TRY(ensure(ctx, ctx->commit, n * 2 * 32));
    TRY(ensure(ctx, ctx->flags, n * 4));
    TRY(ensure(ctx, ctx->res[0], n * 4));
    CU(cudaMemsetAsync(ctx->flags.p, 0, n * 4, ctx->stream));
    in_bufs in;
    memset(&in, 0, sizeof in);
    in.buf[0] = d_cts; in.stride[0] = 64;
    in.buf[1] = d_proofs; in.stride[1] = 64;
    decode_params dp;
    memset(&dp, 0, sizeof dp);
    dp.in = in; dp.n = n; dp.n_slots = 2;
    for (int k = 0; k < 2; k++) { dp.slots[k].buf = 0; dp.slots[k].want_enc = 1; dp.slots[k].enc_index = (uint16_t)k; dp.slots[k].offset = 32 * k; dp.slots[k].p_index = k; }
    dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = (uint32_t *)ctx->flags.p;
    launch_decode(ctx, dp);
    scalars_params sp;
    memset(&sp, 0, sizeof sp);
    sp.in = in; sp.n = n; sp.n_slots = 1; sp.slots[0].buf = 1; sp.slots[0].offset = 0; sp.slots[0].count = 2;
    sp.flags = (uint32_t *)ctx->flags.p;
    launch_scalars(ctx, sp);
    // log_equality.rs:160-164: [x]G = [-c]R + [s]G ; [x]K = [-c]B + [s]K
    commit_params cp;
    memset(&cp, 0, sizeof cp);
    cp.in = in; cp.n = n; cp.n_slots = 2;
    cp.pts = (const uint32_t *)ctx->pts.p; cp.commit = (uint32_t *)ctx->commit.p;
    cp.table_g = ctx->d_table_g; cp.table_k = ctx->d_table_k;
    for (int side = 0; side < 2; side++) {
        commit_slot &s = cp.slots[side];
        s.p_index = side; s.adm_index = -1; s.base = (uint8_t)side; s.e_planar = 0; s.e_buf = 1; s.s_buf = 1;
        s.e_offset = 0; s.s_offset = 32; s.out_index = side;
    }
    launch_commit(ctx, cp);
    logeq_final_params lp;
    memset(&lp, 0, sizeof lp);
    lp.in = in; lp.n = n; lp.pow_enc_index = 0; lp.commit_index = 0; lp.proof_buf = 1; lp.c_offset = 0;
    merlin_new(lp.prefix, EG_LBL("zero_encryption"));              // keys/impls.rs:67
    merlin_append_message(lp.prefix, EG_LBL("dom-sep"), (const uint8_t *)"log_eq", 6);
    merlin_append_message(lp.prefix, EG_LBL("K"), ctx->key, 32);
    lp.enc = (const uint32_t *)ctx->enc.p; lp.commit = (const uint32_t *)ctx->commit.p; lp.result = (uint32_t *)ctx->res[0].p;
    launch_logeq_final(ctx, lp);
    verdict_params vp;
    memset(&vp, 0, sizeof vp);
    vp.n = n; vp.flags = (const uint32_t *)ctx->flags.p; vp.n_checks = 1;
    vp.check[0] = (const uint32_t *)ctx->res[0].p; vp.check_stride[0] = 1; vp.code[0] = EG_V_CHALLENGE_MISMATCH;
    vp.verdicts = d_verdicts;
    launch_verdict(ctx, vp);
    return EG_SUCCESS;
}

extern "C" eg_status eg_verify_zero_batch(eg_ctx *ctx, size_t n, const uint8_t *cts, const uint8_t *proofs, uint8_t *verdicts) {
    TRY(begin_call(ctx));
    if (n == 0) return EG_SUCCESS;
    if (!cts || !proofs || !verdicts) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t chunk = default_chunk(ctx);
    TRY(ensure(ctx, ctx->in[0], std::min(chunk, n) * 64));
    TRY(ensure(ctx, ctx->in[1], std::min(chunk, n) * 64));
    TRY(ensure(ctx, ctx->verdicts, std::min(chunk, n)));
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, cts + 64 * off, m * 64, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, proofs + 64 * off, m * 64, cudaMemcpyHostToDevice, ctx->stream));
        TRY(verify_zero_chunk(ctx, m, (const uint8_t *)ctx->in[0].p, (const uint8_t *)ctx->in[1].p, (uint8_t *)ctx->verdicts.p));
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, m, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== EncryptedChoice::verify + tally

// planar indexes for a choice chunk with m options:
//   points   : 0..2m-1 ciphertexts (R_k at 2k, B_k at 2k+1), 2m: sum R, 2m+1: sum B - G
//   enc      : same numbering
//   commit   : ring r -> 2r, 2r+1 ; sum proof -> 2m, 2m+1
//   chal     : ring r -> r
static eg_status verify_choice_chunk(eg_ctx *ctx, size_t n, uint32_t m, int single, const uint8_t *d_choices, const uint8_t *d_rings,
                                     const uint8_t *d_sums, uint8_t *d_verdicts, bool want_tally, bool first_chunk) {
    const uint32_t np = 2 * m + 2;
    TRY(ensure(ctx, ctx->pts, n * np * 128));
    TRY(ensure(ctx, ctx->enc, n * np * 32));
    TRY(ensure(ctx, ctx->commit, n * np * 32));
    TRY(ensure(ctx, ctx->chal, n * m * 32));
    TRY(ensure(ctx, ctx->flags, n * 4));
    TRY(ensure(ctx, ctx->res[0], n * 4));
    TRY(ensure(ctx, ctx->res[1], n * 4));
    TRY(ensure_bool_adm(ctx));
    CU(cudaMemsetAsync(ctx->flags.p, 0, n * 4, ctx->stream));
    in_bufs in;
    memset(&in, 0, sizeof in);
    in.buf[0] = d_choices; in.stride[0] = 64 * m;
    in.buf[1] = d_rings; in.stride[1] = 32 * (1 + 2 * m);
    in.buf[2] = d_sums; in.stride[2] = 64;

    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    // ---- decode all 2m input elements (serde / from_bytes stage of the reference)
    for (uint32_t k0 = 0; k0 < 2 * m; k0 += EG_MAX_SLOTS) {
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in = in; dp.n = n;
        int ns = 0;
        for (uint32_t k = k0; k < 2 * m && ns < EG_MAX_SLOTS; k++, ns++) {
            dp.slots[ns].buf = 0; dp.slots[ns].want_enc = 1; dp.slots[ns].enc_index = (uint16_t)k;
            dp.slots[ns].offset = 32 * k; dp.slots[ns].p_index = k;
        }
        dp.n_slots = ns;
        dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = (uint32_t *)ctx->flags.p;
        launch_decode(ctx, dp);
    }
    scalars_params sp;
    memset(&sp, 0, sizeof sp);
    sp.in = in; sp.n = n; sp.n_slots = single ? 2 : 1;
    sp.slots[0].buf = 1; sp.slots[0].offset = 0; sp.slots[0].count = 1 + 2 * m;
    sp.slots[1].buf = 2; sp.slots[1].offset = 0; sp.slots[1].count = 2;
    sp.flags = (uint32_t *)ctx->flags.p;
    launch_scalars(ctx, sp);

    commit_slot extra[2];
    int n_extra = 0;
    if (single) {
        // ---- choice.rs:363 + :83-86: sum ciphertext, powers (sum R, sum B - G)
        choice_sum_params cs;
        memset(&cs, 0, sizeof cs);
        cs.n = n; cs.options = m; cs.out_p_index = 2 * m; cs.out_enc_index = 2 * m;
        cs.pts = (uint32_t *)ctx->pts.p; cs.enc = (uint32_t *)ctx->enc.p;
        launch_choice_sum(ctx, cs);
        for (int side = 0; side < 2; side++) {
            commit_slot &s = extra[side];
            memset(&s, 0, sizeof s);
            s.p_index = 2 * m + side; s.adm_index = -1; s.base = (uint8_t)side; s.e_planar = 0; s.e_buf = 2; s.s_buf = 2;
            s.e_offset = 0; s.s_offset = 32; s.out_index = 2 * m + side;
        }
        n_extra = 2;
    }
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));

    ring_job job;
    job.n_rings = m;
    for (uint32_t r = 0; r < m; r++) { job.sizes[r] = 2; job.ct_p_index[r] = 2 * r; job.ct_enc_index[r] = 2 * r; job.adm_index[r] = 0; job.adm_step[r] = 1; }
    job.proof_buf = 1; job.proof_offset = 0; job.commit_index0 = 0; job.chal_index0 = 0;
    merlin_new(job.prefix, EG_LBL("encrypted_choice_ranges"));       // choice.rs:376
    host_ring_initialize(job.prefix, ctx->key);
    job.d_result = (uint32_t *)ctx->res[1].p;
    TRY(run_ring_job(ctx, job, in, n, extra, n_extra, (const uint32_t *)ctx->adm.p));

    if (single) {
        logeq_final_params lp;
        memset(&lp, 0, sizeof lp);
        lp.in = in; lp.n = n; lp.pow_enc_index = 2 * m; lp.commit_index = 2 * m; lp.proof_buf = 2; lp.c_offset = 0;
        merlin_new(lp.prefix, EG_LBL("choice_encryption_sum"));      // choice.rs:91
        merlin_append_message(lp.prefix, EG_LBL("dom-sep"), (const uint8_t *)"log_eq", 6);
        merlin_append_message(lp.prefix, EG_LBL("K"), ctx->key, 32);
        lp.enc = (const uint32_t *)ctx->enc.p; lp.commit = (const uint32_t *)ctx->commit.p; lp.result = (uint32_t *)ctx->res[0].p;
        launch_logeq_final(ctx, lp);
    }
    verdict_params vp;
    memset(&vp, 0, sizeof vp);
    vp.n = n; vp.flags = (const uint32_t *)ctx->flags.p;
    int nc = 0;
    if (single) { vp.check[nc] = (const uint32_t *)ctx->res[0].p; vp.check_stride[nc] = 1; vp.code[nc] = EG_V_CHOICE_SUM; nc++; }
    vp.check[nc] = (const uint32_t *)ctx->res[1].p; vp.check_stride[nc] = 1; vp.code[nc] = EG_V_CHOICE_RANGE; nc++;
    vp.n_checks = nc;
    vp.verdicts = d_verdicts;
    launch_verdict(ctx, vp);
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));

    if (want_tally) {
        // ---- examples/voting.rs:200-203: totals += verified choices
        TRY(ensure(ctx, ctx->partial, (size_t)2 * m * EG_TALLY_BLOCKS * 128));
        TRY(ensure(ctx, ctx->running, (size_t)2 * m * 128));
        const int blocks = launch_tally_partial(ctx, (const uint32_t *)ctx->pts.p, n, (int)(2 * m), d_verdicts, (uint32_t *)ctx->partial.p);
        launch_tally_final(ctx, (const uint32_t *)ctx->partial.p, blocks, (int)(2 * m), (uint32_t *)ctx->running.p, first_chunk ? 0 : 1, nullptr);
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    return EG_SUCCESS;
}

static eg_status tally_emit(eg_ctx *ctx, uint32_t m, uint8_t *d_tally_out) {
    // encode the running totals: reuse k_tally_final with zero new partials
    launch_tally_final(ctx, (const uint32_t *)ctx->partial.p, 0, (int)(2 * m), (uint32_t *)ctx->running.p, 1, d_tally_out);
    return EG_SUCCESS;
}

static void collect_timings(eg_ctx *ctx, float acc[5]) {
    float a = 0, b = 0, c = 0;
    if (cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]) == cudaSuccess) acc[0] += a;
    if (cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]) == cudaSuccess) acc[1] += b;
    if (cudaEventElapsedTime(&c, ctx->ev[2], ctx->ev[3]) == cudaSuccess) acc[3] += c;
    acc[4] += a + b + c;
}

extern "C" eg_status eg_verify_choice_batch_dev(eg_ctx *ctx, size_t n, uint32_t options, int single, const uint8_t *d_choices,
                                                const uint8_t *d_rings, const uint8_t *d_sums, uint8_t *d_verdicts, uint8_t *d_tally) {
    TRY(begin_call(ctx));
    if (options == 0 || options > EG_MAX_RINGS) return fail(ctx, EG_ERR_INVALID_ARG, "options must be in 1..64");
    if (n && (!d_choices || !d_rings || !d_verdicts || (single && !d_sums))) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    float acc[5] = {0, 0, 0, 0, 0};
    const size_t chunk = default_chunk(ctx);
    TRY(ensure(ctx, ctx->partial, (size_t)2 * options * EG_TALLY_BLOCKS * 128));
    TRY(ensure(ctx, ctx->running, (size_t)2 * options * 128));
    bool first = true;
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        TRY(verify_choice_chunk(ctx, m, options, single, d_choices + 64 * options * off, d_rings + 32 * (1 + 2 * options) * off,
                                d_sums ? d_sums + 64 * off : nullptr, d_verdicts + off, d_tally != nullptr, first));
        first = false;
        CU(cudaStreamSynchronize(ctx->stream));
        collect_timings(ctx, acc);
    }
    if (d_tally) {
        if (first) {   // n == 0: the empty sum is the identity ciphertext in every option
            launch_tally_final(ctx, (const uint32_t *)ctx->partial.p, 0, (int)(2 * options), (uint32_t *)ctx->running.p, 0, d_tally);
        } else {
            TRY(tally_emit(ctx, options, d_tally));
        }
    }
    { float keep = ctx->timings[2]; memcpy(ctx->timings, acc, sizeof acc); ctx->timings[2] = keep; }
    return finish_call(ctx);
}

extern "C" eg_status eg_verify_choice_batch(eg_ctx *ctx, size_t n, uint32_t options, int single, const uint8_t *choices,
                                            const uint8_t *rings, const uint8_t *sums, uint8_t *verdicts, uint8_t *tally) {
    TRY(begin_call(ctx));
    if (options == 0 || options > EG_MAX_RINGS) return fail(ctx, EG_ERR_INVALID_ARG, "options must be in 1..64");
    if (n && (!choices || !rings || !verdicts || (single && !sums))) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    float acc[5] = {0, 0, 0, 0, 0};
    const size_t chunk = default_chunk(ctx), cm = std::min(chunk, std::max<size_t>(n, 1));
    const size_t ring_stride = 32 * (1 + 2 * (size_t)options);
    TRY(ensure(ctx, ctx->in[0], cm * 64 * options));
    TRY(ensure(ctx, ctx->in[1], cm * ring_stride));
    TRY(ensure(ctx, ctx->in[2], cm * 64));
    TRY(ensure(ctx, ctx->verdicts, cm));
    TRY(ensure(ctx, ctx->misc, 64 * (size_t)options));
    TRY(ensure(ctx, ctx->partial, (size_t)2 * options * EG_TALLY_BLOCKS * 128));
    TRY(ensure(ctx, ctx->running, (size_t)2 * options * 128));
    // Double-buffered input staging: while chunk k is verified on the compute stream, chunk k + 1 is copied host -> device
    // on the copy stream into the other buffer (its previous reader, chunk k - 1, finished before the end-of-chunk
    // synchronisation of the previous iteration).  With pinned host memory the copies disappear behind the kernels.
    dev_buf *bufs[2][3] = {{&ctx->in[0], &ctx->in[1], &ctx->in[2]}, {&ctx->in2[0], &ctx->in2[1], &ctx->in2[2]}};
    const bool two = n > chunk;
    if (two) {
        TRY(ensure(ctx, ctx->in2[0], cm * 64 * options));
        TRY(ensure(ctx, ctx->in2[1], cm * ring_stride));
        TRY(ensure(ctx, ctx->in2[2], cm * 64));
    }
    auto prefetch = [&](size_t off, int b) -> eg_status {
        const size_t m = std::min(chunk, n - off);
        cudaStream_t cs = two ? ctx->copy_stream : ctx->stream;
        CU(cudaMemcpyAsync(bufs[b][0]->p, choices + 64 * options * off, m * 64 * options, cudaMemcpyHostToDevice, cs));
        CU(cudaMemcpyAsync(bufs[b][1]->p, rings + ring_stride * off, m * ring_stride, cudaMemcpyHostToDevice, cs));
        if (single) CU(cudaMemcpyAsync(bufs[b][2]->p, sums + 64 * off, m * 64, cudaMemcpyHostToDevice, cs));
        if (two) CU(cudaEventRecord(ctx->ev_h2d[b], cs));
        return EG_SUCCESS;
    };
    bool first = true;
    int b = 0;
    if (n) TRY(prefetch(0, 0));
    for (size_t off = 0; off < n; off += chunk, b ^= 1) {
        size_t m = std::min(chunk, n - off);
        if (two) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
        TRY(verify_choice_chunk(ctx, m, options, single, (const uint8_t *)bufs[b][0]->p, (const uint8_t *)bufs[b][1]->p,
                                single ? (const uint8_t *)bufs[b][2]->p : nullptr, (uint8_t *)ctx->verdicts.p, tally != nullptr, first));
        first = false;
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, m, cudaMemcpyDeviceToHost, ctx->stream));
        if (off + chunk < n) TRY(prefetch(off + chunk, b ^ 1));
        CU(cudaStreamSynchronize(ctx->stream));
        collect_timings(ctx, acc);
    }
    if (tally) {
        if (first) {
            launch_tally_final(ctx, (const uint32_t *)ctx->partial.p, 0, (int)(2 * options), (uint32_t *)ctx->running.p, 0, (uint8_t *)ctx->misc.p);
        } else {
            TRY(tally_emit(ctx, options, (uint8_t *)ctx->misc.p));
        }
        CU(cudaMemcpyAsync(tally, ctx->misc.p, 64 * (size_t)options, cudaMemcpyDeviceToHost, ctx->stream));
    }
    { float keep = ctx->timings[2]; memcpy(ctx->timings, acc, sizeof acc); ctx->timings[2] = keep; }
    return finish_call(ctx);
}

// =================================================================== encrypt_bool / EncryptedChoice::new

static eg_status prove_chunk(eg_ctx *ctx, size_t n, uint32_t m, int single, const char *label, uint32_t label_len, const uint8_t *d_values,
                             const uint8_t *d_wide, uint8_t *d_cts, uint8_t *d_ring, uint8_t *d_sum) {
    TRY(ensure(ctx, ctx->pts, n * 2 * m * 128));
    TRY(ensure(ctx, ctx->enc, n * 2 * m * 32));
    TRY(ensure(ctx, ctx->commit, n * 2 * m * 32));
    TRY(ensure(ctx, ctx->res_big, n * 2 * m * 32));
    TRY(ensure(ctx, ctx->chal, n * 32));
    prove_params P;
    memset(&P, 0, sizeof P);
    P.n = n; P.options = m; P.draws = 3 * m + (single ? 1 : 0); P.single = single ? 1 : 0;
    P.values = d_values; P.wide = d_wide; P.cts = d_cts; P.ring = d_ring; P.sum = d_sum;
    merlin_new(P.ring_prefix, label, label_len);
    host_ring_initialize(P.ring_prefix, ctx->key);
    merlin_new(P.sum_prefix, EG_LBL("choice_encryption_sum"));            // choice.rs:72
    merlin_append_message(P.sum_prefix, EG_LBL("dom-sep"), (const uint8_t *)"log_eq", 6);
    merlin_append_message(P.sum_prefix, EG_LBL("K"), ctx->key, 32);
    P.pts = (uint32_t *)ctx->pts.p; P.enc = (uint32_t *)ctx->enc.p; P.sec = (uint32_t *)ctx->res_big.p;
    P.commit = (uint32_t *)ctx->commit.p; P.chal = (uint32_t *)ctx->chal.p;
    P.table_g = ctx->d_table_g; P.table_k = ctx->d_table_k;
    return launch_prove(ctx, P);
}

static eg_status prove_batch(eg_ctx *ctx, size_t n, uint32_t m, int single, const char *label, uint32_t label_len, const uint8_t *values,
                             const uint8_t *wide, uint8_t *cts, uint8_t *ring, uint8_t *sum) {
    TRY(begin_call(ctx));
    if (m == 0 || m > EG_MAX_RINGS) return fail(ctx, EG_ERR_INVALID_ARG, "options must be in 1..64");
    if (n == 0) return EG_SUCCESS;
    if (!values || !wide || !cts || !ring || (single && !sum)) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t draws = 3 * (size_t)m + (single ? 1 : 0), ring_stride = 32 * (1 + 2 * (size_t)m);
    const size_t chunk = std::max<size_t>(1024, default_chunk(ctx) * 5 / m), cm = std::min(chunk, n);
    // device staging: values | wide in in[0], in[1]; outputs in in[2], in[3], misc
    TRY(ensure(ctx, ctx->in[0], cm * m));
    TRY(ensure(ctx, ctx->in[1], cm * draws * 64));
    TRY(ensure(ctx, ctx->in[2], cm * m * 64));
    TRY(ensure(ctx, ctx->in[3], cm * ring_stride));
    TRY(ensure(ctx, ctx->misc, std::max<size_t>(cm * 64, 4096)));
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, values + off * m, k * m, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, wide + off * draws * 64, k * draws * 64, cudaMemcpyHostToDevice, ctx->stream));
        TRY(prove_chunk(ctx, k, m, single, label, label_len, (const uint8_t *)ctx->in[0].p, (const uint8_t *)ctx->in[1].p,
                        (uint8_t *)ctx->in[2].p, (uint8_t *)ctx->in[3].p, (uint8_t *)ctx->misc.p));
        CU(cudaMemcpyAsync(cts + off * m * 64, ctx->in[2].p, k * m * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ring + off * ring_stride, ctx->in[3].p, k * ring_stride, cudaMemcpyDeviceToHost, ctx->stream));
        if (single) CU(cudaMemcpyAsync(sum + off * 64, ctx->misc.p, k * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

extern "C" eg_status eg_encrypt_bool_batch(eg_ctx *ctx, size_t n, const uint8_t *values, const uint8_t *wide_rand, uint8_t *cts,
                                           uint8_t *proofs) {
    return prove_batch(ctx, n, 1, 0, EG_LBL("bool_encryption"), values, wide_rand, cts, proofs, nullptr);      // keys/impls.rs:82
}

extern "C" eg_status eg_encrypt_choice_batch(eg_ctx *ctx, size_t n, uint32_t options, int single, const uint8_t *values,
                                             const uint8_t *wide_rand, uint8_t *choices, uint8_t *ring_proofs, uint8_t *sum_proofs) {
    return prove_batch(ctx, n, options, single, EG_LBL("encrypted_choice_ranges"), values, wide_rand, choices, ring_proofs,  // choice.rs:323
                       sum_proofs);
}

// =================================================================== group-level helpers

extern "C" eg_status eg_elements_validate(eg_ctx *ctx, size_t n, const uint8_t *encodings, uint8_t *ok) {
    if (!ctx || (n && (!encodings || !ok))) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return EG_SUCCESS;
    TRY(ensure(ctx, ctx->in[0], n * 32));
    TRY(ensure(ctx, ctx->verdicts, n));
    CU(cudaMemcpyAsync(ctx->in[0].p, encodings, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    launch_elements_validate(ctx, (const uint8_t *)ctx->in[0].p, n, (uint8_t *)ctx->verdicts.p);
    CU(cudaMemcpyAsync(ok, ctx->verdicts.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    return finish_call(ctx);
}

extern "C" eg_status eg_scalars_validate(eg_ctx *ctx, size_t n, const uint8_t *scalars, uint8_t *ok) {
    if (!ctx || (n && (!scalars || !ok))) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return EG_SUCCESS;
    TRY(ensure(ctx, ctx->in[0], n * 32));
    TRY(ensure(ctx, ctx->verdicts, n));
    CU(cudaMemcpyAsync(ctx->in[0].p, scalars, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    launch_scalars_validate(ctx, (const uint8_t *)ctx->in[0].p, n, (uint8_t *)ctx->verdicts.p);
    CU(cudaMemcpyAsync(ok, ctx->verdicts.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    return finish_call(ctx);
}

extern "C" eg_status eg_scalars_from_wide(eg_ctx *ctx, size_t n, const uint8_t *wide, uint8_t *scalars) {
    if (!ctx || (n && (!wide || !scalars))) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return EG_SUCCESS;
    TRY(ensure(ctx, ctx->in[0], n * 64));
    TRY(ensure(ctx, ctx->in[1], n * 32));
    CU(cudaMemcpyAsync(ctx->in[0].p, wide, n * 64, cudaMemcpyHostToDevice, ctx->stream));
    launch_scalars_from_wide(ctx, (const uint8_t *)ctx->in[0].p, n, (uint8_t *)ctx->in[1].p);
    CU(cudaMemcpyAsync(scalars, ctx->in[1].p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    return finish_call(ctx);
}

static eg_status double_mul_impl(eg_ctx *ctx, size_t n, const uint8_t *a, const uint8_t *A, const uint8_t *b, int mode, uint8_t *out, uint8_t *ok) {
    if (!ctx || (n && (!b || !out || (mode == 0 && (!a || !A))))) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (n == 0) return EG_SUCCESS;
    TRY(ensure(ctx, ctx->in[0], n * 32));
    TRY(ensure(ctx, ctx->in[1], n * 32));
    TRY(ensure(ctx, ctx->in[2], n * 32));
    TRY(ensure(ctx, ctx->in[3], n * 32));
    TRY(ensure(ctx, ctx->verdicts, n));
    if (mode == 0) {
        CU(cudaMemcpyAsync(ctx->in[0].p, a, n * 32, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, A, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    }
    CU(cudaMemcpyAsync(ctx->in[2].p, b, n * 32, cudaMemcpyHostToDevice, ctx->stream));
    launch_double_mul(ctx, (const uint8_t *)ctx->in[0].p, (const uint8_t *)ctx->in[1].p, (const uint8_t *)ctx->in[2].p, n, mode,
                      (uint8_t *)ctx->in[3].p, (uint8_t *)ctx->verdicts.p);
    CU(cudaMemcpyAsync(out, ctx->in[3].p, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
    if (ok) CU(cudaMemcpyAsync(ok, ctx->verdicts.p, n, cudaMemcpyDeviceToHost, ctx->stream));
    return finish_call(ctx);
}

extern "C" eg_status eg_double_mul_generator_batch(eg_ctx *ctx, size_t n, const uint8_t *a, const uint8_t *A, const uint8_t *b,
                                                   uint8_t *out, uint8_t *ok) {
    return double_mul_impl(ctx, n, a, A, b, 0, out, ok);
}

extern "C" eg_status eg_mul_generator_batch(eg_ctx *ctx, size_t n, const uint8_t *k, uint8_t *out, uint8_t *ok) {
    return double_mul_impl(ctx, n, nullptr, nullptr, k, 1, out, ok);
}

extern "C" eg_status eg_ciphertexts_sum(eg_ctx *ctx, size_t n_parts, size_t n_cts, const uint8_t *parts, uint8_t *out, uint8_t *ok) {
    if (!ctx || !out || (n_parts && !parts)) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (n_cts == 0) return EG_SUCCESS;
    TRY(ensure(ctx, ctx->in[0], std::max<size_t>(1, n_parts) * n_cts * 64));
    TRY(ensure(ctx, ctx->in[1], n_cts * 64));
    uint32_t *d_bad = ctx->d_status + 4;
    CU(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    if (n_parts) CU(cudaMemcpyAsync(ctx->in[0].p, parts, n_parts * n_cts * 64, cudaMemcpyHostToDevice, ctx->stream));
    launch_ciphertexts_sum(ctx, (const uint8_t *)ctx->in[0].p, n_parts, n_cts, (uint8_t *)ctx->in[1].p, d_bad);
    uint32_t bad = 0;
    CU(cudaMemcpyAsync(out, ctx->in[1].p, n_cts * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    TRY(finish_call(ctx));
    if (ok) *ok = bad ? 0 : 1;
    return EG_SUCCESS;
}

// device-pointer variant: the combine step after the all_gather of per-GPU partial tallies (asynchronous on the
// context's stream; undecodable parts are reported through *d_bad_flag when it is non-null)
extern "C" eg_status eg_ciphertexts_sum_dev(eg_ctx *ctx, size_t n_parts, size_t n_cts, const uint8_t *d_parts, uint8_t *d_out,
                                            uint32_t *d_bad_flag) {
    if (!ctx || !d_out || (n_parts && !d_parts)) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    if (n_cts == 0) return EG_SUCCESS;
    uint32_t *d_bad = d_bad_flag ? d_bad_flag : ctx->d_status + 4;
    CU(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    launch_ciphertexts_sum(ctx, d_parts, n_parts, n_cts, d_out, d_bad);
    CU(cudaGetLastError());
    return EG_SUCCESS;
}

// =================================================================== RangeDecomposition (host logic)

namespace {

struct opt_entry { uint64_t len; eg_range d; };

uint64_t lower_len_estimate(uint64_t ub) { return (uint64_t)std::ceil(std::log2((double)ub) * 3.0); }   // range.rs:302-305

// RangeDecomposition::optimize (range.rs:238-300)
const opt_entry &range_optimize(uint64_t ub, std::map<uint64_t, opt_entry> &memo) {
    auto it = memo.find(ub);
    if (it != memo.end()) return it->second;
    opt_entry opt;
    memset(&opt.d, 0, sizeof opt.d);
    opt.len = ub + 2;
    opt.d.n_rings = 1; opt.d.size[0] = ub; opt.d.step[0] = 1;
    for (uint64_t first = 2;; first++) {
        if (first + 2 > opt.len) break;
        uint64_t remaining = ub - first;
        for (uint64_t mult = 2; mult <= first; mult++) {
            if (remaining % mult != 0) continue;
            uint64_t inner_ub = remaining / mult + 1;
            if (inner_ub < 2) break;
            if (first + 2 + lower_len_estimate(inner_ub) > opt.len) continue;
            const opt_entry inner = range_optimize(inner_ub, memo);
            uint64_t cand_len = first + 2 + inner.len;
            uint32_t cand_rings = 1 + inner.d.n_rings;
            if ((cand_len < opt.len || (cand_len == opt.len && cand_rings < opt.d.n_rings)) && cand_rings <= 64) {
                opt.len = cand_len;
                opt.d = inner.d;
                for (uint32_t i = 0; i < opt.d.n_rings; i++) opt.d.step[i] *= mult;      // combine_mul range.rs:163-171
                opt.d.size[opt.d.n_rings] = first;
                opt.d.step[opt.d.n_rings] = 1;
                opt.d.n_rings++;
            }
        }
    }
    return memo.emplace(ub, opt).first->second;
}

uint64_t range_rings_size(const eg_range &r) { uint64_t s = 0; for (uint32_t i = 0; i < r.n_rings; i++) s += r.size[i]; return s; }

// RangeDecomposition::upper_bound (range.rs:131-137): 1 + sum (size - 1) * step
uint64_t range_upper_bound(const eg_range &r) { uint64_t u = 1; for (uint32_t i = 0; i < r.n_rings; i++) u += (r.size[i] - 1) * r.step[i]; return u; }

bool range_valid(const eg_range *r) {
    if (!r || r->n_rings == 0 || r->n_rings > 64) return false;
    for (uint32_t i = 0; i < r->n_rings; i++) if (r->size[i] < 1 || r->size[i] > 4096 || r->step[i] == 0) return false;
    return range_rings_size(*r) <= 3500;
}

uint64_t isqrt_u64(uint64_t x) {            // quadratic_voting.rs:127-143
    uint64_t root = 0, p4 = 1ULL << 62;
    while (p4 > x) p4 /= 4;
    while (p4 > 0) {
        if (x >= root + p4) { x -= root + p4; root = root / 2 + p4; } else root /= 2;
        p4 /= 4;
    }
    return root;
}

}  // namespace

extern "C" eg_status eg_range_optimal(uint64_t upper_bound, eg_range *out) {
    if (!out || upper_bound < 2) return EG_ERR_INVALID_ARG;     // range.rs:149 assert
    std::map<uint64_t, opt_entry> memo;
    *out = range_optimize(upper_bound, memo).d;
    return EG_SUCCESS;
}

extern "C" size_t eg_range_display(const eg_range *r, char *buf, size_t cap) {     // range.rs:110-124
    if (!r || !buf || cap == 0) return 0;
    std::string s;
    for (uint32_t i = 0; i < r->n_rings; i++) {
        if (r->step[i] > 1) s += std::to_string(r->step[i]) + " * ";
        s += "0.." + std::to_string(r->size[i]);
        if (i + 1 < r->n_rings) s += " + ";
    }
    size_t n = std::min(cap - 1, s.size());
    memcpy(buf, s.data(), n);
    buf[n] = 0;
    return n;
}

extern "C" eg_status eg_qv_params_new(uint32_t options, uint64_t credits, eg_qv_params *out) {   // quadratic_voting.rs:63-76
    if (!out || options == 0 || credits == 0) return EG_ERR_INVALID_ARG;
    memset(out, 0, sizeof *out);
    out->options = options;
    out->credits = credits;
    eg_status st = eg_range_optimal(isqrt_u64(credits) + 1, &out->vote_range);
    if (st != EG_SUCCESS) return st;
    return eg_range_optimal(credits + 1, &out->credit_range);
}

static size_t range_item_size(const eg_range &r) { return 64 + 64 * (size_t)(r.n_rings - 1) + 32 * (1 + (size_t)range_rings_size(r)); }

extern "C" size_t eg_qv_ballot_size(const eg_qv_params *p) {
    if (!p) return 0;
    return p->options * range_item_size(p->vote_range) + range_item_size(p->credit_range) + 32 * (2 * (size_t)p->options + 2);
}

// =================================================================== RangeProof engine

// admissible values of a range, cached on the device by Display string: adm_base[r] + j = cached([j * step_r] G)
static eg_status ensure_range_adm(eg_ctx *ctx, const eg_range &range, int32_t adm_base[EG_MAX_RINGS]) {
    TRY(ensure_bool_adm(ctx));
    if (ctx->adm_used < 2) ctx->adm_used = 2;
    char key[4096];
    eg_range_display(&range, key, sizeof key);
    auto it = ctx->adm_cache_key.find(key);
    const size_t total = (size_t)range_rings_size(range);
    if (it == ctx->adm_cache_key.end()) {
        if (ctx->adm_used + total > 4096) {      // evict everything but the [O, G] pair
            CU(cudaStreamSynchronize(ctx->stream));
            for (auto i2 = ctx->adm_cache_key.begin(); i2 != ctx->adm_cache_key.end();)
                if (i2->first != "bool") i2 = ctx->adm_cache_key.erase(i2); else ++i2;
            ctx->adm_used = 2;
        }
        std::vector<uint64_t> vals, bases;
        size_t off = ctx->adm_used;
        for (uint32_t r = 0; r < range.n_rings; r++) {
            bases.push_back(off);
            for (uint64_t j = 0; j < range.size[r]; j++) vals.push_back(j * range.step[r]);     // PreparedRange::new range.rs:341-355
            off += (size_t)range.size[r];
        }
        TRY(ensure(ctx, ctx->consts, std::max<size_t>(vals.size() * 8, 4096)));
        CU(cudaMemcpyAsync(ctx->consts.p, vals.data(), vals.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
        launch_admissible(ctx, (const uint64_t *)ctx->consts.p, (int)vals.size(), (uint32_t *)ctx->adm.p + ctx->adm_used * 32);
        CU(cudaStreamSynchronize(ctx->stream));
        ctx->adm_used = off;
        it = ctx->adm_cache_key.emplace(key, bases).first;
    }
    for (uint32_t r = 0; r < range.n_rings; r++) adm_base[r] = (int32_t)it->second[r];
    return EG_SUCCESS;
}

struct range_layout { uint8_t ct_buf, partial_buf, ring_buf; uint32_t ct_off, partial_off, ring_off; };

// RangeProof::verify (range.rs:547-577) for n items addressed through `in` / `lay`; results (1 = ok) to d_result,
// malformed flags to d_flags.  Uses the context scratch from index 0.
static eg_status verify_range_items(eg_ctx *ctx, size_t n, const eg_range &range, const in_bufs &in, const range_layout &lay,
                                    const char *label, uint32_t *d_flags, uint32_t *d_result) {
    const uint32_t R = range.n_rings, T = (uint32_t)range_rings_size(range);
    const uint32_t np = 2 * R + 2;
    TRY(ensure(ctx, ctx->pts, n * np * 128));
    TRY(ensure(ctx, ctx->enc, n * np * 32));
    TRY(ensure(ctx, ctx->commit, n * 2 * R * 32));
    TRY(ensure(ctx, ctx->chal, n * R * 32));
    int32_t adm_base[EG_MAX_RINGS];
    TRY(ensure_range_adm(ctx, range, adm_base));
    CU(cudaMemsetAsync(d_flags, 0, n * 4, ctx->stream));
    // points: partial k -> 2k, 2k+1 ; last ring -> 2(R-1), 2(R-1)+1 (derived) ; main ciphertext -> 2R, 2R+1
    for (uint32_t k0 = 0; k0 < 2 * R; k0 += EG_MAX_SLOTS) {
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in = in; dp.n = n;
        int ns = 0;
        for (uint32_t k = k0; k < 2 * R && ns < EG_MAX_SLOTS; k++, ns++) {
            decode_slot &s = dp.slots[ns];
            if (k < 2) { s.buf = lay.ct_buf; s.offset = lay.ct_off + 32 * k; s.p_index = 2 * R + k; s.want_enc = 0; }
            else { s.buf = lay.partial_buf; s.offset = lay.partial_off + 32 * (k - 2); s.p_index = k - 2; s.want_enc = 1; s.enc_index = (uint16_t)(k - 2); }
        }
        dp.n_slots = ns;
        dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = d_flags;
        launch_decode(ctx, dp);
    }
    scalars_params sp;
    memset(&sp, 0, sizeof sp);
    sp.in = in; sp.n = n; sp.n_slots = 1; sp.slots[0].buf = lay.ring_buf; sp.slots[0].offset = lay.ring_off; sp.slots[0].count = 1 + T;
    sp.flags = d_flags;
    launch_scalars(ctx, sp);
    range_last_params rl;
    memset(&rl, 0, sizeof rl);
    rl.n = n; rl.n_partial = R - 1; rl.ct_p_index = 2 * R; rl.out_p_index = 2 * (R - 1); rl.out_enc_index = 2 * (R - 1);
    rl.pts = (uint32_t *)ctx->pts.p; rl.enc = (uint32_t *)ctx->enc.p;
    launch_range_last(ctx, rl);

    ring_job job;
    job.n_rings = R;
    for (uint32_t r = 0; r < R; r++) {
        job.sizes[r] = (uint32_t)range.size[r]; job.ct_p_index[r] = 2 * r; job.ct_enc_index[r] = 2 * r; job.adm_index[r] = adm_base[r]; job.adm_step[r] = range.step[r];
    }
    job.proof_buf = lay.ring_buf; job.proof_offset = lay.ring_off; job.commit_index0 = 0; job.chal_index0 = 0;
    merlin_new(job.prefix, label, (uint32_t)strlen(label));
    char display[4096];
    size_t dlen = eg_range_display(&range, display, sizeof display);
    merlin_append_message(job.prefix, EG_LBL("dom-sep"), (const uint8_t *)"encryption_range_proof", 22);    // range.rs:561
    merlin_append_message(job.prefix, EG_LBL("range"), (const uint8_t *)display, (uint32_t)dlen);          // range.rs:562
    host_ring_initialize(job.prefix, ctx->key);
    job.d_result = d_result;
    return run_ring_job(ctx, job, in, n, nullptr, 0, (const uint32_t *)ctx->adm.p);
}

static eg_status verify_range_chunk(eg_ctx *ctx, const eg_range &range, const char *label, size_t n, const uint8_t *d_cts,
                                    const uint8_t *d_partial, const uint8_t *d_rings, uint8_t *d_verdicts) {
    const uint32_t R = range.n_rings, T = (uint32_t)range_rings_size(range);
    TRY(ensure(ctx, ctx->flags, n * 4));
    TRY(ensure(ctx, ctx->res[0], n * 4));
    in_bufs in;
    memset(&in, 0, sizeof in);
    in.buf[0] = d_cts; in.stride[0] = 64;
    in.buf[1] = d_partial; in.stride[1] = 64 * (R - 1);
    in.buf[2] = d_rings; in.stride[2] = 32 * (1 + T);
    range_layout lay = {0, 1, 2, 0, 0, 0};
    TRY(verify_range_items(ctx, n, range, in, lay, label, (uint32_t *)ctx->flags.p, (uint32_t *)ctx->res[0].p));
    verdict_params vp;
    memset(&vp, 0, sizeof vp);
    vp.n = n; vp.flags = (const uint32_t *)ctx->flags.p; vp.n_checks = 1;
    vp.check[0] = (const uint32_t *)ctx->res[0].p; vp.check_stride[0] = 1; vp.code[0] = EG_V_CHALLENGE_MISMATCH;
    vp.verdicts = d_verdicts;
    launch_verdict(ctx, vp);
    return EG_SUCCESS;
}

extern "C" eg_status eg_verify_range_batch_dev(eg_ctx *ctx, const eg_range *range, const char *label, size_t n, const uint8_t *d_cts,
                                               const uint8_t *d_partial, const uint8_t *d_rings, uint8_t *d_verdicts) {
    TRY(begin_call(ctx));
    if (!range_valid(range) || !label) return fail(ctx, EG_ERR_INVALID_ARG, "invalid range decomposition or label");
    if (n == 0) return EG_SUCCESS;
    if (!d_cts || !d_rings || !d_verdicts || (range->n_rings > 1 && !d_partial)) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint32_t R = range->n_rings, T = (uint32_t)range_rings_size(*range);
    const size_t chunk = std::max<size_t>(1024, default_chunk(ctx) * 12 / (2 * T + 2));   // keep the scratch footprint of a choice chunk
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        TRY(verify_range_chunk(ctx, *range, label, m, d_cts + 64 * off, d_partial ? d_partial + 64 * (size_t)(R - 1) * off : nullptr,
                               d_rings + 32 * (size_t)(1 + T) * off, d_verdicts + off));
    }
    return finish_call(ctx);
}

extern "C" eg_status eg_verify_range_batch(eg_ctx *ctx, const eg_range *range, const char *label, size_t n, const uint8_t *cts,
                                           const uint8_t *partial, const uint8_t *rings, uint8_t *verdicts) {
    TRY(begin_call(ctx));
    if (!range_valid(range) || !label) return fail(ctx, EG_ERR_INVALID_ARG, "invalid range decomposition or label");
    if (n == 0) return EG_SUCCESS;
    if (!cts || !rings || !verdicts || (range->n_rings > 1 && !partial)) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint32_t R = range->n_rings, T = (uint32_t)range_rings_size(*range);
    const size_t chunk = std::max<size_t>(1024, default_chunk(ctx) * 12 / (2 * T + 2));
    const size_t cm = std::min(chunk, n), pstride = 64 * (size_t)(R - 1), rstride = 32 * (size_t)(1 + T);
    TRY(ensure(ctx, ctx->in[0], cm * 64));
    TRY(ensure(ctx, ctx->in[1], std::max<size_t>(cm * pstride, 64)));
    TRY(ensure(ctx, ctx->in[2], cm * rstride));
    TRY(ensure(ctx, ctx->verdicts, cm));
    for (size_t off = 0; off < n; off += chunk) {
        size_t m = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, cts + 64 * off, m * 64, cudaMemcpyHostToDevice, ctx->stream));
        if (R > 1) CU(cudaMemcpyAsync(ctx->in[1].p, partial + pstride * off, m * pstride, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[2].p, rings + rstride * off, m * rstride, cudaMemcpyHostToDevice, ctx->stream));
        TRY(verify_range_chunk(ctx, *range, label, m, (const uint8_t *)ctx->in[0].p, (const uint8_t *)ctx->in[1].p,
                               (const uint8_t *)ctx->in[2].p, (uint8_t *)ctx->verdicts.p));
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, m, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== QuadraticVotingBallot::verify + tally

static void launch_qv_verdict(eg_ctx *ctx, const qv_verdict_params &P);
static void launch_share_verdict(eg_ctx *ctx, const share_verdict_params &P);

static eg_status upload_slots(eg_ctx *ctx, const std::vector<msm_slot> &slots) {
    TRY(ensure(ctx, ctx->slots, std::max<size_t>(slots.size() * sizeof(msm_slot), 4096)));
    CU(cudaStreamSynchronize(ctx->stream));      // previous launches may still read the slot table
    CU(cudaMemcpyAsync(ctx->slots.p, slots.data(), slots.size() * sizeof(msm_slot), cudaMemcpyHostToDevice, ctx->stream));
    return EG_SUCCESS;
}

static scalar_src src_in(uint8_t buf, uint32_t offset, bool negate) { scalar_src s; memset(&s, 0, sizeof s); s.kind = 0; s.buf = buf; s.offset = offset; s.negate = negate; return s; }
static scalar_src src_const(uint32_t index) { scalar_src s; memset(&s, 0, sizeof s); s.kind = 2; s.offset = index; return s; }

static eg_status verify_qv_chunk(eg_ctx *ctx, const eg_qv_params &qp, size_t n, const uint8_t *d_ballots, uint8_t *d_verdicts,
                                 bool want_tally, bool first_chunk) {
    const uint32_t m = qp.options;
    const size_t vsz = range_item_size(qp.vote_range), csz = range_item_size(qp.credit_range), bsz = eg_qv_ballot_size(&qp);
    const uint32_t Rv = qp.vote_range.n_rings, Rc = qp.credit_range.n_rings;
    // flags / results: [votes n*m][credit n][sumsq n] each for flags and results
    const size_t words = n * m + 2 * n;
    TRY(ensure(ctx, ctx->res_big, 2 * words * 4));
    uint32_t *f_votes = (uint32_t *)ctx->res_big.p, *f_credit = f_votes + n * m, *f_sumsq = f_credit + n;
    uint32_t *r_votes = f_sumsq + n, *r_credit = r_votes + n * m, *r_sumsq = r_credit + n;

    CU(cudaEventRecord(ctx->ev[0], ctx->stream));
    // ---- vote range proofs: items = (ballot, option), quadratic_voting.rs:296-306
    in_bufs in;
    memset(&in, 0, sizeof in);
    in.buf[0] = d_ballots; in.stride[0] = (uint32_t)bsz; in.group[0] = m; in.inner[0] = (uint32_t)vsz;
    range_layout lay_v = {0, 0, 0, 0, 64, (uint32_t)(64 + 64 * (Rv - 1))};
    TRY(verify_range_items(ctx, n * m, qp.vote_range, in, lay_v, "quadratic_voting_variant", f_votes, r_votes));
    // ---- credit range proof, quadratic_voting.rs:308-316
    memset(&in, 0, sizeof in);
    in.buf[0] = d_ballots; in.stride[0] = (uint32_t)bsz;
    range_layout lay_c = {0, 0, 0, (uint32_t)(vsz * m), (uint32_t)(vsz * m + 64), (uint32_t)(vsz * m + 64 + 64 * (Rc - 1))};
    TRY(verify_range_items(ctx, n, qp.credit_range, in, lay_c, "quadratic_voting_credit_range", f_credit, r_credit));
    // ---- sum-of-squares proof over (votes, credit), quadratic_voting.rs:318-326 -> mul.rs:190-260
    const uint32_t np = 2 * m + 2;
    TRY(ensure(ctx, ctx->pts, n * np * 128));
    TRY(ensure(ctx, ctx->enc, n * np * 32));
    TRY(ensure(ctx, ctx->commit, n * np * 32));
    CU(cudaMemsetAsync(f_sumsq, 0, n * 4, ctx->stream));
    const uint32_t proof_off = (uint32_t)(vsz * m + csz);
    for (uint32_t k0 = 0; k0 < np; k0 += EG_MAX_SLOTS) {
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in = in; dp.n = n;
        int ns = 0;
        for (uint32_t k = k0; k < np && ns < EG_MAX_SLOTS; k++, ns++) {
            decode_slot &s = dp.slots[ns];
            s.buf = 0; s.want_enc = 1; s.enc_index = (uint16_t)k; s.p_index = k;
            s.offset = (uint32_t)(vsz * (k / 2) + 32 * (k % 2));      // option k/2 (k/2 == m: the credit ciphertext follows the votes)
        }
        dp.n_slots = ns;
        dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = f_sumsq;
        launch_decode(ctx, dp);
    }
    scalars_params sp;
    memset(&sp, 0, sizeof sp);
    sp.in = in; sp.n = n; sp.n_slots = 1; sp.slots[0].buf = 0; sp.slots[0].offset = proof_off; sp.slots[0].count = 2 * m + 2;
    sp.flags = f_sumsq;
    launch_scalars(ctx, sp);
    // proof = challenge | (r_resp_i, v_resp_i)* | sum_resp
    std::vector<msm_slot> slots;
    const uint32_t c_off = proof_off, sum_off = proof_off + 32 * (1 + 2 * m);
    for (uint32_t i = 0; i < m; i++) {
        const uint32_t r_off = proof_off + 32 * (1 + 2 * i), v_off = r_off + 32;
        msm_slot a;                                    // [e_r]G = [-c]R_x + [r_resp]G, mul.rs:215-219
        memset(&a, 0, sizeof a);
        a.nv = 1; a.nf = 1; a.out_enc = 1; a.out_index = 2 * i;
        a.p_index[0] = 2 * i; a.vs[0] = src_in(0, c_off, true);
        a.fbase[0] = 0; a.fs[0] = src_in(0, r_off, false);
        slots.push_back(a);
        msm_slot b;                                    // [v_resp]G + [r_resp]K + [-c]X, mul.rs:221-228
        memset(&b, 0, sizeof b);
        b.nv = 1; b.nf = 2; b.out_enc = 1; b.out_index = 2 * i + 1;
        b.p_index[0] = 2 * i + 1; b.vs[0] = src_in(0, c_off, true);
        b.fbase[0] = 0; b.fs[0] = src_in(0, v_off, false);
        b.fbase[1] = 1; b.fs[1] = src_in(0, r_off, false);
        slots.push_back(b);
    }
    for (int side = 0; side < 2; side++) {             // mul.rs:232-247: sum_i [v_resp_i]{R_x, X}_i + [sum_resp]{G, K} + [-c]{R_z, Z}
        msm_slot z;
        memset(&z, 0, sizeof z);
        z.nv = (uint8_t)(m + 1); z.nf = 1; z.out_enc = 1; z.out_index = 2 * m + side;
        for (uint32_t i = 0; i < m; i++) { z.p_index[i] = 2 * i + side; z.vs[i] = src_in(0, proof_off + 32 * (2 + 2 * i), false); }
        z.p_index[m] = 2 * m + side; z.vs[m] = src_in(0, c_off, true);
        z.fbase[0] = (uint8_t)side; z.fs[0] = src_in(0, sum_off, false);
        slots.push_back(z);
    }
    TRY(upload_slots(ctx, slots));
    msm_params mp;
    memset(&mp, 0, sizeof mp);
    mp.in = in; mp.n = n; mp.n_slots = (int)slots.size(); mp.slots = (const msm_slot *)ctx->slots.p;
    mp.pts = (const uint32_t *)ctx->pts.p; mp.commit = (uint32_t *)ctx->commit.p; mp.pts_out = (uint32_t *)ctx->pts.p;
    mp.table_g = ctx->d_table_g; mp.table_k = ctx->d_table_k;
    launch_msm(ctx, mp);
    sumsq_final_params fp;
    memset(&fp, 0, sizeof fp);
    fp.in = in; fp.n = n; fp.n_cts = m;
    for (uint32_t i = 0; i < m; i++) fp.ct_enc_index[i] = 2 * i;
    fp.sum_enc_index = 2 * m; fp.commit_index = 0; fp.proof_buf = 0; fp.c_offset = c_off;
    merlin_new(fp.prefix, EG_LBL("quadratic_voting_credit_equiv"));        // quadratic_voting.rs:324
    merlin_append_message(fp.prefix, EG_LBL("dom-sep"), (const uint8_t *)"sum_of_squares", 14);   // mul.rs:96-99
    merlin_append_message(fp.prefix, EG_LBL("K"), ctx->key, 32);
    fp.enc = (const uint32_t *)ctx->enc.p; fp.commit = (const uint32_t *)ctx->commit.p; fp.result = r_sumsq;
    launch_sumsq_final(ctx, fp);
    qv_verdict_params vp;
    memset(&vp, 0, sizeof vp);
    vp.n = n; vp.options = m; vp.flags_votes = f_votes; vp.flags_credit = f_credit; vp.flags_sumsq = f_sumsq;
    vp.res_votes = r_votes; vp.res_credit = r_credit; vp.res_sumsq = r_sumsq; vp.verdicts = d_verdicts;
    launch_qv_verdict(ctx, vp);
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    if (want_tally) {
        TRY(ensure(ctx, ctx->partial, (size_t)2 * m * EG_TALLY_BLOCKS * 128));
        TRY(ensure(ctx, ctx->running, (size_t)2 * m * 128));
        const int blocks = launch_tally_partial(ctx, (const uint32_t *)ctx->pts.p, n, (int)(2 * m), d_verdicts, (uint32_t *)ctx->partial.p);
        launch_tally_final(ctx, (const uint32_t *)ctx->partial.p, blocks, (int)(2 * m), (uint32_t *)ctx->running.p, first_chunk ? 0 : 1, nullptr);
    }
    CU(cudaEventRecord(ctx->ev[3], ctx->stream));
    return EG_SUCCESS;
}

static void launch_qv_verdict(eg_ctx *ctx, const qv_verdict_params &P) {
#ifdef EG_HOSTSIM
    EG_FOR_HOST(P.n, qv_verdict_body(P, tid))
#else
    k_qv_verdict<<<grid_for(P.n, 256), 256, 0, ctx->stream>>>(P);
#endif
    ctx->launches++;
}

static void launch_share_verdict(eg_ctx *ctx, const share_verdict_params &P) {
    size_t total = P.n * (size_t)P.n_shares;
#ifdef EG_HOSTSIM
    EG_FOR_HOST(total, share_verdict_body(P, tid))
#else
    k_share_verdict<<<grid_for(total, 256), 256, 0, ctx->stream>>>(P, total);
#endif
    ctx->launches++;
}

extern "C" eg_status eg_verify_qv_batch(eg_ctx *ctx, const eg_qv_params *params, size_t n, const uint8_t *ballots, uint8_t *verdicts,
                                        uint8_t *tally) {
    TRY(begin_call(ctx));
    if (!params || params->options == 0 || params->options >= EG_MSM_MAXV || !range_valid(&params->vote_range) ||
        !range_valid(&params->credit_range))
        return fail(ctx, EG_ERR_INVALID_ARG, "invalid quadratic voting parameters (options must be in 1..15)");
    if (n && (!ballots || !verdicts)) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint32_t m = params->options;
    const size_t bsz = eg_qv_ballot_size(params);
    const size_t chunk = std::max<size_t>(1024, default_chunk(ctx) / 4), cm = std::min(chunk, std::max<size_t>(n, 1));
    float acc[5] = {0, 0, 0, 0, 0};
    TRY(ensure(ctx, ctx->in[0], cm * bsz));
    TRY(ensure(ctx, ctx->verdicts, cm));
    TRY(ensure(ctx, ctx->misc, 64 * (size_t)m));
    TRY(ensure(ctx, ctx->partial, (size_t)2 * m * EG_TALLY_BLOCKS * 128));
    TRY(ensure(ctx, ctx->running, (size_t)2 * m * 128));
    bool first = true;
    for (size_t off = 0; off < n; off += chunk) {
        size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, ballots + bsz * off, k * bsz, cudaMemcpyHostToDevice, ctx->stream));
        TRY(verify_qv_chunk(ctx, *params, k, (const uint8_t *)ctx->in[0].p, (uint8_t *)ctx->verdicts.p, tally != nullptr, first));
        first = false;
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, k, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        collect_timings(ctx, acc);
    }
    if (tally) {
        launch_tally_final(ctx, (const uint32_t *)ctx->partial.p, 0, (int)(2 * m), (uint32_t *)ctx->running.p, first ? 0 : 1, (uint8_t *)ctx->misc.p);
        CU(cudaMemcpyAsync(tally, ctx->misc.p, 64 * (size_t)m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    { float keep = ctx->timings[2]; memcpy(ctx->timings, acc, sizeof acc); ctx->timings[2] = keep; }
    return finish_call(ctx);
}

// =================================================================== threshold decryption shares

extern "C" eg_status eg_verify_shares_batch(eg_ctx *ctx, const eg_keyset *ks, size_t n, uint32_t n_shares, const uint32_t *indexes,
                                            const uint8_t *cts, const uint8_t *shares, const uint8_t *proofs, uint8_t *verdicts) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear(); ctx->commit_ev_used = 0; ctx->call_commit_tasks = 0; ctx->call_commit_launches = 0; ctx->kind_tasks[0] = ctx->kind_tasks[1] = 0; ctx->kind_launches[0] = ctx->kind_launches[1] = 0;
    if (!ks || !indexes || n_shares == 0 || n_shares > 8 || ks->shares == 0 || ks->shares > 64 || ks->threshold == 0 || ks->threshold > ks->shares)
        return fail(ctx, EG_ERR_INVALID_ARG, "invalid key set / share count (at most 8 shares per call)");
    for (uint32_t j = 0; j < n_shares; j++)
        if (indexes[j] >= ks->shares) return fail(ctx, EG_ERR_INVALID_ARG, "participant index out of bounds");      // key_set.rs:216 panics
    if (n == 0) return EG_SUCCESS;
    if (!cts || !shares || !proofs || !verdicts) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint32_t S = n_shares;
    // participant keys -> constant points (decoded on the device)
    TRY(ensure(ctx, ctx->consts, 64 * 1024));
    uint8_t keys[8 * 32];
    for (uint32_t j = 0; j < S; j++) memcpy(keys + 32 * j, ks->participant_keys[indexes[j]], 32);
    uint8_t *d_keys = (uint8_t *)ctx->consts.p;                       // [0, 256): key encodings
    uint32_t *d_const_pts = (uint32_t *)((uint8_t *)ctx->consts.p + 1024);          // 8 points x 128 B
    uint32_t *d_key_flags = (uint32_t *)((uint8_t *)ctx->consts.p + 4096);
    share_final_params *d_fp = (share_final_params *)((uint8_t *)ctx->consts.p + 8192);
    CU(cudaMemcpyAsync(d_keys, keys, 32 * S, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemsetAsync(d_key_flags, 0, 4, ctx->stream));
    {
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in.buf[0] = d_keys; dp.in.stride[0] = 32 * S; dp.n = 1; dp.n_slots = (int)S;
        for (uint32_t j = 0; j < S; j++) { dp.slots[j].buf = 0; dp.slots[j].offset = 32 * j; dp.slots[j].p_index = j; }
        dp.pts = d_const_pts; dp.enc = nullptr; dp.flags = d_key_flags;
        launch_decode(ctx, dp);
        uint32_t kf = 0;
        CU(cudaMemcpyAsync(&kf, d_key_flags, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
        if (kf) return fail(ctx, EG_ERR_INVALID_ELEMENT, "a participant key does not represent a group element");
    }
    const size_t chunk = default_chunk(ctx), cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * 64));
    TRY(ensure(ctx, ctx->in[1], cm * 32 * S));
    TRY(ensure(ctx, ctx->in[2], cm * 64 * S));
    TRY(ensure(ctx, ctx->verdicts, cm * S));
    TRY(ensure(ctx, ctx->pts, cm * (2 + S) * 128));
    TRY(ensure(ctx, ctx->enc, cm * (2 + S) * 32));
    TRY(ensure(ctx, ctx->commit, cm * 2 * S * 32));
    TRY(ensure(ctx, ctx->flags, cm * (S + 1) * 4));
    TRY(ensure(ctx, ctx->res[0], cm * S * 4));
    for (size_t off = 0; off < n; off += chunk) {
        size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, cts + 64 * off, k * 64, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, shares + 32 * S * off, k * 32 * S, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[2].p, proofs + 64 * S * off, k * 64 * S, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->flags.p, 0, k * (S + 1) * 4, ctx->stream));
        in_bufs in;
        memset(&in, 0, sizeof in);
        in.buf[0] = (const uint8_t *)ctx->in[0].p; in.stride[0] = 64;
        in.buf[1] = (const uint8_t *)ctx->in[1].p; in.stride[1] = 32 * S;
        in.buf[2] = (const uint8_t *)ctx->in[2].p; in.stride[2] = 64 * S;
        // points: 0 = R, 1 = B, 2 + j = share j (CandidateDecryption::from_bytes, decryption.rs:168-177)
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in = in; dp.n = k; dp.n_slots = (int)(2 + S); dp.flag_stride = S + 1;
        for (uint32_t q = 0; q < 2 + S; q++) {
            decode_slot &s = dp.slots[q];
            s.want_enc = 1; s.enc_index = (uint16_t)q; s.p_index = q;
            if (q < 2) { s.buf = 0; s.offset = 32 * q; s.flag_offset = 0; }
            else { s.buf = 1; s.offset = 32 * (q - 2); s.flag_offset = 1 + (q - 2); }
        }
        dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = (uint32_t *)ctx->flags.p;
        launch_decode(ctx, dp);
        scalars_params sp;
        memset(&sp, 0, sizeof sp);
        sp.in = in; sp.n = k; sp.n_slots = (int)S; sp.flag_stride = S + 1;
        for (uint32_t j = 0; j < S; j++) { sp.slots[j].buf = 2; sp.slots[j].offset = 64 * j; sp.slots[j].count = 2; sp.slots[j].flag_offset = 1 + j; }
        sp.flags = (uint32_t *)ctx->flags.p;
        launch_scalars(ctx, sp);
        // log_equality.rs:160-164 with log base R: [x]G = [-c]key_j + [s]G ; [x]K = [-c]dh_j + [s]R
        std::vector<msm_slot> slots;
        for (uint32_t j = 0; j < S; j++) {
            msm_slot a;
            memset(&a, 0, sizeof a);
            a.nv = 1; a.nf = 1; a.out_enc = 1; a.out_index = 2 * j;
            a.p_index[0] = 0x80000000u | j; a.vs[0] = src_in(2, 64 * j, true);
            a.fbase[0] = 0; a.fs[0] = src_in(2, 64 * j + 32, false);
            slots.push_back(a);
            msm_slot b;
            memset(&b, 0, sizeof b);
            b.nv = 2; b.nf = 0; b.out_enc = 1; b.out_index = 2 * j + 1;
            b.p_index[0] = 2 + j; b.vs[0] = src_in(2, 64 * j, true);
            b.p_index[1] = 0; b.vs[1] = src_in(2, 64 * j + 32, false);
            slots.push_back(b);
        }
        TRY(upload_slots(ctx, slots));
        msm_params mp;
        memset(&mp, 0, sizeof mp);
        mp.in = in; mp.n = k; mp.n_slots = (int)slots.size(); mp.slots = (const msm_slot *)ctx->slots.p;
        mp.pts = (const uint32_t *)ctx->pts.p; mp.const_pts = d_const_pts; mp.commit = (uint32_t *)ctx->commit.p;
        mp.pts_out = (uint32_t *)ctx->pts.p; mp.table_g = ctx->d_table_g; mp.table_k = ctx->d_table_k;
        launch_msm(ctx, mp);
        share_final_params fp;
        memset(&fp, 0, sizeof fp);
        fp.in = in; fp.n = k; fp.n_shares = S; fp.r_enc_index = 0; fp.share_enc_index0 = 2; fp.commit_index0 = 0; fp.proof_buf = 2;
        for (uint32_t j = 0; j < S; j++) {
            transcript &t = fp.prefix[j];
            merlin_new(t, EG_LBL("elgamal_decryption_share"));                 // key_set.rs:218
            merlin_append_u64(t, EG_LBL("n"), ks->shares);                     // commit, key_set.rs:167-171
            merlin_append_u64(t, EG_LBL("t"), ks->threshold);
            merlin_append_message(t, EG_LBL("K"), ks->shared_key, 32);
            merlin_append_u64(t, EG_LBL("i"), indexes[j]);                     // key_set.rs:220
            for (int w = 0; w < 8; w++)
                fp.key_words[j][w] = (uint32_t)keys[32 * j + 4 * w] | ((uint32_t)keys[32 * j + 4 * w + 1] << 8) |
                                     ((uint32_t)keys[32 * j + 4 * w + 2] << 16) | ((uint32_t)keys[32 * j + 4 * w + 3] << 24);
        }
        fp.enc = (const uint32_t *)ctx->enc.p; fp.commit = (const uint32_t *)ctx->commit.p; fp.result = (uint32_t *)ctx->res[0].p;
        CU(cudaMemcpyAsync(d_fp, &fp, sizeof fp, cudaMemcpyHostToDevice, ctx->stream));
        launch_share_final(ctx, fp, d_fp);
        share_verdict_params vp;
        memset(&vp, 0, sizeof vp);
        vp.n = k; vp.n_shares = S; vp.flags = (const uint32_t *)ctx->flags.p; vp.result = (const uint32_t *)ctx->res[0].p;
        vp.verdicts = (uint8_t *)ctx->verdicts.p;
        launch_share_verdict(ctx, vp);
        CU(cudaMemcpyAsync(verdicts + S * off, ctx->verdicts.p, k * S, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== PublicKey::encrypt_range / RangeProof::new

static void range_prover_shape(rprove_params &P, const eg_range &range) {
    uint32_t start = 0;
    P.n_rings = range.n_rings;
    for (uint32_t r = 0; r < range.n_rings; r++) {
        P.sizes[r] = (uint16_t)range.size[r]; P.starts[r] = (uint16_t)start; P.steps[r] = range.step[r];
        start += (uint32_t)range.size[r];
    }
    P.total = start;
}

static void range_prover_prefix(transcript &t, const eg_range &range, const char *label, const uint8_t key[32]) {
    merlin_new(t, label, (uint32_t)strlen(label));
    char display[4096];
    size_t dlen = eg_range_display(&range, display, sizeof display);
    merlin_append_message(t, EG_LBL("dom-sep"), (const uint8_t *)"encryption_range_proof", 22);     // range.rs:491
    merlin_append_message(t, EG_LBL("range"), (const uint8_t *)display, (uint32_t)dlen);            // range.rs:492
    host_ring_initialize(t, key);
}

extern "C" size_t eg_range_prover_draws(const eg_range *range) {
    return range ? (size_t)range->n_rings + (size_t)range_rings_size(*range) : 0;
}

// keys/impls.rs:121-141 (encrypt_range) -> range.rs:462-534.  values[i] must be below the range's upper bound (the
// reference panics, range.rs:365-369): checked up front, EG_ERR_INVALID_ARG.
extern "C" eg_status eg_encrypt_range_batch(eg_ctx *ctx, const eg_range *range, const char *label, size_t n, const uint64_t *values,
                                            const uint8_t *wide_rand, uint8_t *cts, uint8_t *partials, uint8_t *rings) {
    TRY(begin_call(ctx));
    if (!range || !range_valid(range)) return fail(ctx, EG_ERR_INVALID_ARG, "invalid range decomposition");
    if (!valid_label(label)) return fail(ctx, EG_ERR_INVALID_ARG, "transcript label must be 1..255 bytes");
    if (n == 0) return EG_SUCCESS;
    const uint32_t R = range->n_rings, T = (uint32_t)range_rings_size(*range);
    if (!values || !wide_rand || !cts || !rings || (R > 1 && !partials)) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint64_t ub = range_upper_bound(*range);
    for (size_t i = 0; i < n; i++)
        if (values[i] >= ub) return fail(ctx, EG_ERR_INVALID_ARG, "a value is outside the range");
    const size_t draws = (size_t)R + T, ring_stride = 32 * (1 + (size_t)T), part_stride = 64 * (size_t)(R - 1);
    const size_t chunk = std::max<size_t>(256, default_chunk(ctx) * 8 / (T + R)), cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * 8));
    TRY(ensure(ctx, ctx->in[1], cm * draws * 64));
    TRY(ensure(ctx, ctx->in[2], cm * 64));
    TRY(ensure(ctx, ctx->in[3], cm * ring_stride));
    TRY(ensure(ctx, ctx->misc, std::max<size_t>(cm * part_stride, 4096)));
    TRY(ensure(ctx, ctx->pts, cm * 2 * R * 128));
    TRY(ensure(ctx, ctx->enc, cm * 2 * R * 32));
    TRY(ensure(ctx, ctx->commit, cm * 2 * R * 32));
    TRY(ensure(ctx, ctx->res_big, cm * 2 * R * 32));
    TRY(ensure(ctx, ctx->chal, cm * 32));
    rprove_params P;
    memset(&P, 0, sizeof P);
    range_prover_shape(P, *range);
    range_prover_prefix(P.prefix, *range, label, ctx->key);
    P.values = (const uint64_t *)ctx->in[0].p; P.value_stride = 1;
    P.wide = (const uint8_t *)ctx->in[1].p; P.wide_stride = draws * 64;
    P.ct_out = (uint8_t *)ctx->in[2].p; P.ct_stride = 64;
    P.partial_out = (uint8_t *)ctx->misc.p; P.partial_stride = part_stride;
    P.ring_out = (uint8_t *)ctx->in[3].p; P.ring_stride = ring_stride;
    P.pts = (uint32_t *)ctx->pts.p; P.enc = (uint32_t *)ctx->enc.p; P.sec = (uint32_t *)ctx->res_big.p;
    P.commit = (uint32_t *)ctx->commit.p; P.chal = (uint32_t *)ctx->chal.p;
    P.table_g = ctx->d_table_g; P.table_k = ctx->d_table_k;
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        P.n = k;
        CU(cudaMemcpyAsync(ctx->in[0].p, values + off, k * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, wide_rand + off * draws * 64, k * draws * 64, cudaMemcpyHostToDevice, ctx->stream));
        TRY(launch_rprove(ctx, P));
        CU(cudaMemcpyAsync(cts + off * 64, ctx->in[2].p, k * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(rings + off * ring_stride, ctx->in[3].p, k * ring_stride, cudaMemcpyDeviceToHost, ctx->stream));
        if (R > 1) CU(cudaMemcpyAsync(partials + off * part_stride, ctx->misc.p, k * part_stride, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== QuadraticVotingBallot::new

extern "C" size_t eg_qv_prover_draws(const eg_qv_params *p) {
    if (!p) return 0;
    return (size_t)p->options * eg_range_prover_draws(&p->vote_range) + eg_range_prover_draws(&p->credit_range) + 1 + 2 * (size_t)p->options;
}

// quadratic_voting.rs:234-284: per option RangeProof::new over the vote range ("quadratic_voting_variant"), RangeProof::new
// of credit = sum votes^2 over the credit range ("quadratic_voting_credit_range"), SumOfSquaresProof::new
// ("quadratic_voting_credit_equiv"); randomness is consumed in that order.  A vote or a credit outside its range makes the
// reference panic (range.rs:365-369): EG_ERR_INVALID_ARG here.
extern "C" eg_status eg_encrypt_qv_batch(eg_ctx *ctx, const eg_qv_params *params, size_t n, const uint64_t *votes, const uint8_t *wide_rand,
                                         uint8_t *ballots) {
    TRY(begin_call(ctx));
    if (!params || params->options == 0 || params->options >= EG_MSM_MAXV || !range_valid(&params->vote_range) ||
        !range_valid(&params->credit_range))
        return fail(ctx, EG_ERR_INVALID_ARG, "invalid quadratic voting parameters (options must be in 1..15)");
    if (n == 0) return EG_SUCCESS;
    if (!votes || !wide_rand || !ballots) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const eg_qv_params &qp = *params;
    const uint32_t m = qp.options, Rv = qp.vote_range.n_rings, Rc = qp.credit_range.n_rings;
    const uint32_t Tv = (uint32_t)range_rings_size(qp.vote_range), Tc = (uint32_t)range_rings_size(qp.credit_range);
    const size_t vsz = range_item_size(qp.vote_range), csz = range_item_size(qp.credit_range), bsz = eg_qv_ballot_size(&qp);
    const size_t Dv = (size_t)Rv + Tv, Dc = (size_t)Rc + Tc, D = m * Dv + Dc + 1 + 2 * (size_t)m;
    const uint64_t ub_v = range_upper_bound(qp.vote_range), ub_c = range_upper_bound(qp.credit_range);
    std::vector<uint64_t> credits(n);
    for (size_t i = 0; i < n; i++) {
        uint64_t credit = 0;
        for (uint32_t o = 0; o < m; o++) {
            const uint64_t v = votes[i * m + o];
            if (v >= ub_v) return fail(ctx, EG_ERR_INVALID_ARG, "a vote is outside the vote range");
            credit += v * v;
        }
        if (credit >= ub_c) return fail(ctx, EG_ERR_INVALID_ARG, "the credit of a ballot is outside the credit range");
        credits[i] = credit;
    }
    const size_t chunk = std::max<size_t>(256, default_chunk(ctx) / 8), cm = std::min(chunk, n);
    const uint32_t Rmax = std::max(Rv, Rc);
    TRY(ensure(ctx, ctx->in[0], cm * m * 8));
    TRY(ensure(ctx, ctx->in[1], cm * D * 64));
    TRY(ensure(ctx, ctx->in[2], cm * bsz));
    TRY(ensure(ctx, ctx->in[3], cm * 8));
    TRY(ensure(ctx, ctx->pts, cm * m * 2 * Rmax * 128));
    TRY(ensure(ctx, ctx->enc, cm * m * 2 * Rmax * 32));
    TRY(ensure(ctx, ctx->commit, cm * m * 2 * Rmax * 32));
    TRY(ensure(ctx, ctx->res_big, cm * m * 2 * Rmax * 32));
    TRY(ensure(ctx, ctx->chal, cm * m * 32));
    TRY(ensure(ctx, ctx->res[1], cm * m * 32));
    TRY(ensure(ctx, ctx->res[2], cm * 32));
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, votes + off * m, k * m * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, wide_rand + off * D * 64, k * D * 64, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[3].p, credits.data() + off, k * 8, cudaMemcpyHostToDevice, ctx->stream));
        uint8_t *d_ballots = (uint8_t *)ctx->in[2].p;
        const uint8_t *d_wide = (const uint8_t *)ctx->in[1].p;
        rprove_params P;
        // ---- votes: items = (ballot, option)
        memset(&P, 0, sizeof P);
        range_prover_shape(P, qp.vote_range);
        range_prover_prefix(P.prefix, qp.vote_range, "quadratic_voting_variant", ctx->key);       // quadratic_voting.rs:253
        P.n = k * m; P.group = m; P.wide_inner = Dv * 64; P.out_inner = vsz;
        P.values = (const uint64_t *)ctx->in[0].p; P.value_stride = 1;
        P.wide = d_wide; P.wide_stride = D * 64;
        P.ct_out = d_ballots; P.ct_stride = bsz;
        P.partial_out = d_ballots + 64; P.partial_stride = bsz;
        P.ring_out = d_ballots + 64 + 64 * (size_t)(Rv - 1); P.ring_stride = bsz;
        P.pts = (uint32_t *)ctx->pts.p; P.enc = (uint32_t *)ctx->enc.p; P.sec = (uint32_t *)ctx->res_big.p;
        P.commit = (uint32_t *)ctx->commit.p; P.chal = (uint32_t *)ctx->chal.p; P.ct_sec = (uint32_t *)ctx->res[1].p;
        P.table_g = ctx->d_table_g; P.table_k = ctx->d_table_k;
        TRY(launch_rprove(ctx, P));
        // ---- credit
        memset(&P, 0, sizeof P);
        range_prover_shape(P, qp.credit_range);
        range_prover_prefix(P.prefix, qp.credit_range, "quadratic_voting_credit_range", ctx->key);   // quadratic_voting.rs:263
        P.n = k; P.group = 0;
        P.values = (const uint64_t *)ctx->in[3].p; P.value_stride = 1;
        P.wide = d_wide + m * Dv * 64; P.wide_stride = D * 64;
        P.ct_out = d_ballots + m * vsz; P.ct_stride = bsz;
        P.partial_out = d_ballots + m * vsz + 64; P.partial_stride = bsz;
        P.ring_out = d_ballots + m * vsz + 64 + 64 * (size_t)(Rc - 1); P.ring_stride = bsz;
        P.pts = (uint32_t *)ctx->pts.p; P.enc = (uint32_t *)ctx->enc.p; P.sec = (uint32_t *)ctx->res_big.p;
        P.commit = (uint32_t *)ctx->commit.p; P.chal = (uint32_t *)ctx->chal.p; P.ct_sec = (uint32_t *)ctx->res[2].p;
        P.table_g = ctx->d_table_g; P.table_k = ctx->d_table_k;
        TRY(launch_rprove(ctx, P));
        // ---- credit equivalence
        sumsq_prove_params S;
        memset(&S, 0, sizeof S);
        S.n = k; S.m = m; S.values = (const uint64_t *)ctx->in[0].p;
        S.wide = d_wide + (m * Dv + Dc) * 64; S.wide_stride = D * 64;
        S.cts = d_ballots; S.ct_stride = bsz; S.ct_inner = vsz;
        S.sum_ct = d_ballots + m * vsz;
        S.proof = d_ballots + m * vsz + csz; S.proof_stride = bsz;
        S.r_cts = (const uint32_t *)ctx->res[1].p; S.r_sum = (const uint32_t *)ctx->res[2].p;
        merlin_new(S.prefix, EG_LBL("quadratic_voting_credit_equiv"));                               // quadratic_voting.rs:272
        merlin_append_message(S.prefix, EG_LBL("dom-sep"), (const uint8_t *)"sum_of_squares", 14);   // mul.rs:96-99
        merlin_append_message(S.prefix, EG_LBL("K"), ctx->key, 32);
        S.table_g = ctx->d_table_g; S.table_k = ctx->d_table_k;
        TRY(launch_sumsq_prove(ctx, S));
        CU(cudaMemcpyAsync(ballots + off * bsz, d_ballots, k * bsz, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== PublicKey::encrypt / encrypt_zero

static eg_status encrypt_batch(eg_ctx *ctx, size_t n, const uint64_t *values, const uint8_t *wide, uint8_t *cts, uint8_t *proofs) {
    TRY(begin_call(ctx));
    const bool zero = values == nullptr;
    if (n == 0) return EG_SUCCESS;
    if (!wide || !cts || (zero && !proofs)) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t draws = zero ? 2 : 1;
    const size_t chunk = (size_t)1 << 20, cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * 8));
    TRY(ensure(ctx, ctx->in[1], cm * draws * 64));
    TRY(ensure(ctx, ctx->in[2], cm * 64));
    TRY(ensure(ctx, ctx->in[3], cm * 64));
    encrypt_params P;
    memset(&P, 0, sizeof P);
    P.with_zero_proof = zero ? 1 : 0;
    P.values = (const uint64_t *)ctx->in[0].p; P.wide = (const uint8_t *)ctx->in[1].p;
    P.cts = (uint8_t *)ctx->in[2].p; P.proofs = (uint8_t *)ctx->in[3].p;
    P.table_g = ctx->d_table_g; P.table_k = ctx->d_table_k;
    if (zero) {
        merlin_new(P.prefix, EG_LBL("zero_encryption"));                  // keys/impls.rs:47
        merlin_append_message(P.prefix, EG_LBL("dom-sep"), (const uint8_t *)"log_eq", 6);
        merlin_append_message(P.prefix, EG_LBL("K"), ctx->key, 32);
    }
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        P.n = k;
        if (!zero) CU(cudaMemcpyAsync(ctx->in[0].p, values + off, k * 8, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, wide + off * draws * 64, k * draws * 64, cudaMemcpyHostToDevice, ctx->stream));
        TRY(launch_encrypt(ctx, P));
        CU(cudaMemcpyAsync(cts + off * 64, ctx->in[2].p, k * 64, cudaMemcpyDeviceToHost, ctx->stream));
        if (zero) CU(cudaMemcpyAsync(proofs + off * 64, ctx->in[3].p, k * 64, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

extern "C" eg_status eg_encrypt_batch(eg_ctx *ctx, size_t n, const uint64_t *values, const uint8_t *wide_rand, uint8_t *cts) {
    if (ctx && n && !values) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    static const uint64_t dummy = 0;
    return encrypt_batch(ctx, n, values ? values : &dummy, wide_rand, cts, nullptr);
}

extern "C" eg_status eg_encrypt_zero_batch(eg_ctx *ctx, size_t n, const uint8_t *wide_rand, uint8_t *cts, uint8_t *proofs) {
    return encrypt_batch(ctx, n, nullptr, wide_rand, cts, proofs);
}

// =================================================================== Group::vartime_multi_mul over a batch

// ristretto.rs:139-146: out[i] = sum_j [scalars[i][j]] points[i][j], `terms` <= 16 per item.  ok[i] = 0 (identity
// encoding) when a point does not decode or a scalar is not canonical.
extern "C" eg_status eg_multi_mul_batch(eg_ctx *ctx, size_t n, uint32_t terms, const uint8_t *scalars, const uint8_t *points, uint8_t *out,
                                        uint8_t *ok) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear(); ctx->commit_ev_used = 0; ctx->call_commit_tasks = 0; ctx->call_commit_launches = 0; ctx->kind_tasks[0] = ctx->kind_tasks[1] = 0; ctx->kind_launches[0] = ctx->kind_launches[1] = 0;
    if (terms == 0 || terms > EG_MSM_MAXV) return fail(ctx, EG_ERR_INVALID_ARG, "terms must be in 1..16");
    if (n == 0) return EG_SUCCESS;
    if (!scalars || !points || !out || !ok) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint32_t T = terms;
    const size_t chunk = std::max<size_t>(1024, default_chunk(ctx) * 2 / T), cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * 32 * T));
    TRY(ensure(ctx, ctx->in[1], cm * 32 * T));
    TRY(ensure(ctx, ctx->in[2], cm * 32));
    TRY(ensure(ctx, ctx->verdicts, cm));
    TRY(ensure(ctx, ctx->pts, cm * T * 128));
    TRY(ensure(ctx, ctx->commit, cm * 32));
    TRY(ensure(ctx, ctx->flags, cm * 4));
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, scalars + off * 32 * T, k * 32 * T, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, points + off * 32 * T, k * 32 * T, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->flags.p, 0, k * 4, ctx->stream));
        in_bufs in;
        memset(&in, 0, sizeof in);
        in.buf[0] = (const uint8_t *)ctx->in[0].p; in.stride[0] = 32 * T;
        in.buf[1] = (const uint8_t *)ctx->in[1].p; in.stride[1] = 32 * T;
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in = in; dp.n = k; dp.n_slots = (int)T;
        for (uint32_t q = 0; q < T; q++) { dp.slots[q].buf = 1; dp.slots[q].offset = 32 * q; dp.slots[q].p_index = q; }
        dp.pts = (uint32_t *)ctx->pts.p; dp.enc = nullptr; dp.flags = (uint32_t *)ctx->flags.p;
        launch_decode(ctx, dp);
        scalars_params sp;
        memset(&sp, 0, sizeof sp);
        sp.in = in; sp.n = k; sp.n_slots = 1; sp.slots[0].buf = 0; sp.slots[0].offset = 0; sp.slots[0].count = T;
        sp.flags = (uint32_t *)ctx->flags.p;
        launch_scalars(ctx, sp);
        std::vector<msm_slot> slots(1);
        msm_slot &z = slots[0];
        memset(&z, 0, sizeof z);
        z.nv = (uint8_t)T; z.nf = 0; z.out_enc = 1; z.out_index = 0;
        for (uint32_t j = 0; j < T; j++) { z.p_index[j] = j; z.vs[j] = src_in(0, 32 * j, false); }
        TRY(upload_slots(ctx, slots));
        msm_params mp;
        memset(&mp, 0, sizeof mp);
        mp.in = in; mp.n = k; mp.n_slots = 1; mp.slots = (const msm_slot *)ctx->slots.p;
        mp.pts = (const uint32_t *)ctx->pts.p; mp.commit = (uint32_t *)ctx->commit.p; mp.pts_out = (uint32_t *)ctx->pts.p;
        mp.table_g = ctx->d_table_g; mp.table_k = ctx->has_receiver ? ctx->d_table_k : ctx->d_table_g;
        launch_msm(ctx, mp);
        unpack_params up;
        up.n = k; up.commit = (const uint32_t *)ctx->commit.p; up.flags = (const uint32_t *)ctx->flags.p;
        up.out = (uint8_t *)ctx->in[2].p; up.ok = (uint8_t *)ctx->verdicts.p;
        launch_unpack(ctx, up);
        CU(cudaMemcpyAsync(out + off * 32, ctx->in[2].p, k * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(ok + off, ctx->verdicts.p, k, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== PublicKeySet::from_participants

// sharing/key_set.rs:87-144 for n_sets key sets of the same (shares, threshold): reconstruct the shared key from the
// first `threshold` participant keys (Lagrange interpolation at 0) and check that every other participant key is the
// interpolation of the same polynomial.  All coefficients depend only on (shares, threshold): computed once on the host
// (lagrange_coefficients sharing/mod.rs:139-170, invert_scalars on 1..=n key_set.rs:106-110), with the common scale folded
// into them ((sum c_i K_i) * s == sum (c_i s) K_i).  threshold <= 16.  No receiver key needed.
extern "C" eg_status eg_keysets_validate_batch(eg_ctx *ctx, uint32_t shares, uint32_t threshold, size_t n_sets, const uint8_t *keys,
                                               uint8_t *shared_keys, uint8_t *verdicts) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear(); ctx->commit_ev_used = 0; ctx->call_commit_tasks = 0; ctx->call_commit_launches = 0; ctx->kind_tasks[0] = ctx->kind_tasks[1] = 0; ctx->kind_launches[0] = ctx->kind_launches[1] = 0;
    if (shares == 0 || shares > 64 || threshold == 0 || threshold > shares || threshold > EG_MSM_MAXV)
        return fail(ctx, EG_ERR_INVALID_ARG, "need 1 <= threshold <= shares <= 64 and threshold <= 16");
    if (n_sets == 0) return EG_SUCCESS;
    if (!keys || !shared_keys || !verdicts) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint32_t N = shares, T = threshold, n_out = 1 + (N - T);
    // ---- constants
    std::vector<uint32_t> coeff((size_t)8 * T * n_out);
    {
        sc denom[EG_MSM_MAXV], scale = sc_from_u64(1), inv[64];
        for (uint32_t a = 0; a < T; a++) {
            bool sign = false;
            sc mag = sc_from_u64(1);
            for (uint32_t b2 = 0; b2 < T; b2++) {
                sc e;
                if (a > b2) { sign = !sign; e = sc_from_u64(a - b2); }
                else if (a < b2) e = sc_from_u64(b2 - a);
                else e = sc_from_u64((uint64_t)a + 1);
                sc_mul(mag, mag, e);
            }
            if (sign) sc_neg(mag, mag);
            sc_invert(denom[a], mag);
            sc e = sc_from_u64((uint64_t)a + 1);
            sc_mul(scale, scale, e);
        }
        for (uint32_t i = 0; i < N; i++) sc_invert(inv[i], sc_from_u64((uint64_t)i + 1));
        for (uint32_t a = 0; a < T; a++) {
            sc c;
            sc_mul(c, denom[a], scale);
            for (int w = 0; w < 8; w++) coeff[8 * a + w] = c.v[w];
        }
        for (uint32_t x = T; x < N; x++) {
            sc key_scale = sc_from_u64(1);
            for (uint32_t idx = 0; idx < T; idx++) sc_mul(key_scale, key_scale, sc_from_u64((uint64_t)(x - idx)));
            if (T % 2 == 0) sc_neg(key_scale, key_scale);
            for (uint32_t idx = 0; idx < T; idx++) {
                sc c;
                sc_mul(c, denom[idx], sc_from_u64((uint64_t)idx + 1));
                sc_mul(c, c, inv[x - idx - 1]);
                sc_mul(c, c, key_scale);
                for (int w = 0; w < 8; w++) coeff[8 * ((size_t)T * (1 + x - T) + idx) + w] = c.v[w];
            }
        }
    }
    TRY(ensure(ctx, ctx->consts, 64 * 1024 + coeff.size() * 4));
    uint32_t *d_coeff = (uint32_t *)((uint8_t *)ctx->consts.p + 64 * 1024);
    CU(cudaMemcpyAsync(d_coeff, coeff.data(), coeff.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    std::vector<msm_slot> slots(n_out);
    for (uint32_t o = 0; o < n_out; o++) {
        msm_slot &z = slots[o];
        memset(&z, 0, sizeof z);
        z.nv = (uint8_t)T; z.nf = 0; z.out_enc = 1; z.out_index = o;
        for (uint32_t j = 0; j < T; j++) { z.p_index[j] = j; z.vs[j] = src_const(T * o + j); }
    }
    TRY(upload_slots(ctx, slots));
    const size_t chunk = std::max<size_t>(256, default_chunk(ctx) / N), cm = std::min(chunk, n_sets);
    TRY(ensure(ctx, ctx->in[0], cm * 32 * N));
    TRY(ensure(ctx, ctx->in[1], cm * 32));
    TRY(ensure(ctx, ctx->verdicts, cm));
    TRY(ensure(ctx, ctx->pts, cm * N * 128));
    TRY(ensure(ctx, ctx->commit, cm * n_out * 32));
    TRY(ensure(ctx, ctx->flags, cm * 4));
    for (size_t off = 0; off < n_sets; off += chunk) {
        const size_t k = std::min(chunk, n_sets - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, keys + off * 32 * N, k * 32 * N, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->flags.p, 0, k * 4, ctx->stream));
        in_bufs in;
        memset(&in, 0, sizeof in);
        in.buf[0] = (const uint8_t *)ctx->in[0].p; in.stride[0] = 32 * N;
        for (uint32_t q0 = 0; q0 < N; q0 += EG_MAX_SLOTS) {
            decode_params dp;
            memset(&dp, 0, sizeof dp);
            dp.in = in; dp.n = k;
            int ns = 0;
            for (uint32_t q = q0; q < N && ns < EG_MAX_SLOTS; q++, ns++) {
                decode_slot &d = dp.slots[ns];
                d.buf = 0; d.reject_identity = 1; d.offset = 32 * q; d.p_index = q;
            }
            dp.n_slots = ns;
            dp.pts = (uint32_t *)ctx->pts.p; dp.enc = nullptr; dp.flags = (uint32_t *)ctx->flags.p;
            launch_decode(ctx, dp);
        }
        msm_params mp;
        memset(&mp, 0, sizeof mp);
        mp.in = in; mp.n = k; mp.n_slots = (int)n_out; mp.slots = (const msm_slot *)ctx->slots.p;
        mp.pts = (const uint32_t *)ctx->pts.p; mp.const_scalars = d_coeff; mp.commit = (uint32_t *)ctx->commit.p;
        mp.pts_out = (uint32_t *)ctx->pts.p;
        mp.table_g = ctx->d_table_g; mp.table_k = ctx->has_receiver ? ctx->d_table_k : ctx->d_table_g;
        launch_msm(ctx, mp);
        keyset_verdict_params vp;
        memset(&vp, 0, sizeof vp);
        vp.n = k; vp.shares = N; vp.threshold = T; vp.keys = (const uint8_t *)ctx->in[0].p;
        vp.commit = (const uint32_t *)ctx->commit.p; vp.flags = (const uint32_t *)ctx->flags.p;
        vp.shared_out = (uint8_t *)ctx->in[1].p; vp.verdicts = (uint8_t *)ctx->verdicts.p; vp.code_mismatch = EG_V_MALFORMED_PARTICIPANT_KEYS;
        launch_keyset_verdict(ctx, vp);
        CU(cudaMemcpyAsync(shared_keys + off * 32, ctx->in[1].p, k * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, k, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== wire format (serde.rs:19-80)

static bool b64_shape(b64_params &P, size_t n, size_t bytes_per_item) {
    if (bytes_per_item == 0 || bytes_per_item > (1u << 20)) return false;
    P.n = n; P.bytes = (uint32_t)bytes_per_item; P.chars = (uint32_t)((4 * bytes_per_item + 2) / 3); P.groups = (P.chars + 3) / 4;
    return true;
}

extern "C" size_t eg_base64url_chars(size_t bytes_per_item) { return (4 * bytes_per_item + 2) / 3; }

extern "C" eg_status eg_base64url_decode_batch_dev(eg_ctx *ctx, size_t n, size_t bytes_per_item, const char *d_text, uint8_t *d_raw,
                                                   uint8_t *d_ok) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    b64_params P;
    if (!b64_shape(P, n, bytes_per_item)) return fail(ctx, EG_ERR_INVALID_ARG, "bytes_per_item must be in 1..2^20");
    if (n == 0) return EG_SUCCESS;
    if (!d_text || !d_raw || !d_ok) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    P.text = (uint8_t *)d_text; P.raw = d_raw; P.ok = d_ok;
    CU(cudaMemsetAsync(d_ok, 1, n, ctx->stream));
    launch_b64url(ctx, P, false);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return EG_SUCCESS;
}

extern "C" eg_status eg_base64url_encode_batch_dev(eg_ctx *ctx, size_t n, size_t bytes_per_item, const uint8_t *d_raw, char *d_text) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    b64_params P;
    if (!b64_shape(P, n, bytes_per_item)) return fail(ctx, EG_ERR_INVALID_ARG, "bytes_per_item must be in 1..2^20");
    if (n == 0) return EG_SUCCESS;
    if (!d_text || !d_raw) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    P.text = (uint8_t *)d_text; P.raw = (uint8_t *)d_raw; P.ok = nullptr;
    launch_b64url(ctx, P, true);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    return EG_SUCCESS;
}

// host-buffer variants: chunked H2D -> kernel -> D2H
static eg_status b64_host(eg_ctx *ctx, size_t n, size_t bytes_per_item, const uint8_t *src, uint8_t *dst, uint8_t *ok, bool encode) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear();
    b64_params shape;
    if (!b64_shape(shape, n, bytes_per_item)) return fail(ctx, EG_ERR_INVALID_ARG, "bytes_per_item must be in 1..2^20");
    if (n == 0) return EG_SUCCESS;
    if (!src || !dst || (!encode && !ok)) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t in_sz = encode ? shape.bytes : shape.chars, out_sz = encode ? shape.chars : shape.bytes;
    const size_t chunk = std::max<size_t>(1, ((size_t)256 << 20) / in_sz), cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * in_sz));
    TRY(ensure(ctx, ctx->in[1], cm * out_sz));
    TRY(ensure(ctx, ctx->verdicts, cm));
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        b64_params P = shape;
        P.n = k;
        CU(cudaMemcpyAsync(ctx->in[0].p, src + off * in_sz, k * in_sz, cudaMemcpyHostToDevice, ctx->stream));
        if (encode) { P.raw = (uint8_t *)ctx->in[0].p; P.text = (uint8_t *)ctx->in[1].p; P.ok = nullptr; }
        else {
            P.text = (uint8_t *)ctx->in[0].p; P.raw = (uint8_t *)ctx->in[1].p; P.ok = (uint8_t *)ctx->verdicts.p;
            CU(cudaMemsetAsync(P.ok, 1, k, ctx->stream));
        }
        launch_b64url(ctx, P, encode);
        CU(cudaMemcpyAsync(dst + off * out_sz, ctx->in[1].p, k * out_sz, cudaMemcpyDeviceToHost, ctx->stream));
        if (!encode) CU(cudaMemcpyAsync(ok + off, ctx->verdicts.p, k, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    CU(cudaGetLastError());
    return EG_SUCCESS;
}

extern "C" eg_status eg_base64url_decode_batch(eg_ctx *ctx, size_t n, size_t bytes_per_item, const char *text, uint8_t *raw, uint8_t *ok) {
    return b64_host(ctx, n, bytes_per_item, (const uint8_t *)text, raw, ok, false);
}

extern "C" eg_status eg_base64url_encode_batch(eg_ctx *ctx, size_t n, size_t bytes_per_item, const uint8_t *raw, char *text) {
    return b64_host(ctx, n, bytes_per_item, raw, (uint8_t *)text, nullptr, true);
}

// =================================================================== CommitmentEquivalenceProof::verify


static void sigma_set_msg(sigma_msg &g, uint8_t kind, const char *label, uint32_t index, uint32_t count) {
    memset(&g, 0, sizeof g);
    g.kind = kind; g.label_len = (uint8_t)strlen(label);
    memcpy(g.label, label, g.label_len);
    g.index = index; g.count = count;
}

// commitment.rs:198-248.  Per item: ciphertext R | B (64 B), commitment C (32 B), proof c | s_r | s_v | s_c (128 B).
//   E_r = [s_r]G - [c]R ; E_b = [s_v]G + [s_r]K - [c]B ; E_c = [s_v]G + [s_c]H - [c]C
extern "C" eg_status eg_verify_commitment_equiv_batch(eg_ctx *ctx, const char *label, size_t n, const uint8_t *cts,
                                                      const uint8_t *commitments, const uint8_t *proofs, uint8_t *verdicts) {
    TRY(begin_call(ctx));
    if (!ctx->has_blinding_base) return fail(ctx, EG_ERR_NO_RECEIVER, "eg_ctx_set_blinding_base has not been called");
    if (!valid_label(label)) return fail(ctx, EG_ERR_INVALID_ARG, "transcript label must be 1..255 bytes");
    if (n == 0) return EG_SUCCESS;
    if (!cts || !commitments || !proofs || !verdicts) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t chunk = default_chunk(ctx), cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * 64));
    TRY(ensure(ctx, ctx->in[1], cm * 32));
    TRY(ensure(ctx, ctx->in[2], cm * 128));
    TRY(ensure(ctx, ctx->verdicts, cm));
    TRY(ensure(ctx, ctx->pts, cm * 3 * 128));
    TRY(ensure(ctx, ctx->enc, cm * 3 * 32));
    TRY(ensure(ctx, ctx->commit, cm * 3 * 32));
    TRY(ensure(ctx, ctx->flags, cm * 4));
    TRY(ensure(ctx, ctx->res[0], cm * 4));
    transcript prefix;
    merlin_new(prefix, label, (uint32_t)strlen(label));
    merlin_append_message(prefix, EG_LBL("dom-sep"), (const uint8_t *)"commitment_equivalence", 22);
    merlin_append_message(prefix, EG_LBL("K"), ctx->key, 32);
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, cts + 64 * off, k * 64, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, commitments + 32 * off, k * 32, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[2].p, proofs + 128 * off, k * 128, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->flags.p, 0, k * 4, ctx->stream));
        in_bufs in;
        memset(&in, 0, sizeof in);
        in.buf[0] = (const uint8_t *)ctx->in[0].p; in.stride[0] = 64;
        in.buf[1] = (const uint8_t *)ctx->in[1].p; in.stride[1] = 32;
        in.buf[2] = (const uint8_t *)ctx->in[2].p; in.stride[2] = 128;
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in = in; dp.n = k; dp.n_slots = 3;
        for (uint32_t q = 0; q < 3; q++) {
            decode_slot &s = dp.slots[q];
            s.want_enc = 1; s.enc_index = (uint16_t)q; s.p_index = q;
            if (q < 2) { s.buf = 0; s.offset = 32 * q; } else { s.buf = 1; s.offset = 0; }
        }
        dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = (uint32_t *)ctx->flags.p;
        launch_decode(ctx, dp);
        scalars_params sp;
        memset(&sp, 0, sizeof sp);
        sp.in = in; sp.n = k; sp.n_slots = 1; sp.slots[0].buf = 2; sp.slots[0].offset = 0; sp.slots[0].count = 4;
        sp.flags = (uint32_t *)ctx->flags.p;
        launch_scalars(ctx, sp);
        std::vector<msm_slot> slots(3);
        for (uint32_t q = 0; q < 3; q++) {
            msm_slot &a = slots[q];
            memset(&a, 0, sizeof a);
            a.nv = 1; a.out_enc = 1; a.out_index = q;
            a.p_index[0] = q; a.vs[0] = src_in(2, 0, true);                         // [-c] {R, B, C}
        }
        slots[0].nf = 1; slots[0].fbase[0] = 0; slots[0].fs[0] = src_in(2, 32, false);              // [s_r] G
        slots[1].nf = 2; slots[1].fbase[0] = 0; slots[1].fs[0] = src_in(2, 64, false);              // [s_v] G
        slots[1].fbase[1] = 1; slots[1].fs[1] = src_in(2, 32, false);                               // [s_r] K
        slots[2].nf = 2; slots[2].fbase[0] = 0; slots[2].fs[0] = src_in(2, 64, false);              // [s_v] G
        slots[2].fbase[1] = 2; slots[2].fs[1] = src_in(2, 96, false);                               // [s_c] H
        TRY(upload_slots(ctx, slots));
        msm_params mp;
        memset(&mp, 0, sizeof mp);
        mp.in = in; mp.n = k; mp.n_slots = 3; mp.slots = (const msm_slot *)ctx->slots.p;
        mp.pts = (const uint32_t *)ctx->pts.p; mp.commit = (uint32_t *)ctx->commit.p; mp.pts_out = (uint32_t *)ctx->pts.p;
        mp.table_g = ctx->d_table_g; mp.table_k = ctx->d_table_k; mp.table_h = ctx->d_table_h;
        launch_msm(ctx, mp);
        sigma_final_params fp;
        memset(&fp, 0, sizeof fp);
        fp.in = in; fp.n = k; fp.prefix = prefix; fp.n_msgs = 6;
        sigma_set_msg(fp.msgs[0], 0, "R", 0, 1);
        sigma_set_msg(fp.msgs[1], 0, "B", 1, 1);
        sigma_set_msg(fp.msgs[2], 0, "C", 2, 1);
        sigma_set_msg(fp.msgs[3], 1, "[e_r]G", 0, 1);
        sigma_set_msg(fp.msgs[4], 1, "[e_v]G + [e_r]K", 1, 1);
        sigma_set_msg(fp.msgs[5], 1, "[e_v]G + [e_c]H", 2, 1);
        fp.proof_buf = 2; fp.c_offset = 0;
        fp.enc = (const uint32_t *)ctx->enc.p; fp.commit = (const uint32_t *)ctx->commit.p; fp.result = (uint32_t *)ctx->res[0].p;
        launch_sigma_final(ctx, fp);
        verdict_params vp;
        memset(&vp, 0, sizeof vp);
        vp.n = k; vp.flags = (const uint32_t *)ctx->flags.p; vp.n_checks = 1;
        vp.check[0] = (const uint32_t *)ctx->res[0].p; vp.check_stride[0] = 1; vp.code[0] = EG_V_CHALLENGE_MISMATCH;
        vp.verdicts = (uint8_t *)ctx->verdicts.p;
        launch_verdict(ctx, vp);
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, k, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== ProofOfPossession::verify

// possession.rs:137-163.  Per item: `keys_per_proof` public keys (32 B each) and the proof c | s_0 .. s_{k-1}.
//   R_j = [s_j]G - [c]K_j ; transcript: start_proof("multi_pop"), "K" x k, "R" x k, challenge "c".
// A key that is undecodable or the identity is malformed (PublicKey::from_bytes, keys/mod.rs:161-176).  No receiver needed.
extern "C" eg_status eg_verify_possession_batch(eg_ctx *ctx, const char *label, uint32_t keys_per_proof, size_t n, const uint8_t *keys,
                                                const uint8_t *proofs, uint8_t *verdicts) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear(); ctx->commit_ev_used = 0; ctx->call_commit_tasks = 0; ctx->call_commit_launches = 0; ctx->kind_tasks[0] = ctx->kind_tasks[1] = 0; ctx->kind_launches[0] = ctx->kind_launches[1] = 0;
    if (!valid_label(label)) return fail(ctx, EG_ERR_INVALID_ARG, "transcript label must be 1..255 bytes");
    const uint32_t K = keys_per_proof;
    if (K == 0 || K > EG_MAX_RINGS) return fail(ctx, EG_ERR_INVALID_ARG, "keys_per_proof must be in 1..64");
    if (n == 0) return EG_SUCCESS;
    if (!keys || !proofs || !verdicts) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const size_t chunk = std::max<size_t>(1, default_chunk(ctx) / K), cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * 32 * K));
    TRY(ensure(ctx, ctx->in[1], cm * 32 * (1 + K)));
    TRY(ensure(ctx, ctx->verdicts, cm));
    TRY(ensure(ctx, ctx->pts, cm * K * 128));
    TRY(ensure(ctx, ctx->enc, cm * K * 32));
    TRY(ensure(ctx, ctx->commit, cm * K * 32));
    TRY(ensure(ctx, ctx->flags, cm * 4));
    TRY(ensure(ctx, ctx->res[0], cm * 4));
    transcript prefix;
    merlin_new(prefix, label, (uint32_t)strlen(label));
    merlin_append_message(prefix, EG_LBL("dom-sep"), (const uint8_t *)"multi_pop", 9);
    for (size_t off = 0; off < n; off += chunk) {
        const size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, keys + 32 * (size_t)K * off, k * 32 * K, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, proofs + 32 * (size_t)(1 + K) * off, k * 32 * (1 + K), cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->flags.p, 0, k * 4, ctx->stream));
        in_bufs in;
        memset(&in, 0, sizeof in);
        in.buf[0] = (const uint8_t *)ctx->in[0].p; in.stride[0] = 32 * K;
        in.buf[1] = (const uint8_t *)ctx->in[1].p; in.stride[1] = 32 * (1 + K);
        for (uint32_t q0 = 0; q0 < K; q0 += EG_MAX_SLOTS) {
            decode_params dp;
            memset(&dp, 0, sizeof dp);
            dp.in = in; dp.n = k;
            int ns = 0;
            for (uint32_t q = q0; q < K && ns < EG_MAX_SLOTS; q++, ns++) {
                decode_slot &s = dp.slots[ns];
                s.buf = 0; s.want_enc = 1; s.reject_identity = 1; s.enc_index = (uint16_t)q; s.offset = 32 * q; s.p_index = q;
            }
            dp.n_slots = ns;
            dp.pts = (uint32_t *)ctx->pts.p; dp.enc = (uint32_t *)ctx->enc.p; dp.flags = (uint32_t *)ctx->flags.p;
            launch_decode(ctx, dp);
        }
        scalars_params sp;
        memset(&sp, 0, sizeof sp);
        sp.in = in; sp.n = k; sp.n_slots = 1; sp.slots[0].buf = 1; sp.slots[0].offset = 0; sp.slots[0].count = 1 + K;
        sp.flags = (uint32_t *)ctx->flags.p;
        launch_scalars(ctx, sp);
        for (uint32_t q0 = 0; q0 < K; q0 += EG_MAX_SLOTS) {
            commit_params cp;
            memset(&cp, 0, sizeof cp);
            cp.in = in; cp.n = k;
            cp.pts = (const uint32_t *)ctx->pts.p; cp.commit = (uint32_t *)ctx->commit.p;
            cp.table_g = ctx->d_table_g; cp.table_k = ctx->d_table_g;
            int ns = 0;
            for (uint32_t q = q0; q < K && ns < EG_MAX_SLOTS; q++, ns++) {
                commit_slot &s = cp.slots[ns];
                s.p_index = q; s.adm_index = -1; s.base = 0; s.e_planar = 0; s.e_buf = 1; s.s_buf = 1;
                s.e_offset = 0; s.s_offset = 32 * (1 + q); s.out_index = q;
            }
            cp.n_slots = ns;
            launch_commit(ctx, cp);
        }
        sigma_final_params fp;
        memset(&fp, 0, sizeof fp);
        fp.in = in; fp.n = k; fp.prefix = prefix; fp.n_msgs = 2;
        sigma_set_msg(fp.msgs[0], 0, "K", 0, K);
        sigma_set_msg(fp.msgs[1], 1, "R", 0, K);
        fp.proof_buf = 1; fp.c_offset = 0;
        fp.enc = (const uint32_t *)ctx->enc.p; fp.commit = (const uint32_t *)ctx->commit.p; fp.result = (uint32_t *)ctx->res[0].p;
        launch_sigma_final(ctx, fp);
        verdict_params vp;
        memset(&vp, 0, sizeof vp);
        vp.n = k; vp.flags = (const uint32_t *)ctx->flags.p; vp.n_checks = 1;
        vp.check[0] = (const uint32_t *)ctx->res[0].p; vp.check_stride[0] = 1; vp.code[0] = EG_V_CHALLENGE_MISMATCH;
        vp.verdicts = (uint8_t *)ctx->verdicts.p;
        launch_verdict(ctx, vp);
        CU(cudaMemcpyAsync(verdicts + off, ctx->verdicts.p, k, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}

// =================================================================== DiscreteLogTable + combine_shares + decrypt

extern "C" eg_status eg_dlog_table_create(eg_ctx *ctx, uint64_t lo, uint64_t hi, eg_dlog_table **out) {
    if (!ctx || !out || hi < lo || hi - lo > (1ULL << 28)) return EG_ERR_INVALID_ARG;
    *out = nullptr;
    CU(cudaSetDevice(ctx->device));
    eg_dlog_table *t = new (std::nothrow) eg_dlog_table();
    if (!t) return EG_ERR_OUT_OF_MEMORY;
    t->ctx = ctx; t->lo = lo; t->hi = hi;
    size_t cap = 16;
    while (cap < 2 * (size_t)(hi - lo) + 2) cap <<= 1;
    t->cap = cap; t->d_keys = nullptr; t->d_vals = nullptr;
    if (cudaMalloc(&t->d_keys, cap * 32) != cudaSuccess || cudaMalloc(&t->d_vals, cap * 8) != cudaSuccess) {
        cudaGetLastError();
        eg_dlog_table_destroy(t);
        return fail(ctx, EG_ERR_OUT_OF_MEMORY, "dlog table allocation");
    }
    CU(cudaMemsetAsync(t->d_vals, 0, cap * 8, ctx->stream));
    dlog_build_params bp;
    bp.lo = lo; bp.hi = hi; bp.per = 32; bp.cap = cap; bp.keys = t->d_keys; bp.vals = (unsigned long long *)t->d_vals;
    size_t threads = (size_t)((hi - lo + bp.per - 1) / bp.per);
    if (threads) launch_dlog_build(ctx, bp, threads);
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaGetLastError());
    *out = t;
    return EG_SUCCESS;
}

extern "C" void eg_dlog_table_destroy(eg_dlog_table *t) {
    if (!t) return;
    if (t->d_keys) cudaFree(t->d_keys);
    if (t->d_vals) cudaFree(t->d_vals);
    delete t;
}

extern "C" eg_status eg_combine_decrypt_batch(eg_ctx *ctx, uint32_t threshold, const uint32_t *indexes, size_t n, uint32_t share_stride,
                                              const uint8_t *cts, const uint8_t *shares, const eg_dlog_table *table, uint64_t *values,
                                              uint8_t *found) {
    if (!ctx) return EG_ERR_INVALID_ARG;
    CU(cudaSetDevice(ctx->device));
    ctx->err.clear(); ctx->commit_ev_used = 0; ctx->call_commit_tasks = 0; ctx->call_commit_launches = 0; ctx->kind_tasks[0] = ctx->kind_tasks[1] = 0; ctx->kind_launches[0] = ctx->kind_launches[1] = 0;
    if (!indexes || !table || threshold == 0 || threshold > EG_MSM_MAXV || share_stride < threshold)
        return fail(ctx, EG_ERR_INVALID_ARG, "invalid threshold / share layout");
    if (n == 0) return EG_SUCCESS;
    if (!cts || !shares || !values || !found) return fail(ctx, EG_ERR_INVALID_ARG, "null pointer");
    const uint32_t t = threshold;
    // lagrange_coefficients (sharing/mod.rs:139-170) on the host: t scalars, folded with the common scale
    std::vector<uint32_t> coeff(8 * t);
    {
        sc scale = sc_from_u64(1);
        for (uint32_t a = 0; a < t; a++) { sc e = sc_from_u64((uint64_t)indexes[a] + 1); sc_mul(scale, scale, e); }
        for (uint32_t a = 0; a < t; a++) {
            bool sign = false;
            sc mag = sc_from_u64(1);
            for (uint32_t b = 0; b < t; b++) {
                sc e;
                if (indexes[a] > indexes[b]) { sign = !sign; e = sc_from_u64(indexes[a] - indexes[b]); }
                else if (indexes[a] < indexes[b]) e = sc_from_u64(indexes[b] - indexes[a]);
                else {
                    if (a != b) return fail(ctx, EG_ERR_INVALID_ARG, "duplicate participant index");
                    e = sc_from_u64((uint64_t)indexes[a] + 1);
                }
                sc_mul(mag, mag, e);
            }
            if (sign) sc_neg(mag, mag);
            sc inv, c;
            sc_invert(inv, mag);
            sc_mul(c, inv, scale);          // (sum lambda_j S_j) * scale == sum (lambda_j * scale) S_j
            for (int w = 0; w < 8; w++) coeff[8 * a + w] = c.v[w];
        }
    }
    TRY(ensure(ctx, ctx->consts, 64 * 1024));
    uint32_t *d_coeff = (uint32_t *)((uint8_t *)ctx->consts.p + 16384);
    CU(cudaMemcpyAsync(d_coeff, coeff.data(), coeff.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
    const size_t chunk = default_chunk(ctx), cm = std::min(chunk, n);
    TRY(ensure(ctx, ctx->in[0], cm * 64));
    TRY(ensure(ctx, ctx->in[1], cm * 32 * share_stride));
    TRY(ensure(ctx, ctx->pts, cm * (3 + t) * 128));
    TRY(ensure(ctx, ctx->flags, cm * 4));
    TRY(ensure(ctx, ctx->res_big, cm * 8));
    TRY(ensure(ctx, ctx->verdicts, cm));
    for (size_t off = 0; off < n; off += chunk) {
        size_t k = std::min(chunk, n - off);
        CU(cudaMemcpyAsync(ctx->in[0].p, cts + 64 * off, k * 64, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemcpyAsync(ctx->in[1].p, shares + 32 * (size_t)share_stride * off, k * 32 * share_stride, cudaMemcpyHostToDevice, ctx->stream));
        CU(cudaMemsetAsync(ctx->flags.p, 0, k * 4, ctx->stream));
        in_bufs in;
        memset(&in, 0, sizeof in);
        in.buf[0] = (const uint8_t *)ctx->in[0].p; in.stride[0] = 64;
        in.buf[1] = (const uint8_t *)ctx->in[1].p; in.stride[1] = 32 * share_stride;
        decode_params dp;
        memset(&dp, 0, sizeof dp);
        dp.in = in; dp.n = k; dp.n_slots = (int)(2 + t);
        for (uint32_t q = 0; q < 2 + t; q++) {
            decode_slot &s = dp.slots[q];
            s.p_index = q;
            if (q < 2) { s.buf = 0; s.offset = 32 * q; } else { s.buf = 1; s.offset = 32 * (q - 2); }
        }
        dp.pts = (uint32_t *)ctx->pts.p; dp.enc = nullptr; dp.flags = (uint32_t *)ctx->flags.p;
        launch_decode(ctx, dp);
        std::vector<msm_slot> slots(1);
        msm_slot &z = slots[0];
        memset(&z, 0, sizeof z);
        z.nv = (uint8_t)t; z.nf = 0; z.out_point = 1; z.out_index = 2 + t;
        for (uint32_t j = 0; j < t; j++) { z.p_index[j] = 2 + j; z.vs[j] = src_const(j); }
        TRY(upload_slots(ctx, slots));
        msm_params mp;
        memset(&mp, 0, sizeof mp);
        mp.in = in; mp.n = k; mp.n_slots = 1; mp.slots = (const msm_slot *)ctx->slots.p;
        mp.pts = (const uint32_t *)ctx->pts.p; mp.const_scalars = d_coeff; mp.pts_out = (uint32_t *)ctx->pts.p;
        mp.table_g = ctx->d_table_g; mp.table_k = ctx->has_receiver ? ctx->d_table_k : ctx->d_table_g;
        launch_msm(ctx, mp);
        dlog_lookup_params lp;
        memset(&lp, 0, sizeof lp);
        lp.n = k; lp.b_p_index = 1; lp.d_p_index = 2 + t; lp.pts = (const uint32_t *)ctx->pts.p;
        lp.cap = table->cap; lp.keys = table->d_keys; lp.vals = (const unsigned long long *)table->d_vals;
        lp.flags = (const uint32_t *)ctx->flags.p; lp.values = (unsigned long long *)ctx->res_big.p; lp.found = (uint8_t *)ctx->verdicts.p;
        launch_dlog_lookup(ctx, lp);
        CU(cudaMemcpyAsync(values + off, ctx->res_big.p, k * 8, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(found + off, ctx->verdicts.p, k, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
    }
    return finish_call(ctx);
}
