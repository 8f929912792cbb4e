/* consumer.c -- a non-Python consumer of the C ABI (include/eg_b200.h), compiled with plain gcc and linked against
 * libeg_b200.so: what a cgo / JNI / Rust -sys binding does, without any of Python's ctypes in between.
 *
 *   consumer <vectors.bin>
 *
 * vectors.bin (written by tests/test_cabi_consumer.py from the committed golden fixtures and the oracle) is a sequence of
 * cases:  u32 options | u32 single | u64 n | key[32] | choices n*options*64 | rings n*(1+2 options)*32 | sums n*64 |
 * expected verdicts n | expected tally options*64.  Case 0 is the reference's own `encrypted-choice` snapshot
 * (tests/snapshots.rs:111-119), the rest a seeded mini batch of the BASELINE config-2 shape with tampered ballots.
 * Every case runs through eg_verify_choice_batch on a single-device context and again on a multi-device context
 * (eg_ctx_create_multi over all visible GPUs).  Exit code 0 = all equal, 77 = no CUDA device (there is no CPU fallback),
 * anything else = failure.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "eg_b200.h"

static int run_case(eg_ctx *ctx, const char *what, uint32_t options, uint32_t single, uint64_t n, const uint8_t *key,
                    const uint8_t *choices, const uint8_t *rings, const uint8_t *sums, const uint8_t *exp_v, const uint8_t *exp_t) {
    uint8_t *v = malloc(n ? n : 1), *t = malloc(64 * (size_t)options);
    eg_status st = eg_ctx_set_receiver(ctx, key);
    if (st == EG_SUCCESS) st = eg_verify_choice_batch(ctx, n, options, (int)single, choices, rings, single ? sums : NULL, v, t);
    int bad = 0;
    if (st != EG_SUCCESS) {
        fprintf(stderr, "%s: status %d: %s\n", what, st, eg_last_error(ctx));
        bad = 1;
    } else {
        if (memcmp(v, exp_v, n) != 0) { fprintf(stderr, "%s: verdict mismatch\n", what); bad = 1; }
        if (memcmp(t, exp_t, 64 * (size_t)options) != 0) { fprintf(stderr, "%s: tally mismatch\n", what); bad = 1; }
    }
    free(v); free(t);
    return bad;
}

int main(int argc, char **argv) {
    if (argc < 2) { fprintf(stderr, "usage: consumer vectors.bin\n"); return 2; }
    FILE *f = fopen(argv[1], "rb");
    if (!f) { perror(argv[1]); return 2; }
    fseek(f, 0, SEEK_END);
    long size = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *buf = malloc((size_t)size);
    if (fread(buf, 1, (size_t)size, f) != (size_t)size) { fprintf(stderr, "short read\n"); return 2; }
    fclose(f);

    printf("%s\n", eg_version());
    eg_ctx *ctx = NULL, *multi = NULL;
    eg_status st = eg_ctx_create(0, &ctx);
    if (st == EG_ERR_NO_DEVICE) { printf("no CUDA device: the library has no CPU fallback\n"); return 77; }
    if (st != EG_SUCCESS) { fprintf(stderr, "eg_ctx_create: %d\n", st); return 1; }
    int devices[8], n_dev = 0;
    for (int d = 0; d < 8; d++) {           /* probe how many devices the library can open */
        eg_ctx *probe = NULL;
        if (eg_ctx_create(d, &probe) != EG_SUCCESS) break;
        eg_ctx_destroy(probe);
        devices[n_dev++] = d;
    }
    st = eg_ctx_create_multi(devices, n_dev, &multi);
    if (st != EG_SUCCESS) { fprintf(stderr, "eg_ctx_create_multi(%d devices): %d\n", n_dev, st); return 1; }

    int failures = 0, cases = 0;
    size_t off = 0;
    while (off + 48 <= (size_t)size) {
        uint32_t options, single;
        uint64_t n;
        memcpy(&options, buf + off, 4); memcpy(&single, buf + off + 4, 4); memcpy(&n, buf + off + 8, 8);
        const uint8_t *key = buf + off + 16;
        const uint8_t *choices = key + 32, *rings = choices + n * options * 64, *sums = rings + n * (1 + 2 * (size_t)options) * 32;
        const uint8_t *exp_v = sums + n * 64, *exp_t = exp_v + n;
        off = (size_t)(exp_t + 64 * (size_t)options - buf);
        if (off > (size_t)size) { fprintf(stderr, "truncated vectors\n"); return 2; }
        char what[64];
        snprintf(what, sizeof what, "case %d (1 device)", cases);
        failures += run_case(ctx, what, options, single, n, key, choices, rings, sums, exp_v, exp_t);
        snprintf(what, sizeof what, "case %d (%d devices)", cases, n_dev);
        failures += run_case(multi, what, options, single, n, key, choices, rings, sums, exp_v, exp_t);
        cases++;
    }
    int devs = 0;
    eg_ctx_comm_info(multi, NULL, NULL, &devs);
    printf("%d cases, %d failures, multi-device context over %d GPU(s), %llu kernel launches\n", cases, failures, devs,
           (unsigned long long)(eg_kernel_launch_count(ctx) + eg_kernel_launch_count(multi)));
    eg_ctx_destroy(multi);
    eg_ctx_destroy(ctx);
    free(buf);
    return failures ? 1 : 0;
}
