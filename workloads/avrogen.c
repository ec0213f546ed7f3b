/*
 * avrogen.c — seeded synthetic Avro datum generators for the benchmark configs (BASELINE.json,
 * SURVEY.md 8(d)).  Benchmark/test infrastructure: produces the packed bytes + offsets a caller
 * would hand to rv_decode_host.  The byte grammar is the one the reference's encoder emits
 * (ruhvro/src/fast_encode.rs:397-599: zigzag varints, positive block count + 0 terminator).
 *
 * Record r is generated from SplitMix64(seed, r) alone, so any range can be produced
 * independently (threads, ranks) and reproducibly.
 *
 *   config 2  FLAT_PRIMITIVES (ruhvro/benches/common/mod.rs:37-63): i=r, l=7r, f=1.5r, d=2.25r,
 *             b=(r%2==0), s="row-{r}"
 *   config 3  generate_avro.py "Kafka" schema (scripts/generate_avro.py:12-62) with the value
 *             distributions of its Faker generator (lengths/probabilities; content is random a-z0-9)
 *   config 4  wide-union + map-heavy divergence stress
 *   config 5  ARRAY_AND_MAP (benches/common/mod.rs:137-165)
 *   config 6  NESTED_STRUCT (benches/common/mod.rs:102-135)
 *   config 7  NULLABLE_PRIMITIVES (benches/common/mod.rs:67-100)
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct { uint64_t s; } rng_t;
static inline uint64_t rnext(rng_t *g) {
    uint64_t z = (g->s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint32_t rbelow(rng_t *g, uint32_t n) { return (uint32_t)((rnext(g) >> 32) * (uint64_t)n >> 32); }
static inline uint32_t rrange(rng_t *g, uint32_t lo, uint32_t hi) { return lo + rbelow(g, hi - lo + 1); } /* inclusive */

static inline uint8_t *put_long(uint8_t *p, int64_t v) {
    uint64_t u = ((uint64_t)v << 1) ^ (uint64_t)(v >> 63);
    while (u >= 0x80) { *p++ = (uint8_t)(u | 0x80); u >>= 7; }
    *p++ = (uint8_t)u;
    return p;
}
static const char ALNUM[] = "abcdefghijklmnopqrstuvwxyz0123456789";
static inline uint8_t *put_rand_str(uint8_t *p, rng_t *g, uint32_t len, int alpha) {
    p = put_long(p, len);
    uint32_t i = 0;
    while (i < len) {
        uint64_t x = rnext(g);
        for (int k = 0; k < 10 && i < len; k++, i++, x >>= 6) *p++ = (uint8_t)ALNUM[(x & 63) % (alpha ? 26 : 36)];
    }
    return p;
}
static inline uint8_t *put_str(uint8_t *p, const char *s) {
    size_t n = strlen(s);
    p = put_long(p, (int64_t)n); memcpy(p, s, n); return p + n;
}
static inline uint8_t *put_f32(uint8_t *p, float f) { memcpy(p, &f, 4); return p + 4; }
static inline uint8_t *put_f64(uint8_t *p, double d) { memcpy(p, &d, 8); return p + 8; }

#define MAX_RECORD 2048

static size_t gen_flat(uint64_t r, uint8_t *out) {
    uint8_t *p = out;
    p = put_long(p, (int32_t)r);
    p = put_long(p, (int64_t)r * 7);
    p = put_f32(p, (float)r * 1.5f);
    p = put_f64(p, (double)r * 2.25);
    *p++ = (r % 2 == 0);
    char buf[32]; snprintf(buf, sizeof buf, "row-%llu", (unsigned long long)r);
    p = put_str(p, buf);
    return (size_t)(p - out);
}

static size_t gen_kafka(uint64_t r, uint64_t seed, uint8_t *out) {
    rng_t g = { seed * 0xD1342543DE82EF95ull + r * 0x9E3779B97F4A7C15ull };
    (void)rnext(&g);
    uint8_t *p = out;
    /* name ["null","string"] */
    if (rbelow(&g, 2)) { p = put_long(p, 1); p = put_rand_str(p, &g, rrange(&g, 8, 20), 1); } else p = put_long(p, 0);
    /* age ["null","int"] */
    if (rbelow(&g, 2)) { p = put_long(p, 1); p = put_long(p, rrange(&g, 18, 80)); } else p = put_long(p, 0);
    /* emails array<string> */
    uint32_t ne = rbelow(&g, 4);
    if (ne) { p = put_long(p, ne); for (uint32_t i = 0; i < ne; i++) p = put_rand_str(p, &g, rrange(&g, 15, 30), 0); }
    p = put_long(p, 0);
    /* address ["null", record{street, city, zipcode}] */
    if (rbelow(&g, 2)) {
        p = put_long(p, 1);
        p = put_rand_str(p, &g, rrange(&g, 12, 25), 0);
        p = put_rand_str(p, &g, rrange(&g, 6, 15), 1);
        p = put_rand_str(p, &g, 5, 0);
    } else p = put_long(p, 0);
    /* phone_numbers map<string> */
    uint32_t np = rbelow(&g, 4);
    if (np) { p = put_long(p, np); for (uint32_t i = 0; i < np; i++) { p = put_rand_str(p, &g, rrange(&g, 3, 10), 1); p = put_rand_str(p, &g, rrange(&g, 10, 20), 0); } }
    p = put_long(p, 0);
    /* preferences ["null", record{contact_method ["null","string"], newsletter boolean}] */
    if (rbelow(&g, 2)) {
        p = put_long(p, 1);
        uint32_t cm = rbelow(&g, 3);
        if (cm == 0) p = put_long(p, 0); else { p = put_long(p, 1); p = put_str(p, cm == 1 ? "email" : "phone"); }
        *p++ = (uint8_t)rbelow(&g, 2);
    } else p = put_long(p, 0);
    /* status ["null","string","int","boolean"] */
    uint32_t sv = rbelow(&g, 4);
    p = put_long(p, sv);
    if (sv == 1) p = put_rand_str(p, &g, rrange(&g, 3, 10), 1);
    else if (sv == 2) p = put_long(p, rrange(&g, 0, 100));
    else if (sv == 3) *p++ = (uint8_t)rbelow(&g, 2);
    /* created_at long: epoch seconds within the last year */
    p = put_long(p, 1726000000ll - (int64_t)rbelow(&g, 31536000u));
    /* class enum{A,B,C} */
    p = put_long(p, rbelow(&g, 3));
    return (size_t)(p - out);
}

static size_t gen_wide(uint64_t r, uint64_t seed, uint8_t *out) {
    rng_t g = { seed * 0xA24BAED4963EE407ull + r * 0x9E3779B97F4A7C15ull };
    (void)rnext(&g);
    uint8_t *p = out;
    p = put_long(p, (int64_t)r);
    for (int f = 0; f < 4; f++) { /* ["null","string","int","long","float","double","boolean",enum] */
        uint32_t v = rbelow(&g, 8);
        p = put_long(p, v);
        switch (v) {
        case 1: p = put_rand_str(p, &g, rrange(&g, 3, 24), 1); break;
        case 2: p = put_long(p, (int32_t)rnext(&g)); break;
        case 3: p = put_long(p, (int64_t)rnext(&g)); break;
        case 4: p = put_f32(p, (float)rbelow(&g, 1000000) * 0.25f); break;
        case 5: p = put_f64(p, (double)rbelow(&g, 1000000) * 0.125); break;
        case 6: *p++ = (uint8_t)rbelow(&g, 2); break;
        case 7: p = put_long(p, rbelow(&g, 4)); break;
        default: break;
        }
    }
    for (int m = 0; m < 3; m++) { /* map<string>, map<long>, map<double> */
        uint32_t n = rbelow(&g, 9);
        if (n) {
            p = put_long(p, n);
            for (uint32_t i = 0; i < n; i++) {
                p = put_rand_str(p, &g, rrange(&g, 3, 10), 1);
                if (m == 0) p = put_rand_str(p, &g, rrange(&g, 4, 16), 0);
                else if (m == 1) p = put_long(p, (int64_t)(rnext(&g) >> 20));
                else p = put_f64(p, (double)rbelow(&g, 1u << 30) * 0.5);
            }
        }
        p = put_long(p, 0);
    }
    return (size_t)(p - out);
}

static size_t gen_array_map(uint64_t r, uint8_t *out) {
    uint8_t *p = out; char b[48];
    p = put_long(p, (int64_t)r);
    p = put_long(p, 3);
    for (int i = 0; i < 3; i++) { snprintf(b, sizeof b, "t-%llu-%c", (unsigned long long)r, 'a' + i); p = put_str(p, b); }
    p = put_long(p, 0);
    p = put_long(p, 2);
    for (int i = 1; i <= 2; i++) {
        snprintf(b, sizeof b, "k%llu-%d", (unsigned long long)r, i); p = put_str(p, b);
        snprintf(b, sizeof b, "v%llu-%d", (unsigned long long)r, i); p = put_str(p, b);
    }
    p = put_long(p, 0);
    return (size_t)(p - out);
}
static size_t gen_nested(uint64_t r, uint8_t *out) {
    uint8_t *p = out; char b[32];
    p = put_long(p, (int64_t)r);
    p = put_long(p, (int32_t)r);
    p = put_long(p, (int32_t)(r * 3));
    snprintf(b, sizeof b, "lbl-%llu", (unsigned long long)r); p = put_str(p, b);
    return (size_t)(p - out);
}
static size_t gen_nullable(uint64_t r, uint8_t *out) {
    uint8_t *p = out; char b[32];
    if (r % 2 == 1) { for (int i = 0; i < 5; i++) p = put_long(p, 0); return (size_t)(p - out); }
    p = put_long(p, 1); p = put_long(p, (int32_t)r);
    p = put_long(p, 1); p = put_long(p, (int64_t)r * 7);
    p = put_long(p, 1); p = put_f64(p, (double)r * 2.25);
    p = put_long(p, 1); *p++ = (r % 4 == 0);
    p = put_long(p, 1); snprintf(b, sizeof b, "row-%llu", (unsigned long long)r); p = put_str(p, b);
    return (size_t)(p - out);
}

static size_t gen_record(int config, uint64_t r, uint64_t seed, uint8_t *out) {
    switch (config) {
    case 2: return gen_flat(r, out);
    case 3: return gen_kafka(r, seed, out);
    case 4: return gen_wide(r, seed, out);
    case 5: return gen_array_map(r, out);
    case 6: return gen_nested(r, out);
    case 7: return gen_nullable(r, out);
    default: return 0;
    }
}

typedef struct { int config; uint64_t seed; int64_t r0, a, b; int64_t *lens; const int64_t *offsets; uint8_t *data; } job_t;

static void *lens_worker(void *arg) {
    job_t *j = arg; uint8_t buf[MAX_RECORD];
    for (int64_t i = j->a; i < j->b; i++) j->lens[i] = (int64_t)gen_record(j->config, (uint64_t)(j->r0 + i), j->seed, buf);
    return NULL;
}
static void *fill_worker(void *arg) {
    job_t *j = arg;
    for (int64_t i = j->a; i < j->b; i++) gen_record(j->config, (uint64_t)(j->r0 + i), j->seed, j->data + (j->offsets[i] - j->offsets[0]));
    return NULL;
}
static void run(void *(*fn)(void *), job_t proto, int64_t n, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 64) threads = 64;
    pthread_t th[64]; job_t jobs[64];
    for (int t = 0; t < threads; t++) {
        jobs[t] = proto; jobs[t].a = n * t / threads; jobs[t].b = n * (t + 1) / threads;
        pthread_create(&th[t], NULL, fn, &jobs[t]);
    }
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
}

/* lens[i] = encoded size of record r0+i */
void avrogen_lens(int config, int64_t r0, int64_t n, uint64_t seed, int64_t *lens, int threads) {
    job_t p = { config, seed, r0, 0, 0, lens, NULL, NULL };
    run(lens_worker, p, n, threads);
}
/* writes record r0+i at data + (offsets[i] - offsets[0]); offsets must be the prefix sum of lens */
void avrogen_fill(int config, int64_t r0, int64_t n, uint64_t seed, const int64_t *offsets, uint8_t *data, int threads) {
    job_t p = { config, seed, r0, 0, 0, NULL, offsets, data };
    run(fill_worker, p, n, threads);
}
