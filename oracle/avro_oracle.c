/*
 * avro_oracle.c — CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C, row-at-a-time restatement of the reference's direct Avro->Arrow
 * decode path.  Only tests/, __graft_entry__.smoke() and bench.py's CPU
 * baseline legs may load this library; the product (pyruhvro_b200/) never does.
 *
 * What is restated (file:line in /root/reference):
 *   - ruhvro/src/fast_decode.rs:38-61    is_supported gate
 *   - ruhvro/src/fast_decode.rs:176-414  decoder-tree construction (make_*)
 *   - ruhvro/src/fast_decode.rs:420-534  FieldDecoder::decode / append_null
 *   - ruhvro/src/fast_decode.rs:536-567  finish (arrow-rs builder semantics, see below)
 *   - ruhvro/src/fast_decode.rs:585-593  union_branch
 *   - ruhvro/src/fast_decode.rs:595-799  Record/Union/List/Map decoders
 *   - ruhvro/src/fast_decode.rs:815-835  decode_with_arrow_schema (entry, trailing bytes ignored)
 *   - ruhvro/src/fast_decode.rs:845-922  wire primitives (varint/zigzag, f32/f64, bool, string)
 *   - ruhvro/src/deserialize.rs:53-68    clamp_chunks / build_slices (chunking)
 *
 * Third-party semantics that live outside /root/reference (restated from their
 * published behaviour, pinned versions from the workspace Cargo.lock):
 *   - arrow 58.3.0 builders: primitive/string/bool builders materialise a
 *     validity bitmap lazily on the first append_null; null slots hold
 *     0 / repeated offset / false; bool values are LSB-first bit-packed.
 *     Nullable record/list/map carry an explicit BooleanBufferBuilder, so their
 *     validity is always present (fast_decode.rs:133,151,165,629,739,790).
 *   - apache-avro 0.21.0 Schema::parse_str: the JSON schema grammar accepted below.
 *
 * Parity status: pinned against the reference's literal golden datums
 * (deserialize.rs:244,303; lib.rs:165-167) in tests/test_oracle_golden.py and
 * cross-checked against an independent pure-Python restatement
 * (oracle/pyoracle.py).  The reference itself cannot be compiled here (Rust
 * toolchain absent), so there is no oracle/_ref.
 *
 * Output format: a flat pre-order list of Arrow arrays, each described by
 * orc_array (kind, length, null_count, up to three buffers, child count).  The
 * Python side (oracle/pyoracle.py) turns that into pyarrow arrays.
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* tiny JSON                                                                  */
/* ------------------------------------------------------------------------- */
typedef enum { J_NULL, J_BOOL, J_NUM, J_STR, J_ARR, J_OBJ } jkind;
typedef struct jval {
    jkind k;
    char *s;              /* J_STR */
    int n;                /* J_ARR / J_OBJ */
    struct jval **items;  /* J_ARR values / J_OBJ values */
    char **keys;          /* J_OBJ */
} jval;

typedef struct { const char *p, *end; int err; } jparser;

static void *xmalloc(size_t n) { void *p = malloc(n ? n : 1); if (!p) abort(); return p; }
static void *xrealloc(void *q, size_t n) { void *p = realloc(q, n ? n : 1); if (!p) abort(); return p; }
static char *xstrdup(const char *s) { size_t n = strlen(s); char *d = xmalloc(n + 1); memcpy(d, s, n + 1); return d; }

static void jskip(jparser *P) {
    while (P->p < P->end && (*P->p == ' ' || *P->p == '\t' || *P->p == '\n' || *P->p == '\r')) P->p++;
}
static jval *jnew(jkind k) { jval *v = xmalloc(sizeof *v); memset(v, 0, sizeof *v); v->k = k; return v; }
static void jfree(jval *v) {
    if (!v) return;
    free(v->s);
    for (int i = 0; i < v->n; i++) { jfree(v->items[i]); if (v->keys) free(v->keys[i]); }
    free(v->items); free(v->keys); free(v);
}
static int hexv(char c) {
    if (c >= '0' && c <= '9') return c - '0';
    if (c >= 'a' && c <= 'f') return c - 'a' + 10;
    if (c >= 'A' && c <= 'F') return c - 'A' + 10;
    return -1;
}
static void put_utf8(char **o, unsigned cp) {
    char *d = *o;
    if (cp < 0x80) *d++ = (char)cp;
    else if (cp < 0x800) { *d++ = (char)(0xC0 | (cp >> 6)); *d++ = (char)(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { *d++ = (char)(0xE0 | (cp >> 12)); *d++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *d++ = (char)(0x80 | (cp & 0x3F)); }
    else { *d++ = (char)(0xF0 | (cp >> 18)); *d++ = (char)(0x80 | ((cp >> 12) & 0x3F)); *d++ = (char)(0x80 | ((cp >> 6) & 0x3F)); *d++ = (char)(0x80 | (cp & 0x3F)); }
    *o = d;
}
static char *jstring(jparser *P) {
    if (P->p >= P->end || *P->p != '"') { P->err = 1; return NULL; }
    P->p++;
    char *out = xmalloc((size_t)(P->end - P->p) + 1), *d = out;
    while (P->p < P->end && *P->p != '"') {
        char c = *P->p++;
        if (c != '\\') { *d++ = c; continue; }
        if (P->p >= P->end) { P->err = 1; break; }
        c = *P->p++;
        switch (c) {
        case 'n': *d++ = '\n'; break; case 't': *d++ = '\t'; break; case 'r': *d++ = '\r'; break;
        case 'b': *d++ = '\b'; break; case 'f': *d++ = '\f'; break;
        case 'u': {
            unsigned cp = 0;
            for (int i = 0; i < 4; i++) { int h = (P->p < P->end) ? hexv(*P->p++) : -1; if (h < 0) { P->err = 1; h = 0; } cp = cp * 16 + (unsigned)h; }
            if (cp >= 0xD800 && cp < 0xDC00 && P->p + 6 <= P->end && P->p[0] == '\\' && P->p[1] == 'u') {
                unsigned lo = 0; P->p += 2;
                for (int i = 0; i < 4; i++) { int h = hexv(*P->p++); if (h < 0) { P->err = 1; h = 0; } lo = lo * 16 + (unsigned)h; }
                cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            put_utf8(&d, cp);
            break;
        }
        default: *d++ = c; break; /* \" \\ \/ */
        }
    }
    if (P->p >= P->end) { P->err = 1; free(out); return NULL; }
    P->p++; *d = 0;
    return out;
}
static jval *jparse(jparser *P) {
    jskip(P);
    if (P->p >= P->end) { P->err = 1; return NULL; }
    char c = *P->p;
    if (c == '"') { jval *v = jnew(J_STR); v->s = jstring(P); return v; }
    if (c == '{' || c == '[') {
        int obj = c == '{'; char close = obj ? '}' : ']';
        jval *v = jnew(obj ? J_OBJ : J_ARR); P->p++;
        int cap = 0;
        jskip(P);
        if (P->p < P->end && *P->p == close) { P->p++; return v; }
        for (;;) {
            if (v->n == cap) { cap = cap ? cap * 2 : 8; v->items = xrealloc(v->items, sizeof(jval *) * (size_t)cap); if (obj) v->keys = xrealloc(v->keys, sizeof(char *) * (size_t)cap); }
            if (obj) {
                jskip(P); char *k = jstring(P); if (P->err) { free(k); return v; }
                jskip(P); if (P->p >= P->end || *P->p != ':') { P->err = 1; free(k); return v; } P->p++;
                v->keys[v->n] = k;
            }
            v->items[v->n] = jparse(P); v->n++;
            if (P->err) return v;
            jskip(P);
            if (P->p < P->end && *P->p == ',') { P->p++; continue; }
            if (P->p < P->end && *P->p == close) { P->p++; return v; }
            P->err = 1; return v;
        }
    }
    if (!strncmp(P->p, "true", 4) && P->end - P->p >= 4) { P->p += 4; return jnew(J_BOOL); }
    if (!strncmp(P->p, "false", 5) && P->end - P->p >= 5) { P->p += 5; return jnew(J_BOOL); }
    if (!strncmp(P->p, "null", 4) && P->end - P->p >= 4) { P->p += 4; return jnew(J_NULL); }
    if (c == '-' || (c >= '0' && c <= '9')) {
        while (P->p < P->end && (strchr("+-0123456789.eE", *P->p) != NULL)) P->p++;
        return jnew(J_NUM);
    }
    P->err = 1; return NULL;
}
static jval *jget(const jval *o, const char *key) {
    if (!o || o->k != J_OBJ) return NULL;
    for (int i = o->n - 1; i >= 0; i--) if (!strcmp(o->keys[i], key)) return o->items[i]; /* a repeated key: the last one wins (serde_json Map::insert) */
    return NULL;
}

/* ------------------------------------------------------------------------- */
/* Avro schema (the subset of apache_avro::Schema the path matches on)        */
/* ------------------------------------------------------------------------- */
typedef enum {
    A_NULL, A_BOOL, A_INT, A_LONG, A_FLOAT, A_DOUBLE, A_STRING,
    A_DATE, A_TS_MILLIS, A_TS_MICROS, A_ENUM, A_RECORD, A_UNION, A_ARRAY, A_MAP,
    A_UNSUPPORTED /* bytes, fixed, decimal, uuid, duration, time-*, local-timestamp-*, Ref */
} akind;

typedef struct asch {
    akind k;
    int n;               /* record fields / union variants / enum symbols */
    struct asch **sub;   /* record field schemas / union variants; [0] = items/values for array/map */
    char **syms;         /* enum symbols */
} asch;

static asch *anew(akind k) { asch *a = xmalloc(sizeof *a); memset(a, 0, sizeof *a); a->k = k; return a; }
static void afree(asch *a) {
    if (!a) return;
    int nsub = (a->k == A_ARRAY || a->k == A_MAP) ? 1 : (a->k == A_ENUM ? 0 : a->n);
    for (int i = 0; i < nsub; i++) afree(a->sub[i]);
    if (a->syms) for (int i = 0; i < a->n; i++) free(a->syms[i]);
    free(a->sub); free(a->syms); free(a);
}

static asch *aparse(const jval *j, int *err);

static asch *aprim(const char *t, const jval *obj, int *err) {
    const jval *lt = obj ? jget(obj, "logicalType") : NULL;
    const char *l = (lt && lt->k == J_STR) ? lt->s : NULL;
    if (!strcmp(t, "null")) return anew(A_NULL);
    if (!strcmp(t, "boolean")) return anew(A_BOOL);
    if (!strcmp(t, "float")) return anew(A_FLOAT);
    if (!strcmp(t, "double")) return anew(A_DOUBLE);
    if (!strcmp(t, "int")) {
        if (l && !strcmp(l, "date")) return anew(A_DATE);
        if (l && !strcmp(l, "time-millis")) return anew(A_UNSUPPORTED);
        return anew(A_INT);
    }
    if (!strcmp(t, "long")) {
        if (l && !strcmp(l, "timestamp-millis")) return anew(A_TS_MILLIS);
        if (l && !strcmp(l, "timestamp-micros")) return anew(A_TS_MICROS);
        if (l && (!strcmp(l, "time-micros") || !strcmp(l, "timestamp-nanos") || !strcmp(l, "local-timestamp-millis") ||
                  !strcmp(l, "local-timestamp-micros") || !strcmp(l, "local-timestamp-nanos")))
            return anew(A_UNSUPPORTED);
        return anew(A_LONG);
    }
    if (!strcmp(t, "string")) {
        if (l && !strcmp(l, "uuid")) return anew(A_UNSUPPORTED);
        return anew(A_STRING);
    }
    if (!strcmp(t, "bytes") || !strcmp(t, "fixed")) return anew(A_UNSUPPORTED);
    (void)err;
    return anew(A_UNSUPPORTED); /* named reference -> Schema::Ref, rejected by the gate (fast_decode.rs:59) */
}

static asch *aparse(const jval *j, int *err) {
    if (!j) { *err = 1; return NULL; }
    if (j->k == J_STR) return aprim(j->s, NULL, err);
    if (j->k == J_ARR) {
        asch *u = anew(A_UNION);
        u->n = j->n; u->sub = xmalloc(sizeof(asch *) * (size_t)(j->n ? j->n : 1));
        for (int i = 0; i < j->n; i++) {
            u->sub[i] = aparse(j->items[i], err);
            if (u->sub[i] && u->sub[i]->k == A_UNION) *err = 1; /* unions may not immediately nest */
        }
        return u;
    }
    if (j->k != J_OBJ) { *err = 1; return NULL; }
    const jval *t = jget(j, "type");
    if (!t) { *err = 1; return NULL; }
    if (t->k != J_STR) return aparse(t, err); /* {"type": {...}} / {"type": [...]} */
    if (!strcmp(t->s, "record") || !strcmp(t->s, "error")) {
        const jval *fs = jget(j, "fields");
        if (!fs || fs->k != J_ARR) { *err = 1; return NULL; }
        asch *r = anew(A_RECORD);
        r->n = fs->n; r->sub = xmalloc(sizeof(asch *) * (size_t)(fs->n ? fs->n : 1));
        for (int i = 0; i < fs->n; i++) {
            /* apache-avro 0.21 RecordField::parse -> Parser::parse_complex(field): with a bare-string "type" the
               type's attributes (items / values / symbols / logicalType) are read from the FIELD object
               (ruhvro/src/serialize.rs:185 relies on it); a bare "record" is a named look-up -> unsupported Ref */
            const jval *f = fs->items[i];
            const jval *ft = f && f->k == J_OBJ ? jget(f, "type") : NULL;
            if (ft && ft->k == J_STR && strcmp(ft->s, "record") && strcmp(ft->s, "error")) r->sub[i] = aparse(f, err);
            else r->sub[i] = aparse(ft, err);
        }
        return r;
    }
    if (!strcmp(t->s, "enum")) {
        const jval *sy = jget(j, "symbols");
        if (!sy || sy->k != J_ARR) { *err = 1; return NULL; }
        asch *e = anew(A_ENUM);
        e->n = sy->n; e->syms = xmalloc(sizeof(char *) * (size_t)(sy->n ? sy->n : 1));
        for (int i = 0; i < sy->n; i++) {
            if (sy->items[i]->k != J_STR) { *err = 1; e->syms[i] = xstrdup(""); } else e->syms[i] = xstrdup(sy->items[i]->s);
        }
        return e;
    }
    if (!strcmp(t->s, "array") || !strcmp(t->s, "map")) {
        asch *a = anew(!strcmp(t->s, "array") ? A_ARRAY : A_MAP);
        a->n = 1; a->sub = xmalloc(sizeof(asch *));
        a->sub[0] = aparse(jget(j, a->k == A_ARRAY ? "items" : "values"), err);
        return a;
    }
    return aprim(t->s, j, err);
}

/* fast_decode.rs:38-61 */
static int is_supported_inner(const asch *a) {
    switch (a->k) {
    case A_INT: case A_LONG: case A_FLOAT: case A_DOUBLE: case A_BOOL: case A_STRING: case A_NULL:
    case A_DATE: case A_TS_MILLIS: case A_TS_MICROS: case A_ENUM: return 1;
    case A_RECORD: case A_UNION:
        for (int i = 0; i < a->n; i++) if (!is_supported_inner(a->sub[i])) return 0;
        return 1;
    case A_ARRAY: case A_MAP: return is_supported_inner(a->sub[0]);
    default: return 0;
    }
}
static int is_supported(const asch *a) { return a->k == A_RECORD && is_supported_inner(a); }

/* ------------------------------------------------------------------------- */
/* arrow-rs builder semantics                                                 */
/* ------------------------------------------------------------------------- */
typedef struct { uint8_t *p; size_t len, cap; } bytebuf;
static void bb_reserve(bytebuf *b, size_t extra) {
    if (b->len + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap * 2 : 64;
    while (nc < b->len + extra) nc *= 2;
    b->p = xrealloc(b->p, nc); b->cap = nc;
}
static void bb_push(bytebuf *b, const void *src, size_t n) { bb_reserve(b, n); if (n) memcpy(b->p + b->len, src, n); b->len += n; }

/* BooleanBufferBuilder: LSB-first packed bits, zero padded to a whole byte. */
typedef struct { bytebuf bytes; size_t nbits; } bitbuf;
static void bit_append(bitbuf *b, int v) {
    if ((b->nbits & 7) == 0) { uint8_t z = 0; bb_push(&b->bytes, &z, 1); }
    if (v) b->bytes.p[b->nbits >> 3] |= (uint8_t)(1u << (b->nbits & 7));
    b->nbits++;
}
/* NullBufferBuilder: no bitmap until the first null, then back-filled with ones. */
typedef struct { bitbuf bits; int materialized; size_t len, nulls; } nullbuilder;
static void nb_append(nullbuilder *n, int valid) {
    if (!valid && !n->materialized) {
        n->materialized = 1;
        for (size_t i = 0; i < n->len; i++) bit_append(&n->bits, 1);
    }
    if (n->materialized) bit_append(&n->bits, valid);
    n->len++; if (!valid) n->nulls++;
}

typedef enum {
    D_INT, D_LONG, D_FLOAT, D_DOUBLE, D_BOOL, D_STRING, D_DATE, D_TS_MILLIS, D_TS_MICROS, D_ENUM,
    D_NULL, D_RECORD, D_UNION, D_LIST, D_MAP
} dkind;

/* One FieldDecoder (fast_decode.rs:73-120).  The ten inlined Nullable* variants and
 * NullableRecord/List/Map are represented by `nullable` + `null_first` on the same node. */
typedef struct dec {
    dkind k;
    int nullable, null_first;
    /* primitive / string / enum builders */
    bytebuf values;        /* fixed-width values, or string data */
    bitbuf bools;          /* BooleanBuilder values */
    bytebuf offsets;       /* i32 offsets (string / list / map) */
    nullbuilder nulls;     /* lazy validity (primitive / string / bool / enum) */
    bitbuf explicit_nulls; /* BooleanBufferBuilder of nullable record / list / map */
    size_t len;            /* rows appended (Null, Record) */
    int32_t cur_offset;    /* list / map */
    char **syms; int nsyms;
    bytebuf type_ids;      /* union */
    int nchild;
    struct dec **child;    /* record fields / union variants / list: [inner] / map: [keys, values] */
} dec;

static __thread size_t g_cap = 0; /* rows in the chunk: the reference pre-sizes builders with it (fast_decode.rs:178-194,824) */
static dec *dnew(dkind k) {
    dec *d = xmalloc(sizeof *d); memset(d, 0, sizeof *d); d->k = k;
    switch (k) {
    case D_INT: case D_DATE: case D_FLOAT: bb_reserve(&d->values, g_cap * 4); break;
    case D_LONG: case D_TS_MILLIS: case D_TS_MICROS: case D_DOUBLE: bb_reserve(&d->values, g_cap * 8); break;
    case D_BOOL: bb_reserve(&d->bools.bytes, g_cap / 8 + 1); break;
    case D_STRING: case D_ENUM: bb_reserve(&d->values, g_cap * 16); bb_reserve(&d->offsets, (g_cap + 1) * 4); break;
    case D_LIST: case D_MAP: bb_reserve(&d->offsets, (g_cap + 1) * 4); break;
    case D_UNION: bb_reserve(&d->type_ids, g_cap); break;
    default: break;
    }
    return d;
}
static void dfree(dec *d) {
    if (!d) return;
    for (int i = 0; i < d->nchild; i++) dfree(d->child[i]);
    free(d->child); free(d->values.p); free(d->bools.bytes.p); free(d->offsets.p);
    free(d->nulls.bits.bytes.p); free(d->explicit_nulls.bytes.p); free(d->type_ids.p); free(d);
}
static void push_i32(bytebuf *b, int32_t v) { bb_push(b, &v, 4); }

static dec *make_decoder(const asch *a);

/* fast_decode.rs:342-370 */
static dec *make_record_decoder(const asch *rs, int nullable) {
    dec *d = dnew(D_RECORD);
    d->nullable = nullable;
    d->nchild = rs->n; d->child = xmalloc(sizeof(dec *) * (size_t)(rs->n ? rs->n : 1));
    for (int i = 0; i < rs->n; i++) d->child[i] = make_decoder(rs->sub[i]);
    return d;
}
/* fast_decode.rs:216-235 */
static dec *make_list_decoder(const asch *items) {
    dec *d = dnew(D_LIST);
    d->nchild = 1; d->child = xmalloc(sizeof(dec *)); d->child[0] = make_decoder(items);
    push_i32(&d->offsets, 0);
    return d;
}
/* fast_decode.rs:237-268 */
static dec *make_map_decoder(const asch *values) {
    dec *d = dnew(D_MAP);
    d->nchild = 2; d->child = xmalloc(sizeof(dec *) * 2);
    d->child[0] = dnew(D_STRING); push_i32(&d->child[0]->offsets, 0); /* dedicated key StringBuilder */
    d->child[1] = make_decoder(values);
    push_i32(&d->offsets, 0);
    return d;
}
/* fast_decode.rs:176-214, 270-340, 372-414 */
static dec *make_decoder(const asch *a) {
    dec *d;
    switch (a->k) {
    case A_INT: return dnew(D_INT);
    case A_LONG: return dnew(D_LONG);
    case A_FLOAT: return dnew(D_FLOAT);
    case A_DOUBLE: return dnew(D_DOUBLE);
    case A_BOOL: return dnew(D_BOOL);
    case A_STRING: d = dnew(D_STRING); push_i32(&d->offsets, 0); return d;
    case A_DATE: return dnew(D_DATE);
    case A_TS_MILLIS: return dnew(D_TS_MILLIS);
    case A_TS_MICROS: return dnew(D_TS_MICROS);
    case A_ENUM: d = dnew(D_ENUM); push_i32(&d->offsets, 0); d->syms = a->syms; d->nsyms = a->n; return d;
    case A_NULL: return dnew(D_NULL);
    case A_RECORD: return make_record_decoder(a, 0);
    case A_ARRAY: return make_list_decoder(a->sub[0]);
    case A_MAP: return make_map_decoder(a->sub[0]);
    case A_UNION: {
        /* split_null_union (fast_decode.rs:404-414) */
        if (a->n == 2 && (a->sub[0]->k == A_NULL || a->sub[1]->k == A_NULL)) {
            int null_first = a->sub[0]->k == A_NULL;
            const asch *inner = null_first ? a->sub[1] : a->sub[0];
            if (inner->k == A_NULL || inner->k == A_UNION) return NULL; /* "unsupported nullable inner type" */
            if (inner->k == A_RECORD) d = make_record_decoder(inner, 1);
            else d = make_decoder(inner);
            if (!d) return NULL;
            d->nullable = 1; d->null_first = null_first;
            return d;
        }
        d = dnew(D_UNION);
        d->nchild = a->n; d->child = xmalloc(sizeof(dec *) * (size_t)(a->n ? a->n : 1));
        for (int i = 0; i < a->n; i++) {
            d->child[i] = make_decoder(a->sub[i]);
            if (!d->child[i]) { d->nchild = i; dfree(d); return NULL; }
        }
        return d;
    }
    default: return NULL;
    }
}

/* error codes (categories of fast_decode.rs bail!/anyhow! sites) */
enum {
    E_OK = 0,
    E_EOF = 1,          /* :849 "unexpected end of buffer" (+ :874, :884, :910) */
    E_VARINT = 2,       /* :866 "zigzag varint too long" */
    E_BOOL = 3,         /* :898 "invalid boolean byte" */
    E_NEG_LEN = 4,      /* :906 "negative string length" */
    E_BRANCH = 5,       /* :591 / :646 union branch index invalid / out of range */
    E_ENUM = 6,         /* :575 enum index out of range */
    E_SCHEMA = 7,       /* schema parse error / unsupported / zero-field record (:633-635) */
    E_OVERFLOW = 8      /* i32 offset overflow (arrow-rs panics; surfaced as an error here) */
};

typedef struct { const uint8_t *p, *end; int err; } cursor;

/* fast_decode.rs:845-869 */
static int64_t read_zigzag_long(cursor *c) {
    uint64_t result = 0; unsigned shift = 0;
    for (;;) {
        if (c->p >= c->end) { c->err = E_EOF; return 0; }
        uint8_t byte = *c->p++;
        result |= (uint64_t)(byte & 0x7F) << shift;
        if ((byte & 0x80) == 0) return (int64_t)(result >> 1) ^ -(int64_t)(result & 1);
        shift += 7;
        if (shift >= 64) { c->err = E_VARINT; return 0; }
    }
}
/* fast_decode.rs:585-593: returns 1 for Value, 0 for Null */
static int union_branch(cursor *c, int null_first) {
    int64_t idx = read_zigzag_long(c);
    if (c->err) return 0;
    if (idx == 0) return null_first ? 0 : 1;
    if (idx == 1) return null_first ? 1 : 0;
    c->err = E_BRANCH; return 0;
}

static void append_null(dec *d);
static void decode(dec *d, cursor *c);

static void append_fixed(dec *d, const void *v, size_t w) { bb_push(&d->values, v, w); nb_append(&d->nulls, 1); }
static void append_string(dec *d, const uint8_t *s, size_t n) {
    bb_push(&d->values, s, n);
    if (d->values.len > (size_t)INT32_MAX) return; /* caller checks */
    push_i32(&d->offsets, (int32_t)d->values.len);
    nb_append(&d->nulls, 1);
}

/* fast_decode.rs:902-922 */
static int read_string(cursor *c, const uint8_t **s, size_t *n) {
    int64_t len = read_zigzag_long(c);
    if (c->err) return 0;
    if (len < 0) { c->err = E_NEG_LEN; return 0; }
    if ((uint64_t)(c->end - c->p) < (uint64_t)len) { c->err = E_EOF; return 0; }
    *s = c->p; *n = (size_t)len; c->p += len;
    return 1;
}

/* value part of FieldDecoder::decode (fast_decode.rs:420-499), after any null-union branch */
static void decode_value(dec *d, cursor *c) {
    switch (d->k) {
    case D_INT: case D_DATE: { int64_t v = read_zigzag_long(c); if (c->err) return; int32_t t = (int32_t)v; append_fixed(d, &t, 4); break; }
    case D_LONG: case D_TS_MILLIS: case D_TS_MICROS: { int64_t v = read_zigzag_long(c); if (c->err) return; append_fixed(d, &v, 8); break; }
    case D_FLOAT: /* :871-879 */
        if (c->end - c->p < 4) { c->err = E_EOF; return; }
        append_fixed(d, c->p, 4); c->p += 4; break;
    case D_DOUBLE: /* :881-891 */
        if (c->end - c->p < 8) { c->err = E_EOF; return; }
        append_fixed(d, c->p, 8); c->p += 8; break;
    case D_BOOL: { /* :893-900 */
        if (c->p >= c->end) { c->err = E_EOF; return; }
        uint8_t b = *c->p++;
        if (b > 1) { c->err = E_BOOL; return; }
        bit_append(&d->bools, b); nb_append(&d->nulls, 1); break;
    }
    case D_STRING: {
        const uint8_t *s; size_t n;
        if (!read_string(c, &s, &n)) return;
        append_string(d, s, n);
        if (d->values.len > (size_t)INT32_MAX) c->err = E_OVERFLOW;
        break;
    }
    case D_ENUM: { /* append_enum :570-578 */
        int64_t idx = read_zigzag_long(c); if (c->err) return;
        if ((uint64_t)idx >= (uint64_t)d->nsyms) { c->err = E_ENUM; return; }
        append_string(d, (const uint8_t *)d->syms[idx], strlen(d->syms[idx]));
        if (d->values.len > (size_t)INT32_MAX) c->err = E_OVERFLOW;
        break;
    }
    case D_NULL: d->len++; break; /* :480 */
    case D_RECORD: /* decode_present :597-606 */
        if (d->nullable) bit_append(&d->explicit_nulls, 1);
        d->len++;
        for (int i = 0; i < d->nchild; i++) { decode(d->child[i], c); if (c->err) return; }
        break;
    case D_UNION: { /* :643-658 */
        int64_t idx = read_zigzag_long(c); if (c->err) return;
        if (idx < 0 || idx >= d->nchild) { c->err = E_BRANCH; return; }
        for (int i = 0; i < d->nchild; i++) {
            if (i == idx) { decode(d->child[i], c); if (c->err) return; } else append_null(d->child[i]);
        }
        int8_t t = (int8_t)idx; bb_push(&d->type_ids, &t, 1);
        break;
    }
    case D_LIST: case D_MAP: /* :703-719, :745-762 */
        for (;;) {
            int64_t n = read_zigzag_long(c); if (c->err) return; /* read_block_count :689-700 */
            if (n < 0) { (void)read_zigzag_long(c); if (c->err) return; n = (int64_t)(0 - (uint64_t)n); } /* `-n` wraps in the release build: i64::MIN stays negative and `0..n` is empty */
            if (n == 0) break;
            for (int64_t i = 0; i < n; i++) {
                if (d->k == D_MAP) {
                    const uint8_t *s; size_t sl;
                    if (!read_string(c, &s, &sl)) return;
                    append_string(d->child[0], s, sl);
                    if (d->child[0]->values.len > (size_t)INT32_MAX) { c->err = E_OVERFLOW; return; }
                    decode(d->child[1], c);
                } else decode(d->child[0], c);
                if (c->err) return;
                if (d->cur_offset == INT32_MAX) { c->err = E_OVERFLOW; return; }
                d->cur_offset++;
            }
        }
        push_i32(&d->offsets, d->cur_offset);
        if (d->nullable) bit_append(&d->explicit_nulls, 1);
        break;
    }
}

/* FieldDecoder::decode (fast_decode.rs:420-499) */
static void decode(dec *d, cursor *c) {
    if (d->nullable) {
        int is_value = union_branch(c, d->null_first);
        if (c->err) return;
        if (!is_value) { append_null(d); return; }
    }
    decode_value(d, c);
}

/* FieldDecoder::append_null (fast_decode.rs:503-534) and the per-type append_null impls */
static void append_null(dec *d) {
    static const uint8_t zeros[8] = {0};
    switch (d->k) {
    case D_INT: case D_DATE: case D_FLOAT: bb_push(&d->values, zeros, 4); nb_append(&d->nulls, 0); break;
    case D_LONG: case D_TS_MILLIS: case D_TS_MICROS: case D_DOUBLE: bb_push(&d->values, zeros, 8); nb_append(&d->nulls, 0); break;
    case D_BOOL: bit_append(&d->bools, 0); nb_append(&d->nulls, 0); break;
    case D_STRING: case D_ENUM: push_i32(&d->offsets, (int32_t)d->values.len); nb_append(&d->nulls, 0); break;
    case D_NULL: d->len++; break;
    case D_RECORD: /* :608-616 */
        if (d->nullable) bit_append(&d->explicit_nulls, 0);
        d->len++;
        for (int i = 0; i < d->nchild; i++) append_null(d->child[i]);
        break;
    case D_UNION: { /* :660-668 */
        for (int i = 0; i < d->nchild; i++) append_null(d->child[i]);
        int8_t t = 0; bb_push(&d->type_ids, &t, 1);
        break;
    }
    case D_LIST: case D_MAP: /* :721-727, :764-770 — children untouched */
        push_i32(&d->offsets, d->cur_offset);
        if (d->nullable) bit_append(&d->explicit_nulls, 0);
        break;
    }
}

/* ------------------------------------------------------------------------- */
/* finish -> flat pre-order array descriptors                                 */
/* ------------------------------------------------------------------------- */
typedef struct {
    int32_t kind;        /* dkind; the map's entries struct is reported as D_RECORD */
    int32_t n_children;
    int64_t length;
    int64_t null_count;
    int32_t has_validity;
    int32_t n_buffers;   /* data buffers after validity: 0..2 */
    const uint8_t *validity; int64_t validity_bytes;
    const uint8_t *buf0; int64_t buf0_bytes;   /* values / offsets / type_ids */
    const uint8_t *buf1; int64_t buf1_bytes;   /* string data */
} orc_array;

typedef struct {
    int err; int64_t err_record;
    int n_arrays, cap;
    orc_array *arrays;
    dec *root;
    int64_t n_rows;
} orc_batch;

static orc_array *emit(orc_batch *b) {
    if (b->n_arrays == b->cap) { b->cap = b->cap ? b->cap * 2 : 32; b->arrays = xrealloc(b->arrays, sizeof(orc_array) * (size_t)b->cap); }
    orc_array *a = &b->arrays[b->n_arrays++]; memset(a, 0, sizeof *a); return a;
}
static size_t popcount_zero(const bitbuf *b) {
    size_t ones = 0, full = b->nbits >> 3;
    for (size_t i = 0; i < full; i++) ones += (size_t)__builtin_popcount(b->bytes.p[i]);
    for (size_t i = full << 3; i < b->nbits; i++) ones += (b->bytes.p[i >> 3] >> (i & 7)) & 1;
    return b->nbits - ones;
}
static void set_lazy_validity(orc_array *a, const dec *d) {
    if (d->nulls.materialized) { a->has_validity = 1; a->validity = d->nulls.bits.bytes.p; a->validity_bytes = (int64_t)d->nulls.bits.bytes.len; a->null_count = (int64_t)d->nulls.nulls; }
}
static void set_explicit_validity(orc_array *a, const dec *d) {
    if (d->nullable) { a->has_validity = 1; a->validity = d->explicit_nulls.bytes.p; a->validity_bytes = (int64_t)d->explicit_nulls.bytes.len; a->null_count = (int64_t)popcount_zero(&d->explicit_nulls); }
}
/* FieldDecoder::finish (fast_decode.rs:536-567) + Record/Union/List/Map finish */
static int finish(orc_batch *b, dec *d) {
    orc_array *a = emit(b);
    a->kind = d->k;
    switch (d->k) {
    case D_INT: case D_DATE: case D_FLOAT: case D_LONG: case D_TS_MILLIS: case D_TS_MICROS: case D_DOUBLE: {
        size_t w = (d->k == D_INT || d->k == D_DATE || d->k == D_FLOAT) ? 4 : 8;
        a->length = (int64_t)(d->values.len / w); a->n_buffers = 1;
        a->buf0 = d->values.p; a->buf0_bytes = (int64_t)d->values.len; set_lazy_validity(a, d); break;
    }
    case D_BOOL: a->length = (int64_t)d->bools.nbits; a->n_buffers = 1; a->buf0 = d->bools.bytes.p; a->buf0_bytes = (int64_t)d->bools.bytes.len; set_lazy_validity(a, d); break;
    case D_STRING: case D_ENUM:
        a->length = (int64_t)(d->offsets.len / 4) - 1; a->n_buffers = 2;
        a->buf0 = d->offsets.p; a->buf0_bytes = (int64_t)d->offsets.len; a->buf1 = d->values.p; a->buf1_bytes = (int64_t)d->values.len;
        set_lazy_validity(a, d); break;
    case D_NULL: a->length = (int64_t)d->len; a->null_count = (int64_t)d->len; break; /* NullArray::new(len) :558 */
    case D_RECORD:
        if (d->nchild == 0) return E_SCHEMA; /* :633-635 */
        a->length = (int64_t)d->len; a->n_children = d->nchild; set_explicit_validity(a, d);
        for (int i = 0; i < d->nchild; i++) { int e = finish(b, d->child[i]); if (e) return e; }
        break;
    case D_UNION:
        a->length = (int64_t)d->type_ids.len; a->n_buffers = 1; a->buf0 = d->type_ids.p; a->buf0_bytes = (int64_t)d->type_ids.len; a->n_children = d->nchild;
        for (int i = 0; i < d->nchild; i++) { int e = finish(b, d->child[i]); if (e) return e; }
        break;
    case D_LIST:
        a->length = (int64_t)(d->offsets.len / 4) - 1; a->n_buffers = 1; a->buf0 = d->offsets.p; a->buf0_bytes = (int64_t)d->offsets.len; a->n_children = 1;
        set_explicit_validity(a, d);
        { int e = finish(b, d->child[0]); if (e) return e; }
        break;
    case D_MAP: {
        a->length = (int64_t)(d->offsets.len / 4) - 1; a->n_buffers = 1; a->buf0 = d->offsets.p; a->buf0_bytes = (int64_t)d->offsets.len; a->n_children = 1;
        set_explicit_validity(a, d);
        orc_array *en = emit(b); /* entries StructArray::try_new(fields,[keys,values],None) :784-788 */
        en->kind = D_RECORD; en->n_children = 2; en->length = (int64_t)(d->child[0]->offsets.len / 4) - 1;
        int e = finish(b, d->child[0]); if (e) return e;
        e = finish(b, d->child[1]); if (e) return e;
        break;
    }
    }
    return 0;
}

/* ------------------------------------------------------------------------- */
/* public C API (ctypes)                                                      */
/* ------------------------------------------------------------------------- */
typedef struct { asch *schema; int parse_err; } orc_schema;

orc_schema *orc_schema_parse(const char *json, int64_t len) {
    jparser P = { json, json + len, 0 };
    jval *j = jparse(&P);
    if (!P.err) { jskip(&P); if (P.p != P.end) P.err = 1; }
    orc_schema *s = xmalloc(sizeof *s); s->schema = NULL; s->parse_err = 0;
    if (P.err || !j) { s->parse_err = 1; jfree(j); return s; }
    int err = 0;
    s->schema = aparse(j, &err);
    if (err || !s->schema) s->parse_err = 1;
    jfree(j);
    return s;
}
void orc_schema_free(orc_schema *s) { if (s) { afree(s->schema); free(s); } }
int orc_schema_ok(const orc_schema *s) { return s && !s->parse_err; }
int orc_schema_is_supported(const orc_schema *s) { return s && !s->parse_err && s->schema && is_supported(s->schema); }

void orc_batch_free(orc_batch *b) { if (b) { dfree(b->root); free(b->arrays); free(b); } }

static int has_null_child(const dec *d) {
    for (int i = 0; i < d->nchild; i++) { if (!d->child[i] || has_null_child(d->child[i])) return 1; }
    return 0;
}
/* A record with no fields always fails at finish: nested -> "RecordDecoder produced a record with 0 fields"
   (fast_decode.rs:633-635); top level -> RecordBatch::try_new with no columns (:834). */
static int has_empty_record(const asch *a) {
    if (!a) return 0;
    if (a->k == A_RECORD && a->n == 0) return 1;
    if (a->k == A_RECORD || a->k == A_UNION || a->k == A_ARRAY || a->k == A_MAP)
        for (int i = 0; i < a->n; i++) if (has_empty_record(a->sub[i])) return 1;
    return 0;
}
/* decode_with_arrow_schema (fast_decode.rs:815-835) over rows [r0, r1) of the packed input */
static orc_batch *decode_range(const orc_schema *s, const uint8_t *data, const int64_t *offsets, int64_t r0, int64_t r1) {
    orc_batch *b = xmalloc(sizeof *b); memset(b, 0, sizeof *b); b->err_record = -1;
    if (!orc_schema_is_supported(s) || has_empty_record(s->schema)) { b->err = E_SCHEMA; return b; }
    g_cap = (size_t)(r1 - r0);
    dec *top = make_record_decoder(s->schema, 0);
    if (has_null_child(top)) { dfree(top); b->err = E_SCHEMA; return b; } /* "unsupported nullable inner type" :338 */
    b->root = top; b->n_rows = r1 - r0;
    for (int64_t r = r0; r < r1; r++) {
        cursor c = { data + offsets[r], data + offsets[r + 1], 0 };
        /* top.decode_present: top-level record, no validity, trailing bytes ignored */
        top->len++;
        for (int i = 0; i < top->nchild; i++) { decode(top->child[i], &c); if (c.err) break; }
        if (c.err) { b->err = c.err; b->err_record = r; return b; }
    }
    for (int i = 0; i < top->nchild; i++) { int e = finish(b, top->child[i]); if (e) { b->err = e; return b; } }
    return b;
}

orc_batch *orc_decode(const orc_schema *s, const uint8_t *data, const int64_t *offsets, int64_t n) {
    return decode_range(s, data, offsets, 0, n);
}
int orc_batch_error(const orc_batch *b) { return b->err; }
int64_t orc_batch_error_record(const orc_batch *b) { return b->err_record; }
int orc_batch_n_arrays(const orc_batch *b) { return b->n_arrays; }
int64_t orc_batch_n_rows(const orc_batch *b) { return b->n_rows; }
const orc_array *orc_batch_array(const orc_batch *b, int i) { return &b->arrays[i]; }

/* clamp_chunks + build_slices (deserialize.rs:53-68) */
int64_t orc_clamp_chunks(int64_t num_chunks, int64_t n) {
    int64_t k = num_chunks < 1 ? 1 : num_chunks;
    int64_t m = n < 1 ? 1 : n;
    return k < m ? k : m;
}
void orc_chunk_bounds(int64_t n, int64_t k, int64_t i, int64_t *r0, int64_t *r1) {
    int64_t cs = n / k;
    *r0 = i * cs; *r1 = (i == k - 1) ? n : (i + 1) * cs;
}

/* Benchmark aid: orc_set_pin(1) pins worker t of orc_decode_threaded to the t-th CPU of the process's affinity mask
 * (steadier timings of the CPU arm on shared boxes).  Off by default. */
static int g_pin = 0;
void orc_set_pin(int on) { g_pin = on; }

typedef struct { const orc_schema *s; const uint8_t *data; const int64_t *offsets; int64_t n, k; int64_t next; pthread_mutex_t mu; orc_batch **out;
                 int next_thread; cpu_set_t allowed; int n_allowed; } job;
static void *worker(void *arg) {
    job *j = arg;
    if (g_pin && j->n_allowed > 0) {
        pthread_mutex_lock(&j->mu); int me = j->next_thread++; pthread_mutex_unlock(&j->mu);
        int want = me % j->n_allowed, seen = 0;
        for (int c = 0; c < CPU_SETSIZE; c++) {
            if (!CPU_ISSET(c, &j->allowed)) continue;
            if (seen++ == want) { cpu_set_t one; CPU_ZERO(&one); CPU_SET(c, &one); pthread_setaffinity_np(pthread_self(), sizeof one, &one); break; }
        }
    }
    for (;;) {
        pthread_mutex_lock(&j->mu); int64_t i = j->next++; pthread_mutex_unlock(&j->mu);
        if (i >= j->k) return NULL;
        int64_t r0, r1; orc_chunk_bounds(j->n, j->k, i, &r0, &r1);
        j->out[i] = decode_range(j->s, j->data, j->offsets, r0, r1);
    }
}
/* per_datum_deserialize_threaded (deserialize.rs:76-121): k chunks decoded on `threads` workers,
 * results in chunk order.  `out` must have room for orc_clamp_chunks(num_chunks, n) pointers. */
int64_t orc_decode_threaded(const orc_schema *s, const uint8_t *data, const int64_t *offsets, int64_t n,
                            int64_t num_chunks, int threads, orc_batch **out) {
    int64_t k = orc_clamp_chunks(num_chunks, n);
    job j;
    memset(&j, 0, sizeof j);
    j.s = s; j.data = data; j.offsets = offsets; j.n = n; j.k = k; j.out = out;
    pthread_mutex_init(&j.mu, NULL);
    CPU_ZERO(&j.allowed);
    j.n_allowed = 0;
    if (g_pin && sched_getaffinity(0, sizeof j.allowed, &j.allowed) == 0) j.n_allowed = CPU_COUNT(&j.allowed);
    if (threads < 1) threads = 1;
    if (threads > k) threads = (int)k;
    pthread_t *th = xmalloc(sizeof(pthread_t) * (size_t)threads);
    for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, worker, &j);
    for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
    free(th);
    return k;
}
