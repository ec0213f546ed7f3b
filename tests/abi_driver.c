/* TEST INFRASTRUCTURE — a caller of the C ABI with no Python and no C++: what an FFI binding (the Rust shim of
 * INTEGRATION.md, cgo, JNI) does.  Parses a schema, decodes three of the reference's literal datums
 * (ruhvro/src/lib.rs:165-167) through rv_decode_host into two batches, exports them through the Arrow C Data Interface
 * and prints what it finds; exit status 0 = everything as expected.
 *
 *   gcc -std=c11 -I include tests/abi_driver.c -L pyruhvro_b200 -lruhvro_b200 -Wl,-rpath,$PWD/pyruhvro_b200 -o abi_driver
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "ruhvro_b200.h"

/* Arrow C Data Interface (https://arrow.apache.org/docs/format/CDataInterface.html) */
struct ArrowSchema {
    const char* format; const char* name; const char* metadata; int64_t flags; int64_t n_children;
    struct ArrowSchema** children; struct ArrowSchema* dictionary; void (*release)(struct ArrowSchema*); void* private_data;
};
struct ArrowArray {
    int64_t length; int64_t null_count; int64_t offset; int64_t n_buffers; int64_t n_children; const void** buffers;
    struct ArrowArray** children; struct ArrowArray* dictionary; void (*release)(struct ArrowArray*); void* private_data;
};

static int hexval(int c) { return c <= '9' ? c - '0' : (c | 32) - 'a' + 10; }

#define CHECK(cond, what) do { if (!(cond)) { fprintf(stderr, "FAIL %s: %s\n", what, rv_last_error()); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 5) { fprintf(stderr, "usage: abi_driver <schema.json> <hex datum> <hex datum> <hex datum>\n"); return 2; }
    FILE* f = fopen(argv[1], "rb");
    if (!f) { perror("schema"); return 2; }
    static char json[1 << 16];
    const size_t jl = fread(json, 1, sizeof json, f);
    fclose(f);

    static uint8_t data[1 << 16];
    int64_t offsets[4] = {0, 0, 0, 0};
    for (int r = 0; r < 3; ++r) {
        const char* h = argv[2 + r];
        const size_t n = strlen(h) / 2;
        for (size_t i = 0; i < n; ++i) data[offsets[r] + (int64_t)i] = (uint8_t)(hexval(h[2 * i]) * 16 + hexval(h[2 * i + 1]));
        offsets[r + 1] = offsets[r] + (int64_t)n;
    }

    rv_schema* s = NULL;
    CHECK(rv_schema_parse(json, jl, &s) == RV_OK, "rv_schema_parse");
    CHECK(rv_schema_is_supported(s) == 1, "rv_schema_is_supported");
    struct ArrowSchema sch;
    CHECK(rv_schema_export_arrow(s, &sch) == RV_OK, "rv_schema_export_arrow");
    printf("schema: format %s, %lld columns:", sch.format, (long long)sch.n_children);
    for (int64_t i = 0; i < sch.n_children; ++i) printf(" %s(%s)", sch.children[i]->name, sch.children[i]->format);
    printf("\n");

    rv_result* res = NULL;
    CHECK(rv_decode_host(s, data, offsets, 3, 2, &res) == RV_OK, "rv_decode_host");
    CHECK(rv_result_num_batches(res) == 2, "two batches");                       /* clamp_chunks / build_slices: rows 1 + 2 */
    CHECK(rv_result_num_rows(res, 0) == 1 && rv_result_num_rows(res, 1) == 2, "chunk bounds");
    int64_t rows = 0;
    for (int64_t b = 0; b < 2; ++b) {
        struct ArrowArray arr;
        CHECK(rv_result_export(res, b, &arr, NULL) == RV_OK, "rv_result_export");
        CHECK(arr.n_children == sch.n_children && arr.release != NULL, "struct array of the columns");
        printf("batch %lld: %lld rows;", (long long)b, (long long)arr.length);
        for (int64_t i = 0; i < arr.n_children; ++i) printf(" %s nulls=%lld", sch.children[i]->name, (long long)arr.children[i]->null_count);
        printf("\n");
        rows += arr.length;
        /* column 1 ("age", nullable int): values buffer of batch 1 holds G4's 28 then G5's null slot (0) */
        if (b == 1) {
            const int32_t* age = (const int32_t*)arr.children[1]->buffers[1];
            CHECK(age[0] == 28 && age[1] == 0 && arr.children[1]->null_count == 1, "age values");
        }
        arr.release(&arr);
    }
    rv_result_free(res);      /* the exported arrays kept the memory alive until their release */
    CHECK(rows == 3, "row total");

    /* a truncated datum is a data error with the reference's category and the record index */
    int64_t bad_off[2] = {0, offsets[1] - 40};
    CHECK(rv_decode_host(s, data, bad_off, 1, 1, &res) == RV_ERR_EOF, "RV_ERR_EOF on a truncated datum");
    printf("truncated datum -> %s\n", rv_last_error());
    sch.release(&sch);
    rv_schema_release(s);
    printf("abi_driver ok\n");
    return 0;
}
