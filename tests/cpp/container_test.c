/* The sharded stream container through the C ABI alone, from plain C (include/ndzip_b200.h must be a C header).
 * Host only: the slab streams come from the CPU oracle (oracle/ndzip_oracle.c). For every profile: plan, create the
 * file, write the segments in reverse rank order (no rank needs another's data), read them back with a different
 * reader count, convert to the single stream and compare it word for word with the oracle's stream of the whole grid. */
#include <ndzip_b200.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

uint32_t ndzo_compress(int dtype, int dims, const uint32_t *size, const void *data, void *stream); /* oracle/ndzip_oracle.c */

#define CHECK(x) do { int rc_ = (x); if (rc_ != NDZB_OK) { printf("FAIL %s:%d %s -> %s\n", __FILE__, __LINE__, #x, ndzb_strerror(rc_)); return 1; } } while (0)
#define EXPECT(c) do { if (!(c)) { printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); return 1; } } while (0)

static uint64_t elements(int dims, const uint32_t *size) {
    uint64_t n = 1;
    for (int d = 0; d < dims; ++d) n *= size[d];
    return n;
}

static int run(const char *path, int dtype, int dims, const uint32_t *size, uint32_t world) {
    const uint64_t wb = dtype == NDZB_F32 ? 4 : 8, n = elements(dims, size);
    unsigned char *data = malloc(n * wb);
    uint32_t seed = 12345u + (uint32_t) dims * 7u + (uint32_t) dtype;
    for (uint64_t i = 0; i < n * wb / 4; ++i) {  /* smooth-ish words: low bits noisy, high bits slowly varying */
        seed = seed * 1664525u + 1013904223u;
        ((uint32_t *) data)[i] = 0x3f800000u + (uint32_t) (i / 64) * 8u + (seed >> 26);
    }
    void *whole = calloc(ndzb_compressed_length_bound(dtype, dims, size), wb);
    const uint64_t whole_words = ndzo_compress(dtype, dims, size, data, whole);

    ndzb_container_info info;
    ndzb_container_segment *segs = calloc(world, sizeof *segs);
    uint64_t *words = calloc(world, sizeof *words);
    void **streams = calloc(world, sizeof *streams);
    uint64_t row = n / size[0];
    for (uint32_t r = 0; r < world; ++r) {
        ndzb_dist_layout l;
        CHECK(ndzb_dist_plan(dtype, dims, size, (int) world, (int) r, &l));
        streams[r] = calloc(l.local_bound_words ? l.local_bound_words : 1, wb);
        words[r] = ndzo_compress(dtype, dims, l.slab_size, data + (uint64_t) l.slab_begin * row * wb, streams[r]);
    }
    CHECK(ndzb_container_plan(dtype, dims, size, world, words, &info, segs));
    EXPECT(info.header_bytes == ndzb_container_header_bytes(world) && info.segments == world);
    CHECK(ndzb_container_create_file(path, &info, segs));
    for (uint32_t r = world; r-- > 0;) CHECK(ndzb_container_write_segment(path, dtype, &segs[r], streams[r]));

    /* reader: table first, then segment by segment */
    ndzb_container_info got;
    CHECK(ndzb_container_read_header(path, &got, NULL, 0));
    EXPECT(got.segments == world && got.total_bytes == info.total_bytes && got.dtype == dtype && got.dims == dims);
    ndzb_container_segment *rsegs = calloc(world, sizeof *rsegs);
    EXPECT(ndzb_container_read_header(path, &got, rsegs, world - 1) == NDZB_ERR_CAPACITY);
    CHECK(ndzb_container_read_header(path, &got, rsegs, world));
    unsigned char *blob = calloc(got.total_bytes, 1);
    unsigned char *head = malloc(got.header_bytes);
    CHECK(ndzb_container_encode_header(&got, rsegs, head, got.header_bytes));
    memcpy(blob, head, got.header_bytes);
    for (uint32_t r = 0; r < world; ++r) {
        EXPECT(rsegs[r].stream_words == words[r] && rsegs[r].byte_offset % 16 == 0);
        CHECK(ndzb_container_read_segment(path, dtype, &rsegs[r], blob + rsegs[r].byte_offset));
        EXPECT(memcmp(blob + rsegs[r].byte_offset, streams[r], words[r] * wb) == 0);
    }
    /* the single stream of the whole grid */
    uint64_t out_words = 0;
    CHECK(ndzb_container_to_global_stream(blob, got.total_bytes, NULL, 0, &out_words));
    EXPECT(out_words == whole_words);
    void *global = malloc(out_words * wb);
    EXPECT(ndzb_container_to_global_stream(blob, got.total_bytes, global, out_words - 1, &out_words) == NDZB_ERR_CAPACITY);
    CHECK(ndzb_container_to_global_stream(blob, got.total_bytes, global, out_words, &out_words));
    EXPECT(memcmp(global, whole, whole_words * wb) == 0);
    /* damage */
    EXPECT(ndzb_container_to_global_stream(blob, got.total_bytes - 1, NULL, 0, &out_words) == NDZB_ERR_CORRUPT_STREAM);
    blob[0] ^= 1;
    EXPECT(ndzb_container_decode_header(blob, got.total_bytes, &got, NULL, 0) == NDZB_ERR_CORRUPT_STREAM);
    printf("%s %dD %u segments: %llu bytes, global stream %llu words identical to the oracle's\n", dtype == NDZB_F32 ? "f32" : "f64", dims, world,
            (unsigned long long) info.total_bytes, (unsigned long long) whole_words);
    for (uint32_t r = 0; r < world; ++r) free(streams[r]);
    free(streams); free(words); free(segs); free(rsegs); free(blob); free(head); free(global); free(whole); free(data);
    return 0;
}

int main(int argc, char **argv) {
    const char *path = argc > 1 ? argv[1] : "/tmp/ndzb_container_test.ndzs";
    const uint32_t s1[3] = {4096 * 5 + 77, 0, 0}, s2[3] = {64 * 5 + 3, 130, 0}, s3[3] = {16 * 4 + 5, 35, 40};
    for (int dtype = 0; dtype < 2; ++dtype) {
        for (uint32_t world = 1; world <= 4; ++world) {
            if (run(path, dtype, 1, s1, world) || run(path, dtype, 2, s2, world) || run(path, dtype, 3, s3, world)) return 1;
        }
    }
    EXPECT(ndzb_container_read_header("/nonexistent/dir/x.ndzs", &(ndzb_container_info){0}, NULL, 0) == NDZB_ERR_IO);
    remove(path);
    printf("PASS\n");
    return 0;
}
