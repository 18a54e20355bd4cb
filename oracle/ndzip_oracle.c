/* TEST INFRASTRUCTURE — CPU oracle for the ndzip hot path. NOT part of the product.
 *
 * A plain, scalar C restatement of the reference algorithm (celerity/ndzip @ ff4e6702):
 * per-hypercube integer Lorenzo transform, B x B bit-plane transpose with zero-plane removal,
 * stream layout (offset header, cubes, raw border) and the inverses. Each function cites the
 * reference file:line it follows. Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * leg may load this library, and only as the checker.
 *
 * Parity status: PINNED. tests/test_oracle.py checks this file against
 *   (1) the golden table of SURVEY.md §8(c) (stream length, CRC32, first words; tests/golden/),
 *   (2) the reference's own known-answer test for border slices (src/test/codec_generic_test.cc:102-111),
 *   (3) the unmodified reference CPU codec built into oracle/_ref/libndzip_ref.so (whole streams,
 *       single-cube transform / transpose / encode primitives), wherever that library is present.
 *
 * Build: make -C oracle oracle   ->  oracle/libndzip_oracle.so
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

enum { cube_elems = 4096 }; /* reference src/ndzip/common.hh:368-381: 4096^1 = 64^2 = 16^3 */

/* reference src/ndzip/common.hh:371-378 */
static uint32_t cube_side(int dims) {
    return dims == 1 ? 4096u : dims == 2 ? 64u : 16u;
}

/* Geometry of one array: sizes slowest dimension first (reference include/ndzip/ndzip.hh:35-160). */
typedef struct {
    int dims;
    uint32_t side;
    uint64_t size[3];     /* elements per dimension */
    uint64_t stride[3];   /* linear-index stride per dimension (row-major, ndzip.hh:172-180) */
    uint64_t inner[3];    /* size rounded down to whole cubes */
    uint32_t cubes[3];    /* cubes per dimension */
    uint32_t num_cubes;   /* reference src/ndzip/common.hh:395-402 */
    uint64_t num_elems;
    uint64_t num_border;  /* reference src/ndzip/common.hh:308-317 */
} layout_t;

static int layout_init(layout_t *g, int dims, const uint32_t *size) {
    if (dims < 1 || dims > 3) return -1;
    memset(g, 0, sizeof *g);
    g->dims = dims;
    g->side = cube_side(dims);
    uint64_t stride = 1, inner_elems = 1;
    g->num_cubes = 1;
    for (int d = dims - 1; d >= 0; --d) {
        g->size[d] = size[d];
        g->stride[d] = stride;
        stride *= size[d];
        g->cubes[d] = size[d] / g->side;
        g->inner[d] = (uint64_t) g->cubes[d] * g->side;
        g->num_cubes *= g->cubes[d];
        inner_elems *= g->inner[d];
    }
    g->num_elems = stride;
    g->num_border = g->num_elems - inner_elems;
    return 0;
}

static uint64_t linear_index(const layout_t *g, const uint64_t *pos) {
    uint64_t l = 0;
    for (int d = 0; d < g->dims; ++d) l += pos[d] * g->stride[d];
    return l;
}

/* hc_index -> element coordinates of the cube's first element. Cubes are numbered row-major over
 * cube coordinates, slowest dimension first (reference src/ndzip/common.hh:414-433, 570-579). */
static void cube_origin(const layout_t *g, uint32_t hc_index, uint64_t *origin) {
    for (int d = g->dims - 1; d >= 0; --d) {
        origin[d] = (uint64_t) (hc_index % g->cubes[d]) * g->side;
        hc_index /= g->cubes[d];
    }
}

/* Border = everything outside whole cubes, emitted as contiguous (offset, count) slices in
 * ascending linear-index order (reference src/ndzip/common.hh:245-282). If any dimension is
 * shorter than one cube the whole array is a single slice (common.hh:271-276). */
typedef struct {
    uint64_t *pairs; /* offset, count, offset, count, ... */
    size_t n, cap, cursor;
} border_iter_t;

static void border_push(border_iter_t *it, uint64_t off, uint64_t cnt) {
    if (it->n == it->cap) {
        it->cap = it->cap ? it->cap * 2 : 64;
        it->pairs = (uint64_t *) realloc(it->pairs, it->cap * 2 * sizeof(uint64_t));
    }
    it->pairs[2 * it->n] = off;
    it->pairs[2 * it->n + 1] = cnt;
    it->n++;
}

/* `base` is the linear offset of the current position with dimensions >= d at coordinate 0.
 * `deepest` is the fastest-varying dimension that has a border: shallower dimensions enumerate
 * their whole-cube coordinates before emitting their own tail slab. */
static void border_collect(border_iter_t *it, const layout_t *g, int d, int deepest, uint64_t base) {
    if (d < deepest) {
        for (uint64_t p = 0; p < g->inner[d]; ++p) border_collect(it, g, d + 1, deepest, base + p * g->stride[d]);
    }
    if (g->inner[d] < g->size[d]) {
        border_push(it, base + g->inner[d] * g->stride[d], (g->size[d] - g->inner[d]) * g->stride[d]);
    }
}

static void border_begin(border_iter_t *it, const layout_t *g) {
    memset(it, 0, sizeof *it);
    int deepest = -1;
    for (int d = 0; d < g->dims; ++d) {
        if (g->cubes[d] == 0) {
            if (g->num_elems) border_push(it, 0, g->num_elems);
            return;
        }
        if (g->inner[d] != g->size[d]) deepest = d;
    }
    if (deepest >= 0) border_collect(it, g, 0, deepest, 0);
}

static int border_next(border_iter_t *it, uint64_t *off, uint64_t *cnt) {
    if (it->cursor == it->n) {
        free(it->pairs);
        memset(it, 0, sizeof *it);
        return 0;
    }
    *off = it->pairs[2 * it->cursor];
    *cnt = it->pairs[2 * it->cursor + 1];
    it->cursor++;
    return 1;
}

#define WORD uint32_t
#define WBITS 32u
#define FN(name) name##_u32
#include "ndzip_oracle_impl.h"
#undef WORD
#undef WBITS
#undef FN

#define WORD uint64_t
#define WBITS 64u
#define FN(name) name##_u64
#include "ndzip_oracle_impl.h"
#undef WORD
#undef WBITS
#undef FN

/* ------------------------------------------------------------------ exported C ABI (ctypes) */
/* dtype: 0 = float / uint32 words, 1 = double / uint64 words. Data is passed as raw IEEE bits. */

uint32_t ndzo_num_hypercubes(int dims, const uint32_t *size) {
    layout_t g;
    if (layout_init(&g, dims, size)) return 0;
    return g.num_cubes;
}

uint64_t ndzo_border_element_count(int dims, const uint32_t *size) {
    layout_t g;
    if (layout_init(&g, dims, size)) return 0;
    return g.num_border;
}

/* reference src/ndzip/common.cc:31-42: header + H * (4096/B * (B+1)) + border */
uint64_t ndzo_compressed_length_bound(int dtype, int dims, const uint32_t *size) {
    layout_t g;
    if (layout_init(&g, dims, size)) return 0;
    const uint32_t bits = dtype == 0 ? 32 : 64;
    const uint64_t hdr = dtype == 0 ? header_words_u32(g.num_cubes) : header_words_u64(g.num_cubes);
    const uint64_t per_cube = (uint64_t) cube_elems / bits * (bits + 1);
    return hdr + (uint64_t) g.num_cubes * per_cube + g.num_border;
}

uint32_t ndzo_compress(int dtype, int dims, const uint32_t *size, const void *data, void *stream) {
    layout_t g;
    if (layout_init(&g, dims, size)) return 0;
    return dtype == 0 ? compress_u32((const uint32_t *) data, &g, (uint32_t *) stream)
                      : compress_u64((const uint64_t *) data, &g, (uint64_t *) stream);
}

uint32_t ndzo_decompress(int dtype, int dims, const uint32_t *size, const void *stream, void *data) {
    layout_t g;
    if (layout_init(&g, dims, size)) return 0;
    return dtype == 0 ? decompress_u32((const uint32_t *) stream, &g, (uint32_t *) data)
                      : decompress_u64((const uint64_t *) stream, &g, (uint64_t *) data);
}

void ndzo_block_transform(int dtype, int dims, void *cube) {
    if (dtype == 0) block_transform_u32((uint32_t *) cube, dims);
    else block_transform_u64((uint64_t *) cube, dims);
}

void ndzo_inverse_block_transform(int dtype, int dims, void *cube) {
    if (dtype == 0) inverse_block_transform_u32((uint32_t *) cube, dims);
    else inverse_block_transform_u64((uint64_t *) cube, dims);
}

void ndzo_transpose_bits(int dtype, const void *in, void *out) {
    if (dtype == 0) transpose_bits_u32((const uint32_t *) in, (uint32_t *) out);
    else transpose_bits_u64((const uint64_t *) in, (uint64_t *) out);
}

/* returns words written */
uint32_t ndzo_zero_bit_encode(int dtype, const void *cube, void *out) {
    return dtype == 0 ? zero_bit_encode_u32((const uint32_t *) cube, (uint32_t *) out)
                      : zero_bit_encode_u64((const uint64_t *) cube, (uint64_t *) out);
}

/* returns words consumed */
uint32_t ndzo_zero_bit_decode(int dtype, const void *in, void *cube) {
    return dtype == 0 ? zero_bit_decode_u32((const uint32_t *) in, (uint32_t *) cube)
                      : zero_bit_decode_u64((const uint64_t *) in, (uint64_t *) cube);
}

/* Gather one cube of raw bits in cube-local order (pins "flattening of hypercubes",
 * reference src/test/codec_profile_test.inl:514-549). */
void ndzo_load_cube(int dtype, int dims, const uint32_t *size, const void *data, uint32_t hc_index, void *cube) {
    layout_t g;
    if (layout_init(&g, dims, size)) return;
    if (dtype == 0) load_cube_u32((const uint32_t *) data, &g, hc_index, (uint32_t *) cube);
    else load_cube_u64((const uint64_t *) data, &g, hc_index, (uint64_t *) cube);
}

/* Border slices with an explicit cube side (the reference's known-answer test uses sides 2/4/5).
 * Writes up to max_pairs (offset,count) pairs; returns the total number of slices. */
int ndzo_border_slices(int dims, const uint32_t *size, uint32_t side, uint64_t *out_pairs, int max_pairs) {
    layout_t g;
    if (layout_init(&g, dims, size)) return -1;
    /* re-derive the cube grid for the requested side */
    g.side = side;
    for (int d = 0; d < dims; ++d) {
        g.cubes[d] = (uint32_t) (g.size[d] / side);
        g.inner[d] = (uint64_t) g.cubes[d] * side;
    }
    border_iter_t it;
    border_begin(&it, &g);
    int n = 0;
    uint64_t off, cnt;
    while (border_next(&it, &off, &cnt)) {
        if (n < max_pairs) {
            out_pairs[2 * n] = off;
            out_pairs[2 * n + 1] = cnt;
        }
        ++n;
    }
    return n;
}
