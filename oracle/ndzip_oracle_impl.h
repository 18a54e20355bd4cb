/* TEST INFRASTRUCTURE — width-generic body of the CPU oracle. Included twice by ndzip_oracle.c with
 *   WORD   = uint32_t / uint64_t      (reference: bits_type<T>, include/ndzip/ndzip.hh:186-212)
 *   WBITS  = 32 / 64
 *   FN(x)  = x##_u32 / x##_u64
 * Every function cites the reference file:line it restates. Nothing here is copied; the reference
 * is templated C++/AVX2, this is scalar C. */

/* reference src/ndzip/common.hh:436-439 */
static inline WORD FN(rotl1)(WORD v) {
    return (WORD) ((v << 1) | (v >> (WBITS - 1)));
}

/* reference src/ndzip/common.hh:441-444 */
static inline WORD FN(rotr1)(WORD v) {
    return (WORD) ((v >> 1) | (v << (WBITS - 1)));
}

/* reference src/ndzip/common.hh:446-449: if the MSB is set, flip every bit below it */
static inline WORD FN(complement_negative)(WORD v) {
    const WORD low_mask = (WORD) (~(WORD) 0 >> 1);
    return (v >> (WBITS - 1)) ? (WORD) (v ^ low_mask) : v;
}

/* x[i] -= x[i-1] along one strided lane, back to front so each step sees the original
 * predecessor (reference src/ndzip/common.hh:451-460 does it front to back with a carry). */
static void FN(diff_lane)(WORD *x, uint32_t n, uint32_t stride) {
    for (uint32_t i = n - 1; i >= 1; --i) {
        x[(size_t) i * stride] = (WORD) (x[(size_t) i * stride] - x[(size_t) (i - 1) * stride]);
    }
}

/* inclusive prefix sum along one strided lane (reference src/ndzip/common.hh:462-467) */
static void FN(sum_lane)(WORD *x, uint32_t n, uint32_t stride) {
    for (uint32_t i = 1; i < n; ++i) {
        x[(size_t) i * stride] = (WORD) (x[(size_t) i * stride] + x[(size_t) (i - 1) * stride]);
    }
}

/* Apply `lane_fn` to every axis-`axis` lane of a side^dims cube (axis 0 = slowest). */
static void FN(for_axis)(WORD *x, int dims, uint32_t side, int axis, void (*lane_fn)(WORD *, uint32_t, uint32_t)) {
    uint32_t stride = 1;
    for (int d = dims - 1; d > axis; --d) stride *= side;
    const uint32_t total = cube_elems;
    /* a lane starts wherever the axis coordinate is 0 */
    for (uint32_t start = 0; start < total; ++start) {
        if ((start / stride) % side == 0) lane_fn(x + start, side, stride);
    }
}

/* Forward integer Lorenzo transform of one cube, in place (reference src/ndzip/common.hh:469-501):
 * rotate sign to LSB, difference along every axis (the per-axis operators commute over Z/2^k, so
 * the axis order is immaterial; the reference itself uses two different orders,
 * common.hh:484-496 vs cpu_codec.inl:215-223), then complement negatives. */
static void FN(block_transform)(WORD *x, int dims) {
    const uint32_t side = cube_side(dims);
    for (uint32_t i = 0; i < cube_elems; ++i) x[i] = FN(rotl1)(x[i]);
    for (int axis = 0; axis < dims; ++axis) FN(for_axis)(x, dims, side, axis, FN(diff_lane));
    for (uint32_t i = 0; i < cube_elems; ++i) x[i] = FN(complement_negative)(x[i]);
}

/* Inverse (reference src/ndzip/common.hh:503-535). */
static void FN(inverse_block_transform)(WORD *x, int dims) {
    const uint32_t side = cube_side(dims);
    for (uint32_t i = 0; i < cube_elems; ++i) x[i] = FN(complement_negative)(x[i]);
    for (int axis = 0; axis < dims; ++axis) FN(for_axis)(x, dims, side, axis, FN(sum_lane));
    for (uint32_t i = 0; i < cube_elems; ++i) x[i] = FN(rotr1)(x[i]);
}

/* B x B bit-matrix transpose: out[i] bit (B-1-j) = in[j] bit (B-1-i)
 * (reference src/ndzip/cpu_codec.inl:355-363). */
static void FN(transpose_bits)(const WORD *in, WORD *out) {
    for (uint32_t i = 0; i < WBITS; ++i) {
        WORD plane = 0;
        for (uint32_t j = 0; j < WBITS; ++j) {
            plane |= (WORD) (((in[j] >> (WBITS - 1 - i)) & 1u) << (WBITS - 1 - j));
        }
        out[i] = plane;
    }
}

/* Residual cube -> compressed cube: C head words, then for every chunk the non-zero bit planes,
 * MSB plane first (reference src/ndzip/cpu_codec.inl:344-352, 514-524, 541-559).
 * Returns the compressed length in words. */
static uint32_t FN(zero_bit_encode)(const WORD *cube, WORD *out) {
    const uint32_t num_chunks = cube_elems / WBITS;
    uint32_t body = num_chunks;
    for (uint32_t c = 0; c < num_chunks; ++c) {
        const WORD *chunk = cube + (size_t) c * WBITS;
        WORD head = 0;
        for (uint32_t j = 0; j < WBITS; ++j) head |= chunk[j];
        out[c] = head;
        if (head != 0) {
            WORD planes[WBITS];
            FN(transpose_bits)(chunk, planes);
            for (uint32_t i = 0; i < WBITS; ++i) {
                if (planes[i] != 0) out[body++] = planes[i];
            }
        }
    }
    return body;
}

/* Inverse of the above (reference src/ndzip/cpu_codec.inl:526-538, 561-578): plane i is present
 * exactly when bit (B-1-i) of the chunk head is set. Returns the words consumed. */
static uint32_t FN(zero_bit_decode)(const WORD *in, WORD *cube) {
    const uint32_t num_chunks = cube_elems / WBITS;
    uint32_t body = num_chunks;
    for (uint32_t c = 0; c < num_chunks; ++c) {
        WORD *chunk = cube + (size_t) c * WBITS;
        const WORD head = in[c];
        if (head == 0) {
            for (uint32_t j = 0; j < WBITS; ++j) chunk[j] = 0;
            continue;
        }
        WORD planes[WBITS];
        for (uint32_t i = 0; i < WBITS; ++i) {
            planes[i] = ((head >> (WBITS - 1 - i)) & 1u) ? in[body++] : 0;
        }
        FN(transpose_bits)(planes, chunk); /* the transpose is an involution, codec_generic_test.cc:65-81 */
    }
    return body;
}

/* Gather / scatter one cube; cube-local order is row-major, slowest dimension first
 * (reference src/ndzip/common.hh:538-568, cpu_codec.inl:74-98). */
static void FN(load_cube)(const WORD *data, const layout_t *g, uint32_t hc_index, WORD *cube) {
    uint64_t origin[3];
    cube_origin(g, hc_index, origin);
    const uint32_t s = g->side;
    const int dims = g->dims;
    for (uint32_t e = 0; e < cube_elems; ++e) {
        uint32_t rem = e;
        uint64_t pos[3] = {0, 0, 0};
        for (int d = dims - 1; d >= 0; --d) {
            pos[d] = origin[d] + rem % s;
            rem /= s;
        }
        cube[e] = data[linear_index(g, pos)];
    }
}

static void FN(store_cube)(WORD *data, const layout_t *g, uint32_t hc_index, const WORD *cube) {
    uint64_t origin[3];
    cube_origin(g, hc_index, origin);
    const uint32_t s = g->side;
    const int dims = g->dims;
    for (uint32_t e = 0; e < cube_elems; ++e) {
        uint32_t rem = e;
        uint64_t pos[3] = {0, 0, 0};
        for (int d = dims - 1; d >= 0; --d) {
            pos[d] = origin[d] + rem % s;
            rem /= s;
        }
        data[linear_index(g, pos)] = cube[e];
    }
}

/* Header area in words: H uint32 offsets, padded to a whole WORD
 * (reference src/ndzip/common.hh:350-352, common.cc:37-38). */
static uint32_t FN(header_words)(uint32_t num_cubes) {
    const uint32_t per_word = WBITS / 32;
    return (num_cubes + per_word - 1) / per_word;
}

/* Whole-array compressor (reference src/ndzip/cpu_codec.inl:597-619 and stream layout
 * common.hh:328-366). Unlike the reference CPU encoder — and like its CUDA/SYCL encoders
 * (cuda_codec.inl:446-452) — the odd-H padding word of a 64-bit header is written as 0.
 * Returns the stream length in words. */
static uint32_t FN(compress)(const WORD *data, const layout_t *g, WORD *stream) {
    const uint32_t H = g->num_cubes;
    const uint32_t hdr = FN(header_words)(H);
    uint32_t *offsets = (uint32_t *) stream;
    if (hdr * (WBITS / 32) > H) offsets[H] = 0;
    WORD cube[4096];
    uint32_t pos = 0; /* words after the header */
    for (uint32_t h = 0; h < H; ++h) {
        FN(load_cube)(data, g, h, cube);
        FN(block_transform)(cube, g->dims);
        pos += FN(zero_bit_encode)(cube, stream + hdr + pos);
        offsets[h] = pos; /* "offset_after", relative to the first cube (common.hh:342-358) */
    }
    /* border: raw bits in ascending linear index order (common.hh:245-294) */
    uint32_t nb = 0;
    border_iter_t it;
    border_begin(&it, g);
    uint64_t off, cnt;
    while (border_next(&it, &off, &cnt)) {
        memcpy(stream + hdr + pos + nb, data + off, (size_t) cnt * sizeof(WORD));
        nb += (uint32_t) cnt;
    }
    return hdr + pos + nb;
}

/* Whole-array decompressor (reference src/ndzip/cpu_codec.inl:640-659). Returns words consumed. */
static uint32_t FN(decompress)(const WORD *stream, const layout_t *g, WORD *data) {
    const uint32_t H = g->num_cubes;
    const uint32_t hdr = FN(header_words)(H);
    const uint32_t *offsets = (const uint32_t *) stream;
    WORD cube[4096];
    for (uint32_t h = 0; h < H; ++h) {
        const uint32_t begin = h ? offsets[h - 1] : 0;
        FN(zero_bit_decode)(stream + hdr + begin, cube);
        FN(inverse_block_transform)(cube, g->dims);
        FN(store_cube)(data, g, h, cube);
    }
    const uint32_t pos = H ? offsets[H - 1] : 0;
    uint32_t nb = 0;
    border_iter_t it;
    border_begin(&it, g);
    uint64_t off, cnt;
    while (border_next(&it, &off, &cnt)) {
        memcpy(data + off, stream + hdr + pos + nb, (size_t) cnt * sizeof(WORD));
        nb += (uint32_t) cnt;
    }
    return hdr + pos + nb;
}
