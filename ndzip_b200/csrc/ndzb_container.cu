// ndzb_container.cu — the sharded stream container behind the C ABI (include/ndzip_b200.h, ndzb_container_*).
//
// New work (the reference has one stream per array and one GPU). A multi-GPU pipeline that
// compresses to storage and decompresses again on several GPUs never needs the global stream — neither the cross-rank
// offset exchange nor the gather to one root. The container keeps every rank's SELF-CONTAINED ndzip stream of its slab
// (what ndzb_dist_compress / ndzb_compress produce for the slab; the reference decoder reads it with the slab's
// extent) behind a small segment table:
//
//   u32 magic "NDZS" | u32 version | u32 dtype (0 f32, 1 f64) | u32 dims | u32 size[3] | u32 segments
//   per segment: u32 slab_begin | u32 slab_end (dimension 0) | u64 stream_words | u64 byte_offset
//   the segments, each starting at a multiple of 16 bytes
//
// Writing needs the other ranks' stream LENGTHS only (one all-gather of an integer: ndzb_dist_gathered_lengths has
// them already), every rank then writes its own segment with pwrite(); reading needs nothing: a rank takes the
// segments whose slabs it owns, in any world size. ndzb_container_to_global_stream converts to the reference's single
// stream (header entries rebased as in reference src/ndzip/common.hh:342-358) when one is wanted.
//
// Host code only (no kernels); ndzb_container_decompress_segment goes through ndzb_offload_decompress.
#include "../../include/ndzip_b200.h"

#include <cerrno>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <sys/stat.h>
#include <unistd.h>
#include <vector>

namespace {

constexpr uint32_t kMagic = 0x535A444Eu;  // "NDZS", little endian
constexpr uint32_t kVersion = 1;
constexpr uint32_t kFixedWords = 8;    // magic, version, dtype, dims, size[3], segments
constexpr uint32_t kSegmentWords = 6;  // begin, end, words (lo, hi), offset (lo, hi)
constexpr uint32_t kMaxSegments = 1u << 20;

uint64_t align16(uint64_t n) { return (n + 15) / 16 * 16; }
uint64_t word_bytes(int dtype) { return dtype == NDZB_F32 ? 4 : 8; }
bool valid_profile(int dtype, int dims) { return (dtype == NDZB_F32 || dtype == NDZB_F64) && dims >= 1 && dims <= 3; }

struct fd_guard {
    int fd;
    ~fd_guard() {
        if (fd >= 0) close(fd);
    }
};

bool full_pwrite(int fd, const void *buf, uint64_t bytes, uint64_t offset) {
    const char *p = static_cast<const char *>(buf);
    while (bytes > 0) {
        const ssize_t n = pwrite(fd, p, bytes < (1ull << 30) ? bytes : (1ull << 30), static_cast<off_t>(offset));
        if (n < 0 && errno == EINTR) continue;
        if (n <= 0) return false;
        p += n;
        offset += static_cast<uint64_t>(n);
        bytes -= static_cast<uint64_t>(n);
    }
    return true;
}

bool full_pread(int fd, void *buf, uint64_t bytes, uint64_t offset) {
    char *p = static_cast<char *>(buf);
    while (bytes > 0) {
        const ssize_t n = pread(fd, p, bytes < (1ull << 30) ? bytes : (1ull << 30), static_cast<off_t>(offset));
        if (n < 0 && errno == EINTR) continue;
        if (n <= 0) return false;  // error or end of file inside the range
        p += n;
        offset += static_cast<uint64_t>(n);
        bytes -= static_cast<uint64_t>(n);
    }
    return true;
}

void slab_size(const ndzb_container_info &info, const ndzb_container_segment &seg, uint32_t out[3]) {
    out[0] = seg.slab_end - seg.slab_begin;
    out[1] = info.dims > 1 ? info.size[1] : 0u;
    out[2] = info.dims > 2 ? info.size[2] : 0u;
}

}  // namespace

extern "C" {

uint64_t ndzb_container_header_bytes(uint32_t segments) {
    return align16(4ull * (kFixedWords + static_cast<uint64_t>(kSegmentWords) * segments));
}

int ndzb_container_plan(int dtype, int dims, const uint32_t *global_size, uint32_t segments, const uint64_t *stream_words,
        ndzb_container_info *info, ndzb_container_segment *out_segments) {
    if (!valid_profile(dtype, dims) || !global_size || !info || segments == 0 || segments > kMaxSegments
            || (!stream_words && out_segments) || (stream_words && !out_segments)) {
        return NDZB_ERR_INVALID_ARGUMENT;
    }
    memset(info, 0, sizeof *info);
    info->dtype = dtype;
    info->dims = dims;
    for (int d = 0; d < dims; ++d) info->size[d] = global_size[d];
    info->segments = segments;
    info->header_bytes = ndzb_container_header_bytes(segments);
    uint64_t offset = info->header_bytes;
    if (out_segments) {
        for (uint32_t r = 0; r < segments; ++r) {
            ndzb_dist_layout l;
            if (int rc = ndzb_dist_plan(dtype, dims, global_size, static_cast<int>(segments), static_cast<int>(r), &l)) return rc;
            out_segments[r].slab_begin = l.slab_begin;
            out_segments[r].slab_end = l.slab_end;
            out_segments[r].stream_words = stream_words[r];
            out_segments[r].byte_offset = offset;
            // the end of the last segment is the end of the container (no padding behind it)
            info->total_bytes = offset + stream_words[r] * word_bytes(dtype);
            offset = align16(info->total_bytes);
        }
    } else {
        info->total_bytes = info->header_bytes;
    }
    return NDZB_OK;
}

int ndzb_container_encode_header(const ndzb_container_info *info, const ndzb_container_segment *segments, void *out, uint64_t out_bytes) {
    if (!info || !segments || !out || !valid_profile(info->dtype, info->dims) || info->segments == 0 || info->segments > kMaxSegments) {
        return NDZB_ERR_INVALID_ARGUMENT;
    }
    const uint64_t need = ndzb_container_header_bytes(info->segments);
    if (out_bytes < need) return NDZB_ERR_CAPACITY;
    memset(out, 0, need);
    uint32_t *w = static_cast<uint32_t *>(out);
    w[0] = kMagic;
    w[1] = kVersion;
    w[2] = static_cast<uint32_t>(info->dtype);
    w[3] = static_cast<uint32_t>(info->dims);
    for (int d = 0; d < 3; ++d) w[4 + d] = d < info->dims ? info->size[d] : 0u;
    w[7] = info->segments;
    for (uint32_t i = 0; i < info->segments; ++i) {
        uint32_t *s = w + kFixedWords + static_cast<size_t>(kSegmentWords) * i;
        s[0] = segments[i].slab_begin;
        s[1] = segments[i].slab_end;
        s[2] = static_cast<uint32_t>(segments[i].stream_words);
        s[3] = static_cast<uint32_t>(segments[i].stream_words >> 32);
        s[4] = static_cast<uint32_t>(segments[i].byte_offset);
        s[5] = static_cast<uint32_t>(segments[i].byte_offset >> 32);
    }
    return NDZB_OK;
}

int ndzb_container_decode_header(const void *buf, uint64_t bytes, ndzb_container_info *info, ndzb_container_segment *out_segments,
        uint32_t max_segments) {
    if (!buf || !info) return NDZB_ERR_INVALID_ARGUMENT;
    memset(info, 0, sizeof *info);
    if (bytes < 4ull * kFixedWords) return NDZB_ERR_CORRUPT_STREAM;  // too short to be a container
    uint32_t fixed[kFixedWords];
    memcpy(fixed, buf, sizeof fixed);  // the buffer need not be aligned
    if (fixed[0] != kMagic || fixed[1] != kVersion) return NDZB_ERR_CORRUPT_STREAM;
    const int dtype = static_cast<int>(fixed[2]), dims = static_cast<int>(fixed[3]);
    const uint32_t count = fixed[7];
    if (!valid_profile(dtype, dims) || count == 0 || count > kMaxSegments) return NDZB_ERR_CORRUPT_STREAM;
    info->dtype = dtype;
    info->dims = dims;
    for (int d = 0; d < dims; ++d) info->size[d] = fixed[4 + d];
    info->segments = count;
    info->header_bytes = ndzb_container_header_bytes(count);
    const uint64_t table_end = 4ull * (kFixedWords + static_cast<uint64_t>(kSegmentWords) * count);
    if (bytes < table_end) return NDZB_ERR_CORRUPT_STREAM;  // truncated table
    const unsigned char *table = static_cast<const unsigned char *>(buf) + 4ull * kFixedWords;
    uint32_t end_of_previous = 0;
    uint64_t end_bytes = info->header_bytes;
    for (uint32_t i = 0; i < count; ++i) {
        uint32_t s[kSegmentWords];
        memcpy(s, table + 4ull * kSegmentWords * i, sizeof s);
        ndzb_container_segment seg;
        seg.slab_begin = s[0];
        seg.slab_end = s[1];
        seg.stream_words = s[2] | (static_cast<uint64_t>(s[3]) << 32);
        seg.byte_offset = s[4] | (static_cast<uint64_t>(s[5]) << 32);
        // slabs tile dimension 0 in order; segments are 16-byte aligned, in order, behind the table, without overlap
        if (seg.slab_begin != end_of_previous || seg.slab_end < seg.slab_begin || seg.byte_offset % 16 != 0
                || seg.byte_offset < end_bytes || seg.byte_offset > (1ull << 56) || seg.stream_words > (1ull << 40)) {
            return NDZB_ERR_CORRUPT_STREAM;
        }
        end_of_previous = seg.slab_end;
        end_bytes = seg.byte_offset + seg.stream_words * word_bytes(dtype);
        if (out_segments && i < max_segments) out_segments[i] = seg;
    }
    if (end_of_previous != info->size[0]) return NDZB_ERR_CORRUPT_STREAM;  // the segments do not cover the grid
    info->total_bytes = end_bytes;
    if (out_segments && max_segments < count) return NDZB_ERR_CAPACITY;  // info->segments says how many there are
    return NDZB_OK;
}

int ndzb_container_to_global_stream(const void *container, uint64_t bytes, void *out_stream, uint64_t capacity_words, uint64_t *out_words) {
    if (!container || !out_words) return NDZB_ERR_INVALID_ARGUMENT;
    ndzb_container_info info;
    if (int rc = ndzb_container_decode_header(container, bytes, &info, nullptr, 0)) return rc;
    std::vector<ndzb_container_segment> segs(info.segments);
    if (int rc = ndzb_container_decode_header(container, bytes, &info, segs.data(), info.segments)) return rc;
    if (bytes < info.total_bytes) return NDZB_ERR_CORRUPT_STREAM;  // truncated
    const uint64_t wb = word_bytes(info.dtype);
    const int world = static_cast<int>(info.segments);
    std::vector<ndzb_dist_layout> lay(info.segments);
    std::vector<uint64_t> cube_words(info.segments, 0), cube_base(info.segments, 0);
    const unsigned char *base = static_cast<const unsigned char *>(container);
    uint64_t total_cubes_words = 0;
    for (int r = 0; r < world; ++r) {
        if (int rc = ndzb_dist_plan(info.dtype, info.dims, info.size, world, r, &lay[r])) return rc;
        // only containers whose slabs are those of the library's own partition concatenate into the global stream
        if (lay[r].slab_begin != segs[r].slab_begin || lay[r].slab_end != segs[r].slab_end) return NDZB_ERR_INVALID_ARGUMENT;
        if (segs[r].stream_words < lay[r].local_header_words) return NDZB_ERR_CORRUPT_STREAM;
        if (lay[r].local_cubes) {
            uint32_t last;
            memcpy(&last, base + segs[r].byte_offset + 4ull * (lay[r].local_cubes - 1), sizeof last);
            cube_words[r] = last;  // "offset_after" of the slab's last cube = its compressed words
        }
        if (segs[r].stream_words != lay[r].local_header_words + cube_words[r] + lay[r].local_border_words) return NDZB_ERR_CORRUPT_STREAM;
        cube_base[r] = total_cubes_words;
        total_cubes_words += cube_words[r];
    }
    const uint64_t ghdr = lay[0].global_header_words;
    const uint64_t total = ghdr + total_cubes_words + lay[0].global_border_words;
    *out_words = total;
    if (total >= (1ull << 32)) return NDZB_ERR_INVALID_ARGUMENT;  // the reference's index_type is uint32
    if (!out_stream) return NDZB_OK;                             // length query
    if (capacity_words < total) return NDZB_ERR_CAPACITY;
    unsigned char *out = static_cast<unsigned char *>(out_stream);
    memset(out, 0, ghdr * wb);  // includes the padding word of an odd double header (cuda_codec.inl:446-452)
    for (int r = 0; r < world; ++r) {
        const unsigned char *seg = base + segs[r].byte_offset;
        const uint32_t rebase = static_cast<uint32_t>(cube_base[r]);
        uint32_t previous = 0;
        for (uint32_t i = 0; i < lay[r].local_cubes; ++i) {
            uint32_t v;
            memcpy(&v, seg + 4ull * i, sizeof v);
            if (v < previous || v > cube_words[r]) return NDZB_ERR_CORRUPT_STREAM;
            previous = v;
            v += rebase;
            memcpy(out + 4ull * (lay[r].cube_index_base + i), &v, sizeof v);
        }
        memcpy(out + (ghdr + cube_base[r]) * wb, seg + lay[r].local_header_words * wb, cube_words[r] * wb);
        memcpy(out + (ghdr + total_cubes_words + lay[r].border_base) * wb, seg + (lay[r].local_header_words + cube_words[r]) * wb,
                lay[r].local_border_words * wb);
    }
    return NDZB_OK;
}

int ndzb_container_create_file(const char *path, const ndzb_container_info *info, const ndzb_container_segment *segments) {
    if (!path || !info || !segments) return NDZB_ERR_INVALID_ARGUMENT;
    std::vector<unsigned char> header(ndzb_container_header_bytes(info->segments));
    if (int rc = ndzb_container_encode_header(info, segments, header.data(), header.size())) return rc;
    fd_guard f{open(path, O_CREAT | O_TRUNC | O_WRONLY, 0644)};
    if (f.fd < 0) return NDZB_ERR_IO;
    if (!full_pwrite(f.fd, header.data(), header.size(), 0)) return NDZB_ERR_IO;
    const ndzb_container_segment &last = segments[info->segments - 1];
    const uint64_t total = last.byte_offset + last.stream_words * word_bytes(info->dtype);
    if (ftruncate(f.fd, static_cast<off_t>(total)) != 0) return NDZB_ERR_IO;
    return NDZB_OK;
}

int ndzb_container_write_segment(const char *path, int dtype, const ndzb_container_segment *segment, const void *h_stream) {
    if (!path || !segment || (dtype != NDZB_F32 && dtype != NDZB_F64) || (!h_stream && segment->stream_words)) return NDZB_ERR_INVALID_ARGUMENT;
    fd_guard f{open(path, O_WRONLY)};
    if (f.fd < 0) return NDZB_ERR_IO;
    return full_pwrite(f.fd, h_stream, segment->stream_words * word_bytes(dtype), segment->byte_offset) ? NDZB_OK : NDZB_ERR_IO;
}

int ndzb_container_read_header(const char *path, ndzb_container_info *info, ndzb_container_segment *out_segments, uint32_t max_segments) {
    if (!path || !info) return NDZB_ERR_INVALID_ARGUMENT;
    fd_guard f{open(path, O_RDONLY)};
    if (f.fd < 0) return NDZB_ERR_IO;
    struct stat st;
    if (fstat(f.fd, &st) != 0) return NDZB_ERR_IO;
    const uint64_t file_bytes = static_cast<uint64_t>(st.st_size);
    uint32_t fixed[kFixedWords] = {0};
    if (file_bytes < sizeof fixed || !full_pread(f.fd, fixed, sizeof fixed, 0)) return NDZB_ERR_CORRUPT_STREAM;
    // the count comes from the file: never read more table than the file can hold
    const uint64_t table_end = 4ull * (kFixedWords + static_cast<uint64_t>(kSegmentWords) * fixed[7]);
    if (fixed[0] != kMagic || fixed[7] > kMaxSegments || table_end > file_bytes) return NDZB_ERR_CORRUPT_STREAM;
    std::vector<unsigned char> head(table_end);
    if (!full_pread(f.fd, head.data(), table_end, 0)) return NDZB_ERR_IO;
    if (int rc = ndzb_container_decode_header(head.data(), table_end, info, out_segments, max_segments)) return rc;
    if (info->total_bytes > file_bytes) return NDZB_ERR_CORRUPT_STREAM;  // truncated file
    return NDZB_OK;
}

int ndzb_container_read_segment(const char *path, int dtype, const ndzb_container_segment *segment, void *h_out) {
    if (!path || !segment || (dtype != NDZB_F32 && dtype != NDZB_F64) || (!h_out && segment->stream_words)) return NDZB_ERR_INVALID_ARGUMENT;
    fd_guard f{open(path, O_RDONLY)};
    if (f.fd < 0) return NDZB_ERR_IO;
    return full_pread(f.fd, h_out, segment->stream_words * word_bytes(dtype), segment->byte_offset) ? NDZB_OK : NDZB_ERR_CORRUPT_STREAM;
}

int ndzb_container_decompress_segment(ndzb_ctx *ctx, const void *container, uint64_t bytes, uint32_t index, void *h_slab, uint64_t *kernel_ns) {
    if (!ctx || !container || !h_slab) return NDZB_ERR_INVALID_ARGUMENT;
    ndzb_container_info info;
    if (int rc = ndzb_container_decode_header(container, bytes, &info, nullptr, 0)) return rc;
    if (index >= info.segments) return NDZB_ERR_INVALID_ARGUMENT;
    std::vector<ndzb_container_segment> segs(info.segments);
    if (int rc = ndzb_container_decode_header(container, bytes, &info, segs.data(), info.segments)) return rc;
    const ndzb_container_segment &seg = segs[index];
    if (seg.byte_offset + seg.stream_words * word_bytes(info.dtype) > bytes || seg.stream_words >= (1ull << 32)) return NDZB_ERR_CORRUPT_STREAM;
    uint32_t slab[3];
    slab_size(info, seg, slab);
    uint32_t consumed = 0;
    // the offloader validates the slab stream's own header against the segment length before touching the device
    return ndzb_offload_decompress(ctx, static_cast<const unsigned char *>(container) + seg.byte_offset, static_cast<uint32_t>(seg.stream_words),
            h_slab, info.dims, slab, &consumed, kernel_ns);
}

}  // extern "C"
