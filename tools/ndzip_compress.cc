// ndzip-compress — command-line front end of the B200 codec, a drop-in for the reference's `compress` tool
// (reference src/compress/compress.cc:130-229: same options, same file formats) on top of the C ABI.
//
//   raw file        = any number of arrays of `--array-size` elements, back to back (binary float / double dump)
//   compressed file = their ndzip streams, back to back (reference compress.cc:17-58)
//
// Differences from the reference tool: there is no CPU encoder in this library, so `-e cuda` is the default and
// the only target (`-e cpu` / `-e sycl` fail with the reference's "Unimplemented target" message); Boost.ProgramOptions
// is replaced by a small parser; file chunks are staged in pinned host memory so the offloader can overlap
// H2D, kernels and D2H.
#include "../include/ndzip_b200.h"

#include <cerrno>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

namespace {

struct usage_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};
struct io_error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

const char *kOptions =
        "Options:\n"
        "  --help                 show this help\n"
        "  -d [ --decompress ]    decompress (default compress)\n"
        "  -n [ --array-size ] N… array size (one value per dimension, first-major)\n"
        "  -t [ --data-type ] T   float|double (default float)\n"
        "  -e [ --target ] E      cuda (default cuda; this build has no other encoder)\n"
        "  -T [ --threads ] N     number of CPU threads (accepted for compatibility, unused)\n"
        "  -i [ --input ] FILE    input file (default '-' is stdin)\n"
        "  -o [ --output ] FILE   output file (default '-' is stdout)\n"
        "  --no-mmap              do not use memory-mapped I/O\n";

struct options {
    bool decompress = false, no_mmap = false, help = false;
    std::vector<uint32_t> size;
    int dtype = NDZB_F32;
    std::string input = "-", output = "-";
};

options parse(int argc, char **argv) {
    options o;
    auto value_of = [&](int &i, const std::string &name) -> std::string {
        if (i + 1 >= argc) throw usage_error("the required argument for option '" + name + "' is missing");
        return argv[++i];
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--help") {
            o.help = true;
        } else if (a == "-d" || a == "--decompress") {
            o.decompress = true;
        } else if (a == "--no-mmap") {
            o.no_mmap = true;
        } else if (a == "-n" || a == "--array-size") {
            // multitoken: every following argument that is a number
            while (i + 1 < argc && argv[i + 1][0] >= '0' && argv[i + 1][0] <= '9') {
                char *end = nullptr;
                errno = 0;
                const unsigned long long v = strtoull(argv[++i], &end, 10);
                if (errno || *end || v >= (1ull << 32)) throw usage_error("the argument ('" + std::string(argv[i]) + "') for option '--array-size' is invalid");
                o.size.push_back(static_cast<uint32_t>(v));
            }
        } else if (a == "-t" || a == "--data-type") {
            const std::string t = value_of(i, "--data-type");
            if (t == "float") o.dtype = NDZB_F32;
            else if (t == "double") o.dtype = NDZB_F64;
            else throw usage_error("Invalid data type " + t);
        } else if (a == "-e" || a == "--target") {
            const std::string e = value_of(i, "--target");
            if (e != "cuda") throw usage_error("Unimplemented target " + e);
        } else if (a == "-T" || a == "--threads") {
            value_of(i, "--threads");
        } else if (a == "-i" || a == "--input") {
            o.input = value_of(i, "--input");
        } else if (a == "-o" || a == "--output") {
            o.output = value_of(i, "--output");
        } else {
            throw usage_error("unrecognised option '" + a + "'");
        }
    }
    if (o.help) return o;
    if (o.size.empty()) throw usage_error("the option '--array-size' is required but missing");
    if (o.size.size() > 3) throw usage_error("Expected between 1 and 3 dimensions, got " + std::to_string(o.size.size()));
    return o;
}

void check(int status, const char *what) {
    if (status == NDZB_OK) return;
    std::string msg = std::string(what) + ": " + ndzb_strerror(status);
    if (status == NDZB_ERR_CUDA) msg += std::string(" (") + ndzb_last_cuda_error() + ")";
    throw std::runtime_error(msg);
}

// Pinned host buffer (falls back to malloc when there is no device: then the first codec call reports it).
struct host_buffer {
    void *p = nullptr;
    size_t bytes = 0;
    bool pinned = false;
    explicit host_buffer(size_t n) : bytes(n ? n : 1) {
        if (ndzb_host_alloc(&p, bytes) == NDZB_OK) {
            pinned = true;
        } else {
            p = malloc(bytes);
            if (!p) throw std::bad_alloc();
        }
    }
    ~host_buffer() {
        if (pinned) ndzb_host_free(p);
        else free(p);
    }
    host_buffer(const host_buffer &) = delete;
    host_buffer &operator=(const host_buffer &) = delete;
};

// Input: either the whole file mapped (chunks are windows into the mapping) or read(2) into a staging buffer.
class reader {
  public:
    reader(const std::string &name, size_t chunk_bytes, bool allow_mmap) : _chunk_bytes(chunk_bytes) {
        if (!name.empty() && name != "-") {
            _fd = open(name.c_str(), O_RDONLY);
            if (_fd < 0) throw io_error("open: " + name + ": " + strerror(errno));
            _owns_fd = true;
        }
        struct stat st {};
        if (allow_mmap && fstat(_fd, &st) == 0 && S_ISREG(st.st_mode) && st.st_size > 0) {
            void *m = mmap(nullptr, static_cast<size_t>(st.st_size), PROT_READ, MAP_PRIVATE | MAP_POPULATE, _fd, 0);
            if (m != MAP_FAILED) {
                _map = static_cast<const unsigned char *>(m);
                _map_bytes = static_cast<size_t>(st.st_size);
            }
        }
    }
    ~reader() {
        if (_map) munmap(const_cast<unsigned char *>(_map), _map_bytes);
        if (_owns_fd) close(_fd);
    }
    reader(const reader &) = delete;
    reader &operator=(const reader &) = delete;

    // Up to chunk_bytes bytes into `dst`, of which the first `keep` are the unconsumed tail of the previous chunk
    // (reference io.cc:43-53, read_some). Returns the number of valid bytes in dst.
    size_t fill(unsigned char *dst, size_t keep, size_t prev_valid) {
        if (keep) memmove(dst, dst + (prev_valid - keep), keep);
        size_t have = keep;
        while (have < _chunk_bytes) {
            size_t got;
            if (_map) {
                got = _map_bytes - _map_pos < _chunk_bytes - have ? _map_bytes - _map_pos : _chunk_bytes - have;
                memcpy(dst + have, _map + _map_pos, got);
                _map_pos += got;
            } else {
                const ssize_t r = read(_fd, dst + have, _chunk_bytes - have);
                if (r < 0) {
                    if (errno == EINTR) continue;
                    throw io_error(std::string("read: ") + strerror(errno));
                }
                got = static_cast<size_t>(r);
            }
            if (got == 0) break;
            have += got;
        }
        return have;
    }

  private:
    int _fd = STDIN_FILENO;
    bool _owns_fd = false;
    size_t _chunk_bytes;
    const unsigned char *_map = nullptr;
    size_t _map_bytes = 0, _map_pos = 0;
};

class writer {
  public:
    explicit writer(const std::string &name) {
        if (!name.empty() && name != "-") {
            _fd = open(name.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0644);
            if (_fd < 0) throw io_error("open: " + name + ": " + strerror(errno));
            _owns_fd = true;
        }
    }
    ~writer() {
        if (_owns_fd) close(_fd);
    }
    writer(const writer &) = delete;
    writer &operator=(const writer &) = delete;
    void put(const void *p, size_t bytes) {
        const unsigned char *c = static_cast<const unsigned char *>(p);
        while (bytes) {
            const ssize_t w = write(_fd, c, bytes);
            if (w < 0) {
                if (errno == EINTR) continue;
                throw io_error(std::string("write: ") + strerror(errno));
            }
            c += w;
            bytes -= static_cast<size_t>(w);
        }
    }

  private:
    int _fd = STDOUT_FILENO;
    bool _owns_fd = false;
};

int run(const options &o) {
    const int dims = static_cast<int>(o.size.size());
    const uint32_t *size = o.size.data();
    const size_t word = o.dtype == NDZB_F32 ? 4 : 8;
    uint64_t elements = 1;
    for (uint32_t n : o.size) elements *= n;
    if (elements >= (1ull << 32)) throw std::runtime_error("array-size: more than 2^32-1 elements (the stream format's index type is 32 bits)");
    const size_t array_bytes = static_cast<size_t>(elements) * word;
    const uint64_t bound_words = ndzb_compressed_length_bound(o.dtype, dims, size);
    if (bound_words >= (1ull << 32)) throw std::runtime_error("array-size: compressed length bound exceeds the 32-bit index type");
    const size_t bound_bytes = static_cast<size_t>(bound_words) * word;

    ndzb_ctx *ctx = nullptr;
    check(ndzb_ctx_create(&ctx, o.dtype, dims, ndzb_num_hypercubes(dims, size), nullptr), "ndzb_ctx_create");
    struct ctx_guard {
        ndzb_ctx *c;
        ~ctx_guard() { ndzb_ctx_destroy(c); }
    } guard{ctx};

    reader in(o.input, o.decompress ? bound_bytes : array_bytes, !o.no_mmap);
    writer out(o.output);
    host_buffer raw(array_bytes), packed(bound_bytes);
    uint64_t kernel_ns_total = 0;

    if (!o.decompress) {
        size_t chunks = 0, compressed_bytes = 0;
        for (;;) {
            const size_t got = in.fill(static_cast<unsigned char *>(raw.p), 0, 0);
            if (got == 0) break;
            if (got != array_bytes) throw io_error("Input file size is not a multiple of the chunk size");  // reference io.cc:62
            uint32_t length = 0;
            uint64_t ns = 0;
            check(ndzb_offload_compress(ctx, raw.p, dims, size, packed.p, &length, &ns), "compress");
            out.put(packed.p, static_cast<size_t>(length) * word);
            compressed_bytes += static_cast<size_t>(length) * word;
            kernel_ns_total += ns;
            ++chunks;
        }
        const double raw_bytes = static_cast<double>(chunks) * array_bytes;
        fprintf(stderr, "raw = %.0f bytes", raw_bytes);
        if (chunks > 1) fprintf(stderr, " (%zu chunks of %zu bytes)", chunks, array_bytes);
        fprintf(stderr, ", compressed = %zu bytes, ratio = %.4f, time = %.3fs\n", compressed_bytes,
                raw_bytes > 0 ? compressed_bytes / raw_bytes : 0.0, kernel_ns_total * 1e-9);
    } else {
        // a compressed file has no chunk table: read up to one bound's worth, decode one array, keep what the
        // decoder did not consume for the next round (reference compress.cc:73-88)
        size_t keep = 0, valid = 0;
        for (;;) {
            valid = in.fill(static_cast<unsigned char *>(packed.p), keep, valid);
            if (valid == 0) break;
            uint32_t consumed = 0;
            uint64_t ns = 0;
            const int status = ndzb_offload_decompress(ctx, packed.p, static_cast<uint32_t>(valid / word), raw.p, dims, size, &consumed, &ns);
            if (status == NDZB_ERR_CORRUPT_STREAM) throw io_error("Compressed input ends inside a stream");
            check(status, "decompress");
            const size_t consumed_bytes = static_cast<size_t>(consumed) * word;
            if (consumed_bytes > valid) throw io_error("Compressed input ends inside a stream");
            out.put(raw.p, array_bytes);
            keep = valid - consumed_bytes;
            kernel_ns_total += ns;
        }
    }
    return EXIT_SUCCESS;
}

}  // namespace

int main(int argc, char **argv) {
    const std::string usage = std::string("Usage: ") + argv[0] + " [options]\n\n";
    options o;
    try {
        o = parse(argc, argv);
    } catch (const usage_error &e) {
        fprintf(stderr, "%s\n\n%s%s", e.what(), usage.c_str(), kOptions);
        return EXIT_FAILURE;
    }
    if (o.help) {
        printf("Compress or decompress binary float dump\n\n%s%s\n", usage.c_str(), kOptions);
        return EXIT_SUCCESS;
    }
    try {
        return run(o);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return EXIT_FAILURE;
    }
}
