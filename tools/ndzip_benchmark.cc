// ndzip-benchmark — the `ndzip-cuda` column of the reference's benchmark driver (reference
// src/benchmark/benchmark.cc:300-349, 1395-1500) on top of the C ABI: same dataset description file
// (docs/benchmarking.md: `file;float|double;n0 [n1 [n2]]`, files relative to the CSV), same repetition rules
// (-t / -r / -R / --no-warmup), same output rows, so the result can be fed to the reference's plot_benchmark.py
// next to rows produced by the reference binary for the other algorithms.
//
//   dataset;data type;dimensions;algorithm;tunable;number of threads;compression times (microseconds);
//   decompression times (microseconds);uncompressed bytes;compressed bytes
//
// Times are the cudaEvent interval around the kernels of one offloader call (the reference's kernel_duration,
// cuda_codec.inl:687-704), host<->device copies excluded, exactly what the reference records for GPU algorithms
// (benchmark.cc:329-342). Third-party compressors and the CPU / SYCL ndzip targets are out of scope.
#include "../include/ndzip_b200.h"

#include <cerrno>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace {

struct dataset {
    std::string path, name;
    int dtype = NDZB_F32;
    std::vector<uint32_t> extent;
};

struct options {
    std::string csv;
    double time_each_ms = 1000;
    unsigned min_reps = 1, max_reps = 100;
    bool warmup = true, help = false;
};

const char *kUsage =
        "Usage: ndzip-benchmark [options] csv-file\n\n"
        "Options:\n"
        "  --help                   show this help\n"
        "  -a [ --algorithms ] A…   algorithms to evaluate: this build has ndzip-cuda only\n"
        "  -t [ --time-each ] MS    repeat each for at least t ms (default 1000)\n"
        "  -r [ --min-reps ] N      repeat each at least n times (default 1)\n"
        "  -R [ --max-reps ] N      repeat each at most n times (default 100)\n"
        "  --no-warmup              do not perform an additional warm-up step per benchmark\n"
        "  --no-mmap                accepted for compatibility (files are read into pinned memory)\n";

options parse(int argc, char **argv) {
    options o;
    auto number = [&](int &i, const char *name) -> double {
        if (i + 1 >= argc) throw std::invalid_argument(std::string("the required argument for option '") + name + "' is missing");
        char *end = nullptr;
        const double v = strtod(argv[++i], &end);
        if (*end || v < 0) throw std::invalid_argument(std::string("the argument for option '") + name + "' is invalid");
        return v;
    };
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        if (a == "--help") o.help = true;
        else if (a == "--no-warmup") o.warmup = false;
        else if (a == "--no-mmap") {}
        else if (a == "-t" || a == "--time-each") o.time_each_ms = number(i, "--time-each");
        else if (a == "-r" || a == "--min-reps") o.min_reps = static_cast<unsigned>(number(i, "--min-reps"));
        else if (a == "-R" || a == "--max-reps") o.max_reps = static_cast<unsigned>(number(i, "--max-reps"));
        else if (a == "-a" || a == "--algorithms") {
            bool ours = false;
            while (i + 1 < argc && argv[i + 1][0] != '-') {
                const std::string algo = argv[++i];
                if (algo == "ndzip-cuda" || algo == "ndzip-gpu") ours = true;
                else fprintf(stderr, "ndzip-benchmark: algorithm %s is not part of this build, skipped\n", algo.c_str());
            }
            if (!ours) throw std::invalid_argument("no algorithm of this build selected (available: ndzip-cuda)");
        } else if (!a.empty() && a[0] == '-') {
            throw std::invalid_argument("unrecognised option '" + a + "'");
        } else if (o.csv.empty()) {
            o.csv = a;
        } else {
            throw std::invalid_argument("too many positional options");
        }
    }
    if (!o.help && o.csv.empty()) throw std::invalid_argument("the option '--csv-file' is required but missing");
    if (o.max_reps < o.min_reps) o.max_reps = o.min_reps;
    return o;
}

// docs/benchmarking.md / benchmark.cc load_metadata_file: `name;float|double;n0 n1 n2`
std::vector<dataset> load_metadata(const std::string &csv) {
    std::ifstream in(csv);
    if (!in) throw std::runtime_error(csv + ": " + strerror(errno));
    const size_t slash = csv.find_last_of('/');
    const std::string dir = slash == std::string::npos ? std::string() : csv.substr(0, slash + 1);
    std::vector<dataset> sets;
    std::string line;
    for (size_t lineno = 1; std::getline(in, line); ++lineno) {
        if (line.empty()) continue;
        std::istringstream fields(line);
        std::string name, type, dims;
        if (!std::getline(fields, name, ';') || !std::getline(fields, type, ';') || !std::getline(fields, dims, ';')) {
            throw std::runtime_error(csv + ":" + std::to_string(lineno) + ": expected 3 parameters");
        }
        dataset d;
        d.name = name;
        d.path = dir + name;
        if (type == "float") d.dtype = NDZB_F32;
        else if (type == "double") d.dtype = NDZB_F64;
        else throw std::runtime_error(csv + ":" + std::to_string(lineno) + ": invalid data type " + type);
        std::istringstream ext(dims);
        unsigned long long n;
        while (ext >> n) d.extent.push_back(static_cast<uint32_t>(n));
        if (d.extent.empty() || d.extent.size() > 3) throw std::runtime_error(csv + ":" + std::to_string(lineno) + ": expected 1 to 3 dimensions");
        sets.push_back(std::move(d));
    }
    return sets;
}

void check(int status, const char *what) {
    if (status == NDZB_OK) return;
    std::string msg = std::string(what) + ": " + ndzb_strerror(status);
    if (status == NDZB_ERR_CUDA) msg += std::string(" (") + ndzb_last_cuda_error() + ")";
    throw std::runtime_error(msg);
}

struct pinned {
    void *p = nullptr;
    explicit pinned(size_t bytes) { check(ndzb_host_alloc(&p, bytes), "ndzb_host_alloc"); }
    ~pinned() { ndzb_host_free(p); }
    pinned(const pinned &) = delete;
    pinned &operator=(const pinned &) = delete;
};

std::string join_us(const std::vector<uint64_t> &ns) {
    std::string s;
    for (size_t i = 0; i < ns.size(); ++i) {
        if (i) s += ',';
        s += std::to_string(ns[i] / 1000);  // the reference prints whole microseconds
    }
    return s;
}

// benchmark.cc:190-230: (warm-up,) then repeat until min-reps AND the time budget are reached, at most max-reps
template<typename F>
std::vector<uint64_t> repeat(const options &o, F &&once) {
    if (o.warmup) once();
    std::vector<uint64_t> times;
    double total_ms = 0;
    while (times.size() < o.max_reps && (times.size() < o.min_reps || total_ms < o.time_each_ms)) {
        const uint64_t ns = once();
        times.push_back(ns);
        total_ms += ns * 1e-6;
    }
    return times;
}

void run_one(const options &o, const dataset &d) {
    const int dims = static_cast<int>(d.extent.size());
    const size_t word = d.dtype == NDZB_F32 ? 4 : 8;
    uint64_t elements = 1;
    for (uint32_t n : d.extent) elements *= n;
    const uint64_t bound = ndzb_compressed_length_bound(d.dtype, dims, d.extent.data());
    if (elements >= (1ull << 32) || bound >= (1ull << 32)) throw std::runtime_error(d.name + ": larger than the stream format's 32-bit index type");
    const size_t bytes = static_cast<size_t>(elements) * word;

    pinned input(bytes), stream(static_cast<size_t>(bound) * word), output(bytes);
    FILE *f = fopen(d.path.c_str(), "rb");
    if (!f) throw std::runtime_error(d.path + ": " + strerror(errno));
    const size_t got = fread(input.p, 1, bytes, f);
    fclose(f);
    if (got != bytes) throw std::runtime_error(d.path + ": file is shorter than " + std::to_string(bytes) + " bytes");

    ndzb_ctx *ctx = nullptr;
    check(ndzb_ctx_create(&ctx, d.dtype, dims, ndzb_num_hypercubes(dims, d.extent.data()), nullptr), "ndzb_ctx_create");
    struct guard {
        ndzb_ctx *c;
        ~guard() { ndzb_ctx_destroy(c); }
    } g{ctx};

    uint32_t length = 0;
    const auto ctimes = repeat(o, [&] {
        uint64_t ns = 0;
        check(ndzb_offload_compress(ctx, input.p, dims, d.extent.data(), stream.p, &length, &ns), "compress");
        return ns;
    });
    const auto dtimes = repeat(o, [&] {
        uint64_t ns = 0;
        uint32_t consumed = 0;
        check(ndzb_offload_decompress(ctx, stream.p, length, output.p, dims, d.extent.data(), &consumed, &ns), "decompress");
        return ns;
    });
    if (memcmp(input.p, output.p, bytes) != 0) {  // benchmark.cc:1321-1326
        throw std::logic_error("mismatch between input and decompressed buffer for " + d.name + " with ndzip-cuda (tunable=1)");
    }
    // benchmark.cc:1332-1337 prints metadata.extent.size() (the NUMBER of dimensions) in the third column
    printf("%s;%s;%zu;ndzip-cuda;1;1;%s;%s;%zu;%zu\n", d.name.c_str(), d.dtype == NDZB_F32 ? "float" : "double", d.extent.size(),
            join_us(ctimes).c_str(), join_us(dtimes).c_str(), bytes, static_cast<size_t>(length) * word);
    fflush(stdout);
}

}  // namespace

int main(int argc, char **argv) {
    options o;
    try {
        o = parse(argc, argv);
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n\n%s", e.what(), kUsage);
        return EXIT_FAILURE;
    }
    if (o.help) {
        printf("Benchmark the ndzip CUDA codec on float data sets\n\n%s\nAvailable algorithms: ndzip-cuda\n", kUsage);
        return EXIT_SUCCESS;
    }
    try {
        const auto sets = load_metadata(o.csv);
        printf("dataset;data type;dimensions;algorithm;tunable;number of threads;"
               "compression times (microseconds);decompression times (microseconds);"
               "uncompressed bytes;compressed bytes\n");
        for (const auto &d : sets) run_one(o, d);
        return EXIT_SUCCESS;
    } catch (const std::exception &e) {
        fprintf(stderr, "%s\n", e.what());
        return EXIT_FAILURE;
    }
}
