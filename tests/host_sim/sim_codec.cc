// CPU simulation of the CUDA cube codec: drives the __host__ __device__ building blocks of
// ndzip_b200/csrc/ndzb_cube.cuh "thread by thread" and "lane by lane" in the same phase order as
// the kernels (phases separated by __syncthreads there are separate loops here). Used by
// tests/test_host_sim.py to check the device logic bit for bit against the oracle without a GPU.
// This is a TEST harness; the product has no CPU path.
#include "../../ndzip_b200/csrc/ndzb_cube.cuh"

#include <cstring>
#include <vector>

using namespace ndzb;

namespace {

template<typename Bits>
uint32_t popc_bits(Bits v) {
    if constexpr (sizeof(Bits) == 4) return popc32(v);
    else return popc32(static_cast<uint32_t>(v >> 32)) + popc32(static_cast<uint32_t>(v));
}

template<typename Bits, int Dims>
uint32_t encode_cube(const Bits *cube, Bits *out) {
    using tr = codec_traits<Bits>;
    constexpr int tile_words = tr::cube_words32 > tr::image_words32 ? tr::cube_words32 : tr::image_words32;
    std::vector<uint32_t> tile(tile_words, 0xdeadbeefu);
    for (int e = 0; e < kCubeElems; ++e) {
        const int w = input_layout<Bits, Dims>::elem(e);
        tile[w] = static_cast<uint32_t>(cube[e]);
        if constexpr (sizeof(Bits) == 8) tile[w + 1] = static_cast<uint32_t>(cube[e] >> 32);
    }

    // phase 1: residuals (reads the input tile only), heads, plane counts
    std::vector<Bits> res(kCubeElems);
    for (int u = 0; u < kCubeThreads; ++u) residual_run<Bits, Dims>(tile.data(), u, &res[32 * u]);
    if constexpr (Dims == 3) {
        // compress_ws_kernel's 3-D path (residual_run_3d_warp): z difference, the second half of run u-1 by
        // __shfl_up_sync(.., 1) within the warp (lane 0 keeps its own value), then the y / x differences. Must give
        // exactly what the five-neighbour stencil above gives.
        std::vector<Bits> warp(kCubeElems), above(16 * kCubeThreads);
        for (int u = 0; u < kCubeThreads; ++u) residual3_zdiff<Bits>(tile.data(), u, &warp[32 * u]);
        for (int u = 0; u < kCubeThreads; ++u) {
            const int src = (u & 31) == 0 ? u : u - 1;
            memcpy(&above[16 * u], &warp[32 * src + 16], 16 * sizeof(Bits));
        }
        for (int u = 0; u < kCubeThreads; ++u) residual3_finish<Bits>(u, &above[16 * u], &warp[32 * u]);
        if (memcmp(warp.data(), res.data(), kCubeElems * sizeof(Bits)) != 0) return 0;  // no cube compresses to 0 words: the test fails
        res = warp;
    }
    Bits heads[tr::chunks];
    for (int c = 0; c < tr::chunks; ++c) heads[c] = 0;
    for (int e = 0; e < kCubeElems; ++e) heads[e / tr::bits] |= res[e];
    uint32_t body[tr::chunks];
    uint32_t total = tr::chunks;
    for (int c = 0; c < tr::chunks; ++c) {
        body[c] = total;
        total += popc_bits(heads[c]);
    }

    // phase 2: planes, compacted into the cube image (aliases the input tile, as in the kernel)
    if constexpr (sizeof(Bits) == 4) {
        for (int u = 0; u < kCubeThreads; ++u) {
            uint32_t planes[32];
            planes_of_run(&res[32 * u], planes);
            compact_planes(tile.data(), u, heads[u], body[u], planes);
        }
    } else {
        for (int u = 0; u < kCubeThreads; ++u) {
            uint32_t ph[32], pl[32];
            planes_of_run(&res[32 * u], ph, pl);
            compact_planes(tile.data(), u >> 1, (u & 1) == 0, heads[u >> 1], body[u >> 1], ph, pl);
        }
    }

    // phase 3: linear copy-out
    memcpy(out, tile.data(), static_cast<size_t>(total) * sizeof(Bits));
    return total;
}

template<typename Bits, int Dims>
uint32_t decode_cube(const Bits *in, Bits *cube) {
    using tr = codec_traits<Bits>;
    constexpr int tile_words = tr::cube_words32 > tr::image_words32 ? tr::cube_words32 : tr::image_words32;
    std::vector<uint32_t> stage(tile_words, 0xdeadbeefu);
    uint32_t body[tr::chunks];
    uint32_t total = tr::chunks;
    for (int c = 0; c < tr::chunks; ++c) {
        body[c] = total;
        total += popc_bits(in[c]);
    }
    memcpy(stage.data(), in, static_cast<size_t>(total) * sizeof(Bits));  // coalesced copy-in

    // per-thread: image -> residual run, x-direction prefix inside the run
    std::vector<Bits> res(kCubeElems);
    for (int u = 0; u < kCubeThreads; ++u) {
        Bits *r = &res[32 * u];
        if constexpr (sizeof(Bits) == 4) {
            run_of_image(stage.data(), in[u], body[u], r);
        } else {
            run_of_image(stage.data(), (u & 1) == 0, in[u >> 1], body[u >> 1], r);
        }
        if constexpr (Dims == 3) {
            for (int i = 1; i < 16; ++i) { r[i] += r[i - 1]; r[16 + i] += r[16 + i - 1]; }
        } else {
            for (int j = 1; j < 32; ++j) r[j] += r[j - 1];
        }
    }
    // cross-thread part of the x prefix (device: block scan / shuffle)
    if constexpr (Dims == 1) {
        Bits carry = 0;
        for (int u = 0; u < kCubeThreads; ++u) {
            const Bits total = res[32 * u + 31];
            for (int j = 0; j < 32; ++j) res[32 * u + j] += carry;
            carry += total;
        }
    } else if constexpr (Dims == 2) {
        for (int u = 1; u < kCubeThreads; u += 2) {
            const Bits left = res[32 * (u - 1) + 31];
            for (int j = 0; j < 32; ++j) res[32 * u + j] += left;
        }
    }
    // the value tile aliases the compressed image (barrier in the kernel: all reads before any write)
    for (int u = 0; u < kCubeThreads; ++u) store_run(stage.data(), u, &res[32 * u]);

    // remaining axes: column passes in shared memory
    auto column_pass = [&](int first, int stride, int n) {
        Bits acc = tile_load<Bits>(stage.data(), first);
        for (int k = 1; k < n; ++k) {
            acc += tile_load<Bits>(stage.data(), first + k * stride);
            tile_store<Bits>(stage.data(), first + k * stride, acc);
        }
    };
    if constexpr (Dims == 2) {
        for (int x = 0; x < 64; ++x) column_pass(x, 64, 64);
    } else if constexpr (Dims == 3) {
        for (int z = 0; z < 16; ++z)
            for (int x = 0; x < 16; ++x) column_pass(z * 256 + x, 16, 16);
        for (int yx = 0; yx < 256; ++yx) column_pass(yx, 256, 16);
    }
    for (int e = 0; e < kCubeElems; ++e) cube[e] = rotr1(tile_load<Bits>(stage.data(), e));
    return total;
}

template<typename Bits>
uint32_t encode_dispatch(int dims, const void *cube, void *out) {
    switch (dims) {
        case 1: return encode_cube<Bits, 1>(static_cast<const Bits *>(cube), static_cast<Bits *>(out));
        case 2: return encode_cube<Bits, 2>(static_cast<const Bits *>(cube), static_cast<Bits *>(out));
        default: return encode_cube<Bits, 3>(static_cast<const Bits *>(cube), static_cast<Bits *>(out));
    }
}
template<typename Bits>
uint32_t decode_dispatch(int dims, const void *in, void *cube) {
    switch (dims) {
        case 1: return decode_cube<Bits, 1>(static_cast<const Bits *>(in), static_cast<Bits *>(cube));
        case 2: return decode_cube<Bits, 2>(static_cast<const Bits *>(in), static_cast<Bits *>(cube));
        default: return decode_cube<Bits, 3>(static_cast<const Bits *>(in), static_cast<Bits *>(cube));
    }
}

grid_geom make_geom(int dims, const uint32_t *size) {
    grid_geom g{};
    const uint32_t side = dims == 1 ? 4096 : dims == 2 ? 64 : 16;
    for (int d = 0; d < 3; ++d) { g.n[d] = 1; g.cubes[d] = 1; }
    g.num_cubes = 1;
    for (int d = 0; d < dims; ++d) {
        g.n[3 - dims + d] = size[d];
        g.cubes[3 - dims + d] = size[d] / side;
        g.num_cubes *= size[d] / side;
    }
    g.div_x = make_fastdiv(g.cubes[2]);
    g.div_y = make_fastdiv(g.cubes[1]);
    return g;
}

}  // namespace

template<typename Bits>
int strip_address_mismatches() {
    int bad = 0;
    for (int o = 0; o < 16; ++o) {
        for (int xq = 0; xq < 8; ++xq) {
            const strip_addr_y3<Bits> ay(o, xq);
            const strip_addr_z3<Bits> az(o, xq);
            for (int k = 0; k < 16; ++k) {
                bad += ay.at(k) != tile3_elem<Bits>(o * 256 + k * 16 + xq * 2);
                bad += az.at(k) != tile3_elem<Bits>(k * 256 + o * 16 + xq * 2);
            }
        }
    }
    for (int seg = 0; seg < 4; ++seg) {
        for (int xq = 0; xq < 32; ++xq) {
            const strip_addr_y2<Bits> a(seg, xq);
            for (int k = 0; k < 16; ++k) bad += a.at(k) != tile_elem<Bits>((seg * 16 + k) * 64 + xq * 2);
        }
    }
    return bad;
}

extern "C" {

// number of strip addresses (decoder column passes) that differ from tile_elem(); must be 0
int sim_strip_address_mismatches() { return strip_address_mismatches<uint32_t>() + strip_address_mismatches<uint64_t>(); }

// cube: 4096 raw words in cube-local order; out: compressed cube. Returns words written.
uint32_t sim_encode_cube(int dtype, int dims, const void *cube, void *out) {
    return dtype == 0 ? encode_dispatch<uint32_t>(dims, cube, out) : encode_dispatch<uint64_t>(dims, cube, out);
}

uint32_t sim_decode_cube(int dtype, int dims, const void *in, void *cube) {
    return dtype == 0 ? decode_dispatch<uint32_t>(dims, in, cube) : decode_dispatch<uint64_t>(dims, in, cube);
}

void sim_transpose32(uint32_t *a) { transpose32(a); }

// linear element index of cube-local element e of hypercube hc
uint64_t sim_cube_element_index(int dims, const uint32_t *size, uint32_t hc, int e) {
    const grid_geom g = make_geom(dims, size);
    switch (dims) {
        case 1: return cube_origin<1>(g, hc) + cube_local_offset<1>(g, e);
        case 2: return cube_origin<2>(g, hc) + cube_local_offset<2>(g, e);
        default: return cube_origin<3>(g, hc) + cube_local_offset<3>(g, e);
    }
}

}  // extern "C"
