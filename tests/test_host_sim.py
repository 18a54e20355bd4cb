"""CPU simulation of the CUDA cube codec (tests/host_sim/sim_codec.cc drives the __host__ __device__
functions of ndzip_b200/csrc/ndzb_cube.cuh thread by thread) checked bit for bit against the oracle.
Catches layout / stencil / transpose / compaction bugs before any GPU time is spent."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from ndzip_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIM_DIR = os.path.join(ROOT, "tests", "host_sim")
SIDE = {1: 4096, 2: 64, 3: 16}
PROFILES = [(dt, d) for dt in ("float32", "float64") for d in (1, 2, 3)]


@pytest.fixture(scope="module")
def sim():
    so = os.path.join(SIM_DIR, "libsim_codec.so")
    src = os.path.join(SIM_DIR, "sim_codec.cc")
    hdr = os.path.join(ROOT, "ndzip_b200", "csrc", "ndzb_cube.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", so, src], check=True)
    L = ctypes.CDLL(so)
    L.sim_encode_cube.restype = ctypes.c_uint32
    L.sim_encode_cube.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.sim_decode_cube.restype = ctypes.c_uint32
    L.sim_decode_cube.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    L.sim_transpose32.argtypes = [ctypes.c_void_p]
    L.sim_cube_element_index.restype = ctypes.c_uint64
    L.sim_cube_element_index.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.c_int]
    return L


def _inputs(dtype):
    bits = np.uint32 if dtype == "float32" else np.uint64
    yield "raw", synth.raw_bits((4096,), dtype, seed=5).view(bits)
    yield "hashed", synth.hashed((4096,), dtype, seed=6).view(bits)
    yield "quantised", synth.quantised((4096,), dtype, seed=7).view(bits)
    yield "ramp", synth.ramp((4096,), dtype).view(bits)
    yield "zeros", np.zeros(4096, dtype=bits)
    z = synth.hashed((4096,), dtype, seed=8).view(bits).copy()
    z[:128] = 0
    yield "zero_head", z


def test_transpose32_is_lsb_transpose(sim):
    a = synth.raw_bits((32,), "float32", seed=3).view(np.uint32).copy()
    t = a.copy()
    sim.sim_transpose32(t.ctypes.data)
    for k in range(32):
        for b in range(32):
            assert (int(t[k]) >> b) & 1 == (int(a[b]) >> k) & 1


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_simulated_cube_encode_matches_oracle(sim, oracle, dtype, dims):
    code = 0 if dtype == "float32" else 1
    bits = np.uint32 if code == 0 else np.uint64
    for name, cube in _inputs(dtype):
        expect = oracle.zero_bit_encode(oracle.block_transform(cube, dims))
        out = np.zeros(4096 + 128, dtype=bits)
        n = sim.sim_encode_cube(code, dims, cube.ctypes.data, out.ctypes.data)
        assert n == expect.size, name
        assert np.array_equal(out[:n], expect), name


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_simulated_cube_decode_matches_oracle(sim, oracle, dtype, dims):
    code = 0 if dtype == "float32" else 1
    bits = np.uint32 if code == 0 else np.uint64
    for name, cube in _inputs(dtype):
        enc = oracle.zero_bit_encode(oracle.block_transform(cube, dims))
        enc_padded = np.concatenate([enc, np.zeros(8, dtype=bits)])
        back = np.zeros(4096, dtype=bits)
        n = sim.sim_decode_cube(code, dims, enc_padded.ctypes.data, back.ctypes.data)
        assert n == enc.size, name
        assert np.array_equal(back, cube), name


@pytest.mark.parametrize("shape", [(3 * 4096 + 5,), (130, 200), (33, 50, 70)])
def test_cube_addressing_matches_oracle(sim, oracle, shape):
    # reference src/test/codec_profile_test.inl:514-549 "Flattening of hypercubes identical"
    dims = len(shape)
    data = np.arange(int(np.prod(shape)), dtype=np.uint32).view(np.float32).reshape(shape)
    size = (ctypes.c_uint32 * 3)(*(list(shape) + [0] * (3 - dims)))
    H = oracle.num_hypercubes(shape)
    assert H >= 2
    for hc in {0, 1, H - 1}:
        cube = oracle.load_cube(data, hc)
        for e in (0, 1, 15, 16, 63, 64, 255, 256, 1000, 4095):
            assert sim.sim_cube_element_index(dims, size, hc, e) == int(cube[e])


def test_decoder_strip_addresses_match_the_tile_layout(sim):
    # strength-reduced addresses of the y / z column passes (strip_addr_*) == tile_elem() for every strip
    sim.sim_strip_address_mismatches.restype = ctypes.c_int
    assert sim.sim_strip_address_mismatches() == 0
