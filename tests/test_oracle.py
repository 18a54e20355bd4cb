"""Pins the CPU oracle (oracle/ndzip_oracle.c) — CPU only, no GPU.

1. against the committed golden fixtures (generated from the unmodified reference CPU codec by
   tests/golden/make_golden.py; the first eight rows are the table of SURVEY.md §8c),
2. against the reference's own known-answer test for border slices
   (reference src/test/codec_generic_test.cc:102-111),
3. against the compiled reference itself (oracle/_ref) on primitives and whole streams, mirroring
   the bold parity tests of SURVEY.md §4 (reference src/test/codec_profile_test.inl).
"""
import json
import os
import zlib

import numpy as np
import pytest

from ndzip_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIDE = {1: 4096, 2: 64, 3: 16}
PROFILES = [(dt, d) for dt in ("float32", "float64") for d in (1, 2, 3)]


def _rows():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["rows"]


def _row_id(r):
    return f"{r['generator']}-{r['dtype']}-{'x'.join(map(str, r['shape']))}"


# SURVEY.md §8(c) golden table, verbatim: (dtype, shape, bound, words, crc32, first words)
SURVEY_TABLE = [
    ("float32", (8195,), 8453, 2782, "b5554034", "56a ad9 7dfe0000 f0000"),
    ("float32", (131, 131), 17677, 13051, "4764a137", "bfc 17db 2404 2fee"),
    ("float32", (35, 35, 35), 43907, 38563, "6d4456b8", "e07 1bd5 29a5 379f"),
    ("float64", (12291,), 12485, 5464, "07ef610c", "e3100000717 1553 7fbfe00000000000 1f00000000000"),
    ("float64", (195, 195), 38606, 34803, "9c4bf089", "1d3200000e6d 3a6600002bbb 578b000048f2 74bb0000661c"),
    ("float64", (51, 51, 51), 134393, 124489, "ab8477d8", "1dcb00000ed5 3b5a00002c8a 591500004a14 768b000067d6"),
    ("float32", (100,), 100, 100, "30f5f92f", "0 3ee00000 3f600000 3fa80000"),
    ("float64", (15, 15, 15), 3375, 3375, "d589e5e9", "0 3fdc000000000000 3fec000000000000 3ff5000000000000"),
]


@pytest.mark.parametrize("dtype,shape,bound,words,crc,first", SURVEY_TABLE)
def test_oracle_reproduces_survey_table(oracle, dtype, shape, bound, words, crc, first):
    data = synth.ramp(shape, dtype)
    assert oracle.compressed_length_bound(dtype, shape) == bound
    stream = oracle.compress(data)
    assert stream.size == words
    assert "%08x" % zlib.crc32(stream.tobytes()) == crc
    assert " ".join("%x" % int(w) for w in stream[:4]) == first
    back, consumed = oracle.decompress(stream, dtype, shape)
    assert consumed == words
    assert back.tobytes() == data.tobytes()


@pytest.mark.parametrize("row", _rows(), ids=_row_id)
def test_oracle_matches_golden_fixture(oracle, row):
    data = synth.make(row["generator"], tuple(row["shape"]), row["dtype"], **row["kwargs"])
    assert "%08x" % zlib.crc32(data.tobytes()) == row["input_crc32"], "generator drifted"
    assert oracle.compressed_length_bound(row["dtype"], row["shape"]) == row["bound"]
    stream = oracle.compress(data)
    assert stream.size == row["stream_words"]
    assert "%08x" % zlib.crc32(stream.tobytes()) == row["stream_crc32"]
    assert ["%x" % int(w) for w in stream[:4]] == row["first_words"]
    back, consumed = oracle.decompress(stream, row["dtype"], tuple(row["shape"]))
    assert consumed == stream.size
    assert back.tobytes() == data.tobytes()


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_oracle_matches_golden_cube_stream(oracle, dtype, dims):
    cubes = np.load(os.path.join(ROOT, "tests", "golden", "cubes.npz"))
    data = synth.hashed((SIDE[dims],) * dims, dtype, seed=11)
    data.reshape(-1)[: (32 if dtype == "float32" else 64)] = 0
    expect = cubes[f"{dtype}_{dims}d_stream"]
    got = oracle.compress(data)
    assert np.array_equal(got, expect)


def test_border_slices_known_answers(oracle):
    # reference src/test/codec_generic_test.cc:102-111, verbatim expectations
    assert oracle.border_slices((4, 4), 4) == []
    assert oracle.border_slices((4, 6), 2) == []
    assert oracle.border_slices((5, 4), 4) == [(16, 4)]
    assert oracle.border_slices((4, 5), 4) == [(4, 1), (9, 1), (14, 1), (19, 1)]
    assert oracle.border_slices((4, 5), 2) == [(4, 1), (9, 1), (14, 1), (19, 1)]
    assert oracle.border_slices((4, 6), 4) == [(4, 2), (10, 2), (16, 2), (22, 2)]
    assert oracle.border_slices((4, 6), 5) == [(0, 24)]
    assert oracle.border_slices((6, 4), 5) == [(0, 24)]


@pytest.mark.parametrize("bits", [np.uint32, np.uint64])
def test_transpose_is_involution(oracle, bits):
    # reference src/test/codec_generic_test.cc:65-81
    B = np.dtype(bits).itemsize * 8
    w = synth.engineered_cube(bits)[:B]
    t = oracle.transpose_bits(w)
    assert np.array_equal(oracle.transpose_bits(t), w)
    # definition check on one entry: out[i] bit (B-1-j) == in[j] bit (B-1-i)
    for i, j in [(0, 0), (3, 17), (B - 1, 5), (7, B - 1)]:
        assert (int(t[i]) >> (B - 1 - j)) & 1 == (int(w[j]) >> (B - 1 - i)) & 1


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_block_transform_reversible(oracle, dtype, dims):
    # reference src/test/codec_profile_test.inl:23-34
    bits = np.uint32 if dtype == "float32" else np.uint64
    cube = synth.raw_bits((4096,), dtype, seed=21).view(bits)
    fwd = oracle.block_transform(cube, dims)
    assert not np.array_equal(fwd, cube)
    assert np.array_equal(oracle.block_transform(fwd, dims, inverse=True), cube)


@pytest.mark.parametrize("bits", [np.uint32, np.uint64])
def test_zero_bit_encode_reversible(oracle, bits):
    # reference src/test/codec_generic_test.cc:38-62 + codec_profile_test.inl:552-567 (engineered zeros)
    cube = synth.engineered_cube(bits)
    enc = oracle.zero_bit_encode(cube)
    B = np.dtype(bits).itemsize * 8
    assert 4096 // B <= enc.size < 4096 + 4096 // B
    dec, consumed = oracle.zero_bit_decode(enc)
    assert consumed == enc.size
    assert np.array_equal(dec, cube)


def test_zero_hypercube_extents(oracle):
    # reference src/test/codec_profile_test.inl:1045-1082
    for dtype in ("float32", "float64"):
        for dims in (1, 2, 3):
            for n in (0, 1):
                shape = (n,) * dims
                data = synth.ramp(shape, dtype)
                stream = oracle.compress(data)
                assert stream.size == data.size == oracle.compressed_length_bound(dtype, shape)
                back, consumed = oracle.decompress(stream, dtype, shape)
                assert consumed == stream.size and back.tobytes() == data.tobytes()


# ---------------------------------------------------------------- against the compiled reference

@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_primitives_match_reference(oracle, reference, dtype, dims):
    bits = np.uint32 if dtype == "float32" else np.uint64
    B = np.dtype(bits).itemsize * 8
    cube = synth.raw_bits((4096,), dtype, seed=33).view(bits)
    # scalar normative transform, and the AVX2 transform the CPU encoder really runs
    # (reference src/test/codec_profile_test.inl:889-947 pins these equal)
    fwd_ref = reference.block_transform(cube, dims)
    assert np.array_equal(reference.block_transform(cube, dims, simd=True), fwd_ref)
    assert np.array_equal(oracle.block_transform(cube, dims), fwd_ref)
    assert np.array_equal(oracle.block_transform(fwd_ref, dims, inverse=True),
                          reference.block_transform(fwd_ref, dims, inverse=True))
    # transpose
    assert np.array_equal(oracle.transpose_bits(cube[:B]), reference.transpose_bits(cube[:B]))
    # residual encoding of the engineered cube (codec_profile_test.inl:552-729)
    eng = synth.engineered_cube(bits)
    assert np.array_equal(oracle.zero_bit_encode(eng), reference.zero_bit_encode(eng))
    dec_o, n_o = oracle.zero_bit_decode(reference.zero_bit_encode(eng))
    assert np.array_equal(dec_o, eng)


@pytest.mark.parametrize("shape,side", [((4, 5), 4), ((255, 255), 64), ((63, 63, 63), 16), ((33, 16, 48), 16),
                                        ((16, 35, 32), 16), ((32, 16, 21), 16), ((9, 40, 40), 16), ((16383,), 4096)])
def test_border_slices_match_reference(oracle, reference, shape, side):
    assert oracle.border_slices(shape, side) == reference.border_slices(shape, side)


@pytest.mark.parametrize("dtype,dims", PROFILES)
@pytest.mark.parametrize("gen", ["hashed", "raw_bits", "quantised", "poly"])
def test_streams_match_reference(oracle, reference, dtype, dims, gen):
    # shapes of reference src/test/codec_profile_test.inl:37-140, 952-995: 4*side-1 has a border
    n = SIDE[dims] * 4 - 1 if dims > 1 else SIDE[dims] * 2 + 77
    shape = (n,) * dims
    kw = {} if gen == "poly" else {"seed": 101}
    data = synth.make(gen, shape, dtype, **kw)
    data.reshape(-1)[: (32 if dtype == "float32" else 64)] = 0
    expect = reference.compress(data, threads=1)
    got = oracle.compress(data)
    assert got.size == expect.size
    assert np.array_equal(got, expect)
    back, consumed = oracle.decompress(expect, dtype, shape)
    assert consumed == expect.size and back.tobytes() == data.tobytes()
    # cross pairing: reference decodes the oracle's stream
    back_ref, consumed_ref = reference.decompress(got, dtype, shape)
    assert consumed_ref == got.size and back_ref.tobytes() == data.tobytes()


def test_reference_openmp_equals_serial(reference):
    data = synth.hashed((48, 48, 48), "float32", seed=8)
    assert np.array_equal(reference.compress(data, threads=1), reference.compress(data, threads=4))
