"""CPU-only checks of the drop-in boundary: the C-ABI library builds/loads, exports every symbol that
include/ndzip_b200.h declares, its host-side stream arithmetic matches the oracle, and the compute
entry points fail loudly (no CPU fallback) when no CUDA device is present."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from ndzip_b200 import build as nzbuild
    from ndzip_b200 import _lib
    nzbuild.build()  # no-op when up to date
    return _lib.load()


def test_header_symbols_are_exported(lib):
    header = open(os.path.join(ROOT, "include", "ndzip_b200.h")).read()
    declared = set(re.findall(r"\b(ndzb_[a-z0-9_]+)\s*\(", header))
    declared -= {"ndzb_status"}
    assert len(declared) >= 18
    raw = ctypes.CDLL(os.path.join(ROOT, "ndzip_b200", "libndzip_b200.so"))
    for name in sorted(declared):
        assert hasattr(raw, name), f"{name} declared in include/ndzip_b200.h but not exported"
    from ndzip_b200 import _lib
    assert declared == {s[0] for s in _lib.SYMBOLS}


def test_no_product_dependency_on_oracle():
    # the product package must never import or link the checkers
    pkg = os.path.join(ROOT, "ndzip_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "from oracle" not in text, f
                assert "ndzip_oracle" not in text and "libndzip_ref" not in text, f


SHAPES = [(0,), (1,), (4095,), (4096,), (3 * 4096 + 5,), (63, 64), (64, 64), (255, 255), (130, 64), (70, 10),
          (16, 16, 16), (63, 63, 63), (33, 16, 48), (9, 40, 40), (512, 512, 512), (1 << 24,), (8192, 8192)]


@pytest.mark.parametrize("shape", SHAPES)
def test_host_arithmetic_matches_oracle(lib, oracle, shape):
    import ndzip_b200 as nz
    from ndzip_b200 import _lib
    dims, sz = _lib.size3(shape)
    assert nz.num_hypercubes(shape) == oracle.num_hypercubes(shape)
    assert lib.ndzb_border_element_count(dims, sz) == oracle.border_element_count(shape)
    for dtype in ("float32", "float64"):
        assert nz.compressed_length_bound(dtype, shape) == oracle.compressed_length_bound(dtype, shape)
    H = nz.num_hypercubes(shape)
    assert lib.ndzb_header_words(0, H) == H
    assert lib.ndzb_header_words(1, H) == (H + 1) // 2
    assert lib.ndzb_compressed_cube_bound(0) == 4224 and lib.ndzb_compressed_cube_bound(1) == 4160


def test_bound_matches_reference(lib, reference):
    import ndzip_b200 as nz
    for shape in SHAPES:
        if int(np.prod(shape)) == 0:
            continue
        for dtype in ("float32", "float64"):
            assert nz.compressed_length_bound(dtype, shape) == reference.compressed_length_bound(dtype, shape)


def test_compute_fails_loudly_without_a_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import ndzip_b200 as nz
    with pytest.raises(nz.NdzipB200Error, match="CUDA"):
        nz.make_cuda_compressor("float32", (64, 64))
    with pytest.raises(nz.NdzipB200Error, match="CUDA"):
        nz.make_cuda_offloader("float64", 3)


def test_argument_errors_mirror_the_reference(lib):
    import ndzip_b200 as nz
    with pytest.raises(nz.NdzipB200Error):
        nz.compressor_requirements([(4, 4), (4,)])          # reference src/ndzip/common.cc:12-15
    with pytest.raises(nz.NdzipB200Error, match="empty requirements"):
        nz.compressor_requirements().dimensions              # reference src/ndzip/common.hh:319-322
    with pytest.raises(nz.NdzipB200Error, match="dimensionality"):
        nz.num_hypercubes((2, 2, 2, 2))                      # reference src/ndzip/common.hh:642
    assert lib.ndzb_strerror(-2).decode().startswith("data dimensionality does not match")
    assert b"sm_100a" in lib.ndzb_version()


def test_host_placement_is_a_no_op_without_topology(lib):
    # ndzb_device_numa_node / ndzb_bind_host_to_device: no CUDA device, or a box that exposes no NUMA node for it
    # (numa_node = -1, like the pool's VMs): -1 and the calling thread's affinity is left alone
    import torch
    import ndzip_b200 as nz
    before = os.sched_getaffinity(0)
    if not torch.cuda.is_available():
        assert nz.device_numa_node(0) == -1 and nz.bind_host_to_device(0) == -1
        assert os.sched_getaffinity(0) == before
    assert nz.device_numa_node(10 ** 6) == -1 and nz.bind_host_to_device(10 ** 6) == -1   # no such device
    assert os.sched_getaffinity(0) == before


def test_container_header_arithmetic(lib):
    # host-only entry points of the sharded container (the data paths are covered in test_dist_cpu.py / container_test.c)
    from ndzip_b200 import _lib
    assert lib.ndzb_container_header_bytes(1) == 64 and lib.ndzb_container_header_bytes(2) == 80 and lib.ndzb_container_header_bytes(8) == 224
    info = _lib.ContainerInfo()
    dims, sz = _lib.size3((4096 * 4,))
    assert lib.ndzb_container_plan(0, dims, sz, 4, None, ctypes.byref(info), None) == 0
    assert info.header_bytes == info.total_bytes == lib.ndzb_container_header_bytes(4) and info.segments == 4
    assert lib.ndzb_container_plan(0, dims, sz, 0, None, ctypes.byref(info), None) == -1       # no segments
    assert lib.ndzb_container_plan(2, dims, sz, 4, None, ctypes.byref(info), None) == -1       # bad dtype
    assert lib.ndzb_container_decode_header(None, 0, ctypes.byref(info), None, 0) == -1
    assert lib.ndzb_strerror(-7) == b"file input / output failed"


def _chunk_plan(lib, dtype, shape, decompress):
    from ndzip_b200 import _lib
    dims, sz = _lib.size3(shape)
    n = ctypes.c_uint32(0)
    assert lib.ndzb_offload_chunk_plan(dtype, dims, sz, int(decompress), None, 0, ctypes.byref(n)) == 0
    if n.value == 0:
        return []
    rows = (ctypes.c_uint32 * (n.value + 1))()
    assert lib.ndzb_offload_chunk_plan(dtype, dims, sz, int(decompress), rows, n.value, ctypes.byref(n)) == -3   # capacity: n + 1 entries
    assert lib.ndzb_offload_chunk_plan(dtype, dims, sz, int(decompress), rows, n.value + 1, ctypes.byref(n)) == 0
    return list(rows)


@pytest.mark.parametrize("decompress", [False, True])
def test_offloader_chunk_plan_invariants(lib, decompress, monkeypatch):
    # ndzb_offload_chunk_plan = the plan ndzb_offload_compress / _decompress execute (csrc/ndzb_capi.cu plan_chunks): chunks of whole
    # cube rows that tile dimension 0 in order, never more than the 128 the per-chunk events / totals are sized for, single-row
    # pieces at the small end of the pipeline (compression: the back, decompression: the front)
    side = {1: 4096, 2: 64, 3: 16}
    cases = [(0, (512, 512, 512)), (1, (8192, 8192)), (1, (128, 1024, 1024)), (0, (1 << 28,)), (0, (1 << 24,)), (0, (1 << 31,)),
             (1, (1024, 1024, 1024)), (0, (64 * 300, 64 * 40)), (1, (16 * 7, 256, 256)), (0, (4096 * 1025,))]
    for chunk_env in (None, "4096", str(1 << 20), str(1 << 30)):
        if chunk_env is None:
            monkeypatch.delenv("NDZB_CHUNK_BYTES", raising=False)
        else:
            monkeypatch.setenv("NDZB_CHUNK_BYTES", chunk_env)
        for dtype, shape in cases:
            rows = _chunk_plan(lib, dtype, shape, decompress)
            cube_rows = shape[0] // side[len(shape)]
            if not rows:
                continue
            lens = [b - a for a, b in zip(rows, rows[1:])]
            assert rows[0] == 0 and rows[-1] == cube_rows and all(n > 0 for n in lens), (shape, rows)
            assert 2 <= len(lens) <= 127, (shape, len(lens))
            small = lens[0] if decompress else lens[-1]
            assert 2 * small <= max(lens) + 1 or max(lens) == 1, (shape, lens)  # the small end really is small
    monkeypatch.delenv("NDZB_CHUNK_BYTES", raising=False)
    # what the headline workload gets: 32 MiB chunks (2 cube rows of 16 MiB), the last one in two pieces
    assert [b - a for a, b in zip(*(lambda r: (r, r[1:]))(_chunk_plan(lib, 0, (512, 512, 512), False)))] == [2] * 15 + [1, 1]
    assert [b - a for a, b in zip(*(lambda r: (r, r[1:]))(_chunk_plan(lib, 0, (512, 512, 512), True)))] == [1] * 4 + [4] * 7
    # not pipelined: a border, or a small array
    assert _chunk_plan(lib, 0, (513, 512, 512), False) == [] and _chunk_plan(lib, 0, (64, 64, 64), False) == []
