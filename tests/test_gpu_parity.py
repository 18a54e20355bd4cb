"""-m gpu parity tests: the CUDA path (through the C ABI) against the oracle, the committed golden
fixtures and — where oracle/_ref/libndzip_ref.so travelled with the repo — the unmodified reference
CPU codec. Bit-exact everywhere: this is integer work.

Mirrors the reference's own parity tests (src/test/codec_profile_test.inl): identical
streams between encoders for one cube (:952-995), for bordered shapes (:37-140), 0-hypercube extents
(:1045-1082), decode(encode(x)) == x for every encoder/decoder pairing.
"""
import json
import os
import zlib

import numpy as np
import pytest

from ndzip_b200 import synth

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SIDE = {1: 4096, 2: 64, 3: 16}
PROFILES = [(dt, d) for dt in ("float32", "float64") for d in (1, 2, 3)]
PATHS = ["tma", "vec16", "scalar"]


def _rows():
    with open(os.path.join(ROOT, "tests", "golden", "golden.json")) as f:
        return json.load(f)["rows"]


def _row_id(r):
    return f"{r['generator']}-{r['dtype']}-{'x'.join(map(str, r['shape']))}"


@pytest.fixture(scope="module")
def nz():
    import torch
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import ndzip_b200
    from ndzip_b200 import _lib
    _lib.load()  # fails loudly if the extension is missing
    return ndzip_b200


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("row", _rows(), ids=_row_id)
def test_golden_fixture(nz, row, path):
    from gpu_util import gpu_compress, gpu_decompress, load_path
    shape = tuple(row["shape"])
    data = synth.make(row["generator"], shape, row["dtype"], **row["kwargs"])
    assert nz.compressed_length_bound(row["dtype"], shape) == row["bound"]
    with load_path(path):
        stream, _ = gpu_compress(data)
    assert stream.size == row["stream_words"]
    assert "%08x" % zlib.crc32(stream.tobytes()) == row["stream_crc32"]
    assert ["%x" % int(w) for w in stream[:4]] == row["first_words"]
    back = gpu_decompress(stream, row["dtype"], shape)
    assert back.tobytes() == data.tobytes()


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_golden_single_cube_streams(nz, dtype, dims):
    # word-by-word comparison against streams produced by the reference CPU encoder
    from gpu_util import gpu_compress
    cubes = np.load(os.path.join(ROOT, "tests", "golden", "cubes.npz"))
    data = synth.hashed((SIDE[dims],) * dims, dtype, seed=11)
    data.reshape(-1)[: (32 if dtype == "float32" else 64)] = 0
    expect = cubes[f"{dtype}_{dims}d_stream"]
    stream, _ = gpu_compress(data)
    assert stream.size == expect.size
    mismatch = np.nonzero(stream != expect)[0]
    assert mismatch.size == 0, f"first mismatching word {mismatch[:8]}"


@pytest.mark.parametrize("path", PATHS)
@pytest.mark.parametrize("gen", ["hashed", "raw_bits", "quantised", "poly", "smooth"])
@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_stream_equals_oracle(nz, oracle, dtype, dims, gen, path):
    from gpu_util import gpu_compress, gpu_decompress, load_path
    n = {1: 5 * 4096 + 123, 2: 4 * 64 - 1, 3: 4 * 16 - 1}[dims]  # bordered (codec_profile_test.inl:44-50)
    shape = (n,) * dims
    kw = {} if gen in ("poly",) else {"seed": 77}
    data = synth.make(gen, shape, dtype, **kw)
    data.reshape(-1)[: (32 if dtype == "float32" else 64)] = 0  # regression: first chunk zero (:54-58)
    expect = oracle.compress(data)
    with load_path(path):
        stream, full = gpu_compress(data)
    assert stream.size == expect.size
    assert np.array_equal(stream, expect)
    back = gpu_decompress(expect, dtype, shape)  # GPU decodes the CPU stream
    assert back.tobytes() == data.tobytes()
    back_cpu, consumed = oracle.decompress(stream, dtype, shape)  # CPU decodes the GPU stream
    assert consumed == stream.size and back_cpu.tobytes() == data.tobytes()


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_aligned_multi_cube_equals_reference(nz, oracle, dtype, dims):
    # no border, many cubes, TMA path; compared against the unmodified reference when present
    from oracle import get_reference
    from gpu_util import gpu_compress, gpu_decompress
    shape = {1: (64 * 4096,), 2: (512, 768), 3: (64, 96, 128)}[dims]
    data = synth.smooth(shape, dtype, seed=5)
    ref = get_reference()
    # threads=1: the reference's serial encoder is the parity oracle. Its OpenMP encoder is racy
    # (cpu_codec.inl:826-836 computes a chunk's destination from a header entry another thread may
    # not have written yet) and intermittently corrupts streams; see DESIGN.md.
    expect = ref.compress(data, threads=1) if ref is not None else oracle.compress(data)
    stream, _ = gpu_compress(data)
    assert np.array_equal(stream, expect)
    assert gpu_decompress(stream, dtype, shape).tobytes() == data.tobytes()


# (kernel, variant): compress_kernel ("v1") and every tuning variant of the warp-specialised kernel
KERNELS = [("v1", None)] + [("ws", v) for v in range(5)]


@pytest.mark.parametrize("kernel,variant", KERNELS, ids=[k if v is None else f"{k}{v}" for k, v in KERNELS])
@pytest.mark.parametrize("gen", ["hashed", "smooth", "zeros"])
@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_tma_compatible_bordered_shapes_all_kernels(nz, oracle, dtype, dims, gen, kernel, variant):
    # bordered shapes whose row pitch is still 16-byte aligned, so the TMA input path (and with it the
    # warp-specialised kernel) is taken: incompressible cubes (longest images), smooth data, and all-zero
    # cubes (shortest images: heads only)
    from gpu_util import gpu_compress, gpu_decompress, compress_kernel
    shape = {1: (9 * 4096 + 123,), 2: (200, 260), 3: (40, 52, 68)}[dims]
    data = np.zeros(shape, dtype) if gen == "zeros" else synth.make(gen, shape, dtype, seed=11)
    expect = oracle.compress(data)
    with compress_kernel(kernel, variant):
        stream, _ = gpu_compress(data)
    assert stream.size == expect.size
    assert np.array_equal(stream, expect)
    assert gpu_decompress(stream, dtype, shape).tobytes() == data.tobytes()


@pytest.mark.parametrize("kernel", ["v1", "ws"])
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_many_incompressible_cubes_and_misaligned_stream_pointer(nz, oracle, dtype, kernel):
    # several cubes per encoder group and SM, all of maximum length; the stream buffer starts one word
    # into an allocation, so neither the header nor the cubes are 16-byte aligned
    import torch
    from gpu_util import to_device, compress_kernel
    shape = (560 * 4096,) if dtype == "float32" else (300 * 4096,)
    data = synth.make("hashed", shape, dtype, seed=3)
    expect = oracle.compress(data)
    tbits = torch.int32 if dtype == "float32" else torch.int64
    with compress_kernel(kernel, None):
        comp = nz.make_cuda_compressor(dtype, nz.compressor_requirements(shape))
        backing = torch.zeros(nz.compressed_length_bound(dtype, shape) + 1, dtype=tbits, device="cuda")
        d_stream = backing[1:]
        d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
        comp.compress(to_device(data), shape, d_stream, d_len)
        torch.cuda.synchronize()
    n = int(d_len.cpu().numpy().view(np.uint32)[0])
    assert n == expect.size
    got = d_stream[:n].cpu().numpy().view(expect.dtype)
    assert np.array_equal(got, expect)
    assert int(backing[0].item()) == 0


def test_repeated_launches_on_one_dimensional_input_stay_identical(nz, monkeypatch):
    # regression: 1-D inputs have the shortest cube period; a TMA load slower than one round of the slot ring
    # let an encoder group pass its mbarrier parity wait one phase early (about one launch in 30 at this size),
    # the cube's length was never published and the kernel spun. NDZB_WS_CHECK turns a raised watchdog into
    # an error instead of a trap, so a recurrence fails this test instead of poisoning the CUDA context.
    import torch
    monkeypatch.setenv("NDZB_WS_CHECK", "1")
    shape = (16384 * 4096,)
    d_in = torch.from_numpy(synth.smooth((1 << 20,), "float32", seed=9)).cuda().repeat(64)
    d_in += torch.arange(d_in.numel(), device="cuda", dtype=torch.float32) * 1e-7  # no two cubes alike
    comp = nz.make_cuda_compressor("float32", nz.compressor_requirements(shape))
    d_stream = torch.zeros(nz.compressed_length_bound("float32", shape), dtype=torch.int32, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
    first = None
    for i in range(40):
        comp.compress(d_in, shape, d_stream, d_len)
        torch.cuda.synchronize()
        n = int(d_len.item())
        if i in (0, 1, 39):
            digest = (n, zlib.crc32(d_stream[:n].cpu().numpy().tobytes()))
            first = first or digest
            assert digest == first
    back = torch.zeros_like(d_in)
    nz.make_cuda_decompressor("float32", 1).decompress(d_stream, back, shape)
    torch.cuda.synchronize()
    assert torch.equal(back.view(torch.int32), d_in.view(torch.int32))


@pytest.mark.parametrize("dtype,dims", PROFILES)
@pytest.mark.parametrize("n", [0, 1])
def test_zero_hypercube_extents(nz, oracle, dtype, dims, n):
    # reference src/test/codec_profile_test.inl:1045-1082
    from gpu_util import gpu_compress, gpu_decompress
    shape = (n,) * dims
    data = synth.ramp(shape, dtype) + np.dtype(dtype).type(1.5)
    stream, _ = gpu_compress(data)
    assert np.array_equal(stream, oracle.compress(data))
    assert stream.size == data.size
    assert gpu_decompress(stream, dtype, shape).tobytes() == data.tobytes()


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_stream_equals_reference_cuda_encoder(nz, dtype, dims):
    # second, GPU-side oracle: the reference's own CUDA encoder recompiled for sm_100a (when prebuilt)
    import torch
    from oracle import ReferenceCuda
    if not ReferenceCuda.available():
        pytest.skip("oracle/_ref/libndzip_refcuda.so not prebuilt")
    shape = {1: (7 * 4096 + 33,), 2: (4 * 64 - 1, 3 * 64 + 5), 3: (63, 40, 50)}[dims]
    data = synth.smooth(shape, dtype, seed=15)
    tbits = torch.int32 if dtype == "float32" else torch.int64
    d_in = torch.from_numpy(data).cuda()
    bound = nz.compressed_length_bound(dtype, shape)
    ours = torch.zeros(bound, dtype=tbits, device="cuda")
    theirs = torch.zeros(bound, dtype=tbits, device="cuda")
    n_ours = torch.zeros(1, dtype=torch.int32, device="cuda")
    n_theirs = torch.zeros(1, dtype=torch.int32, device="cuda")
    nz.make_cuda_compressor(dtype, shape).compress(d_in, shape, ours, n_ours)
    ref = ReferenceCuda(dtype, shape)
    ref.compress(d_in.data_ptr(), theirs.data_ptr(), n_theirs.data_ptr())
    torch.cuda.synchronize()
    n = int(n_ours.item())
    assert n == int(n_theirs.item())
    assert torch.equal(ours[:n], theirs[:n])
    # cross pairing: their decoder on our stream, our decoder on their stream
    back_a = torch.zeros_like(d_in)
    back_b = torch.zeros_like(d_in)
    ref.decompress(ours.data_ptr(), back_a.data_ptr())
    nz.make_cuda_decompressor(dtype, dims).decompress(theirs, back_b, shape)
    torch.cuda.synchronize()
    assert torch.equal(back_a.view(tbits), d_in.view(tbits))
    assert torch.equal(back_b.view(tbits), d_in.view(tbits))


def test_header_padding_word_is_zero(nz):
    # f64 with an odd cube count: the GPU encoders write 0 to the padding word (cuda_codec.inl:446-452)
    from gpu_util import gpu_compress
    data = synth.hashed((3 * 4096,), "float64", seed=2)
    stream, full = gpu_compress(data, fill=0xFF)
    header = stream[:2].view(np.uint32)
    assert header[3] == 0
    assert header[2] == stream.size - 2


def test_context_reuse_and_varying_extents(nz, oracle):
    # one compressor object, many calls (epoch / ticket bookkeeping), as the CLI does per chunk
    from gpu_util import gpu_compress
    req = nz.compressor_requirements([(96, 96, 96), (33, 50, 70), (16, 16, 16)])
    comp = nz.make_cuda_compressor("float32", req)
    for rep in range(3):
        for shape in [(96, 96, 96), (16, 16, 16), (33, 50, 70), (48, 64, 80)]:
            data = synth.hashed(shape, "float32", seed=10 + rep)
            stream, _ = gpu_compress(data, compressor=comp)
            assert np.array_equal(stream, oracle.compress(data)), (rep, shape)


def test_context_grows_beyond_requirements(nz, oracle):
    from gpu_util import gpu_compress
    comp = nz.make_cuda_compressor("float64", nz.compressor_requirements((64, 64)))
    data = synth.poly((256, 320), "float64")
    stream, _ = gpu_compress(data, compressor=comp)
    assert np.array_equal(stream, oracle.compress(data))


def test_length_pointer_is_optional(nz, oracle):
    from gpu_util import gpu_compress
    data = synth.quantised((130, 200), "float32", seed=4)
    _, full = gpu_compress(data, want_length_tensor=False)
    expect = oracle.compress(data)
    assert np.array_equal(full[: expect.size], expect)


def test_unaligned_device_pointer_uses_fallback_path(nz, oracle):
    import torch
    import ndzip_b200 as nzb
    shape = (32, 32, 48)
    data = synth.hashed(shape, "float32", seed=6)
    backing = torch.zeros(data.size + 1, dtype=torch.float32, device="cuda")
    view = backing[1:]  # 4-byte aligned only: TMA and 16-byte loads are impossible
    view.copy_(torch.from_numpy(data.reshape(-1)))
    bound = nzb.compressed_length_bound("float32", shape)
    out_backing = torch.zeros(bound + 1, dtype=torch.int32, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
    comp = nzb.make_cuda_compressor("float32", shape)
    comp.compress(view, shape, out_backing[1:], d_len)
    torch.cuda.synchronize()
    n = int(d_len.item())
    got = out_backing[1:1 + n].cpu().numpy().view(np.uint32)
    assert np.array_equal(got, oracle.compress(data))
    dec = nzb.make_cuda_decompressor("float32", 3)
    back = torch.zeros(data.size + 1, dtype=torch.float32, device="cuda")
    dec.decompress(out_backing[1:], back[1:], shape)
    torch.cuda.synchronize()
    assert back[1:].cpu().numpy().tobytes() == data.tobytes()


def test_dimension_mismatch_raises(nz):
    import torch
    comp = nz.make_cuda_compressor("float32", (64, 64))
    x = torch.zeros(4096, device="cuda")
    out = torch.zeros(8192, dtype=torch.int32, device="cuda")
    with pytest.raises(nz.NdzipB200Error, match="dimensionality"):
        comp.compress(x, (4096,), out, None)  # reference cuda_codec.inl:557-559
    with pytest.raises(nz.NdzipB200Error):
        nz.compressor_requirements([(4, 4), (4,)])  # reference common.cc:12-15
    with pytest.raises(nz.NdzipB200Error, match="empty requirements"):
        nz.make_cuda_compressor("float32", nz.compressor_requirements())


@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_offloader_host_api(nz, oracle, dtype, dims):
    # reference include/ndzip/offload.hh:16-24; tests use it at codec_profile_test.inl:64-93
    shape = {1: (3 * 4096 + 17,), 2: (200, 130), 3: (40, 33, 50)}[dims]
    data = synth.smooth(shape, dtype, seed=9)
    off = nz.make_cuda_offloader(dtype, dims)
    bits = np.uint32 if dtype == "float32" else np.uint64
    stream = np.zeros(nz.compressed_length_bound(dtype, shape), dtype=bits)
    n = off.compress(data, shape, stream)
    assert off.kernel_duration_ns is not None and off.kernel_duration_ns > 0
    expect = oracle.compress(data)
    assert n == expect.size and np.array_equal(stream[:n], expect)
    back = np.zeros(shape, dtype=dtype)
    consumed = off.decompress(stream, n, back, shape)
    assert consumed == n
    assert back.tobytes() == data.tobytes()


@pytest.mark.parametrize("dtype,shape", [("float32", (96, 64, 80)), ("float64", (5 * 4096,)), ("float32", (320, 192)),
                                         ("float64", (64, 48, 32))])
def test_pipelined_offloader_path(nz, oracle, dtype, shape, monkeypatch):
    # border-free extents take the chunked three-stream path; force small chunks so
    # that several chained launches (running base offset across launches) are exercised
    import torch
    monkeypatch.setenv("NDZB_PIPELINE_MIN_BYTES", "1")
    side = {1: 4096, 2: 64, 3: 16}[len(shape)]
    row_bytes = side * int(np.prod(shape[1:])) * np.dtype(dtype).itemsize if len(shape) > 1 else side * np.dtype(dtype).itemsize
    monkeypatch.setenv("NDZB_CHUNK_BYTES", str(2 * row_bytes))
    data = synth.smooth(shape, dtype, seed=21)
    expect = oracle.compress(data)
    bits = np.uint32 if dtype == "float32" else np.uint64
    off = nz.make_cuda_offloader(dtype, len(shape))
    for pinned in (False, True):
        if pinned:
            h_in = torch.from_numpy(data.copy()).pin_memory()
            h_stream = torch.zeros(nz.compressed_length_bound(dtype, shape), dtype=torch.int32 if dtype == "float32" else torch.int64).pin_memory()
            h_back = torch.zeros(shape, dtype=h_in.dtype).pin_memory()
            n = off.compress(h_in, shape, h_stream)
            got = h_stream.numpy().view(bits)[:n]
        else:
            stream = np.zeros(nz.compressed_length_bound(dtype, shape), dtype=bits)
            n = off.compress(data, shape, stream)
            got = stream[:n]
        assert off.last_launch_count >= 2, "expected the chunked path"
        assert n == expect.size and np.array_equal(got, expect)
        if pinned:
            consumed = off.decompress(h_stream, n, h_back, shape)
            back = h_back.numpy()
        else:
            back = np.zeros(shape, dtype=dtype)
            consumed = off.decompress(stream, n, back, shape)
        assert consumed == n and back.tobytes() == data.tobytes()
        assert off.kernel_duration_ns > 0


def test_sharded_cube_ranges_stitch_to_the_full_stream(nz, oracle):
    # single-GPU check of the multi-GPU building blocks
    import torch
    shape = (64, 48, 80)
    dtype = "float32"
    data = synth.smooth(shape, dtype, seed=12)
    expect = oracle.compress(data)
    H = nz.num_hypercubes(shape)
    d_in = torch.from_numpy(data).cuda()
    comp = nz.make_cuda_compressor(dtype, shape)
    bounds = [0, H // 3, H // 3, H - 5, H]
    pieces, offsets, totals = [], [], []
    for b, e in zip(bounds[:-1], bounds[1:]):
        cubes = torch.zeros(max(1, (e - b) * 4224), dtype=torch.int32, device="cuda")
        offs = torch.zeros(max(1, e - b), dtype=torch.int32, device="cuda")
        tot = torch.full((1,), -1, dtype=torch.int32, device="cuda")
        comp.compress_cubes(d_in.data_ptr(), shape, b, e, cubes, offs, tot)
        torch.cuda.synchronize()
        t = int(tot.item())
        pieces.append(cubes[:t].cpu().numpy().view(np.uint32))
        offsets.append(offs[: e - b].cpu().numpy().view(np.uint32))
        totals.append(t)
    base = np.concatenate([[0], np.cumsum(totals)[:-1]]).astype(np.uint32)
    header = np.concatenate([o + b for o, b in zip(offsets, base)])
    stitched = np.concatenate([header] + pieces)
    assert np.array_equal(stitched, expect[: stitched.size])
    # ranged decompression of the full stream
    dec = nz.make_cuda_decompressor(dtype, 3)
    d_stream = torch.from_numpy(expect.view(np.int32).copy()).cuda()
    out = torch.zeros(data.size, dtype=torch.float32, device="cuda")
    for b, e in zip(bounds[:-1], bounds[1:]):
        dec.decompress_cubes(d_stream, out.data_ptr(), shape, b, e)
    torch.cuda.synchronize()
    assert out.cpu().numpy().tobytes() == data.tobytes()


# ------------------------------------------------------------------ BASELINE.json sizes
# Full-size configs: stream equality against the unmodified reference (multi-threaded) when it
# travelled with the repo, plus size-independent properties otherwise.

BASELINE_CASES = [
    ("float32", (1 << 24,)),            # config 1: 1D fp32 16 Mi
    ("float32", (512, 512, 512)),       # config 2: 3D fp32 512^3
    ("float64", (8192, 8192)),          # config 3: 2D fp64 8192^2
]


@pytest.mark.parametrize("dtype,shape", BASELINE_CASES, ids=["cfg1-1d-f32-16Mi", "cfg2-3d-f32-512", "cfg3-2d-f64-8192"])
def test_baseline_config_roundtrip_and_reference_parity(nz, dtype, shape):
    import torch
    from bench import make_device_input  # same generator the benchmark uses
    from oracle import get_reference
    d_in = make_device_input(dtype, shape, seed=0x5EED0002)
    bound = nz.compressed_length_bound(dtype, shape)
    tbits = torch.int32 if dtype == "float32" else torch.int64
    d_stream = torch.empty(bound, dtype=tbits, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
    comp = nz.make_cuda_compressor(dtype, shape)
    comp.compress(d_in, shape, d_stream, d_len)
    torch.cuda.synchronize()
    n = int(d_len.cpu().numpy().view(np.uint32)[0])
    H = nz.num_hypercubes(shape)
    hdr_words = H if dtype == "float32" else (H + 1) // 2
    header = d_stream[:hdr_words].cpu().numpy().view(np.uint32)[:H].astype(np.int64)
    # properties: offsets strictly increase by [C, bound] per cube; length = header + last offset
    steps = np.diff(np.concatenate([[0], header]))
    C, cube_bound = (128, 4224) if dtype == "float32" else (64, 4160)
    assert steps.min() >= C and steps.max() <= cube_bound
    assert n == hdr_words + int(header[-1])
    # round trip on the device
    dec = nz.make_cuda_decompressor(dtype, len(shape))
    d_back = torch.empty_like(d_in)
    dec.decompress(d_stream, d_back, shape)
    torch.cuda.synchronize()
    assert torch.equal(d_in.view(tbits), d_back.view(tbits))
    # second compression into a fresh buffer is bit-identical (determinism despite dynamic scheduling)
    d_stream2 = torch.empty(bound, dtype=tbits, device="cuda")
    comp.compress(d_in, shape, d_stream2, d_len)
    torch.cuda.synchronize()
    assert torch.equal(d_stream[:n], d_stream2[:n])
    ref = get_reference()
    if ref is not None:
        host = d_in.cpu().numpy()
        bits = np.uint32 if dtype == "float32" else np.uint64
        expect = np.zeros(bound, dtype=bits)
        n_ref = ref.compress_into(host, expect, threads=1)  # serial encoder = parity oracle (OpenMP one is racy)
        assert n_ref == n
        got = d_stream[:n].cpu().numpy().view(bits)
        assert zlib.crc32(got.tobytes()) == zlib.crc32(expect[:n].tobytes())
        assert np.array_equal(got, expect[:n])


# Configs 4 and 5 of BASELINE.json at their full single-GPU-resident sizes (8 GiB of input): too large
# for a CPU oracle pass inside a test, so they are checked through size-independent properties —
# exact device round trip, header monotonicity within [C, bound], stream length = header + last offset,
# run-to-run determinism — plus stream equality against the oracle on the leading cube rows.
LARGE_CASES = [
    ("float64", (1024, 1024, 1024)),    # config 4: 3D fp64 1024^3 (8 GiB)
    ("float32", (1 << 31,)),            # config 5: 1D fp32 2 Gi elements (8 GiB)
]


@pytest.mark.parametrize("dtype,shape", LARGE_CASES, ids=["cfg4-3d-f64-1024", "cfg5-1d-f32-2Gi"])
def test_large_configs_properties(nz, oracle, dtype, shape):
    import torch
    from bench import make_device_input
    torch.cuda.empty_cache()  # mem_get_info does not see what torch's caching allocator holds from earlier tests
    free, _ = torch.cuda.mem_get_info()
    itemsize = np.dtype(dtype).itemsize
    n_bytes = int(np.prod(shape)) * itemsize
    bound = nz.compressed_length_bound(dtype, shape)
    need = 2 * n_bytes + 2 * bound * itemsize + (2 << 30)
    if free < need:
        pytest.skip(f"needs {need >> 30} GiB of device memory")
    assert bound < 2 ** 32
    tbits = torch.int32 if dtype == "float32" else torch.int64
    d_in = make_device_input(dtype, shape, seed=0x5EED0004)
    d_stream = torch.empty(bound, dtype=tbits, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
    comp = nz.make_cuda_compressor(dtype, shape)
    comp.compress(d_in, shape, d_stream, d_len)
    torch.cuda.synchronize()
    n = int(d_len.cpu().numpy().view(np.uint32)[0])
    H = nz.num_hypercubes(shape)
    hdr_words = H if dtype == "float32" else (H + 1) // 2
    header = d_stream[:hdr_words].cpu().numpy().view(np.uint32)[:H].astype(np.int64)
    steps = np.diff(np.concatenate([[0], header]))
    C, cube_bound = (128, 4224) if dtype == "float32" else (64, 4160)
    assert steps.min() >= C and steps.max() <= cube_bound
    assert n == hdr_words + int(header[-1])
    # leading cube rows against the oracle (cube contents do not depend on the rest of the grid)
    rows = {1: 4096 * 8, 3: 16}[len(shape)]
    lead_shape = (rows,) + tuple(shape[1:])
    lead = d_in[:rows].cpu().numpy()
    expect = oracle.compress(lead)
    H_lead = nz.num_hypercubes(lead_shape)
    hdr_lead = H_lead if dtype == "float32" else (H_lead + 1) // 2
    bits = np.uint32 if dtype == "float32" else np.uint64
    assert np.array_equal(header[:H_lead].astype(np.uint32), expect[:hdr_lead].view(np.uint32)[:H_lead])
    words_lead = int(header[H_lead - 1])
    got = d_stream[hdr_words: hdr_words + words_lead].cpu().numpy().view(bits)
    assert np.array_equal(got, expect[hdr_lead: hdr_lead + words_lead])
    del lead, expect, got
    # exact round trip on the device
    d_back = torch.empty_like(d_in)
    nz.make_cuda_decompressor(dtype, len(shape)).decompress(d_stream, d_back, shape)
    torch.cuda.synchronize()
    assert torch.equal(d_in.view(tbits), d_back.view(tbits))
    del d_back
    # determinism
    d_stream2 = torch.empty(bound, dtype=tbits, device="cuda")
    comp.compress(d_in, shape, d_stream2, d_len)
    torch.cuda.synchronize()
    assert int(d_len.cpu().numpy().view(np.uint32)[0]) == n
    assert torch.equal(d_stream[:n], d_stream2[:n])
    del d_stream2
    # the WHOLE 8 GiB stream against the reference's own CUDA encoder (its device API is safe up to 2^32 - 1 elements,
    # SURVEY.md §3.3; the CPU reference would take minutes here), and the reference decoding OUR stream
    from oracle import ReferenceCuda
    if ReferenceCuda.available():
        torch.cuda.empty_cache()
        free, _ = torch.cuda.mem_get_info()
        scratch = nz.num_hypercubes(shape) * (4096 + 128) * itemsize  # the reference's chunk scratch, cuda_codec.inl:543-552
        if free > bound * itemsize + scratch + n_bytes + (4 << 30):
            rc = ReferenceCuda(dtype, shape)
            r_stream = torch.zeros(bound, dtype=tbits, device="cuda")
            r_len = torch.zeros(1, dtype=torch.int32, device="cuda")
            rc.compress(d_in.data_ptr(), r_stream.data_ptr(), r_len.data_ptr())
            torch.cuda.synchronize()
            assert int(r_len.cpu().numpy().view(np.uint32)[0]) == n
            assert torch.equal(r_stream[:n], d_stream[:n])
            del r_stream
            r_back = torch.empty_like(d_in)
            rc.decompress(d_stream.data_ptr(), r_back.data_ptr())
            torch.cuda.synchronize()
            assert torch.equal(r_back.view(tbits), d_in.view(tbits))


@pytest.mark.parametrize("dtype,shape", [("float32", (96, 64, 80)), ("float64", (300, 200)), ("float32", (9 * 4096 + 77,))])
def test_sharded_container_of_gpu_streams(nz, dtype, shape):
    # every "rank" (here: one GPU, slab after slab) compresses its slab into a self-contained stream;
    # the container needs no offset exchange and no gather, and converts to the single-GPU stream bit for bit
    from gpu_util import gpu_compress, gpu_decompress
    from ndzip_b200 import dist as nzd
    data = synth.smooth(shape, dtype, seed=13)
    spans = nzd.slab_partition(shape, 3)
    local = [gpu_compress(np.ascontiguousarray(data[b:e]))[0] for b, e in spans]
    from oracle import get_oracle
    for (b, e), got in zip(spans, local):  # the GPU's slab streams are the oracle's, not merely self-consistent
        assert np.array_equal(got, get_oracle().compress(np.ascontiguousarray(data[b:e])))
    buf = nzd.pack_sharded(dtype, shape, local)
    hdr = nzd.decode_sharded_header(buf)
    for i, (b, e) in enumerate(spans):
        back = gpu_decompress(nzd.sharded_segment(buf, hdr, i), dtype, hdr.slab_shape(i))
        assert back.tobytes() == data[b:e].tobytes()
    whole, _ = gpu_compress(data)
    assert np.array_equal(nzd.to_global_stream(buf), whole)
    assert np.array_equal(whole, get_oracle().compress(data))
    # sharded decompression through the C ABI: ndzb_container_decompress_segment reads the segment out of the container
    off = nz.make_cuda_offloader(dtype, len(shape))
    for i, (b, e) in enumerate(spans):
        out = np.zeros(hdr.slab_shape(i), dtype=dtype)
        nzd.decompress_segment(off, buf, i, out)
        assert out.tobytes() == data[b:e].tobytes()
    # a container whose segment is cut short fails before anything reaches the device
    with pytest.raises(nz.NdzipB200Error):
        nzd.decompress_segment(off, buf[: hdr.segments[-1].byte_offset + 8], len(spans) - 1, np.zeros(hdr.slab_shape(len(spans) - 1), dtype=dtype))


# ---- reference "Residual encodings equivalent" (src/test/codec_profile_test.inl:552-729): a cube whose RESIDUALS are
# engineered to contain all-zero bit columns, all-zero words and fully dense chunks (pattern at :561-567). The GPU kernel
# fuses transform and encoding, so the input is the oracle's inverse transform of that cube: its residuals are then the
# engineered cube exactly, and the cube's stream must be the oracle's zero-bit encoding of it, word for word.

@pytest.mark.parametrize("path", ["tma", "vec16", "scalar"])
@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_engineered_residual_cube_encodes_like_the_oracle(nz, oracle, dtype, dims, path):
    from gpu_util import gpu_compress, gpu_decompress, load_path
    bits = np.uint32 if dtype == "float32" else np.uint64
    side = {1: 4096, 2: 64, 3: 16}[dims]
    eng = synth.engineered_cube(bits)
    values = oracle.block_transform(eng, dims, inverse=True)           # value domain (bit patterns, may be NaNs)
    assert np.array_equal(oracle.block_transform(values, dims), eng)   # the transform is a bijection
    data = values.view(dtype).reshape((side,) * dims)
    with load_path(path):
        stream, _ = gpu_compress(data)
        back = gpu_decompress(stream, dtype, data.shape)   # the decoder's output path follows `path` too
    assert back.tobytes() == data.tobytes()
    encoded = oracle.zero_bit_encode(eng)
    hdr = 1  # one cube: one offset word (f64: offset + padding in one 64-bit word)
    assert stream.size == hdr + encoded.size
    assert int(stream[:1].view(np.uint32)[0]) == encoded.size
    assert np.array_equal(stream[hdr:], encoded)
    assert np.array_equal(stream, oracle.compress(data))
    back = gpu_decompress(stream, dtype, data.shape)
    assert back.tobytes() == data.tobytes()


# ---- host pointers beyond 4 GiB: the reference's offloader computes byte sizes in 32 bits and re-allocates per call
# (src/ndzip/cuda_codec.inl:679-685); ndzb_offload_* uses 64-bit sizes and the chunked three-stream pipeline. The stream
# it returns must be the device API's, and the round trip exact.

def test_offloader_beyond_four_gib(nz):
    import torch
    n = (1 << 30) + 3 * 4096                       # 1D f32, 4 GiB + 48 KiB, 262,147 hypercubes
    shape = (n,)
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    bound = nz.compressed_length_bound("float32", shape)
    if free < 3 * 4 * n + 2 * 4 * bound + (4 << 30):
        pytest.skip("needs ~30 GiB of device memory")
    try:
        h_in = torch.empty(n, dtype=torch.float32, pin_memory=True)
        h_stream = torch.empty(bound, dtype=torch.int32, pin_memory=True)
        h_back = torch.empty(n, dtype=torch.float32, pin_memory=True)
    except RuntimeError:
        pytest.skip("cannot pin 13 GiB of host memory on this box")
    from bench import make_device_input
    d_in = make_device_input("float32", shape, seed=0x5EED0005)
    h_in.copy_(d_in)
    off = nz.make_cuda_offloader("float32", 1)
    words = off.compress(h_in, shape, h_stream)
    assert off.kernel_duration_ns and off.kernel_duration_ns > 0
    consumed = off.decompress(h_stream, words, h_back, shape)
    assert consumed == words
    assert torch.equal(h_back.view(torch.int32), h_in.view(torch.int32))
    # against the device-pointer API on the same data
    d_stream = torch.empty(bound, dtype=torch.int32, device="cuda")
    d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
    nz.make_cuda_compressor("float32", shape).compress(d_in, shape, d_stream, d_len)
    torch.cuda.synchronize()
    assert int(d_len.cpu().numpy().view(np.uint32)[0]) == words
    assert torch.equal(d_stream[:words].cpu(), h_stream[:words])


def test_offloader_rejects_truncated_and_corrupt_streams(nz, oracle):
    # the header is validated on the host before anything reaches the device (include/ndzip_b200.h,
    # NDZB_ERR_CORRUPT_STREAM): a stream cut short, a header that runs backwards, a cube longer than its bound
    from ndzip_b200 import NdzipB200Error
    data = synth.smooth((3 * 4096 + 5,), "float32", seed=3)
    stream = oracle.compress(data)
    off = nz.make_cuda_offloader("float32", 1)
    out = np.empty_like(data)
    assert off.decompress(stream, stream.size, out, data.shape) == stream.size and out.tobytes() == data.tobytes()
    for bad_len in (0, 2, stream.size - 1, stream.size - 6):
        with pytest.raises(NdzipB200Error, match="ends inside a stream"):
            off.decompress(stream, bad_len, out, data.shape)
    backwards = stream.copy()
    backwards[1] = backwards[0] - 1
    with pytest.raises(NdzipB200Error, match="corrupt"):
        off.decompress(backwards, backwards.size, out, data.shape)
    too_long = stream.copy()
    too_long[0] = 5000
    with pytest.raises(NdzipB200Error, match="corrupt"):
        off.decompress(too_long, too_long.size, out, data.shape)
    # the library is still usable afterwards (no sticky CUDA error)
    assert off.decompress(stream, stream.size, out, data.shape) == stream.size and out.tobytes() == data.tobytes()


def test_misaligned_stream_pointer_with_a_short_header(nz, oracle):
    # two cubes behind a two-word header in a buffer that starts 4 bytes past a 16-byte boundary: the decoder's bulk copy
    # must not round the first cube's source down to in front of the buffer (reference API: word alignment only)
    import torch
    from gpu_util import to_device
    data = synth.smooth((2 * 4096,), "float32", seed=9)
    expect = oracle.compress(data)
    backing = torch.zeros(expect.size + 8, dtype=torch.int32, device="cuda")
    for shift in (1, 2, 3):
        d_stream = backing[shift: shift + expect.size]
        d_stream.copy_(torch.from_numpy(expect.view(np.int32)).cuda())
        d_out = torch.zeros(data.shape, dtype=torch.float32, device="cuda")
        nz.make_cuda_decompressor("float32", 1).decompress(d_stream, d_out, data.shape)
        torch.cuda.synchronize()
        assert d_out.cpu().numpy().tobytes() == data.tobytes()


# ---- scan primitives in isolation (reference src/test/cuda_bits_test.cu:37-114 tests its warp / hierarchical scans the
# same way): the compress kernel's decoupled look-back and the warp scan, driven by plain numbers

@pytest.mark.parametrize("mode", [0, 1, 2], ids=["two-level", "window32", "window64"])
@pytest.mark.parametrize("count", [1, 31, 32, 33, 1000, 70001])
def test_lookback_primitive_is_an_exclusive_scan(nz, mode, count):
    import torch
    from ndzip_b200 import _lib
    rng = np.random.default_rng(count * 3 + mode)
    lengths = rng.integers(0, 4225, size=count, dtype=np.uint32)
    lengths[rng.integers(0, count, size=max(1, count // 7))] = 0
    base = 12345
    comp = nz.make_cuda_compressor("float32", (4096 * 4,))       # small context: the self test grows the descriptors
    d_len = torch.from_numpy(lengths.view(np.int32)).cuda()
    d_out = torch.full((count,), -1, dtype=torch.int32, device="cuda")
    for _ in range(3):                                            # epochs: stale descriptors must read as invalid
        _lib.check(comp._lib.ndzb_selftest_lookback(comp._handle, mode, d_len.data_ptr(), count, base, d_out.data_ptr()))
        torch.cuda.synchronize()
        expect = (base + np.concatenate([[0], np.cumsum(lengths.astype(np.uint64))[:-1]])).astype(np.uint32)
        assert np.array_equal(d_out.cpu().numpy().view(np.uint32), expect)
        d_out.fill_(-1)
    # the context still compresses correctly afterwards (descriptors, tickets and block words left consistent)
    from oracle import get_oracle
    data = synth.smooth((4096 * 4,), "float32", seed=2)
    d_stream = torch.zeros(nz.compressed_length_bound("float32", data.shape), dtype=torch.int32, device="cuda")
    d_n = torch.zeros(1, dtype=torch.int32, device="cuda")
    comp.compress(torch.from_numpy(data).cuda(), data.shape, d_stream, d_n)
    torch.cuda.synchronize()
    n = int(d_n.item())
    assert np.array_equal(d_stream[:n].cpu().numpy().view(np.uint32), get_oracle().compress(data))


def test_warp_scan_primitive(nz):
    import torch
    from ndzip_b200 import _lib
    rng = np.random.default_rng(5)
    v = rng.integers(0, 65, size=1000, dtype=np.uint32)
    comp = nz.make_cuda_compressor("float32", (4096,))
    d_in = torch.from_numpy(v.view(np.int32)).cuda()
    d_out = torch.zeros_like(d_in)
    _lib.check(comp._lib.ndzb_selftest_warp_scan(comp._handle, d_in.data_ptr(), d_out.data_ptr(), v.size))
    torch.cuda.synchronize()
    pad = np.zeros(1024, dtype=np.uint32)
    pad[:1000] = v
    expect = np.cumsum(pad.reshape(-1, 32), axis=1).reshape(-1)[:1000]
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), expect.astype(np.uint32))


@pytest.mark.parametrize("store", ["tma", "vec16", "scalar"])
@pytest.mark.parametrize("gen", ["hashed", "smooth", "zeros"])
@pytest.mark.parametrize("dtype,dims", PROFILES)
def test_decoder_output_paths_on_tma_compatible_shapes(nz, oracle, dtype, dims, gen, store):
    # the three ways the decoded cube reaches global memory (TMA tensor store of the tile / vector stores / element-wise)
    # on bordered shapes whose pitch is 16-byte aligned, decoding the ORACLE's stream; the untouched border of the
    # output buffer around the cubes must come from the stream's border section, not from the tensor store
    from gpu_util import gpu_decompress, load_path
    shape = {1: (7 * 4096 + 124,), 2: (150, 196), 3: (36, 52, 40)}[dims]
    data = np.zeros(shape, dtype) if gen == "zeros" else synth.make(gen, shape, dtype, seed=23)
    stream = oracle.compress(data)
    with load_path(store):
        back = gpu_decompress(stream, dtype, shape)
    assert back.tobytes() == data.tobytes()


@pytest.mark.parametrize("decoder", ["ws", "v1"])
@pytest.mark.parametrize("gen", ["hashed", "smooth", "zeros"])
@pytest.mark.parametrize("dims", [2, 3])
def test_both_float_decoders_with_tensor_store(nz, oracle, dims, gen, decoder, monkeypatch):
    # float 2-D / 3-D with a TMA-addressable output: decompress_ws_kernel (persistent CTA, loader warp + 7 decode groups over
    # a slot ring; default) and decompress_kernel (NDZB_DECOMPRESS_KERNEL=v1) must both reproduce the data from the ORACLE's
    # stream; many cubes per CTA, a border, and a stream whose cubes end at every alignment
    from gpu_util import gpu_decompress
    monkeypatch.setenv("NDZB_DECOMPRESS_KERNEL", decoder)
    shape = {2: (64 * 37 + 12, 64 * 9 + 4), 3: (16 * 11 + 4, 16 * 13, 16 * 9 + 8)}[dims]
    data = np.zeros(shape, "float32") if gen == "zeros" else synth.make(gen, shape, "float32", seed=31)
    stream = oracle.compress(data)
    back = gpu_decompress(stream, "float32", shape)
    assert back.tobytes() == data.tobytes()


def test_paired_tickets_keep_the_context_in_step(nz, oracle):
    # 3-D float loaders draw tickets in pairs; the host mirrors the device's free-running ticket counter (2 * grid - (count & 1)
    # tickets beyond `count` per launch). Odd and even cube counts, one cube, more cubes than SMs, back to back on ONE context:
    # a miscounted launch would leave the next one waiting for tickets that were never drawn (or skipping cubes).
    import torch
    from gpu_util import to_device
    shapes = [(48, 48, 48), (16, 16, 16), (16, 16, 80), (32, 48, 64), (16, 16, 16), (112, 48, 48), (48, 48, 48)]
    comp = nz.make_cuda_compressor("float32", nz.compressor_requirements(shapes))
    for i, shape in enumerate(shapes + shapes):
        data = synth.make("smooth" if i % 2 else "hashed", shape, "float32", seed=40 + i)
        expect = oracle.compress(data)
        d_stream = torch.zeros(nz.compressed_length_bound("float32", shape), dtype=torch.int32, device="cuda")
        d_len = torch.zeros(1, dtype=torch.int32, device="cuda")
        comp.compress(to_device(data), shape, d_stream, d_len)
        torch.cuda.synchronize()
        n = int(d_len.cpu().numpy().view(np.uint32)[0])
        assert n == expect.size, (shape, n, expect.size)
        assert np.array_equal(d_stream[:n].cpu().numpy().view(np.uint32), expect), shape
