"""Host-side logic of the multi-GPU path on CPU: slab partitioning, the cross-rank offset exchange
(all_gather + exclusive scan) and the stream gather, with world_size 2 and 3 over ``gloo``.
Per-rank compression is done by the oracle here (test infrastructure); on GPUs the same host logic
wraps the CUDA path (bench.py, tests/test_gpu_parity.py)."""
import os
import socket

import numpy as np
import pytest

from ndzip_b200 import dist as nzd
from ndzip_b200 import synth

CASES = [
    ("float32", (5 * 4096 + 100,)),
    ("float32", (200, 130)),
    ("float64", (67, 40, 50)),
    ("float64", (3 * 4096,)),
    ("float32", (48, 32, 32)),
]


def test_slab_partition_properties():
    for shape in [(4096 * 9 + 5,), (1000, 64), (16 * 7 + 3, 16, 16), (15, 64, 64), (64, 64, 64)]:
        side = nzd.SIDE[len(shape)]
        for world in (1, 2, 3, 8):
            spans = nzd.slab_partition(shape, world)
            assert spans[0][0] == 0 and spans[-1][1] == shape[0]
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            assert all(s[1] % side == 0 for s in spans[:-1])
            assert sum(nzd.cubes_in(nzd.slab_shape(shape, s)) for s in spans) == nzd.cubes_in(shape)
            assert sum(nzd.border_in(nzd.slab_shape(shape, s)) for s in spans) == nzd.border_in(shape)
            counts = [nzd.cubes_in(nzd.slab_shape(shape, s)) for s in spans]
            assert max(counts) - min(counts) <= nzd.cubes_in((side,) + tuple(shape[1:]))


@pytest.mark.parametrize("dtype,shape", CASES)
@pytest.mark.parametrize("world", [1, 2, 3])
def test_stitched_shards_equal_global_stream(oracle, dtype, shape, world):
    data = synth.hashed(shape, dtype, seed=31)
    expect = oracle.compress(data)
    spans = nzd.slab_partition(shape, world)
    local = [oracle.compress(np.ascontiguousarray(data[b:e])) for b, e in spans]
    got = nzd.stitch_global_stream(dtype, shape, local)
    assert got.size == expect.size
    assert np.array_equal(got, expect)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, dtype, shape, result_dir):
    import torch
    import torch.distributed as dist
    from oracle import get_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = get_oracle()
        data = synth.hashed(shape, dtype, seed=31)
        b, e = nzd.slab_partition(shape, world)[rank]
        local = oracle.compress(np.ascontiguousarray(data[b:e]))
        cubes = nzd.cubes_in(nzd.slab_shape(shape, (b, e)))
        hdr = nzd.header_words(dtype, cubes)
        local_words = int(local[:hdr].view(np.uint32)[cubes - 1]) if cubes else 0
        layout = nzd.exchange_layout(dtype, shape, local_words)
        assert layout.local_cubes == cubes and layout.local_cube_words == local_words
        tdt = torch.int32 if np.dtype(dtype).itemsize == 4 else torch.int64
        t_local = torch.from_numpy(local.view(np.int32 if tdt == torch.int32 else np.int64).copy())
        header32 = t_local[:hdr].view(torch.int32)[:cubes].clone()
        header32 += layout.cube_word_base  # the header fix-up (ndzb_add_offset on the GPU)
        out = nzd.gather_global_stream(layout, t_local, header32, root=0)
        if rank == 0:
            np.save(os.path.join(result_dir, "global.npy"), out.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype,shape", CASES[:3])
@pytest.mark.parametrize("world", [2, 3])
def test_gloo_offset_exchange_and_gather(oracle, tmp_path, dtype, shape, world):
    import torch.multiprocessing as mp
    port = _free_port()
    mp.spawn(_worker, args=(world, port, dtype, shape, str(tmp_path)), nprocs=world, join=True)
    got = np.load(os.path.join(str(tmp_path), "global.npy"))
    bits = np.uint32 if dtype == "float32" else np.uint64
    expect = oracle.compress(synth.hashed(shape, dtype, seed=31))
    assert np.array_equal(got.view(bits), expect)


# ---- sharded stream container -------------------------------------------------------------------

@pytest.mark.parametrize("dtype,shape", CASES)
@pytest.mark.parametrize("world", [1, 2, 3])
def test_sharded_container_round_trip_and_global_stream(oracle, dtype, shape, world):
    data = synth.hashed(shape, dtype, seed=32)
    spans = nzd.slab_partition(shape, world)
    local = [oracle.compress(np.ascontiguousarray(data[b:e])) for b, e in spans]
    buf = nzd.pack_sharded(dtype, shape, local)
    hdr = nzd.decode_sharded_header(buf)
    assert hdr.dtype == dtype and hdr.shape == tuple(shape) and len(hdr.segments) == world
    assert hdr.total_bytes == len(buf) and all(s.byte_offset % 16 == 0 for s in hdr.segments)
    for i, (b, e) in enumerate(spans):  # every segment is a self-contained stream of its slab
        assert hdr.segments[i].slab == (b, e)
        back, consumed = oracle.decompress(nzd.sharded_segment(buf, hdr, i), dtype, hdr.slab_shape(i))
        assert consumed == hdr.segments[i].stream_words and back.tobytes() == data[b:e].tobytes()
    assert np.array_equal(nzd.to_global_stream(buf), oracle.compress(data))  # the reference's single stream


def test_sharded_container_rejects_garbage(oracle):
    data = synth.hashed((200, 130), "float32", seed=1)
    buf = bytearray(nzd.pack_sharded("float32", data.shape, [oracle.compress(data)]))
    for bad in (b"", bytes(64), bytes(buf[:20])):
        with pytest.raises(ValueError):
            nzd.decode_sharded_header(bad)
    buf[4] = 9  # version
    with pytest.raises(ValueError):
        nzd.decode_sharded_header(bytes(buf))
    good = nzd.pack_sharded("float32", data.shape, [oracle.compress(data)])
    hdr = nzd.decode_sharded_header(good)
    with pytest.raises(ValueError):
        nzd.sharded_segment(good[:-8], hdr, 0)
    assert [nzd.segments_of_rank(nzd.sharded_header("float32", (4096 * 8,), [1] * 8), r, 3) for r in range(3)] == [[0, 1], [2, 3, 4], [5, 6, 7]]


def test_sharded_container_table_is_validated_by_the_library(oracle):
    # ndzb_container_decode_header / ndzb_container_to_global_stream (csrc/ndzb_container.cu): every way a table can lie
    shape = (4096 * 6,)
    data = synth.hashed(shape, "float32", seed=5)
    spans = nzd.slab_partition(shape, 3)
    local = [oracle.compress(np.ascontiguousarray(data[b:e])) for b, e in spans]
    good = nzd.pack_sharded("float32", shape, local)
    hdr = nzd.decode_sharded_header(good)
    words = np.frombuffer(good, dtype=np.uint32).copy()

    def table(seg, field):  # u32 index of a table field: begin 0, end 1, words 2-3, offset 4-5
        return 8 + 6 * seg + field

    def broken(index, value):
        w = words.copy()
        w[index] = value
        return w.tobytes()

    for bad in (
        broken(table(1, 0), hdr.segments[1].slab[0] + 4096),           # slabs do not tile dimension 0
        broken(table(2, 1), hdr.segments[2].slab[1] - 4096),           # ... or do not reach its end
        broken(table(1, 4), hdr.segments[1].byte_offset + 4),          # misaligned segment
        broken(table(1, 4), hdr.segments[0].byte_offset),              # overlapping segments
        broken(7, 1 << 24),                                            # absurd segment count
        broken(7, 4),                                                  # table longer than it is
        broken(3, 4), broken(2, 2),                                    # dims / dtype out of range
    ):
        with pytest.raises(ValueError):
            nzd.decode_sharded_header(bad)
    with pytest.raises(ValueError):   # a segment that ends beyond the buffer
        nzd.to_global_stream(good[:-16])
    w = words.copy()                  # a slab stream whose own header disagrees with the segment length
    w[hdr.segments[1].byte_offset // 4 + nzd.cubes_in((spans[1][1] - spans[1][0],)) - 1] += 1
    with pytest.raises(ValueError):
        nzd.to_global_stream(w.tobytes())
    # slabs other than the library's own partition cannot be concatenated (they can still be decoded one by one)
    first = 80  # header of two segments: 4 * (8 + 2 * 6) = 80 bytes
    uneven = nzd.ShardedHeader("float32", shape, [nzd.Segment((0, 4096), local[0].size, first),
                                                  nzd.Segment((4096, shape[0]), 0, first + 4 * ((local[0].size + 3) // 4 * 4))])
    blob = bytearray(uneven.total_bytes)
    blob[: len(nzd.encode_sharded_header(uneven))] = nzd.encode_sharded_header(uneven)
    assert len(nzd.decode_sharded_header(bytes(blob)).segments) == 2
    with pytest.raises(ValueError):
        nzd.to_global_stream(bytes(blob))
    # unaligned buffers parse (the table is read with memcpy)
    shifted = np.frombuffer(b"x" + good, dtype=np.uint8)[1:]
    assert nzd.decode_sharded_header(shifted).segments == hdr.segments


def test_sharded_container_files_fail_cleanly(tmp_path, oracle):
    with pytest.raises(OSError):
        nzd.read_sharded(str(tmp_path / "missing.ndzs"))
    junk = tmp_path / "junk.ndzs"
    junk.write_bytes(b"NDZS" + bytes(60))
    with pytest.raises(ValueError):
        nzd.read_sharded(str(junk))
    data = synth.hashed((4096 * 2,), "float32", seed=6)
    path = str(tmp_path / "one.ndzs")
    nzd.write_sharded(path, "float32", data.shape, oracle.compress(data))
    hdr, segs = nzd.read_sharded(path)
    assert len(segs) == 1 and np.array_equal(segs[0][2], oracle.compress(data))
    with open(path, "r+b") as f:   # file cut inside the segment
        f.truncate(hdr.total_bytes - 8)
    with pytest.raises(ValueError):
        nzd.read_sharded(path)


def _container_worker(rank, world, port, dtype, shape, path):
    import torch.distributed as dist
    from oracle import get_oracle
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        oracle = get_oracle()
        data = synth.hashed(shape, dtype, seed=33)
        b, e = nzd.slab_partition(shape, world)[rank]
        nzd.write_sharded(path, dtype, shape, oracle.compress(np.ascontiguousarray(data[b:e])))
        # read back with the SAME world size: exactly the own slab, no communication
        hdr, mine = nzd.read_sharded(path, rank, world)
        assert len(mine) == 1 and mine[0][0] == (b, e)
        back, _ = oracle.decompress(mine[0][2], dtype, mine[0][1])
        assert back.tobytes() == data[b:e].tobytes()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("dtype,shape", CASES[:3])
@pytest.mark.parametrize("world", [2, 3])
def test_gloo_sharded_write_then_read_with_any_world_size(oracle, tmp_path, dtype, shape, world):
    import torch.multiprocessing as mp
    path = str(tmp_path / "grid.ndzs")
    mp.spawn(_container_worker, args=(world, _free_port(), dtype, shape, path), nprocs=world, join=True)
    data = synth.hashed(shape, dtype, seed=33)
    raw = open(path, "rb").read()
    assert np.array_equal(nzd.to_global_stream(raw), oracle.compress(data))
    # a reader with a different world size takes several (or no) segments per rank; together they cover the grid
    for readers in (1, 2, 5):
        rows = 0
        for r in range(readers):
            hdr, segs = nzd.read_sharded(path, r, readers)
            for (b, e), slab_shape, words in segs:
                back, _ = oracle.decompress(words, dtype, slab_shape)
                assert back.tobytes() == data[b:e].tobytes()
                rows += e - b
        assert rows == shape[0]


# ---- the library's slab plan (ndzb_dist_plan, pure geometry: runs without a GPU) against this module's layout arithmetic,
# which the gloo tests above pin against the oracle

@pytest.mark.parametrize("dtype,shape,world", [
    ("float32", (130, 70, 40), 3), ("float64", (4096 * 5 + 7,), 2), ("float64", (200, 333), 4), ("float32", (16, 16, 16), 5),
    ("float64", (1024, 1024, 1024), 8), ("float32", (1 << 31,), 8), ("float32", (7,), 2)])
def test_library_plan_matches_layout_arithmetic(dtype, shape, world):
    from ndzip_b200 import dist as nzd
    words = [1000 + 17 * r for r in range(world)]
    spans = nzd.slab_partition(shape, world)
    for r in range(world):
        L = nzd.plan(dtype, shape, world, r)
        P = nzd.layout_from_counts(dtype, shape, words, r)
        assert (L.slab_begin, L.slab_end) == spans[r]
        assert tuple(L.slab_size[: len(shape)]) == P.local_shape
        assert L.local_cubes == P.local_cubes and L.cube_index_base == P.cube_index_base
        assert L.local_header_words == P.local_header_words and L.local_border_words == P.local_border
        assert L.border_base == P.border_base and L.global_cubes == P.global_cubes
        assert L.global_header_words == P.global_header_words and L.global_border_words == P.total_border


def test_library_plan_rejects_bad_arguments():
    from ndzip_b200 import NdzipB200Error
    from ndzip_b200 import dist as nzd
    with pytest.raises(NdzipB200Error):
        nzd.plan("float32", (64, 64), 0, 0)
    with pytest.raises(NdzipB200Error):
        nzd.plan("float32", (64, 64), 2, 2)
