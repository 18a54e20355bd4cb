"""Multi-GPU sharding of the hot path (new work, SURVEY.md §5 / §8e — the reference is single-GPU).

Hypercubes are independent, so the grid is cut into slabs along the slowest dimension, one slab per
rank (one process per GPU). Every rank compresses its slab into a self-contained local ndzip stream.
The only data-path exchange is the cross-rank exclusive scan of the per-rank compressed word counts
(one ``all_gather`` of a single integer per rank over NCCL/NVLink): it turns the local "offset_after"
header entries into the global ones (reference src/ndzip/common.hh:342-358). The final stream gather
(cube segments to one root over NVLink) is optional and timed separately: it is bounded by one GPU's
NVLink ingest, far below the compression rate.

Global stream of the whole grid == [fixed-up headers, rank order][cube segments, rank order]
[border segments, rank order] — bit-identical to the single-GPU / CPU stream, because with slabs of
whole cube rows both the cube order and the ascending-linear-index border order concatenate.

The host-side logic here is backend-agnostic (``gloo`` on CPU in tests, ``nccl`` on GPUs).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Sequence, Tuple

import numpy as np

SIDE = {1: 4096, 2: 64, 3: 16}


def slab_partition(shape: Sequence[int], world_size: int) -> List[Tuple[int, int]]:
    """[begin, end) along dimension 0 for every rank. Interior boundaries are multiples of the cube
    side; cube rows are dealt out as evenly as possible; the last rank also takes the trailing
    partial rows (the border slab)."""
    dims = len(shape)
    side = SIDE[dims]
    cube_rows = shape[0] // side
    bounds = [0]
    for r in range(world_size):
        rows = cube_rows // world_size + (1 if r < cube_rows % world_size else 0)
        bounds.append(bounds[-1] + rows * side)
    bounds[-1] = shape[0]
    return [(bounds[r], bounds[r + 1]) for r in range(world_size)]


def slab_shape(shape: Sequence[int], span: Tuple[int, int]) -> Tuple[int, ...]:
    return (span[1] - span[0],) + tuple(shape[1:])


def cubes_in(shape: Sequence[int]) -> int:
    side = SIDE[len(shape)]
    n = 1
    for s in shape:
        n *= s // side
    return n


def border_in(shape: Sequence[int]) -> int:
    side = SIDE[len(shape)]
    total, inner = 1, 1
    for s in shape:
        total *= s
        inner *= s // side * side
    return total - inner


def header_words(dtype, num_cubes: int) -> int:
    return num_cubes if np.dtype(dtype).itemsize == 4 else (num_cubes + 1) // 2


@dataclass
class ShardLayout:
    """Where one rank's pieces live in its local stream and in the global stream (all in words)."""
    rank: int
    world_size: int
    local_shape: Tuple[int, ...]
    local_cubes: int
    cube_index_base: int      # global hypercube index of the rank's first cube
    local_header_words: int
    local_cube_words: int     # compressed cube words of this rank
    local_border: int
    global_cubes: int
    global_header_words: int
    cube_word_base: int       # exclusive scan of local_cube_words over ranks
    total_cube_words: int
    border_base: int          # exclusive scan of local_border over ranks
    total_border: int
    global_shape: Tuple[int, ...] = ()
    all_cube_words: Tuple[int, ...] = ()   # every rank's compressed cube words (the gathered vector)
    dtype_itemsize: int = 4

    def peer(self, r: int) -> "ShardLayout":
        """Layout of rank r, derived from the same gathered counts (no further communication)."""
        return layout_from_counts(np.float32 if self.dtype_itemsize == 4 else np.float64, self.global_shape,
                                  self.all_cube_words, r)

    @property
    def global_stream_words(self) -> int:
        return self.global_header_words + self.total_cube_words + self.total_border

    @property
    def global_cube_offset(self) -> int:
        return self.global_header_words + self.cube_word_base

    @property
    def global_border_offset(self) -> int:
        return self.global_header_words + self.total_cube_words + self.border_base


def exchange_layout(dtype, global_shape: Sequence[int], local_cube_words, group=None, device=None) -> ShardLayout:
    """The exchange step: all-gather the per-rank compressed word counts and scan them.
    ``local_cube_words`` is an int (host) or a one-element integer tensor (device; stays on device
    until the gathered vector is read back once)."""
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    spans = slab_partition(global_shape, world)
    if isinstance(local_cube_words, torch.Tensor):
        mine = local_cube_words.to(torch.int64).reshape(1)
    else:
        mine = torch.tensor([int(local_cube_words)], dtype=torch.int64, device=device or "cpu")
    if world > 1:
        gathered = torch.empty(world, dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered = mine
    words = [int(w) for w in gathered.cpu().tolist()]
    return layout_from_counts(dtype, global_shape, words, rank)


def layout_from_counts(dtype, global_shape: Sequence[int], cube_words: Sequence[int], rank: int) -> ShardLayout:
    world = len(cube_words)
    spans = slab_partition(global_shape, world)
    shapes = [slab_shape(global_shape, s) for s in spans]
    cubes = [cubes_in(s) for s in shapes]
    borders = [border_in(s) for s in shapes]
    assert sum(cubes) == cubes_in(global_shape), "slabs must tile the cube grid"
    assert sum(borders) == border_in(global_shape), "slab borders must tile the global border"
    return ShardLayout(
        rank=rank, world_size=world, local_shape=shapes[rank], local_cubes=cubes[rank],
        cube_index_base=sum(cubes[:rank]), local_header_words=header_words(dtype, cubes[rank]),
        local_cube_words=int(cube_words[rank]), local_border=borders[rank],
        global_cubes=sum(cubes), global_header_words=header_words(dtype, sum(cubes)),
        cube_word_base=int(sum(cube_words[:rank])), total_cube_words=int(sum(cube_words)),
        border_base=sum(borders[:rank]), total_border=sum(borders),
        global_shape=tuple(int(x) for x in global_shape), all_cube_words=tuple(int(w) for w in cube_words),
        dtype_itemsize=np.dtype(dtype).itemsize)


def stitch_global_stream(dtype, global_shape: Sequence[int], local_streams: Sequence[np.ndarray]) -> np.ndarray:
    """Host-side assembly of the global stream from every rank's complete local stream (numpy, bits
    dtype). Used for validation and by the CPU tests; on GPUs the same arithmetic drives the NVLink
    gather (``gather_global_stream``)."""
    bits = np.uint32 if np.dtype(dtype).itemsize == 4 else np.uint64
    world = len(local_streams)
    spans = slab_partition(global_shape, world)
    shapes = [slab_shape(global_shape, s) for s in spans]
    cubes = [cubes_in(s) for s in shapes]
    borders = [border_in(s) for s in shapes]
    hdrs = [header_words(dtype, c) for c in cubes]
    cube_words = []
    headers32 = []
    for r, s in enumerate(local_streams):
        s = np.ascontiguousarray(s, dtype=bits)
        h32 = s[: hdrs[r]].view(np.uint32)[: cubes[r]]
        cube_words.append(int(h32[-1]) if cubes[r] else 0)
        headers32.append(h32)
    layouts = [layout_from_counts(dtype, global_shape, cube_words, r) for r in range(world)]
    out = np.zeros(layouts[0].global_stream_words, dtype=bits)
    gh32 = out[: layouts[0].global_header_words].view(np.uint32)
    for r, s in enumerate(local_streams):
        L = layouts[r]
        s = np.ascontiguousarray(s, dtype=bits)
        gh32[L.cube_index_base: L.cube_index_base + L.local_cubes] = headers32[r] + np.uint32(L.cube_word_base)
        out[L.global_cube_offset: L.global_cube_offset + L.local_cube_words] = s[hdrs[r]: hdrs[r] + L.local_cube_words]
        b0 = hdrs[r] + L.local_cube_words
        out[L.global_border_offset: L.global_border_offset + L.local_border] = s[b0: b0 + L.local_border]
    return out


def gather_global_stream(layout: ShardLayout, local_stream, fixed_header32, root: int = 0, group=None):
    """Final stream gather over the process group (NCCL send/recv on NVLink; gloo in the CPU tests).
    ``local_stream``: the rank's complete local stream tensor (bits held as int32 / int64);
    ``fixed_header32``: int32 tensor with the rank's ``local_cubes`` GLOBAL header entries.
    Returns the assembled global stream tensor on ``root`` (None elsewhere)."""
    import torch
    import torch.distributed as dist

    world, rank = layout.world_size, layout.rank
    hdr = layout.local_header_words
    cubes_seg = local_stream[hdr: hdr + layout.local_cube_words]
    border_seg = local_stream[hdr + layout.local_cube_words: hdr + layout.local_cube_words + layout.local_border]
    out = None
    ops = []
    if rank == root:
        out = torch.zeros(layout.global_stream_words, dtype=local_stream.dtype, device=local_stream.device)
        header32 = out[: layout.global_header_words].view(torch.int32)
        for r in range(world):
            L = layout.peer(r)
            h_dst = header32[L.cube_index_base: L.cube_index_base + L.local_cubes]
            c_dst = out[L.global_cube_offset: L.global_cube_offset + L.local_cube_words]
            b_dst = out[L.global_border_offset: L.global_border_offset + L.local_border]
            if r == root:
                h_dst.copy_(fixed_header32[: L.local_cubes])
                c_dst.copy_(cubes_seg)
                b_dst.copy_(border_seg)
            else:
                if L.local_cubes:
                    ops.append(dist.P2POp(dist.irecv, h_dst, r, group))
                if L.local_cube_words:
                    ops.append(dist.P2POp(dist.irecv, c_dst, r, group))
                if L.local_border:
                    ops.append(dist.P2POp(dist.irecv, b_dst, r, group))
    else:
        if layout.local_cubes:
            ops.append(dist.P2POp(dist.isend, fixed_header32[: layout.local_cubes].contiguous(), root, group))
        if layout.local_cube_words:
            ops.append(dist.P2POp(dist.isend, cubes_seg.contiguous(), root, group))
        if layout.local_border:
            ops.append(dist.P2POp(dist.isend, border_seg.contiguous(), root, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return out


# --------------------------------------------------------------------------------------------------
# Sharded stream container (new work, nothing to replace in the reference)
#
# An 8-GPU pipeline that compresses to storage and decompresses again on 8 GPUs never needs the global
# stream: neither the cross-rank offset exchange nor the gather to one root. The container keeps every
# rank's SELF-CONTAINED local ndzip stream (the stream of its slab, which the reference decoder can read
# with `-n <slab shape>`) behind a small segment table:
#
#   u32 magic "NDZS" | u32 version | u32 dtype (0 f32, 1 f64) | u32 dims | u32 shape[3] | u32 segments
#   per segment: u32 slab_begin | u32 slab_end (dimension 0) | u64 stream_words | u64 byte_offset
#   segments, each starting at a multiple of 16 bytes
#
# The format, its validation, the conversion to the reference's single stream and the file I/O live in the
# library (csrc/ndzb_container.cu, include/ndzip_b200.h ndzb_container_*); what follows is the ctypes mirror.
# Writing needs one all-gather of the stream lengths (for the byte offsets); reading needs nothing: a
# rank takes the segments whose slabs it owns, in any world size.

SHARDED_MAGIC = 0x535A444E  # "NDZS" little endian
SHARDED_VERSION = 1
_FIXED_WORDS = 8            # magic, version, dtype, dims, shape[3], segments
_SEGMENT_WORDS = 6          # begin, end, words (2), offset (2)


@dataclass
class Segment:
    slab: Tuple[int, int]      # [begin, end) along dimension 0 of the global grid
    stream_words: int          # length of the slab's ndzip stream in bits_type words
    byte_offset: int           # where it starts in the container


@dataclass
class ShardedHeader:
    dtype: str
    shape: Tuple[int, ...]
    segments: List[Segment]

    @property
    def header_bytes(self) -> int:
        from . import _lib
        return int(_lib.load().ndzb_container_header_bytes(len(self.segments)))

    @property
    def total_bytes(self) -> int:
        if not self.segments:
            return self.header_bytes
        last = self.segments[-1]
        return last.byte_offset + last.stream_words * np.dtype(self.dtype).itemsize

    def slab_shape(self, i: int) -> Tuple[int, ...]:
        return slab_shape(self.shape, self.segments[i].slab)


def _container_error(status: int, what: str):
    """ValueError for anything that is not a well-formed container, like the Python prototype raised."""
    from . import _lib
    msg = _lib.load().ndzb_strerror(status).decode()
    return ValueError(f"{what}: {msg}") if status in (-6, -1, -3) else OSError(f"{what}: {msg}")


def _c_structs(hdr: ShardedHeader):
    import ctypes
    from . import _lib
    info = _lib.ContainerInfo()
    info.dtype = 0 if hdr.dtype == "float32" else 1
    info.dims = len(hdr.shape)
    for d, n in enumerate(hdr.shape):
        info.size[d] = int(n)
    info.segments = len(hdr.segments)
    info.header_bytes = hdr.header_bytes
    info.total_bytes = hdr.total_bytes
    segs = (_lib.ContainerSegment * max(1, len(hdr.segments)))()
    for i, s in enumerate(hdr.segments):
        segs[i].slab_begin, segs[i].slab_end = s.slab
        segs[i].stream_words = s.stream_words
        segs[i].byte_offset = s.byte_offset
    return info, segs, ctypes


def _from_c(info, segs) -> ShardedHeader:
    return ShardedHeader("float32" if info.dtype == 0 else "float64", tuple(int(info.size[d]) for d in range(info.dims)),
                         [Segment((int(s.slab_begin), int(s.slab_end)), int(s.stream_words), int(s.byte_offset))
                          for s in segs[: info.segments]])


def sharded_header(dtype, global_shape: Sequence[int], stream_words: Sequence[int]) -> ShardedHeader:
    """Segment table for `len(stream_words)` ranks that own the slabs of slab_partition(). ndzb_container_plan."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    n = len(stream_words)
    dims, sz = _lib.size3(tuple(int(x) for x in global_shape))
    words = (ctypes.c_uint64 * max(1, n))(*[int(w) for w in stream_words])
    info = _lib.ContainerInfo()
    segs = (_lib.ContainerSegment * max(1, n))()
    rc = lib.ndzb_container_plan(0 if np.dtype(dtype).itemsize == 4 else 1, dims, sz, n, words, ctypes.byref(info), segs)
    if rc:
        raise _container_error(rc, "ndzb_container_plan")
    return _from_c(info, segs)


def encode_sharded_header(hdr: ShardedHeader) -> bytes:
    """ndzb_container_encode_header"""
    from . import _lib
    info, segs, ctypes = _c_structs(hdr)
    out = ctypes.create_string_buffer(hdr.header_bytes)
    rc = _lib.load().ndzb_container_encode_header(ctypes.byref(info), segs, out, hdr.header_bytes)
    if rc:
        raise _container_error(rc, "ndzb_container_encode_header")
    return out.raw


def _as_u8(buf) -> np.ndarray:
    return np.ascontiguousarray(np.frombuffer(buf, dtype=np.uint8))


def decode_sharded_header(buf) -> ShardedHeader:
    """Parses the table at the start of a container (bytes / memoryview / uint8 array); raises ValueError on
    anything that is not one. ndzb_container_decode_header."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    raw = _as_u8(buf)
    info = _lib.ContainerInfo()
    rc = lib.ndzb_container_decode_header(raw.ctypes.data, raw.size, ctypes.byref(info), None, 0)
    if rc:
        raise _container_error(rc, "not an ndzip sharded stream")
    segs = (_lib.ContainerSegment * info.segments)()
    rc = lib.ndzb_container_decode_header(raw.ctypes.data, raw.size, ctypes.byref(info), segs, info.segments)
    if rc:
        raise _container_error(rc, "not an ndzip sharded stream")
    return _from_c(info, segs)


def pack_sharded(dtype, global_shape: Sequence[int], local_streams: Sequence[np.ndarray]) -> bytes:
    """One process has every rank's local stream (numpy, bits dtype): the whole container."""
    bits = np.uint32 if np.dtype(dtype).itemsize == 4 else np.uint64
    hdr = sharded_header(dtype, global_shape, [np.asarray(s).size for s in local_streams])
    out = bytearray(hdr.total_bytes)
    out[: hdr.header_bytes] = encode_sharded_header(hdr)
    for seg, s in zip(hdr.segments, local_streams):
        raw = np.ascontiguousarray(s, dtype=bits).tobytes()
        out[seg.byte_offset: seg.byte_offset + len(raw)] = raw
    return bytes(out)


def sharded_segment(buf, hdr: ShardedHeader, i: int) -> np.ndarray:
    """Segment i of a container as an array of bits_type words (a view into `buf`)."""
    bits = np.uint32 if hdr.dtype == "float32" else np.uint64
    seg = hdr.segments[i]
    raw = np.frombuffer(buf, dtype=np.uint8)
    end = seg.byte_offset + seg.stream_words * np.dtype(bits).itemsize
    if end > raw.size:
        raise ValueError("truncated sharded stream")
    return raw[seg.byte_offset: end].view(bits)


def segments_of_rank(hdr: ShardedHeader, rank: int, world_size: int) -> List[int]:
    """Which segments a rank of a (possibly different) world size decodes: contiguous, as even as possible."""
    n = len(hdr.segments)
    lo = rank * n // world_size
    hi = (rank + 1) * n // world_size
    return list(range(lo, hi))


def to_global_stream(buf) -> np.ndarray:
    """The reference's single stream of the whole grid, from a container whose slabs are those of
    slab_partition(shape, segments) (what write_sharded / pack_sharded produce). ndzb_container_to_global_stream."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    raw = _as_u8(buf)
    hdr = decode_sharded_header(raw)
    bits = np.uint32 if hdr.dtype == "float32" else np.uint64
    words = ctypes.c_uint64(0)
    rc = lib.ndzb_container_to_global_stream(raw.ctypes.data, raw.size, None, 0, ctypes.byref(words))
    if rc == -1:
        raise ValueError("segments are not the slabs of slab_partition(); decode them one by one instead")
    if rc:
        raise _container_error(rc, "ndzb_container_to_global_stream")
    out = np.empty(words.value, dtype=bits)
    rc = lib.ndzb_container_to_global_stream(raw.ctypes.data, raw.size, out.ctypes.data, out.size, ctypes.byref(words))
    if rc:
        raise _container_error(rc, "ndzb_container_to_global_stream")
    return out


def decompress_segment(offloader, buf, index: int, out_slab) -> None:
    """Sharded decompression: segment `index` of a container in host memory -> its slab (numpy array or pinned torch
    tensor of the slab's shape) on the GPU behind `offloader` (make_cuda_offloader). ndzb_container_decompress_segment."""
    from . import _lib
    raw = _as_u8(buf)
    ptr = out_slab.ctypes.data if isinstance(out_slab, np.ndarray) else out_slab.data_ptr()
    _lib.check(_lib.load().ndzb_container_decompress_segment(offloader._handle, raw.ctypes.data, raw.size, int(index), ptr, None))


def write_sharded(path: str, dtype, global_shape: Sequence[int], local_stream, group=None) -> ShardedHeader:
    """Collective: every rank writes its own segment of the container at `path` (a file all ranks can reach) with
    pwrite() — ndzb_container_create_file on rank 0, ndzb_container_write_segment everywhere. `local_stream`: the rank's
    complete local stream (numpy bits array, or a torch tensor on any device — it is brought to the host here). The
    only communication is one all-gather of the stream lengths."""
    import torch
    import torch.distributed as dist
    from . import _lib

    lib = _lib.load()
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if isinstance(local_stream, torch.Tensor):
        host = local_stream.detach().cpu().numpy()
        device = local_stream.device if local_stream.is_cuda else "cpu"
    else:
        host = np.ascontiguousarray(local_stream)
        device = "cpu"
    if world > 1 and device == "cpu" and dist.get_backend(group) == "nccl":
        device = torch.device("cuda", torch.cuda.current_device())  # NCCL moves device tensors only
    mine = torch.tensor([int(host.size)], dtype=torch.int64, device=device)
    if world > 1:
        gathered = torch.empty(world, dtype=torch.int64, device=mine.device)
        dist.all_gather_into_tensor(gathered, mine, group=group)
    else:
        gathered = mine
    hdr = sharded_header(dtype, global_shape, [int(w) for w in gathered.cpu().tolist()])
    info, segs, ctypes = _c_structs(hdr)

    def agree(rc: int) -> int:
        """Every rank learns the worst status (instead of a bare barrier: a rank that failed must not leave the others
        waiting in the next collective)."""
        if world == 1:
            return rc
        t = torch.tensor([rc], dtype=torch.int64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
        return int(t.item())

    rc = lib.ndzb_container_create_file(path.encode(), ctypes.byref(info), segs) if rank == 0 else 0
    rc = agree(rc)
    if rc:
        raise _container_error(rc, f"creating {path}")
    rc = agree(lib.ndzb_container_write_segment(path.encode(), info.dtype, ctypes.byref(segs[rank]), host.ctypes.data))
    if rc:
        raise _container_error(rc, f"writing the segments of {path}")
    return hdr


def read_sharded(path: str, rank: int = 0, world_size: int = 1):
    """The segments a rank owns: [(slab span, slab shape, stream words as numpy bits array), ...]. No communication.
    ndzb_container_read_header / ndzb_container_read_segment."""
    import ctypes
    from . import _lib
    lib = _lib.load()
    info = _lib.ContainerInfo()
    rc = lib.ndzb_container_read_header(path.encode(), ctypes.byref(info), None, 0)
    if rc:
        raise _container_error(rc, f"reading {path}")
    segs = (_lib.ContainerSegment * info.segments)()
    rc = lib.ndzb_container_read_header(path.encode(), ctypes.byref(info), segs, info.segments)
    if rc:
        raise _container_error(rc, f"reading {path}")
    hdr = _from_c(info, segs)
    bits = np.uint32 if hdr.dtype == "float32" else np.uint64
    out = []
    for i in segments_of_rank(hdr, rank, world_size):
        words = np.empty(hdr.segments[i].stream_words, dtype=bits)
        rc = lib.ndzb_container_read_segment(path.encode(), info.dtype, ctypes.byref(segs[i]), words.ctypes.data)
        if rc:
            raise ValueError("truncated sharded stream") if rc == -6 else _container_error(rc, f"reading segment {i} of {path}")
        out.append((hdr.segments[i].slab, hdr.slab_shape(i), words))
    return hdr, out


# --------------------------------------------------------------------------------------------------
# The data plane proper lives in the library (csrc/ndzb_dist.cu, include/ndzip_b200.h ndzb_dist_*): slab
# compression, the NCCL count exchange on a side stream, the header fix-up kernel and the NCCL gather. This
# class is its ctypes mirror; torch.distributed is only used to ship the 128-byte NCCL unique id.

def plan(dtype, global_shape: Sequence[int], world_size: int, rank: int):
    """ndzb_dist_plan: a rank's slab and where its pieces go in the global stream (pure geometry, no GPU)."""
    import ctypes
    from . import _lib
    from .api import _code
    dims, sz = _lib.size3(global_shape)
    out = _lib.DistLayout()
    _lib.check(_lib.load().ndzb_dist_plan(_code(dtype), dims, sz, world_size, rank, ctypes.byref(out)))
    return out


class DistCodec:
    """One rank of the multi-GPU hot path. Collective constructor (NCCL communicator over the ranks of
    ``group``, or of the default process group)."""

    def __init__(self, dtype, global_shape: Sequence[int], group=None, stream=None):
        import ctypes
        import torch
        import torch.distributed as dist
        from . import _lib
        from .api import _code, _stream_handle

        self._lib = _lib.load()
        self.dtype = np.dtype(dtype)
        self.global_shape = tuple(int(x) for x in global_shape)
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        dims, sz = _lib.size3(self.global_shape)
        uid = torch.zeros(128, dtype=torch.uint8)
        if self.world > 1:
            if self.rank == 0:
                _lib.check(self._lib.ndzb_dist_unique_id(uid.data_ptr()))
            backend = dist.get_backend(group)
            carrier = uid.cuda() if backend == "nccl" else uid
            dist.broadcast(carrier, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
            uid = carrier.cpu()
        self._handle = ctypes.c_void_p()
        _lib.check(self._lib.ndzb_dist_create(ctypes.byref(self._handle), _code(dtype), dims, sz, uid.data_ptr(), self.rank,
                                               self.world, _stream_handle(stream)))
        self.layout = _lib.DistLayout()
        _lib.check(self._lib.ndzb_dist_layout_of(self._handle, self.rank, ctypes.byref(self.layout)))
        self.slab_shape = tuple(int(self.layout.slab_size[d]) for d in range(dims))

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.ndzb_dist_destroy(self._handle)
            self._handle = None

    __del__ = close

    def compress(self, d_slab, d_local_stream, d_local_length=None) -> None:
        from . import _lib
        from .api import _ptr
        _lib.check(self._lib.ndzb_dist_compress(self._handle, _ptr(d_slab), _ptr(d_local_stream), _ptr(d_local_length)))

    def decompress(self, d_local_stream, d_slab) -> None:
        from . import _lib
        from .api import _ptr
        _lib.check(self._lib.ndzb_dist_decompress(self._handle, _ptr(d_local_stream), _ptr(d_slab)))

    def wait_exchange(self) -> None:
        from . import _lib
        _lib.check(self._lib.ndzb_dist_wait_exchange(self._handle))

    def gather(self, d_local_stream, d_global_stream=None, root: int = 0) -> int:
        """Collective. Returns the global stream length in words (known on every rank)."""
        import ctypes
        from . import _lib
        from .api import _ptr
        total = ctypes.c_uint64(0)
        _lib.check(self._lib.ndzb_dist_gather(self._handle, _ptr(d_local_stream), _ptr(d_global_stream), root, ctypes.byref(total)))
        return total.value

    @property
    def last_gather_path(self) -> str:
        """How the last gather moved the data (ndzb_dist_last_gather_path)."""
        return {0: "nccl send/recv", 1: "peer stores over NVLink (CUDA IPC mapping)", 2: "peer stores over NVLink (same process)"}[
            self._lib.ndzb_dist_last_gather_path(self._handle)]

    def global_header(self):
        """This rank's slice of the global header as a torch int32 tensor (valid after wait_exchange on the stream)."""
        import torch
        n = int(self.layout.local_cubes)
        ptr = self._lib.ndzb_dist_global_header(self._handle)
        return _device_view(ptr, n, torch.int32)

    def gathered_lengths(self):
        import torch
        return _device_view(self._lib.ndzb_dist_gathered_lengths(self._handle), self.world, torch.int32)


def _device_view(ptr: int, n: int, tdtype):
    """A torch tensor over device memory owned by the library (no copy)."""
    import torch

    class _Holder:
        pass

    h = _Holder()
    item = torch.tensor([], dtype=tdtype).element_size()
    typestr = {4: "<i4", 8: "<i8"}[item]
    h.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (int(ptr) if n else 0, False), "version": 2}
    return torch.as_tensor(h, device="cuda") if n else torch.empty(0, dtype=tdtype, device="cuda")
