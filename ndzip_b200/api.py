"""Host-side mirror of the reference's CUDA-facing interface, on top of the C ABI.

Names, argument meaning and error behaviour follow the reference (celerity/ndzip @ ff4e6702):

* ``compressor_requirements``                       include/ndzip/ndzip.hh:255-269, src/ndzip/common.cc:8-28
* ``compressed_length_bound(dtype, extent)``         include/ndzip/ndzip.hh:224-225, src/ndzip/common.cc:31-55
* ``make_cuda_compressor`` / ``cuda_compressor``     include/ndzip/cuda.hh:10-22, 36-38  (device pointers, async)
* ``make_cuda_decompressor`` / ``cuda_decompressor`` include/ndzip/cuda.hh:24-41
* ``make_cuda_offloader`` / ``cuda_offloader``       include/ndzip/offload.hh:8-57       (host pointers, sync)

Device memory and streams are torch tensors / torch streams (plumbing only); every compute call goes
through ``libndzip_b200.so``. Stream words are ``bits_type``: uint32 for float32, uint64 for float64
(held in torch int32 / int64 tensors, same bits).
"""
from __future__ import annotations

import ctypes
from typing import Iterable, Optional, Sequence, Union

import numpy as np

from . import _lib
from ._lib import NdzipB200Error

_DTYPE_CODE = {np.dtype(np.float32): 0, np.dtype(np.float64): 1}


def _code(dtype) -> int:
    try:
        import torch
        if isinstance(dtype, torch.dtype):
            dtype = {torch.float32: np.float32, torch.float64: np.float64}[dtype]
    except (ImportError, KeyError):
        pass
    try:
        return _DTYPE_CODE[np.dtype(dtype)]
    except (KeyError, TypeError):
        raise NdzipB200Error(f"ndzip supports float32 and float64, not {dtype}")


def bits_numpy_dtype(dtype):
    return (np.uint32, np.uint64)[_code(dtype)]


def num_hypercubes(extent: Sequence[int]) -> int:
    """reference src/ndzip/common.hh:395-412"""
    dims, sz = _lib.size3(extent)
    return _lib.load().ndzb_num_hypercubes(dims, sz)


def compressed_length_bound(dtype, extent: Sequence[int]) -> int:
    """Upper bound of the stream length in words (reference src/ndzip/common.cc:31-55)."""
    dims, sz = _lib.size3(extent)
    return _lib.load().ndzb_compressed_length_bound(_code(dtype), dims, sz)


class compressor_requirements:
    """Maximum hypercube count over the extents a compressor will see (reference ndzip.hh:255-269)."""

    def __init__(self, extents: Union[None, Sequence[int], Iterable[Sequence[int]]] = None):
        self._dims = -1
        self._max_num_hypercubes = 0
        if extents is None:
            return
        extents = list(extents)
        if extents and not isinstance(extents[0], (list, tuple)):
            extents = [extents]
        for e in extents:
            self.include(e)

    def include(self, extent: Sequence[int]) -> None:
        dims = len(extent)
        if self._dims == -1:
            self._dims = dims
        elif dims != self._dims:
            # reference src/ndzip/common.cc:12-15
            raise NdzipB200Error(f"Cannot add a {dims}-dimensional extent to {self._dims}-dimensional compressor_requirements")
        self._max_num_hypercubes = max(self._max_num_hypercubes, num_hypercubes(extent))

    @property
    def dimensions(self) -> int:
        if self._dims == -1:
            # reference src/ndzip/common.hh:319-322
            raise NdzipB200Error("Cannot construct a compressor with empty requirements")
        return self._dims

    @property
    def max_num_hypercubes(self) -> int:
        return self._max_num_hypercubes


def _stream_handle(stream) -> int:
    if stream is None:
        return 0
    if isinstance(stream, int):
        return stream
    return int(stream.cuda_stream)  # torch.cuda.Stream


def _ptr(t) -> int:
    if t is None:
        return 0
    if isinstance(t, int):
        return t
    if not t.is_cuda:
        raise NdzipB200Error("expected a CUDA tensor (device pointer API, reference include/ndzip/cuda.hh)")
    if not t.is_contiguous():
        raise NdzipB200Error("tensors passed to ndzip_b200 must be contiguous")
    return t.data_ptr()


class _context:
    def __init__(self, dtype, dims: int, max_hypercubes: int, stream):
        if dims not in (1, 2, 3):
            raise NdzipB200Error("Invalid dimensionality")  # reference src/ndzip/common.hh:642
        self._lib = _lib.load()
        self.dtype_code = _code(dtype)
        self.dims = dims
        self._handle = ctypes.c_void_p()
        _lib.check(self._lib.ndzb_ctx_create(ctypes.byref(self._handle), self.dtype_code, dims, max_hypercubes,
                                             _stream_handle(stream)))

    def close(self):
        if getattr(self, "_handle", None):
            self._lib.ndzb_ctx_destroy(self._handle)
            self._handle = None

    __del__ = close

    @property
    def last_launch_count(self) -> int:
        return self._lib.ndzb_last_launch_count(self._handle)


class cuda_compressor(_context):
    """Device-pointer compressor bound to one CUDA stream (reference include/ndzip/cuda.hh:10-22)."""

    def __init__(self, dtype, requirements: compressor_requirements, stream=None):
        super().__init__(dtype, requirements.dimensions, requirements.max_num_hypercubes, stream)

    def compress(self, in_device_data, data_size: Sequence[int], out_device_stream, out_device_stream_length=None) -> None:
        """Asynchronous; ``out_device_stream`` holds compressed_length_bound words,
        ``out_device_stream_length`` (optional, one uint32) receives the stream length in words."""
        dims, sz = _lib.size3(data_size)
        _lib.check(self._lib.ndzb_compress(self._handle, _ptr(in_device_data), dims, sz, _ptr(out_device_stream),
                                           _ptr(out_device_stream_length)))

    # sharded (multi-GPU) entry points, see include/ndzip_b200.h
    def compress_cubes(self, in_device_data_base: int, data_size, hc_begin: int, hc_end: int, out_cubes, out_offsets, out_local_words) -> None:
        dims, sz = _lib.size3(data_size)
        _lib.check(self._lib.ndzb_compress_cubes(self._handle, in_device_data_base, dims, sz, hc_begin, hc_end,
                                                 _ptr(out_cubes), _ptr(out_offsets), _ptr(out_local_words)))

    def add_offset(self, offsets, count: int, base_words) -> None:
        _lib.check(self._lib.ndzb_add_offset(self._handle, _ptr(offsets), count, _ptr(base_words)))

    def fixup_header(self, local_header, global_header, count: int, gathered_lengths, overhead_words, rank: int) -> None:
        """Multi-GPU exchange step after the all-gather of stream lengths (include/ndzip_b200.h)."""
        _lib.check(self._lib.ndzb_fixup_header(self._handle, _ptr(local_header), _ptr(global_header), count,
                                               _ptr(gathered_lengths), _ptr(overhead_words), rank))

    def pack_border(self, in_device_data, data_size, out) -> None:
        dims, sz = _lib.size3(data_size)
        _lib.check(self._lib.ndzb_pack_border(self._handle, _ptr(in_device_data), dims, sz, _ptr(out)))


class cuda_decompressor(_context):
    """reference include/ndzip/cuda.hh:24-34"""

    def __init__(self, dtype, dims: int, stream=None):
        super().__init__(dtype, dims, 0, stream)

    def decompress(self, in_device_stream, out_device_data, data_size: Sequence[int]) -> None:
        dims, sz = _lib.size3(data_size)
        _lib.check(self._lib.ndzb_decompress(self._handle, _ptr(in_device_stream), _ptr(out_device_data), dims, sz))

    def decompress_cubes(self, in_device_stream, out_device_data_base: int, data_size, hc_begin: int, hc_end: int) -> None:
        dims, sz = _lib.size3(data_size)
        _lib.check(self._lib.ndzb_decompress_cubes(self._handle, _ptr(in_device_stream), out_device_data_base, dims, sz,
                                                   hc_begin, hc_end))


class cuda_offloader(_context):
    """Host-pointer, synchronous (reference include/ndzip/offload.hh:8-34, cuda_codec.inl:654-761)."""

    def __init__(self, dtype, dims: int):
        super().__init__(dtype, dims, 0, None)
        self.kernel_duration_ns: Optional[int] = None

    @staticmethod
    def _host_ptr(a) -> int:
        if isinstance(a, np.ndarray):
            if not a.flags["C_CONTIGUOUS"]:
                raise NdzipB200Error("host arrays must be C-contiguous")
            return a.ctypes.data
        return a.data_ptr()  # pinned / CPU torch tensor

    def compress(self, data, data_size: Sequence[int], stream) -> int:
        """Returns the stream length in words; ``stream`` holds compressed_length_bound words."""
        dims, sz = _lib.size3(data_size)
        length = ctypes.c_uint32(0)
        ns = ctypes.c_uint64(0)
        _lib.check(self._lib.ndzb_offload_compress(self._handle, self._host_ptr(data), dims, sz, self._host_ptr(stream),
                                                   ctypes.byref(length), ctypes.byref(ns)))
        self.kernel_duration_ns = ns.value
        return length.value

    def decompress(self, stream, length: int, data, data_size: Sequence[int]) -> int:
        """Returns the number of stream words consumed."""
        dims, sz = _lib.size3(data_size)
        consumed = ctypes.c_uint32(0)
        ns = ctypes.c_uint64(0)
        _lib.check(self._lib.ndzb_offload_decompress(self._handle, self._host_ptr(stream), length, self._host_ptr(data),
                                                     dims, sz, ctypes.byref(consumed), ctypes.byref(ns)))
        self.kernel_duration_ns = ns.value
        return consumed.value


def device_numa_node(device: int) -> int:
    """NUMA node of CUDA device `device` (-1: unknown / not a NUMA box). ndzb_device_numa_node."""
    return int(_lib.load().ndzb_device_numa_node(int(device)))


def bind_host_to_device(device: int) -> int:
    """Restrict this process to the cores of the device's NUMA node and prefer its memory for new pages; call once per
    rank before allocating pinned buffers. Returns the node or -1 (nothing changed). ndzb_bind_host_to_device."""
    return int(_lib.load().ndzb_bind_host_to_device(int(device)))


def make_cuda_compressor(dtype, requirements: Union[compressor_requirements, Sequence[int]], stream=None) -> cuda_compressor:
    """reference include/ndzip/cuda.hh:36-38, src/ndzip/cuda_factory.cu:4-9"""
    if not isinstance(requirements, compressor_requirements):
        requirements = compressor_requirements(requirements)
    return cuda_compressor(dtype, requirements, stream)


def make_cuda_decompressor(dtype, dims: int, stream=None) -> cuda_decompressor:
    """reference include/ndzip/cuda.hh:40-41, src/ndzip/cuda_factory.cu:11-14"""
    return cuda_decompressor(dtype, dims, stream)


def make_cuda_offloader(dtype, dims: int) -> cuda_offloader:
    """reference include/ndzip/offload.hh:56-57, src/ndzip/cuda_factory.cu:16-19"""
    return cuda_offloader(dtype, dims)
