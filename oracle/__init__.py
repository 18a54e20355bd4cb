"""TEST INFRASTRUCTURE — ctypes loaders for the CPU checkers. NOT part of the product path.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline / ``--impl reference``
legs may import this package (the product package ``ndzip_b200`` never does).

* :class:`Oracle`     — ``oracle/libndzip_oracle.so``, the plain-C restatement (``ndzip_oracle.c``).
* :class:`Reference`  — ``oracle/_ref/libndzip_ref.so``, the UNMODIFIED reference CPU codec
  (serial + OpenMP) compiled from ``/root/reference`` by ``oracle/Makefile``; present wherever it
  was prebuilt (it travels to the GPU box with the repo snapshot, ``/root/reference`` does not).
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libndzip_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libndzip_ref.so")
REFERENCE_TREE = "/root/reference"

_VALUE = {0: np.float32, 1: np.float64}
_BITS = {0: np.uint32, 1: np.uint64}


def dtype_code(dtype) -> int:
    dt = np.dtype(dtype)
    if dt in (np.dtype(np.float32), np.dtype(np.uint32)):
        return 0
    if dt in (np.dtype(np.float64), np.dtype(np.uint64)):
        return 1
    raise TypeError(f"ndzip handles float32/float64 only, got {dt}")


def bits_dtype(code: int):
    return _BITS[code]


def value_dtype(code: int):
    return _VALUE[code]


def _size3(shape: Sequence[int]):
    dims = len(shape)
    if not 1 <= dims <= 3:
        raise ValueError("1 <= dims <= 3")
    return dims, (ctypes.c_uint32 * 3)(*(list(shape) + [0] * (3 - dims)))


def build(target: str = "all", quiet: bool = True) -> None:
    """Run the committed recipe (oracle/Makefile). ``ref`` needs /root/reference and is skipped
    (keeping any prebuilt _ref) where the tree is absent."""
    targets = ["oracle"]
    if target in ("all", "ref") and os.path.isdir(os.path.join(REFERENCE_TREE, "src", "ndzip")):
        targets.append("ref")
    for t in targets:
        subprocess.run(["make", "-C", _HERE, t], check=True,
                       stdout=subprocess.DEVNULL if quiet else None)


class Oracle:
    """Plain-C restatement (oracle/ndzip_oracle.c)."""

    def __init__(self, path: str = ORACLE_SO):
        if not os.path.exists(path):
            build("oracle")
        L = ctypes.CDLL(path)
        vp, u32, u64, ci = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
        L.ndzo_num_hypercubes.restype = u32
        L.ndzo_num_hypercubes.argtypes = [ci, vp]
        L.ndzo_border_element_count.restype = u64
        L.ndzo_border_element_count.argtypes = [ci, vp]
        L.ndzo_compressed_length_bound.restype = u64
        L.ndzo_compressed_length_bound.argtypes = [ci, ci, vp]
        L.ndzo_compress.restype = u32
        L.ndzo_compress.argtypes = [ci, ci, vp, vp, vp]
        L.ndzo_decompress.restype = u32
        L.ndzo_decompress.argtypes = [ci, ci, vp, vp, vp]
        for name in ("ndzo_block_transform", "ndzo_inverse_block_transform"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [ci, ci, vp]
        L.ndzo_transpose_bits.restype = None
        L.ndzo_transpose_bits.argtypes = [ci, vp, vp]
        L.ndzo_zero_bit_encode.restype = u32
        L.ndzo_zero_bit_encode.argtypes = [ci, vp, vp]
        L.ndzo_zero_bit_decode.restype = u32
        L.ndzo_zero_bit_decode.argtypes = [ci, vp, vp]
        L.ndzo_load_cube.restype = None
        L.ndzo_load_cube.argtypes = [ci, ci, vp, vp, u32, vp]
        L.ndzo_border_slices.restype = ci
        L.ndzo_border_slices.argtypes = [ci, vp, u32, vp, ci]
        self.L = L

    # -- sizes
    def num_hypercubes(self, shape) -> int:
        dims, sz = _size3(shape)
        return self.L.ndzo_num_hypercubes(dims, sz)

    def border_element_count(self, shape) -> int:
        dims, sz = _size3(shape)
        return self.L.ndzo_border_element_count(dims, sz)

    def compressed_length_bound(self, dtype, shape) -> int:
        dims, sz = _size3(shape)
        return self.L.ndzo_compressed_length_bound(dtype_code(dtype), dims, sz)

    # -- whole arrays
    def compress(self, data: np.ndarray) -> np.ndarray:
        code = dtype_code(data.dtype)
        data = np.ascontiguousarray(data)
        dims, sz = _size3(data.shape)
        out = np.zeros(max(1, self.compressed_length_bound(data.dtype, data.shape)), dtype=_BITS[code])
        n = self.L.ndzo_compress(code, dims, sz, data.ctypes.data, out.ctypes.data)
        return out[:n].copy()

    def decompress(self, stream: np.ndarray, dtype, shape) -> tuple[np.ndarray, int]:
        code = dtype_code(dtype)
        stream = np.ascontiguousarray(stream, dtype=_BITS[code])
        dims, sz = _size3(shape)
        out = np.zeros(shape, dtype=_VALUE[code])
        n = self.L.ndzo_decompress(code, dims, sz, stream.ctypes.data, out.ctypes.data)
        return out, n

    # -- single-cube primitives (4096 words of bits_type)
    def block_transform(self, cube: np.ndarray, dims: int, inverse: bool = False) -> np.ndarray:
        code = dtype_code(cube.dtype)
        c = np.ascontiguousarray(cube, dtype=_BITS[code]).copy()
        assert c.size == 4096
        (self.L.ndzo_inverse_block_transform if inverse else self.L.ndzo_block_transform)(code, dims, c.ctypes.data)
        return c

    def transpose_bits(self, words: np.ndarray) -> np.ndarray:
        code = dtype_code(words.dtype)
        w = np.ascontiguousarray(words, dtype=_BITS[code])
        assert w.size == (32, 64)[code]
        out = np.zeros_like(w)
        self.L.ndzo_transpose_bits(code, w.ctypes.data, out.ctypes.data)
        return out

    def zero_bit_encode(self, cube: np.ndarray) -> np.ndarray:
        code = dtype_code(cube.dtype)
        c = np.ascontiguousarray(cube, dtype=_BITS[code])
        out = np.zeros(4096 + 4096 // (32, 64)[code], dtype=_BITS[code])
        n = self.L.ndzo_zero_bit_encode(code, c.ctypes.data, out.ctypes.data)
        return out[:n].copy()

    def zero_bit_decode(self, words: np.ndarray) -> tuple[np.ndarray, int]:
        code = dtype_code(words.dtype)
        w = np.ascontiguousarray(words, dtype=_BITS[code])
        out = np.zeros(4096, dtype=_BITS[code])
        n = self.L.ndzo_zero_bit_decode(code, w.ctypes.data, out.ctypes.data)
        return out, n

    def load_cube(self, data: np.ndarray, hc_index: int) -> np.ndarray:
        code = dtype_code(data.dtype)
        data = np.ascontiguousarray(data)
        dims, sz = _size3(data.shape)
        out = np.zeros(4096, dtype=_BITS[code])
        self.L.ndzo_load_cube(code, dims, sz, data.ctypes.data, hc_index, out.ctypes.data)
        return out

    def border_slices(self, shape, side: int) -> list[tuple[int, int]]:
        dims, sz = _size3(shape)
        n = self.L.ndzo_border_slices(dims, sz, side, None, 0)
        buf = np.zeros(max(1, 2 * n), dtype=np.uint64)
        self.L.ndzo_border_slices(dims, sz, side, buf.ctypes.data, n)
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)]


class Reference:
    """The unmodified reference CPU codec (oracle/_ref/libndzip_ref.so)."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REF_SO)

    def __init__(self, path: str = REF_SO):
        if not os.path.exists(path):
            build("ref")
        if not os.path.exists(path):
            raise FileNotFoundError(f"{path} missing and {REFERENCE_TREE} not present to build it")
        L = ctypes.CDLL(path)
        vp, u32, u64, ci, cu = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, ctypes.c_uint
        L.ndzr_offloader_create.restype = vp
        L.ndzr_offloader_create.argtypes = [ci, ci, cu]
        L.ndzr_offloader_destroy.restype = None
        L.ndzr_offloader_destroy.argtypes = [vp]
        L.ndzr_offloader_compress.restype = u32
        L.ndzr_offloader_compress.argtypes = [vp, vp, vp, vp]
        L.ndzr_offloader_decompress.restype = u32
        L.ndzr_offloader_decompress.argtypes = [vp, vp, vp, u32, vp]
        L.ndzr_compressed_length_bound.restype = u32
        L.ndzr_compressed_length_bound.argtypes = [ci, ci, vp]
        L.ndzr_num_hypercubes.restype = u32
        L.ndzr_num_hypercubes.argtypes = [ci, vp]
        for name in ("ndzr_block_transform", "ndzr_inverse_block_transform",
                     "ndzr_block_transform_simd", "ndzr_inverse_block_transform_simd"):
            getattr(L, name).restype = None
            getattr(L, name).argtypes = [ci, ci, vp]
        L.ndzr_transpose_bits.restype = None
        L.ndzr_transpose_bits.argtypes = [ci, vp, vp]
        L.ndzr_zero_bit_encode.restype = u64
        L.ndzr_zero_bit_encode.argtypes = [ci, vp, vp]
        L.ndzr_zero_bit_decode.restype = u64
        L.ndzr_zero_bit_decode.argtypes = [ci, vp, vp]
        L.ndzr_border_slices.restype = ci
        L.ndzr_border_slices.argtypes = [ci, vp, u32, vp, ci]
        L.ndzr_has_openmp.restype = ci
        L.ndzr_physical_concurrency.restype = cu
        self.L = L
        self._offloaders: dict[tuple[int, int, int], int] = {}

    def __del__(self):
        try:
            for h in self._offloaders.values():
                self.L.ndzr_offloader_destroy(h)
        except Exception:
            pass

    def _offloader(self, code: int, dims: int, threads: int):
        key = (code, dims, threads)
        if key not in self._offloaders:
            h = self.L.ndzr_offloader_create(code, dims, threads)
            if not h:
                raise RuntimeError("reference make_cpu_offloader failed")
            self._offloaders[key] = h
        return self._offloaders[key]

    def physical_concurrency(self) -> int:
        return int(self.L.ndzr_physical_concurrency())

    def compressed_length_bound(self, dtype, shape) -> int:
        dims, sz = _size3(shape)
        return self.L.ndzr_compressed_length_bound(dtype_code(dtype), dims, sz)

    def num_hypercubes(self, shape) -> int:
        dims, sz = _size3(shape)
        return self.L.ndzr_num_hypercubes(dims, sz)

    def compress_into(self, data: np.ndarray, out: np.ndarray, threads: int = 1) -> int:
        """``out`` must be zero-filled by the caller (SURVEY.md §0 padding-word caveat) and hold
        compressed_length_bound words. Returns the stream length in words."""
        code = dtype_code(data.dtype)
        dims, sz = _size3(data.shape)
        return self.L.ndzr_offloader_compress(self._offloader(code, dims, threads), sz,
                                              data.ctypes.data, out.ctypes.data)

    def compress(self, data: np.ndarray, threads: int = 1) -> np.ndarray:
        code = dtype_code(data.dtype)
        data = np.ascontiguousarray(data)
        out = np.zeros(max(1, self.compressed_length_bound(data.dtype, data.shape)), dtype=_BITS[code])
        n = self.compress_into(data, out, threads)
        return out[:n].copy()

    def decompress_into(self, stream: np.ndarray, out: np.ndarray, threads: int = 1) -> int:
        code = dtype_code(out.dtype)
        dims, sz = _size3(out.shape)
        return self.L.ndzr_offloader_decompress(self._offloader(code, dims, threads), sz,
                                                stream.ctypes.data, stream.size, out.ctypes.data)

    def decompress(self, stream: np.ndarray, dtype, shape, threads: int = 1) -> tuple[np.ndarray, int]:
        code = dtype_code(dtype)
        stream = np.ascontiguousarray(stream, dtype=_BITS[code])
        out = np.zeros(shape, dtype=_VALUE[code])
        n = self.decompress_into(stream, out, threads)
        return out, n

    def block_transform(self, cube: np.ndarray, dims: int, inverse: bool = False, simd: bool = False) -> np.ndarray:
        code = dtype_code(cube.dtype)
        c = np.ascontiguousarray(cube, dtype=_BITS[code]).copy()
        name = "ndzr_" + ("inverse_" if inverse else "") + "block_transform" + ("_simd" if simd else "")
        getattr(self.L, name)(code, dims, c.ctypes.data)
        return c

    def transpose_bits(self, words: np.ndarray) -> np.ndarray:
        code = dtype_code(words.dtype)
        w = np.ascontiguousarray(words, dtype=_BITS[code])
        out = np.zeros_like(w)
        self.L.ndzr_transpose_bits(code, w.ctypes.data, out.ctypes.data)
        return out

    def zero_bit_encode(self, cube: np.ndarray) -> np.ndarray:
        code = dtype_code(cube.dtype)
        c = np.ascontiguousarray(cube, dtype=_BITS[code])
        out = np.zeros(2 * 4096, dtype=_BITS[code])
        nbytes = self.L.ndzr_zero_bit_encode(code, c.ctypes.data, out.ctypes.data)
        return out[: nbytes // out.itemsize].copy()

    def zero_bit_decode(self, words: np.ndarray) -> tuple[np.ndarray, int]:
        code = dtype_code(words.dtype)
        w = np.ascontiguousarray(words, dtype=_BITS[code])
        out = np.zeros(4096, dtype=_BITS[code])
        nbytes = self.L.ndzr_zero_bit_decode(code, w.ctypes.data, out.ctypes.data)
        return out, nbytes // out.itemsize

    def border_slices(self, shape, side: int) -> list[tuple[int, int]]:
        dims, sz = _size3(shape)
        n = self.L.ndzr_border_slices(dims, sz, side, None, 0)
        buf = np.zeros(max(1, 2 * n), dtype=np.uint32)
        self.L.ndzr_border_slices(dims, sz, side, buf.ctypes.data, n)
        return [(int(buf[2 * i]), int(buf[2 * i + 1])) for i in range(n)]


REFCUDA_SO = os.path.join(_HERE, "_ref", "libndzip_refcuda.so")


class ReferenceCuda:
    """The unmodified reference CUDA codec recompiled for sm_100a (oracle/_ref/libndzip_refcuda.so,
    `make -C oracle refcuda`): device-pointer API of reference include/ndzip/cuda.hh:10-41."""

    @staticmethod
    def available() -> bool:
        return os.path.exists(REFCUDA_SO)

    def __init__(self, dtype, shape, stream: int = 0, path: str = REFCUDA_SO):
        L = ctypes.CDLL(path)
        vp, ci = ctypes.c_void_p, ctypes.c_int
        L.ndzrc_create.restype = vp
        L.ndzrc_create.argtypes = [ci, ci, vp, vp]
        L.ndzrc_destroy.argtypes = [vp]
        L.ndzrc_compress.restype = ci
        L.ndzrc_compress.argtypes = [vp, vp, vp, vp, vp]
        L.ndzrc_decompress.restype = ci
        L.ndzrc_decompress.argtypes = [vp, vp, vp, vp]
        self.L = L
        self.dims, self.size = _size3(shape)
        self.h = L.ndzrc_create(dtype_code(dtype), self.dims, self.size, stream)
        if not self.h:
            raise RuntimeError("reference make_cuda_compressor failed")

    def compress(self, d_in_ptr: int, d_stream_ptr: int, d_len_ptr: int) -> None:
        if self.L.ndzrc_compress(self.h, d_in_ptr, self.size, d_stream_ptr, d_len_ptr) != 0:
            raise RuntimeError("reference cuda compress failed")

    def decompress(self, d_stream_ptr: int, d_out_ptr: int) -> None:
        if self.L.ndzrc_decompress(self.h, d_stream_ptr, d_out_ptr, self.size) != 0:
            raise RuntimeError("reference cuda decompress failed")

    def __del__(self):
        try:
            if self.h:
                self.L.ndzrc_destroy(self.h)
                self.h = None
        except Exception:
            pass


_oracle: Optional[Oracle] = None
_reference: Optional[Reference] = None


def get_oracle() -> Oracle:
    global _oracle
    if _oracle is None:
        _oracle = Oracle()
    return _oracle


def get_reference() -> Optional[Reference]:
    """The compiled reference if available (prebuilt or buildable here), else None."""
    global _reference
    if _reference is None:
        try:
            _reference = Reference()
        except (FileNotFoundError, OSError, subprocess.CalledProcessError):
            return None
    return _reference
