"""ctypes binding of the C ABI in include/ndzip_b200.h. Fails loudly if the CUDA library is missing:
there is no CPU fallback (BASELINE.json north_star)."""
from __future__ import annotations

import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# NDZB_LIB selects another build of the same library (A/B runs of experimental kernels, scripts/build_variant.py)
LIB_PATH = os.environ.get("NDZB_LIB") or os.path.join(HERE, "libndzip_b200.so")

F32, F64 = 0, 1

# every symbol include/ndzip_b200.h declares: (name, restype, argtypes)
_vp, _u32, _u64, _i = ctypes.c_void_p, ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int
_pu32, _pu64 = ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint64)
SYMBOLS = [
    ("ndzb_ctx_create", _i, [ctypes.POINTER(_vp), _i, _i, _u32, _vp]),
    ("ndzb_ctx_destroy", None, [_vp]),
    ("ndzb_compress", _i, [_vp, _vp, _i, _vp, _vp, _vp]),
    ("ndzb_decompress", _i, [_vp, _vp, _vp, _i, _vp]),
    ("ndzb_offload_compress", _i, [_vp, _vp, _i, _vp, _vp, _pu32, _pu64]),
    ("ndzb_offload_decompress", _i, [_vp, _vp, _u32, _vp, _i, _vp, _pu32, _pu64]),
    ("ndzb_offload_chunk_plan", _i, [_i, _i, _vp, _i, _vp, _u32, _pu32]),
    ("ndzb_host_alloc", _i, [ctypes.POINTER(_vp), ctypes.c_size_t]),
    ("ndzb_host_free", None, [_vp]),
    ("ndzb_device_numa_node", _i, [_i]),
    ("ndzb_bind_host_to_device", _i, [_i]),
    ("ndzb_compress_cubes", _i, [_vp, _vp, _i, _vp, _u32, _u32, _vp, _vp, _vp]),
    ("ndzb_add_offset", _i, [_vp, _vp, _u32, _vp]),
    ("ndzb_fixup_header", _i, [_vp, _vp, _vp, _u32, _vp, _vp, _u32]),
    ("ndzb_pack_border", _i, [_vp, _vp, _i, _vp, _vp]),
    ("ndzb_decompress_cubes", _i, [_vp, _vp, _vp, _i, _vp, _u32, _u32]),
    ("ndzb_num_hypercubes", _u32, [_i, _vp]),
    ("ndzb_compressed_length_bound", _u64, [_i, _i, _vp]),
    ("ndzb_border_element_count", _u64, [_i, _vp]),
    ("ndzb_header_words", _u32, [_i, _u32]),
    ("ndzb_compressed_cube_bound", _u32, [_i]),
    ("ndzb_strerror", ctypes.c_char_p, [_i]),
    ("ndzb_last_cuda_error", ctypes.c_char_p, []),
    ("ndzb_version", ctypes.c_char_p, []),
    ("ndzb_last_launch_count", _u32, [_vp]),
    ("ndzb_fixup_header_on", _i, [_vp, _vp, _vp, _u32, _vp, _vp, _u32]),
    ("ndzb_selftest_lookback", _i, [_vp, _i, _vp, _u32, _u32, _vp]),
    ("ndzb_selftest_warp_scan", _i, [_vp, _vp, _vp, _u32]),
    # multi-GPU data plane
    ("ndzb_dist_plan", _i, [_i, _i, _vp, _i, _i, _vp]),
    ("ndzb_dist_unique_id", _i, [_vp]),
    ("ndzb_dist_create", _i, [ctypes.POINTER(_vp), _i, _i, _vp, _vp, _i, _i, _vp]),
    ("ndzb_dist_create_local", _i, [ctypes.POINTER(_vp), _i, _i, _vp, _i, _vp]),
    ("ndzb_dist_destroy", None, [_vp]),
    ("ndzb_dist_layout_of", _i, [_vp, _i, _vp]),
    ("ndzb_dist_stream", _vp, [_vp]),
    ("ndzb_dist_compress", _i, [_vp, _vp, _vp, _vp]),
    ("ndzb_dist_decompress", _i, [_vp, _vp, _vp]),
    ("ndzb_dist_wait_exchange", _i, [_vp]),
    ("ndzb_dist_global_header", _vp, [_vp]),
    ("ndzb_dist_gathered_lengths", _vp, [_vp]),
    ("ndzb_dist_gather", _i, [_vp, _vp, _vp, _i, _pu64]),
    ("ndzb_dist_last_error", ctypes.c_char_p, []),
    ("ndzb_dist_last_gather_path", _i, [_vp]),
    # sharded stream container
    ("ndzb_container_header_bytes", _u64, [_u32]),
    ("ndzb_container_plan", _i, [_i, _i, _vp, _u32, _vp, _vp, _vp]),
    ("ndzb_container_encode_header", _i, [_vp, _vp, _vp, _u64]),
    ("ndzb_container_decode_header", _i, [_vp, _u64, _vp, _vp, _u32]),
    ("ndzb_container_to_global_stream", _i, [_vp, _u64, _vp, _u64, _pu64]),
    ("ndzb_container_create_file", _i, [ctypes.c_char_p, _vp, _vp]),
    ("ndzb_container_write_segment", _i, [ctypes.c_char_p, _i, _vp, _vp]),
    ("ndzb_container_read_header", _i, [ctypes.c_char_p, _vp, _vp, _u32]),
    ("ndzb_container_read_segment", _i, [ctypes.c_char_p, _i, _vp, _vp]),
    ("ndzb_container_decompress_segment", _i, [_vp, _vp, _u64, _u32, _vp, _pu64]),
]


class ContainerSegment(ctypes.Structure):
    """struct ndzb_container_segment (include/ndzip_b200.h)"""
    _fields_ = [("slab_begin", _u32), ("slab_end", _u32), ("stream_words", _u64), ("byte_offset", _u64)]


class ContainerInfo(ctypes.Structure):
    """struct ndzb_container_info (include/ndzip_b200.h)"""
    _fields_ = [("dtype", ctypes.c_int32), ("dims", ctypes.c_int32), ("size", _u32 * 3), ("segments", _u32),
                ("header_bytes", _u64), ("total_bytes", _u64)]


class DistLayout(ctypes.Structure):
    """struct ndzb_dist_layout (include/ndzip_b200.h)"""
    _fields_ = [
        ("slab_begin", _u32), ("slab_end", _u32), ("slab_size", _u32 * 3), ("local_cubes", _u32),
        ("cube_index_base", _u32), ("local_header_words", _u32), ("local_border_words", _u64), ("border_base", _u64),
        ("local_bound_words", _u64), ("global_cubes", _u32), ("global_header_words", _u32),
        ("global_border_words", _u64), ("global_bound_words", _u64),
    ]

_lib = None


class NdzipB200Error(RuntimeError):
    """Mirrors the reference's std::runtime_error (src/ndzip/cuda_bits.cuh:165-169, cuda_codec.inl:557-559)."""


def load() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NdzipB200Error(
                f"{LIB_PATH} is missing: build it with `python -m ndzip_b200.build` "
                "(or __graft_entry__.build()). ndzip_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, res, args in SYMBOLS:
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(status: int) -> None:
    if status != 0:
        lib = load()
        msg = lib.ndzb_strerror(status).decode()
        if status == -4:
            detail = lib.ndzb_last_cuda_error().decode()
            dist_detail = lib.ndzb_dist_last_error().decode()
            msg += ": " + (dist_detail if dist_detail != "no error" and detail == "no error" else detail)
        raise NdzipB200Error(msg)


def size3(shape):
    dims = len(shape)
    if not 1 <= dims <= 3:
        raise NdzipB200Error("Invalid dimensionality")  # reference src/ndzip/common.hh:642
    return dims, (ctypes.c_uint32 * 3)(*(list(int(s) for s in shape) + [0] * (3 - dims)))
