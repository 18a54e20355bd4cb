"""Helpers for the -m gpu parity tests: run the CUDA path through the C ABI (ndzip_b200.api)."""
import os
from contextlib import contextmanager

import numpy as np

import ndzip_b200 as nz


@contextmanager
def load_path(name):
    """Force the compress input path AND the decompress output path (tma | vec16 | scalar); read by ndzb_ctx_create
    (NDZB_LOAD_PATH: TMA tensor load / 16-byte loads / element-wise loads; NDZB_STORE_PATH: TMA tensor store of the
    decoded tile / 8-16-byte stores / element-wise stores)."""
    keys = ("NDZB_LOAD_PATH", "NDZB_STORE_PATH")
    old = {k: os.environ.get(k) for k in keys}
    for k in keys:
        if name is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = name
    try:
        yield
    finally:
        for k in keys:
            if old[k] is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = old[k]


@contextmanager
def compress_kernel(kind=None, variant=None):
    """Select the compress kernel for TMA-compatible inputs (read by ndzb_ctx_create):
    kind "v1" = compress_kernel, anything else = compress_ws_kernel in tuning variant `variant`.
    "v1" and variants > 0 only exist in -DNDZB_TUNING builds of the library (NDZB_EXTRA_NVCC_FLAGS=-DNDZB_TUNING);
    with the default build those parametrisations are skipped."""
    if kind == "v1" or (variant or 0) > 0:
        import pytest
        if not tuning_build():
            pytest.skip("needs a -DNDZB_TUNING build of libndzip_b200.so")
    saved = {k: os.environ.get(k) for k in ("NDZB_COMPRESS_KERNEL", "NDZB_WS_VARIANT")}
    for k, v in (("NDZB_COMPRESS_KERNEL", kind), ("NDZB_WS_VARIANT", None if variant is None else str(variant))):
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    try:
        yield
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v


def tuning_build():
    from ndzip_b200 import _lib
    return b"tuning" in _lib.load().ndzb_version()


def _torch_bits(dtype):
    import torch
    return torch.int32 if np.dtype(dtype) == np.float32 else torch.int64


def to_device(data: np.ndarray):
    import torch
    if data.size == 0:
        return torch.empty(data.shape, dtype=torch.float32 if data.dtype == np.float32 else torch.float64, device="cuda")
    # go through the integer view so that NaN payloads survive untouched
    bits = np.ascontiguousarray(data).view(np.int32 if data.dtype == np.float32 else np.int64)
    t = torch.from_numpy(bits.copy()).cuda()
    return t.view(torch.float32 if data.dtype == np.float32 else torch.float64)


def gpu_compress(data: np.ndarray, compressor=None, want_length_tensor=True, fill=0x5A):
    """Returns (stream words as numpy bits array, full device buffer as numpy) — the device stream
    buffer is pre-filled with a junk byte so that unwritten words are detected."""
    import torch
    dtype, shape = data.dtype, data.shape
    bound = nz.compressed_length_bound(dtype, shape)
    comp = compressor or nz.make_cuda_compressor(dtype, nz.compressor_requirements(shape))
    d_in = to_device(data)
    d_stream = torch.full((max(bound, 1),), 0, dtype=_torch_bits(dtype), device="cuda")
    d_stream.view(torch.uint8).fill_(fill)
    d_len = torch.full((1,), -1, dtype=torch.int32, device="cuda") if want_length_tensor else None
    comp.compress(d_in, shape, d_stream, d_len)
    torch.cuda.synchronize()
    host = d_stream.cpu().numpy().view(nz.api.bits_numpy_dtype(dtype))
    if want_length_tensor:
        n = int(d_len.cpu().numpy().view(np.uint32)[0])
        return host[:n].copy(), host
    return None, host


def gpu_decompress(stream: np.ndarray, dtype, shape, decompressor=None):
    import torch
    dec = decompressor or nz.make_cuda_decompressor(dtype, len(shape))
    words = np.ascontiguousarray(stream).view(np.int32 if np.dtype(dtype) == np.float32 else np.int64)
    d_stream = torch.from_numpy(words.copy()).cuda() if words.size else torch.empty((1,), dtype=_torch_bits(dtype), device="cuda")
    n = int(np.prod(shape)) if len(shape) else 0
    d_out = torch.full((max(n, 1),), 0, dtype=_torch_bits(dtype), device="cuda")
    d_out.view(torch.uint8).fill_(0xA5)
    dec.decompress(d_stream, d_out, shape)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy()[:n].view(dtype).reshape(shape)
    return out
