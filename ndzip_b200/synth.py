"""Deterministic synthetic grids for parity tests and bench.py.

All integer-arithmetic generators (``ramp``, ``hashed``, ``raw_bits``, ``engineered_cube``) are
bit-reproducible on any host and are what the committed golden fixtures are keyed on. ``smooth``
uses libm ``sin`` and may differ in the last bit across hosts: generate it once per run and feed
the same buffer to every implementation being compared.
"""
from __future__ import annotations

import numpy as np

_MASK64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x: np.ndarray) -> np.ndarray:
    """Counter-based hash (Steele et al. splitmix64 finaliser) over uint64 arrays."""
    with np.errstate(over="ignore"):
        z = (x.astype(np.uint64) + np.uint64(0x9E3779B97F4A7C15)) & _MASK64
        z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _MASK64
        z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _MASK64
        return z ^ (z >> np.uint64(31))


def _count(shape) -> int:
    n = 1
    for s in shape:
        n *= int(s)
    return n


def ramp(shape, dtype) -> np.ndarray:
    """The golden-table input of SURVEY.md §8(c): v[i] = T((i*7 + (i/5)*3) % 1024) / 16."""
    i = np.arange(_count(shape), dtype=np.uint64)
    v = ((i * np.uint64(7) + (i // np.uint64(5)) * np.uint64(3)) % np.uint64(1024)).astype(dtype) / np.dtype(dtype).type(16)
    return v.reshape(shape)


def hashed(shape, dtype, seed: int = 1) -> np.ndarray:
    """Uniform [0,1) from splitmix64(seed, index): incompressible mantissas, r ~ 1."""
    i = np.arange(_count(shape), dtype=np.uint64)
    h = splitmix64(i ^ (np.uint64(seed) << np.uint64(32)))
    if np.dtype(dtype) == np.float32:
        v = (h >> np.uint64(40)).astype(np.float32) * np.float32(2.0 ** -24)
    else:
        v = (h >> np.uint64(11)).astype(np.float64) * np.float64(2.0 ** -53)
    return v.reshape(shape)


def raw_bits(shape, dtype, seed: int = 1) -> np.ndarray:
    """Arbitrary bit patterns (NaNs/denormals included) viewed as floats: the codec is integer-only."""
    i = np.arange(_count(shape), dtype=np.uint64)
    h = splitmix64(i ^ (np.uint64(seed) << np.uint64(32)))
    if np.dtype(dtype) == np.float32:
        return (h >> np.uint64(32)).astype(np.uint32).view(np.float32).reshape(shape)
    return h.view(np.float64).reshape(shape)


def quantised(shape, dtype, seed: int = 1, levels: int = 37) -> np.ndarray:
    """Small integers as floats: many zero low planes and holes in the chunk heads."""
    i = np.arange(_count(shape), dtype=np.uint64)
    h = splitmix64(i ^ (np.uint64(seed) << np.uint64(32)))
    return ((h >> np.uint64(20)) % np.uint64(levels)).astype(dtype).reshape(shape)


def engineered_cube(bits_dtype, seed: int = 7) -> np.ndarray:
    """One 4096-word cube of *residual bits* with engineered zero bit-columns and zero words, after
    the idea of the reference test src/test/codec_profile_test.inl:561-567."""
    B = np.dtype(bits_dtype).itemsize * 8
    i = np.arange(4096, dtype=np.uint64)
    h = splitmix64(i ^ (np.uint64(seed) << np.uint64(32)))
    w = h.astype(bits_dtype) if B == 64 else (h >> np.uint64(32)).astype(np.uint32)
    chunk = (i // np.uint64(B)).astype(np.uint64)
    one = np.dtype(bits_dtype).type(1)
    for idx in (0, 12, 13, 29, B - 2):
        sh = ((np.uint64(idx) * chunk) % np.uint64(B)).astype(bits_dtype)
        w &= ~(one << sh)
    w = w.reshape(-1, B)
    for idx in (0, 12, 13, 29, B - 2):
        w[:, idx] = 0
    return w.reshape(-1).copy()


def smooth_constants(seed: int, dims: int, modes: int = 6):
    """(amplitude, integer wave vector, phase) per mode — shared with bench.py's device generator."""
    consts = splitmix64(np.arange(8 * (modes + 1), dtype=np.uint64) ^ (np.uint64(seed) << np.uint64(32)))
    u01 = (consts >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    out = []
    for k in range(1, modes + 1):
        wave = [float(np.floor(u01[8 * k + 1 + d] * (k + 1))) for d in range(dims)]  # components 0..k
        if not any(wave):
            wave[-1] = float(k)
        out.append((k ** (-5.0 / 3.0), wave, 2 * np.pi * float(u01[8 * k])))
    return out


def smooth(shape, dtype, seed: int = 0x5EED0002, noise: float = 1e-4, coord_shape=None, index_offset: int = 0,
           threads: int = 1) -> np.ndarray:
    """Turbulence-like field: six sine modes with a k^(-5/3) amplitude
    spectrum, hashed integer wave vectors |m_d| <= k and phases, plus `noise` * u(index), u in [-1,1).
    Evaluated in float64, rounded to ``dtype``. Mid-range ratios (~0.6 for 3D float32).
    ``coord_shape`` (default: ``shape``) is the extent the coordinates are normalised by and ``index_offset`` the
    linear index of element 0, so that any run of rows of a larger grid can be generated on its own; ``threads`` > 1
    evaluates slabs side by side (numpy releases the GIL inside its loops)."""
    dims = len(shape)
    coord_shape = tuple(coord_shape) if coord_shape is not None else tuple(shape)
    n = _count(shape)
    if n == 0:
        return np.zeros(shape, dtype=dtype)
    modes = smooth_constants(seed, dims)
    out = np.empty(n, dtype=dtype)
    strides = [1] * dims
    for d in range(dims - 2, -1, -1):
        strides[d] = strides[d + 1] * int(shape[d + 1])
    step = 1 << 22  # slabs bound the temporaries

    def fill(lo):
        idx = np.arange(lo, min(n, lo + step), dtype=np.uint64) + np.uint64(index_offset)
        coords = []
        rem = idx
        for d in range(dims):
            coords.append((rem // np.uint64(strides[d])).astype(np.float64) / float(coord_shape[d]))
            rem = rem % np.uint64(strides[d])
        acc = np.zeros(idx.size, dtype=np.float64)
        for amp, wave, phase in modes:
            arg = np.zeros(idx.size, dtype=np.float64)
            for d in range(dims):
                if wave[d]:
                    arg += wave[d] * coords[d]
            acc += amp * np.sin(2 * np.pi * arg + phase)
        jitter = (splitmix64(idx ^ (np.uint64(seed) << np.uint64(32))) >> np.uint64(11)).astype(np.float64) * 2.0 ** -52 - 1.0
        out[lo:lo + idx.size] = (acc + noise * jitter).astype(dtype)

    starts = range(0, n, step)
    if threads > 1 and len(starts) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(max_workers=threads) as pool:
            list(pool.map(fill, starts))
    else:
        for lo in starts:
            fill(lo)
    return out.reshape(shape)


def poly(shape, dtype) -> np.ndarray:
    """Integer-only smooth stand-in: a low-order polynomial in the coordinates, /64 (bit-reproducible)."""
    idx = np.indices(shape, dtype=np.int64)
    acc = np.zeros(shape, dtype=np.int64)
    for d, g in enumerate(idx):
        acc += (d + 1) * g * g + 3 * g
    return (acc.astype(dtype) / np.dtype(dtype).type(64)).reshape(shape)


GENERATORS = {
    "poly": poly,
    "ramp": ramp,
    "hashed": hashed,
    "raw_bits": raw_bits,
    "quantised": quantised,
    "smooth": smooth,
    "zeros": lambda shape, dtype, **kw: np.zeros(shape, dtype=dtype),
    "constant": lambda shape, dtype, **kw: np.full(shape, 3.25, dtype=dtype),
}


def make(name: str, shape, dtype, **kw) -> np.ndarray:
    return GENERATORS[name](shape, dtype, **kw)
