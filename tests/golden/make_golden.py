"""Regenerates tests/golden/golden.json and tests/golden/cubes.npz from the UNMODIFIED reference
CPU codec (oracle/_ref/libndzip_ref.so, built by `make -C oracle ref` from /root/reference).

Run in the build container only (it needs /root/reference):  python tests/golden/make_golden.py

golden.json rows: generator name + kwargs, dtype, shape, CRC32 of the input bytes (guards against
generator drift), stream length in words, CRC32 (zlib/IEEE) of the stream bytes, first four words.
The first eight rows are the table of SURVEY.md §8(c) and must reproduce it verbatim.
cubes.npz: for every profile, one full single-cube stream (input = `hashed`, seed 11) so that a
mismatch can be localised word by word without the reference being present.
"""
import json
import os
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from ndzip_b200 import synth  # noqa: E402
from oracle import Reference  # noqa: E402

SIDE = {1: 4096, 2: 64, 3: 16}


def cases():
    # (generator, kwargs, dtype, shape)
    yield from [
        ("ramp", {}, "float32", (8195,)), ("ramp", {}, "float32", (131, 131)), ("ramp", {}, "float32", (35, 35, 35)),
        ("ramp", {}, "float64", (12291,)), ("ramp", {}, "float64", (195, 195)), ("ramp", {}, "float64", (51, 51, 51)),
        ("ramp", {}, "float32", (100,)), ("ramp", {}, "float64", (15, 15, 15)),
    ]
    for dt in ("float32", "float64"):
        for dims in (1, 2, 3):
            s = SIDE[dims]
            for n in (s, 3 * s if dims > 1 else 2 * s, 4 * s - 1):
                shape = (n,) * dims
                yield ("hashed", {"seed": 3}, dt, shape)
            yield ("quantised", {"seed": 5}, dt, (2 * s + 3,) * dims)
            yield ("raw_bits", {"seed": 9}, dt, (s + 1,) * dims)
            yield ("poly", {}, dt, (2 * s,) * dims)
            yield ("zeros", {}, dt, (s,) * dims)
            yield ("ramp", {}, dt, (1,) * dims)
        # anisotropic shapes: borders in some dimensions only
        yield ("hashed", {"seed": 4}, dt, (130, 64))
        yield ("hashed", {"seed": 4}, dt, (64, 130))
        yield ("hashed", {"seed": 4}, dt, (33, 16, 48))
        yield ("hashed", {"seed": 4}, dt, (16, 35, 32))
        yield ("hashed", {"seed": 4}, dt, (32, 16, 21))
        yield ("hashed", {"seed": 4}, dt, (70, 10))       # one dimension shorter than a cube: all border
        yield ("hashed", {"seed": 4}, dt, (9, 40, 40))


def generate(name, kwargs, dtype, shape):
    return synth.make(name, shape, dtype, **kwargs)


def main():
    ref = Reference()
    rows = []
    for name, kwargs, dtype, shape in cases():
        data = generate(name, kwargs, dtype, shape)
        stream = ref.compress(data, threads=1)
        stream_mt = ref.compress(data, threads=3)
        assert np.array_equal(stream, stream_mt), (name, dtype, shape)
        back, consumed = ref.decompress(stream, dtype, shape)
        assert consumed == stream.size and back.tobytes() == data.tobytes(), (name, dtype, shape)
        rows.append({
            "generator": name, "kwargs": kwargs, "dtype": dtype, "shape": list(shape),
            "input_crc32": "%08x" % zlib.crc32(data.tobytes()),
            "bound": ref.compressed_length_bound(dtype, shape),
            "stream_words": int(stream.size),
            "stream_crc32": "%08x" % zlib.crc32(stream.tobytes()),
            "first_words": ["%x" % int(w) for w in stream[:4]],
        })
    here = os.path.dirname(os.path.abspath(__file__))
    with open(os.path.join(here, "golden.json"), "w") as f:
        json.dump({"reference_commit": "ff4e6702", "rows": rows}, f, indent=1)
    cubes = {}
    for dt in ("float32", "float64"):
        for dims in (1, 2, 3):
            data = synth.hashed((SIDE[dims],) * dims, dt, seed=11)
            data.reshape(-1)[: (32 if dt == "float32" else 64)] = 0  # first chunk all-zero (codec_profile_test.inl:54-58)
            cubes[f"{dt}_{dims}d_stream"] = ref.compress(data, threads=1)
    np.savez_compressed(os.path.join(here, "cubes.npz"), **cubes)
    print(f"wrote {len(rows)} rows, {len(cubes)} cube streams")


if __name__ == "__main__":
    main()
