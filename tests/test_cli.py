"""tools/ndzip_compress.cc — the counterpart of the reference's `compress` tool (reference
src/compress/compress.cc:130-229): same options, raw file = arrays back to back, compressed file = their
streams back to back. CPU tests cover the option parsing and the loud failure without a device; the -m gpu
tests compare the files it writes with the oracle's streams byte for byte and round-trip them."""
import os
import subprocess

import numpy as np
import pytest

from ndzip_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def tool():
    from ndzip_b200 import build
    build.build()
    return build.build_tool()


def run(tool, *args, stdin=None):
    return subprocess.run([tool, *args], input=stdin, stdout=subprocess.PIPE, stderr=subprocess.PIPE, timeout=300)


def test_help_lists_the_reference_options(tool):
    r = run(tool, "--help")
    assert r.returncode == 0
    out = r.stdout.decode()
    for opt in ("--decompress", "--array-size", "--data-type", "--target", "--threads", "--input", "--output", "--no-mmap"):
        assert opt in out


@pytest.mark.parametrize("args,message", [
    ((), "--array-size"),                                   # required (compress.cc:144)
    (("-n", "1", "2", "3", "4"), "between 1 and 3 dimensions"),   # compress.cc:188-190
    (("-n", "64", "-t", "half"), "Invalid data type half"),       # compress.cc:201
    (("-n", "64", "-e", "cpu"), "Unimplemented target cpu"),      # compress.cc:185: no CPU encoder in this library
    (("-n", "64", "--frobnicate"), "unrecognised option"),
])
def test_usage_errors_exit_with_failure(tool, args, message):
    r = run(tool, *args)
    assert r.returncode == 1
    assert message in r.stderr.decode()
    assert "Usage:" in r.stderr.decode()


def test_fails_loudly_without_a_device(tool, tmp_path):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    src = tmp_path / "in.bin"
    src.write_bytes(np.zeros(4096, np.float32).tobytes())
    r = run(tool, "-n", "4096", "-i", str(src), "-o", str(tmp_path / "out.ndz"))
    assert r.returncode == 1 and b"CUDA" in r.stderr  # no CPU fallback


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,shape", [("float32", (40, 52, 68)), ("float64", (130, 200)), ("float32", (3 * 4096 + 9,))])
@pytest.mark.parametrize("mmap", [True, False])
def test_files_equal_oracle_streams_and_round_trip(tool, oracle, tmp_path, dtype, shape, mmap):
    chunks = [synth.make(gen, shape, dtype, seed=21 + i) for i, gen in enumerate(("smooth", "hashed", "quantised"))]
    raw = b"".join(c.tobytes() for c in chunks)
    expect = b"".join(oracle.compress(c).tobytes() for c in chunks)
    src, packed, back = tmp_path / "in.bin", tmp_path / "out.ndz", tmp_path / "back.bin"
    src.write_bytes(raw)
    common = ["-n", *map(str, shape), "-t", "float" if dtype == "float32" else "double"] + ([] if mmap else ["--no-mmap"])
    r = run(tool, *common, "-i", str(src), "-o", str(packed))
    assert r.returncode == 0, r.stderr.decode()
    assert b"3 chunks" in r.stderr and b"ratio" in r.stderr
    assert packed.read_bytes() == expect
    r = run(tool, "-d", *common, "-i", str(packed), "-o", str(back))
    assert r.returncode == 0, r.stderr.decode()
    assert back.read_bytes() == raw


@pytest.mark.gpu
def test_pipes_and_partial_chunks(tool, oracle):
    shape = (2 * 4096 + 77,)
    data = synth.make("smooth", shape, "float32", seed=5)
    r = run(tool, "-n", str(shape[0]), stdin=data.tobytes())  # stdin -> stdout
    assert r.returncode == 0, r.stderr.decode()
    assert r.stdout == oracle.compress(data).tobytes()
    r2 = run(tool, "-d", "-n", str(shape[0]), stdin=r.stdout)
    assert r2.returncode == 0 and r2.stdout == data.tobytes()
    r3 = run(tool, "-n", str(shape[0]), stdin=data.tobytes()[:-4])  # reference io.cc:62
    assert r3.returncode == 1 and b"not a multiple of the chunk size" in r3.stderr


# ---- tools/ndzip_benchmark.cc: the ndzip-cuda column of the reference's benchmark driver -------------------------
BENCH_HEADER = ("dataset;data type;dimensions;algorithm;tunable;number of threads;compression times (microseconds);"
                "decompression times (microseconds);uncompressed bytes;compressed bytes")  # reference benchmark.cc:1487-1489


@pytest.fixture(scope="module")
def bench_tool(tool):
    from ndzip_b200 import build
    return build.BENCHMARK_TOOL


def test_benchmark_driver_usage(bench_tool, tmp_path):
    assert run(bench_tool, "--help").returncode == 0
    r = run(bench_tool)
    assert r.returncode == 1 and b"csv-file" in r.stderr
    r = run(bench_tool, str(tmp_path / "missing.csv"))
    assert r.returncode == 1
    r = run(bench_tool, "-a", "zfp", str(tmp_path / "x.csv"))
    assert r.returncode == 1 and b"ndzip-cuda" in r.stderr
    bad = tmp_path / "bad.csv"
    bad.write_text("a.f32;half;16\n")
    r = run(bench_tool, str(bad))
    assert r.returncode == 1 and b"invalid data type" in r.stderr


@pytest.mark.gpu
def test_benchmark_driver_rows_match_the_reference_schema(bench_tool, oracle, tmp_path):
    import io
    import pandas as pd
    sets = [("smooth.f32", "float32", (48, 64, 80)), ("grid.f64", "float64", (130, 200)), ("line.f32", "float32", (5 * 4096 + 3,))]
    lines = []
    for name, dtype, shape in sets:
        synth.smooth(shape, dtype, seed=31).tofile(tmp_path / name)
        lines.append(f"{name};{'float' if dtype == 'float32' else 'double'};{' '.join(map(str, shape))}")
    (tmp_path / "sets.csv").write_text("\n".join(lines) + "\n")
    r = run(bench_tool, str(tmp_path / "sets.csv"), "-r", "3", "-t", "1", "-a", "ndzip-cuda")
    assert r.returncode == 0, r.stderr.decode()
    out = r.stdout.decode()
    assert out.splitlines()[0] == BENCH_HEADER
    df = pd.read_csv(io.StringIO(out), sep=";")  # what the reference's plot_benchmark.py does
    assert list(df["dataset"]) == [s[0] for s in sets]
    assert set(df["algorithm"]) == {"ndzip-cuda"}
    for (name, dtype, shape), (_, row) in zip(sets, df.iterrows()):
        data = np.fromfile(tmp_path / name, dtype=dtype).reshape(shape)
        assert row["dimensions"] == len(shape) and row["data type"] == ("float" if dtype == "float32" else "double")
        assert row["uncompressed bytes"] == data.nbytes
        assert row["compressed bytes"] == oracle.compress(data).nbytes
        assert len(str(row["compression times (microseconds)"]).split(",")) >= 3
        assert len(str(row["decompression times (microseconds)"]).split(",")) >= 3
