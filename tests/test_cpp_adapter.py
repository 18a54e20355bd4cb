"""The C++ drop-in boundary: tests/cpp/adapter_test.cc is written against the reference's public API
and built (a) against this repo's include/ndzip headers and (b) against the REFERENCE's own headers
(only where /root/reference exists), both linked to libndzip_b200.so. Building is a CPU test; the
binaries (in build/, which travels to the GPU box) run under -m gpu."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BUILD = os.path.join(ROOT, "build")
SRC = os.path.join(ROOT, "tests", "cpp", "adapter_test.cc")
CUDA = "/usr/local/cuda"
VARIANTS = {
    "own": [f"-I{ROOT}/include"],
    "ref": ["-I/root/reference/include", "-DNDZIP_CUDA_SUPPORT=1"],
}


def _build(variant):
    from ndzip_b200 import build as nzbuild
    import oracle
    nzbuild.build()
    oracle.build("oracle")
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, f"adapter_test_{variant}")
    libdir = os.path.join(ROOT, "ndzip_b200")
    oradir = os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++17", "-O1", SRC, *VARIANTS[variant], f"-I{CUDA}/include", "-o", out,
           f"-L{libdir}", "-lndzip_b200", f"-L{oradir}", "-lndzip_oracle", f"-L{CUDA}/lib64", "-lcudart",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{oradir}", f"-Wl,-rpath,{CUDA}/lib64"]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(cmd, check=True, env=env)
    return out


def test_builds_against_own_headers():
    assert os.path.exists(_build("own"))


def test_builds_against_reference_headers():
    if not os.path.isdir("/root/reference/include/ndzip"):
        pytest.skip("/root/reference not present")
    assert os.path.exists(_build("ref"))


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["own", "ref"])
def test_adapter_program_passes(variant):
    exe = os.path.join(BUILD, f"adapter_test_{variant}")
    if not os.path.exists(exe):
        if variant == "ref":
            pytest.skip("reference-header build was not prebuilt")
        exe = _build(variant)
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PASS" in res.stdout, res.stdout + res.stderr


# ---- multi-GPU data plane through the C ABI alone (tests/cpp/dist_test.cc): one process, one thread per GPU ----

DIST_SRC = os.path.join(ROOT, "tests", "cpp", "dist_test.cc")


def _build_dist():
    from ndzip_b200 import build as nzbuild
    import oracle
    nzbuild.build()
    oracle.build("oracle")
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "dist_test")
    libdir = os.path.join(ROOT, "ndzip_b200")
    oradir = os.path.join(ROOT, "oracle")
    cmd = ["g++", "-std=c++17", "-O1", "-pthread", DIST_SRC, f"-I{ROOT}/include", f"-I{CUDA}/include", "-o", out,
           f"-L{libdir}", "-lndzip_b200", f"-L{oradir}", "-lndzip_oracle", f"-L{CUDA}/lib64", "-lcudart",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{oradir}", f"-Wl,-rpath,{CUDA}/lib64"]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(cmd, check=True, env=env)
    return out


def test_dist_program_builds():
    assert os.path.exists(_build_dist())


@pytest.mark.gpu
def test_dist_program_passes_on_all_visible_gpus():
    """world = every visible GPU (1 on the single-GPU test box: the slab / gather arithmetic with one rank; the
    N > 1 run is scripts/gpu_multi.sh under `gpurun --gpus N`)."""
    exe = os.path.join(BUILD, "dist_test")
    if not os.path.exists(exe):
        exe = _build_dist()
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PASS" in res.stdout, res.stdout + res.stderr


# ---- sharded stream container through the C ABI from plain C (tests/cpp/container_test.c): host only, runs here ----

def test_container_program_in_c_passes(tmp_path):
    from ndzip_b200 import build as nzbuild
    import oracle
    nzbuild.build()
    oracle.build("oracle")
    os.makedirs(BUILD, exist_ok=True)
    out = os.path.join(BUILD, "container_test")
    libdir = os.path.join(ROOT, "ndzip_b200")
    oradir = os.path.join(ROOT, "oracle")
    cmd = ["gcc", "-std=c11", "-O1", "-Wall", "-Wextra", "-Werror", os.path.join(ROOT, "tests", "cpp", "container_test.c"),
           f"-I{ROOT}/include", "-o", out, f"-L{libdir}", "-lndzip_b200", f"-L{oradir}", "-lndzip_oracle",
           f"-Wl,-rpath,{libdir}", f"-Wl,-rpath,{oradir}"]
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.run(cmd, check=True, env=env)
    res = subprocess.run([out, str(tmp_path / "c.ndzs")], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0 and "PASS" in res.stdout, res.stdout + res.stderr
