"""Builds libndzip_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.environ.get("NDZB_CSRC") or os.path.join(HERE, "csrc")  # NDZB_CSRC: build another revision of the sources (A/B runs)
LIB = os.path.join(HERE, "libndzip_b200.so")
SOURCES = ["ndzb_kernels.cu", "ndzb_capi.cu", "ndzb_dist.cu", "ndzb_container.cu", "ndzip_adapter.cu"]
HEADERS = ["ndzb_cube.cuh", "ndzb_ptx.cuh", "ndzb_kernels.cuh"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-shared", "-cudart", "static",
    "-diag-suppress=177",  # "declared but never referenced": locals that only some template instantiations use
]


def _nvcc() -> str:
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(nvcc):
        raise RuntimeError("nvcc not found; cannot build libndzip_b200.so")
    return nvcc


def _stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, f) for f in SOURCES + HEADERS if os.path.exists(os.path.join(CSRC, f))]
    deps += [os.path.join(os.path.dirname(HERE), "include", "ndzip_b200.h")]
    inc = os.path.join(os.path.dirname(HERE), "include", "ndzip")
    if os.path.isdir(inc):
        deps += [os.path.join(inc, f) for f in os.listdir(inc)]
    return any(os.path.getmtime(d) > t for d in deps)


TOOLS_DIR = os.path.join(os.path.dirname(HERE), "tools")
BIN_DIR = os.path.join(HERE, "bin")
TOOLS = {"ndzip-compress": "ndzip_compress.cc", "ndzip-benchmark": "ndzip_benchmark.cc"}
TOOL = os.path.join(BIN_DIR, "ndzip-compress")
BENCHMARK_TOOL = os.path.join(BIN_DIR, "ndzip-benchmark")


def build_tool(force: bool = False) -> str:
    """Compile the command-line tools (tools/*.cc, counterparts of the reference's `compress` and `benchmark`)
    against the C ABI with plain g++. Returns the path of ndzip-compress."""
    os.makedirs(BIN_DIR, exist_ok=True)
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    for name, src in TOOLS.items():
        out, src = os.path.join(BIN_DIR, name), os.path.join(TOOLS_DIR, src)
        deps = [src, LIB, os.path.join(os.path.dirname(HERE), "include", "ndzip_b200.h")]
        if not force and os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps if os.path.exists(d)):
            continue
        subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-o", out, src, "-L" + HERE, "-lndzip_b200",
                        "-Wl,-rpath,$ORIGIN/.."], check=True, env=env)
    return TOOL


def build(force: bool = False, verbose: bool = False, out: str | None = None, extra_flags: list[str] | None = None) -> str:
    """Compile the CUDA library if missing or older than its sources. Returns the .so path.
    `out` / `extra_flags` build a variant elsewhere (tuning builds: extra_flags=["-DNDZB_TUNING"])."""
    lib = out or LIB
    if out is None and not force and not _stale():
        return lib
    from concurrent.futures import ThreadPoolExecutor
    srcs = [os.path.join(CSRC, f) for f in SOURCES if os.path.exists(os.path.join(CSRC, f))]
    flags = list(extra_flags or [])
    if os.environ.get("NDZB_EXTRA_NVCC_FLAGS"):  # tuning builds, e.g. -DNDZB_TUNING -DNDZB_DESC_STRIDE=4
        flags += os.environ["NDZB_EXTRA_NVCC_FLAGS"].split()
    if verbose:
        flags.append("-Xptxas=-v")
    # nvcc must use the system g++ (the CXX in this image's environment points at a wrapper without libgomp specs)
    env = dict(os.environ)
    env.pop("CXX", None)
    env.pop("CC", None)
    objdir = os.path.join(os.path.dirname(HERE), "build", "obj_" + os.path.basename(lib).replace(".so", ""))
    os.makedirs(objdir, exist_ok=True)
    compile_flags = [f for f in NVCC_FLAGS if f != "-shared"]
    inc = ["-I", os.path.join(os.path.dirname(HERE), "include")]

    def compile_one(src):
        obj = os.path.join(objdir, os.path.basename(src) + ".o")
        subprocess.run([_nvcc(), *flags, *compile_flags, *inc, "-c", "-o", obj, src], check=True, env=env)
        return obj

    with ThreadPoolExecutor(max_workers=len(srcs)) as pool:  # the three translation units compile side by side
        objs = list(pool.map(compile_one, srcs))
    subprocess.run([_nvcc(), *NVCC_FLAGS, "-o", lib, *objs, "-ldl"], check=True, env=env)
    return lib


if __name__ == "__main__":
    print(build(force=True, verbose=True))
    print(build_tool(force=True))
