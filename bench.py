#!/usr/bin/env python
"""bench.py — measures BASELINE.json's metric for the ndzip hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfgN]

metric  : uncompressed GB/s of one compress + decompress round trip (uncompressed bytes / (t_c + t_d)),
          whole-job aggregate over all ranks; compress and decompress are also reported separately.
workload: N=1 -> BASELINE.json configs[1]: 3D fp32 512^3 synthetic turbulence-like grid (512 MiB).
          N>1 -> weak scaling: every rank owns one 512^3 slab of a (512*N) x 512 x 512 grid; the data path has
          one exchange step (NCCL all-gather of the per-rank stream lengths + header fix-up, inside the library,
          on a side stream). After the timed loop the N>1 run also measures BASELINE configs[3] (3D fp64, 128 x 1024^2
          per rank = 1024^3 at N=8) and configs[4] (1D fp32, 256 Mi per rank = 2 Gi at N=8) with and without the final
          NCCL stream gather and asserts that the gathered stream equals the one ONE GPU produces for the whole grid
          ("configs" in the JSON line; also "stream_gather" for the headline grid).
A "step" is one pass of the hot path (compress, offset exchange, decompress) over resident inputs.
Inputs (512 MiB per rank) are larger than the 126 MB L2, so no explicit L2 flush is needed.

One JSON line on stdout (rank 0). Keys beyond the base contract:
  roofline     dominant kernel (compress_ws_kernel): algorithmic bytes (input + stream) / CUDA-event time
               of that launch alone, against MEASURED_PEAKS.json's hbm_gbs.
  cpu_baseline the UNMODIFIED reference CPU codec (oracle/_ref, OpenMP, all host threads) on a bounded
               sample of the same grid.
  e2e          same metric through the host-pointer offloader API (pinned host buffers, H2D + D2H
               inside the timed region).
`--impl reference` times the reference CPU implementation itself (rank 0 only).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dtype, per-rank shape, description)
    "cfg1": ("float32", (1 << 24,), "1D fp32 16 Mi elements"),
    "cfg2": ("float32", (512, 512, 512), "3D fp32 512^3 synthetic turbulence-like grid"),
    "cfg3": ("float64", (8192, 8192), "2D fp64 8192x8192 grid"),
    "cfg4": ("float64", (128, 1024, 1024), "3D fp64 1024^3 grid, one 128-plane slab per rank"),
    "cfg5": ("float32", (1 << 28,), "1D fp32 2 Gi-element stream, 256 Mi elements per rank"),
}
SEED = 0x5EED0002


# --------------------------------------------------------------------------------------------------
# synthetic input, generated on the device (same field as ndzip_b200.synth.smooth; torch's sin may
# differ from libm in the last bit, so every implementation compared in one run gets this buffer)

def _splitmix64_torch(x):
    import torch
    m = (1 << 64) - 1

    def c(v):  # as signed int64 constant
        v &= m
        return v - (1 << 64) if v >= (1 << 63) else v

    def lsr(v, s):  # logical shift right on int64
        return (v >> s) & ((1 << (64 - s)) - 1)

    z = x + c(0x9E3779B97F4A7C15)
    z = (z ^ lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
    z = (z ^ lsr(z, 27)) * c(0x94D049BB133111EB)
    return z ^ lsr(z, 31)


def make_device_input(dtype, shape, seed=SEED, noise=1e-4, device="cuda", index_offset=0):
    """Turbulence-like field on the device. `index_offset` shifts the slowest coordinate / hash index so
    that the ranks of a multi-GPU run hold different slabs of one larger grid."""
    import torch
    from ndzip_b200 import synth
    tdt = torch.float32 if np.dtype(dtype) == np.float32 else torch.float64
    dims = len(shape)
    modes = synth.smooth_constants(seed, dims)
    n0 = shape[0]
    rest = int(np.prod(shape[1:])) if dims > 1 else 1
    out = torch.empty(shape, dtype=tdt, device=device)
    flat = out.view(n0, rest) if dims > 1 else out.view(-1, 1)
    rows_per_chunk = max(1, (1 << 24) // rest) if dims > 1 else (1 << 24)
    total0 = n0
    seed_c = (seed << 32) & ((1 << 63) - 1)
    for lo in range(0, total0, rows_per_chunk):
        hi = min(total0, lo + rows_per_chunk)
        if dims == 1:
            idx = torch.arange(lo, hi, dtype=torch.int64, device=device)
            coords = [(idx + index_offset).to(torch.float64) / float(shape[0])]
            lin = idx + index_offset
        else:
            i0 = torch.arange(lo, hi, dtype=torch.int64, device=device)
            inner = torch.arange(rest, dtype=torch.int64, device=device)
            lin = ((i0 + index_offset)[:, None] * rest + inner[None, :]).reshape(-1)
            coords = [((i0 + index_offset).to(torch.float64) / float(shape[0]))[:, None].expand(hi - lo, rest).reshape(-1)]
            if dims == 2:
                coords.append((inner.to(torch.float64) / float(shape[1]))[None, :].expand(hi - lo, rest).reshape(-1))
            else:
                y = (inner // shape[2]).to(torch.float64) / float(shape[1])
                x = (inner % shape[2]).to(torch.float64) / float(shape[2])
                coords.append(y[None, :].expand(hi - lo, rest).reshape(-1))
                coords.append(x[None, :].expand(hi - lo, rest).reshape(-1))
        acc = torch.zeros(lin.numel(), dtype=torch.float64, device=device)
        for amp, wave, phase in modes:
            arg = torch.zeros_like(acc)
            for d in range(dims):
                if wave[d]:
                    arg += wave[d] * coords[d]
            acc += amp * torch.sin(2 * np.pi * arg + phase)
        h = _splitmix64_torch(lin ^ seed_c)
        jitter = ((h >> 11) & ((1 << 53) - 1)).to(torch.float64) * 2.0 ** -52 - 1.0
        vals = (acc + noise * jitter).to(tdt)
        if dims == 1:
            out[lo:hi] = vals
        else:
            flat[lo:hi] = vals.view(hi - lo, rest)
    return out


# --------------------------------------------------------------------------------------------------
# clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)

class ClockSampler:
    """Samples SM clock and clock-event reasons through NVML every few ms while the timed region runs."""
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown"}

    def __init__(self, gpu_index=0, period_s=0.004):
        self.gpu_index, self.period_s = gpu_index, period_s
        self.samples, self.reason_bits, self.power = [], 0, []
        self._stop = threading.Event()
        self.thread = None
        self.handle = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            try:  # CUDA_VISIBLE_DEVICES-proof: look the device up by PCI bus id
                bus = torch.cuda.get_device_properties(gpu_index).pci_bus_id
                dom = torch.cuda.get_device_properties(gpu_index).pci_domain_id
                dev = torch.cuda.get_device_properties(gpu_index).pci_device_id
                self.handle = pynvml.nvmlDeviceGetHandleByPciBusId(f"{dom:08x}:{bus:02x}:{dev:02x}.0")
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(gpu_index)
            self.nvml = pynvml
        except Exception:
            self.handle = None

    def _loop(self):
        nv = self.nvml
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM))
                try:
                    self.reason_bits |= nv.nvmlDeviceGetCurrentClocksEventReasons(self.handle)
                except Exception:
                    self.reason_bits |= nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle)
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.handle) / 1000.0)
            except Exception:
                pass
            time.sleep(self.period_s)

    def start(self):
        if self.handle is not None:
            self.thread = threading.Thread(target=self._loop, daemon=True)
            self.thread.start()

    def stop(self):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self._stop.set()
        self.thread.join(timeout=2)
        try:
            mx = self.nvml.nvmlDeviceGetMaxClockInfo(self.handle, self.nvml.NVML_CLOCK_SM)
        except Exception:
            mx = None
        reasons = sorted(name for bit, name in self.REASONS.items() if self.reason_bits & bit)
        return {"sm_mhz": statistics.median(self.samples) if self.samples else None, "sm_max_mhz": mx,
                "samples": len(self.samples), "power_w_max": max(self.power) if self.power else None, "reasons": reasons}


# --------------------------------------------------------------------------------------------------

def measured_hbm_peak():
    """HBM GB/s from the driver-written MEASURED_PEAKS.json (burst figure: the roofline kernel is timed alone, one launch
    between synchronisations), else B200_PROFILING.md's fallback. The file's schema is not ours, so any numeric entry
    whose key path mentions hbm is accepted, preferring one that also says burst."""
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            doc = json.load(f)
    except (OSError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"
    found = []

    def walk(node, trail):
        if isinstance(node, dict):
            for k, v in node.items():
                walk(v, trail + [str(k).lower()])
        elif isinstance(node, (int, float)) and not isinstance(node, bool):
            key = ".".join(trail)
            if "hbm" in key and "tf" not in key:
                v = float(node)
                if 500.0 < v < 20000.0:          # GB/s
                    found.append((key, v))
                elif 0.5 < v < 20.0:             # TB/s
                    found.append((key, v * 1000.0))

    walk(doc, [])
    if not found:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s; no hbm entry in MEASURED_PEAKS.json)"
    for want in ("burst", "hbm_gbs", ""):
        for key, v in found:
            if want in key:
                return v, f"measured (MEASURED_PEAKS.json {key})"
    return found[0][1], f"measured (MEASURED_PEAKS.json {found[0][0]})"


def ncu_traffic_per_launch(workload):
    """dram bytes read+written per compress launch from the committed ncu capture, if any."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(workload, {}).get("compress_dram_bytes")
    except (OSError, ValueError):
        return None


def host_input(dtype, shape, threads=0):
    """The workload's grid on the host (numpy twin of the device generator), generated with all cores."""
    from ndzip_b200 import synth
    return synth.smooth(shape, dtype, seed=SEED, threads=threads or (os.cpu_count() or 1))


def run_cpu_reference(dtype, shape, data, steps, warmup, threads=0):
    """Times the unmodified reference CPU codec (oracle/_ref; falls back to the C oracle port) on `data`.
    threads = 0: all host threads (make_cpu_offloader(dims, physical_concurrency), the reference's `-e cpu-mt`);
    threads = 1: the serial encoder (`-e cpu`, reference src/ndzip/cpu_factory.cc:89-92)."""
    from oracle import get_oracle, get_reference
    ref = get_reference()
    bits = np.uint32 if np.dtype(dtype) == np.float32 else np.uint64
    times_c, times_d = [], []
    if ref is not None:
        kind = "reference"
        cores = ref.physical_concurrency() if threads == 0 else threads
        bound = ref.compressed_length_bound(dtype, shape)
        stream = np.zeros(bound, dtype=bits)
        back = np.empty(shape, dtype=dtype)
        for it in range(warmup + steps):
            stream[: 1 + bound // 64] = 0
            t0 = time.perf_counter()
            n = ref.compress_into(data, stream, threads=threads)
            t1 = time.perf_counter()
            ref.decompress_into(stream[:n], back, threads=threads)
            t2 = time.perf_counter()
            if it >= warmup:
                times_c.append(t1 - t0)
                times_d.append(t2 - t1)
        # The reference's OpenMP compressor has a data race (cpu_codec.inl:826-836) that occasionally
        # corrupts its own stream; timing is unaffected, so this is reported rather than fatal.
        roundtrip_ok = back.tobytes() == data.tobytes()
    else:
        kind, cores, roundtrip_ok = "port", 1, True
        oracle = get_oracle()
        for it in range(max(1, warmup // 3) + max(1, steps // 3)):
            t0 = time.perf_counter()
            s = oracle.compress(data)
            t1 = time.perf_counter()
            oracle.decompress(s, dtype, shape)
            t2 = time.perf_counter()
            times_c.append(t1 - t0)
            times_d.append(t2 - t1)
        n = s.size
    nbytes = data.nbytes
    tc, td = statistics.median(times_c), statistics.median(times_d)
    return {
        "kind": kind, "cores": int(cores), "value": nbytes / (tc + td) / 1e9,
        "compress_gbs": nbytes / tc / 1e9, "decompress_gbs": nbytes / td / 1e9,
        "ratio": n * np.dtype(bits).itemsize / nbytes, "ms_per_step": (tc + td) * 1e3,
        "steps": len(times_c), "roundtrip_ok": bool(roundtrip_ok),
    }


def l2_note(nbytes):
    if nbytes > 2 * 126e6:
        return f"inputs ({nbytes >> 20} MiB per rank) exceed the 126 MB L2; no explicit flush"
    return (f"inputs are {nbytes >> 20} MiB per rank: comparable to the 126 MB L2, so part of every step is served from L2 "
            "(the BASELINE config is this small; no flush between steps)")


def time_call(fn, reps):
    import torch
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return ts


def pcie_ceiling(h2d_bytes, d2h_bytes, dev, reps=3):
    """Pinned cudaMemcpyAsync of the step's H2D bytes on one stream while the step's D2H bytes go the other way on a
    second stream: the time no host-pointer API can beat on this box. Returns milliseconds."""
    import torch
    h_a = torch.empty(h2d_bytes, dtype=torch.uint8, pin_memory=True)
    d_a = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    h_b = torch.empty(d2h_bytes, dtype=torch.uint8, pin_memory=True)
    d_b = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    best = None
    for _ in range(reps + 1):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        with torch.cuda.stream(s1):
            d_a.copy_(h_a, non_blocking=True)
        with torch.cuda.stream(s2):
            h_b.copy_(d_b, non_blocking=True)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3
        best = ms if best is None else min(best, ms)
    return best


def sharded_container_record(nz, nzd, dist, dtype, global_shape, slab_shape, rank, world, dev, tbits, d_in, d_stream, n_words,
                             d_global, total_words):
    """N > 1: the headline grid through the sharded container instead of the gather. Timed with the wall clock (file
    I/O), max over ranks: (a) D2H of the slab stream + pwrite of the rank's segment, (b) table read + the rank's segment
    decoded from the (memory-mapped) container by ndzb_container_decompress_segment."""
    import tempfile
    import torch
    itemsize = np.dtype(dtype).itemsize
    need = int(world * n_words * itemsize * 1.2) + (1 << 20)
    box = [None]
    if rank == 0:
        for cand in ("/dev/shm", tempfile.gettempdir()):
            try:
                st = os.statvfs(cand)
                if st.f_bavail * st.f_frsize > 2 * need and os.access(cand, os.W_OK):
                    box[0] = os.path.join(cand, f"ndzb_bench_{os.getpid()}.ndzs")
                    break
            except OSError:
                pass
    dist.broadcast_object_list(box, src=0)
    path = box[0]
    if path is None:
        return {"skipped": "no writable directory with %d MB free" % (2 * need >> 20)}

    def wall_max(t):
        v = torch.tensor([t], dtype=torch.float64, device=dev)
        dist.all_reduce(v, op=dist.ReduceOp.MAX)
        return float(v.item())

    h_stream = torch.empty(n_words, dtype=tbits, pin_memory=True)   # staging buffers and the offloader: not timed
    h_back = torch.empty(slab_shape, dtype=d_in.dtype, pin_memory=True)
    off = nz.make_cuda_offloader(dtype, len(slab_shape))
    torch.cuda.synchronize()
    dist.barrier()
    t0 = time.perf_counter()
    h_stream.copy_(d_stream[:n_words])
    hdr = nzd.write_sharded(path, dtype, global_shape, h_stream.numpy())  # collective; fails on every rank or on none
    write_s = wall_max(time.perf_counter() - t0)
    ok, same, err = False, None, None
    t0 = time.perf_counter()
    try:  # rank-local from here to the next collective: a failure is reported, not raised
        blob = np.memmap(path, dtype=np.uint8, mode="r")
        nzd.decompress_segment(off, blob, rank, h_back)
        read_local = time.perf_counter() - t0
        ok = bool(torch.equal(h_back.view(tbits), d_in.cpu().view(tbits)))
        if rank == 0:
            stitched = torch.from_numpy(nzd.to_global_stream(blob).view(np.int32 if itemsize == 4 else np.int64))
            same = bool(stitched.numel() == total_words and torch.equal(stitched.to(dev), d_global[:total_words]))
            del stitched
        del blob
    except Exception as exc:
        err, read_local = repr(exc)[:200], time.perf_counter() - t0
    read_s = wall_max(read_local)
    flags = torch.tensor([1 if ok else 0, 1 if (same or rank != 0) else 0], device=dev)
    dist.all_reduce(flags, op=dist.ReduceOp.MIN)
    if rank == 0 and os.path.exists(path):
        os.remove(path)
    nbytes = int(np.prod(global_shape)) * itemsize
    out = {"what": "every rank writes / reads / decodes its own segment of ONE container file; no offset exchange, no gather "
                   "(ndzb_container_create_file, _write_segment, _decompress_segment, _to_global_stream)",
           "file": os.path.dirname(path), "container_bytes": int(hdr.total_bytes), "segments": len(hdr.segments),
           "write_ms": write_s * 1e3, "write_gbs": nbytes / write_s / 1e9,
           "read_decompress_ms": read_s * 1e3, "read_decompress_gbs": nbytes / read_s / 1e9,
           "segments_round_trip": bool(flags[0].item()), "single_stream_from_container_identical_to_gathered": bool(flags[1].item())}
    if err:
        out["error"] = err
    return out


def run_dist_config(name, world, rank, dev, steps=5, identity=True):
    """One multi-GPU BASELINE config through the library's data plane (ndzb_dist_*): every rank owns one slab of
    WORKLOADS[name] x world. Timed: compress + exchange + decompress, the same plus the final NCCL gather into a
    pre-allocated buffer on rank 0, and the compress launch alone; checked: round trip on every rank and the gathered
    stream against the stream ONE GPU produces for the whole grid (rank 0 receives every slab for that)."""
    import torch
    import torch.distributed as dist
    import ndzip_b200 as nz
    from ndzip_b200 import dist as nzd

    dtype, shape, desc = WORKLOADS[name]
    itemsize = np.dtype(dtype).itemsize
    tbits = torch.int32 if itemsize == 4 else torch.int64
    global_shape = (shape[0] * world,) + tuple(shape[1:])
    nbytes_rank = int(np.prod(shape)) * itemsize
    d_in = make_device_input(dtype, shape, seed=SEED, device=dev, index_offset=rank * shape[0])
    codec = nzd.DistCodec(dtype, global_shape)
    assert codec.slab_shape == tuple(shape), (codec.slab_shape, shape)
    L = codec.layout
    d_stream = torch.empty(int(L.local_bound_words), dtype=tbits, device=dev)
    d_len = torch.zeros(1, dtype=torch.int32, device=dev)
    d_back = torch.empty_like(d_in)
    d_global = torch.empty(int(L.global_bound_words), dtype=tbits, device=dev) if rank == 0 else None

    def sync_all():
        dist.barrier()
        torch.cuda.synchronize()

    def step():
        codec.compress(d_in, d_stream, d_len)
        codec.decompress(d_stream, d_back)
        codec.wait_exchange()

    def step_gather():
        codec.compress(d_in, d_stream, d_len)
        return codec.gather(d_stream, d_global, root=0)

    def timed(fn, reps):
        """Median over `reps` individually timed calls on every rank (device events), then the max over ranks; the
        K-call block time (max over ranks) is returned beside it. (A fresh NCCL communicator's first collectives and the
        tear-down of the previous one show up as outliers in a plain block average: 1.87 instead of 0.57 ms per step
        for cfg5 on 8 GPUs when it ran right after cfg4.)"""
        sync_all()
        marks = [torch.cuda.Event(enable_timing=True) for _ in range(reps + 1)]
        marks[0].record()
        for i in range(reps):
            fn()
            marks[i + 1].record()
        sync_all()
        each = sorted(marks[i].elapsed_time(marks[i + 1]) for i in range(reps))
        t = torch.tensor([each[len(each) // 2], marks[0].elapsed_time(marks[reps]) / reps], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    for _ in range(5):
        step()
    total_words = step_gather()  # warm-up: NCCL point-to-point connections
    for _ in range(3):
        step()
    sync_all()
    roundtrip = bool(torch.equal(d_in.view(tbits), d_back.view(tbits)))
    ms_step, ms_step_block = timed(step, max(steps, 9))
    ms_gather, _ = timed(step_gather, max(3, steps // 2))
    tc = time_call(lambda: codec.compress(d_in, d_stream, d_len), steps)
    codec.wait_exchange()
    torch.cuda.synchronize()
    n_words = int(d_len.cpu().numpy().view(np.uint32)[0])
    algo = nbytes_rank + n_words * itemsize
    frac = torch.tensor([algo / (sum(tc) / len(tc) * 1e-3) / 1e9], dtype=torch.float64, device=dev)
    frac_min = frac.clone()
    dist.all_reduce(frac_min, op=dist.ReduceOp.MIN)
    dist.all_reduce(frac, op=dist.ReduceOp.SUM)

    identical = None
    if identity:
        step_gather()
        slabs = [torch.empty_like(d_in) for _ in range(world)] if rank == 0 else None
        dist.gather(d_in, slabs, dst=0)
        if rank == 0:
            whole = torch.cat(slabs, dim=0)
            del slabs
            comp = nz.make_cuda_compressor(dtype, nz.compressor_requirements(global_shape))
            ref = torch.empty(nz.compressed_length_bound(dtype, global_shape), dtype=tbits, device=dev)
            ref_len = torch.zeros(1, dtype=torch.int32, device=dev)
            comp.compress(whole, global_shape, ref, ref_len)
            torch.cuda.synchronize()
            n_ref = int(ref_len.cpu().numpy().view(np.uint32)[0])
            identical = bool(n_ref == total_words and torch.equal(ref[:n_ref], d_global[:n_ref]))
            del whole, ref, comp
        flag = torch.tensor([1 if (identical or rank != 0) else 0], device=dev)
        dist.broadcast(flag, src=0)
        identical = bool(flag.item())
    ok = torch.tensor([1 if roundtrip else 0], device=dev)
    dist.all_reduce(ok, op=dist.ReduceOp.MIN)
    peak, _ = measured_hbm_peak()
    out = {
        "workload": f"{name}: {desc} x {world} ranks = {global_shape}", "per_rank_bytes": nbytes_rank,
        "value": nbytes_rank * world / (ms_step * 1e-3) / 1e9, "unit": "GB/s", "ms_per_step": ms_step,
        "ms_per_step_block_average": ms_step_block,
        "timing": "median of individually event-timed steps per rank, max over ranks (block average beside it)",
        "with_gather_gbs": nbytes_rank * world / (ms_gather * 1e-3) / 1e9, "with_gather_ms": ms_gather,
        "global_stream_bytes": int(total_words) * itemsize, "ratio": n_words * itemsize / nbytes_rank,
        "roofline": {"kernel": "compress_ws_kernel", "achieved_per_gpu_avg": float(frac.item()) / world, "peak": peak,
                     "frac": float(frac.item()) / world / peak, "frac_min_over_ranks": float(frac_min.item()) / peak},
        "global_stream_identical": identical, "roundtrip_ok": bool(ok.item()), "gather_path": codec.last_gather_path,
    }
    codec.close()
    del d_in, d_stream, d_back, d_global
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="N>1: skip the cfg4 / cfg5 records")
    ap.add_argument("--no-sustained", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dtype, shape, desc = WORKLOADS[args.workload]
    itemsize = np.dtype(dtype).itemsize
    nbytes_rank = int(np.prod(shape)) * itemsize
    unit = "GB/s"
    metric = "uncompressed GB/s (compress+decompress round trip)"

    # ------------------------------------------------------------------ reference arm (CPU, rank 0)
    if args.impl == "reference":
        if rank != 0:
            return 0
        # one rank's grid of the workload (for N = 1 that IS the config; for N > 1 a bounded sample of it: one slab)
        data = host_input(dtype, shape)
        r = run_cpu_reference(dtype, shape, data, args.steps, args.warmup, threads=0)
        sample = f"the whole {shape} {dtype} grid" if args.gpus == 1 else f"one rank's {shape} {dtype} slab of the {args.gpus}-slab grid"
        line = {
            "impl": "reference", "metric": metric, "value": r["value"], "unit": unit, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"], "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32" if itemsize == 4 else "u64", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {desc}", "sample": sample},
            "compress_gbs": r["compress_gbs"], "decompress_gbs": r["decompress_gbs"], "ratio": r["ratio"],
            "roundtrip_ok": r["roundtrip_ok"],
            "cpu_baseline": {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": r["kind"],
                             "sample": f"{sample}, reference OpenMP codec on all host threads, median of {r['steps']}",
                             "roundtrip_ok": r["roundtrip_ok"]},
            "e2e": {"value": r["value"], "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0,
        }
        print(json.dumps(line))
        return 0

    # ------------------------------------------------------------------ our arm
    import torch
    import torch.distributed as dist
    import ndzip_b200 as nz
    from ndzip_b200 import dist as nzd

    torch.cuda.set_device(local_rank)
    # host placement: this rank's thread and the pinned buffers it allocates go to the NUMA node of its GPU
    # (ndzb_bind_host_to_device; -1 = the box exposes no NUMA topology, nothing changed)
    affinity_before = os.sched_getaffinity(0)
    numa_node = nz.bind_host_to_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)
    tbits = torch.int32 if itemsize == 4 else torch.int64

    global_shape = (shape[0] * world,) + tuple(shape[1:])
    d_in = make_device_input(dtype, shape, seed=SEED, device=dev, index_offset=rank * shape[0])
    bound = nz.compressed_length_bound(dtype, shape)
    d_stream = torch.empty(bound, dtype=tbits, device=dev)
    d_len = torch.zeros(1, dtype=torch.int32, device=dev)
    d_back = torch.empty_like(d_in)
    launches = 0

    if world > 1:
        # the library's multi-GPU data plane: slab compression, NCCL count exchange on a side stream (it overlaps the
        # decompression, which does not depend on it), header fix-up kernel; no Python in the exchange
        codec = nzd.DistCodec(dtype, global_shape)
        assert codec.slab_shape == tuple(shape)

        def step():
            codec.compress(d_in, d_stream, d_len)   # compress_ws_kernel + (side stream) ncclAllGather + fixup_header_kernel
            codec.decompress(d_stream, d_back)      # decompress_kernel
            codec.wait_exchange()
            return 3                                # our kernels per step (the all-gather kernel is NCCL's)

        def compress_only():
            codec.compress(d_in, d_stream, d_len)

        def decompress_only():
            codec.decompress(d_stream, d_back)
    else:
        comp = nz.make_cuda_compressor(dtype, nz.compressor_requirements(shape))
        dec = nz.make_cuda_decompressor(dtype, len(shape))

        def step():
            comp.compress(d_in, shape, d_stream, d_len)
            n = comp.last_launch_count
            dec.decompress(d_stream, d_back, shape)
            return n + dec.last_launch_count

        def compress_only():
            comp.compress(d_in, shape, d_stream, d_len)

        def decompress_only():
            dec.decompress(d_stream, d_back, shape)

    def sync_all():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    sync_all()
    n_words = int(d_len.cpu().numpy().view(np.uint32)[0])
    stream_bytes = n_words * itemsize
    assert torch.equal(d_in.view(tbits), d_back.view(tbits)), "round trip mismatch"

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sync_all()
    ev0.record()
    for _ in range(args.steps):
        launches += step()
    ev1.record()
    sync_all()
    total_ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms = float(t.item())
    ms_per_step = total_ms / args.steps

    # ---- per-kernel timings (same stream, CUDA events around single launches), rank-local
    reps = max(5, min(args.steps, 20))
    tc = time_call(compress_only, reps)
    td = time_call(decompress_only, reps)
    clocks = sampler.stop() if rank == 0 else None
    tc_avg, td_avg = sum(tc) / len(tc), sum(td) / len(td)

    # ---- sustained: the same step back to back for at least a second (the reference benchmark's protocol is >= 1 s of
    # repetitions, src/benchmark/benchmark.cc:197-227), clocks and power sampled: the 20-step figure above is a burst
    sustained = None
    if not args.no_sustained:
        n_sus = max(args.steps, int(1200.0 / max(ms_per_step, 1e-3)))
        s2 = ClockSampler(local_rank, period_s=0.01)
        if rank == 0:
            s2.start()
        sync_all()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n_sus):
            step()
        b.record()
        sync_all()
        sus_ms = a.elapsed_time(b)
        if world > 1:
            t = torch.tensor([sus_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sus_ms = float(t.item())
        c2 = s2.stop() if rank == 0 else None
        sustained = {"steps": n_sus, "seconds": sus_ms * 1e-3, "value": nbytes_rank * world * n_sus / (sus_ms * 1e-3) / 1e9,
                     "unit": unit, "ms_per_step": sus_ms / n_sus, "clocks": c2}

    # ---- the reference's own CUDA kernels (recompiled for sm_100a) on the same buffers, rank 0 only
    ref_cuda = None
    if rank == 0 and not args.no_cpu_baseline:
        try:
            from oracle import ReferenceCuda
            if ReferenceCuda.available():
                rc = ReferenceCuda(dtype, shape)
                r_stream = torch.empty(bound, dtype=tbits, device=dev)
                r_len = torch.zeros(1, dtype=torch.int32, device=dev)
                r_back = torch.empty_like(d_in)
                rtc = time_call(lambda: rc.compress(d_in.data_ptr(), r_stream.data_ptr(), r_len.data_ptr()), 5)
                rtd = time_call(lambda: rc.decompress(r_stream.data_ptr(), r_back.data_ptr()), 5)
                n_ref = int(r_len.cpu().numpy().view(np.uint32)[0])
                identical = bool(n_ref == n_words and torch.equal(r_stream[:n_ref], d_stream[:n_words])
                                 and torch.equal(r_back.view(tbits), d_in.view(tbits)))
                ref_cuda = {"what": "reference cuda_compressor/cuda_decompressor (src/ndzip/cuda_codec.inl) recompiled for sm_100a, same buffers",
                            "compress_ms": min(rtc), "decompress_ms": min(rtd),
                            "compress_gbs": nbytes_rank / (min(rtc) * 1e-3) / 1e9, "decompress_gbs": nbytes_rank / (min(rtd) * 1e-3) / 1e9,
                            "stream_identical_to_ours": identical}
                del rc, r_stream, r_back
        except Exception as exc:  # the baseline is optional; never fail the bench because of it
            ref_cuda = {"error": repr(exc)[:200]}

    # ---- N > 1: final stream gather of the headline grid, timed and checked against ONE GPU compressing the whole grid
    headline_gather = None
    sharded = None
    if world > 1:
        L = codec.layout
        d_global = torch.empty(int(L.global_bound_words), dtype=tbits, device=dev) if rank == 0 else None
        codec.compress(d_in, d_stream, d_len)
        total_words = codec.gather(d_stream, d_global, root=0)  # warm-up: NCCL point-to-point connections
        times = []
        for _ in range(3):
            sync_all()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            codec.compress(d_in, d_stream, d_len)
            codec.gather(d_stream, d_global, root=0)
            g1.record()
            sync_all()
            gt = torch.tensor([g0.elapsed_time(g1)], dtype=torch.float64, device=dev)
            dist.all_reduce(gt, op=dist.ReduceOp.MAX)
            times.append(float(gt.item()))
        slabs = [torch.empty_like(d_in) for _ in range(world)] if rank == 0 else None
        dist.gather(d_in, slabs, dst=0)
        identical = None
        if rank == 0:
            whole = torch.cat(slabs, dim=0)
            del slabs
            c1 = nz.make_cuda_compressor(dtype, nz.compressor_requirements(global_shape))
            ref = torch.empty(nz.compressed_length_bound(dtype, global_shape), dtype=tbits, device=dev)
            ref_len = torch.zeros(1, dtype=torch.int32, device=dev)
            c1.compress(whole, global_shape, ref, ref_len)
            torch.cuda.synchronize()
            n_ref = int(ref_len.cpu().numpy().view(np.uint32)[0])
            identical = bool(n_ref == total_words and torch.equal(ref[:n_ref], d_global[:n_ref]))
            del whole, ref, c1
        headline_gather = {"compress_exchange_gather_ms": min(times), "global_stream_bytes": int(total_words) * itemsize,
                           "with_gather_gbs": nbytes_rank * world / (min(times) * 1e-3) / 1e9,
                           "global_stream_identical": identical, "gather_path": codec.last_gather_path,
                           "note": "ndzb_dist_compress + ndzb_dist_gather into a pre-allocated buffer on rank 0; "
                                   "bounded by one GPU's NVLink ingest"}
        # ---- the same grid without any gather: the sharded container (ndzb_container_*). Every rank
        # writes its own slab stream into one file, reads it back and decodes it from the container; rank 0 converts
        # the container to the single stream and compares it with the gathered one.
        try:
            sharded = sharded_container_record(nz, nzd, dist, dtype, global_shape, shape, rank, world, dev, tbits,
                                               d_in, d_stream, n_words, d_global, int(total_words))
        except Exception as exc:  # optional record: never lose the headline line
            sharded = {"error": repr(exc)[:300]}
        del d_global
        torch.cuda.empty_cache()

    # ---- e2e: host-pointer offloader API, pinned buffers, copies inside the timed region (rank-local)
    e2e = None
    if not args.no_e2e:
        h_in = torch.empty(shape, dtype=d_in.dtype, pin_memory=True)
        h_in.copy_(d_in)
        h_stream = torch.empty(bound, dtype=tbits, pin_memory=True)
        h_back = torch.empty(shape, dtype=d_in.dtype, pin_memory=True)
        off = nz.make_cuda_offloader(dtype, len(shape))
        for _ in range(2):
            n_off = off.compress(h_in, shape, h_stream)
            off.decompress(h_stream, n_off, h_back, shape)
        sync_all()
        e2e_steps = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(e2e_steps):
            n_off = off.compress(h_in, shape, h_stream)
            off.decompress(h_stream, n_off, h_back, shape)
        e1.record()
        torch.cuda.synchronize()
        e2e_ms = e0.elapsed_time(e1) / e2e_steps
        wall_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
        e2e_ms = max(e2e_ms, wall_ms)  # host-synchronous API: wall clock is the honest figure
        assert n_off == n_words and torch.equal(h_in.view(tbits), h_back.view(tbits))
        h2d, d2h = int(nbytes_rank + stream_bytes), int(stream_bytes + nbytes_rank + 4)
        # the box's ceiling for these byte counts: compress call = H2D(input) || D2H(stream), decompress call =
        # H2D(stream) || D2H(output); every rank measures at the same time, like the e2e loop itself
        sync_all()
        ceil_ms = pcie_ceiling(nbytes_rank, stream_bytes, dev) + pcie_ceiling(stream_bytes, nbytes_rank, dev)
        if world > 1:
            t = torch.tensor([e2e_ms, ceil_ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            e2e_ms, ceil_ms = float(t[0].item()), float(t[1].item())
        e2e = {"value": nbytes_rank * world / (e2e_ms * 1e-3) / 1e9, "unit": unit,
               "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
               "ms_per_step": e2e_ms, "api": "make_cuda_offloader(...).compress/.decompress (pinned host buffers)",
               "pcie_ceiling_gbs": nbytes_rank * world / (ceil_ms * 1e-3) / 1e9, "pcie_ceiling_ms": ceil_ms,
               "frac_of_ceiling": ceil_ms / e2e_ms, "host_numa_node": numa_node,
               "pcie_ceiling_how": "pinned cudaMemcpyAsync of the same byte counts, H2D and D2H on two streams at once, "
                                   "compress-call pair + decompress-call pair, all ranks simultaneously"}
        del h_in, h_stream, h_back, off

    # ---- N > 1: the multi-GPU BASELINE configs (configs[3], configs[4]) outside the timed cfg2 loop
    configs = None
    if world > 1 and not args.no_configs:
        del d_stream, d_back
        torch.cuda.empty_cache()
        configs = {}
        for name in ("cfg4", "cfg5"):
            try:
                configs[name] = run_dist_config(name, world, rank, dev)
            except Exception as exc:  # report, do not lose the headline line
                configs[name] = {"error": repr(exc)[:300]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = measured_hbm_peak()
    algo_bytes = nbytes_rank + stream_bytes
    achieved = algo_bytes / (tc_avg * 1e-3) / 1e9
    dec_achieved = algo_bytes / (td_avg * 1e-3) / 1e9
    line = {
        "metric": metric, "value": nbytes_rank * world / (ms_per_step * 1e-3) / 1e9, "unit": unit,
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "u32" if itemsize == 4 else "u64", "data": "synthetic",
        "config": {
            "workload": f"{args.workload}: {desc}" + (f" x {world} ranks, slabs of {global_shape}" if world > 1 else ""),
            "per_rank_bytes": nbytes_rank, "ratio": stream_bytes / nbytes_rank,
            "l2": l2_note(nbytes_rank),
            "exchange": ("ndzb_dist_compress: ncclAllGather of the per-rank stream lengths on a high-priority side stream "
                         "(overlaps the decompression) + header fix-up kernel, all inside libndzip_b200.so") if world > 1 else "none (1 GPU)",
        },
        "compress_gbs": nbytes_rank / (tc_avg * 1e-3) / 1e9, "decompress_gbs": nbytes_rank / (td_avg * 1e-3) / 1e9,
        "compress_ms": tc_avg, "decompress_ms": td_avg, "compress_ms_min": min(tc), "decompress_ms_min": min(td),
        "roofline": {"bound": "hbm", "kernel": "compress_ws_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak, "peak_source": peak_src, "algorithmic_bytes": algo_bytes,
                     "traffic": ncu_traffic_per_launch(args.workload),
                     "read_only_frac": nbytes_rank / (tc_avg * 1e-3) / 1e9 / peak,
                     "decompress_kernel": {"achieved": dec_achieved, "frac": dec_achieved / peak}},
        "clocks": clocks, "gpu_launches": launches,
    }
    if sustained:
        line["sustained"] = sustained
    if e2e:
        line["e2e"] = e2e
    if headline_gather:
        line["stream_gather"] = headline_gather
    if sharded:
        line["sharded_container"] = sharded
    if configs:
        line["configs"] = configs
    if ref_cuda:
        line["reference_cuda"] = ref_cuda
    if not args.no_cpu_baseline:
        host = d_in.cpu().numpy()
        os.sched_setaffinity(0, affinity_before)  # the CPU baseline gets every host core again, not just the GPU's NUMA node
        r = run_cpu_reference(dtype, shape, host, steps=3, warmup=1, threads=0)
        r1 = run_cpu_reference(dtype, shape, host, steps=1, warmup=0, threads=1)
        line["cpu_baseline"] = {"value": r["value"], "unit": unit, "cores": r["cores"], "kind": r["kind"],
                                "compress_gbs": r["compress_gbs"], "decompress_gbs": r["decompress_gbs"],
                                "roundtrip_ok": r["roundtrip_ok"],
                                "sample": f"the same {shape} {dtype} buffer (whole grid of this rank), reference OpenMP codec (-e cpu-mt), median of {r['steps']}",
                                "single_thread": {"value": r1["value"], "cores": 1, "compress_gbs": r1["compress_gbs"],
                                                  "decompress_gbs": r1["decompress_gbs"], "roundtrip_ok": r1["roundtrip_ok"],
                                                  "what": "reference serial codec (-e cpu, make_cpu_offloader(dims, 1)), same buffer, 1 repetition"}}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
