#!/usr/bin/env python3
"""Benchmark of the scan hot path (BASELINE.json metric: scanned GB/s; configs[1] = 32 mixed patterns over 10 GiB).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--gib G]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One step = one pass of the hot path over one batch of synthetic syslog text (per rank: --gib GiB, default 10).
  value  : device-resident whole-job throughput, CUDA events on the stream the kernels run on, max over ranks.
  e2e    : same pass through the C ABI from PINNED HOST memory, H2D of the input and D2H of the records inside the
           timed region, full delivery path (matched lines copied into result slots, native discard callback).
  roofline: the streaming kernel (k_stream) against the measured HBM copy peak (MEASURED_PEAKS.json).
  cpu_baseline / --impl reference: the oracle port (reference loop shape, PCRE2-JIT matcher; the reference itself
           needs Hyperscan, which is not installable here) on the box's host cores, bounded sample.
Multi-GPU: ranks scan independent newline-aligned shards (rank-specific seed), no collective on the data path.
"""

from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

WORKLOAD = "configs[1]: 32 mixed literal + character-class patterns, synthetic syslog text"


def parse_args() -> argparse.Namespace:
    parser = argparse.ArgumentParser()
    parser.add_argument("--gpus", type=int, default=1)
    parser.add_argument("--steps", type=int, default=5)
    parser.add_argument("--warmup", type=int, default=3)
    parser.add_argument("--impl", default="ours", choices=["ours", "reference"])
    parser.add_argument("--gib", type=float, default=float(os.environ.get("GPUGREP_BENCH_GIB", "10")))
    parser.add_argument("--cpu-seconds", type=float, default=15.0, help="target CPU time of the cpu_baseline sample")
    return parser.parse_args()


def load_peaks() -> tuple[float, str]:
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path, encoding="utf-8") as handle:
            return float(json.load(handle)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except (OSError, KeyError, ValueError):
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons while the timed region runs."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int) -> None:
        self.device = device
        self.proc = None
        self.lines: list[tuple[float, str]] = []
        self.windows: list[tuple[float, float]] = []   # timed regions (perf_counter), samples outside them are dropped

    def wait_first(self, timeout: float = 10.0) -> None:
        """nvidia-smi takes a second or more to print its first sample; the timed regions are shorter than that."""
        deadline = time.perf_counter() + timeout
        while self.proc is not None and not self.lines and time.perf_counter() < deadline:
            time.sleep(0.02)

    def start(self) -> None:
        try:
            self.proc = subprocess.Popen(  # pylint: disable=consider-using-with
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.device)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._drain, daemon=True).start()
        except OSError:
            self.proc = None

    def _drain(self) -> None:
        for line in self.proc.stdout:
            self.lines.append((time.perf_counter(), line.strip()))

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for when, line in self.lines:
            if self.windows and not any(lo <= when <= hi for lo, hi in self.windows):
                continue
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                mx.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[5:9]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def oracle_throughput(data_ptr: int, size: int, patterns, threads: int, seconds: float):
    """Oracle port (reference loop + PCRE2-JIT) over newline-aligned shards of a bounded sample, `threads` at once.

    Returns (GB/s, bytes, matches, wall seconds).  The sample grows until it costs about `seconds` of wall time.
    """
    from gpu_api import marshal  # pylint: disable=import-outside-toplevel
    from oracle_api import load_oracle  # pylint: disable=import-outside-toplevel

    oracle = load_oracle()
    oracle.oracle_count_buffer.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                           ctypes.c_uint, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]
    pa, fa, ia, n = marshal(patterns)
    view = (ctypes.c_ubyte * size).from_address(data_ptr)

    def run(sample: int):
        bounds = [0]
        for t in range(1, threads):
            pos = sample * t // threads
            while pos < sample and view[pos - 1] != 10:
                pos += 1
            bounds.append(pos)
        bounds.append(sample)
        matches = [0] * threads

        def work(t: int) -> None:
            m, ln = ctypes.c_ulonglong(0), ctypes.c_ulonglong(0)
            lo, hi = bounds[t], bounds[t + 1]
            rc = oracle.oracle_count_buffer(data_ptr + lo, hi - lo, pa, fa, ia, n, 262140, ctypes.byref(m), ctypes.byref(ln))
            assert rc == 0
            matches[t] = m.value

        t0 = time.perf_counter()
        pool = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
        for th in pool:
            th.start()
        for th in pool:
            th.join()
        return time.perf_counter() - t0, sum(matches)

    sample = min(size, (8 << 20) * threads)
    while sample > 1 and view[sample - 1] != 10:
        sample -= 1
    elapsed, matches = run(sample)
    if elapsed < seconds / 3 and sample < size:
        scaled = min(size, int(sample * seconds / max(elapsed, 1e-3)))
        while scaled > 1 and view[scaled - 1] != 10:
            scaled -= 1
        sample = scaled
        elapsed, matches = run(sample)
    return sample / elapsed / 1e9, sample, matches, elapsed


def main() -> None:  # pylint: disable=too-many-locals,too-many-statements
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import numpy as np  # pylint: disable=import-outside-toplevel

    from hypergrep_b200 import synth  # pylint: disable=import-outside-toplevel

    patterns = synth.C2_PATTERNS
    config = {"workload": WORKLOAD, "patterns": len(patterns), "gib_per_gpu": args.gib, "line_bytes_mean": 151,
              "l2": "inputs (GiBs) are far larger than the 126 MB L2, no flush needed", "parallelism": f"shard{world}",
              "buffer_size": 262140}

    # ------------------------------------------------------------------ reference arm (CPU, rank 0 only)
    if args.impl == "reference":
        if rank != 0:
            return
        lib = ctypes.CDLL(os.path.join(ROOT, "hypergrep_b200", "lib", "libgpugrep.so"))
        threads = os.cpu_count() or 1
        size = min(int(args.gib * (1 << 30)), (64 << 20) * threads)
        text = np.empty(size, dtype=np.uint8)
        synth.fill_syslog(text, seed=1234, lib=lib)
        rates = []
        sample = matches = 0
        per_step = max(2.0, min(args.cpu_seconds, 120.0 / max(1, args.steps + args.warmup)))
        for step in range(args.warmup + args.steps):
            gbs, sample, matches, _ = oracle_throughput(text.ctypes.data, size, patterns, threads, per_step)
            if step >= args.warmup:
                rates.append(gbs)
        value = sum(rates) / len(rates)
        line = {
            "impl": "reference", "metric": "scanned GB/s", "value": value, "unit": "GB/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sample / value / 1e6, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": value, "unit": "GB/s", "cores": threads, "kind": "port",
                             "sample": f"{sample / (1 << 20):.0f} MiB of the same text per step, {threads} newline-aligned shards in parallel; "
                                       "oracle port = reference loop (gzgets split, per-line match, strcpy) with PCRE2-JIT standing in for Hyperscan",
                             "matched_lines_per_s": matches / (sample / value / 1e9)},
            "e2e": {"value": value, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        }
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm
    import torch  # pylint: disable=import-outside-toplevel
    import torch.distributed as dist  # pylint: disable=import-outside-toplevel

    from gpu_api import Stats, marshal  # pylint: disable=import-outside-toplevel
    from hypergrep_b200 import utils  # pylint: disable=import-outside-toplevel

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = utils._get_hyperscanner_lib()  # pylint: disable=protected-access
    lib.gpugrep_set_device(local_rank)
    lib.gpugrep_scan_buffer.argtypes = [
        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint,
        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p,
    ]
    discard = ctypes.cast(lib.gpugrep_discard_results, ctypes.c_void_p)
    pa, fa, ia, npat = marshal(patterns)

    size = int(args.gib * (1 << 30))
    host = torch.empty(size, dtype=torch.uint8).pin_memory()
    t0 = time.perf_counter()
    lines = synth.fill_syslog(host.numpy(), seed=1234 + 1000 * rank, lib=lib)
    gen_s = time.perf_counter() - t0
    dev = host.cuda(non_blocking=False)
    torch.cuda.synchronize()
    # a dedicated (non-default) stream: the library enqueues every kernel of the device-resident passes on it, so the
    # CUDA events below bracket exactly the work that is timed
    stream = torch.cuda.Stream()

    def scan(ptr: int, location: int, callback) -> Stats:
        st = Stats()
        rc = lib.gpugrep_scan_buffer(ptr, size, location, pa, fa, ia, npat, callback, 262140, 4096, 0,
                                     ctypes.c_void_p(stream.cuda_stream) if location == 1 else None, ctypes.byref(st))
        if rc != 0:
            raise RuntimeError(f"gpugrep_scan_buffer failed with code {rc}: {lib.gpugrep_last_error()}")
        return st

    def barrier() -> None:
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(value: float) -> float:
        if world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(value: float) -> float:
        if world == 1:
            return value
        t = torch.tensor([value], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # ---- device-resident: value + roofline
    sampler = ClockSampler(local_rank)
    sampler.start()
    sampler.wait_first()
    for _ in range(args.warmup):
        scan(dev.data_ptr(), 1, None)
    barrier()
    window_begin = time.perf_counter()
    begin, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches = stream_launches = 0
    stream_ms = gpu_ms = 0.0
    matches = 0
    stream.wait_stream(torch.cuda.current_stream())
    begin.record(stream)
    for _ in range(args.steps):
        st = scan(dev.data_ptr(), 1, None)
        launches += st.launches
        stream_launches += st.stream_launches
        stream_ms += st.stream_kernel_ms
        gpu_ms += st.gpu_ms
        matches = st.matches
    end.record(stream)
    barrier()
    sampler.windows.append((window_begin, time.perf_counter()))
    dev_ms = max_over_ranks(begin.elapsed_time(end))
    total_bytes = sum_over_ranks(float(size))
    value = total_bytes * args.steps / (dev_ms / 1e3) / 1e9
    total_matches = sum_over_ranks(float(matches))

    # ---- end to end from pinned host memory through the C ABI (H2D + kernels + D2H + delivery)
    for _ in range(max(1, args.warmup // 2)):
        scan(host.data_ptr(), 0, discard)
    barrier()
    t0 = time.perf_counter()
    h2d = d2h = 0
    e2e_launches = 0
    for _ in range(args.steps):
        st = scan(host.data_ptr(), 0, discard)
        h2d, d2h = st.h2d_bytes, st.d2h_bytes
        e2e_launches += st.launches
    barrier()
    sampler.windows.append((t0, time.perf_counter()))
    clocks = sampler.stop()   # samples of both timed regions (device-resident passes and end-to-end passes)
    e2e_s = max_over_ranks(sampler.windows[-1][1] - t0)
    e2e_value = total_bytes * args.steps / e2e_s / 1e9

    # ---- the reference-facing call itself: hyperscan(path) on a file in tmpfs (read() into pinned memory + H2D + ...)
    e2e_file = None
    if world == 1:
        try:
            file_bytes = min(size, 4 << 30)
            while file_bytes > 1 and host[file_bytes - 1].item() != 10:
                file_bytes -= 1
            shm = "/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"
            path = os.path.join(shm, f"gpugrep_bench_{os.getpid()}.log")
            host.numpy()[:file_bytes].tofile(path)
            lib.gpugrep_scan_file.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p,
                                              ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p]
            fst = Stats()
            for it in range(3):
                t0 = time.perf_counter()
                rc = lib.gpugrep_scan_file(path.encode(), pa, fa, ia, npat, discard, 262140, 4096, 0, ctypes.byref(fst))
                file_s = time.perf_counter() - t0
                assert rc == 0
            os.unlink(path)
            e2e_file = {"value": file_bytes / file_s / 1e9, "unit": "GB/s", "bytes": file_bytes, "matches": int(fst.matches),
                        "source": "hyperscan(path)-equivalent gpugrep_scan_file on a tmpfs file, native discard callback, third of 3 runs"}
        except Exception as error:  # pylint: disable=broad-except
            e2e_file = {"value": None, "error": str(error)}

    peak, peak_source = load_peaks()
    kernel_bytes = size / max(1, stream_launches // max(1, args.steps))   # algorithmic bytes per k_stream launch
    avg_launch_ms = stream_ms / max(1, stream_launches)
    achieved = kernel_bytes / (avg_launch_ms / 1e3) / 1e9 if avg_launch_ms > 0 else 0.0

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # DRAM traffic of the roofline kernel from the committed ncu capture (profiles/), scaled to this run's launch size
    traffic, traffic_source = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "r1_k_stream_traffic.json"), encoding="utf-8") as handle:
            cap = json.load(handle)
        traffic = (cap["dram_bytes_read"] + cap["dram_bytes_write"]) * kernel_bytes / cap["launch_bytes"]
        traffic_source = cap["source"]
    except (OSError, KeyError, ValueError):
        pass

    line = {
        "metric": "scanned GB/s", "value": value, "unit": "GB/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
        "data": "synthetic", "config": config,
        "matched_lines_per_s": total_matches * args.steps / (dev_ms / 1e3),
        "matched_lines_per_step": total_matches, "lines_per_step": sum_over_ranks(float(lines)) if world == 1 else None,
        "e2e": {"value": e2e_value, "unit": "GB/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": e2e_s * 1e3 / args.steps, "source": "pinned host memory -> gpugrep_scan_buffer (C ABI) -> native discard callback"},
        "e2e_file": e2e_file,
        "gpu_launches": int(launches),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                     "traffic_source": traffic_source,
                     "kernel": "k_stream (newline count + literal prefilter)", "peak_source": peak_source,
                     "algorithmic_bytes_per_launch": kernel_bytes, "avg_launch_ms": avg_launch_ms,
                     "kernel_share_of_gpu_time": stream_ms / gpu_ms if gpu_ms else None,
                     "whole_pipeline_frac": (size * args.steps / (gpu_ms / 1e3) / 1e9) / peak if gpu_ms else None},
        "clocks": clocks,
        "corpus_generation_s": gen_s,
    }
    if world == 1:
        try:
            threads = 1
            gbs, sample, cpu_matches, secs = oracle_throughput(host.data_ptr(), size, patterns, threads, args.cpu_seconds)
            line["cpu_baseline"] = {
                "value": gbs, "unit": "GB/s", "cores": threads, "kind": "port",
                "sample": f"first {sample / (1 << 20):.0f} MiB of the same text, {secs:.1f} s; oracle port = reference loop shape with "
                          "PCRE2-JIT standing in for Hyperscan (not installable here)",
                "matched_lines_per_s": cpu_matches / secs, "host_cores_available": os.cpu_count()}
        except Exception as error:  # pylint: disable=broad-except
            line["cpu_baseline"] = {"value": None, "unit": "GB/s", "cores": 1, "kind": "port", "sample": f"failed: {error}"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
