#!/usr/bin/env python3
"""BASELINE configs[3] shape: a corpus of gzip / zstd / plain files through multiscanner.parallel_grep (one host thread
per file, files round-robin over the visible GPUs).  Prints throughput in GB/s of DECOMPRESSED text per format."""
import contextlib
import ctypes
import io
import os
import sys
import time
import zlib
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from hypergrep_b200 import multiscanner, synth, utils  # noqa: E402

files_n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
mib_each = int(sys.argv[2]) if len(sys.argv) > 2 else 64
lib = utils._get_hyperscanner_lib()
root = "/dev/shm/gpugrep_c4" if os.path.isdir("/dev/shm") else "/tmp/gpugrep_c4"
os.makedirs(root, exist_ok=True)
zstd = ctypes.CDLL("libzstd.so.1")
zstd.ZSTD_compressBound.restype = ctypes.c_size_t
zstd.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
zstd.ZSTD_compress.restype = ctypes.c_size_t
zstd.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]


def make(index: int) -> tuple[int, int, int]:
    buf = np.zeros(mib_each << 20, dtype=np.uint8)
    synth.fill_syslog(buf, seed=1000 + index, lib=lib)
    raw = bytes(buf[: int(np.flatnonzero(buf == 10)[-1]) + 1])
    with open(f"{root}/f{index}.log", "wb") as handle:
        handle.write(raw)
    gz = zlib.compressobj(6, zlib.DEFLATED, 31)
    packed = gz.compress(raw) + gz.flush()
    with open(f"{root}/f{index}.log.gz", "wb") as handle:
        handle.write(packed)
    bound = zstd.ZSTD_compressBound(len(raw))
    out = ctypes.create_string_buffer(bound)
    size = zstd.ZSTD_compress(out, bound, raw, len(raw), 3)
    with open(f"{root}/f{index}.log.zst", "wb") as handle:
        handle.write(out.raw[:size])
    return len(raw), len(packed), size


t0 = time.time()
with ThreadPoolExecutor(max_workers=os.cpu_count()) as pool:
    sizes = list(pool.map(make, range(files_n)))
total = sum(s[0] for s in sizes)
print(f"corpus: {files_n} files x {mib_each} MiB, gzip ratio {total / sum(s[1] for s in sizes):.1f}, zstd ratio {total / sum(s[2] for s in sizes):.1f}, "
      f"built in {time.time() - t0:.0f} s, {os.cpu_count()} host cores")
for suffix in (".log", ".log.zst", ".log.gz", ".log"):
    names = [f"{root}/f{i}{suffix}" for i in range(files_n)]
    best = 0.0
    count = ""
    rates = []
    for _ in range(3):
        sink = io.StringIO()
        t0 = time.perf_counter()
        with contextlib.redirect_stdout(sink):
            rc = multiscanner.parallel_grep(names, synth.C2_PATTERNS, total_results=True)
        dt = time.perf_counter() - t0
        rates.append(total / dt / 1e9)
        count = sink.getvalue().strip()
    print(f"{suffix:9s} parallel_grep -c total={count} rc={rc}  GB/s of text per run: " + " ".join(f"{r:.2f}" for r in rates))
for name in os.listdir(root):
    os.remove(os.path.join(root, name))
