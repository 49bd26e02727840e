#!/usr/bin/env python3
"""A/B probe for the plain-file read path: one 4 GiB file through gpugrep_scan_file, and eight 64 MiB files at once
through multiscanner.parallel_grep, with the library given in GPUGREP_LIBRARY."""
import contextlib, ctypes, io, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gpu_api import Stats, marshal
from hypergrep_b200 import multiscanner, synth, utils
lib = utils._get_hyperscanner_lib()
root = "/dev/shm/gpugrep_prp"; os.makedirs(root, exist_ok=True)
big = np.empty(4 << 30, dtype=np.uint8); synth.fill_syslog(big, seed=1234)
end = int(np.flatnonzero(big[-4096:] == 10)[-1]) + len(big) - 4096 + 1
big[:end].tofile(f"{root}/big.log")
names = []
for k in range(8):
    lo = k * (64 << 20); hi = lo + (64 << 20)
    while big[hi - 1] != 10: hi -= 1
    names.append(f"{root}/p{k}.log"); big[lo:hi].tofile(names[-1])
pa, fa, ia, n = marshal(synth.C2_PATTERNS)
lib.gpugrep_scan_file.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p]
discard = ctypes.cast(lib.gpugrep_discard_results, ctypes.c_void_p)
single = []
for _ in range(4):
    st = Stats(); t0 = time.perf_counter()
    rc = lib.gpugrep_scan_file(f"{root}/big.log".encode(), pa, fa, ia, n, discard, 262140, 4096, 0, ctypes.byref(st))
    single.append(end / (time.perf_counter() - t0) / 1e9)
multi = []
total = sum(os.path.getsize(p) for p in names)
for _ in range(6):
    sink = io.StringIO(); t0 = time.perf_counter()
    with contextlib.redirect_stdout(sink):
        multiscanner.parallel_grep(names, synth.C2_PATTERNS, count_results=True, total_results=True)
    multi.append(total / (time.perf_counter() - t0) / 1e9)
print(os.environ.get("GPUGREP_LIBRARY", "default"), "single GB/s", [round(x, 1) for x in single], "8 files GB/s", [round(x, 1) for x in multi])
for p in names + [f"{root}/big.log"]: os.unlink(p)
