#!/usr/bin/env python3
"""Per-call cost of the file entry point on small files: one file at a time, then many at once."""
import ctypes, os, sys, time
from concurrent.futures import ThreadPoolExecutor
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from gpu_api import Stats, marshal
from hypergrep_b200 import synth, utils
lib = utils._get_hyperscanner_lib()
n_files, mib = int(sys.argv[1]), int(sys.argv[2])
root = "/dev/shm/gpugrep_fp"; os.makedirs(root, exist_ok=True)
names = []
for i in range(n_files):
    buf = np.zeros(mib << 20, dtype=np.uint8); synth.fill_syslog(buf, seed=500 + i, lib=lib)
    raw = bytes(buf[: int(np.flatnonzero(buf == 10)[-1]) + 1])
    names.append(f"{root}/f{i}.log"); open(names[-1], "wb").write(raw)
pa, fa, ia, n = marshal(synth.C2_PATTERNS)
lib.gpugrep_scan_file.argtypes = [ctypes.c_char_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p]
discard = ctypes.cast(lib.gpugrep_discard_results, ctypes.c_void_p)
def one(name):
    st = Stats(); t0 = time.perf_counter()
    rc = lib.gpugrep_scan_file(name.encode(), pa, fa, ia, n, discard, 262140, 4096, 0, ctypes.byref(st))
    return rc, time.perf_counter() - t0, st.matches
for rnd in range(3):
    ts = [one(nm)[1] for nm in names[:4]]
    print(f"round {rnd}: sequential per-call ms: " + " ".join(f"{t * 1e3:.1f}" for t in ts))
for workers in (4, 15):
    for rnd in range(2):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=workers) as pool: res = list(pool.map(one, names))
        dt = time.perf_counter() - t0
        print(f"{workers} workers round {rnd}: {n_files * mib / 1024 / dt * 1.0737:.2f} GB/s, slowest call {max(r[1] for r in res) * 1e3:.0f} ms")
for nm in names: os.remove(nm)
