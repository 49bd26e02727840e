#!/usr/bin/env python3
"""End-to-end probe: pinned host buffer -> gpugrep_scan_buffer, for a few chunk sizes and callback modes."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402

from gpu_api import Stats, marshal  # noqa: E402
from hypergrep_b200 import synth, utils  # noqa: E402

lib = utils._get_hyperscanner_lib()
lib.gpugrep_scan_buffer.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint,
                                    ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p]
size = int(float(sys.argv[1]) * (1 << 30)) if len(sys.argv) > 1 else 4 << 30
host = torch.empty(size, dtype=torch.uint8).pin_memory()
synth.fill_syslog(host.numpy(), seed=1234, lib=lib)
pa, fa, ia, n = marshal(synth.C2_PATTERNS)
discard = ctypes.cast(lib.gpugrep_discard_results, ctypes.c_void_p)
# raw H2D ceiling
dev = torch.empty(size, dtype=torch.uint8, device="cuda")
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter(); dev.copy_(host, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"raw H2D copy: {size / dt / 1e9:.1f} GB/s")
for chunk in (32, 64, 128, 256, 512):
    os.environ["GPUGREP_CHUNK_MB"] = str(chunk)
    for name, cb in (("discard", discard), ("count", None)):
        best = 0
        for _ in range(3):
            st = Stats()
            t0 = time.perf_counter()
            rc = lib.gpugrep_scan_buffer(host.data_ptr(), size, 0, pa, fa, ia, n, cb, 262140, 4096, 0, None, ctypes.byref(st))
            dt = time.perf_counter() - t0
            assert rc == 0
            best = max(best, size / dt / 1e9)
        print(f"chunk {chunk:4d} MiB  callback={name:8s} e2e {best:6.1f} GB/s  matches {st.matches} segments {st.segments}")
