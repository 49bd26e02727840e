"""ctypes access to the CPU oracle (oracle/hyperscanner_port.c).  Test infrastructure only."""

from __future__ import annotations

import ctypes
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_SO = os.path.join(ROOT, "oracle", "_build", "liboracle_hyperscanner.so")


class Result(ctypes.Structure):
    _fields_ = [("id", ctypes.c_uint), ("line_number", ctypes.c_ulonglong), ("line", ctypes.c_char_p)]


CALLBACK = ctypes.CFUNCTYPE(None, ctypes.POINTER(Result), ctypes.c_int)


def load_oracle() -> ctypes.CDLL:
    if not os.path.exists(ORACLE_SO):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")], stdout=subprocess.DEVNULL)
    return ctypes.CDLL(ORACLE_SO)


def run_scan(lib: ctypes.CDLL, path: str, patterns, flags=None, ids=None, buffer_size=262140, buffer_count=16,
             max_match_count=0):
    """Call `hyperscan` of any library with the reference ABI; returns (rc, [(id, line_number, line)], [batch sizes])."""
    n = len(patterns)
    flags = list(flags) if flags else [14] * n
    ids = list(ids) if ids else [0] * n
    pa = (ctypes.c_char_p * n)(*[p if isinstance(p, bytes) else p.encode() for p in patterns])
    fa = (ctypes.c_uint * n)(*flags)
    ia = (ctypes.c_uint * n)(*ids)
    got, batches = [], []

    def cb(results, count):
        batches.append(count)
        for i in range(count):
            r = results[i]
            got.append((r.id, r.line_number, r.line))

    ccb = CALLBACK(cb)
    rc = lib.hyperscan(path.encode(), pa, fa, ia, n, ccb, buffer_size, buffer_count, ctypes.c_ulonglong(max_match_count))
    return rc, got, batches


def scan_bytes(lib: ctypes.CDLL, data: bytes, patterns, **kw):
    with tempfile.NamedTemporaryFile(suffix=".txt", delete=False) as f:
        f.write(data)
        path = f.name
    try:
        return run_scan(lib, path, patterns, **kw)
    finally:
        os.unlink(path)


def check(lib: ctypes.CDLL, patterns, flags=None, ids=None) -> int:
    n = len(patterns)
    flags = list(flags) if flags else [14] * n
    ids = list(ids) if ids else [0] * n
    pa = (ctypes.c_char_p * n)(*[p if isinstance(p, bytes) else p.encode() for p in patterns])
    fa = (ctypes.c_uint * n)(*flags)
    ia = (ctypes.c_uint * n)(*ids)
    return lib.check_patterns(pa, fa, ia, n)
