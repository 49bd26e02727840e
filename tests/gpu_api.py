"""ctypes helpers for the extension entry points of include/gpugrep.h (test infrastructure)."""

from __future__ import annotations

import ctypes

from oracle_api import CALLBACK


class Stats(ctypes.Structure):
    _fields_ = [
        ("bytes_scanned", ctypes.c_ulonglong), ("lines", ctypes.c_ulonglong), ("matches", ctypes.c_ulonglong),
        ("candidates", ctypes.c_ulonglong), ("h2d_bytes", ctypes.c_ulonglong), ("d2h_bytes", ctypes.c_ulonglong),
        ("gpu_ms", ctypes.c_double), ("stream_kernel_ms", ctypes.c_double), ("wall_ms", ctypes.c_double),
        ("launches", ctypes.c_uint), ("stream_launches", ctypes.c_uint), ("segments", ctypes.c_uint), ("path", ctypes.c_uint),
        ("split_segments", ctypes.c_uint), ("reserved", ctypes.c_uint),
    ]


def marshal(patterns, flags=None, ids=None):
    n = len(patterns)
    flags = list(flags) if flags else [14] * n
    ids = list(ids) if ids else [0] * n
    pa = (ctypes.c_char_p * n)(*[p if isinstance(p, bytes) else p.encode() for p in patterns])
    return pa, (ctypes.c_uint * n)(*flags), (ctypes.c_uint * n)(*ids), n


def scan_buffer(lib, ptr: int, size: int, location: int, patterns, flags=None, ids=None, collect=True, buffer_size=262140,
                buffer_count=16, max_match_count=0, stream: int = 0):
    """gpugrep_scan_buffer -> (rc, records, stats).  records is None when collect=False (count-only, NULL callback)."""
    pa, fa, ia, n = marshal(patterns, flags, ids)
    lib.gpugrep_scan_buffer.argtypes = [
        ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint,
        ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_ulonglong, ctypes.c_void_p, ctypes.c_void_p,
    ]
    got = []

    def cb(results, count):
        for i in range(count):
            r = results[i]
            got.append((r.id, r.line_number, r.line))

    ccb = CALLBACK(cb) if collect else None
    stats = Stats()
    rc = lib.gpugrep_scan_buffer(ptr, size, location, pa, fa, ia, n, ctypes.cast(ccb, ctypes.c_void_p) if ccb else None, buffer_size,
                                 buffer_count, max_match_count, stream or None, ctypes.byref(stats))
    return rc, (got if collect else None), stats
