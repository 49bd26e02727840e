"""Test-only CPU walker for the tables the pattern compiler emits (never used by the product).

It loads a compiled database through the introspection entry points of include/gpugrep.h and walks the DFA
groups over pseudo-lines with the same rules the CUDA kernels implement, so the compiler can be checked against
the oracle on a machine without a GPU.
"""

from __future__ import annotations

import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class DbInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint) for n in (
        "patterns", "groups", "simple", "simple_id", "prefilter", "prefilter_stride", "prefilter_fold",
        "prefilter_log2_bits", "prefilter_grams", "prefilter_min_factor", "total_states", "reserved")]


class GroupInfo(ctypes.Structure):
    _fields_ = [("states", ctypes.c_uint), ("classes", ctypes.c_uint), ("stride", ctypes.c_uint),
                ("first_accept", ctypes.c_uint), ("sink_match", ctypes.c_int), ("dead", ctypes.c_int),
                ("accept_sets", ctypes.c_uint), ("members", ctypes.c_uint)]


def marshal(patterns, flags=None, ids=None):
    n = len(patterns)
    flags = list(flags) if flags else [14] * n
    ids = list(ids) if ids else [0] * n
    pa = (ctypes.c_char_p * n)(*[p if isinstance(p, bytes) else p.encode() for p in patterns])
    fa = (ctypes.c_uint * n)(*flags)
    ia = (ctypes.c_uint * n)(*ids)
    return pa, fa, ia, n


class CompiledDb:
    def __init__(self, lib: ctypes.CDLL, patterns, flags=None, ids=None):
        self.lib = lib
        lib.gpugrep_db_compile.restype = ctypes.c_void_p
        lib.gpugrep_last_error.restype = ctypes.c_char_p
        lib.gpugrep_db_prefilter_note.restype = ctypes.c_char_p
        lib.gpugrep_db_prefilter_note.argtypes = [ctypes.c_void_p]
        lib.gpugrep_db_copy_prefilter.restype = ctypes.c_size_t
        pa, fa, ia, n = marshal(patterns, flags, ids)
        rc = ctypes.c_int(0)
        self.handle = lib.gpugrep_db_compile(pa, fa, ia, n, ctypes.byref(rc))
        self.rc = rc.value
        self.error = lib.gpugrep_last_error().decode()
        self.groups = []
        if not self.handle:
            return
        h = ctypes.c_void_p(self.handle)
        self.info = DbInfo()
        lib.gpugrep_db_get_info(h, ctypes.byref(self.info))
        for g in range(self.info.groups):
            gi = GroupInfo()
            lib.gpugrep_db_get_group(h, g, ctypes.byref(gi))
            cls = np.zeros(256, dtype=np.uint8)
            trans = np.zeros(gi.states * gi.stride, dtype=np.uint32)
            acc = np.zeros(gi.states, dtype=np.uint32)
            lib.gpugrep_db_copy_group(h, g, cls.ctypes.data_as(ctypes.c_void_p), trans.ctypes.data_as(ctypes.c_void_p),
                                      acc.ctypes.data_as(ctypes.c_void_p))
            reports = []
            for k in range(gi.accept_sets):
                ids_buf = (ctypes.c_uint * 4096)()
                sm_buf = (ctypes.c_uint * 4096)()
                cnt = lib.gpugrep_db_accept_reports(h, g, k, ids_buf, sm_buf, 4096)
                reports.append([(ids_buf[i], sm_buf[i]) for i in range(cnt)])
            self.groups.append((gi, cls, trans.reshape(gi.states, gi.stride), acc, reports))
        self.prefilter_note = lib.gpugrep_db_prefilter_note(h).decode()
        self.bitmap = None
        if self.info.prefilter:
            words = (1 << self.info.prefilter_log2_bits) // 32
            bm = np.zeros(words, dtype=np.uint32)
            mul = ctypes.c_uint32(0)
            lib.gpugrep_db_copy_prefilter(h, bm.ctypes.data_as(ctypes.c_void_p), words, ctypes.byref(mul))
            self.bitmap = bm
            self.hash_mul = mul.value

    def __del__(self):
        if getattr(self, "handle", None):
            self.lib.gpugrep_db_free(ctypes.c_void_p(self.handle))
            self.handle = None

    # ---- line semantics shared with the kernels ----
    @staticmethod
    def block_of(pseudo_line: bytes) -> bytes:
        """Leading NULs stripped, cut at the next NUL (reference hyperscanner.c:205-217)."""
        i = 0
        while i < len(pseudo_line) and pseudo_line[i] == 0:
            i += 1
        rest = pseudo_line[i:]
        cut = rest.find(b"\0")
        return rest if cut < 0 else rest[:cut]

    def events(self, block: bytes):
        """All (end_offset, id, singlematch) reports of a block, unordered."""
        out = []
        for gi, cls, trans, acc, reports in self.groups:
            s = 0
            for pos, b in enumerate(block):
                s = int(trans[s, cls[b]])
                a = int(acc[s])
                if a:
                    for rid, sm in reports[a]:
                        out.append((pos, rid, sm))   # match ended just before byte `pos`
                    if gi.sink_match >= 0 and s == gi.sink_match:
                        break
                if s == gi.dead:
                    break
            else:
                s = int(trans[s, gi.classes])
                a = int(acc[s])
                if a:
                    for rid, sm in reports[a]:
                        out.append((len(block), rid, sm))
        return out

    def line_reports(self, pseudo_line: bytes):
        """Ordered ids delivered for one pseudo-line (dedupe per (id,end); SINGLEMATCH ids once)."""
        block = self.block_of(pseudo_line)
        evs = sorted(set(self.events(block)))
        if self.info.simple:
            return [self.info.simple_id] if evs else []
        seen = set()
        out = []
        for end, rid, sm in evs:
            if sm:
                if rid in seen:
                    continue
                seen.add(rid)
            out.append(rid)
        return out

    def prefilter_hits(self, text: bytes) -> bool:
        """True if any sampled 4-gram of `text` hits the bitmap (superset test used by the streaming kernel)."""
        if self.bitmap is None:
            return True
        st = self.info.prefilter_stride
        lb = self.info.prefilter_log2_bits
        for q in range(0, max(0, len(text) - 3), 1):
            gram = int.from_bytes(text[q:q + 4], "little")
            if self.info.prefilter_fold:
                gram |= 0x20202020
            h = ((gram * self.hash_mul) & 0xFFFFFFFF) >> (32 - lb)
            if (int(self.bitmap[h >> 5]) >> (h & 31)) & 1:
                if q % st == 0:
                    return True
        return False


def split_pseudo_lines(data: bytes, buffer_size: int):
    """gzgets() splitting (reference hyperscanner.c:199): up to buffer_size-1 bytes, stop after '\\n'."""
    out = []
    pos = 0
    lim = buffer_size - 1
    n = len(data)
    while pos < n:
        nl = data.find(b"\n", pos, pos + lim)
        end = nl + 1 if nl >= 0 else min(n, pos + lim)
        out.append(data[pos:end])
        pos = end
    return out
