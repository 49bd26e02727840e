"""Differential harness: any library with the reference ABI vs the CPU oracle, on the same inputs."""

from __future__ import annotations

import gzip
import os
import random
import tempfile

from oracle_api import run_scan
from regex_gen import gen_regex, gen_text


def compare(lib, oracle, data: bytes | None, patterns, flags=None, ids=None, buffer_size=262140, buffer_count=16,
            max_match_count=0, path: str | None = None):
    """Run both libraries on the same file and assert identical return code, records, order and batch sizes."""
    own = path is None
    if own:
        with tempfile.NamedTemporaryFile(suffix=".log", delete=False) as handle:
            handle.write(data)
            path = handle.name
    try:
        kw = {"flags": flags, "ids": ids, "buffer_size": buffer_size, "buffer_count": buffer_count, "max_match_count": max_match_count}
        exp_rc, exp, exp_batches = run_scan(oracle, path, patterns, **kw)
        got_rc, got, got_batches = run_scan(lib, path, patterns, **kw)
    finally:
        if own:
            os.unlink(path)
    assert got_rc == exp_rc, f"return code {got_rc} != oracle {exp_rc} for {patterns}"
    if got != exp:
        for k, (a, b) in enumerate(zip(got, exp)):
            if a != b:
                raise AssertionError(f"record {k}: got {a} expected {b}; patterns={patterns} flags={flags} ids={ids} bs={buffer_size}")
        raise AssertionError(f"record count {len(got)} != oracle {len(exp)}; patterns={patterns} flags={flags} ids={ids} bs={buffer_size}")
    assert got_batches == exp_batches, f"batch sizes differ: {got_batches[:8]} vs {exp_batches[:8]}"
    return len(exp)


def has_all_nul_pseudo_line(data: bytes, buffer_size: int) -> bool:
    """The reference reads stale buffer bytes for a pseudo-line made only of NULs (SURVEY.md §8a-2 rule 8): excluded."""
    pos, lim = 0, buffer_size - 1
    while pos < len(data):
        nl = data.find(b"\n", pos, pos + lim)
        end = nl + 1 if nl >= 0 else min(len(data), pos + lim)
        chunk = data[pos:end]
        if chunk and chunk[0] == 0 and not chunk.strip(b"\0"):
            return True
        pos = end
    return False


def random_case(seed: int):
    """One seeded random pattern set / text / parameter combination (patterns may be rejected by both sides)."""
    rng = random.Random(seed)
    k = rng.choice([1, 1, 1, 2, 3, 5])
    patterns = [gen_regex(rng) for _ in range(k)]
    mode = rng.choice(["simple", "simple", "ids", "nosm", "mixed"])
    base = rng.choice([14, 14, 14, 10, 12, 8, 15, 14])
    flags = [base] * k
    if mode == "simple":
        ids = [rng.choice([0, 7])] * k
    elif mode == "ids":
        ids = [rng.randint(0, 2) for _ in range(k)]
    elif mode == "nosm":
        flags = [f & ~8 for f in flags]
        ids = [rng.randint(0, 1) for _ in range(k)]
    else:
        ids = list(range(k))
        flags = [f & ~8 if i % 2 else f for i, f in enumerate(flags)]
    buffer_size = rng.choice([262140, 262140, 64, 9, 5])
    data = gen_text(rng, rng.choice([0, 1, 30, 200]), nul_rate=rng.choice([0, 0, 0.05]))
    buffer_count = rng.choice([16, 1, 3])
    max_match = rng.choice([0, 0, 0, 1, 4])
    return patterns, flags, ids, buffer_size, data, buffer_count, max_match


def gz_members(parts: list[bytes]) -> bytes:
    return b"".join(gzip.compress(p) for p in parts)


def zstd_frame(data: bytes, level: int = 3) -> bytes:
    """One zstd frame (system libzstd through ctypes; the image has no zstd module)."""
    import ctypes

    lib = ctypes.CDLL("libzstd.so.1")
    lib.ZSTD_compressBound.restype = ctypes.c_size_t
    lib.ZSTD_compressBound.argtypes = [ctypes.c_size_t]
    lib.ZSTD_compress.restype = ctypes.c_size_t
    lib.ZSTD_compress.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    bound = lib.ZSTD_compressBound(len(data))
    buf = ctypes.create_string_buffer(bound)
    size = lib.ZSTD_compress(buf, bound, data, len(data), level)
    return buf.raw[:size]


def zstd_skippable(payload: bytes = b"user data that is not text") -> bytes:
    """A skippable frame: magic 0x184D2A50, little-endian length, payload."""
    return b"\x50\x2a\x4d\x18" + len(payload).to_bytes(4, "little") + payload


def zstd_cases(text: bytes) -> dict:
    """Compressed layouts of `text` (which ends with a newline) and what the reference's zlibWrapper makes of them."""
    half = text.rfind(b"\n", 0, len(text) // 2) + 1
    a, b = text[:half], text[half:]
    return {
        "one_frame": zstd_frame(text),
        "two_frames": zstd_frame(a) + zstd_frame(b),                              # read through
        "three_frames_levels": zstd_frame(a, 1) + zstd_frame(b[: len(b) // 2], 9) + zstd_frame(b[len(b) // 2:], 3),
        "skippable_between": zstd_frame(a) + zstd_skippable() + zstd_frame(b),   # stops after the first frame
        "trailing_garbage": zstd_frame(a) + b"this is not zstd\n",               # garbage ignored
        "empty_frame_then_text": zstd_frame(b"") + zstd_frame(text),
        "truncated": zstd_frame(text)[: -7],                                      # what decodes before the error is scanned
    }


def member_layout_cases(kind: str, text: bytes, seed: int = 5) -> dict:
    """Files of MANY gzip members / zstd frames (what bgzip, pzstd or `cat a.gz b.gz` write), with the situations the
    multi-threaded decode (ingest_members.cpp) has to get right: bytes inside a member that look like a member header,
    empty members, one large member among small ones, garbage, a corrupt and a truncated member.  Every layout is read by
    the reference through ONE gzgets() stream (hyperscanner.c:189-199): that is what the oracle does."""
    rng = random.Random(seed)
    magic = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" if kind == "gzip" else b"\x28\xb5\x2f\xfd\x04\x58\x01\x00\x00"

    def pack(data: bytes, level: int | None = None) -> bytes:
        if kind == "gzip":
            return gzip.compress(data, compresslevel=6 if level is None else level)
        return zstd_frame(data, 3 if level is None else level)

    def stored(data: bytes) -> bytes:
        """A member whose payload is kept verbatim, so that header look-alikes inside `data` appear in the file."""
        if kind == "gzip":
            return gzip.compress(data, compresslevel=0)
        return zstd_frame(data, -5) if False else zstd_frame(data, 1)

    lines = text.splitlines(keepends=True)
    pieces, at = [], 0
    while at < len(lines):
        step = rng.choice([1, 7, 40, 200, 900])
        pieces.append(b"".join(lines[at:at + step]))
        at += step
    many = b"".join(pack(p, rng.choice([1, 6, 9]) if kind == "gzip" else rng.choice([1, 3, 9])) for p in pieces)
    # incompressible lines that contain member headers: the bytes show up verbatim inside the member
    noise = b"".join(magic + rng.randbytes(rng.randint(30, 200)).replace(b"\n", b" ").replace(b"\0", b" ") + b" ERROR port 77\n" for _ in range(300))
    third = len(pieces) // 3
    head, mid, tail = b"".join(pieces[:third]), b"".join(pieces[third:2 * third]), b"".join(pieces[2 * third:])
    small = lambda blob: b"".join(pack(q) for q in [blob[i:i + 20000] for i in range(0, len(blob), 20000)])   # cuts inside lines are fine
    corrupt = bytearray(small(head) + pack(mid) + small(tail))
    hit = len(small(head)) + len(pack(mid)) // 2
    corrupt[hit:hit + 8] = bytes(b ^ 0x5a for b in corrupt[hit:hit + 8])
    return {
        "many_members": many,
        "header_lookalikes": small(head) + stored(noise) + small(noise) + stored(noise + mid) + small(tail),
        "empty_members": pack(b"") + small(head) + pack(b"") + pack(b"") + small(mid + tail) + pack(b""),
        "large_member_among_small": small(head) + pack(mid + noise + mid) + small(tail),
        "garbage_after_members": small(head) + b"\n no member here \n" + small(tail),
        "garbage_that_starts_like_a_member": small(head) + magic[:2] + b"\xff\xff not a member" + small(tail),
        "corrupt_member": bytes(corrupt),
        "truncated_last_member": (small(head) + pack(mid))[:-9],
    }


ONLY_MATCHING_FIXED = ["ERROR", "port [0-9]{4}", "[a-f0-9]{8}", "(?:GET|PUT) /", r"user=\w\w\w", "a.c", r"\bsshd\b", "^Oct", "aa",
                       "(ab|cd)e", "x{3}", r"\d\d:\d\d", "[^ ]{5} ", r"\.\.", r"\x41B", "[]x]y", r"\Afoo", "(a|b)(c|d)", r"\Bo\B", r"\bA"]
ONLY_MATCHING_OTHER = ["ERR.*", "a+", "foo|barbaz", "(?i)error", "end$", r"\s\s", "colou?r", "x{2,3}", "[[:alpha:]]x", "^", r"\w+@"]


def only_matching_text(seed: int = 3) -> bytes:
    """Lines that exercise `grep -o`: overlapping candidates, several patterns on one line, lines that are not ASCII, a
    line with NULs, a last line without newline."""
    rng = random.Random(seed)
    words = ["ERROR", "error", "port 8080", "port 80", "deadbeef", "0123abcd", "GET /index", "PUT /x", "user=bob", "user=al", "abc", "axc",
             "sshd", "xsshd", "Oct 12", "aaaaa", "aa", "abe", "cde", "xxxxxxx", "12:34:56", "hello world ", "....", "AB", "]y", "foo", "ac", "bd",
             "colour", "color", "end", "tab\there", "caf\u00e9".encode().decode("latin-1"), "A"]
    lines = []
    for _ in range(1500):
        line = " ".join(rng.choice(words) for _ in range(rng.randint(1, 9)))
        lines.append(line.encode("latin-1") + b"\n")
    lines.insert(7, b"aaaaaaaaa ERROR aaaa\n")
    lines.insert(11, b"foo ERROR port 1234 with \xff\xfe bytes ERROR\n")
    lines.insert(13, b"\0\0ERROR after nuls ERROR\0 hidden ERROR\n")
    lines.insert(17, b"Oct 12 Oct 13 ERRORERROR\n")
    lines.append(b"ERROR last line without newline aa")
    return b"".join(lines)
