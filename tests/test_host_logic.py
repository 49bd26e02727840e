"""Host logic (capi.cpp segment cutting / batching / stop rule, ingest, pattern compiler) against the oracle, on CPU,
through the mock engine.  The same cases run against the CUDA engine in test_gpu_parity.py."""

import ctypes
import os

import pytest

import parity
from oracle_api import scan_bytes as run_scan_bytes
from hypergrep_b200 import synth


@pytest.mark.parametrize("seed", range(120))
def test_random_cases_match_oracle(seed, hostmock_lib, oracle_lib):
    patterns, flags, ids, buffer_size, data, buffer_count, max_match = parity.random_case(seed)
    if parity.has_all_nul_pseudo_line(data, buffer_size):
        pytest.skip("all-NUL pseudo-line: reference reads stale bytes, excluded from parity")
    parity.compare(hostmock_lib, oracle_lib, data, patterns, flags, ids, buffer_size, buffer_count, max_match)


EDGE_TEXTS = {
    "empty": b"",
    "one_line_no_newline": b"foo bar",
    "only_newlines": b"\n\n\n\n",
    "crlf": b"foo\r\nbar foo\r\n\r\nfoo",
    "leading_nuls": b"\0\0foo\nbar\n\0foo bar\n",
    "embedded_nul": b"foo\0bar\nbar\0foo\nfoo\n",
    "long_line": b"x" * 700 + b"foo" + b"y" * 300 + b"\nfoo\n",
    "boundary_lengths": b"a" * 62 + b"\n" + b"b" * 63 + b"\n" + b"c" * 64 + b"\n" + b"foo" * 21 + b"\n" + b"foo" * 42 + b"\nfoo",
    "match_split_by_buffer": b"0123456fo" + b"obar\n" + b"foobar\n",
}


@pytest.mark.parametrize("name", sorted(EDGE_TEXTS))
@pytest.mark.parametrize("buffer_size", [262140, 64, 8])
def test_edge_texts(name, buffer_size, hostmock_lib, oracle_lib):
    for patterns in (["foo"], ["foo$", "^bar"], ["o+b", "x{3}"], ["foo."]):
        parity.compare(hostmock_lib, oracle_lib, EDGE_TEXTS[name], patterns, buffer_size=buffer_size)
    parity.compare(hostmock_lib, oracle_lib, EDGE_TEXTS[name], ["foo", "bar", "o"], flags=[14, 14, 6], ids=[1, 2, 3], buffer_size=buffer_size)


def test_multi_segment_syslog(hostmock_lib, oracle_lib, monkeypatch, tmp_path):
    """3 MiB of syslog text cut into ~530 KiB segments: line numbers and records must not depend on the cuts."""
    monkeypatch.setenv("GPUGREP_CHUNK_BYTES", "1")
    data = synth.syslog_bytes(3 << 20, seed=5, lib=hostmock_lib)
    n = parity.compare(hostmock_lib, oracle_lib, data, synth.C1_PATTERNS)
    assert n > 100
    parity.compare(hostmock_lib, oracle_lib, data, synth.C2_PATTERNS)
    parity.compare(hostmock_lib, oracle_lib, data, synth.C2_PATTERNS, max_match_count=1000, buffer_count=7)
    parity.compare(hostmock_lib, oracle_lib, data[: 1 << 20], ["ERROR", "port [0-9]+", "WARN"], flags=[14, 14, 14], ids=[3, 1, 2])
    # small gzgets buffer: every line is split into pseudo-lines, across segment cuts as well
    parity.compare(hostmock_lib, oracle_lib, data[: 1 << 20], ["ERROR", "ssh2$"], buffer_size=50)


def test_compressed_inputs(hostmock_lib, oracle_lib, tmp_path):
    text = synth.syslog_bytes(1 << 20, seed=9, lib=hostmock_lib)
    half = text.rfind(b"\n", 0, len(text) // 2) + 1
    multi = tmp_path / "multi.log.gz"
    multi.write_bytes(parity.gz_members([text[:half], text[half:]]))
    parity.compare(hostmock_lib, oracle_lib, None, ["ERROR"], path=str(multi))
    garbage = tmp_path / "garbage.log.gz"
    garbage.write_bytes(parity.gz_members([text[:half]]) + b"this is not gzip")
    parity.compare(hostmock_lib, oracle_lib, None, ["ERROR"], path=str(garbage))


def test_missing_file_and_directory(hostmock_lib, oracle_lib, tmp_path):
    parity.compare(hostmock_lib, oracle_lib, None, ["foo"], path=str(tmp_path / "nope.txt"))   # both: 6 (GZ_OPEN)
    parity.compare(hostmock_lib, oracle_lib, None, ["foo"], path=str(tmp_path))                # both: 0, no lines
    assert os.path.isdir(tmp_path)


@pytest.mark.parametrize("seed", range(200, 240))
def test_nfa_fallback_random_cases(seed, hostmock_lib, oracle_lib, monkeypatch):
    """Every pattern forced through the bit-parallel NFA fallback (DFA state budget of 3): same reports as the oracle."""
    monkeypatch.setenv("GPUGREP_MAX_DFA_STATES", "3")
    patterns, flags, ids, buffer_size, data, buffer_count, max_match = parity.random_case(seed)
    if parity.has_all_nul_pseudo_line(data, buffer_size):
        pytest.skip("all-NUL pseudo-line")
    parity.compare(hostmock_lib, oracle_lib, data, patterns, flags, ids, buffer_size, buffer_count, max_match)


def test_nfa_fallback_for_exploding_patterns(hostmock_lib, oracle_lib):
    """Patterns whose DFA is exponential (a gap of n arbitrary bytes) take the NFA path next to ordinary DFA patterns."""
    text = synth.syslog_bytes(256 << 10, seed=13, lib=hostmock_lib)
    patterns = [r"e.{60}d\b", r"ERROR", r"\bport .{40,80}x"]
    assert parity.compare(hostmock_lib, oracle_lib, text, patterns) > 10
    parity.compare(hostmock_lib, oracle_lib, text[: 64 << 10], patterns, flags=[14, 14, 6], ids=[1, 2, 3])


def test_default_buffer_boundary_lines(hostmock_lib, oracle_lib):
    """Lines of exactly buffer_size-2, -1, 0, +1 bytes around the default gzgets buffer (262,140): the pseudo-line split
    (SURVEY.md §8a-2 rule 1) with matches on both sides of every cut, between ordinary lines."""
    limit = 262140 - 1
    parts = [b"foo first\n"]
    for total in (limit - 1, limit, limit + 1, limit + 2, 2 * limit, 2 * limit + 5):
        body = bytearray(b"x" * (total - 1))
        body[10:13] = b"foo"
        body[limit - 2:limit + 1] = b"foo" if len(body) > limit + 1 else body[limit - 2:limit + 1]   # straddles the first cut
        if len(body) > limit + 20:
            body[limit + 5:limit + 8] = b"foo"
        parts.append(bytes(body) + b"\n")
        parts.append(b"bar between\n")
    parts.append(b"foo last")
    data = b"".join(parts)
    assert parity.compare(hostmock_lib, oracle_lib, data, ["foo"]) >= 8
    parity.compare(hostmock_lib, oracle_lib, data, ["foo", "bar"], flags=[14, 6], ids=[1, 2])
    parity.compare(hostmock_lib, oracle_lib, data, ["o{2}$", "^x+foo"])


def test_grep_count_only_paths_agree(hostmock_lib, tmp_path, monkeypatch):
    """grep(count_only=True) counts inside the library (gpugrep_scan_file with a NULL callback, no Python frame per batch);
    with a library that lacks that entry point (a plain libhyperscanner.so) it counts through the callback like the
    reference (utils.py:199-203).  Both must equal the number of collected lines, with and without max_match_count."""
    from hypergrep_b200 import utils

    text = b"".join(b"line %d %s\n" % (k, b"foobar" if k % 3 == 0 else b"nothing") for k in range(5000))
    path = tmp_path / "count.log"
    path.write_bytes(text)
    monkeypatch.setattr(utils, "_get_hyperscanner_lib", lambda: hostmock_lib)
    lines, rc = utils.grep(str(path), ["foobar", "^line 7 "])
    assert rc == 0 and len(lines) == 1667 + 1   # line 7 is no multiple of 3
    for limit in (0, 1, 16, 17, 1000):
        expected = len(lines) if limit == 0 else min(limit, len(lines))
        assert utils.grep(str(path), ["foobar", "^line 7 "], count_only=True, max_match_count=limit) == (expected, 0)

    class CallbackOnly:   # what a library exporting only the reference's two symbols looks like
        hyperscan = hostmock_lib.hyperscan
        check_patterns = hostmock_lib.check_patterns

        def __getattr__(self, name):
            raise AttributeError(name)

    monkeypatch.setattr(utils, "_get_hyperscanner_lib", CallbackOnly)
    assert utils._count_matches(str(path), ["foobar"], [14], 0) is None
    for limit in (0, 17):
        expected = len(lines) if limit == 0 else limit
        assert utils.grep(str(path), ["foobar", "^line 7 "], count_only=True, max_match_count=limit) == (expected, 0)


def test_multiline_circumflex_after_the_final_newline(hostmock_lib, oracle_lib):
    """Hyperscan's multiline ^ holds after ANY newline, also the one that ends the scanned block; PCRE's
    default does not, so the oracle compiles with PCRE2_ALT_CIRCUMFLEX.  `x\\W^` therefore matches the line "x\\n"."""
    data = b"x\n\t xxx\t_a1AB\nyx\nx"
    rc, got, _ = run_scan_bytes(oracle_lib, data, ["x\\W^"])
    assert rc == 0 and [ln for (_i, ln, _t) in got] == [0, 2]
    for patterns in (["x\\W^"], ["x\\W\\Z\\z^([^a]a)*?"], ["^x"], ["x\\s^", "^y"]):
        parity.compare(hostmock_lib, oracle_lib, data, patterns, flags=[14] * len(patterns), ids=list(range(len(patterns))), buffer_size=64)
        parity.compare(hostmock_lib, oracle_lib, data, patterns)


@pytest.mark.parametrize("devices", ["all", "0,1,2"])
def test_one_input_over_several_devices_merges_like_one(devices, hostmock_lib, oracle_lib, monkeypatch, tmp_path):
    """$GPUGREP_DEVICES: hyperscan(path) on a plain file and gpugrep_scan_buffer on host memory split the input into one
    newline-aligned range per device and merge on the calling thread: same records, same line
    numbers, same batch sizes, same max_match_count cut as one device (the mock engine pretends three devices)."""
    from gpu_api import scan_buffer

    monkeypatch.setenv("GPUGREP_MOCK_DEVICES", "3")
    monkeypatch.setenv("GPUGREP_DEVICES", devices)
    monkeypatch.setenv("GPUGREP_MIN_SHARD_BYTES", "100000")
    data = synth.syslog_bytes(1 << 20, seed=77)
    for patterns, kw in ((synth.C2_PATTERNS, {}), (synth.C2_PATTERNS, {"max_match_count": 1000, "buffer_count": 7}),
                         (["ERROR", "port [0-9]+", "o"], {"flags": [14, 14, 6], "ids": [3, 1, 2], "max_match_count": 5000}),
                         (["ERROR", "ssh2$"], {"buffer_size": 50}), (["no such text anywhere"], {})):
        assert parity.compare(hostmock_lib, oracle_lib, data, patterns, **kw) >= 0
    # host-buffer entry point, count-only and collecting
    rc, expected, _ = run_scan_bytes(oracle_lib, data, synth.C2_PATTERNS)
    buf = ctypes.create_string_buffer(data, len(data))
    rc, got, stats = scan_buffer(hostmock_lib, ctypes.addressof(buf), len(data), 0, synth.C2_PATTERNS)
    assert rc == 0 and got == expected and stats.lines == data.count(b"\n") and stats.segments >= 3
    rc, none, stats = scan_buffer(hostmock_lib, ctypes.addressof(buf), len(data), 0, synth.C2_PATTERNS, collect=False)
    assert rc == 0 and none is None and stats.matches == len(expected)
    # a text without any newline cannot be split: the boundaries collapse and one shard does the work
    blob = b"x" * 300000 + b"foo" + b"y" * 300000
    parity.compare(hostmock_lib, oracle_lib, blob, ["foo"])
    parity.compare(hostmock_lib, oracle_lib, b"", ["foo"])


def test_zstd_frame_layouts(hostmock_lib, oracle_lib, tmp_path):
    """zstd inputs (reference: gzopen/gzgets of zstd's zlibWrapper, hyperscanner.c:189-199): several frames are read
    through; a skippable frame or garbage after a frame ends the data (gz_look() only continues on a gzip / zstd header)."""
    text = synth.syslog_bytes(512 << 10, seed=19)
    counts = {}
    for name, blob in parity.zstd_cases(text).items():
        path = tmp_path / f"{name}.log.zst"
        path.write_bytes(blob)
        counts[name] = parity.compare(hostmock_lib, oracle_lib, None, ["ERROR", "port [0-9]+"], path=str(path))
    assert counts["one_frame"] == counts["two_frames"] == counts["three_frames_levels"] == counts["empty_frame_then_text"] > 50
    assert 0 < counts["skippable_between"] == counts["trailing_garbage"] < counts["one_frame"]


@pytest.mark.parametrize("kind", ["gzip", "zstd"])
def test_many_members_decoded_by_several_threads(hostmock_lib, oracle_lib, tmp_path, monkeypatch, kind):
    """One file of many gzip members / zstd frames, decoded ahead of the reader by helper threads from speculative
    starts (ingest_members.cpp): the records equal those of the reference's single gzgets() stream (hyperscanner.c:189-199)
    in every layout, and those of the one-thread decode of the library itself."""
    text = synth.syslog_bytes(1 << 20, seed=23)
    patterns = ["ERROR", "port [0-9]+"]
    suffix = ".log.gz" if kind == "gzip" else ".log.zst"
    for name, blob in parity.member_layout_cases(kind, text).items():
        path = tmp_path / f"{name}{suffix}"
        path.write_bytes(blob)
        monkeypatch.setenv("GPUGREP_DECODE_THREADS", "0")
        one = parity.compare(hostmock_lib, oracle_lib, None, patterns, path=str(path))
        for spacing, chain in (("512", "20000"), ("4096", "200000"), ("100000", "1")):
            monkeypatch.setenv("GPUGREP_DECODE_THREADS", "4")
            monkeypatch.setenv("GPUGREP_DECODE_MIN_BYTES", "0")
            monkeypatch.setenv("GPUGREP_DECODE_SPACING", spacing)
            monkeypatch.setenv("GPUGREP_DECODE_CHAIN", chain)
            many = parity.compare(hostmock_lib, oracle_lib, None, patterns, path=str(path))
            assert many == one, (kind, name, spacing)
        assert one > 0, (kind, name)


def test_ingest_probe_same_text_with_any_thread_count(hostmock_lib, tmp_path, monkeypatch):
    """gpugrep_ingest_probe (the ingest alone): byte count and hash of the text do not depend on the decode threads,
    and equal Python's own decode."""
    import gzip

    text = synth.syslog_bytes(3 << 20, seed=29)
    files = {"t.log": text, "t.log.gz": parity.member_layout_cases("gzip", text)["many_members"],
             "t.log.zst": parity.member_layout_cases("zstd", text)["many_members"], "one.log.gz": gzip.compress(text, 1)}
    hostmock_lib.gpugrep_ingest_probe.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]
    monkeypatch.setenv("GPUGREP_DECODE_MIN_BYTES", "0")
    monkeypatch.setenv("GPUGREP_DECODE_SPACING", "3000")
    seen = set()
    for name, blob in files.items():
        path = tmp_path / name
        path.write_bytes(blob)
        for threads in ("0", "2", "5"):
            monkeypatch.setenv("GPUGREP_DECODE_THREADS", threads)
            size, digest = ctypes.c_ulonglong(), ctypes.c_ulonglong()
            assert hostmock_lib.gpugrep_ingest_probe(str(path).encode(), ctypes.byref(size), ctypes.byref(digest)) == 0
            assert size.value == len(text)
            seen.add(digest.value)
    assert len(seen) == 1
    assert hostmock_lib.gpugrep_ingest_probe(str(tmp_path / "missing").encode(), None, None) == 6


def check_only_matching(lib, tmp_path, monkeypatch):
    """grep(only_matching=True) with spans from the engine's match ends == the reference's algorithm (re.finditer over
    every matched line, utils.py:205-212), record for record."""
    from hypergrep_b200 import utils

    monkeypatch.setattr(utils, "_get_hyperscanner_lib", lambda: lib)
    path = tmp_path / "only.log"
    path.write_bytes(parity.only_matching_text())
    widths = utils._span_widths(parity.ONLY_MATCHING_FIXED + parity.ONLY_MATCHING_OTHER)
    assert all(w > 0 for w in widths[:len(parity.ONLY_MATCHING_FIXED)]), widths
    assert all(w == -1 for w in widths[len(parity.ONLY_MATCHING_FIXED):]), widths
    used = []
    original = utils._fill_spans

    def counting(groups, deferred, *rest):
        used.append(len(deferred))
        return original(groups, deferred, *rest)

    sets = [[p] for p in parity.ONLY_MATCHING_FIXED] + [parity.ONLY_MATCHING_FIXED, parity.ONLY_MATCHING_FIXED[:4] + parity.ONLY_MATCHING_OTHER[:3],
                                                     parity.ONLY_MATCHING_OTHER[:3]]
    for patterns in sets:
        for ignore_case in (False, True):
            monkeypatch.setattr(utils, "_fill_spans", counting)
            got = utils.grep(str(path), patterns, only_matching=True, ignore_case=ignore_case)
            with monkeypatch.context() as plain:   # the reference's way: no widths, finditer for every record
                plain.setattr(utils, "_span_widths", lambda _patterns: None)
                expected = utils.grep(str(path), patterns, only_matching=True, ignore_case=ignore_case)
            assert got == expected, (patterns, ignore_case)
    assert sum(used) > 3000   # most records did take the engine's ends
    # random fixed-width patterns over random text
    import random

    rng = random.Random(11)
    atoms = ["a", "b", "ab", "[ab]", ".", "[^a]", "(a|b)", "(?:ab|ba)", "a{2}", "[a-c]{3}", r"\w", r"\d", r"\bb", "c"]
    for round_number in range(40):
        patterns = ["".join(rng.choice(atoms) for _ in range(rng.randint(1, 4))) for _ in range(rng.randint(1, 3))]
        text = b"".join(bytes(rng.choice(b"abc1 ") for _ in range(rng.randint(0, 30))) + b"\n" for _ in range(200))
        path.write_bytes(text)
        got = utils.grep(str(path), patterns, only_matching=True)
        with monkeypatch.context() as plain:
            plain.setattr(utils, "_span_widths", lambda _patterns: None)
            expected = utils.grep(str(path), patterns, only_matching=True)
        assert got == expected, (round_number, patterns)


def test_only_matching_spans_from_match_ends(hostmock_lib, tmp_path, monkeypatch):
    check_only_matching(hostmock_lib, tmp_path, monkeypatch)


def test_match_ends_entry_point(hostmock_lib):
    """gpugrep_match_ends: (line, id, end) for every match end, hs_scan order, SINGLEMATCH ignored, capacity respected."""
    from hypergrep_b200 import utils

    text = b"abcabc x\n\nzzabc\nno match here\nbcbc"
    patterns, flags, ids = utils.prepare_patterns(["abc", "bc", "c x"], flags=[8, 0, 0], ids=[5, 6, 7])
    entry = hostmock_lib.gpugrep_match_ends
    entry.restype = ctypes.c_int
    entry.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_uint,
                      ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]
    expected = [(0, 5, 3), (0, 6, 3), (0, 5, 6), (0, 6, 6), (0, 7, 8), (2, 5, 5), (2, 6, 5), (4, 6, 2), (4, 6, 4)]
    for capacity in (64, 4, 0):
        out = (utils._MatchEnd * max(1, capacity))()
        found = ctypes.c_size_t()
        assert entry(text, len(text), 0, patterns, flags, ids, 3, 262140, out, capacity, ctypes.byref(found), None) == 0
        assert found.value == len(expected)
        got = [(out[k].line_number, out[k].id, out[k].end) for k in range(min(capacity, found.value))]
        assert got == expected[:capacity]
    bad, bad_flags, bad_ids = utils.prepare_patterns(["(unclosed"], flags=[0], ids=[0])
    assert entry(text, len(text), 0, bad, bad_flags, bad_ids, 1, 262140, None, 0, None, None) == 4


def test_every_end_of_a_trailing_repeat_is_reported(hostmock_lib, oracle_lib):
    """Without HS_FLAG_SINGLEMATCH Hyperscan reports EVERY end offset: `user=\\w+` over "user=bob" ends at 6, 7 and 8 - three
    reports, three callbacks with the same line (hyperscanner.c:83-102 copies the line per report).  (The oracle needs
    PCRE2_NO_AUTO_POSSESS for this: PCRE2 otherwise makes the trailing repeat possessive and its DFA matcher returns the
    longest match only - found by tests/guarded_device_scan.py in round 2.)"""
    text = b"user=bob and user=al\nnothing here\nport 8080 user=x\n12345\n"
    patterns, flags, ids = ["user=\\w+", "[0-9]+", "port [0-9]+"], [6, 6, 14], [1, 2, 3]
    count = parity.compare(hostmock_lib, oracle_lib, text, patterns, flags=flags, ids=ids)
    # line 0: 3 + 2 ends of user=..; line 2: port (once, SINGLEMATCH), 4 digit ends, 1 user= end; line 3: 5 digit ends
    assert count == 5 + (1 + 4 + 1) + 5
    rc, got, _ = run_scan_bytes(hostmock_lib, text, patterns, flags=flags, ids=ids)
    assert rc == 0 and [r[0] for r in got if r[1] == 0] == [1, 1, 1, 1, 1]


@pytest.mark.parametrize("kind", ["gzip", "zstd"])
def test_random_member_layouts_decode_like_one_stream(hostmock_lib, tmp_path, monkeypatch, kind):
    """Differential test of the multi-threaded member decode (ingest_members.cpp) against the one-thread decode of the same
    library: random member sizes (empty ones included), stored members whose payload contains header look-alikes, and
    random damage - a flipped byte, a truncation, garbage between members - under random start spacings, hand-over
    sizes and thread counts.  Byte count and hash of the delivered text must be identical."""
    import gzip
    import random

    hostmock_lib.gpugrep_ingest_probe.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_ulonglong), ctypes.POINTER(ctypes.c_ulonglong)]
    rng = random.Random(1234 if kind == "gzip" else 4321)
    text = synth.syslog_bytes(1 << 20, seed=47)
    magic = b"\x1f\x8b\x08\x00\x00\x00\x00\x00\x00\x03" if kind == "gzip" else b"\x28\xb5\x2f\xfd\x04\x58\x01\x00\x00"
    pack = (lambda d, lvl: gzip.compress(d, lvl)) if kind == "gzip" else (lambda d, lvl: parity.zstd_frame(d, max(1, lvl)))
    path = tmp_path / ("r.log.gz" if kind == "gzip" else "r.log.zst")

    def probe():
        size, digest = ctypes.c_ulonglong(), ctypes.c_ulonglong()
        assert hostmock_lib.gpugrep_ingest_probe(str(path).encode(), ctypes.byref(size), ctypes.byref(digest)) == 0
        return size.value, digest.value

    for case in range(40):
        members, at = [], 0
        while at < len(text) and len(members) < 60:
            step = rng.choice([0, 1, 300, 5000, 40000, 200000])
            piece = text[at:at + step]
            at += step
            if rng.random() < 0.15:   # incompressible payload with member headers inside
                piece = b"".join(magic + rng.randbytes(50) for _ in range(40)) + piece
                members.append(pack(piece, 0 if kind == "gzip" else 1))
            else:
                members.append(pack(piece, rng.choice([1, 6, 9])))
        blob = bytearray(b"".join(members))
        damage = rng.choice(["none", "none", "flip", "truncate", "garbage", "garbage_magic"])
        if damage == "flip" and len(blob) > 100:
            spot = rng.randrange(20, len(blob) - 1)
            blob[spot] ^= 0x55
        elif damage == "truncate" and len(blob) > 100:
            del blob[rng.randrange(10, len(blob)):]
        elif damage in ("garbage", "garbage_magic") and len(members) > 2:
            cut = len(b"".join(members[:rng.randrange(1, len(members))]))
            blob[cut:cut] = (magic[:2] if damage == "garbage_magic" else b"") + rng.randbytes(rng.randrange(1, 40))
        path.write_bytes(bytes(blob))
        monkeypatch.setenv("GPUGREP_DECODE_THREADS", "0")
        expected = probe()
        for _ in range(2):
            monkeypatch.setenv("GPUGREP_DECODE_THREADS", str(rng.choice([2, 3, 5, 9])))
            monkeypatch.setenv("GPUGREP_DECODE_MIN_BYTES", "0")
            monkeypatch.setenv("GPUGREP_DECODE_SPACING", str(rng.choice([64, 1000, 20000, 300000])))
            monkeypatch.setenv("GPUGREP_DECODE_CHAIN", str(rng.choice([1, 5000, 100000, 16 << 20])))
            assert probe() == expected, (kind, case, damage)
