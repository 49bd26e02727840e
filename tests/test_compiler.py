"""Pattern compiler checks on CPU: the tables libgpugrep.so emits, walked by tests/dfa_sim.py, against the oracle."""

import ctypes
import random

import pytest

import parity
from dfa_sim import CompiledDb, fast_path_matched_line_starts, split_pseudo_lines
from oracle_api import check, scan_bytes
from regex_gen import gen_regex, gen_text

ACCEPTED = ["foobar", "fo{2}bar", "fo+bar", "^foo$", r"\bfoo\b", "[[:alpha:]]+[0-9]", r"a\x41\n", r"\Qa.b\E+", "(?i)abc", "(?x) a b c # comment",
            "(?P<name>ab)+c", "a{2,3}?b", r"[^\n]x", r"\d{1,3}\.\d{1,3}", "(a|b)*abb", r"x\z", r"x\Z", r"\Afoo", "(?s:a.b)", "a(?#note)b"]
REJECTED = ["(?<!foo)bar", "(?=a)b", "(?!a)b", "(?<=a)b", r"(a)\1", "(?>ab)c", "a*+", "a++", "(?(1)a|b)", "(?R)", r"a\Kb", r"\X", r"\C", r"\R",
            "(*UTF)a", "a*", "foo|", "^", "$", r"\b", "(a", "a)", "[a", "a{3,2}", r"\p{L}", "(?C1)a", "a{40000}", r"\x{100}", "(?P=n)", "", "x*?"]


def test_accept_reject_table(gpu_lib, oracle_lib):
    """Accept/reject decisions (reference check_patterns, hyperscanner.c:154-167) agree with the oracle's Hyperscan rules."""
    for pattern in ACCEPTED:
        assert CompiledDb(gpu_lib, [pattern]).rc == 0, pattern
        assert check(oracle_lib, [pattern]) == 0, pattern
    for pattern in REJECTED:
        if pattern == "":
            continue
        assert CompiledDb(gpu_lib, [pattern]).rc == 4, pattern
        assert check(oracle_lib, [pattern]) == 4, pattern
    # ids sharing SINGLEMATCH inconsistently, unknown flag bits
    assert CompiledDb(gpu_lib, ["a", "b"], [14, 6], [1, 1]).rc == 4
    assert CompiledDb(gpu_lib, ["a"], [14 | 32]).rc == 4


@pytest.mark.parametrize("seed", range(60))
def test_tables_match_oracle(seed, gpu_lib, oracle_lib):
    """DFA tables (all groups, accept sets, report lists) reproduce the oracle's per-line reports, incl. multiple ids,
    non-SINGLEMATCH patterns, NULs and gzgets splitting."""
    patterns, flags, ids, buffer_size, data, _, _ = parity.random_case(1000 + seed)
    if parity.has_all_nul_pseudo_line(data, buffer_size):
        pytest.skip("all-NUL pseudo-line")
    db = CompiledDb(gpu_lib, patterns, flags, ids)
    assert (db.rc == 0) == (check(oracle_lib, patterns, flags, ids) == 0)
    if db.rc:
        return
    _, got, _ = scan_bytes(oracle_lib, data, patterns, flags=flags, ids=ids, buffer_size=buffer_size)
    expected = [(i, ln) for (i, ln, _line) in got]
    mine = []
    for ln, pl in enumerate(split_pseudo_lines(data, buffer_size)):
        reports = db.line_reports(pl)
        mine += [(rid, ln) for rid in reports]
        assert not reports or db.prefilter_hits(db.block_of(pl)), "prefilter must be a superset filter"
    assert mine == expected


LITS = ["abc", "bca", "a b", "Ab1", "x_1", "c.a", "ab", "bb", "1x"]


def _factor_pattern(rng: random.Random) -> str:
    core = "".join(rng.choice(LITS) for _ in range(rng.randint(2, 3)))
    pre = rng.choice(["", "", "^", r"\b", "x?", "[ab]{0,2}", "(?:a|b_)", r"\w+", ".*", "a{2,}"])
    suf = rng.choice(["", "", "$", r"\b", "c*", "[^a]", "(?:x|1)+", r"\s", r"\d{1,2}", ".b"])
    return pre + core.replace(".", r"\.") + suf


@pytest.mark.parametrize("seed", range(60))
def test_fast_path_model_is_exact(seed, gpu_lib):
    """The fast-path ALGORITHM (gram hits -> local DFA walk with look-back, mid-line entry states, idle stop, NUL
    re-check), modelled on CPU over the compiler's tables, finds exactly the lines the full per-line walk finds."""
    rng = random.Random(seed)
    k = rng.choice([1, 2, 4])
    patterns = [_factor_pattern(rng) for _ in range(k)]
    db = CompiledDb(gpu_lib, patterns, [rng.choice([14, 14, 15, 10])] * k, [0] * k)
    assert db.rc == 0
    if not db.info.prefilter:
        pytest.skip("no usable factor")
    lines = []
    for _ in range(40):
        text = "".join(rng.choice("abcAB _x1.\t") for _ in range(rng.randint(0, 50)))
        if rng.random() < 0.5:
            at = rng.randint(0, len(text))
            text = text[:at] + "".join(rng.choice(LITS) for _ in range(rng.randint(1, 3))) + text[at:]
        if rng.random() < 0.1:
            at = rng.randint(0, len(text))
            text = text[:at] + "\0" + text[at:]
        lines.append(text)
    data = ("\n".join(lines) + ("\n" if rng.random() < 0.7 else "")).encode("latin1")
    if parity.has_all_nul_pseudo_line(data, 262140):
        pytest.skip("all-NUL pseudo-line")
    if rng.random() < 0.5:
        gpu_lib.gpugrep_db_tune.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        db.retune(data[: max(8, len(data) // 2)])
    truth, pos = set(), 0
    for pl in split_pseudo_lines(data, 262140):
        if db.line_reports(pl):
            truth.add(pos)
        pos += len(pl)
    assert fast_path_matched_line_starts(db, data) == truth


def test_mixed_sampling_is_chosen_for_few_short_factors(gpu_lib):
    """One 5-byte factor next to long ones: table lookups at stride 4 plus register compares at offsets 2 (mod 4)
    (gpugrep.h, gpugrep_db_copy_odd_compares) instead of stride 2 for everything."""
    db = CompiledDb(gpu_lib, ["ERROR", "connection reset by peer", "segfault at"])
    assert db.rc == 0 and db.info.prefilter_stride == 4
    assert len(db.odd) == 2   # "ERRO" and "RROR", exact (mul == 1)
    assert sorted((-add) & 0xFFFFFFFF for mul, add in db.odd) == sorted(
        int.from_bytes(g, "little") for g in (b"ERRO", b"RROR"))
    for shift in range(8):   # every alignment of the short factor is seen
        assert db.prefilter_hits(b"x" * shift + b"an ERROR here\n")
        assert db.prefilter_hits(b"y" * shift + b"segfault at 0\n")
    assert not db.prefilter_hits(b"nothing to see, move along please\n")
    # too many short factors: plain stride 2
    many = CompiledDb(gpu_lib, ["ERROR", "WARN:", "FATAL", "PANIC", "ALERT"])
    assert many.info.prefilter_stride == 2 and not many.odd


ANCHORED = [["^GET "], ["^abcd"], ["^foo:"], ["^.abcdefg"], [r"\AGET /index"], ["^GET ", "POST /submit"], [r"^\s*warn: disk"]]


def anchored_alignment_text(head: bytes, body: bytes = b" /index.html HTTP/1.1 200\n") -> bytes:
    """Lines that start with `head` at every offset 0..63 (mod 64), separated by filler lines of growing length."""
    out = bytearray()
    for align in range(64):
        while len(out) % 64 != align:
            pad = (align - len(out)) % 64
            out += b"x" * (pad - 1) + b"\n" if pad > 1 else b"\n"
        out += head + body
        out += b"not at the start: " + head + b"tail\n"
    return bytes(out)


@pytest.mark.parametrize("patterns", ANCHORED)
def test_fast_path_model_anchored_at_every_alignment(patterns, gpu_lib):
    """Regression (round-1 advisor finding): a ^-anchored match whose line starts exactly where the local walk starts
    (the '\\n' is the byte BEFORE the look-back window) must still be found, at every alignment of the line start."""
    db = CompiledDb(gpu_lib, patterns)
    assert db.rc == 0 and db.info.prefilter
    heads = {"^GET ": b"GET ", "^abcd": b"abcd", "^foo:": b"foo:", "^.abcdefg": b"Zabcdefg", r"\AGET /index": b"GET",
             r"^\s*warn: disk": b"  warn: disk"}
    data = anchored_alignment_text(heads[patterns[0]])
    truth, pos = set(), 0
    for pl in split_pseudo_lines(data, 262140):
        if db.line_reports(pl):
            truth.add(pos)
        pos += len(pl)
    assert len(truth) >= 64 or patterns[0].startswith(r"\A")
    assert fast_path_matched_line_starts(db, data) == truth


@pytest.mark.parametrize("seed", range(40))
def test_state_depth_bounds_the_age_of_partial_matches(seed, gpu_lib):
    """Dfa::depth (automata.hpp): nothing that is in progress in state s began more than depth[s] bytes ago.  Then a walk
    that starts only depth[s] bytes earlier - in the mid-line entry state that the byte in front of it selects - must
    arrive in the very same state.  (This is what lets a verification walk stop early, and what would break it.)"""
    rng = random.Random(1000 + seed)
    k = rng.choice([1, 2, 4])
    patterns = [_factor_pattern(rng) for _ in range(k)]
    db = CompiledDb(gpu_lib, patterns, [rng.choice([14, 14, 15, 10])] * k, [0] * k)
    assert db.rc == 0
    checked = bounded = 0
    for _ in range(30):
        text = "".join(rng.choice("abcAB _x1.\t") for _ in range(rng.randint(0, 40)))
        at = rng.randint(0, len(text))
        line = (text[:at] + "".join(rng.choice(LITS) for _ in range(rng.randint(1, 3))) + text[at:]).encode("latin1")
        for (gi, cls, trans, acc, _reports), depth in zip(db.groups, db.depths):
            s = 0
            for p, b in enumerate(line):
                s = int(trans[s, cls[b]])
                if s >= gi.first_accept or s == gi.dead:
                    break   # matched (absorbing in simple mode) or dead: nothing to bound
                d = int(depth[s])
                checked += 1
                if d == 255 or d > p + 1:
                    continue
                bounded += 1
                q = p + 1 - d   # the local walk consumes line[q .. p]
                local = 0 if q == 0 else (gi.entry_mid_word if chr(line[q - 1]).isalnum() or line[q - 1] == 95 else gi.entry_mid_other)
                for b2 in line[q:p + 1]:
                    local = int(trans[local, cls[b2]])
                assert local == s, (patterns, line, p, d)
    if checked and not bounded:
        pytest.skip("every state of this set sits behind a loop (depth 255)")
