#!/usr/bin/env python3
"""One-off stress run: random parity cases (tests/parity.py) for a range of seeds against the oracle, on the GPU box."""
import ctypes
import os
import sys
import traceback

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import parity  # noqa: E402
from hypergrep_b200 import utils  # noqa: E402
from oracle_api import load_oracle  # noqa: E402

first, last = int(sys.argv[1]), int(sys.argv[2])
gpu = utils._get_hyperscanner_lib()
oracle = load_oracle()
bad = skipped = 0
for seed in (range(first, last) if not (len(sys.argv) > 3 and sys.argv[3] == "text") else ()):
    patterns, flags, ids, buffer_size, data, buffer_count, max_match = parity.random_case(seed)
    if parity.has_all_nul_pseudo_line(data, buffer_size):
        skipped += 1
        continue
    try:
        parity.compare(gpu, oracle, data, patterns, flags, ids, buffer_size, buffer_count, max_match)
    except Exception:  # pylint: disable=broad-except
        bad += 1
        print(f"seed {seed} FAILED: patterns={patterns!r} flags={flags} ids={ids} buffer_size={buffer_size}")
        traceback.print_exc(limit=2)
        if bad >= 5:
            break
print(f"seeds {first}..{last}: {bad} failures, {skipped} skipped")


def text_mode(first: int, last: int) -> None:
    """Random subsets of the benchmark pattern sets over 0.5-3 MiB of synthetic syslog / JSON-ish text, random
    buffer sizes and segment sizes: exercises segment cuts, dense candidate lists, several DFA groups, long lines."""
    import random

    from hypergrep_b200 import synth

    c3, plants = synth.c3_patterns()
    c5 = synth.c5_patterns(400)
    pool = synth.C2_PATTERNS + synth.C1_PATTERNS + c3[:300] + c5[:200] + [r"\d{3,5} in lib\w+", r"(?i)failed \w+ for", r"^\w{3} +\d+ ", r"ssh2$"]
    failures = 0
    for seed in range(first, last):
        rng = random.Random(seed)
        k = rng.choice([1, 1, 2, 3, 8, 32, 120])
        patterns = rng.sample(pool, k)
        flags = [rng.choice([14, 14, 15])] * k
        size = rng.choice([1 << 19, 1 << 20, 3 << 20])
        if rng.random() < 0.3:
            data = synth.jsonish_bytes(size // 2, seed=seed, patterns_to_plant=["session_4242 failed", "code=E31337abcd"], plant_rate=0.3)
        else:
            data = synth.syslog_bytes(size, seed=seed, plants=plants[:50] if rng.random() < 0.5 else None, plant_ppm=20000, lib=gpu)
        os.environ["GPUGREP_CHUNK_BYTES"] = str(rng.choice([1, 1, 300000, 1 << 20]))   # test hook: tiny segments
        buffer_size = rng.choice([262140, 262140, 4096, 300, 64])
        try:
            parity.compare(gpu, oracle, data, patterns, flags=flags, buffer_size=buffer_size, buffer_count=rng.choice([16, 7, 1000]),
                           max_match_count=rng.choice([0, 0, 5, 1000]))
        except Exception:  # pylint: disable=broad-except
            failures += 1
            print(f"text seed {seed} FAILED: k={k} buffer_size={buffer_size} chunk={os.environ['GPUGREP_CHUNK_BYTES']} patterns={patterns[:4]!r}...")
            traceback.print_exc(limit=2)
            if failures >= 5:
                break
    print(f"text seeds {first}..{last}: {failures} failures")


if len(sys.argv) > 3 and sys.argv[3] == "text":
    text_mode(first, last)
