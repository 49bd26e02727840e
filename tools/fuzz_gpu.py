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
for seed in range(first, last):
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
