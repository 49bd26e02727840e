"""Seeded random regex / text generators for the differential tests (test infrastructure)."""

from __future__ import annotations

import random

ATOMS = ["a", "b", "c", "A", " ", "_", ".", "[ab]", "[^a]", "[a-c]", "\\w", "\\s", "\\d", "\\W", "x", "1", "[[:alpha:]]", "\\.", "[^\\n]"]
ANCHORS = ["^", "$", "\\b", "\\B", "\\A", "\\z", "\\Z"]


def gen_regex(rng: random.Random, depth: int = 0) -> str:
    r = rng.random()
    if depth > 3 or r < 0.35:
        parts = [rng.choice(ATOMS) for _ in range(rng.randint(1, 4))]
        return "".join(parts)
    if r < 0.5:
        return gen_regex(rng, depth + 1) + gen_regex(rng, depth + 1)
    if r < 0.65:
        return "(?:" + gen_regex(rng, depth + 1) + "|" + gen_regex(rng, depth + 1) + ")"
    if r < 0.85:
        q = rng.choice(["?", "*", "+", "{2}", "{1,3}", "{2,}", "{0,2}", "+?", "*?"])
        inner = gen_regex(rng, depth + 1)
        return "(" + inner + ")" + q
    if r < 0.95:
        return rng.choice(ANCHORS) + gen_regex(rng, depth + 1)
    return gen_regex(rng, depth + 1) + rng.choice(["$", "\\b", "\\z", "\\Z", "\\B"])


def gen_text(rng: random.Random, lines: int, alphabet: str = "abcAB _x1.\t", max_len: int = 24, nul_rate: float = 0.0) -> bytes:
    out = bytearray()
    for _ in range(lines):
        n = rng.randint(0, max_len)
        for _ in range(n):
            if nul_rate and rng.random() < nul_rate:
                out.append(0)
            else:
                out.append(ord(rng.choice(alphabet)))
        out.append(10)
    if rng.random() < 0.3 and out:
        out.pop()  # no trailing newline
    return bytes(out)
