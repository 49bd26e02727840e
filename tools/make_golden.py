#!/usr/bin/env python3
"""Extract the reference's own golden vectors for the scan path into tests/golden/reference_cases.json.

Run in the build container (needs /root/reference):  python tools/make_golden.py
It imports the reference's test module (hypergrep/test/test_hypergrep.py) only to read its TEST_CASES table and
its fixture files; nothing from the reference is executed natively (its libhs blob is missing).  The output holds
  - the fixture files (base64) the cases read,
  - every `scan`, `grep`, `parallel_grep` and `check_hyperscan_compatibility` case: args, kwargs, expected value.
tests/test_golden.py replays them against the oracle (CPU) and against libgpugrep.so (GPU).
"""

import base64
import json
import os
import sys

REF = os.environ.get("GPUGREP_REFERENCE", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(0, os.path.join(REF, "hypergrep", "test"))

import test_hypergrep as ref_tests  # noqa: E402  pylint: disable=wrong-import-position

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TEST_ROOT = ref_tests.TEST_ROOT


def encode(value):
    """JSON-encode args: fixture paths become {"fixture": name}; callables become their name."""
    if isinstance(value, str):
        if value.startswith(TEST_ROOT):
            rel = value[len(TEST_ROOT):].lstrip("/")
            return {"fixture": rel}
        return value
    if isinstance(value, (list, tuple)):
        return [encode(v) for v in value]
    if isinstance(value, dict):
        return {k: encode(v) for k, v in value.items()}
    if callable(value):
        return {"callable": value.__name__}
    return value


def main() -> None:
    fixtures = {}
    for name in sorted(os.listdir(TEST_ROOT)):
        path = os.path.join(TEST_ROOT, name)
        if os.path.isfile(path) and (name.startswith("greptest") or name.startswith("samplefile")):
            with open(path, "rb") as handle:
                fixtures[name] = base64.b64encode(handle.read()).decode()
    cases = {}
    for func in ("scan", "grep", "parallel_grep", "check_hyperscan_compatibility", "to_basic_regular_expressions",
                 "to_gnu_regular_expressions"):
        out = {}
        for case_name, case in ref_tests.TEST_CASES[func].items():
            entry = {"args": encode(case.get("args", [])), "kwargs": encode(case.get("kwargs", {}))}
            if "raises" in case:
                entry["raises"] = case["raises"].__name__
            else:
                entry["returns"] = encode(case["returns"])
            out[case_name] = entry
        cases[func] = out
    doc = {
        "source": "pyranha-labs/hypergrep v3.2.0 hypergrep/test/test_hypergrep.py TEST_CASES (scan/grep/parallel_grep/check)",
        "generator": "tools/make_golden.py",
        "fixtures": fixtures,
        "cases": cases,
    }
    out_path = os.path.join(ROOT, "tests", "golden", "reference_cases.json")
    with open(out_path, "w", encoding="utf-8") as handle:
        json.dump(doc, handle, indent=1, sort_keys=True)
    print(out_path, {k: len(v) for k, v in cases.items()})


if __name__ == "__main__":
    main()
