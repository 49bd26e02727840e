"""Reference golden vectors replayed on CPU: pins the ORACLE and the HOST LOGIC (Python layer + capi.cpp + compiler
+ ingest on the mock engine) to the reference's own expected outputs (reference hypergrep/test/test_hypergrep.py)."""

import json
import os

import pytest

from golden_replay import replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_cases.json"), encoding="utf-8") as _handle:
    _CASES = json.load(_handle)["cases"]
_PARAMS = [(func, name) for func, cases in _CASES.items() for name in cases]


@pytest.mark.parametrize("func,name", _PARAMS, ids=[f"{f}:{n}" for f, n in _PARAMS])
def test_oracle_matches_reference_golden(func, name, oracle_lib, fixture_dir, monkeypatch, capsys):
    replay(func, _CASES[func][name], oracle_lib, fixture_dir, monkeypatch, capsys)


@pytest.mark.parametrize("func,name", _PARAMS, ids=[f"{f}:{n}" for f, n in _PARAMS])
def test_host_logic_matches_reference_golden(func, name, hostmock_lib, fixture_dir, monkeypatch, capsys):
    replay(func, _CASES[func][name], hostmock_lib, fixture_dir, monkeypatch, capsys)
