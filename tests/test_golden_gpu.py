"""Reference golden vectors replayed through libgpugrep.so on the GPU (the drop-in acceptance suite)."""

import json
import os

import pytest

from golden_replay import replay

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
with open(os.path.join(ROOT, "tests", "golden", "reference_cases.json"), encoding="utf-8") as _handle:
    _CASES = json.load(_handle)["cases"]
_PARAMS = [(func, name) for func, cases in _CASES.items() for name in cases]


@pytest.mark.gpu
@pytest.mark.parametrize("func,name", _PARAMS, ids=[f"{f}:{n}" for f, n in _PARAMS])
def test_gpu_matches_reference_golden(func, name, gpu_lib, fixture_dir, monkeypatch, capsys):
    replay(func, _CASES[func][name], gpu_lib, fixture_dir, monkeypatch, capsys)
