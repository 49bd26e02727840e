"""Replay of the reference's golden cases (tests/golden/reference_cases.json) through hypergrep_b200's Python layer
on top of an arbitrary native library with the reference ABI (the oracle, the host mock, or libgpugrep.so)."""

from __future__ import annotations

import builtins
import ctypes
import os

import pytest

import hypergrep_b200
from hypergrep_b200 import multiscanner, utils


def _decode(value, fixture_dir):
    if isinstance(value, dict):
        if "fixture" in value:
            return os.path.join(fixture_dir, value["fixture"]) if value["fixture"] else fixture_dir
        if "callable" in value:
            return _print_callback
        return {k: _decode(v, fixture_dir) for k, v in value.items()}
    if isinstance(value, list):
        return [_decode(v, fixture_dir) for v in value]
    return value


def _print_callback(matches, count):
    for index in range(count):
        match = matches[index]
        print(f"{match.line_number}:{match.line.decode(errors='ignore').rstrip()}")


def _normalise(value):
    if isinstance(value, tuple):
        return [_normalise(v) for v in value]
    if isinstance(value, list):
        return [_normalise(v) for v in value]
    return value


def replay(func: str, case: dict, lib: ctypes.CDLL, fixture_dir: str, monkeypatch: pytest.MonkeyPatch, capsys) -> None:
    monkeypatch.setattr(utils, "_get_hyperscanner_lib", lambda: lib)
    args = _decode(case["args"], fixture_dir)
    kwargs = _decode(case["kwargs"], fixture_dir)

    def call():
        if func == "scan":
            hypergrep_b200.scan(*args, **kwargs)
            return capsys.readouterr().out.splitlines()
        if func == "grep":
            return hypergrep_b200.grep(*args, **kwargs)
        if func == "check_hyperscan_compatibility":
            return hypergrep_b200.check_compatibility(*args, **kwargs)
        if func == "parallel_grep":
            code = multiscanner.parallel_grep(*args, **kwargs)
            lines = [line.replace(f"{fixture_dir}/", "") for line in capsys.readouterr().out.splitlines()]
            return lines, code
        return getattr(multiscanner, func)(*args, **kwargs)

    if "raises" in case:
        with pytest.raises(getattr(builtins, case["raises"])):
            call()
    else:
        assert _normalise(call()) == _normalise(case["returns"])
