"""Shared fixtures.  `gpu` marks tests that need a B200; everything else runs on CPU."""

from __future__ import annotations

import base64
import ctypes
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GPU_LIB = os.path.join(ROOT, "hypergrep_b200", "lib", "libgpugrep.so")
HOSTMOCK_LIB = os.path.join(ROOT, "tests", "_build", "libgpugrep_hostmock.so")
ORACLE_LIB = os.path.join(ROOT, "oracle", "_build", "liboracle_hyperscanner.so")


def pytest_configure(config: pytest.Config) -> None:
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu on the GPU box")


def _make(directory: str) -> None:
    subprocess.check_call(["make", "-C", directory], stdout=subprocess.DEVNULL)


@pytest.fixture(scope="session")
def oracle_lib() -> ctypes.CDLL:
    if not os.path.exists(ORACLE_LIB):
        _make(os.path.join(ROOT, "oracle"))
    return ctypes.CDLL(ORACLE_LIB)


@pytest.fixture(scope="session")
def hostmock_lib() -> ctypes.CDLL:
    """Host logic of the boundary linked against the CPU mock of the CUDA engine (tests/mock_engine)."""
    _make(os.path.join(ROOT, "tests", "mock_engine"))
    return ctypes.CDLL(HOSTMOCK_LIB)


@pytest.fixture(scope="session")
def gpu_lib() -> ctypes.CDLL:
    if not os.path.exists(GPU_LIB):
        _make(os.path.join(ROOT, "hypergrep_b200", "csrc"))
    return ctypes.CDLL(GPU_LIB)


@pytest.fixture(scope="session")
def golden() -> dict:
    with open(os.path.join(ROOT, "tests", "golden", "reference_cases.json"), encoding="utf-8") as handle:
        return json.load(handle)


@pytest.fixture(scope="session")
def fixture_dir(golden: dict, tmp_path_factory: pytest.TempPathFactory) -> str:
    """The reference's fixture files, materialised from the golden JSON."""
    directory = tmp_path_factory.mktemp("fixtures")
    for name, payload in golden["fixtures"].items():
        (directory / name).write_bytes(base64.b64decode(payload))
    return str(directory)
