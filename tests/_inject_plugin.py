"""pytest plugin used ONLY by tests/test_reference_suite.py.

Runs the reference's own, unmodified test-suite (/root/reference/hypergrep/test/test_hypergrep.py) with the
reference's loader hook (hypergrep/utils.py:67-81, `_get_hyperscanner_lib`) redirected to a C-ABI library
named by $GPUGREP_INJECT_LIB: either the CPU oracle (pins the oracle to the reference's golden vectors) or
libgpugrep.so (proves the drop-in boundary).  Test infrastructure; never imported by the product.
"""

import ctypes
import os

import pytest

_LIB = None


def _lib() -> ctypes.CDLL:
    global _LIB  # pylint: disable=global-statement
    if _LIB is None:
        _LIB = ctypes.CDLL(os.environ["GPUGREP_INJECT_LIB"])
    return _LIB


@pytest.fixture(autouse=True)
def _inject_native_lib(monkeypatch: pytest.MonkeyPatch) -> None:
    from hypergrep import utils  # the REFERENCE package (PYTHONPATH=/root/reference)

    monkeypatch.setattr(utils, "_get_hyperscanner_lib", _lib)
