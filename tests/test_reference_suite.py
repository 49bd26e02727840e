"""The reference's OWN, unmodified test-suite with its loader hook redirected to another native library.

The reference's Python layer and test module are read from /root/reference in the build container, and from the copy
that `make -C oracle ref` (run by __graft_entry__.build()) stages under oracle/_ref/ - git-ignored, but shipped to the
GPU box - everywhere else.  The CPU legs pin the oracle and the host logic to every golden case the reference holds
for the scan path; the GPU leg is the drop-in acceptance test for libgpugrep.so: the reference's own unmodified
utils.py / multiscanner.py / test_hypergrep.py on top of the new library.
"""

import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("GPUGREP_REFERENCE", "/root/reference")
if not os.path.exists(os.path.join(REFERENCE, "hypergrep", "test", "test_hypergrep.py")):
    REFERENCE = os.path.join(ROOT, "oracle", "_ref")
SUITE = os.path.join(REFERENCE, "hypergrep", "test", "test_hypergrep.py")

needs_reference = pytest.mark.skipif(not os.path.exists(SUITE), reason="reference checkout not present")


def _run(library_path: str) -> None:
    env = dict(os.environ)
    env["GPUGREP_INJECT_LIB"] = library_path
    env["PYTHONPATH"] = os.pathsep.join([REFERENCE, os.path.join(ROOT, "tests")])
    proc = subprocess.run(
        [sys.executable, "-m", "pytest", SUITE, "-p", "_inject_plugin", "-p", "no:cacheprovider", "-q", "-x"],
        cwd="/tmp", env=env, capture_output=True, text=True, timeout=600, check=False)
    assert proc.returncode == 0, proc.stdout[-3000:] + proc.stderr[-2000:]
    assert "75 passed" in proc.stdout, proc.stdout[-500:]


@needs_reference
def test_oracle_passes_reference_suite(oracle_lib):
    _run(oracle_lib._name)


@needs_reference
def test_host_logic_passes_reference_suite(hostmock_lib):
    _run(hostmock_lib._name)


@needs_reference
@pytest.mark.gpu
def test_libgpugrep_passes_reference_suite(gpu_lib):
    _run(gpu_lib._name)
