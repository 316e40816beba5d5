"""GPU: the reference's integration tests run through the C++ host mirror
(include/oxmpl_cuda.hpp -> liboxg C ABI); see tests/cpp/host_mirror_test.cpp."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BIN = os.path.join(ROOT, "tests", "cpp", "host_mirror_test")


@pytest.mark.gpu
def test_reference_integration_tests_through_host_mirror():
    import __graft_entry__ as g

    g.build_host_mirror_test()
    r = subprocess.run([BIN], capture_output=True, text=True, timeout=300)
    print(r.stdout)
    print(r.stderr)
    assert r.returncode == 0, r.stdout[-4000:]
    assert "ALL OK" in r.stdout
