"""Rehearsal of the driver-level GPU tests' CODE in a container without a GPU: tests/test_gpu_driver.py and tests/test_native_driver_gpu.py
are run against the oracle-backed TEST DOUBLE of libgimic_b200.so (tests/mock_backend/, built into a temporary directory).  This checks
the tests themselves (paths, regular expressions, golden comparisons, the driver's orchestration) -- it is NOT a GPU result: on a B200
the same tests run against the CUDA library.  Usage: python tools/rehearse_gpu_driver_tests.py"""
import os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
if os.environ.get("GIMIC_REHEARSAL") != "1":
    subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
    d = tempfile.mkdtemp(prefix="gimic_mock_")
    so = os.path.join(d, "libgimic_b200.so")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", so, os.path.join(ROOT, "tests", "mock_backend", "mock_api.cpp"),
                           os.path.join(ROOT, "gimic_b200", "csrc", "host_basis.cpp"), "-L" + os.path.join(ROOT, "oracle"), "-l:libgimic_oracle.so",
                           "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    # re-exec with the loader path of the gimic-b200 program pointing at the test double
    env = dict(os.environ, GIMIC_REHEARSAL="1", LD_LIBRARY_PATH=d + os.pathsep + os.environ.get("LD_LIBRARY_PATH", ""), GIMIC_MOCK_SO=so)
    sys.exit(subprocess.call([sys.executable, os.path.abspath(__file__)], env=env))
from gimic_b200 import _lib
_lib.SO_PATH = os.environ["GIMIC_MOCK_SO"]
assert "TEST DOUBLE" in _lib.lib().gimic_b200_version().decode()
import pytest
sys.exit(pytest.main(["-q", "-m", "gpu", os.path.join(ROOT, "tests", "test_gpu_driver.py"), os.path.join(ROOT, "tests", "test_native_driver_gpu.py"),
                      "-p", "no:cacheprovider"]))
