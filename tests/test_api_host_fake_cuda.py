"""Host-side orchestration of every C-ABI entry point, executed in a container without a GPU: the library's own objects (api.cu,
host_basis.cpp, the nvcc launch stubs of the kernels) are linked with tests/fake_cudart/api_harness.cpp against a FAKE CUDA runtime
(tests/fake_cudart/fake_cudart.cpp: zeroed host memory as device memory, kernel launches are no-ops that report success).  Nothing is
computed -- every result is zero; what is checked is that contexts, staging, tile / batch bookkeeping, quadrature and property drivers,
the legacy symbols and all error paths run and return the documented codes.  Built into a temporary directory; test infrastructure only.
(tools/sanitize_api_host.sh runs the same harness under ASan / UBSan / LSan.)"""
import os
import subprocess
import pytest

import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "gimic_b200", "csrc")


def test_every_entry_point_runs_on_the_fake_cuda_runtime(tmp_path, cases):
    import __graft_entry__ as ge
    objs = [os.path.join(CSRC, f) for f in ("api.o", "host_basis.o", "k_prepare.o", "k_jtensor.o", "k_fields.o")]
    if not all(os.path.exists(o) for o in objs):
        ge.build()
    if not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("CUDA headers not installed")
    fake = tmp_path / "fakecuda"
    fake.mkdir()
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", str(fake / "libcudart.so.12"),
                           os.path.join(ROOT, "tests", "fake_cudart", "fake_cudart.cpp")])
    os.symlink("libcudart.so.12", fake / "libcudart.so")
    exe = tmp_path / "api_harness"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "fake_cudart", "api_harness.cpp"), *objs,
                           "-I" + os.path.join(ROOT, "include"), "-L" + str(fake), "-lcudart", "-lpthread"])
    env = dict(os.environ, LD_LIBRARY_PATH=str(fake))          # the fake runtime must win over any real libcudart.so.12 on the loader path
    p = subprocess.run([str(exe), cases["c4h4"]["mol"], cases["c4h4"]["xdens"], cases["open_shell"]["mol"], cases["open_shell"]["xdens"]],
                       capture_output=True, text=True, timeout=300, env=env)
    assert p.returncode == 0 and "api harness: 0 failure(s)" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]
    # and with no device reported, creation fails loudly instead of computing anything
    q = subprocess.run([str(exe), cases["c4h4"]["mol"], cases["c4h4"]["xdens"], cases["open_shell"]["mol"], cases["open_shell"]["xdens"]],
                       capture_output=True, text=True, timeout=300, env=dict(env, FAKE_CUDA_NO_DEVICE="1"))
    assert q.returncode == 3 and "create refused: no CUDA device available" in q.stdout
