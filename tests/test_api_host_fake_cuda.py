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
    # drain groups forced onto the small point sets of the harness (the emulated plan takes the chunk size from the environment): a call
    # with host outputs launches the contraction group by group and the groups cover every batch
    for chunk in ("30000", "200000"):
        g = subprocess.run(p.args, capture_output=True, text=True, timeout=300, env=dict(env, FAKE_DRAIN_CHUNK=chunk, GIMIC_B200_SLICES="0"))   # few tiles: no column slices here
        assert g.returncode == 0 and "api harness: 0 failure(s)" in g.stdout, (chunk, g.stdout[-2000:] + g.stderr[-2000:])
    # and with no device reported, creation fails loudly instead of computing anything
    q = subprocess.run([str(exe), cases["c4h4"]["mol"], cases["c4h4"]["xdens"], cases["open_shell"]["mol"], cases["open_shell"]["xdens"]],
                       capture_output=True, text=True, timeout=300, env=dict(env, FAKE_CUDA_NO_DEVICE="1"))
    assert q.returncode == 3 and "create refused: no CUDA device available" in q.stdout


WRAPPER_SCRIPT = r'''
import sys, numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {tests!r})
import gimic_b200, fixtures
mol, xd, molu, xdu = sys.argv[1:5]
g = gimic_b200.Gimic(mol, xd, screening_thrs=1e-8)
assert (g.nbf, g.natoms, g.uhf) == (168, 8, False)
r = np.random.default_rng(0).normal(size=(1000, 3))
g.set_profiling(True)
t = g.jtensors(r)
st = g.stats()
assert t.shape == (1000, 9) and st["n_points"] == 1000 and st["n_tiles"] >= 8 and st["launches"] > 0
assert st["sum_nact"] > 0 and 0 < st["useful_flops"] <= st["executed_flops"] <= st["dense_flops"] * 128      # emulated tile counts: real bookkeeping
f = g.fields(r, [0, 0, 1.0], "total", tens=True, jvec=True, jmod=True, acid=True, edens=True, divj=True)
assert {{k: v.shape for k, v in f.items()}} == dict(tens=(1000, 9), jvec=(1000, 3), jmod=(1000,), acid=(1000,), edens=(1000,), divj=(1000,))
assert g.fields(r, [0, 0, 1.0], jvec=True, jmod=True)["jvec"].shape == (1000, 3)                     # J path
assert g.fields_from_tensors(r, t, [0, 0, 1.0], jvec=True, jmod=True, acid=True)["acid"].shape == (1000,)
bf, dr = g.basis(r[:5])
assert bf.shape == (5, 168) and dr.shape == (5, 3, 168) and g.jmod_from_jvec(r, f["jvec"], [0, 0, 1.0]).shape == (1000,)
xyz = g.atom_coords()
p0, w0, p1, w1 = np.zeros(9), np.zeros(9), np.zeros(9), np.zeros(9)
gimic_b200.gausspoints(0.0, 4.0, 9, p0, w0); gimic_b200.gausspoints(0.0, 4.0, 9, p1, w1)
gr = gimic_b200.Grid(xyz[0] - [2.0, 1.0, 0.0], np.eye(3), [p0, p1, np.zeros(1)], [w0, w1, np.ones(1)])      # a 9 x 9 Gauss plane
assert g.integrate(gr, [0, 0, 1.0], "total", 7).shape == (7,) and g.integrate_batch([gr, gr, gr], [0, 0, 1.0]).shape == (3, 7)
assert g.jtensors_grid(gr).shape == (81, 9) and g.jtensors_grid(gr, 10, 30).shape == (20, 9)
res = g.property(r, np.full(1000, 0.01), t, xyz, [400, 600])
assert res["sigma"].shape == (8, 3) and res["chi_atoms"].shape == (2, 3) and g.property_integrand(r, t, xyz[0]).shape == (1000, 4)
assert g.jtensor(r[0]).shape == (9,) and len(g.jvector(r[0])) == 3
for bad in ("beta", "spindens"):
    try:
        g.jtensors(r, bad); raise SystemExit("closed-shell context accepted " + bad)
    except gimic_b200.GimicB200Error as e:
        assert e.code == -4
g.close()
u = gimic_b200.Gimic(molu, xdu, uhf=True, screening_thrs=1e-8)
for sc in ("alpha", "beta", "total", "spindens"):
    assert u.jtensors(r[:300], sc).shape == (300, 9)
sh, da, nbf = fixtures.synthetic_case(5, "flake", seed=5)
a = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fixtures.dens_to_colmajor(da), dens_beta=fixtures.dens_to_colmajor(da[::-1].copy()), **sh)
assert (a.nbf, a.natoms, a.uhf) == (nbf, 5, True) and a.jtensors(r[:200], "spindens").shape == (200, 9)
a.close()
assert "torch" not in sys.modules
print("python wrapper ok")
'''


def test_python_wrapper_signatures_on_the_fake_cuda_runtime(tmp_path, cases):
    """gimic_b200.Gimic (ctypes over include/gimic_b200.h) against the REAL libgimic_b200.so whose libcudart is replaced by the fake runtime:
    every wrapper method is called once -- argument marshalling, the Stats / Opts / Grid struct layouts and the error codes must match the
    library (results are zeros: kernels are no-ops).  Runs in a fresh interpreter that never imports torch, so that the fake runtime is the
    only libcudart.so.12 in the process."""
    import __graft_entry__ as ge
    import sys
    from gimic_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        ge.build()
    if not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("CUDA headers not installed")
    fake = tmp_path / "fakecuda"
    fake.mkdir()
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-shared", "-fPIC", "-I/usr/local/cuda/include", "-o", str(fake / "libcudart.so.12"),
                           os.path.join(ROOT, "tests", "fake_cudart", "fake_cudart.cpp")])
    code = WRAPPER_SCRIPT.format(root=ROOT, tests=os.path.join(ROOT, "tests"))
    p = subprocess.run([sys.executable, "-c", code, cases["c4h4"]["mol"], cases["c4h4"]["xdens"], cases["open_shell"]["mol"], cases["open_shell"]["xdens"]],
                       capture_output=True, text=True, timeout=300, env=dict(os.environ, LD_LIBRARY_PATH=str(fake)))
    assert p.returncode == 0 and "python wrapper ok" in p.stdout, p.stdout[-1500:] + p.stderr[-3000:]


def test_plan_helper_properties(tmp_path):
    """slice_width / piece_cost / drain_group (kernels.cuh, shared by the plan kernels, the contraction and the host): slices cover a tile's
    columns exactly once in multiples of 32, follow the tile's cost and never exceed the slot count; costs are monotone; drain groups are
    ordered along a batch (tests/fake_cudart/plan_props.cpp, 20 000 random tiles)."""
    if not os.path.isdir("/usr/local/cuda/include"):
        pytest.skip("CUDA headers not installed")
    exe = tmp_path / "plan_props"
    subprocess.check_call(["g++", "-O1", "-std=c++17", "-I/usr/local/cuda/include", "-o", str(exe), os.path.join(ROOT, "tests", "fake_cudart", "plan_props.cpp")])
    p = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "plan props: 0 failure(s)" in p.stdout, p.stdout[-2000:]
