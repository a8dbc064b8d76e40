"""End-to-end runs of the host-side layers above the C ABI in a container WITHOUT a GPU: the compiled driver (the gimic-b200 program,
and libgimic_b200_driver.so through the Python launcher) is pointed at a TEST DOUBLE of libgimic_b200.so (tests/mock_backend/mock_api.cpp, compiled into a temporary
directory, never in-tree) whose compute entry points are answered by the CPU oracle.

Checked here: (1) the program and the launcher path write byte-identical reports and files for every run mode (cdens closed / open
shell, ACID, property, integrals, edens / divj, scan, appended VTK); (2) the single-process multi-device partition (two host
threads, two contexts) reproduces the single-device run; (3) native driver + oracle backend reproduce the reference's own golden
outputs (c4h4 jvec.vtu at 10 digits, the integration stdout windows, the eight open-shell .vti files); (4) the launcher under
torch.distributed with two ranks over gloo (the driver's rank mode: row gather for cdens, all-reduce for integrals through the two
callbacks of gimic_b200_run_opts) writes what a single process writes.
This says nothing about the CUDA kernels: tests/test_gpu_*.py hold those to the oracle on a B200."""
import filecmp
import io
import os
import re
import shutil
import subprocess
import sys
import numpy as np
import pytest

import fixtures
import oracle_lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = fixtures.GOLD
INPUTS = os.path.join(GOLD, "inputs")
EXE = os.path.join(ROOT, "gimic_b200", "gimic-b200")
sys.path.insert(0, GOLD)


@pytest.fixture(scope="module")
def mock_dir(tmp_path_factory):
    import __graft_entry__ as ge
    from gimic_b200 import _lib
    if not (os.path.exists(_lib.SO_PATH) and os.path.exists(EXE)):
        ge.build()
    oracle_lib.lib()                                            # builds oracle/libgimic_oracle.so if needed
    d = tmp_path_factory.mktemp("mock_backend")
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-pthread", "-o", str(d / "libgimic_b200.so"),
                           os.path.join(ROOT, "tests", "mock_backend", "mock_api.cpp"), os.path.join(ROOT, "gimic_b200", "csrc", "host_basis.cpp"),
                           "-L" + os.path.join(ROOT, "oracle"), "-l:libgimic_oracle.so", "-Wl,-rpath," + os.path.join(ROOT, "oracle")])
    return d


def _env(mock_dir):
    e = dict(os.environ)
    e["LD_LIBRARY_PATH"] = str(mock_dir) + os.pathsep + e.get("LD_LIBRARY_PATH", "")     # searched before the program's RUNPATH ($ORIGIN)
    e["OMP_NUM_THREADS"] = "2"
    return e


def _native(mock_dir, args, timeout=900):
    p = subprocess.run([EXE, *[str(a) for a in args]], capture_output=True, text=True, timeout=timeout, env=_env(mock_dir))
    assert p.returncode == 0, p.stderr
    return fixtures.strip_clock(p.stdout)


PY_RUNNER = r'''
import io, os, sys
sys.path.insert(0, {root!r})
from gimic_b200 import _lib
_lib.SO_PATH = {so!r}                      # the test double instead of the CUDA library
assert "TEST DOUBLE" in _lib.lib().gimic_b200_version().decode()
from gimic_b200 import driver
args = sys.argv[1:]
appended = "--appended" in args
files = [a for a in args if not a.startswith("--")]
if len(files) > 1:
    driver.run_scan(files)
else:
    out = io.StringIO()
    driver.Driver(files[0], out=out, vtk_appended=appended).run()
    sys.stdout.write(out.getvalue())
'''


def _python(mock_dir, args, timeout=900):
    """the driver library through the Python launcher, in a fresh interpreter bound to the test double (this process keeps the real library)"""
    code = PY_RUNNER.format(root=ROOT, so=str(mock_dir / "libgimic_b200.so"))
    p = subprocess.run([sys.executable, "-c", code, *[str(a) for a in args]], capture_output=True, text=True, timeout=timeout, env=_env(mock_dir))
    assert p.returncode == 0, p.stderr[-2000:]
    return fixtures.strip_clock(p.stdout)


def _pair(tmp_path, name, mol, xdens, edit=None, extra=None):
    dirs = []
    for k in ("nat", "py"):
        d = tmp_path / k / name
        d.mkdir(parents=True)
        shutil.copy(mol, d / "MOL"); shutil.copy(xdens, d / "XDENS")
        txt = open(os.path.join(INPUTS, name + ".inp")).read()
        (d / "gimic.inp").write_text(edit(txt) if edit else txt)
        if extra:
            extra(d)
        dirs.append(d)
    return dirs


def _same_dirs(dn, dp):
    assert sorted(os.listdir(dn)) == sorted(os.listdir(dp))
    for f in sorted(os.listdir(dn)):
        if f.endswith(".out"):             # scan reports carry wall-clock times
            assert fixtures.strip_clock(open(dn / f).read()) == fixtures.strip_clock(open(dp / f).read()), f
        else:
            assert filecmp.cmp(dn / f, dp / f, shallow=False), f


def test_c4h4_read_grid_native_vs_python_and_golden(mock_dir, tmp_path, cases):
    """test/c4h4/read-grid through gimic-b200: jvec.vtu equals the launcher path's bytes and the reference's golden at its 10 digits"""
    from make_golden import read_vtu_vectors
    gold = fixtures.golden_npz("c4h4_readgrid.npz")

    def extra(d):
        np.savetxt(d / "gridfile.grd", gold["grid"], fmt="%.6f")
        with open(d / "grid.1.ele", "w") as f:
            f.write("3  4  0\n    1    1475  1730  1474  1717\n    2     100     8   112   245\n    3     5     6     7     8\n")
    dn, dp = _pair(tmp_path, "c4h4_read-grid", cases["c4h4"]["mol"], cases["c4h4"]["xdens"], extra=extra)
    a = _native(mock_dir, [dn / "gimic.inp"])
    b = _python(mock_dir, [dp / "gimic.inp"])
    assert a == b and "Closed-shell calculation" in a
    _same_dirs(dn, dp)
    pts, vec = read_vtu_vectors(str(dn / "jvec.vtu"))
    ref = gold["jvec"]
    assert pts.shape == vec.shape == (4110, 3) and np.allclose(pts, gold["grid"], atol=1e-9)
    big = np.abs(ref) > 1e-3 * np.abs(ref).max()
    assert (np.abs(vec - ref)[big] / np.abs(ref)[big]).max() < 4e-9
    assert np.abs(vec - ref).max() < 1e-9 * np.abs(ref).max() + 1e-14


@pytest.mark.parametrize("case,name,gold_json", [("c4h4", "c4h4_integration", "c4h4_integration.json"),
                                                  ("open_shell", "open-shell_integration", "open_shell_integration.json")])
def test_integration_reports_native_vs_python_and_golden(mock_dir, tmp_path, cases, case, name, gold_json):
    """test/c4h4/integration and test/open-shell/integration (36 x 36 Gauss plane, all spin cases): the gimic-b200 report is the
    launcher path's, and its numbers are the ones the reference printed"""
    dn, dp = _pair(tmp_path, name, cases[case]["mol"], cases[case]["xdens"])
    a = _native(mock_dir, [dn / "gimic.inp"])
    b = _python(mock_dir, [dp / "gimic.inp"])
    assert a == b
    _same_dirs(dn, dp)
    gold = fixtures.golden_json(gold_json)
    cur = [blk for blk in gold["blocks"] if blk["section"] == "current"]
    found = [float(x) for x in re.findall(r"Induced current \(au\)\s+:\s*([-\d.]+)", a)]
    assert len(found) == len(cur) and np.allclose(found, [blk["au"] for blk in cur], rtol=0, atol=1.01e-6)
    tail = a[a.index("*** Integrating current"):]                      # the window the reference's runtest compares
    pos = [float(x) for x in re.findall(r"Positive contribution:\s*([-\d.]+)", tail)]
    neg = [float(x) for x in re.findall(r"Negative contribution:\s*([-\d.]+)", tail)]
    assert np.allclose(pos, [blk["pos"] for blk in cur], rtol=0, atol=1.01e-6) and np.allclose(neg, [blk["neg"] for blk in cur], rtol=0, atol=1.01e-6)


def test_open_shell_3d_native_vs_python_and_golden(mock_dir, tmp_path, cases):
    """test/open-shell/3d: eight .vti files (UHF, J path, alpha/beta combined by linearity) on a 17^3 subgrid of the reference's 33^3
    grid (every second point, so the CPU oracle finishes in seconds): native == Python bytes, values == the reference's golden"""
    from make_golden import read_vti
    txt = open(os.path.join(INPUTS, "open-shell_3d.inp")).read()
    assert "spacing=[0.5, 0.5, 0.5]" in txt
    edit = lambda t: t.replace("spacing=[0.5, 0.5, 0.5]", "spacing=[1.0, 1.0, 1.0]")
    dn, dp = _pair(tmp_path, "open-shell_3d", cases["open_shell"]["mol"], cases["open_shell"]["xdens"], edit=edit)
    a = _native(mock_dir, [dn / "gimic.inp"])
    b = _python(mock_dir, [dp / "gimic.inp"])
    assert a == b and "Open-shell calculation" in a
    _same_dirs(dn, dp)
    assert all(os.path.exists(dn / f"jvec{t}.vti") and os.path.exists(dn / f"jmod{t}.vti") for t in ("", "alpha", "beta", "spindens"))
    if True:
        gold = fixtures.golden_npz("open_shell_3d.npz")
        idx = gold["index"]
        i, j, k = idx % 33, (idx // 33) % 33, idx // (33 * 33)
        on = (i % 2 == 0) & (j % 2 == 0) & (k % 2 == 0)                  # golden sample points that lie on the 17^3 subgrid
        sub = (i[on] // 2) + 17 * ((j[on] // 2) + 17 * (k[on] // 2))
        assert on.sum() > 50
        for tag in ("", "alpha", "beta", "spindens"):
            jv = read_vti(str(dn / f"jvec{tag}.vti"))
            gj = gold["jvec" + tag][on]
            assert (np.abs(jv[sub] - gj) <= 2e-6 * np.abs(gj) + 1e-9 * np.abs(gj).max()).all(), tag


def test_benzene_modes_native_vs_python(mock_dir, tmp_path, cases):
    """benzene inputs with synthetic densities (nbf = 252) on shrunken grids: the ACID / tensor path (3d), a Gauss bond grid in cdens
    mode (jmod.txt), the radius / rotation keywords, and the magnetizability input with prop=on (report + integrand plots)"""
    from gimic_b200.driver import mol_geometry as read_mol_geometry
    xd = tmp_path / "XDENS"
    fixtures.write_xdens(str(xd), fixtures.dens_to_colmajor(fixtures.synthetic_density(252, seed=21)))
    shrink3 = lambda t: re.sub(r"grid_points=\[\s*\d+\s*,\s*\d+\s*,\s*\d+\s*\]", "grid_points=[6,5,4]", t)
    shrink2 = lambda t: re.sub(r"grid_points=\[\s*\d+\s*,\s*\d+\s*,\s*0\s*\]", "grid_points=[9, 9, 0]", t)

    def prop_files(d):
        _, coords = read_mol_geometry(str(d / "MOL"))
        rng = np.random.default_rng(5)
        counts = rng.integers(6, 12, size=coords.shape[0])
        pts = np.vstack([coords[a] + rng.normal(scale=1.5, size=(c, 3)) for a, c in enumerate(counts)])
        np.savetxt(d / "gridfile.grd", pts, fmt="%.10f"); np.savetxt(d / "grid_w.grd", rng.uniform(0, 0.1, size=pts.shape[0]), fmt="%.12e")
        shutil.copy(os.path.join(GOLD, "benzene_coord.au"), d / "coord.au")
        np.savetxt(d / "nelpts.info", np.column_stack([np.arange(1, len(counts) + 1), counts]), fmt="%d")
        with open(d / "grid.1.ele", "w") as f:
            f.write("2  4  0\n    1    1  2  3  4\n    2     5     6     7     8\n")
    for name, edit, extra in (("benzene_3d", shrink3, None), ("benzene_int-cdens", shrink2, None), ("benzene_keyword-radius", shrink2, None),
                              ("benzene_keyword-rotation", shrink2, None), ("benzene_integration-lobatto", shrink2, None),
                              ("benzene_magnetizability", None, prop_files)):
        dn, dp = _pair(tmp_path, name, cases["benzene_mol"], xd, edit=edit, extra=extra)
        a = _native(mock_dir, [dn / "gimic.inp"])
        b = _python(mock_dir, [dp / "gimic.inp"])
        assert a == b, name
        _same_dirs(dn, dp)
        if name == "benzene_3d":
            assert os.path.exists(dn / "acid.vti") and os.path.exists(dn / "jmod.vti") and os.path.exists(dn / "jvec.vti")
        if name == "benzene_magnetizability":
            assert "isotropic magnetizability chi" in a and os.path.exists(dn / "intchi.vtu") and os.path.exists(dn / "sigma_zz12.vtu")


def test_scalar_modes_appended_vtk_and_scan_native_vs_python(mock_dir, tmp_path, cases):
    mol, xd = cases["c4h4"]["mol"], cases["c4h4"]["xdens"]
    for calc in ("edens", "divj"):
        def edit(txt, calc=calc):
            txt = txt.replace("calc=integral", "calc=" + calc)
            return re.sub(r"Grid\(bond\) \{.*?\n\}", "Grid(base) {\n type=even\n origin=[-4.0,-3.0,-1.0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n"
                          " lengths=[3.0,3.0,1.0]\n spacing=[0.5,0.5,0.5]\n}", txt, flags=re.S)
        dn, dp = _pair(tmp_path / calc, "c4h4_integration", mol, xd, edit=edit)
        assert _native(mock_dir, [dn / "gimic.inp"]) == _python(mock_dir, [dp / "gimic.inp"])
        _same_dirs(dn, dp)
        assert os.path.exists(dn / f"{calc}.vti")
    # --vtk appended on a small open-shell cdens grid
    edit = lambda t: t.replace("spacing=[0.5, 0.5, 0.5]", "spacing=[4.0, 4.0, 8.0]")
    dn, dp = _pair(tmp_path / "app", "open-shell_3d", cases["open_shell"]["mol"], cases["open_shell"]["xdens"], edit=edit)
    assert _native(mock_dir, ["--vtk", "appended", dn / "gimic.inp"]) == _python(mock_dir, ["--appended", dp / "gimic.inp"])
    _same_dirs(dn, dp)
    # scan: six slices of the c4h4 plane, one context, batched integrals; reports in gimic.N.out
    dn, dp = _pair(tmp_path / "scan", "c4h4_integration", mol, xd)
    base = open(dn / "gimic.inp").read()
    edges = np.linspace(-1.25614, 6.0, 7)
    names = {"nat": [], "py": []}
    for k in range(6):
        txt = base.replace("width=[-1.25614, 6.0]", f"width=[{edges[k]:.6f}, {edges[k + 1]:.6f}]").replace("grid_points=[30, 30, 0]", "grid_points=[9, 9, 0]")
        assert txt != base
        for key, d in (("nat", dn), ("py", dp)):
            (d / f"gimic.{k}.inp").write_text(txt)
            names[key].append(d / f"gimic.{k}.inp")
    for d in (dn, dp):
        (d / "calculation.dat").write_text("atom1=1\natom2=2\nin=0.0 out=7.0\ndelta=1.209357 nsteps=6\n")
    _native(mock_dir, names["nat"])
    _python(mock_dir, names["py"])
    # current_profile.dat (what jobscripts/src/gradient.sh.in pastes together from the gimic.N.out files): same bytes from the program and the launcher path,
    # and its columns are the numbers printed in the reports (there rounded to 6 decimals)
    assert filecmp.cmp(dn / "current_profile.dat", dp / "current_profile.dat", shallow=False)
    prof = np.loadtxt(dn / "current_profile.dat")
    assert prof.shape == (6, 4) and np.allclose(prof[:, 0], np.round(np.arange(6) * 1.209357, 2))
    for k in range(6):
        rep = open(dn / f"gimic.{k}.out").read()
        tail = rep[rep.index("*** Integrating current"):]
        si = float(re.search(r"Induced current \(nA/T\)\s+:\s*([-\d.]+)", tail).group(1))
        assert abs(prof[k, 1] - si) < 1.01e-6 and abs(prof[k, 1] - (prof[k, 2] + prof[k, 3])) < 1e-7
    for k in range(6):
        a = fixtures.strip_clock(open(dn / f"gimic.{k}.out").read())
        assert a == fixtures.strip_clock(open(dp / f"gimic.{k}.out").read()), k
        assert "wall time:" in a                                        # what the jobscripts grep for in a finished slice
        assert a == _native(mock_dir, [names["nat"][k]]), k             # what a separate run prints
        assert "Induced current (au)" in a


@pytest.mark.parametrize("case,name,edit", [("c4h4", "c4h4_integration", None), ("open_shell", "open-shell_3d", "shrink"), ("c4h4", "c4h4_read-grid", None)])
def test_multi_device_partition_equals_single_device(mock_dir, tmp_path, cases, case, name, edit):
    """--devices 0,1 / all (the test double reports two devices): one context and one host thread per device, point slabs / plane
    rows split like schedule() (parallel.F90:66-84).  The oracle evaluates every point independently, so files must be byte-identical
    to the single-device run; integrals agree to the last printed digit (partial sums are added in device order)."""
    gold = fixtures.golden_npz("c4h4_readgrid.npz")
    ed = (lambda t: t.replace("spacing=[0.5, 0.5, 0.5]", "spacing=[2.0, 2.0, 4.0]")) if edit else None

    def extra(d):
        if name == "c4h4_read-grid":
            np.savetxt(d / "gridfile.grd", gold["grid"][:501], fmt="%.6f")
    dn, dp = _pair(tmp_path, name, cases[case]["mol"], cases[case]["xdens"], edit=ed, extra=extra)
    one = _native(mock_dir, [dn / "gimic.inp"])
    two = _native(mock_dir, ["--devices", "0,1", dp / "gimic.inp"])
    if name == "c4h4_integration":
        na = np.array([float(x) for x in re.findall(r"[-+]?\d+\.\d+", one)]); nb = np.array([float(x) for x in re.findall(r"[-+]?\d+\.\d+", two)])
        assert re.sub(r"[-+]?\d+\.\d+", "#", one) == re.sub(r"[-+]?\d+\.\d+", "#", two) and np.allclose(na, nb, rtol=0, atol=1.01e-6)
    else:
        assert one == two
    _same_dirs(dn, dp)
    d3 = tmp_path / "all"
    shutil.copytree(dp, d3)
    for f in os.listdir(d3):
        if f not in ("MOL", "XDENS", "gimic.inp", "gridfile.grd"):
            os.remove(d3 / f)
    assert _native(mock_dir, ["--devices", "all", d3 / "gimic.inp"]) == two
    _same_dirs(dp, d3)


def _dist_worker(rank, world, port, so, inpfile, outfile):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), OMP_NUM_THREADS="2")
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from gimic_b200 import _lib
    _lib.SO_PATH = so                                           # the test double
    from gimic_b200 import driver
    dist.init_process_group("gloo", rank=rank, world_size=world)
    out = io.StringIO()
    driver.Driver(inpfile, out=out).run()
    if rank == 0:
        open(outfile, "w").write(out.getvalue())
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("case,name,edit", [("open_shell", "open-shell_3d", lambda t: t.replace("spacing=[0.5, 0.5, 0.5]", "spacing=[2.0, 2.0, 4.0]")),
                                            ("c4h4", "c4h4_integration", None), ("c4h4", "c4h4_read-grid", None)])
def test_launcher_world2_gloo_equals_single_process(mock_dir, tmp_path, cases, case, name, edit):
    """`torchrun -m gimic_b200 gimic.inp` semantics with two ranks over gloo (CPU): the compiled driver in rank mode (gimic_b200_run_opts
    rank / nranks): cdens evaluates each rank's share of the partition and gathers the rows on rank 0 through the launcher's
    allgather_rows callback (jfield.f90:90-137), integral mode splits the plane rows and adds the <= 7 sums through allreduce_sum
    (parallel.F90:66-84).  Rank 0 must write what a single process writes."""
    import torch.multiprocessing as mp
    gold = fixtures.golden_npz("c4h4_readgrid.npz")

    def extra(d):
        if name == "c4h4_read-grid":
            np.savetxt(d / "gridfile.grd", gold["grid"][:301], fmt="%.6f")
    d1, d2 = _pair(tmp_path, name, cases[case]["mol"], cases[case]["xdens"], edit=edit, extra=extra)
    single = _python(mock_dir, [d1 / "gimic.inp"])
    ctx = mp.get_context("spawn")
    port = 29500 + (os.getpid() * 7 + len(name)) % 2000
    rep = tmp_path / "rank0.out"
    procs = [ctx.Process(target=_dist_worker, args=(r, 2, port, str(mock_dir / "libgimic_b200.so"), str(d2 / "gimic.inp"), str(rep))) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    dist_out = fixtures.strip_clock(rep.read_text())
    if name == "c4h4_integration":
        num = r"[-+]?\d+\.\d+"
        assert re.sub(num, "#", single) == re.sub(num, "#", dist_out)
        assert np.allclose([float(x) for x in re.findall(num, single)], [float(x) for x in re.findall(num, dist_out)], rtol=0, atol=1.01e-6)
    else:
        assert single == dist_out
    _same_dirs(d1, d2)


_NUM = re.compile(r"[-+]?\d+\.\d+(?:[EeDd][-+]?\d+)?|[-+]?\d+")


def _normalised_reference_report(name, ours):
    """The reference's stdout for one of its tests (tests/golden/stdout/, from two lines above 'TITLE:' on) brought to what THIS commit of
    the reference would print, so that it can be compared line by line.  The goldens were written by 13 different revisions between 2017
    and 2020; what changed since the older ones (all verified against the current source):
      * GTO counts are printed with i6 instead of i4 (basis.f90:59-62): those two lines are compared by tokens only;
      * the |J| pass only runs with Essential.jmod=on, otherwise ' Jmod integration skipped.' (gimic.F90:228-243);
      * times carry a ' ( h )' suffix (timer.f90:23-40): dates and times are blanked on both sides;
      * the 'maxi, mini' lines of the old ACID / |J| plots are gone (no such print in jfield.f90);
      * one revision (the keyword-rotation_origin golden) printed a blank as its empty line.
    Left out on our side: the 'Estimated CPU time for single core calculation' note of calc_jtensors (+ its blank line) and the front end's
    closing 'This is F-GIMIC.'"""
    gold = fixtures.strip_clock(open(os.path.join(GOLD, "stdout", name + ".txt")).read()).split("\n")
    gold = [l for l in gold if l != "This is F-GIMIC." and "maxi, mini" not in l]
    if name == "benzene_keyword-rotation_origin":
        gold = [l if l.strip() else "" for l in gold]
    for k, l in enumerate(gold):
        if "Estimated CPU time" in l:
            del gold[k:k + 2]
            break
    if " Jmod integration skipped." in ours and " Jmod integration skipped." not in gold and " *** Integrating |J|" in gold:
        a = gold.index(" *** Integrating |J|"); b = gold.index(" *** Integrating current")
        gold[a:b] = [" Jmod integration skipped."]
    while gold and gold[-1] == "":
        gold.pop()
    return gold


REPORTS = [("c4h4", "c4h4_integration"), ("open_shell", "open-shell_integration")] + \
          [("benzene", f[:-4]) for f in sorted(os.listdir(os.path.join(GOLD, "stdout"))) if f.startswith("benzene_")]


@pytest.mark.parametrize("case,name", REPORTS)
def test_report_equals_the_reference_stdout_line_by_line(mock_dir, tmp_path, cases, case, name):
    """The complete report of gimic-b200 (oracle-backed test double) against what the reference printed for the same input -- all 17 stdout
    goldens of the reference tree: integrals with every grid keyword, ACID, |J|, diamag/paramag/GIAO switches, two cdens runs, the
    property run.  Same lines in the same order, identical text, identical LENGTH of every line (all formats are fixed-width).  For the two
    cases whose densities exist (c4h4, open-shell) every number must also agree within one unit of its last printed digit; the benzene
    densities are missing from the reference tree, so there the numbers come from synthetic densities and only the layout is compared."""
    from gimic_b200.driver import mol_geometry as read_mol_geometry
    if case == "benzene":
        xd = tmp_path / "XDENS"
        fixtures.write_xdens(str(xd), fixtures.dens_to_colmajor(fixtures.synthetic_density(252, seed=21)))
        mol = cases["benzene_mol"]
    else:
        mol, xd = cases[case]["mol"], cases[case]["xdens"]

    def extra(d):
        if name != "benzene_magnetizability":
            return
        _, coords = read_mol_geometry(str(d / "MOL"))
        rng = np.random.default_rng(5)
        counts = rng.integers(6, 12, size=coords.shape[0])
        pts = np.vstack([coords[a] + rng.normal(scale=1.5, size=(c, 3)) for a, c in enumerate(counts)])
        np.savetxt(d / "gridfile.grd", pts, fmt="%.10f"); np.savetxt(d / "grid_w.grd", rng.uniform(0, 0.1, size=pts.shape[0]), fmt="%.12e")
        shutil.copy(os.path.join(GOLD, "benzene_coord.au"), d / "coord.au")
        np.savetxt(d / "nelpts.info", np.column_stack([np.arange(1, len(counts) + 1), counts]), fmt="%d")
    dn, _ = _pair(tmp_path, name, mol, xd, extra=extra)
    ours = _native(mock_dir, [dn / "gimic.inp"]).split("\n")
    while ours and ours[-1] == "":
        ours.pop()
    gold = _normalised_reference_report(name, ours)
    assert len(ours) == len(gold), (len(ours), len(gold), [(x, y) for x, y in zip(ours, gold) if _NUM.sub("#", x).split() != _NUM.sub("#", y).split()][:3])
    for k, (x, y) in enumerate(zip(ours, gold)):
        if "Total number of" in x and "GTO's" in x:
            assert x.split() == y.split(), (k, x, y)
            continue
        assert _NUM.sub("#", x).split() == _NUM.sub("#", y).split() or (_NUM.sub("#", x).replace("(#", "( #").split() == _NUM.sub("#", y).replace("(#", "( #").split()), (k, x, y)
        if "Total number of grid points" in x and name == "benzene_magnetizability":
            continue                                                    # the stand-in point set is smaller than the reference's NumGrid file
        assert len(x) == len(y), (k, x, y)
        if case != "benzene":
            for p, q in zip(_NUM.findall(x), _NUM.findall(y)):
                unit = 10.0 ** -len(q.split(".")[1]) if "." in q and "E" not in q.upper() else (0.0 if "." not in q else 1e-4 * abs(float(q)))
                assert abs(float(p) - float(q)) <= 1.01 * unit, (k, x, y)
