"""The `gimic-b200` program on the GPU: it must write the same report and the same files as the driver library called through the
Python launcher (gimic_b200.driver.Driver -> gimic_b200_run), which tests/test_gpu_driver.py pins against the reference's goldens and
the oracle; plus the reference's goldens straight from the program, the scan mode and the single-process multi-device partition.
Runs last among the GPU tests (file name order)."""
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

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = fixtures.GOLD
INPUTS = os.path.join(GOLD, "inputs")
EXE = os.path.join(ROOT, "gimic_b200", "gimic-b200")
sys.path.insert(0, GOLD)
NUM = re.compile(r"[-+]?(?:\d+\.\d*|\.\d+|\d+)(?:[eEdD]?[-+]\d+)?")


@pytest.fixture(scope="module", autouse=True)
def _built():
    if not os.path.exists(EXE):
        import __graft_entry__ as ge
        ge.build()


def _unit_last_digit(tok):
    """value of one unit in the last printed digit of a numeric token (Fortran E/D/F forms, plain integers)"""
    t = tok.replace("D", "E").replace("d", "e").replace("e", "E")
    mant, _, ex = t.partition("E")
    if not ex and len(mant) > 1 and (mant.rfind("+") > 0 or mant.rfind("-") > 0):      # gfortran drops the E: 0.123456-100
        k = max(mant.rfind("+"), mant.rfind("-"))
        mant, ex = mant[:k], mant[k:]
    ndec = len(mant.split(".")[1]) if "." in mant else 0
    return 10.0 ** ((int(ex) if ex else 0) - ndec)


def _same_text(a, b, what, floor_rel=0.0):
    """byte-identical; should two runs ever differ in the last printed digit of a number (different tiles => different summation
    order inside a point), the text around the numbers must still be identical and every number within one unit of its last
    printed digit.  floor_rel: extra absolute allowance for E-format data values, relative to the largest one in the text (values
    that cancel to almost nothing near a node only agree at the scale of the summed terms, like the 1e-10 parity contract)"""
    if a == b:
        return
    ta, tb = NUM.split(a), NUM.split(b)
    assert [x.split() for x in ta] == [x.split() for x in tb], f"{what}: layout differs"      # blanks move with a number's sign
    fa, fb = NUM.findall(a), NUM.findall(b)
    conv = lambda x: float(x.replace("D", "E").replace("d", "e")) if ("E" in x.upper() or "D" in x.upper() or not re.search(r"\d[-+]\d", x)) \
        else float(re.sub(r"(\d)([-+]\d)", r"\1E\2", x))
    na, nb = np.array([conv(x) for x in fa]), np.array([conv(x) for x in fb])
    tol = 1.01 * np.maximum([_unit_last_digit(x) for x in fa], [_unit_last_digit(x) for x in fb])
    if floor_rel > 0.0:
        is_e = np.array([bool(re.search(r"\d[EeDd]?[-+]\d", x)) for x in fa])
        if is_e.any():
            tol = tol + np.where(is_e, floor_rel * np.abs(nb[is_e]).max(), 0.0)
    bad = np.abs(na - nb) > tol
    assert not bad.any(), (what, [(fa[i], fb[i]) for i in np.flatnonzero(bad)[:5]])


def _pair(tmp_path, name, mol, xdens, extra=None):
    dirs = []
    for k in ("nat", "py"):
        d = tmp_path / k / name
        d.mkdir(parents=True)
        shutil.copy(mol, d / "MOL"); shutil.copy(xdens, d / "XDENS")
        shutil.copy(os.path.join(INPUTS, name + ".inp"), d / "gimic.inp")
        if extra:
            extra(d)
        dirs.append(d)
    return dirs


def _run_both(dn, dp, args=()):
    from gimic_b200.driver import Driver
    p = subprocess.run([EXE, *args, str(dn / "gimic.inp")], capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stderr
    out = io.StringIO()
    Driver(str(dp / "gimic.inp"), out=out, vtk_appended=("appended" in args)).run()
    _same_text(fixtures.strip_clock(p.stdout), fixtures.strip_clock(out.getvalue()), "report")
    assert sorted(os.listdir(dn)) == sorted(os.listdir(dp))
    for f in sorted(os.listdir(dn)):
        if not filecmp.cmp(dn / f, dp / f, shallow=False):
            _same_text(fixtures.strip_clock(open(dn / f, errors="replace").read()), fixtures.strip_clock(open(dp / f, errors="replace").read()), f,
                       floor_rel=1e-9)
    return p.stdout


def test_native_c4h4_read_grid_and_integration(tmp_path, cases):
    """the two c4h4 reference cases: jvec.vtu on the 4110-point file grid (golden: 10 digits) and the bond-plane integral"""
    from make_golden import read_vtu_vectors
    gold = fixtures.golden_npz("c4h4_readgrid.npz")

    def extra(d):
        np.savetxt(d / "gridfile.grd", gold["grid"], fmt="%.6f")
        with open(d / "grid.1.ele", "w") as f:
            f.write("3  4  0\n    1    1475  1730  1474  1717\n    2     100     8   112   245\n    3     5     6     7     8\n")
    dn, dp = _pair(tmp_path, "c4h4_read-grid", cases["c4h4"]["mol"], cases["c4h4"]["xdens"], extra)
    _run_both(dn, dp)
    _, vec = read_vtu_vectors(str(dn / "jvec.vtu"))
    ref = gold["jvec"]
    assert np.abs(vec - ref).max() < 1e-9 * np.abs(ref).max() + 1e-14          # the reference's own golden, straight from the C++ driver
    dn, dp = _pair(tmp_path, "c4h4_integration", cases["c4h4"]["mol"], cases["c4h4"]["xdens"])
    text = _run_both(dn, dp)
    want = next(b for b in fixtures.golden_json("c4h4_integration.json")["blocks"] if b["section"] == "current" and b["spin"] == "total")
    m = re.search(r"Induced current \(au\)\s+:\s*([-\d.]+)", text)
    assert m is not None and abs(float(m.group(1)) - want["au"]) < 1.01e-6       # the reference's printed value, from the C++ driver


def test_native_open_shell_cases(tmp_path, cases):
    """UHF: the 33^3 cdens run (8 .vti files through the J path, alpha/beta combined by linearity) and the four-spin-case integral"""
    for name in ("open-shell_3d", "open-shell_integration"):
        dn, dp = _pair(tmp_path, name, cases["open_shell"]["mol"], cases["open_shell"]["xdens"])
        text = _run_both(dn, dp)
        assert "Open-shell calculation" in text
    assert os.path.exists(tmp_path / "nat" / "open-shell_3d" / "jvecspindens.vti")


@pytest.mark.parametrize("inp_name", sorted(f[:-4] for f in os.listdir(INPUTS) if f.startswith("benzene_")))
def test_native_every_benzene_input(tmp_path, cases, inp_name):
    """all 19 test/benzene inputs (synthetic densities, nbf = 252): program == library entry, incl. the ACID / tensor path, jmod.txt on
    Gauss grids, rotation / radius / spacing keywords and the property report of the magnetizability input"""
    xd = tmp_path / "XDENS"
    fixtures.write_xdens(str(xd), fixtures.dens_to_colmajor(fixtures.synthetic_density(252, seed=21)))

    def extra(d):
        if inp_name != "benzene_magnetizability":
            return
        from gimic_b200.driver import mol_geometry as read_mol_geometry
        _, coords = read_mol_geometry(str(d / "MOL"))
        rng = np.random.default_rng(5)
        counts = rng.integers(60, 120, size=coords.shape[0])
        pts = np.vstack([coords[a] + rng.normal(scale=1.5, size=(c, 3)) for a, c in enumerate(counts)])
        np.savetxt(d / "gridfile.grd", pts, fmt="%.10f"); np.savetxt(d / "grid_w.grd", rng.uniform(0, 0.1, size=pts.shape[0]), fmt="%.12e")
        shutil.copy(os.path.join(GOLD, "benzene_coord.au"), d / "coord.au")
        np.savetxt(d / "nelpts.info", np.column_stack([np.arange(1, len(counts) + 1), counts]), fmt="%d")
        with open(d / "grid.1.ele", "w") as f:
            f.write("2  4  0\n    1    1  2  3  4\n    2     5     6     7     8\n")
    dn, dp = _pair(tmp_path, inp_name, cases["benzene_mol"], xd, extra)
    text = _run_both(dn, dp)
    if inp_name == "benzene_magnetizability":
        assert "isotropic magnetizability chi" in text and os.path.exists(dn / "intchi.vtu") and os.path.exists(dn / "sigma_zz12.vtu")


def test_native_appended_vtk_and_scalar_modes(tmp_path, cases):
    """--vtk appended, calc=edens and calc=divj"""
    dn, dp = _pair(tmp_path, "open-shell_3d", cases["open_shell"]["mol"], cases["open_shell"]["xdens"])
    _run_both(dn, dp, ("--vtk", "appended"))
    for calc in ("edens", "divj"):
        def extra(d, calc=calc):
            txt = open(d / "gimic.inp").read().replace("calc=integral", "calc=" + calc)
            txt = re.sub(r"Grid\(bond\) \{.*?\n\}", "Grid(base) {\n type=even\n origin=[-4.0,-3.0,-1.0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n"
                         " lengths=[3.0,3.0,1.0]\n spacing=[0.5,0.5,0.5]\n}", txt, flags=re.S)
            open(d / "gimic.inp", "w").write(txt)
        dn, dp = _pair(tmp_path / calc, "c4h4_integration", cases["c4h4"]["mol"], cases["c4h4"]["xdens"], extra)
        _run_both(dn, dp)
        assert os.path.exists(dn / f"{calc}.vti")


def test_native_scan_equals_launcher_scan(tmp_path, cases):
    """a current-profile scan: `gimic-b200 gimic.0.inp ... gimic.5.inp` (one context, one batched tensor pass) writes the same
    gimic.N.out reports as gimic_b200_run_scan through the Python launcher and as separate runs of the program"""
    from gimic_b200.driver import run_scan
    dn, dp = _pair(tmp_path, "c4h4_integration", cases["c4h4"]["mol"], cases["c4h4"]["xdens"])
    base = open(dn / "gimic.inp").read()
    edges = np.linspace(-1.25614, 6.0, 7)
    names = {"nat": [], "py": []}
    for k in range(6):
        txt = base.replace("width=[-1.25614, 6.0]", f"width=[{edges[k]:.6f}, {edges[k + 1]:.6f}]").replace("grid_points=[30, 30, 0]", "grid_points=[30, 9, 0]")
        for key, d in (("nat", dn), ("py", dp)):
            (d / f"gimic.{k}.inp").write_text(txt)
            names[key].append(str(d / f"gimic.{k}.inp"))
    p = subprocess.run([EXE, *names["nat"]], capture_output=True, text=True, timeout=240)
    assert p.returncode == 0, p.stderr
    run_scan(names["py"])
    for k in range(6):
        a = fixtures.strip_clock(open(dn / f"gimic.{k}.out").read())
        _same_text(a, fixtures.strip_clock(open(dp / f"gimic.{k}.out").read()), f"gimic.{k}.out")
        single = subprocess.run([EXE, names["nat"][k]], capture_output=True, text=True, timeout=240)
        assert single.returncode == 0, single.stderr
        _same_text(a, fixtures.strip_clock(single.stdout), f"separate run {k}")


@pytest.mark.parametrize("name,case", [("c4h4_integration", "c4h4"), ("open-shell_3d", "open_shell"), ("c4h4_read-grid", "c4h4")])
def test_native_multi_device_partition_equals_single_device(tmp_path, cases, name, case):
    """--devices 0,0: two contexts (here on the same GPU), point slabs / plane rows split between them -- the single-process form of
    schedule() (parallel.F90:66-84).  Same report and files as the single-device run at print precision (tiles differ, so the
    summation order inside a point may)."""
    gold = fixtures.golden_npz("c4h4_readgrid.npz")

    def extra(d):
        if name == "c4h4_read-grid":
            np.savetxt(d / "gridfile.grd", gold["grid"], fmt="%.6f")
    dn, dp = _pair(tmp_path, name, cases[case]["mol"], cases[case]["xdens"], extra)
    one = subprocess.run([EXE, str(dn / "gimic.inp")], capture_output=True, text=True, timeout=240)
    two = subprocess.run([EXE, "--devices", "0,0", str(dp / "gimic.inp")], capture_output=True, text=True, timeout=240)
    assert one.returncode == 0 and two.returncode == 0, (one.stderr, two.stderr)
    _same_text(fixtures.strip_clock(one.stdout), fixtures.strip_clock(two.stdout), "report")
    assert sorted(os.listdir(dn)) == sorted(os.listdir(dp))
    for f in sorted(os.listdir(dn)):
        if not filecmp.cmp(dn / f, dp / f, shallow=False):
            _same_text(open(dn / f, errors="replace").read(), open(dp / f, errors="replace").read(), f, floor_rel=1e-9)


def test_rank_mode_over_nccl_when_two_gpus():
    """`torchrun --nproc-per-node 2 -m gimic_b200` semantics on hardware: the compiled driver in rank mode, rows gathered and integral sums
    reduced over NCCL through the launcher's two callbacks; rank 0's files and report against the reference's goldens
    (tools/dist_driver_check.py).  The same glue runs on CPU over gloo in tests/test_native_driver_mock.py."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (recorded on a 2-GPU box in profiles/r02_dist_driver_check_n2.txt)")
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29573", os.path.join(ROOT, "tools", "dist_driver_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and '"cdens_open_shell_files_match_golden": true' in p.stdout and '"integral_matches_golden": true' in p.stdout, \
        p.stdout[-2000:] + p.stderr[-2000:]
