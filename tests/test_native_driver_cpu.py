"""CPU-side checks of the native (C++) run-mode driver, libgimic_b200_driver.so + the gimic-b200 program
(include/gimic_b200_driver.h): the compiled counterpart of `gimic gimic.inp` (src/gimic.in:25-159 + src/fgimic/gimic.F90).
Input parsing, grid geometry, field direction, report text and every file format must equal the Python driver above the same
C ABI (which the other tests pin against the oracle and the reference's goldens) byte for byte; what needs the GPU is in
tests/test_native_driver_gpu.py."""
import ctypes as C
import filecmp
import io
import os
import re
import shutil
import subprocess
import numpy as np
import pytest

import fixtures

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = fixtures.GOLD
INPUTS = os.path.join(GOLD, "inputs")
EXE = os.path.join(ROOT, "gimic_b200", "gimic-b200")
DRV_SO = os.path.join(ROOT, "gimic_b200", "libgimic_b200_driver.so")
ALL_INPUTS = sorted(f[:-4] for f in os.listdir(INPUTS))


@pytest.fixture(scope="module")
def D():
    import __graft_entry__ as ge
    from gimic_b200 import _lib
    if not (os.path.exists(_lib.SO_PATH) and os.path.exists(DRV_SO) and os.path.exists(EXE)):
        ge.build()
    _lib.lib()                                   # libgimic_b200.so first (RTLD_GLOBAL), the driver links against it
    L = C.CDLL(DRV_SO)
    L.gimic_b200_run_input.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p]
    L.gimic_b200_run_scan.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_int, C.c_int]
    L.gimic_b200_write_field.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.POINTER(C.c_double), C.c_long, C.c_char_p, C.c_int]
    L.gimic_b200_driver_last_error.restype = C.c_char_p
    return L


def _mol_for(name):
    return "c4h4_MOL" if name.startswith("c4h4") else "open_shell_MOL" if name.startswith("open") else "benzene_MOL"


def _workdir(base, name, text=None):
    d = base / name
    d.mkdir(parents=True)
    shutil.copy(os.path.join(GOLD, _mol_for(name)), d / "MOL")
    if text is None:
        shutil.copy(os.path.join(INPUTS, name + ".inp"), d / "gimic.inp")
    else:
        (d / "gimic.inp").write_text(text)
    if "read-grid" in name or "magnetizability" in name:
        np.savetxt(d / "gridfile.grd", fixtures.golden_npz("c4h4_readgrid.npz")["grid"][:64], fmt="%.6f")
    return d


def test_driver_header_symbols_exported(D):
    hdr = open(os.path.join(ROOT, "include", "gimic_b200_driver.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(gimic_b200_\w+)\s*\(", hdr))
    assert names == {"gimic_b200_run_input", "gimic_b200_run", "gimic_b200_run_scan", "gimic_b200_write_field", "gimic_b200_cache_xdens",
                     "gimic_b200_driver_last_error"}
    for n in names:
        assert hasattr(D, n), n


@pytest.mark.parametrize("name", ALL_INPUTS)
def test_native_dry_run_equals_python_driver(D, tmp_path, name):
    """-y on every reference input (19 benzene, 2 c4h4, 2 open-shell): report text, mol.xyz and grid.xyz byte-identical to the
    Python driver -- covers the gimic.inp reader, std / bond / file grids, even / gauss / lobatto axes, rotation, radius and
    get_magnet.  No XDENS in the directory, no GPU in this container: the dry run reaches no compute entry point."""
    from gimic_b200 import driver
    dn, dp = _workdir(tmp_path / "nat", name), _workdir(tmp_path / "py", name)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    out = io.StringIO()
    driver.Driver(str(dp / "gimic.inp"), out=out, dryrun=True).run()
    assert fixtures.strip_clock(p.stdout) == fixtures.strip_clock(out.getvalue())
    assert sorted(os.listdir(dn)) == sorted(os.listdir(dp))
    for f in os.listdir(dn):
        assert filecmp.cmp(dn / f, dp / f, shallow=False), f
    # the library entry point with a report file gives the same text
    rep = tmp_path / "report.txt"
    assert D.gimic_b200_run_input(os.fsencode(dn / "gimic.inp"), None, -1, 1, os.fsencode(rep)) == 0
    assert fixtures.strip_clock(rep.read_text()) == fixtures.strip_clock(out.getvalue())
    assert "wall time:" in p.stdout and p.stdout.rstrip().endswith("done.")            # the trailer the jobscripts grep for


def _py_grid(d):
    from gimic_b200 import inp as _inp, grids, driver
    I = _inp.parse_file(str(d / "gimic.inp"))
    _, xyz = driver.mol_geometry(str(d / "MOL"))
    return grids.from_input(I, xyz, str(d))


def _write(D, d, kind, data, fname, flags=0):
    v = np.ascontiguousarray(data, dtype=np.float64).ravel()
    rc = D.gimic_b200_write_field(os.fsencode(d / "gimic.inp"), None, kind.encode(), v.ctypes.data_as(C.POINTER(C.c_double)), v.size,
                                  fname.encode(), flags)
    assert rc == 0, D.gimic_b200_driver_last_error().decode()


@pytest.mark.parametrize("name", ["benzene_3d", "benzene_2d", "benzene_keyword-radius", "benzene_keyword-rotation", "open-shell_3d"])
def test_native_vti_writers_equal_python_writers(D, tmp_path, name):
    """write_vtk_imagedata / write_vtk_vector_imagedata (vtkplot.f90:14-234) incl. the CellData block of cell-averaged |J|, the
    radius mask of 2-D bond grids (jfield.f90:310-346) and the appended-binary extra: native bytes == Python writer bytes"""
    from gimic_b200 import writers
    text = open(os.path.join(INPUTS, name + ".inp")).read()
    if name == "benzene_keyword-radius":        # make it the 2-D even bond grid the mask applies to
        text = text.replace("calc=integral", "calc=cdens").replace("type=gauss", "type=even")
        text = re.sub(r"gauss_order\s*=\s*\d+", "", text)
    d = _workdir(tmp_path, name, text)
    g = _py_grid(d)
    rng = np.random.default_rng(11)
    vec = rng.normal(size=(g.n, 3)) * 10.0 ** rng.integers(-12, 3, size=(g.n, 1))
    vec[::17] = 0.0
    sca = rng.normal(size=g.n) * 10.0 ** rng.integers(-120, 3, size=g.n)     # three-digit exponents drop the 'E'
    for appended in (False, True):
        tag = "a" if appended else ""
        writers.write_vti_vector(str(d / f"py_vec{tag}.vti"), g, writers.radius_masked_vectors(g, vec), appended)
        writers.write_vti_scalar(str(d / f"py_sca{tag}.vti"), g, sca, appended)
        _write(D, d, "vti_vector", vec, f"nat_vec{tag}.vti", 2 if appended else 0)
        _write(D, d, "vti_scalar", sca, f"nat_sca{tag}.vti", 2 if appended else 0)
        assert filecmp.cmp(d / f"py_vec{tag}.vti", d / f"nat_vec{tag}.vti", shallow=False), (name, appended)
        assert filecmp.cmp(d / f"py_sca{tag}.vti", d / f"nat_sca{tag}.vti", shallow=False), (name, appended)
    if name == "benzene_keyword-radius":
        masked = writers.radius_masked_vectors(g, vec)
        assert masked is not vec and (np.abs(masked).sum(1) == 0).sum() > (np.abs(vec).sum(1) == 0).sum()    # the mask did bite


def test_native_jmod_txt_and_vtu_writers_equal_python_writers(D, tmp_path):
    """jmod.txt ('(6f11.7)' rows with a blank line per i-row on bond grids, jfield.f90:356-376,531-541) and the UnstructuredGrid
    writers (vtkplot.f90:241-391) on a Grid(file) input with a TetGen .ele file"""
    from gimic_b200 import writers
    rng = np.random.default_rng(12)
    d = _workdir(tmp_path, "benzene_integration-gauss", open(os.path.join(INPUTS, "benzene_integration-gauss.inp")).read().replace("calc=integral", "calc=cdens"))
    g = _py_grid(d)
    vec = rng.normal(size=(g.n, 3))
    writers.write_jmod_txt(str(d / "py_jmod.txt"), g, vec, regular=True)
    _write(D, d, "jmod_txt", vec, "nat_jmod.txt")
    assert filecmp.cmp(d / "py_jmod.txt", d / "nat_jmod.txt", shallow=False)
    d2 = _workdir(tmp_path, "c4h4_read-grid")
    with open(d2 / "grid.1.ele", "w") as f:
        f.write("3  4  0\n    1    14  17  4  7\n    2     10     8   12   45\n    3     5     6     7     8\n")
    g2 = _py_grid(d2)
    assert g2.n == 64
    vec2 = rng.normal(size=(g2.n, 3)) * 1e-3
    cells = writers.read_ele(str(d2 / "grid.1.ele"))
    writers.write_vtu_vector(str(d2 / "py.vtu"), g2.points(), vec2, cells)
    writers.write_vtu_scalar(str(d2 / "pys.vtu"), g2.points(), vec2[:, 0], cells)
    _write(D, d2, "vtu_vector", vec2, "nat.vtu")
    _write(D, d2, "vtu_scalar", vec2[:, 0], "nats.vtu")
    assert filecmp.cmp(d2 / "py.vtu", d2 / "nat.vtu", shallow=False) and filecmp.cmp(d2 / "pys.vtu", d2 / "nats.vtu", shallow=False)
    # wrong length and unknown kind are errors with a message, not crashes
    v = np.zeros(5)
    rc = D.gimic_b200_write_field(os.fsencode(d2 / "gimic.inp"), None, b"vtu_vector", v.ctypes.data_as(C.POINTER(C.c_double)), 5, b"x.vtu", 0)
    assert rc < 0 and b"expected 192 values" in D.gimic_b200_driver_last_error()
    rc = D.gimic_b200_write_field(os.fsencode(d2 / "gimic.inp"), None, b"nonsense", v.ctypes.data_as(C.POINTER(C.c_double)), 64, b"x", 0)
    assert rc < 0 and b"unknown kind" in D.gimic_b200_driver_last_error()


BAD_INPUTS = [
    ("calc=cdens\nmagnet=[0,0,1]\n", "no Grid section"),
    ("calc=nonsense\nmagnet=[0,0,1]\nGrid(std){\n type=even\n origin=[0,0,0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n lengths=[1,1,1]\n spacing=[1,1,1]\n}\n", "unknown option calc"),
    ("calc=cdens\nGrid(std){\n type=even\n origin=[0,0,0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n lengths=[1,1,1]\n spacing=[1,1,1]\n}\n", "Direction of magnetic field"),
    ("calc=cdens\nmagnet=[0,0,1]\nmagnet_axis=z\nGrid(std){\n type=even\n origin=[0,0,0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n lengths=[1,1,1]\n spacing=[1,1,1]\n}\n", "Both magnet vector and axis"),
    ("calc=cdens\nmagnet=[0,0,1]\nGrid(std){\n type=even\n origin=[0,0,0]\n ivec=[1,0,0]\n lengths=[1,1,1]\n spacing=[1,1,1]\n}\n", "Required option 'jvec'"),
    ("calc=cdens\nmagnet=[0,0,1]\nGrid(bond){\n type=even\n bond=[1,2]\n fixpoint=3\n distance=1.0\n spacing=[1,1,1]\n}\n", "width and height"),
    ("calc=cdens\nmagnet=[0,0,1]\nfrobnicate=3\nGrid(file){\n file=gridfile.grd\n}\n", "unknown keyword 'frobnicate'"),
    ("calc=cdens\nmagnet=[0,0,1]\nGrid(std){\n type=even\n origin=[0,0,0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n lengths=[1,1,1]\n spacing=[1,1,1]\n", "unbalanced"),
    ("calc=cdens\nmagnet=[0,0,0]\nGrid(std){\n type=even\n origin=[0,0,0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n lengths=[1,1,1]\n spacing=[1,1,1]\n}\n", "Magnetic field is zero"),
    ("calc=integral\nmagnet_axis=X\nGrid(bond){\n type=gauss\n gauss_order=0\n bond=[1,2]\n fixpoint=3\n distance=1.0\n height=[-1,1]\n width=[-1,1]\n grid_points=[9,9,0]\n}\n", "gauss_order must be positive"),
    ("calc=cdens\nmagnet=[0,0,1]\nGrid(bond){\n type=even\n bond=[1,99]\n fixpoint=3\n distance=1.0\n height=[-1,1]\n width=[-1,1]\n spacing=[1,1,1]\n}\n", "out of range"),
]


@pytest.mark.parametrize("text,needle", BAD_INPUTS)
def test_native_rejects_bad_input_like_the_front_end(D, tmp_path, text, needle):
    """check_top / check_grid (src/gimic.in:161-283) and the grid set-up errors: a negative code and the front end's message;
    the Python reader raises on the same inputs"""
    from gimic_b200 import inp as _inp, grids, driver
    d = _workdir(tmp_path, "benzene_bad", 'basis="MOL"\n' + text)
    rc = D.gimic_b200_run_input(os.fsencode(d / "gimic.inp"), None, -1, 1, os.fsencode(tmp_path / "rep"))
    assert rc < 0
    assert needle in D.gimic_b200_driver_last_error().decode()
    with pytest.raises(Exception):
        I = _inp.parse_file(str(d / "gimic.inp"))
        _, xyz = driver.mol_geometry(str(d / "MOL"))
        grids.get_magnet(grids.from_input(I, xyz, str(d)), I.get("magnet_axis"), I.get("magnet"))
    p = subprocess.run([EXE, "-y", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and needle in p.stderr


def test_native_driver_has_no_cpu_fallback(D, tmp_path):
    """without -y the run needs the device context: in a GPU-less container it fails loudly (nothing is computed on the host)"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    d = _workdir(tmp_path, "c4h4_integration")
    fixtures.write_xdens(str(d / "XDENS"), fixtures.golden_npz("c4h4_xdens.npz")["xdens"])
    p = subprocess.run([EXE, str(d / "gimic.inp")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 1 and "Induced current" not in p.stdout
    assert "CUDA" in p.stderr or "cuda" in p.stderr
    rc = D.gimic_b200_run_input(os.fsencode(d / "gimic.inp"), None, -1, 0, os.fsencode(tmp_path / "rep"))
    assert rc == -3                                           # GIMIC_B200_ECUDA
    missing = subprocess.run([EXE, str(tmp_path / "nothing.inp")], capture_output=True, text=True, timeout=60)
    assert missing.returncode == 1 and "cannot open input file" in missing.stderr


def test_python_repr_layout_of_the_appended_header_numbers(D, tmp_path):
    """the appended-VTK extra prints Origin / Spacing like Python's repr(float): fixed notation for 1e-4 <= |x| < 1e16"""
    from gimic_b200 import writers
    text = ("basis=MOL\ncalc=cdens\nmagnet=[0,0,1]\nGrid(std){\n type=even\n origin=[-100000.0, 1.0e-5, 0.30000000000000004]\n ivec=[1,0,0]\n jvec=[0,1,0]\n"
            " lengths=[2.0e5, 1.0e-4, 3.0]\n grid_points=[3,3,3]\n}\n")
    d = _workdir(tmp_path, "benzene_repr", text)
    g = _py_grid(d)
    v = np.arange(g.n, dtype=np.float64)
    writers.write_vti_scalar(str(d / "py.vti"), g, v, True)
    _write(D, d, "vti_scalar", v, "nat.vti", 2)
    assert filecmp.cmp(d / "py.vti", d / "nat.vti", shallow=False)
    head = open(d / "nat.vti", "rb").read(400).decode("latin1")
    assert 'Origin="-100000.0 1e-05 0.30000000000000004"' in head and 'Spacing="100000.0 5e-05 1.5"' in head


SYNTAX_VARIANTS = {
    "baseline": lambda t: t,
    "crlf_free_spacing": lambda t: t.replace("=", "  =  ").replace("Grid(bond) {", "Grid ( bond )\n{"),
    "continuation": lambda t: t.replace("grid_points=[30, 30, 0]", "grid_points=[30, |\n      30, |   \n\n 0]"),
    "quotes_and_case": lambda t: t.replace("type=gauss", "type='gauss'").replace("acid=on", "acid=TRUE").replace("openshell=false", "openshell=No"),
    "fortran_exponents": lambda t: t.replace("distance=1.32", "distance=0.132D+01").replace("height=[-5.0, 5.0]", "height=[-0.5d1 5.e0]"),
    "comment_chars_in_strings": lambda t: t.replace('title=""', 'title="ring # 1 {test}"'),
    "one_line_sections": lambda t: t.replace("Essential {\n    acid=on\n}", "Essential { acid=on jmod=off }"),
    "gimlet_section_ignored": lambda t: t + "\nGimlet {\n  foo=1\n  bar=[1,2]\n}\n",
    "trailing_garbage": lambda t: t + "\n}\n",                                   # unbalanced: both must refuse
    "scalar_for_array": lambda t: t.replace("bond=[1,2]", "bond=1"),            # both accept a bare scalar for an array keyword? they must agree
    "array_for_scalar": lambda t: t.replace("fixpoint=4", "fixpoint=[4,5]"),   # both must refuse
    "bad_bool": lambda t: t.replace("acid=on", "acid=maybe"),
    "bad_number": lambda t: t.replace("distance=1.32", "distance=1.3.2"),
    "unknown_section": lambda t: t + "\nNonsense {\n a=1\n}\n",
}


@pytest.mark.parametrize("variant", sorted(SYNTAX_VARIANTS))
def test_gimic_inp_surface_syntax_agrees_with_the_python_reader(D, tmp_path, variant):
    """doc/input.rst:4-19 surface syntax (free spacing, '|' continuation, quotes, Fortran d-exponents, '#' inside strings, one-line
    sections, the ignored Gimlet section) and its error cases: the native reader accepts exactly what the Python reader accepts and
    then produces the identical dry-run report and grid.xyz"""
    from gimic_b200 import driver
    text = SYNTAX_VARIANTS[variant](open(os.path.join(INPUTS, "benzene_integration-gauss.inp")).read())
    dn, dp = _workdir(tmp_path / "nat", "benzene_v", text), _workdir(tmp_path / "py", "benzene_v", text)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    out = io.StringIO()
    try:
        driver.Driver(str(dp / "gimic.inp"), out=out, dryrun=True).run()
        py_ok = True
    except Exception:
        py_ok = False
    assert (p.returncode == 0) == py_ok, (variant, p.stderr)
    if py_ok:
        assert fixtures.strip_clock(p.stdout) == fixtures.strip_clock(out.getvalue())
        assert filecmp.cmp(dn / "grid.xyz", dp / "grid.xyz", shallow=False)
    expect_ok = variant not in ("trailing_garbage", "array_for_scalar", "bad_bool", "bad_number", "unknown_section")
    if variant != "scalar_for_array":
        assert py_ok == expect_ok, variant


def test_driver_header_is_c99_and_a_c_caller_links(D, tmp_path):
    """include/gimic_b200_driver.h from plain C: a dry run through gimic_b200_run_input and the error path"""
    d = _workdir(tmp_path, "benzene_2d")
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include "gimic_b200_driver.h"
int main(int argc, char **argv) {
    int rc = gimic_b200_run_input(argv[1], NULL, -1, GIMIC_B200_RUN_DRYRUN, argv[2]);
    int bad = gimic_b200_run_input("/nonexistent/gimic.inp", NULL, -1, GIMIC_B200_RUN_DRYRUN, NULL);
    printf("%d %d %s\n", rc, bad, gimic_b200_driver_last_error());
    (void)argc;
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.join(ROOT, "gimic_b200")
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libgimic_b200_driver.so", "-l:libgimic_b200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe), str(d / "gimic.inp"), str(tmp_path / "rep.txt")], text=True)
    assert out.startswith("0 -2 cannot open input file"), out       # GIMIC_B200_EIO
    assert "Dry run, not calculating" in (tmp_path / "rep.txt").read_text() and (d / "grid.xyz").exists()


def test_multi_device_switch_without_gpus(D, tmp_path):
    """--devices (one process, one context per GPU): the dry run ignores it, a real run without GPUs fails loudly, a malformed list
    is a usage error"""
    import torch
    d = _workdir(tmp_path, "c4h4_integration")
    p = subprocess.run([EXE, "-y", "--devices", "all", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and "Dry run, not calculating" in p.stdout
    bad = subprocess.run([EXE, "--devices", "0,x", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert bad.returncode == 2 and "bad device list" in bad.stderr
    if not torch.cuda.is_available():
        fixtures.write_xdens(str(d / "XDENS"), fixtures.golden_npz("c4h4_xdens.npz")["xdens"])
        for devs in ("all", "0,1"):
            q = subprocess.run([EXE, "--devices", devs, str(d / "gimic.inp")], capture_output=True, text=True, timeout=120)
            assert q.returncode == 1 and ("CUDA" in q.stderr or "cuda" in q.stderr) and "Induced current" not in q.stdout
        from gimic_b200.driver import RunOpts                  # gimic_b200_run_opts
        D.gimic_b200_run.argtypes = [C.c_char_p, C.POINTER(RunOpts)]
        # a rank of a multi-process run without the two collective callbacks is a usage error, not a crash
        o = RunOpts(flags=0, device=-1, rank=1, nranks=2, report_path=os.fsencode(tmp_path / "rep"))
        assert D.gimic_b200_run(os.fsencode(d / "gimic.inp"), C.byref(o)) == -1 and b"callbacks" in D.gimic_b200_driver_last_error()
        o = RunOpts(flags=0, device=-1, ndevices=-1, report_path=os.fsencode(tmp_path / "rep"))
        assert D.gimic_b200_run(os.fsencode(d / "gimic.inp"), C.byref(o)) == -3          # GIMIC_B200_ECUDA
        o = RunOpts(flags=1, device=-1, ndevices=-1, title=b"via the struct", report_path=os.fsencode(tmp_path / "rep"))
        assert D.gimic_b200_run(os.fsencode(d / "gimic.inp"), C.byref(o)) == 0 and "TITLE: via the struct" in (tmp_path / "rep").read_text()


@pytest.mark.parametrize("mutation,needle", [
    (lambda m: m.replace("\n      2  1\n", "\n   99999999999     2  1\n", 1), "bad contraction block"),       # found by fuzzing: used to allocate 16 GB
    (lambda m: m.replace(" 6.0    1 4  5  3  2  1", " 6.0    1 -4  5  3  2  1", 1), "shell count"),
    (lambda m: m.replace(" 6.0    1 4  5  3  2  1", " 6.0    1 4  5  3  2  100", 1), "block counts"),
    (lambda m: m[: len(m) // 3], "truncated"),
    (lambda m: "", "INTGRL"),
])
def test_malformed_mol_files_are_errors_not_crashes(D, tmp_path, mutation, needle):
    """MOL reader (intgrl.f90:20-264) on damaged files: a message and a negative code within a second, no crash, no huge allocation"""
    import time
    mol = open(os.path.join(GOLD, "c4h4_MOL")).read()
    bad = mutation(mol)
    assert bad != mol
    d = _workdir(tmp_path, "c4h4_integration")
    (d / "MOL").write_text(bad)
    t0 = time.perf_counter()
    p = subprocess.run([EXE, "-y", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 1 and needle in p.stderr, p.stderr
    assert time.perf_counter() - t0 < 5.0


def test_reference_front_end_switches_are_accepted(D, tmp_path):
    """`gimic -t title -d 1 -b fgimic -o out -y gimic.inp` (src/gimic.in:36-57) keeps working with both drivers"""
    from gimic_b200 import driver
    d = _workdir(tmp_path, "benzene_2d")
    p = subprocess.run([EXE, "-t", "my job", "-d", "1", "-b", "fgimic", "-o", "out", "-y", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and "Dry run, not calculating" in p.stdout
    q = subprocess.run([EXE, "-b", "pygimic", "-y", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert q.returncode == 2 and "not provided" in q.stderr
    os.remove(d / "grid.xyz")
    assert driver.main([str(d / "gimic.inp"), "-t", "my job", "-d", "1", "-b", "fgimic", "-o", "out", "-y"]) == 0 and (d / "grid.xyz").exists()


@pytest.mark.parametrize("name,gold_json", [("c4h4_integration", "c4h4_integration.json"), ("open-shell_integration", "open_shell_integration.json")])
def test_report_preamble_equals_the_reference_stdout(D, tmp_path, name, gold_json):
    """What the reference prints before the results (test/*/integration/reference/stdout: TURBOMOLE note, atom and GTO counts, screening
    threshold, the 'Integration grid data' block, grid mode, rotation, adjusted quadrature sizes, point counts, grid plot note, field
    direction): the dry-run report carries the same lines, and the numbers of the grid block are the golden's"""
    d = _workdir(tmp_path, name)
    p = subprocess.run([EXE, "-y", str(d / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    out = p.stdout
    geo = fixtures.golden_json(gold_json)["geometry"]
    for key, label in (("center", " center "), ("origin", " origin "), ("basv1", " basv1  "), ("basv2", " basv2  "), ("basv3", " basv3  ")):
        line = next(l for l in out.split("\n") if l.startswith(label))
        assert np.allclose([float(x) for x in line[len(label):].split()], geo[key], rtol=0, atol=1.01e-6), (key, line)
    for needle in (" Integration grid data", " Grid mode = bond", " INFO: Integration grid selected.",
                   " INFO: Adjusted number of grid points for quadrature:    36   36    0", "   Number of grid points <v1,v2>:   36   36    1",
                   "   Total number of grid points  :      1296", " *** Grid plot in grid.xyz", " Number of atoms =   8", " Normalizing basis",
                   " *** Calculating screening coefficients", " INFO: Screening threshold:   0.1000E-07", "  TITLE:"):
        assert needle + "\n" in out, needle
    assert ("  Total number of contracted GTO's    168\n" in out) and (" INFO: Detected TURBOMOLE input\n" in out) == name.startswith("c4h4")


def test_python_repr_layout_on_random_magnitudes(D, tmp_path):
    """py_repr (shortest round-trip digits, repr() layout) against Python itself for 60 random origins / spacings over 40 decades"""
    from gimic_b200 import writers
    rng = np.random.default_rng(3)
    for it in range(60):
        o = rng.normal(size=3) * 10.0 ** rng.integers(-20, 20, size=3)
        l = np.abs(rng.normal(size=3)) * 10.0 ** rng.integers(-6, 17, size=3) + 1e-7
        text = ("basis=MOL\ncalc=cdens\nmagnet=[0,0,1]\nGrid(std){\n type=even\n origin=[%r, %r, %r]\n ivec=[1,0,0]\n jvec=[0,1,0]\n"
                " lengths=[%r, %r, %r]\n grid_points=[2,2,2]\n}\n" % (*map(float, o), *map(float, l)))
        d = _workdir(tmp_path, f"benzene_r{it}", text)
        g = _py_grid(d)
        v = np.arange(8, dtype=np.float64)
        writers.write_vti_scalar(str(d / "py.vti"), g, v, True)
        _write(D, d, "vti_scalar", v, "nat.vti", 2)
        assert filecmp.cmp(d / "py.vti", d / "nat.vti", shallow=False), (it, open(d / "py.vti", "rb").read(300), open(d / "nat.vti", "rb").read(300))
        # the ASCII header prints the same numbers list-directed (gfortran layout, 17 significant digits): ld_real on both sides
        writers.write_vti_scalar(str(d / "pya.vti"), g, v, False)
        _write(D, d, "vti_scalar", v, "nata.vti", 0)
        assert filecmp.cmp(d / "pya.vti", d / "nata.vti", shallow=False), (it, open(d / "pya.vti", "rb").read(400), open(d / "nata.vti", "rb").read(400))


def test_cache_xdens_switch(D, tmp_path):
    """gimic-b200 --cache-xdens gimic.inp: the text XDENS named in the input becomes <xdens>.bin (GB2XDENS magic, sizes from the MOL file and
    the openshell keyword, values bit for bit what the text parser reads); host only"""
    d = _workdir(tmp_path, "open-shell_integration")
    vals = fixtures.golden_npz("open_shell_xdens.npz")["xdens"]
    fixtures.write_xdens(str(d / "XDENS"), vals)
    p = subprocess.run([EXE, "--cache-xdens", str(d / "gimic.inp")], capture_output=True, text=True, timeout=120)
    assert p.returncode == 0 and "XDENS.bin" in p.stdout, p.stderr
    raw = open(d / "XDENS.bin", "rb").read()
    assert raw[:8] == b"GB2XDENS"
    nbf, nmat = np.frombuffer(raw[8:24], dtype=np.int64)
    assert (nbf, nmat) == (168, 8)
    got = np.frombuffer(raw[24:], dtype=np.float64)
    want = np.array([float("%.14E" % v) for v in vals])
    assert got.shape == want.shape and np.array_equal(got, want)
    (d / "XDENS").write_text("1.0\n")
    q = subprocess.run([EXE, "--cache-xdens", str(d / "gimic.inp")], capture_output=True, text=True, timeout=120)
    assert q.returncode == 1 and "too short" in q.stderr


FRAME_CASES = [("benzene/2d/reference/jvec.vti", "benzene_2d", "vti_vector"), ("benzene/2d-keyword-magnet/reference/jvec.vti", "benzene_2d-keyword-magnet", "vti_vector"),
               ("benzene/vectors/reference/jvec.vti", "benzene_vectors", "vti_vector"), ("benzene/3d/reference/jvec.vti", "benzene_3d", "vti_vector"),
               ("benzene/3d/reference/jmod.vti", "benzene_3d", "vti_scalar"), ("benzene/3d/reference/acid.vti", "benzene_3d", "vti_scalar"),
               ("benzene/3d-keyword-magnet/reference/jvec.vti", "benzene_3d-keyword-magnet", "vti_vector"),
               ("benzene/int-cdens/reference/jmod.txt", "benzene_int-cdens", "jmod_txt"),
               ("open-shell/3d/reference/jvec.vti", "open-shell_3d", "vti_vector"), ("open-shell/3d/reference/jmodspindens.vti", "open-shell_3d", "vti_scalar")]


@pytest.mark.parametrize("rel,name,kind", FRAME_CASES)
def test_output_files_have_the_skeleton_of_the_reference_files(D, tmp_path, rel, name, kind):
    """Every byte of the reference's own output files that is not a data value -- XML boilerplate, the list-directed WholeExtent / Origin /
    Spacing numbers as gfortran printed them, block sizes, values per line, line lengths, the blank lines of jmod.txt and (for jmod.txt) the
    coordinate columns themselves -- is reproduced by both writers.  The skeleton only depends on MOL + gimic.inp, so the benzene cases are
    checkable although their densities are missing from the reference tree (tests/golden/file_frames.json, made by make_golden.py)."""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import file_frame
    from gimic_b200 import writers
    want = fixtures.golden_json("file_frames.json")[rel]
    d = _workdir(tmp_path, name)
    g = _py_grid(d)
    ncomp = 1 if kind == "vti_scalar" else 3
    v = np.zeros((g.n, ncomp)) if ncomp == 3 else np.zeros(g.n)
    if kind == "vti_vector":
        writers.write_vti_vector(str(d / "py.out"), g, v)
    elif kind == "vti_scalar":
        writers.write_vti_scalar(str(d / "py.out"), g, v)
    else:
        writers.write_jmod_txt(str(d / "py.out"), g, v, regular=True)
    _write(D, d, kind, v, "nat.out")
    assert filecmp.cmp(d / "py.out", d / "nat.out", shallow=False)
    got = file_frame(str(d / "nat.out"), 33 if kind == "jmod_txt" else 0)
    assert got["rle"] == want["rle"], [(a, b) for a, b in zip(got["rle"], want["rle"]) if a != b][:3]
    if "coord_sha256" in want:
        assert got["coord_sha256"] == want["coord_sha256"]


def test_vtu_file_has_the_skeleton_of_the_reference_file(D, tmp_path):
    """test/c4h4/read-grid/reference/jvec.vtu: boilerplate, NumberOfPoints / NumberOfCells field widths, the 3e20.10 Points block (hashed: the
    coordinates come from gridfile.grd), connectivity / offsets / types / CellData line layouts for the reference's 23221 cells (stand-in
    connectivity: the TetGen file is not shipped with the tests, its values do not change any line length)"""
    import hashlib, sys
    sys.path.insert(0, GOLD)
    from make_golden import file_frame
    from gimic_b200 import writers
    want = fixtures.golden_json("file_frames.json")["c4h4/read-grid/reference/jvec.vtu"]
    d = _workdir(tmp_path, "c4h4_read-grid")
    np.savetxt(d / "gridfile.grd", fixtures.golden_npz("c4h4_readgrid.npz")["grid"], fmt="%.6f")
    ncells = 23221
    with open(d / "grid.1.ele", "w") as f:
        f.write(f"{ncells}  4  0\n" + "".join(f"{c + 1:6d} {1 + c % 4000:6d} {2 + c % 4000:6d} {3 + c % 4000:6d} {4 + c % 4000:6d}\n" for c in range(ncells)))
    g = _py_grid(d)
    v = np.zeros((g.n, 3))
    writers.write_vtu_vector(str(d / "py.vtu"), g.points(), v, writers.read_ele(str(d / "grid.1.ele")))
    _write(D, d, "vtu_vector", v, "nat.vtu")
    assert filecmp.cmp(d / "py.vtu", d / "nat.vtu", shallow=False)
    got = file_frame(str(d / "nat.vtu"))
    assert got["rle"] == want["rle"], [(a, b) for a, b in zip(got["rle"], want["rle"]) if a != b][:3]
    h = hashlib.sha256()
    for line in open(d / "nat.vtu").read().split("\n")[6:6 + 4110]:
        h.update(line.encode() + b"\n")
    assert h.hexdigest() == want["points_sha256"]


def test_python_entry_can_hand_the_run_to_the_compiled_driver(D, tmp_path):
    """python -m gimic_b200 --native: the same report file and grid.xyz as the Python driver (dry run here; no GPU needed)"""
    from gimic_b200 import driver
    d1, d2 = _workdir(tmp_path / "a", "benzene_keyword-rotation"), _workdir(tmp_path / "b", "benzene_keyword-rotation")
    assert driver.run_native(str(d1 / "gimic.inp"), dryrun=True, title="handed over", report=str(tmp_path / "native.txt")) == 0
    out = io.StringIO()
    driver.Driver(str(d2 / "gimic.inp"), out=out, dryrun=True, title="handed over").run()
    assert fixtures.strip_clock((tmp_path / "native.txt").read_text()) == fixtures.strip_clock(out.getvalue())
    assert filecmp.cmp(d1 / "grid.xyz", d2 / "grid.xyz", shallow=False)
    with pytest.raises(RuntimeError, match="cannot open input file"):
        driver.run_native(str(tmp_path / "missing.inp"), dryrun=True)


def test_one_grid_point_along_an_axis_gives_nan_coordinates_like_the_reference(D, tmp_path):
    """grid_points = 1 along an axis: the reference divides the length by npts - 1 = 0 (grid.f90:157) and carries on with NaN / Inf
    coordinates; both drivers do the same and print them the way gfortran does ('NaN', never printf's '-nan')"""
    from gimic_b200 import driver
    text = ('calc=cdens\ntitle=""\nbasis="MOL"\nxdens="XDENS"\ndebug=1\nopenshell=false\nmagnet=[0.0, 0.0, 1.0]\n'
            "Grid(base) {\n type=even\n origin=[-4.0, -4.0, -3.0]\n ivec=[1, 0, 0]\n jvec=[0, 1, 0]\n lengths=[4.5, 8.0, 0]\n grid_points=[3, 20, 1]\n}\n")
    dn, dp = _workdir(tmp_path / "nat", "benzene_nan", text), _workdir(tmp_path / "py", "benzene_nan", text)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    out = io.StringIO()
    with np.errstate(all="ignore"):
        driver.Driver(str(dp / "gimic.inp"), out=out, dryrun=True).run()
    assert fixtures.strip_clock(p.stdout) == fixtures.strip_clock(out.getvalue())
    assert filecmp.cmp(dn / "grid.xyz", dp / "grid.xyz", shallow=False)
    xyz = open(dn / "grid.xyz").read()
    assert "             NaN" in xyz and "nan" not in xyz


def test_randomized_grid_inputs_agree_between_the_drivers(D, tmp_path):
    """tools/fuzz_dryrun_drivers.py, two fixed seeds x 40 random gimic.inp files over the grid / magnet keywords: same accept / refuse
    decision, same dry-run report, same grid.xyz byte for byte; plus the input on which this comparison once failed: a field along a rotated
    in-plane basis vector (magnet_axis=i), where check_field's x > 0 (magnet.f90:75) is decided by the last bit of the summation"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("fuzz_dryrun_drivers", os.path.join(ROOT, "tools", "fuzz_dryrun_drivers.py"))
    fz = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(fz)
    with np.errstate(all="ignore"):
        assert fz.run(1, 40, tmp_path / "a") == 0
        assert fz.run(11, 40, tmp_path / "b") == 0
    text = ('calc=integral\ntitle=""\nbasis="MOL"\nxdens="XDENS"\ndebug=1\nopenshell=false\nmagnet_axis=i\nGrid(bond) {\n type=even\n bond=[5,11]\n'
            " fixpoint=10\n distance=2.48005\n height=[-4.49534, 5.27421]\n width=[-0.772161, 4.24791]\n grid_points=[9, 21, 0]\n"
            " rotation=[2.42421, -43.0696, -19.5473]\n}\n")
    from gimic_b200 import driver
    dn, dp = _workdir(tmp_path / "nat", "benzene_inplane", text), _workdir(tmp_path / "py", "benzene_inplane", text)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    out = io.StringIO()
    driver.Driver(str(dp / "gimic.inp"), out=out, dryrun=True).run()
    assert fixtures.strip_clock(p.stdout) == fixtures.strip_clock(out.getvalue())
