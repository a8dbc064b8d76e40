"""CPU-side checks of the native (C++) run-mode driver, libgimic_b200_driver.so + the gimic-b200 program
(include/gimic_b200_driver.h): the compiled counterpart of `gimic gimic.inp` (src/gimic.in:25-159 + src/fgimic/gimic.F90).
Input parsing against an independent reader of the grammar, grid geometry and field direction against the oracle's grid code, report
text against snapshots and the reference's stdout goldens, every file format against the skeletons of the reference's own files and a
round trip through the parser of the goldens; what needs the GPU is in tests/test_native_driver_gpu.py."""
import ctypes as C
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
                     "gimic_b200_input_grid", "gimic_b200_driver_last_error"}
    for n in names:
        assert hasattr(D, n), n


@pytest.mark.parametrize("name", ALL_INPUTS)
def test_native_dry_run_report_and_files(D, tmp_path, name):
    """-y on every reference input (19 benzene, 2 c4h4, 2 open-shell): report text, mol.xyz and grid.xyz equal the snapshot under
    tests/golden/native_dryrun (tests/golden/make_dryrun_snapshots.py says what it pins) -- covers the gimic.inp reader, std / bond / file
    grids, even / gauss / lobatto axes, rotation, radius and get_magnet.  The program, the library entry point and the Python launcher
    print the same text.  No XDENS in the directory, no GPU in this container: the dry run reaches no compute entry point."""
    from gimic_b200 import driver
    snap = os.path.join(GOLD, "native_dryrun")
    dn, dp = _workdir(tmp_path / "nat", name), _workdir(tmp_path / "py", name)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    assert fixtures.strip_clock(p.stdout) == open(os.path.join(snap, name + ".out")).read()
    for f in ("mol.xyz", "grid.xyz"):
        assert filecmp.cmp(dn / f, os.path.join(snap, name + "." + f), shallow=False), f
    out = io.StringIO()
    driver.Driver(str(dp / "gimic.inp"), out=out, dryrun=True).run()
    assert fixtures.strip_clock(p.stdout) == fixtures.strip_clock(out.getvalue())
    assert sorted(os.listdir(dn)) == sorted(os.listdir(dp))
    rep = tmp_path / "report.txt"
    assert D.gimic_b200_run_input(os.fsencode(dn / "gimic.inp"), None, -1, 1, os.fsencode(rep)) == 0
    assert fixtures.strip_clock(rep.read_text()) == fixtures.strip_clock(out.getvalue())
    assert "wall time:" in p.stdout and p.stdout.rstrip().endswith("done.")            # the trailer the jobscripts grep for


def _grid(d):
    """the grid gimic.inp describes, from the driver library itself (gimic_b200_input_grid): (Grid | points, info)"""
    from gimic_b200 import driver
    g, _, info = driver.input_grid(str(d / "gimic.inp"))
    return g, info


AU2A = float(np.float32(0.52917726))     # globals.f90:51 is a single-precision literal


def _write(D, d, kind, data, fname, flags=0):
    v = np.ascontiguousarray(data, dtype=np.float64).ravel()
    rc = D.gimic_b200_write_field(os.fsencode(d / "gimic.inp"), None, kind.encode(), v.ctypes.data_as(C.POINTER(C.c_double)), v.size,
                                  fname.encode(), flags)
    assert rc == 0, D.gimic_b200_driver_last_error().decode()


@pytest.mark.parametrize("name", ["benzene_3d", "benzene_2d", "benzene_keyword-radius", "benzene_keyword-rotation", "open-shell_3d"])
def test_native_vti_writers_round_trip(D, tmp_path, name):
    """write_vtk_imagedata / write_vtk_vector_imagedata (vtkplot.f90:14-234): the files parse back, with the parser that reads the
    reference's golden files, to the numbers that went in at the 6 significant digits of e14.6; the vector file carries the CellData block; the
    radius mask of 2-D bond grids (jfield.f90:310-346, with the reference's Angstrom-vs-bohr comparison) zeroes the far vectors;
    the appended-binary extra holds the same doubles bit for bit at the offsets its header names"""
    sys.path.insert(0, GOLD)
    from make_golden import read_vti
    text = open(os.path.join(INPUTS, name + ".inp")).read()
    if name == "benzene_keyword-radius":        # make it the 2-D even bond grid the mask applies to
        text = text.replace("calc=integral", "calc=cdens").replace("type=gauss", "type=even")
        text = re.sub(r"gauss_order\s*=\s*\d+", "", text)
    d = _workdir(tmp_path, name, text)
    g, info = _grid(d)
    rng = np.random.default_rng(11)
    vec = rng.normal(size=(g.n, 3)) * 10.0 ** rng.integers(-12, 3, size=(g.n, 1))
    vec[::17] = 0.0
    sca = rng.normal(size=g.n) * 10.0 ** rng.integers(-60, 3, size=g.n)
    expect = vec
    if name == "benzene_keyword-radius":
        pts = g.points().reshape(-1, 3)
        a, b = pts[g.npts[0] - 1], pts[g.npts[0] * (g.npts[1] - 1)]                 # grid_center, grid.f90:529-541
        far = np.sqrt(((pts * AU2A - 0.5 * (a + b)) ** 2).sum(1)) > info["radius"]
        assert far.any() and (~far).any()
        expect = np.where(far[:, None], 0.0, vec)
    _write(D, d, "vti_vector", vec, "vec.vti")
    _write(D, d, "vti_scalar", sca, "sca.vti")
    got_v, got_s = read_vti(str(d / "vec.vti")), read_vti(str(d / "sca.vti"))
    assert got_v.shape == expect.shape and np.allclose(got_v, expect, rtol=5.1e-6, atol=0)
    assert got_s.shape == sca.shape and np.allclose(got_s, sca, rtol=5.1e-6, atol=0)
    txt = open(d / "vec.vti").read()
    ext = "".join(f"{v:12d}" for v in (0, g.npts[0] - 1, 0, g.npts[1] - 1, 0, g.npts[2] - 1))
    assert f'WholeExtent="{ext} "' in txt and "<CellData" in txt
    ncell = max(g.npts[0] - 1, 1) * max(g.npts[1] - 1, 1) * max(g.npts[2] - 1, 1) if g.npts[2] > 1 else (g.npts[0] - 1) * (g.npts[1] - 1)
    assert txt.count("\n") >= 6 + g.n + ncell
    _write(D, d, "vti_vector", vec, "veca.vti", 2)
    _write(D, d, "vti_scalar", sca, "scaa.vti", 2)
    for fname, arrs in (("veca.vti", [expect.ravel(), None]), ("scaa.vti", [sca])):
        raw = open(d / fname, "rb").read()
        head, data = raw.split(b'<AppendedData encoding="raw">\n_', 1)
        offs = [int(x) for x in re.findall(rb'offset="(\d+)"', head)]
        assert len(offs) == len(arrs) and b'header_type="UInt64"' in head
        for off, ref in zip(offs, arrs):
            nbytes = int(np.frombuffer(data[off:off + 8], np.uint64)[0])
            got = np.frombuffer(data[off + 8:off + 8 + nbytes], "<f8")
            if ref is not None:
                assert np.array_equal(got, ref)
            else:
                assert (got >= 0).all()                                             # cell-averaged |J|


def test_native_jmod_txt_and_vtu_writers(D, tmp_path):
    """jmod.txt ('(6f11.7)' rows with a blank line per i-row on bond grids, jfield.f90:356-376,531-541) against a per-row restatement,
    and the UnstructuredGrid writers (vtkplot.f90:241-391) on a Grid(file) input with a TetGen .ele file, parsed back with the
    parser of the reference's golden jvec.vtu"""
    sys.path.insert(0, GOLD)
    from make_golden import read_vtu_vectors
    rng = np.random.default_rng(12)
    d = _workdir(tmp_path, "benzene_integration-gauss", open(os.path.join(INPUTS, "benzene_integration-gauss.inp")).read().replace("calc=integral", "calc=cdens"))
    g, _ = _grid(d)
    vec = rng.normal(size=(g.n, 3))
    _write(D, d, "jmod_txt", vec, "nat_jmod.txt")
    r = g.points().reshape(-1, 3) * AU2A; jm = np.sqrt((vec ** 2).sum(1))
    ref = ""
    for n in range(g.n):
        ref += "".join(f"{x:11.7f}" for x in (*r[n], jm[n])) + "\n"
        if (n + 1) % g.npts[0] == 0:
            ref += "\n"
    assert open(d / "nat_jmod.txt").read() == ref
    d2 = _workdir(tmp_path, "c4h4_read-grid")
    with open(d2 / "grid.1.ele", "w") as f:
        f.write("3  4  0\n    1    14  17  4  7\n    2     10     8   12   45\n    3     5     6     7     8\n")
    pts2, _ = _grid(d2)
    assert pts2.shape == (64, 3)
    vec2 = rng.normal(size=(64, 3)) * 1e-3
    _write(D, d2, "vtu_vector", vec2, "nat.vtu")
    _write(D, d2, "vtu_scalar", vec2[:, 0], "nats.vtu")
    p, v = read_vtu_vectors(str(d2 / "nat.vtu"))
    assert np.allclose(p, pts2, atol=1e-9) and np.allclose(v, vec2, rtol=1.01e-9, atol=0)       # e20.10
    body = open(d2 / "nats.vtu").read().split('<DataArray Name="scalars"')[1].split("</DataArray>")[0].split("\n", 1)[1]
    assert np.allclose([float(x) for x in body.split()], vec2[:, 0], rtol=1.01e-9, atol=0)
    cells = open(d2 / "nat.vtu").read().split('Name="connectivity"')[1].split("</DataArray>")[0].split("\n", 1)[1]
    assert [int(x) for x in cells.split()] == [13, 16, 3, 6, 9, 7, 11, 44, 4, 5, 6, 7]                  # 0-based node numbers
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
    """check_top / check_grid (src/gimic.in:161-283) and the grid set-up errors: a negative code and the front end's message, from
    the library entry points and from the program"""
    from gimic_b200 import driver
    d = _workdir(tmp_path, "benzene_bad", 'basis="MOL"\n' + text)
    rc = D.gimic_b200_run_input(os.fsencode(d / "gimic.inp"), None, -1, 1, os.fsencode(tmp_path / "rep"))
    assert rc < 0
    assert needle in D.gimic_b200_driver_last_error().decode()
    with pytest.raises(RuntimeError, match=re.escape(needle)):
        driver.input_grid(str(d / "gimic.inp"))
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
    text = ("basis=MOL\ncalc=cdens\nmagnet=[0,0,1]\nGrid(std){\n type=even\n origin=[-100000.0, 1.0e-5, 0.30000000000000004]\n ivec=[1,0,0]\n jvec=[0,1,0]\n"
            " lengths=[2.0e5, 1.0e-4, 3.0]\n grid_points=[3,3,3]\n}\n")
    d = _workdir(tmp_path, "benzene_repr", text)
    g, _ = _grid(d)
    v = np.arange(g.n, dtype=np.float64)
    _write(D, d, "vti_scalar", v, "nat.vti", 2)
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
def test_gimic_inp_surface_syntax(D, tmp_path, variant):
    """doc/input.rst:4-19 surface syntax (free spacing, '|' continuation, quotes, Fortran d-exponents, '#' inside strings, one-line
    sections, the ignored Gimlet section) and its error cases: the driver accepts exactly what an independent reader of the grammar
    (tests/inp_reader.py) accepts, and spellings that mean the same produce the baseline's grid.xyz"""
    import inp_reader
    base = open(os.path.join(INPUTS, "benzene_integration-gauss.inp")).read()
    text = SYNTAX_VARIANTS[variant](base)
    dn, db = _workdir(tmp_path / "nat", "benzene_v", text), _workdir(tmp_path / "base", "benzene_v", base)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    try:
        inp_reader.parse_text(text)
        py_ok = True
    except (inp_reader.InputError, ValueError):
        py_ok = False
    if variant == "scalar_for_array":           # bond=1 reads as a one-element array; the grid set-up then refuses it
        assert p.returncode == 1 and "needs two atom indices" in p.stderr
        return
    assert (p.returncode == 0) == py_ok, (variant, p.stderr)
    assert py_ok == (variant not in ("trailing_garbage", "array_for_scalar", "bad_bool", "bad_number", "unknown_section")), variant
    if py_ok and variant != "one_line_sections":          # same meaning as the baseline: same grid
        q = subprocess.run([EXE, "-y", str(db / "gimic.inp")], capture_output=True, text=True, timeout=60)
        assert q.returncode == 0 and filecmp.cmp(dn / "grid.xyz", db / "grid.xyz", shallow=False)


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
    """`gimic -t title -d 1 -b fgimic -o out -y gimic.inp` (src/gimic.in:36-57) keeps working, from the program and from `python -m gimic_b200`"""
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
    """py_repr (shortest round-trip digits, repr() layout) against Python itself, and the ASCII header's list-directed numbers (gfortran
    layout, 17 significant digits) against a per-value restatement, for 60 random origins / spacings over 40 decades"""
    from fortran_fmt import ld_real
    rng = np.random.default_rng(3)
    for it in range(60):
        o = rng.normal(size=3) * 10.0 ** rng.integers(-20, 20, size=3)
        l = np.abs(rng.normal(size=3)) * 10.0 ** rng.integers(-6, 17, size=3) + 1e-7
        text = ("basis=MOL\ncalc=cdens\nmagnet=[0,0,1]\nGrid(std){\n type=even\n origin=[%r, %r, %r]\n ivec=[1,0,0]\n jvec=[0,1,0]\n"
                " lengths=[%r, %r, %r]\n grid_points=[2,2,2]\n}\n" % (*map(float, o), *map(float, l)))
        d = _workdir(tmp_path, f"benzene_r{it}", text)
        g, _ = _grid(d)
        pts = g.points().reshape(-1, 3)
        qmin, step = pts[0], pts[-1] - pts[0]                                        # vtkplot.f90:33-38: 2 points per axis
        v = np.arange(8, dtype=np.float64)
        _write(D, d, "vti_scalar", v, "nat.vti", 2)
        head = open(d / "nat.vti", "rb").read(600).decode("latin1")
        m = re.search(r'Origin="([^"]*)" Spacing="([^"]*)"', head)
        assert m.group(1).split() == [repr(float(x)) for x in qmin], (it, head)
        assert [float(x) for x in m.group(2).split()] == [float(x) if x > 1e-8 else float(x) for x in step], (it, head)
        assert all(t == repr(float(t)) for t in m.group(2).split()), (it, head)
        _write(D, d, "vti_scalar", v, "nata.vti", 0)
        heada = open(d / "nata.vti", "rb").read(700).decode("latin1")
        assert 'Origin="' + "".join(ld_real(x) for x in qmin) + ' "' in heada, (it, heada)


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
    coordinate columns themselves -- is reproduced by the writers.  The skeleton only depends on MOL + gimic.inp, so the benzene cases are
    checkable although their densities are missing from the reference tree (tests/golden/file_frames.json, made by make_golden.py)."""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import file_frame
    want = fixtures.golden_json("file_frames.json")[rel]
    d = _workdir(tmp_path, name)
    g, _ = _grid(d)
    ncomp = 1 if kind == "vti_scalar" else 3
    v = np.zeros((g.n, ncomp)) if ncomp == 3 else np.zeros(g.n)
    _write(D, d, kind, v, "nat.out")
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
    want = fixtures.golden_json("file_frames.json")["c4h4/read-grid/reference/jvec.vtu"]
    d = _workdir(tmp_path, "c4h4_read-grid")
    np.savetxt(d / "gridfile.grd", fixtures.golden_npz("c4h4_readgrid.npz")["grid"], fmt="%.6f")
    ncells = 23221
    with open(d / "grid.1.ele", "w") as f:
        f.write(f"{ncells}  4  0\n" + "".join(f"{c + 1:6d} {1 + c % 4000:6d} {2 + c % 4000:6d} {3 + c % 4000:6d} {4 + c % 4000:6d}\n" for c in range(ncells)))
    pts, _ = _grid(d)
    v = np.zeros((pts.shape[0], 3))
    _write(D, d, "vtu_vector", v, "nat.vtu")
    got = file_frame(str(d / "nat.vtu"))
    assert got["rle"] == want["rle"], [(a, b) for a, b in zip(got["rle"], want["rle"]) if a != b][:3]
    h = hashlib.sha256()
    for line in open(d / "nat.vtu").read().split("\n")[6:6 + 4110]:
        h.update(line.encode() + b"\n")
    assert h.hexdigest() == want["points_sha256"]


def test_python_entry_runs_the_compiled_driver(D, tmp_path):
    """python -m gimic_b200 / gimic_b200.driver.run: the program's report file and grid.xyz (dry run here; no GPU needed)"""
    from gimic_b200 import driver
    d1, d2 = _workdir(tmp_path / "a", "benzene_keyword-rotation"), _workdir(tmp_path / "b", "benzene_keyword-rotation")
    assert driver.run(str(d1 / "gimic.inp"), dryrun=True, title="handed over", report=str(tmp_path / "native.txt")) == 0
    p = subprocess.run([EXE, "-y", "-t", "handed over", str(d2 / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and fixtures.strip_clock((tmp_path / "native.txt").read_text()) == fixtures.strip_clock(p.stdout)
    assert "TITLE: handed over" in p.stdout and filecmp.cmp(d1 / "grid.xyz", d2 / "grid.xyz", shallow=False)
    with pytest.raises(RuntimeError, match="cannot open input file"):
        driver.run(str(tmp_path / "missing.inp"), dryrun=True)


def test_one_grid_point_along_an_axis_gives_nan_coordinates_like_the_reference(D, tmp_path):
    """grid_points = 1 along an axis: the reference divides the length by npts - 1 = 0 (grid.f90:157) and carries on with NaN / Inf
    coordinates; the driver does the same and prints them the way gfortran does ('NaN', never printf's '-nan')"""
    text = ('calc=cdens\ntitle=""\nbasis="MOL"\nxdens="XDENS"\ndebug=1\nopenshell=false\nmagnet=[0.0, 0.0, 1.0]\n'
            "Grid(base) {\n type=even\n origin=[-4.0, -4.0, -3.0]\n ivec=[1, 0, 0]\n jvec=[0, 1, 0]\n lengths=[4.5, 8.0, 0]\n grid_points=[3, 20, 1]\n}\n")
    dn = _workdir(tmp_path / "nat", "benzene_nan", text)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0, p.stderr
    xyz = open(dn / "grid.xyz").read()
    assert "             NaN" in xyz and "nan" not in xyz and "nan" not in p.stdout


def test_randomized_grid_inputs_against_the_oracle_grids(D, tmp_path):
    """tools/fuzz_dryrun_drivers.py, two fixed seeds x 40 random gimic.inp files over the grid / magnet keywords: the driver accepts
    what an independent reader of the grammar accepts, and its grid (points, weights) and field direction are the oracle's; plus the
    input on which round 1's two-driver comparison once failed: a field along a rotated in-plane basis vector (magnet_axis=i),
    where check_field's x > 0 (magnet.f90:75) is decided by the last bit of the summation -- the dry run must simply succeed"""
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
    dn = _workdir(tmp_path / "nat", "benzene_inplane", text)
    p = subprocess.run([EXE, "-y", str(dn / "gimic.inp")], capture_output=True, text=True, timeout=60)
    assert p.returncode == 0 and "Magnetic field <x,y,z>" in p.stdout, p.stderr
