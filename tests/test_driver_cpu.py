"""CPU tests of the native driver pieces: gimic.inp reader, grid geometry and magnet
against the oracle AND against what the reference printed for its benzene keyword tests, output formatting."""
import io
import json
import os
import numpy as np
import pytest

import fixtures
import oracle_lib as O

GOLD = fixtures.GOLD
INPUTS = os.path.join(GOLD, "inputs")


def _inp(name):
    from gimic_b200 import inp
    return inp.parse_file(os.path.join(INPUTS, name + ".inp"))


def test_parser_reads_reference_inputs():
    names = sorted(f[:-4] for f in os.listdir(INPUTS))
    assert len(names) >= 14
    for n in names:
        I = _inp(n)
        assert I.get("calc") in ("cdens", "integral")
        assert I.get("Advanced.spherical") is False and I.get("Advanced.screening") is True
        assert abs(I.get("Advanced.screening_thrs") - 1e-8) < 1e-20          # written as 1.d-8
    I = _inp("c4h4_integration")
    assert I.grid_arg == "bond" and I.get("Grid.bond") == [2, 1] and I.get("Grid.fixpoint") == 4
    assert I.get("Grid.width") == [-1.25614, 6.0] and I.is_set("Grid.rotation") and not I.is_set("Grid.radius")
    assert I.get("magnet_axis") == "z" and not I.is_set("magnet")
    I = _inp("c4h4_read-grid")
    assert I.grid_arg == "file" and I.get("Grid.file") == "gridfile.grd" and I.get("magnet") == [0.0, 0.0, -1.0]
    I = _inp("open-shell_3d")
    assert I.get("openshell") is True and I.get("Essential.jmod") is True and I.get("Grid.spacing") == [0.5, 0.5, 0.5]


def test_parser_rejects_bad_input():
    from gimic_b200 import inp
    base = 'calc=cdens\nmagnet_axis=z\nGrid(std){ origin=[0,0,0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n lengths=[1,1,1]\n spacing=[0.5,0.5,0.5] }\n'
    assert inp.parse_text(base).get("Grid.lengths") == [1.0, 1.0, 1.0]
    for bad in (base.replace("calc=cdens", "calc=foo"), base.replace("magnet_axis=z", ""), base + "magnet=[0,0,1]\n",
                base.replace("spacing=[0.5,0.5,0.5]", ""), base.replace("Grid(std)", "Grid(cube)"),
                base.replace("spacing=", "grid_points=[3,3,3]\n spacing=")):
        with pytest.raises(inp.InputError):
            inp.parse_text(bad)
    assert inp.parse_text(base.replace("calc=cdens", "calc=cdens # comment\ntitle=\"a # b\"")).get("title") == "a # b"


def _grid_from(name, coords):
    from gimic_b200 import grids
    I = _inp(name)
    g = grids.from_input(I, coords, INPUTS)
    return I, g, grids.get_magnet(g, I.get("magnet_axis"), I.get("magnet"))


def _oracle_grid(I, coords):
    G = lambda k: I.get("Grid." + k)
    S = lambda k: I.is_set("Grid." + k)
    kw = dict(type=G("type"), gauss_order=G("gauss_order"), grid_points=G("grid_points") if S("grid_points") else None,
              spacing=G("spacing") if S("spacing") else None, rotation=G("rotation") if S("rotation") else None,
              rotation_origin=G("rotation_origin") if S("rotation_origin") else None)
    if I.grid_arg == "bond":
        b = G("bond")
        return O.grid_bond(coords[b[0] - 1], coords[b[1] - 1], coords[G("fixpoint") - 1], G("distance"), height=G("height"),
                           width=G("width"), radius=G("radius") if S("radius") else None,
                           magnet=I.get("magnet") if I.is_set("magnet") else None, **kw)
    return O.grid_std(G("origin"), G("ivec"), G("jvec"), G("lengths"), **kw)


@pytest.mark.parametrize("name,mol", [("c4h4_integration", "c4h4_MOL"), ("open-shell_integration", "open_shell_MOL"),
                                      ("open-shell_3d", "open_shell_MOL"), ("benzene_int-grid-bond-even", "benzene_MOL"),
                                      ("benzene_keyword-rotation", "benzene_MOL"), ("benzene_keyword-rotation_origin", "benzene_MOL"),
                                      ("benzene_keyword-radius", "benzene_MOL"), ("benzene_keyword-spacing", "benzene_MOL"),
                                      ("benzene_keyword-magnet", "benzene_MOL"), ("benzene_integration-lobatto", "benzene_MOL"),
                                      ("benzene_3d", "benzene_MOL"), ("benzene_2d", "benzene_MOL"), ("benzene_int-cdens", "benzene_MOL")])
def test_grid_and_magnet_match_oracle(name, mol):
    from gimic_b200.driver import read_mol_geometry
    _, coords = read_mol_geometry(os.path.join(GOLD, mol))
    I, g, mag = _grid_from(name, coords)
    og = _oracle_grid(I, coords)
    assert g.npts == og.npts
    assert np.allclose(g.origin, og.origin, rtol=0, atol=1e-13) and np.allclose(g.basv, og.basv, rtol=0, atol=1e-14)
    for d in range(3):
        p, w = og.axis(d)
        assert np.allclose(g.pts[d], p, rtol=0, atol=1e-13) and np.allclose(g.wgt[d], w, rtol=0, atol=1e-14)
    assert np.allclose(g.points(), og.points(), rtol=0, atol=1e-12)
    assert np.allclose(mag, og.magnet(I.get("magnet_axis"), I.get("magnet")), rtol=0, atol=1e-14)
    if I.grid_arg == "bond":
        assert g.radius == og.radius and np.allclose(g.center(), og.center(), atol=1e-13)


def test_grid_matches_what_the_reference_printed():
    """'Integration grid data' block, point counts and field direction of the reference's benzene stdout goldens"""
    from gimic_b200.driver import read_mol_geometry
    _, coords = read_mol_geometry(os.path.join(GOLD, "benzene_MOL"))
    gold = json.load(open(os.path.join(GOLD, "benzene_grids.json")))
    checked = 0
    for name, ref in gold.items():
        if not os.path.exists(os.path.join(INPUTS, f"benzene_{name}.inp")):
            continue
        I, g, mag = _grid_from(f"benzene_{name}", coords)
        if "npts" in ref:
            assert list(g.npts) == ref["npts"], name
        if ref.get("magnet"):
            assert np.allclose(mag, ref["magnet"], atol=1e-5), name
        geo = ref["geometry"]
        if geo and not I.is_set("Grid.rotation"):      # the block is printed before the rotation is applied
            assert np.allclose(g.center_bond, geo["center"], atol=1e-6), name
            assert np.allclose(g.origin, geo["origin"], atol=1e-6), name
            for v in range(3):
                assert np.allclose(g.basv[v], geo[f"basv{v + 1}"], atol=1e-6), name
            assert np.allclose(g.lengths, geo["lenghts"], atol=1e-6), name
        checked += 1
    assert checked >= 8


def test_fortran_number_formats():
    from gimic_b200.writers import fortran_e, _ld_real
    assert fortran_e(0.648806e-13, 14, 6) == "  0.648806E-13" and fortran_e(-0.671910e-13, 14, 6) == " -0.671910E-13"
    assert fortran_e(0.0, 14, 6) == "  0.000000E+00" and fortran_e(-6.480976, 20, 10) == "   -0.6480976000E+01"
    assert fortran_e(9.9999996e-5, 14, 6) == "  0.100000E-03"                       # rounding carries into the exponent
    assert _ld_real(-8.0) == "  -8.0000000000000000     " and _ld_real(0.5) == "  0.50000000000000000     "


def test_vti_writers_roundtrip(tmp_path):
    """files written in the reference's layout parse back (with the reference-golden parser) to the same numbers"""
    import sys
    sys.path.insert(0, GOLD)
    from make_golden import read_vti
    from gimic_b200 import grids, writers
    g = grids.std_grid([-1, -1, -1], [1, 0, 0], [0, 1, 0], [2, 2, 2], "even", spacing=[0.5, 1.0, 2.0])
    assert g.npts == (5, 3, 2)
    rng = np.random.default_rng(0)
    v = rng.normal(size=(g.n, 3)) * 1e-3
    s = rng.normal(size=g.n)
    writers.write_vti_vector(tmp_path / "jvec.vti", g, v)
    writers.write_vti_scalar(tmp_path / "jmod.vti", g, s)
    assert np.allclose(read_vti(str(tmp_path / "jvec.vti")), v, rtol=1e-5, atol=0)
    assert np.allclose(read_vti(str(tmp_path / "jmod.vti")), s, rtol=1e-5, atol=0)
    txt = open(tmp_path / "jvec.vti").read()
    assert 'WholeExtent="           0           4           0           2           0           1 "' in txt
    assert txt.count("\n") == 6 + g.n + 4 + (4 * 2 * 1) + 4          # header, vectors, CellData of (p1-1)(p2-1)(p3-1) cells, footer
    head = open(os.path.join(GOLD, "..", "golden", "open_shell_MOL")).readline()
    assert head.startswith("INTGRL")


def test_native_bulk_formatter_equals_the_python_edit_descriptor():
    """gimic_b200_format_e (threaded, std::to_chars) == fortran_e value by value: line grouping of the vti scalar block
    (break after value l when l % 4 == 0), 3 per line, prefixed vtu rows; zeros, three-digit exponents, rounding carries"""
    from gimic_b200.writers import fortran_e, format_e
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(size=20000) * 10.0 ** rng.integers(-30, 30, size=20000),
                        [0.0, -0.0, 1e-100, -3.5e120, 1.0, -1.0, 9.9999995e-5, 0.99999995, 123456.5, 0.1234565, 0.1234575]])
    for w, d, per_line, first, prefix in [(14, 6, 4, 1, ""), (14, 6, 3, 0, ""), (20, 10, 3, 0, "        "), (14, 6, 1, 0, ""), (20, 10, 1, 0, "  ")]:
        got = format_e(v, w, d, per_line, first, prefix).decode()
        lines, line = [], []
        for x in v:
            line.append(fortran_e(x, w, d))
            if len(line) == ((first or per_line) if not lines else per_line):
                lines.append(prefix + "".join(line) + "\n"); line = []
        assert got == "".join(lines) + (prefix + "".join(line) if line else ""), (w, d, per_line, first)
    assert format_e([], 14, 6, 3) == b"" and fortran_e(1e-101, 14, 6) == "  0.100000-100" and fortran_e(1e200, 8, 6) == "*" * 8
    # non-finite values and the ends of the double range: both formatters print them like gfortran and agree with each other
    odd = [float("nan"), float("inf"), -float("inf"), 5e-324, 2.2250738585072014e-308, 1.7976931348623157e308, -1.7976931348623157e308, 4.9e-310]
    for w, d in ((14, 6), (24, 16)):
        assert format_e(odd, w, d, 1).decode().split("\n")[:-1] == [fortran_e(x, w, d) for x in odd]
    assert fortran_e(float("nan"), 14, 6) == " " * 11 + "NaN" and fortran_e(-float("inf"), 14, 6) == "     -Infinity"
    assert fortran_e(5e-324, 14, 6) == "  0.494066-323"
    assert fortran_e(-0.0, 14, 6) == " -0.000000E+00" and format_e([-0.0, 0.0], 14, 6, 2) == b" -0.000000E+00  0.000000E+00\n"


def test_vti_appended_extra_holds_the_same_numbers(tmp_path):
    """--vtk appended (an extra, not a reference format): raw Float64 blocks at the offsets the header names"""
    import re
    from gimic_b200 import grids, writers
    g = grids.std_grid([-1, -1, -1], [1, 0, 0], [0, 1, 0], [2, 2, 2], "even", spacing=[0.5, 1.0, 2.0])
    rng = np.random.default_rng(1)
    v = rng.normal(size=(g.n, 3)); s = rng.normal(size=g.n)
    writers.write_vti_vector(tmp_path / "jvec.vti", g, v, appended=True)
    writers.write_vti_scalar(tmp_path / "jmod.vti", g, s, appended=True)
    for name, arrs in (("jvec.vti", [v.ravel(), None]), ("jmod.vti", [s])):
        raw = open(tmp_path / name, "rb").read()
        head, data = raw.split(b'<AppendedData encoding="raw">\n_', 1)
        offs = [int(x) for x in re.findall(rb'offset="(\d+)"', head)]
        assert len(offs) == len(arrs) and b'header_type="UInt64"' in head
        geo = re.search(rb'Origin="([^"]*)" Spacing="([^"]*)"', head)        # plain numbers a VTK reader can parse
        assert [float(x) for x in geo.group(1).split()] == [-1.0, -1.0, -1.0]
        assert [float(x) for x in geo.group(2).split()] == [0.5, 1.0, 2.0]
        for off, ref in zip(offs, arrs):
            nbytes = int(np.frombuffer(data[off:off + 8], np.uint64)[0])
            got = np.frombuffer(data[off + 8:off + 8 + nbytes], "<f8")
            if ref is not None:
                assert np.array_equal(got, ref)
            else:
                assert got.size == 4 * 2 * 1 and (got >= 0).all()          # cell-averaged |J|


def test_native_f_format_and_jmod_txt_writer(tmp_path):
    """gimic_b200_format_f == '%w.df' (Fortran Fw.d) incl. asterisks on overflow; jmod.txt ('(6f11.7)' rows, a blank line after
    each i-row on regular grids, jfield.f90:356-376,540) is byte-identical to a per-row reference implementation"""
    from gimic_b200 import grids, writers
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(size=3000) * 10.0 ** rng.integers(-9, 3, size=3000),
                        [0.0, 0.5, -0.5, 1e-8, -1e-8, 123.45678949999, 999.99999995, -999.99999995, 99999.9999999]])
    got = writers.format_f(v, 11, 7, 1).decode().split("\n")[:-1]
    assert got == [("%11.7f" % x) if len("%11.7f" % x) <= 11 else "*" * 11 for x in v]
    g = grids.std_grid([-1, -1, 0], [1, 0, 0], [0, 1, 0], [2, 2, 0], "gauss", grid_points=[9, 9, 0], gauss_order=9)
    vec = rng.normal(size=(g.n, 3)) * 1e-2
    r = g.points() * writers.AU2A; jm = np.sqrt((vec ** 2).sum(1))
    for regular in (True, False):
        writers.write_jmod_txt(tmp_path / "jmod.txt", g, vec, regular=regular)
        ref = ""
        for n in range(g.n):
            ref += "".join(f"{x:11.7f}" for x in (*r[n], jm[n])) + "\n"
            if regular and (n + 1) % g.npts[0] == 0:
                ref += "\n"
        assert open(tmp_path / "jmod.txt").read() == ref


def test_radius_keyword_masks_vectors_of_2d_bond_grids_like_the_reference():
    """jfield.f90:310-346: 2-D bond grid + radius keyword -> vectors beyond the radius are zeroed in jvec.vti, with the reference's
    Angstrom-vs-bohr comparison; no-op for the default radius (1e10), 3-D grids and base grids"""
    from gimic_b200 import grids, writers
    c1, c2, fix = np.array([0.0, 0.0, 0.0]), np.array([2.6, 0.0, 0.0]), np.array([1.3, 2.2, 0.0])
    g = grids.bond_grid(c1, c2, fix, 1.3, [-3.0, 3.0], [-3.0, 3.0], "even", spacing=[0.5, 0.5, 0.5], radius=1.5)
    v = np.ones((g.n, 3))
    m = writers.radius_masked_vectors(g, v)
    far = np.sqrt(((g.points() * writers.AU2A - g.center()) ** 2).sum(1)) > 1.5
    assert far.any() and (~far).any() and (m[far] == 0).all() and (m[~far] == 1).all() and (v == 1).all()
    g0 = grids.bond_grid(c1, c2, fix, 1.3, [-3.0, 3.0], [-3.0, 3.0], "even", spacing=[0.5, 0.5, 0.5])
    assert writers.radius_masked_vectors(g0, v) is v
    gb = grids.std_grid([-1, -1, -1], [1, 0, 0], [0, 1, 0], [2, 2, 2], "even", spacing=[0.5, 0.5, 0.5])
    vb = np.ones((gb.n, 3))
    assert writers.radius_masked_vectors(gb, vb) is vb


def test_dry_run_needs_no_device(tmp_path):
    """-y / dryrun=on (gimic.F90:150-204, src/gimic.in:53-54,139-140): basis geometry and grid only, no densities and
    here no device context; mol.xyz and grid.xyz are written, the run mode's banner is printed, nothing is calculated.
    Runs in this GPU-less container: proof that the dry run does not reach a compute entry point."""
    import shutil
    from gimic_b200 import driver
    shutil.copy(os.path.join(GOLD, "benzene_MOL"), tmp_path / "MOL")      # no XDENS on purpose
    shutil.copy(os.path.join(INPUTS, "benzene_integration-gauss.inp"), tmp_path / "gimic.inp")
    out = io.StringIO()
    d = driver.Driver(str(tmp_path / "gimic.inp"), out=out, dryrun=True)
    assert d.g is None
    d.run()
    text = out.getvalue()
    assert "Dry run, not calculating" in text and "Integrating current density" in text
    assert "Magnetic field <x,y,z>" not in text      # magnet_axis=X: get_magnet skips check_field and its printout (magnet.f90:60-63)
    assert "Induced current" not in text
    syms, coords = driver.read_mol_geometry(str(tmp_path / "MOL"))
    assert [s.strip() for s in d.symbols] == [s.strip() for s in syms] and np.allclose(d.xyz, coords, rtol=0, atol=0)
    assert (tmp_path / "mol.xyz").exists() and (tmp_path / "grid.xyz").exists()
    assert int(open(tmp_path / "mol.xyz").readline()) == len(syms)
    # the command-line switch
    os.remove(tmp_path / "mol.xyz")
    assert driver.main([str(tmp_path / "gimic.inp"), "--dryrun"]) == 0 and (tmp_path / "mol.xyz").exists()
    # without the switch the same input needs the device context and fails loudly here (no CPU fallback)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(Exception):
            driver.Driver(str(tmp_path / "gimic.inp"), out=io.StringIO())


def test_native_formatters_on_arbitrary_doubles():
    """property test (hypothesis): for any double, any reasonable Ew.d / Fw.d, gimic_b200_format_e / _f print what the
    per-value Python edit descriptors print (correct rounding of the exact binary value, exponent carries, three-digit
    exponents, asterisks on overflow, non-finite values)"""
    from hypothesis import given, settings, strategies as st
    from gimic_b200.writers import fortran_e, format_e, format_f

    @settings(max_examples=300, deadline=None)
    @given(st.lists(st.floats(allow_nan=True, allow_infinity=True, width=64), min_size=1, max_size=40),
           st.integers(min_value=1, max_value=17), st.integers(min_value=0, max_value=9), st.integers(min_value=1, max_value=5))
    def check(vals, d, extra, per_line):
        w = d + 7 + extra
        got = format_e(vals, w, d, per_line).decode()
        toks = [got_line[i:i + w] for got_line in got.split("\n") for i in range(0, len(got_line), w)]
        assert toks == [fortran_e(x, w, d) for x in vals]
        finite = [x for x in vals if x == x and abs(x) < 1e15]
        if finite:
            gf = format_f(finite, w, d, 1).decode().split("\n")[:-1]
            assert gf == [("%*.*f" % (w, d, x)) if len("%*.*f" % (w, d, x)) <= w else "*" * w for x in finite]
    check()


def test_report_sink_prints_nan_like_gfortran():
    """a NaN in a report line (e.g. an integral over a grid with one point on an axis, where the reference's step is l/0) reads 'NaN' as
    gfortran's F edit writes it, at the same width; words containing 'nan' are left alone"""
    import io
    from gimic_b200.driver import _GfortranNaN
    buf = io.StringIO()
    out = _GfortranNaN(buf)
    out.write(f" Induced current (au)    :{float('nan'):14.6f}\n resonance nanoring\n")
    assert buf.getvalue() == " Induced current (au)    :           NaN\n resonance nanoring\n"
    assert out.getvalue() == buf.getvalue()                      # everything else is the wrapped stream's


@pytest.mark.filterwarnings("ignore")
def test_random_grids_equal_the_oracle_grids():
    """200 random std / bond grids (even, gauss, lobatto; grid_points or spacing; rotation, rotation_origin, radius): the product's
    grid code gives the oracle's points (1e-11 bohr), weights and fields for magnet_axis = X, k, -k, z, -x, y.  Not compared: magnet_axis
    = i / j on such grids -- check_field (magnet.f90:75) then tests the sign of a dot product of two orthogonal vectors, i.e. of
    rounding noise, in the reference as well; the two drivers agree with each other there (same summation order), the oracle need not."""
    from gimic_b200 import grids
    rng = np.random.default_rng(1); N = 200; bad = 0
    for k in range(N):
        typ=str(rng.choice(["even","gauss","lobatto"])); order=int(rng.integers(2,12))
        kw={}
        if typ!="even" or rng.random()<0.5: kw["grid_points"]=[int(rng.integers(2,25)),int(rng.integers(2,25)),0]
        else: kw["spacing"]=list(rng.uniform(0.2,1.5,size=3))
        if rng.random()<0.5:
            kw["rotation"]=list(rng.uniform(-90,90,size=3))
            if rng.random()<0.5: kw["rotation_origin"]=list(rng.normal(size=3))
        try:
            if rng.random()<0.5:
                c1,c2,fix=rng.normal(size=3)*2,rng.normal(size=3)*2+1,rng.normal(size=3)*3
                d=float(rng.uniform(0.1,2)); h=[-float(rng.uniform(0.5,5)),float(rng.uniform(0.5,5))]; w=[-float(rng.uniform(0.5,5)),float(rng.uniform(0.5,5))]
                rad=float(rng.uniform(1,4)) if rng.random()<0.3 else None
                g=grids.bond_grid(c1,c2,fix,d,h,w,gtype=typ,gauss_order=order,radius=rad,**kw)
                o=O.grid_bond(c1,c2,fix,d,height=h,width=w,radius=rad,type=typ,gauss_order=order,**kw)
            else:
                org=rng.uniform(-5,0,size=3); i=rng.normal(size=3); j=np.cross(i,rng.normal(size=3)); L=rng.uniform(1,8,size=3)
                if "grid_points" in kw: kw["grid_points"][2]=int(rng.integers(0,8))
                if "grid_points" in kw and kw["grid_points"][2]==1: kw["grid_points"][2]=2
                g=grids.std_grid(org,i,j,L,gtype=typ,gauss_order=order,**kw)
                o=O.grid_std(org,i,j,L,type=typ,gauss_order=order,**kw)
        except Exception as e:
            print("EXC",k,repr(e)[:200]); bad+=1; continue
        ok = list(g.npts)==list(o.npts)
        if ok:
            pg,po=g.points(),o.points()
            ok = np.allclose(pg,po,rtol=0,atol=1e-11)
            for d in range(3):
                pw=o.axis(d); ok = ok and np.allclose(g.wgt[d],pw[1],rtol=1e-12,atol=1e-14)
            for ax in ("X","k","-k","z","-x","y"):
                ok = ok and np.allclose(grids.get_magnet(g,ax,None),o.magnet(ax),atol=1e-12)
        if not ok: bad+=1; print("MISMATCH",k,typ,kw,list(g.npts),list(o.npts))
    assert bad == 0

