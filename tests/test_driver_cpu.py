"""CPU tests of the driver's host layer through the driver library's own entry points: grid geometry
and field direction (gimic_b200_input_grid) against the oracle AND against what the reference printed for its benzene keyword
tests; the bulk number formatters against per-value restatements of the Fortran edit descriptors; the dry run.  The gimic.inp
reader used to configure the oracle side (tests/inp_reader.py) is test infrastructure and is checked here too."""
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
    import inp_reader
    return inp_reader.parse_file(os.path.join(INPUTS, name + ".inp"))


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
    import inp_reader as inp
    base = 'calc=cdens\nmagnet_axis=z\nGrid(std){ origin=[0,0,0]\n ivec=[1,0,0]\n jvec=[0,1,0]\n lengths=[1,1,1]\n spacing=[0.5,0.5,0.5] }\n'
    assert inp.parse_text(base).get("Grid.lengths") == [1.0, 1.0, 1.0]
    for bad in (base.replace("calc=cdens", "calc=foo"), base.replace("magnet_axis=z", ""), base + "magnet=[0,0,1]\n",
                base.replace("spacing=[0.5,0.5,0.5]", ""), base.replace("Grid(std)", "Grid(cube)"),
                base.replace("spacing=", "grid_points=[3,3,3]\n spacing=")):
        with pytest.raises(inp.InputError):
            inp.parse_text(bad)
    assert inp.parse_text(base.replace("calc=cdens", "calc=cdens # comment\ntitle=\"a # b\"")).get("title") == "a # b"


def _grid_from(tmp, name, mol):
    """(parsed input, product grid, field direction, info) of a reference input: the grid comes from the driver library"""
    import shutil
    from gimic_b200 import driver
    d = tmp / name
    d.mkdir()
    shutil.copy(os.path.join(GOLD, mol), d / "MOL")
    shutil.copy(os.path.join(INPUTS, name + ".inp"), d / "gimic.inp")
    g, mag, info = driver.input_grid(str(d / "gimic.inp"))
    return _inp(name), g, mag, info


@pytest.mark.parametrize("name,mol", [("c4h4_integration", "c4h4_MOL"), ("open-shell_integration", "open_shell_MOL"),
                                      ("open-shell_3d", "open_shell_MOL"), ("benzene_int-grid-bond-even", "benzene_MOL"),
                                      ("benzene_keyword-rotation", "benzene_MOL"), ("benzene_keyword-rotation_origin", "benzene_MOL"),
                                      ("benzene_keyword-radius", "benzene_MOL"), ("benzene_keyword-spacing", "benzene_MOL"),
                                      ("benzene_keyword-magnet", "benzene_MOL"), ("benzene_integration-lobatto", "benzene_MOL"),
                                      ("benzene_3d", "benzene_MOL"), ("benzene_2d", "benzene_MOL"), ("benzene_int-cdens", "benzene_MOL")])
def test_grid_and_magnet_match_oracle(tmp_path, name, mol):
    from gimic_b200.driver import mol_geometry
    from oracle_grid import oracle_grid
    _, coords = mol_geometry(os.path.join(GOLD, mol))
    I, g, mag, info = _grid_from(tmp_path, name, mol)
    og = oracle_grid(I, coords)
    assert g.npts == tuple(og.npts)
    assert np.allclose(g.origin, og.origin, rtol=0, atol=1e-13) and np.allclose(g.basv, og.basv, rtol=0, atol=1e-14)
    for d in range(3):
        p, w = og.axis(d)
        assert np.allclose(g.pts[d], p, rtol=0, atol=1e-13) and np.allclose(g.wgt[d], w, rtol=0, atol=1e-14)
    assert np.allclose(g.points().reshape(-1, 3), og.points(), rtol=0, atol=1e-12)
    assert np.allclose(mag, og.magnet(I.get("magnet_axis"), I.get("magnet")), rtol=0, atol=1e-14)
    if I.grid_arg == "bond":
        assert g.radius == og.radius == info["radius"]


def test_grid_matches_what_the_reference_printed(tmp_path):
    """'Integration grid data' block, point counts and field direction of the reference's benzene stdout goldens"""
    gold = json.load(open(os.path.join(GOLD, "benzene_grids.json")))
    checked = 0
    for name, ref in gold.items():
        if not os.path.exists(os.path.join(INPUTS, f"benzene_{name}.inp")):
            continue
        I, g, mag, info = _grid_from(tmp_path, f"benzene_{name}", "benzene_MOL")
        if "npts" in ref:
            assert list(g.npts) == ref["npts"], name
        if ref.get("magnet"):
            assert np.allclose(mag, ref["magnet"], atol=1e-5), name
        geo = ref["geometry"]
        if geo and not I.is_set("Grid.rotation"):      # the block is printed before the rotation is applied
            assert np.allclose(info["center_bond"], geo["center"], atol=1e-6), name
            assert np.allclose(g.origin, geo["origin"], atol=1e-6), name
            for v in range(3):
                assert np.allclose(g.basv[v], geo[f"basv{v + 1}"], atol=1e-6), name
            assert np.allclose(info["lengths"], geo["lenghts"], atol=1e-6), name
        checked += 1
    assert checked >= 8


def test_fortran_number_formats():
    """the per-value restatements of Ew.d and list-directed real(8) (tests/fortran_fmt.py) on values taken from the reference's files"""
    from fortran_fmt import fortran_e, ld_real as _ld_real
    assert fortran_e(0.648806e-13, 14, 6) == "  0.648806E-13" and fortran_e(-0.671910e-13, 14, 6) == " -0.671910E-13"
    assert fortran_e(0.0, 14, 6) == "  0.000000E+00" and fortran_e(-6.480976, 20, 10) == "   -0.6480976000E+01"
    assert fortran_e(9.9999996e-5, 14, 6) == "  0.100000E-03"                       # rounding carries into the exponent
    assert _ld_real(-8.0) == "  -8.0000000000000000     " and _ld_real(0.5) == "  0.50000000000000000     "


def test_native_bulk_formatter_equals_the_python_edit_descriptor():
    """gimic_b200_format_e (threaded, std::to_chars) == fortran_e value by value: line grouping of the vti scalar block
    (break after value l when l % 4 == 0), 3 per line, prefixed vtu rows; zeros, three-digit exponents, rounding carries"""
    from fortran_fmt import fortran_e, format_e
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


def test_native_f_format():
    """gimic_b200_format_f == '%w.df' (Fortran Fw.d) incl. asterisks on overflow"""
    from fortran_fmt import format_f
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.normal(size=3000) * 10.0 ** rng.integers(-9, 3, size=3000),
                        [0.0, 0.5, -0.5, 1e-8, -1e-8, 123.45678949999, 999.99999995, -999.99999995, 99999.9999999]])
    got = format_f(v, 11, 7, 1).decode().split("\n")[:-1]
    assert got == [("%11.7f" % x) if len("%11.7f" % x) <= 11 else "*" * 11 for x in v]


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
    d.run()
    text = out.getvalue()
    assert text == d.report and "Dry run, not calculating" in text and "Integrating current density" in text
    assert "Magnetic field <x,y,z>" not in text      # magnet_axis=X: get_magnet skips check_field and its printout (magnet.f90:60-63)
    assert "Induced current" not in text
    syms, coords = driver.mol_geometry(str(tmp_path / "MOL"))
    assert (tmp_path / "mol.xyz").exists() and (tmp_path / "grid.xyz").exists()
    xyz = open(tmp_path / "mol.xyz").read().split("\n")
    assert int(xyz[0]) == len(syms) == 12 and [l.split()[0] for l in xyz[2:14]] == [s.strip() for s in syms]
    assert np.allclose([[float(t) for t in l.split()[1:4]] for l in xyz[2:14]], coords * 0.52917726, atol=1e-6)      # Angstrom
    # the command-line switch
    os.remove(tmp_path / "mol.xyz")
    assert driver.main([str(tmp_path / "gimic.inp"), "--dryrun"]) == 0 and (tmp_path / "mol.xyz").exists()
    # without the switch the same input needs the device context and fails loudly here (no CPU fallback)
    import torch
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            driver.Driver(str(tmp_path / "gimic.inp"), out=io.StringIO()).run()


def test_native_formatters_on_arbitrary_doubles():
    """property test (hypothesis): for any double, any reasonable Ew.d / Fw.d, gimic_b200_format_e / _f print what the
    per-value Python edit descriptors print (correct rounding of the exact binary value, exponent carries, three-digit
    exponents, asterisks on overflow, non-finite values)"""
    from hypothesis import given, settings, strategies as st
    from fortran_fmt import fortran_e, format_e, format_f

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


