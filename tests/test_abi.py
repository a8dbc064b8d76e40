"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the header
declares, its host-only entry points agree with the oracle, and compute entry points fail loudly
(no silent CPU fallback) when there is no GPU."""
import ctypes as C
import os
import re
import numpy as np
import pytest

import oracle_lib as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def L():
    import __graft_entry__ as ge
    from gimic_b200 import _lib
    if not os.path.exists(_lib.SO_PATH):
        ge.build()
    return _lib.lib()


def test_header_symbols_exported(L):
    from gimic_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "gimic_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b((?:gimic_|mkgausspoints)\w*)\s*\(", hdr))
    assert names, "no prototypes found"
    for n in sorted(names):
        assert hasattr(L, n), f"{n} declared in include/gimic_b200.h but not exported"
    assert names == set(_lib.LEGACY_SYMBOLS + _lib.API_SYMBOLS)


def test_legacy_signatures_match_reference_header():
    # src/libgimic/gimic_interface.h:9-18 and gausspoints.h:10-13 (names only; signatures are cited in the header)
    want = ["gimic_init", "gimic_finalize", "gimic_set_uhf", "gimic_set_magnet", "gimic_set_spin", "gimic_set_screening",
            "gimic_calc_jtensor", "gimic_calc_jvector", "gimic_calc_modj", "gimic_get_gauss_points", "mkgausspoints"]
    from gimic_b200 import _lib
    assert _lib.LEGACY_SYMBOLS == want


@pytest.mark.parametrize("npts,order,quadr", [(36, 9, "gauss"), (7, 7, "gauss"), (30, 5, "gauss"), (1, 9, "gauss"),
                                              (27, 9, "lobatto"), (13, 13, "lobatto")])
def test_gauss_points_match_oracle(L, npts, order, quadr):
    import gimic_b200
    p = np.zeros(npts); w = np.zeros(npts)
    gimic_b200.gausspoints(0.0, 7.25614, order, p, w, quadrature=quadr)
    po, wo = O.gauss_points(0.0, 7.25614, npts, order, quadr)
    assert np.allclose(p, po, rtol=0, atol=1e-14) and np.allclose(w, wo, rtol=0, atol=1e-14)


@pytest.mark.parametrize("turbomole", [False, True])
def test_c2s_rows_match_oracle(L, turbomole):
    """spherical=on projection (cao2sao.f90:163-231): the product's integer construction == the oracle's statement-by-statement
    restatement (float coefficients + renorm), l = 0..5, both cartesian component orders; includes the reference's
    store-instead-of-accumulate quirk in the |m| >= 2 rows of g and h shells"""
    from gimic_b200.gimic import c2s_rows
    sh = dict(coords=np.zeros((1, 3)), nctr_per_atom=np.array([1], np.int32), ctr_l=np.array([0], np.int32),
              ctr_npf=np.array([1], np.int32), xp=np.array([1.0]), cc=np.array([1.0]))
    o = O.Oracle.from_arrays(dens_a=np.zeros(4), turbomole_order=turbomole, **sh)
    for l in range(6):
        got, ref = c2s_rows(l, turbomole), o.c2s(l)
        assert got.shape == ref.shape and np.array_equal(got, np.round(ref)) and np.abs(ref - np.round(ref)).max() < 1e-9, l
    g4 = c2s_rows(4, False)   # standard order: component 3 is x^2 y^2; S_4,2 ~ (x^2-y^2)(6z^2-x^2-y^2) has no x^2y^2 term
    assert g4[4 + 2, 3] == -1.0   # ... but the reference stores -1 there (last (u,v) term wins, cao2sao.f90:188)


def test_xdens_text_parser_and_binary_cache(L, tmp_path):
    """XDENS reader (dens.f90:129-135 list-directed reals): Fortran D exponents, leading '+', commas, blank lines, CRLF;
    threaded parse == Python float(); the binary cache holds the parsed values bit for bit; size mismatches are errors"""
    import gimic_b200
    nbf = 3
    vals = np.random.default_rng(0).uniform(-1, 1, 4 * nbf * nbf)
    toks = []
    for i, v in enumerate(vals):
        toks.append(["%.14E" % v, ("%.14E" % v).replace("E", "D"), "  +%.14e" % abs(v), "%.16g," % v, "\n%.17g\r" % v,
                     ("%.10e" % v).replace("e", "d")][i % 6])
    (tmp_path / "X").write_text("\n".join(toks) + "\n")
    gimic_b200.convert_xdens(tmp_path / "X", nbf, tmp_path / "X.bin")
    raw = (tmp_path / "X.bin").read_bytes()
    assert raw[:8] == b"GB2XDENS" and np.frombuffer(raw[8:24], np.int64).tolist() == [nbf, 4]
    ref = np.array([float(t.replace("D", "e").replace("d", "e").replace(",", "").strip()) for t in toks])
    assert np.array_equal(np.frombuffer(raw[24:], np.float64), ref)
    # a file large enough to be cut into per-thread pieces
    n = 260
    big = np.random.default_rng(1).uniform(-1, 1, 4 * n * n)
    (tmp_path / "B").write_text("\n".join("%.14E" % v for v in big) + "\n")
    gimic_b200.convert_xdens(tmp_path / "B", n, tmp_path / "B.bin")
    got = np.frombuffer((tmp_path / "B.bin").read_bytes()[24:], np.float64)
    assert np.array_equal(got, np.array([float("%.14E" % v) for v in big]))
    with pytest.raises(gimic_b200.GimicB200Error, match="too short"):
        gimic_b200.convert_xdens(tmp_path / "X", nbf + 1, tmp_path / "Y.bin")
    (tmp_path / "bad").write_text("0.1\nabc\n" * 18)
    with pytest.raises(gimic_b200.GimicB200Error, match="malformed"):
        gimic_b200.convert_xdens(tmp_path / "bad", nbf, tmp_path / "Y.bin")


def test_legacy_gauss_entry(L):
    a, b, n, o = C.c_double(0.0), C.c_double(2.0), C.c_int(18), C.c_int(9)
    p = np.zeros(18); w = np.zeros(18)
    L.mkgausspoints(C.byref(a), C.byref(b), C.byref(n), C.byref(o), p.ctypes.data_as(C.POINTER(C.c_double)),
                    w.ctypes.data_as(C.POINTER(C.c_double)))
    po, wo = O.gauss_points(0.0, 2.0, 18, 9)
    assert np.allclose(p, po, atol=1e-14) and np.allclose(w, wo, atol=1e-14)


def test_gauss_bad_order_is_an_error(L):
    import gimic_b200
    with pytest.raises(gimic_b200.GimicB200Error):
        gimic_b200.gausspoints(0.0, 1.0, 9, np.zeros(10), np.zeros(10))


def test_io_errors_are_reported(L, cases):
    import gimic_b200
    with pytest.raises(gimic_b200.GimicB200Error) as e:
        gimic_b200.Gimic(cases["c4h4"]["mol"], "/nonexistent/XDENS")
    assert e.value.code == -2 and "Density file not found" in str(e.value)
    with pytest.raises(gimic_b200.GimicB200Error) as e:
        gimic_b200.Gimic(cases["c4h4"]["xdens"], cases["c4h4"]["xdens"])
    assert e.value.code == -2 and "INTGRL" in str(e.value)


def test_no_cpu_fallback(L, cases):
    """Without a CUDA device the product must fail loudly, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import gimic_b200
    with pytest.raises(gimic_b200.GimicB200Error) as e:
        gimic_b200.Gimic(cases["c4h4"]["mol"], cases["c4h4"]["xdens"])
    assert e.value.code == -3 and "no CPU path" in str(e.value)


def test_product_does_not_touch_the_oracle():
    """oracle/ is test infrastructure: nothing under gimic_b200/ or include/ may reference it."""
    bad = []
    for base in ("gimic_b200", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                    txt = open(os.path.join(dp, f), errors="replace").read()
                    if re.search(r"oracle_lib|libgimic_oracle|gimic_oracle|\boracle/", txt):
                        bad.append(os.path.join(dp, f))
    assert not bad, bad


def test_header_is_plain_c_and_links(L, tmp_path):
    """include/gimic_b200.h compiles as C99 (no C++/torch types in the boundary) and a C caller of the legacy symbols
    (what GimicInterface.cpp / gimic.pyx bind, gimic_interface.h:9-18) links against libgimic_b200.so; the host-only
    entry points run without a GPU"""
    import subprocess
    from gimic_b200 import _lib
    src = tmp_path / "caller.c"
    src.write_text(r'''
#include <stdio.h>
#include "gimic_b200.h"
int main(void) {
    double a = 0.0, b = 2.0, pts[18], wgts[18];
    int npts = 18, order = 9;
    void (*legacy[])(void) = {(void (*)(void))gimic_init, (void (*)(void))gimic_finalize, (void (*)(void))gimic_set_uhf,
        (void (*)(void))gimic_set_magnet, (void (*)(void))gimic_set_spin, (void (*)(void))gimic_set_screening,
        (void (*)(void))gimic_calc_jtensor, (void (*)(void))gimic_calc_jvector, (void (*)(void))gimic_calc_modj};
    gimic_b200_opts o;
    gimic_b200_default_opts(&o);
    mkgausspoints(&a, &b, &npts, &order, pts, wgts);
    double rows[5 * 6];
    if (gimic_b200_c2s_rows(2, 0, rows)) return 2;
    printf("%d %d %.15f %.15f %g %s\n", (int)(sizeof legacy / sizeof legacy[0]), o.giao, pts[0], wgts[17], rows[2 * 6 + 0], gimic_b200_version());
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["gcc", "-std=c99", "-pedantic", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libgimic_b200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe)], text=True).split()
    po, wo = O.gauss_points(0.0, 2.0, 18, 9)
    assert out[0] == "9" and out[1] == "1" and abs(float(out[2]) - po[0]) < 1e-14 and abs(float(out[3]) - wo[17]) < 1e-14
    assert float(out[4]) == -1.0            # d shell, m = 0 row: 2zz - xx - yy -> coefficient of xx


def test_cpp_wrapper_mirrors_gimicinterface(L, tmp_path, cases):
    """include/GimicB200Interface.h: the method set of the reference's GimicInterface (GimicInterface.h:4-15) over the handle API; compiles as
    strict C++11, links, and -- in a container without a GPU -- construction fails loudly with the library's message (no CPU fallback)"""
    import subprocess
    import torch
    from gimic_b200 import _lib
    src = tmp_path / "caller.cpp"
    src.write_text(r'''
#include <cstdio>
#include <exception>
#include "GimicB200Interface.h"
int main(int argc, char **argv) {
    try {
        GimicB200Interface g(argv[1], argv[2]);
        const double b[3] = {0.0, 0.0, 1.0}, r[6] = {0.1, 0.2, 0.3, 1.0, -0.5, 0.25};
        double jt[9], jv[3], jvs[6], mj;
        g.set_magnet(b); g.set_spin("total");
        g.calc_jtensor(r, jt); g.calc_jvector(r, jv); g.calc_modj(r, &mj); g.calc_jvectors(2, r, jvs);
        double want = 0.0;
        for (int m = 0; m < 3; ++m) want += (jt[m + 6] - jv[m]) * (jt[m + 6] - jv[m]) + (jvs[m] - jv[m]) * (jvs[m] - jv[m]);   // J = T.B, B = e_z
        std::printf("ok nbf=%d natoms=%d consistent=%d modj=%.12e\n", g.nbf(), g.natoms(), want < 1e-18 * mj * mj || want < 1e-22, mj);   /* tensor path vs J path: 1e-10 relative / 1e-12 absolute per component */
        try { g.set_spin("sideways"); } catch (const std::invalid_argument &e) { std::printf("spin refused\n"); }
    } catch (const std::exception &e) {
        std::printf("refused: %s\n", e.what());
    }
    (void)argc;
    return 0;
}
''')
    exe = tmp_path / "caller"
    libdir = os.path.dirname(_lib.SO_PATH)
    subprocess.check_call(["g++", "-std=c++11", "-pedantic", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe),
                           "-L", libdir, "-l:libgimic_b200.so", f"-Wl,-rpath,{libdir}"])
    out = subprocess.check_output([str(exe), cases["c4h4"]["mol"], cases["c4h4"]["xdens"]], text=True)
    if torch.cuda.is_available():
        assert out.startswith("ok nbf=168 natoms=8 consistent=1") and "spin refused" in out, out
    else:
        assert out.startswith("refused: gimic_b200: no CUDA device available"), out


def test_product_is_independent_of_the_oracle(L):
    """oracle/ is test infrastructure: no product source names it (neither an import, an include, a path nor a link line), the
    built libraries and the program have no dependency on libgimic_oracle.so, and importing the package does not load it"""
    import subprocess
    import sys
    pkg = os.path.join(ROOT, "gimic_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) and f != "Makefile":
                continue
            text = open(os.path.join(d, f), errors="replace").read()
            assert not re.search(r"oracle[_/.](?:lib|gimic|_ref)|libgimic_oracle|import\s+oracle|from\s+oracle|\.\./oracle|oracle/", text), os.path.join(d, f)
    for h in os.listdir(os.path.join(ROOT, "include")):
        assert "oracle/" not in open(os.path.join(ROOT, "include", h), errors="replace").read(), h
    for name in ("libgimic_b200.so", "libgimic_b200_driver.so", "gimic-b200"):
        dyn = subprocess.check_output(["readelf", "-d", os.path.join(pkg, name)], text=True)
        assert "oracle" not in dyn, name
    code = ("import sys; sys.path.insert(0, %r); import gimic_b200, gimic_b200.driver, gimic_b200.gimic; "
            "maps = open('/proc/self/maps').read(); assert 'libgimic_oracle' not in maps; "
            "assert not [m for m in sys.modules if 'oracle' in m], [m for m in sys.modules if 'oracle' in m]") % ROOT
    subprocess.check_call([sys.executable, "-c", code])
