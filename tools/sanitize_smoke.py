"""Small tensor + J-path + integral run for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures
import gimic_b200


def gauss_plane(origin, basv, l0, l1, n0, n1, order=9):
    """n0 x n1 Gauss-Legendre points over [0, l0] x [0, l1] in the plane of basv[0], basv[1]"""
    p0, w0, p1, w1 = np.zeros(n0), np.zeros(n0), np.zeros(n1), np.zeros(n1)
    gimic_b200.gausspoints(0.0, l0, order, p0, w0); gimic_b200.gausspoints(0.0, l1, order, p1, w1)
    return gimic_b200.Grid(origin, basv, [p0, p1, np.zeros(1)], [w0, w1, np.ones(1)])


cases = fixtures.materialize(tempfile.mkdtemp())
g = gimic_b200.Gimic(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-8)
rng = np.random.default_rng(0)
r = np.vstack([rng.uniform(-5, 5, size=(300, 3)), rng.uniform(-40, 40, size=(40, 3))])
B = np.array([0.0, 0.0, 1.0])
t = g.jtensors(r)
f = g.fields(r, B, "total", jvec=True, jmod=True, edens=True)
f2 = g.fields(r, B, "total", tens=True, jvec=True, jmod=True, acid=True)
xyz = g.atom_coords()
gr = gauss_plane(0.5 * (xyz[0] + xyz[1]) - [0.0, 1.25614, 5.0], [[0, 1, 0], [0, 0, 1], [1, 0, 0]], 7.25614, 10.0, 9, 9)
s = g.integrate(gr, B, "total", 7)
sh, dens, nbf = fixtures.synthetic_case(6, "flake", seed=1)
g2 = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fixtures.dens_to_colmajor(dens), giao=False, **sh)
t2 = g2.jtensors(r[:200])
print("ok", float(np.abs(t).max()), float(np.abs(f["jvec"] - f2["jvec"]).max()), s[:3], float(np.abs(t2).max()))
