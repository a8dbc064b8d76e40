"""Small tensor + J-path + integral run for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures
import gimic_b200
from gimic_b200 import grids

cases = fixtures.materialize(tempfile.mkdtemp())
g = gimic_b200.Gimic(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-8)
rng = np.random.default_rng(0)
r = np.vstack([rng.uniform(-5, 5, size=(300, 3)), rng.uniform(-40, 40, size=(40, 3))])
B = np.array([0.0, 0.0, 1.0])
t = g.jtensors(r)
f = g.fields(r, B, "total", jvec=True, jmod=True, edens=True)
f2 = g.fields(r, B, "total", tens=True, jvec=True, jmod=True, acid=True)
xyz = g.atom_coords()
gr = grids.bond_grid(xyz[1], xyz[0], xyz[3], 1.48794, [-5.0, 5.0], [-1.25614, 6.0], "gauss", grid_points=[9, 9, 0], gauss_order=9)
s = g.integrate(gr, B, "total", 7)
sh, dens, nbf = fixtures.synthetic_case(6, "flake", seed=1)
g2 = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fixtures.dens_to_colmajor(dens), giao=False, **sh)
t2 = g2.jtensors(r[:200])
print("ok", float(np.abs(t).max()), float(np.abs(f["jvec"] - f2["jvec"]).max()), s[:3], float(np.abs(t2).max()))
