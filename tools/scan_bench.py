"""Current-profile scan (jobscripts/src/current-profile-*): N thin slices of one bond plane, each a Gauss grid of 9 x 36 points.
Compares one gimic_b200_integrate call per slice with ONE gimic_b200_integrate_batch call (all slices in one tensor pass).
The reference starts a gimic process per slice (MOL/XDENS re-read each time; c4h4/integration takes 2.13 s for one plane)."""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures
import gimic_b200
from gimic_b200 import grids

cases = fixtures.materialize(tempfile.mkdtemp())
g = gimic_b200.Gimic(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-8)
xyz = g.atom_coords()
out = {}
for nsl in (50, 400):
    edges = np.linspace(-1.25614, 6.0, nsl + 1)
    gs = [grids.bond_grid(xyz[1], xyz[0], xyz[3], 1.48794, [-5.0, 5.0], [edges[i], edges[i + 1]], "gauss", grid_points=[30, 9, 0], gauss_order=9)
          for i in range(nsl)]
    B = np.array([0.0, 0.0, 1.0])
    g.integrate_batch(gs[:4], B, "total", 3); g.integrate(gs[0], B, "total", 3)
    t0 = time.perf_counter(); single = np.array([g.integrate(x, B, "total", 3) for x in gs]); t1 = time.perf_counter()
    batch = g.integrate_batch(gs, B, "total", 3); t2 = time.perf_counter()
    out[f"{nsl}_slices"] = {"points_per_slice": gs[0].n, "loop_of_integrate_s": t1 - t0, "integrate_batch_s": t2 - t1,
                            "max_abs_diff": float(np.abs(single - batch).max()), "total_current_au": float(batch[:, 0].sum())}
print(json.dumps(out, indent=1))
