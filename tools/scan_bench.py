"""Current-profile scan (jobscripts/src/current-profile-*): N thin slices of one bond plane, each a Gauss grid of 9 x 36 points.
Compares one gimic_b200_integrate call per slice with ONE gimic_b200_integrate_batch call (all slices in one tensor pass).
The reference starts a gimic process per slice (MOL/XDENS re-read each time; c4h4/integration takes 2.13 s for one plane)."""
import json, os, sys, tempfile, time
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
xyz = g.atom_coords()
out = {}
for nsl in (50, 400):
    edges = np.linspace(-1.25614, 6.0, nsl + 1)
    mid = 0.5 * (xyz[0] + xyz[1])
    gs = [gauss_plane(mid + [0.0, edges[i], -5.0], [[0, 1, 0], [0, 0, 1], [1, 0, 0]], edges[i + 1] - edges[i], 10.0, 9, 36) for i in range(nsl)]
    B = np.array([0.0, 0.0, 1.0])
    g.integrate_batch(gs, B, "total", 3); g.integrate(gs[0], B, "total", 3)      # warm: workspace growth, lazily loaded kernels (0.9 s the first time)
    t0 = time.perf_counter(); single = np.array([g.integrate(x, B, "total", 3) for x in gs]); t1 = time.perf_counter()
    batch = g.integrate_batch(gs, B, "total", 3); t2 = time.perf_counter()
    out[f"{nsl}_slices"] = {"points_per_slice": gs[0].n, "loop_of_integrate_s": t1 - t0, "integrate_batch_s": t2 - t1,
                            "max_abs_diff": float(np.abs(single - batch).max()), "total_current_au": float(batch[:, 0].sum())}
print(json.dumps(out, indent=1))
