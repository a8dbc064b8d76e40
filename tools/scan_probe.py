"""Where the time of one gimic_b200_integrate_batch call goes (current-profile scan, c4h4, nbf = 168): profiled stage times."""
import os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
import gimic_b200


def gauss_plane(origin, basv, l0, l1, n0, n1, order=9):
    p0, w0, p1, w1 = np.zeros(n0), np.zeros(n0), np.zeros(n1), np.zeros(n1)
    gimic_b200.gausspoints(0.0, l0, order, p0, w0); gimic_b200.gausspoints(0.0, l1, order, p1, w1)
    return gimic_b200.Grid(origin, basv, [p0, p1, np.zeros(1)], [w0, w1, np.ones(1)])


cases = fixtures.materialize(tempfile.mkdtemp())
g = gimic_b200.Gimic(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-8)
xyz = g.atom_coords()
B = np.array([0.0, 0.0, 1.0])
for nsl in (50, 400):
    edges = np.linspace(-1.25614, 6.0, nsl + 1)
    mid = 0.5 * (xyz[0] + xyz[1])
    gs = [gauss_plane(mid + [0.0, edges[i], -5.0], [[0, 1, 0], [0, 0, 1], [1, 0, 0]], edges[i + 1] - edges[i], 10.0, 9, 36) for i in range(nsl)]
    for what in (3, 7):
        g.integrate_batch(gs, B, "total", what)
        t0 = time.perf_counter(); g.integrate_batch(gs, B, "total", what); t1 = time.perf_counter()
        g.set_profiling(True); g.integrate_batch(gs, B, "total", what); st = g.stats(); g.set_profiling(False)
        print(f"{nsl} slices what={what}: wall {1e3 * (t1 - t0):.2f} ms; tiles {st['n_tiles']} mean nact {st['sum_nact'] / max(st['n_tiles'], 1):.0f}; "
              f"profiled ms: sort {st['ms_sort']:.3f} tiles {st['ms_tiles']:.3f} basis {st['ms_basis']:.3f} contract {st['ms_contract']:.3f}; launches {st['launches']}")
    r = np.vstack([x.points().reshape(-1, 3) for x in gs])
    g.fields(r, B, "total", jvec=True)
    t0 = time.perf_counter(); g.fields(r, B, "total", jvec=True); t1 = time.perf_counter()
    print(f"   the same {r.shape[0]} points through calc_fields(jvec): {1e3 * (t1 - t0):.2f} ms")
