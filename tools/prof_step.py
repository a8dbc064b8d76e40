"""Small driver for ncu captures: one warm + one measured tensor pass over a chunk of the bench workload."""
import argparse, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gimic_b200
from gimic_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--natoms", type=int, default=278)
ap.add_argument("--grid", type=int, default=256)
ap.add_argument("--points", type=int, default=148 * 8 * 128)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--no-giao", action="store_true")
ap.add_argument("--jvec", action="store_true", help="J = T.B path (fields with jvec only) instead of the tensor path")
a = ap.parse_args()
sh, dens, nbf, origin, basv, pts = bench.build_workload(a.natoms, a.grid)
g = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=synthetic.dens_to_colmajor(dens), giao=not a.no_giao, **sh)
r = bench.slab_points(origin, basv, pts, 0)
r = np.ascontiguousarray(r[-a.points:])   # the planes of octant 0 closest to the molecular plane
g.set_profiling(True)
for i in range(a.reps):
    t0 = time.perf_counter(); t = g.fields(r, np.array([0.0, 0.0, 1.0]), "total", jvec=True) if a.jvec else g.jtensors(r); dt = time.perf_counter() - t0
    s = g.stats()
    print(f"rep {i}: {r.shape[0]} pts {dt*1e3:.1f} ms wall, contract {s['ms_contract']:.1f} ms, basis {s['ms_basis']:.2f} ms, "
          f"{s['executed_flops']/max(s['ms_contract'],1e-9)/1e9:.2f} TF executed, mean nact {s['sum_nact']/max(s['n_tiles'],1):.0f}")
