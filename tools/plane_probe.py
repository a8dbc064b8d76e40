"""One 36 x 36 Gauss plane integral through the headline flake, a few times: for an ncu launch list of the latency-bound integral mode
(tools/gpu_r02_o.sh).  Prints the wall time per call and the profiled stage times."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gimic_b200
from gimic_b200 import synthetic

n = int(sys.argv[1]) if len(sys.argv) > 1 else 36
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
sh, dens, nbf, origin, basv, pts = bench.build_workload(278, 16)
g = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=synthetic.dens_to_colmajor(dens), **sh)
B = np.array([0.0, 0.0, 1.0])
plane = bench.flake_plane(sh, n)
g.integrate(plane, B, "total", 3)
ts = []
for _ in range(reps):
    t0 = time.perf_counter(); s = g.integrate(plane, B, "total", 3); ts.append(time.perf_counter() - t0)
g.set_profiling(True); g.integrate(plane, B, "total", 3); st = g.stats(); g.set_profiling(False)
print("plane %dx%d nbf=%d: wall %.1f us/call; tiles %d mean nact %.0f; profiled ms: sort %.3f tiles %.3f basis %.3f contract %.3f; launches %d; sums %s" % (
    n, n, nbf, min(ts) * 1e6, st["n_tiles"], st["sum_nact"] / max(st["n_tiles"], 1), st["ms_sort"], st["ms_tiles"], st["ms_basis"], st["ms_contract"],
    st["launches"], np.array2string(np.asarray(s[:3]), precision=10)))
