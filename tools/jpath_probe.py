"""J = T.B path on one octant of the headline workload (2.1 M points, nbf = 10008): time, and an ncu target for k_jtensor_e<GIAO,JVEC>:
    ncu --set full --import-source on --clock-control none -k regex:k_jtensor_e -s 1 -c 1 -o gpurun_out/jpath python tools/jpath_probe.py"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
import gimic_b200
from gimic_b200 import synthetic

sh, dens, nbf, origin, basv, pts = bench.build_workload(278, 256)
g = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=synthetic.dens_to_colmajor(dens), **sh)
r = bench.slab_points(origin, basv, pts, 0)
B = np.array([0.0, 0.0, 1.0])
g.set_profiling(True)
for k in range(3):
    t0 = time.perf_counter(); out = g.fields(r, B, "total", jvec=True, jmod=True); t1 = time.perf_counter()
    st = g.stats()
    print(f"J path, {r.shape[0]} points: wall {1e3 * (t1 - t0):.2f} ms; contract {st['ms_contract']:.2f} ms basis {st['ms_basis']:.2f} plan {st['ms_sort'] + st['ms_tiles']:.2f}; "
          f"executed {st['executed_flops'] / 1e12:.3f} TF -> {st['executed_flops'] / st['ms_contract'] / 1e9:.2f} TFLOP/s")
