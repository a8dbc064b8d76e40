"""Latency of the legacy one-point-per-call boundary (gimic_interface.h: gimic_calc_jtensor / gimic_calc_jvector), the way
src/pygimic/field.py:82-93 and tools/PyGimicTest.py.in drive the reference, next to the same points through ONE batched call.
The reference itself needs ~0.65 ms per tensor at nbf=168 on one core (1549 evals/s, test/open-shell/integration stdout)."""
import json, os, sys, tempfile, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
from gimic_b200 import _lib, Gimic

L = _lib.lib()
tmp = tempfile.mkdtemp(); cases = fixtures.materialize(tmp)
mol, xd = cases["c4h4"]["mol"], cases["c4h4"]["xdens"]
pts = np.ascontiguousarray(fixtures.golden_npz("c4h4_readgrid.npz")["grid"][::8][:500])
L.gimic_init(mol.encode(), xd.encode())
b = np.array([0.0, 0.0, 1.0]); L.gimic_set_magnet(b.ctypes.data_as(_lib.dp))
jt = np.zeros(9); jv = np.zeros(3)
for p in pts[:20]:
    L.gimic_calc_jtensor(p.ctypes.data_as(_lib.dp), jt.ctypes.data_as(_lib.dp))
t0 = time.perf_counter()
for p in pts:
    L.gimic_calc_jtensor(p.ctypes.data_as(_lib.dp), jt.ctypes.data_as(_lib.dp))
t1 = time.perf_counter()
for p in pts:
    L.gimic_calc_jvector(p.ctypes.data_as(_lib.dp), jv.ctypes.data_as(_lib.dp))
t2 = time.perf_counter()
L.gimic_finalize()
g = Gimic(mol, xd, screening_thrs=1e-6)
g.jtensors(pts)
t3 = time.perf_counter(); g.jtensors(pts); t4 = time.perf_counter()
print(json.dumps({"nbf": g.nbf, "points": int(pts.shape[0]), "legacy_calc_jtensor_us_per_call": (t1 - t0) / len(pts) * 1e6,
                  "legacy_calc_jvector_us_per_call": (t2 - t1) / len(pts) * 1e6, "batched_call_us_total": (t4 - t3) * 1e6,
                  "reference_cpu_us_per_tensor": 1e6 / 1549.0}))
