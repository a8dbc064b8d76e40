"""Config 4 of BASELINE.json: coronene-size synthetic (42 centres, nbf=1512), ACID + jmod (+ jvec, edens) on a 128^3 grid,
1 GPU.  Reports the tensor pass and the HBM-bound field pass (k_fields) separately; field-pass roofline =
algorithmic bytes (72 T + 24 r + 24 jvec + 8 jmod + 8 acid per point) / CUDA-event time vs MEASURED_PEAKS hbm_gbs."""
import json, os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import gimic_b200
from gimic_b200 import synthetic, _lib
import ctypes as C

natoms, n1 = 42, (int(sys.argv[1]) if len(sys.argv) > 1 else 128)
sh, dens, nbf = synthetic.synthetic_case(natoms, "flake", seed=1234)
g = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=synthetic.dens_to_colmajor(dens), **sh)
origin, basv, pts = synthetic.box_grid(sh["coords"], (n1, n1, n1))
grid = gimic_b200.Grid(origin, basv, pts)
n = grid.n
dev = torch.device("cuda", 0)
tens = torch.empty((n, 9), dtype=torch.float64, device=dev)
g.set_profiling(True)
for _ in range(2):
    g.jtensors_grid(grid, 0, n, "total", out=tens)
st = g.stats()
t_tens = st["ms_sort"] + st["ms_tiles"] + st["ms_basis"] + st["ms_contract"]
r = torch.from_numpy(grid.points()).to(dev)
jvec = torch.empty((n, 3), dtype=torch.float64, device=dev); jmod = torch.empty(n, dtype=torch.float64, device=dev); acid = torch.empty(n, dtype=torch.float64, device=dev)
B = np.array([0.0, 0.0, 1.0])
L = _lib.lib()
ms = []
for it in range(8):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    _lib.check(L.gimic_b200_fields_from_tensors(g._h, n, C.c_void_p(r.data_ptr()), C.c_void_p(tens.data_ptr()), B.ctypes.data_as(_lib.dp),
                                                C.c_void_p(jvec.data_ptr()), C.c_void_p(jmod.data_ptr()), C.c_void_p(acid.data_ptr()), _lib.DEVICE_PTR))
    ms.append((time.perf_counter() - t0) * 1e3)
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
bytes_pt = 72 + 24 + 24 + 8 + 8
best = min(ms[2:])
out = {"workload": f"synthetic flake nbf={nbf}, {n1}^3 grid, ACID+jmod+jvec", "points": n, "tensor_pass_ms": t_tens,
       "tensor_points_per_s": n / (t_tens * 1e-3), "contract_tflops_executed": st["executed_flops"] / (st["ms_contract"] * 1e-3) / 1e12,
       "mean_nact": st["sum_nact"] / st["n_tiles"], "fields_pass_ms_wall_incl_launch_sync": best, "fields_bytes_per_point": bytes_pt,
       "fields_achieved_gbs": n * bytes_pt / (best * 1e-3) / 1e9, "hbm_peak_gbs": peaks["hbm_gbs"],
       "fields_frac_of_measured_hbm": n * bytes_pt / (best * 1e-3) / 1e9 / peaks["hbm_gbs"]}
print(json.dumps(out))
