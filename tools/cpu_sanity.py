"""Sanity link between the CPU oracle and the real Fortran reference: tensor evaluations per second
of the oracle on ONE core for the reference's own small cases, next to the rates the reference's golden stdout files imply
(unknown developer hardware, serial gfortran build).  Runs on the CPU only.

  benzene basis (test/benzene/MOL, nbf=252) with synthetic densities (the XDENS of that test is not in the reference tree),
      30^3 = 27 000 points like test/benzene/3d-keyword-magnet: reference 27000 / 34.28 s = 788 pts/s
  c4h4 (nbf=168, real XDENS), 1296 points like test/c4h4/integration: reference 1217 evals/s
"""
import json, os, sys, tempfile, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import fixtures
import oracle_lib as O

cases = fixtures.materialize(tempfile.mkdtemp())
out = {}
# c4h4: real MOL + XDENS
ob = O.Oracle.from_files(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-8)
r = np.random.default_rng(0).uniform(-5, 5, size=(1296, 3))
for threads in (1, O.max_threads()):
    ob.ctensor(r[:64], "total", nthreads=threads)
    t0 = time.perf_counter(); ob.ctensor(r, "total", nthreads=threads); dt = time.perf_counter() - t0
    out[f"c4h4_nbf168_threads{threads}_pts_per_s"] = r.shape[0] / dt
out["c4h4_reference_stdout_evals_per_s_1core"] = 1217
# benzene: basis from the MOL, synthetic densities
tmp = tempfile.mkdtemp()
xd = os.path.join(tmp, "XDENS")
dens = fixtures.synthetic_density(252, seed=1)
fixtures.write_xdens(xd, fixtures.dens_to_colmajor(dens))
oz = O.Oracle.from_files(cases["benzene_mol"], xd, screening_thrs=1e-8)
assert oz.nbf == 252
g = np.linspace(-8, 8, 30)
r = np.stack(np.meshgrid(g, g, g, indexing="ij"), -1).reshape(-1, 3)
for threads in (1, O.max_threads()):
    oz.ctensor(r[:256], "total", nthreads=threads)
    t0 = time.perf_counter(); oz.ctensor(r, "total", nthreads=threads); dt = time.perf_counter() - t0
    out[f"benzene_nbf252_threads{threads}_pts_per_s"] = r.shape[0] / dt
out["benzene_reference_stdout_pts_per_s_1core"] = 788
out["host"] = {"threads": O.max_threads(), "cpu": next((l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")), "?")}
print(json.dumps(out, indent=1))
