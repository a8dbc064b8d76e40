"""tests/test_gpu_parity.py::test_spherical_basis_vs_oracle[False] outside pytest (same call sequence), for compute-sanitizer:
    compute-sanitizer --tool initcheck python tools/repro_spherical_test.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
import oracle_lib as O
import gimic_b200

rng = np.random.default_rng(23)
coords = np.array([[0.0, 0.0, 0.0], [1.9, 0.4, -0.3], [-0.7, 2.1, 0.8]])
shells = [(0, [3.1, 0.7], [0.4, 0.7]), (1, [1.3], [1.0]), (2, [0.9, 0.35], [0.6, 0.5]), (3, [0.8], [1.0]), (4, [0.7], [1.0]), (5, [0.6], [1.0])]
nat = coords.shape[0]
sh = dict(coords=coords, nctr_per_atom=np.full(nat, len(shells), np.int32), ctr_l=np.array([s[0] for s in shells] * nat, np.int32),
          ctr_npf=np.array([len(s[1]) for s in shells] * nat, np.int32), xp=np.array([x for s in shells for x in s[1]] * nat),
          cc=np.array([x for s in shells for x in s[2]] * nat))
nsph = nat * sum(2 * l + 1 for l, _, _ in shells)
da = fixtures.dens_to_colmajor(fixtures.synthetic_density(nsph, seed=3, general_p=True))
db = fixtures.dens_to_colmajor(fixtures.synthetic_density(nsph, seed=4, general_p=True))
g = gimic_b200.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=da, dens_beta=db, turbomole_order=False, spherical=True, **sh)
o = O.Oracle.from_arrays(dens_a=da, dens_b=db, turbomole_order=False, spherical=True, **sh)
r = rng.uniform(-3, 4, size=(300, 3))
bf, dr = g.basis(r[:10])
for sc in ("alpha", "beta", "total", "spindens"):
    t, ref = g.jtensors(r, sc), o.ctensor(r, sc)
    err = np.abs(t - ref) / (1e-10 * np.abs(ref) + 1e-12)
    rows = np.flatnonzero(err.max(1) > 1.0)
    print(sc, "max scaled err %.3g" % err.max(), "bad rows", rows.size, rows[:10], "nan", int(np.isnan(t).sum()))
