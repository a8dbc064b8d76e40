"""Run-to-run reproducibility of small point sets (fewer tiles than SMs: column slices, k_basis atom groups): the same call repeated must
return the same bits (fixed reduction order).  Used to hunt a flaky parity failure (round 2, call S); also a compute-sanitizer target:
    compute-sanitizer --tool racecheck python tools/repro_small_sets.py 3
"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import fixtures
import gimic_b200

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 200
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
r = rng.uniform(-3, 4, size=(300, 3))
bad = 0
for sc in ("alpha", "beta", "total", "spindens"):
    ref = g.jtensors(r, sc)
    for k in range(reps):
        t = g.jtensors(r, sc)
        if not np.array_equal(t, ref):
            bad += 1
            d = np.abs(t - ref)
            if bad <= 5:
                rows = np.flatnonzero(d.max(1) > 0)
                print(f"{sc} rep {k}: {rows.size} rows differ (first {rows[:8]}), max rel {float((d / (np.abs(ref) + 1e-300)).max()):.3g}")
B = np.array([0.3, -0.2, 0.9])
ref = g.fields(r, B, "total", jvec=True, jmod=True)["jvec"]
for k in range(reps):
    if not np.array_equal(g.fields(r, B, "total", jvec=True, jmod=True)["jvec"], ref):
        bad += 1
print("slices", os.environ.get("GIMIC_B200_SLICES", "default"), "reps", reps, "non-reproducible calls", bad)
