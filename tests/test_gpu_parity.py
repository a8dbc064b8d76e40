"""GPU parity: the CUDA path, called through the C ABI, against the CPU oracle on the same inputs.

Tolerance (BASELINE.json north_star): |gpu - ref| <= 1e-10 |ref| + 1e-12 in FP64.
The oracle itself is pinned to the reference's goldens in test_oracle_golden.py; the goldens that
are runnable are checked here directly as well.
"""
import ctypes as C
import numpy as np
import pytest

import fixtures
import oracle_lib as O

pytestmark = pytest.mark.gpu

RTOL, ATOL = 1e-10, 1e-12


def scaled_err(a, ref):
    return float((np.abs(a - ref) / (RTOL * np.abs(ref) + ATOL)).max()) if a.size else 0.0


def assert_close(a, ref, what=""):
    a, ref = np.asarray(a), np.asarray(ref)
    e = scaled_err(a, ref)
    if e > 1.0 and a.ndim >= 2:         # which rows (points): one tile, a stripe, everything?
        rows = np.flatnonzero((np.abs(a - ref) / (RTOL * np.abs(ref) + ATOL)).reshape(a.shape[0], -1).max(1) > 1.0)
        what += f" [{rows.size} of {a.shape[0]} rows off, first {rows[:12].tolist()}, nan {int(np.isnan(a).sum())}]"
    assert e <= 1.0, f"{what}: max |a-ref| / (1e-10|ref| + 1e-12) = {e:.3g}"


def _jscale_close(j, jref, tref, what):
    """J = T.B compared at the scale of the tensor components that are summed (1e-10 of the point's largest |T B| term)"""
    scale = np.abs(tref).max(axis=1, keepdims=True)
    err = np.abs(j - jref) / (RTOL * np.maximum(np.abs(jref), 1e-3 * scale) + ATOL)
    assert err.max() <= 1.0, f"{what}: {err.max():.3g}"


@pytest.fixture(scope="module")
def gb():
    import gimic_b200
    return gimic_b200


@pytest.fixture(scope="module")
def c4h4(gb, cases):
    m, x = cases["c4h4"]["mol"], cases["c4h4"]["xdens"]
    return gb.Gimic(m, x, screening_thrs=1e-8), O.Oracle.from_files(m, x, screening_thrs=1e-8)


@pytest.fixture(scope="module")
def opensh(gb, cases):
    m, x = cases["open_shell"]["mol"], cases["open_shell"]["xdens"]
    return gb.Gimic(m, x, uhf=True, screening_thrs=1e-8), O.Oracle.from_files(m, x, uhf=True, screening_thrs=1e-8)


def to_product_grid(gb, og):
    pts, wgt = zip(*[og.axis(d) for d in range(3)])
    return gb.Grid(og.origin, og.basv, pts, wgt, radius=og.radius)


# ---- config 1/2: c4h4 (closed shell, Turbomole order) ------------------------------------------------
def test_c4h4_read_grid_tensors_and_golden(c4h4):
    g, o = c4h4
    gold = fixtures.golden_npz("c4h4_readgrid.npz")
    r, B = gold["grid"], gold["magnet"]
    out = g.fields(r, B, "total", tens=True, jvec=True, jmod=True, acid=True, edens=True)
    ref, ed = o.ctensor(r, "total", want_edens=True)
    assert_close(out["tens"], ref, "tensors")
    assert_close(out["edens"], ed, "edens")
    jv = O.jvectors(ref, B)
    assert_close(out["jvec"], jv, "jvec")
    assert_close(out["acid"], O.acid_field(ref), "acid")
    jm = O.jmod_signed(r, jv, B)
    assert_close(np.abs(out["jmod"]), np.abs(jm), "|jmod|")
    assert (np.sign(out["jmod"]) != np.sign(jm)).mean() < 1e-3
    # the reference's own golden (10 printed digits)
    gj = gold["jvec"]
    big = np.abs(gj) > 1e-3 * np.abs(gj).max()
    assert (np.abs(out["jvec"] - gj)[big] / np.abs(gj)[big]).max() < 2e-9


def test_c4h4_integration_golden(gb, c4h4):
    g, o = c4h4
    xyz = o.atom_coords()
    assert np.allclose(g.atom_coords(), xyz)
    og = O.grid_bond(xyz[1], xyz[0], xyz[3], 1.48794, height=[-5.0, 5.0], width=[-1.25614, 6.0], type="gauss",
                     gauss_order=9, grid_points=[30, 30, 0], rotation=[0.0, 0.0, 0.0])
    bb = og.magnet("z")
    grid = to_product_grid(gb, og)
    out = g.integrate(grid, bb, "total", what=7)
    cur = o.integrate(og, bb, "total", 0); mod = o.integrate(og, bb, "total", 1); acid = o.integrate(og, bb, "total", 2)
    assert_close(out[0:3], cur, "current"); assert_close(out[3:6], mod, "modulus"); assert_close(out[6], acid[1], "acid")
    gold = fixtures.golden_json("c4h4_integration.json")["blocks"]
    for blk in gold:
        got = out[3:6] if blk["section"] == "modulus" else out[0:3]
        assert abs(got[0] - blk["au"]) < 1.01e-6 and abs(got[1] - blk["pos"]) < 1.01e-6 and abs(got[2] - blk["neg"]) < 1.01e-6
    # row slabs add up (what a multi-GPU run all-reduces)
    parts = sum(g.integrate(grid, bb, "total", 7, lo, hi) for lo, hi in [(0, 11), (11, 30), (30, 36)])
    assert np.allclose(parts, out, rtol=1e-12, atol=1e-14)


def test_integrate_batch_equals_single_calls(gb, c4h4):
    """current-profile scan: many thin slices of one bond plane in ONE tensor pass == one integrate() per slice == oracle"""
    g, o = c4h4
    xyz = o.atom_coords()
    edges = np.linspace(-1.0, 6.0, 9)
    ogs = [O.grid_bond(xyz[0], xyz[1], xyz[2], 1.3, height=[-3.0, 3.0], width=[edges[i], edges[i + 1]], type="gauss", gauss_order=9,
                       spacing=[0.5, 0.5, 0.5]) for i in range(8)]
    grids = [to_product_grid(gb, og) for og in ogs]
    B = np.array([0.0, 0.0, 1.0])
    batch = g.integrate_batch(grids, B, "total", what=7)
    assert batch.shape == (8, 7)
    for i, (og, pg) in enumerate(zip(ogs, grids)):
        single = g.integrate(pg, B, "total", what=7)
        assert_close(batch[i], single, f"slice {i} batch vs single")
        assert_close(batch[i, 0:3], o.integrate(og, B, "total", what=0), f"slice {i} current vs oracle")
    assert g.integrate_batch([], B).shape == (0, 7)


def test_c4h4_radius_mask(gb, c4h4):
    g, o = c4h4
    xyz = o.atom_coords()
    og = O.grid_bond(xyz[1], xyz[0], xyz[3], 1.48794, height=[-5.0, 5.0], width=[-1.25614, 6.0], type="gauss",
                     gauss_order=9, grid_points=[18, 18, 0], radius=3.0)
    bb = og.magnet("z")
    out = g.integrate(to_product_grid(gb, og), bb, "total", what=7)
    assert_close(out[0:3], o.integrate(og, bb, "total", 0)); assert_close(out[3:6], o.integrate(og, bb, "total", 1))
    assert_close(out[6], o.integrate(og, bb, "total", 2)[1])


@pytest.mark.parametrize("kw", [dict(giao=False), dict(diamag=False), dict(paramag=False), dict(screening=False),
                                dict(screening_thrs=1e-6), dict(giao=False, diamag=False)])
def test_c4h4_option_switches(gb, cases, kw):
    m, x = cases["c4h4"]["mol"], cases["c4h4"]["xdens"]
    okw = dict(screening_thrs=1e-8); okw.update(kw)
    g = gb.Gimic(m, x, **okw)
    o = O.Oracle.from_files(m, x, **okw)
    r = fixtures.golden_npz("c4h4_readgrid.npz")["grid"][::9]
    assert_close(g.jtensors(r), o.ctensor(r), str(kw))
    g.close()


def test_edge_cases(c4h4):
    g, o = c4h4
    assert g.jtensors(np.zeros((0, 3))).shape == (0, 9)                      # empty
    r1 = np.array([[0.3, -0.2, 0.7]])
    assert_close(g.jtensors(r1), o.ctensor(r1), "single point")
    rng = np.random.default_rng(5)
    r = rng.uniform(-6, 6, size=(129, 3))                                    # ragged: one full tile + 1 point
    assert_close(g.jtensors(r), o.ctensor(r), "ragged")
    far = np.array([[200.0, 0, 0], [0, -300.0, 5.0], [1e4, 1e4, 1e4]])      # everything screened -> exact zeros
    t = g.jtensors(far)
    assert (t == 0).all() and (o.ctensor(far) == 0).all()
    mix = np.vstack([far, r[:5], far])                                       # zeros and non-zeros in one tile
    assert_close(g.jtensors(mix), o.ctensor(mix), "mixed")
    dup = np.repeat(r[:3], 50, axis=0)                                       # duplicates / collisions
    td = g.jtensors(dup)
    assert_close(td, o.ctensor(dup), "duplicates")
    on_atom = o.atom_coords()                                                # points on nuclei (0**0 = 1 paths)
    assert_close(g.jtensors(on_atom), o.ctensor(on_atom), "on nuclei")
    plane = r.copy(); plane[:, 2] = 0.0                                      # molecular plane: exact zeros of odd-z functions
    assert_close(g.jtensors(plane), o.ctensor(plane), "nuclear plane")


def test_edge_cases_j_path(gb, c4h4):
    """the same edge cases through the J = T.B kernel (fields with jvec only): empty, single, ragged, screened, on nuclei"""
    import torch
    g, o = c4h4
    B = np.array([0.1, -0.3, 0.95])
    J = lambda x: g.fields(x, B, "total", jvec=True)["jvec"]
    Jo = lambda x: O.jvectors(o.ctensor(x), B)
    assert J(np.zeros((0, 3))).shape == (0, 3)
    r1 = np.array([[0.3, -0.2, 0.7]])
    assert_close(J(r1), Jo(r1), "single point")
    rng = np.random.default_rng(5)
    r = rng.uniform(-6, 6, size=(129, 3))
    far = np.array([[200.0, 0, 0], [0, -300.0, 5.0], [1e4, 1e4, 1e4]])
    assert (J(far) == 0).all()
    mix = np.vstack([far, r[:5], far, r, o.atom_coords()])
    tref = o.ctensor(mix)
    _jscale_close(J(mix), O.jvectors(tref, B), tref, "mixed tile")
    a = J(mix); p = rng.permutation(mix.shape[0])
    assert np.array_equal(J(mix[p]), a[p])                                   # fixed reduction order: bit-reproducible
    f = g.fields(mix, B, "total", jvec=True, jmod=True)
    assert np.array_equal(g.jmod_from_jvec(mix, f["jvec"], B), f["jmod"])


def test_point_order_invariance(c4h4):
    g, _ = c4h4
    rng = np.random.default_rng(11)
    r = rng.uniform(-7, 7, size=(1000, 3))
    t = g.jtensors(r)
    p = rng.permutation(1000)
    t2 = g.jtensors(r[p])
    assert_close(t2, t[p], "permuted points")
    t3 = g.jtensors(r)
    assert (t3 == t).all(), "same input twice must be bit-identical (deterministic reduction order)"


def test_legacy_single_point_api(gb, cases, c4h4):
    _, o = c4h4
    from gimic_b200 import _lib
    L = _lib.lib()
    L.gimic_init(cases["c4h4"]["mol"].encode(), cases["c4h4"]["xdens"].encode())
    o6 = O.Oracle.from_files(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-6)   # gimic_init uses 1e-6
    b = np.array([0.0, 0.0, 1.0]); r = np.array([0.5, 0.25, 1.0]); jt = np.zeros(9); jv = np.zeros(3); mj = C.c_double()
    dp = C.POINTER(C.c_double)
    L.gimic_set_magnet(b.ctypes.data_as(dp))
    L.gimic_set_spin(b"total")
    L.gimic_calc_jtensor(r.ctypes.data_as(dp), jt.ctypes.data_as(dp))
    L.gimic_calc_jvector(r.ctypes.data_as(dp), jv.ctypes.data_as(dp))
    L.gimic_calc_modj(r.ctypes.data_as(dp), C.byref(mj))
    ref = o6.ctensor(r[None])[0]
    assert_close(jt, ref, "legacy jtensor")
    assert_close(jv, O.jvectors(ref[None], b)[0], "legacy jvector")
    assert abs(mj.value - np.linalg.norm(jv)) < 1e-15
    L.gimic_finalize()
    # the Cython-class mirror
    g = gb.Gimic(cases["c4h4"]["mol"], cases["c4h4"]["xdens"])
    g.set_property("magnet", [0.0, 0.0, 1.0])
    assert_close(np.array(g.jvector(r)), O.jvectors(ref[None], b)[0], "Gimic.jvector")
    assert_close(g.jtensor(r), ref, "Gimic.jtensor")
    with pytest.raises(ValueError):
        g.set_spin("gamma")
    g.close()


def test_basis_vectors_vs_oracle(c4h4, opensh):
    """rows A1-A3 of the scope table on their own: Phi, dPhi/dr and exact screening zeros (bfeval.f90, caos.f90, basis.f90:118-136)"""
    for (g, o), seed in ((c4h4, 1), (opensh, 2)):       # Turbomole and standard component order
        rng = np.random.default_rng(seed)
        r = np.vstack([rng.uniform(-7, 7, size=(40, 3)), o.atom_coords()[:2], [[0.3, -0.2, 0.0]], [[30.0, 0, 0]]])
        bf, dr = g.basis(r)
        for i, p in enumerate(r):
            obf, odr, _, _ = o.calc_basis(p)
            assert_close(bf[i], obf, "bf"); assert_close(dr[i], odr, "dr")
            assert ((bf[i] == 0) == (obf == 0)).all(), "screening pattern differs"
        assert (bf[-1] == 0).all() and (dr[-1] == 0).all()


# ---- config 3: open shell ----------------------------------------------------------------------------
def test_open_shell_spin_cases_and_golden(gb, opensh):
    g, o = opensh
    gold = fixtures.golden_npz("open_shell_3d.npz")
    og = O.grid_std([-8.0, -8.0, -8.0], [1.0, 0, 0], [0, 1.0, 0], [16.0, 16.0, 16.0], type="even", spacing=[0.5, 0.5, 0.5])
    bb = og.magnet("X")
    grid = to_product_grid(gb, og)
    assert grid.n == 35937
    r_all = og.points()
    assert np.allclose(grid.points(), r_all, atol=1e-14)
    idx = gold["index"]
    for tag, sc in (("", "total"), ("alpha", "alpha"), ("beta", "beta"), ("spindens", "spindens")):
        tens = g.jtensors_grid(grid, 0, grid.n, sc)                      # device-generated grid points, full 33^3
        ref = o.ctensor(r_all[idx], sc)
        assert_close(tens[idx], ref, f"open-shell {sc}")
        f = g.fields_from_tensors(r_all, tens, bb, jvec=True, jmod=True)
        gj = gold["jvec" + tag]
        tol = 1e-5 * np.abs(gj) + 1e-9 * np.abs(gj).max()
        assert (np.abs(f["jvec"][idx] - gj) <= tol).all(), tag
        gm = gold["jmod" + tag]
        assert (np.abs(np.abs(f["jmod"][idx]) - np.abs(gm)) <= 1e-5 * np.abs(gm) + 1e-9 * np.abs(gm).max()).all(), tag


def test_open_shell_integration_golden(gb, opensh):
    g, o = opensh
    xyz = o.atom_coords()
    og = O.grid_bond(xyz[0], xyz[1], xyz[3], 1.32, height=[-5.0, 5.0], width=[-2.2, 5.0], type="gauss", gauss_order=9,
                     grid_points=[30, 30, 0])
    bb = og.magnet("X")
    grid = to_product_grid(gb, og)
    gold = fixtures.golden_json("open_shell_integration.json")["blocks"]
    res = {sc: g.integrate(grid, bb, sc, what=3) for sc in ("total", "alpha", "beta", "spindens")}
    for blk in gold:
        got = res[blk["spin"]][3:6] if blk["section"] == "modulus" else res[blk["spin"]][0:3]
        assert abs(got[0] - blk["au"]) < 1.01e-6 and abs(got[1] - blk["pos"]) < 1.01e-6 and abs(got[2] - blk["neg"]) < 1.01e-6, blk
    assert_close(res["alpha"][0:3], o.integrate(og, bb, "alpha", 0), "alpha current vs oracle")
    assert_close(res["spindens"][3:6], o.integrate(og, bb, "spindens", 1), "spindens modulus vs oracle")


def test_closed_shell_rejects_beta(gb, c4h4):
    g, _ = c4h4
    with pytest.raises(gb.GimicB200Error) as e:
        g.jtensors(np.zeros((1, 3)), "beta")
    assert e.value.code == -4


# ---- config 4/5 shape: synthetic carbon flakes --------------------------------------------------------
@pytest.mark.parametrize("natoms,npts,general_p", [(6, 700, False), (14, 500, True), (42, 300, False)])
def test_synthetic_flake_vs_oracle(gb, natoms, npts, general_p):
    sh, dens, nbf = fixtures.synthetic_case(natoms, "flake", general_p=general_p)
    flat = fixtures.dens_to_colmajor(dens)
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, **sh)
    o = O.Oracle.from_arrays(dens_a=flat, **sh)
    assert g.nbf == nbf == o.nbf
    rng = np.random.default_rng(3)
    lo, hi = sh["coords"].min(0) - 6.0, sh["coords"].max(0) + 6.0
    r = rng.uniform(lo, hi, size=(npts, 3)); r[:, 2] = rng.uniform(-4, 4, size=npts)
    ref, ed = o.ctensor(r, want_edens=True)
    out = g.fields(r, [0, 0, 1.0], tens=True, edens=True)
    assert_close(out["tens"], ref, f"flake nbf={nbf}")
    assert_close(out["edens"], ed, "edens")
    st = g.stats()
    assert st["n_points"] == npts and st["executed_flops"] > 0 and st["executed_flops"] <= 1.2 * st["dense_flops"] * 128
    g.close()


def test_synthetic_ring_far_origin(gb):
    """atoms ~120 bohr from the origin: the gauge-difference operand must not lose digits"""
    sh, dens, nbf = fixtures.synthetic_case(10, "ring")
    flat = fixtures.dens_to_colmajor(dens)
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, **sh)
    o = O.Oracle.from_arrays(dens_a=flat, **sh)
    rng = np.random.default_rng(4)
    r = sh["coords"][rng.integers(0, 10, 400)] + rng.uniform(-5, 5, size=(400, 3))
    assert_close(g.jtensors(r), o.ctensor(r), "ring")
    g.close()


def test_linearity_in_density(gb):
    """size-independent property: T is linear in (D, P): T[a X + b Y] = a T[X] + b T[Y]"""
    sh, d1, nbf = fixtures.synthetic_case(20, "flake", seed=1)
    d2 = fixtures.synthetic_density(nbf, seed=2, general_p=True)
    rng = np.random.default_rng(8)
    r = rng.uniform(-12, 12, size=(4000, 3)); r[:, 2] *= 0.3
    ts = []
    for d in (d1, d2, 0.5 * d1 - 2.0 * d2):
        g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fixtures.dens_to_colmajor(d), **sh)
        ts.append(g.jtensors(r)); g.close()
    comb = 0.5 * ts[0] - 2.0 * ts[1]
    scale = np.abs(ts[0]).max() + np.abs(ts[1]).max()
    assert np.abs(ts[2] - comb).max() < 1e-12 * scale


def test_uhf_total_is_alpha_plus_beta(gb):
    sh, da, nbf = fixtures.synthetic_case(8, "flake", seed=5)
    db = fixtures.synthetic_density(nbf, seed=6)
    fa, fb = fixtures.dens_to_colmajor(da), fixtures.dens_to_colmajor(db)
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fa, dens_beta=fb, **sh)
    o = O.Oracle.from_arrays(dens_a=fa, dens_b=fb, **sh)
    rng = np.random.default_rng(9)
    r = rng.uniform(-8, 8, size=(600, 3))
    for sc in ("alpha", "beta", "total", "spindens"):
        assert_close(g.jtensors(r, sc), o.ctensor(r, sc), sc)
    g.close()


def test_divj_small_for_giao_and_device_pointers(gb, c4h4):
    """no reference semantics for divj at this commit (parity unpinned): self-test that the central-difference
    divergence of J = T.B matches an independent finite-difference of the oracle's J; and the torch zero-copy path."""
    import torch
    g, o = c4h4
    rng = np.random.default_rng(21)
    r = rng.uniform(-3, 3, size=(64, 3))
    B = np.array([0.0, 0.0, 1.0]); h = 1e-3
    out = g.fields(r, B, divj=True, jvec=True, divj_h=h)
    div = np.zeros(64)
    for ax in range(3):
        e = np.zeros(3); e[ax] = h
        div += (O.jvectors(o.ctensor(r + e), B)[:, ax] - O.jvectors(o.ctensor(r - e), B)[:, ax]) / (2 * h)
    assert np.abs(out["divj"] - div).max() < 1e-9
    rt = torch.from_numpy(r).cuda()
    tt = g.jtensors(rt)
    assert tt.is_cuda and tt.shape == (64, 9)
    assert_close(tt.cpu().numpy(), o.ctensor(r), "device-pointer path")


@pytest.mark.parametrize("turbomole", [False, True])
def test_high_angular_momentum_shells(gb, turbomole):
    """s..h shells (l = 0..5, MAX_L of globals.f90:30) in the standard and the Turbomole component order (gtodefs.f90:86-123)"""
    rng = np.random.default_rng(17)
    coords = np.array([[0.0, 0.0, 0.0], [1.9, 0.4, -0.3], [-0.7, 2.1, 0.8]])
    shells = [(0, [3.1, 0.7], [0.4, 0.7]), (1, [1.3], [1.0]), (2, [0.9, 0.35], [0.6, 0.5]), (3, [0.8], [1.0]), (4, [0.7], [1.0]), (5, [0.6], [1.0])]
    nat = coords.shape[0]
    sh = dict(coords=coords, nctr_per_atom=np.full(nat, len(shells), np.int32), ctr_l=np.array([s[0] for s in shells] * nat, np.int32),
              ctr_npf=np.array([len(s[1]) for s in shells] * nat, np.int32), xp=np.array([x for s in shells for x in s[1]] * nat),
              cc=np.array([x for s in shells for x in s[2]] * nat))
    nbf = nat * sum((l + 1) * (l + 2) // 2 for l, _, _ in shells)
    assert nbf == 3 * 56
    flat = fixtures.dens_to_colmajor(fixtures.synthetic_density(nbf, seed=3, general_p=True))
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, turbomole_order=turbomole, **sh)
    o = O.Oracle.from_arrays(dens_a=flat, turbomole_order=turbomole, **sh)
    r = rng.uniform(-3, 4, size=(300, 3))
    bf, dr = g.basis(r[:20])
    for i in range(20):
        obf, odr, _, _ = o.calc_basis(r[i])
        assert_close(bf[i], obf, "bf l<=5"); assert_close(dr[i], odr, "dr l<=5")
    assert_close(g.jtensors(r), o.ctensor(r), f"l<=5 turbomole={turbomole}")
    g.close()


@pytest.mark.parametrize("turbomole", [False, True])
def test_spherical_basis_vs_oracle(gb, turbomole):
    """Advanced.spherical=on (cao2sao.f90): densities over 2l+1 components per shell.  The product folds the projection into
    the densities once; the oracle projects the basis vectors at every point like the reference.  s..h shells."""
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
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=da, dens_beta=db, turbomole_order=turbomole, spherical=True, **sh)
    o = O.Oracle.from_arrays(dens_a=da, dens_b=db, turbomole_order=turbomole, spherical=True, **sh)
    assert g.nbf == o.nbf == nsph
    r = rng.uniform(-3, 4, size=(300, 3))
    bf, dr = g.basis(r[:10])
    for i in range(10):
        obf, odr, _, _ = o.calc_basis(r[i])
        assert_close(bf[i], obf, "spherical bf"); assert_close(dr[i], odr, "spherical dr")
    for sc in ("alpha", "beta", "total", "spindens"):
        assert_close(g.jtensors(r, sc), o.ctensor(r, sc), f"spherical {sc} turbomole={turbomole}")
    g.close()


def test_spherical_mol_xdens_files(gb, cases, tmp_path):
    """spherical=on through the file path: a Turbomole-ordered MOL (c4h4, s p d f shells) with an XDENS over the spherical
    components; checks the SAO-space Turbomole permutation (reorder.f90:54-96 on 2l+1 counts) and the XDENS reader sizes."""
    o_cart = O.Oracle.from_files(cases["c4h4"]["mol"], cases["c4h4"]["xdens"], screening_thrs=1e-8)
    sh = o_cart.export_shells()
    nsph = int(sum(2 * l + 1 for l in sh["ctr_l"]))
    rng = np.random.default_rng(9)
    xd = tmp_path / "XDENS_sph"
    np.savetxt(xd, rng.uniform(-0.3, 0.3, size=4 * nsph * nsph), fmt="%.12e")
    g = gb.Gimic(cases["c4h4"]["mol"], str(xd), screening_thrs=1e-8, spherical=True)
    o = O.Oracle.from_files(cases["c4h4"]["mol"], str(xd), screening_thrs=1e-8, spherical=True)
    assert g.nbf == o.nbf == nsph
    r = rng.uniform(-4, 4, size=(200, 3))
    assert_close(g.jtensors(r), o.ctensor(r), "spherical MOL/XDENS")
    g.close()


def test_binary_xdens_cache_gives_identical_tensors(gb, cases, tmp_path, opensh):
    """the binary XDENS cache (gimic_b200_convert_xdens) is read back bit for bit: UHF halving and Turbomole reorder included"""
    g_text, _ = opensh
    gb.convert_xdens(cases["open_shell"]["xdens"], g_text.nbf, tmp_path / "XDENS.bin", uhf=True)
    g_bin = gb.Gimic(cases["open_shell"]["mol"], tmp_path / "XDENS.bin", uhf=True, screening_thrs=1e-8)
    r = np.random.default_rng(2).uniform(-4, 4, size=(257, 3))
    for sc in ("alpha", "spindens"):
        assert np.array_equal(g_bin.jtensors(r, sc), g_text.jtensors(r, sc))
    g_bin.close()
    with pytest.raises(gb.GimicB200Error, match="binary XDENS cache"):
        gb.Gimic(cases["benzene_mol"], tmp_path / "XDENS.bin", screening_thrs=1e-8)   # nbf 252 vs a cache for 168


def test_more_active_atoms_than_the_shared_memory_table(gb):
    """k_jtensor stages the GIAO tap weights of up to 1024 active atoms per tile in shared memory and reads them from global
    memory beyond that: 1100 one-function atoms with screening off (every atom active in every tile), plus a p shell on a few"""
    rng = np.random.default_rng(31)
    nat = 1100
    coords = rng.uniform(-6, 6, size=(nat, 3))
    nctr = np.ones(nat, np.int32); nctr[:5] = 2
    ls, npf, xp, cc = [], [], [], []
    for a in range(nat):
        ls.append(0); npf.append(1); xp.append(0.4 + 0.001 * a); cc.append(1.0)
        if a < 5:
            ls.append(1); npf.append(1); xp.append(0.7); cc.append(1.0)
    sh = dict(coords=coords, nctr_per_atom=nctr, ctr_l=np.array(ls, np.int32), ctr_npf=np.array(npf, np.int32), xp=np.array(xp), cc=np.array(cc))
    nbf = nat + 15
    flat = fixtures.dens_to_colmajor(fixtures.synthetic_density(nbf, seed=8, general_p=True) * 0.05)
    g = gb.Gimic.from_arrays(dens_alpha=flat, screening=False, **sh)
    o = O.Oracle.from_arrays(dens_a=flat, screening_thrs=-1.0, **sh)
    r = rng.uniform(-5, 5, size=(150, 3))
    assert_close(g.jtensors(r), o.ctensor(r), "1100 active atoms per tile")
    assert g.stats()["n_tiles"] >= 1
    g.close()


@pytest.mark.parametrize("seed", [101, 102, 103, 104, 105, 106])
def test_random_molecules_vs_oracle(gb, seed):
    """randomised inputs: 2..40 atoms in a box of random size, every atom with its own random shell set (l = 0..4, 1..4
    primitives, exponents from tight to diffuse so the screening radii differ widely), random component order, open or closed
    shell, random option switches; points inside and far outside the molecule (tiles with no, few and all atoms active)"""
    rng = np.random.default_rng(seed)
    nat = int(rng.integers(2, 41))
    box = rng.uniform(2.0, 25.0)
    coords = rng.uniform(-box, box, size=(nat, 3)) + rng.uniform(-50, 50, size=3) * (seed % 2)   # odd seeds: far from the origin
    nctr, ls, npf, xp, cc = [], [], [], [], []
    for a in range(nat):
        k = int(rng.integers(1, 6)); nctr.append(k)
        for _ in range(k):
            l = int(rng.integers(0, 5)); n = int(rng.integers(1, 5))
            ls.append(l); npf.append(n)
            xp += list(10.0 ** rng.uniform(-1.3, 2.0, size=n)); cc += list(rng.uniform(0.1, 1.0, size=n))
    sh = dict(coords=coords, nctr_per_atom=np.array(nctr, np.int32), ctr_l=np.array(ls, np.int32), ctr_npf=np.array(npf, np.int32),
              xp=np.array(xp), cc=np.array(cc))
    nbf = int(sum((l + 1) * (l + 2) // 2 for l in ls))
    uhf = bool(rng.integers(0, 2)); tm = bool(rng.integers(0, 2))
    kw = dict(giao=bool(rng.integers(0, 4)), diamag=bool(rng.integers(0, 4)), paramag=bool(rng.integers(0, 4)))
    thr = float(10.0 ** rng.uniform(-10, -5))
    da = fixtures.dens_to_colmajor(fixtures.synthetic_density(nbf, seed=seed, general_p=bool(rng.integers(0, 2))))
    db = fixtures.dens_to_colmajor(fixtures.synthetic_density(nbf, seed=seed + 1000, general_p=True)) if uhf else None
    g = gb.Gimic.from_arrays(dens_alpha=da, dens_beta=db, turbomole_order=tm, screening_thrs=thr, **kw, **sh)
    o = O.Oracle.from_arrays(dens_a=da, dens_b=db, turbomole_order=tm, screening_thrs=thr, **kw, **sh)
    ctr = coords.mean(0)
    r = np.vstack([ctr + rng.uniform(-box - 3, box + 3, size=(400, 3)), ctr + rng.uniform(-box - 40, box + 40, size=(150, 3)),
                   coords[rng.integers(0, nat, size=50)] + rng.normal(scale=0.05, size=(50, 3))])
    for sc in (("alpha", "beta", "total", "spindens") if uhf else ("total",)):
        res = g.fields(r, np.array([0.3, -0.2, 0.9]), sc, tens=True, edens=True)
        to, eo = o.ctensor(r, sc, want_edens=True)
        what = f"seed {seed} {sc} nat={nat} nbf={nbf} uhf={uhf} tm={tm} {kw} thr={thr:.1e}"
        # Random (unphysical) densities make single tensor components cancel to 1e-4 of the point's scale, and total/spindens
        # are sums of alpha and beta parts of either sign: the 1e-10 bound is taken at the scale of the quantities that are
        # summed (the largest component of the point, alpha and beta separately), where FP64 noise limits the oracle as well.
        scale = np.abs(to).max(axis=1, keepdims=True)
        bound = RTOL * np.maximum(np.abs(to), 1e-3 * scale) + ATOL
        if uhf and sc in ("total", "spindens"):
            # T(alpha) +- T(beta): each part carries the 1e-10 bound at its own scale and the two add (one contraction with the
            # operand D_alpha +- D_beta sums terms of the size of both), so the bound of the combination is the sum of the bounds
            ta, tb = o.ctensor(r, "alpha"), o.ctensor(r, "beta")
            sa, sb = np.abs(ta).max(axis=1, keepdims=True), np.abs(tb).max(axis=1, keepdims=True)
            bound = RTOL * (np.maximum(np.abs(ta), 1e-3 * sa) + np.maximum(np.abs(tb), 1e-3 * sb)) + ATOL
        err = np.abs(res["tens"] - to) / bound
        assert err.max() <= 1.0, f"{what}: {err.max():.3g}"
        if not uhf:
            assert_close(res["edens"], eo, what + " edens")
        Bf = np.array([0.3, -0.2, 0.9])
        # J = sum_b T(m,b) B_b cancels once more between the b terms: floor at 1e-2 of the point's tensor scale (1e-12 relative)
        jbound = RTOL * np.maximum(np.abs(O.jvectors(to, Bf)), 1e-2 * scale) + ATOL
        if uhf and sc in ("total", "spindens"):      # the sum of the alpha and beta bounds, as for the tensor above
            jbound = RTOL * (np.maximum(np.abs(O.jvectors(ta, Bf)), 1e-2 * sa) + np.maximum(np.abs(O.jvectors(tb, Bf)), 1e-2 * sb)) + ATOL
        jerr = np.abs(g.fields(r, Bf, sc, jvec=True)["jvec"] - O.jvectors(to, Bf)) / jbound
        assert jerr.max() <= 1.0, f"{what} J path: {jerr.max():.3g}"
    g.close()


def test_jvec_only_path_vs_oracle(gb, cases, c4h4, opensh):
    """fields(jvec/jmod without tens/acid) takes the J = T.B kernel (operands (D, sum_b B_b P_b), one tap weight per row):
    same J, signed |J| and rho as contracting the oracle's tensor with B afterwards (compute_jvectors, jfield.f90:167-184)"""
    rng = np.random.default_rng(12)
    r = np.vstack([rng.uniform(-5, 5, size=(700, 3)), rng.uniform(-30, 30, size=(100, 3))])
    for B in (np.array([0.0, 0.0, 1.0]), np.array([0.3, -0.5, 0.8]), np.array([-1.0, 0.0, 0.0])):
        g, o = c4h4
        tref, eref = o.ctensor(r, "total", want_edens=True)
        f = g.fields(r, B, "total", jvec=True, jmod=True, edens=True)
        jref = O.jvectors(tref, B)
        _jscale_close(f["jvec"], jref, tref, f"c4h4 J path B={B}")
        assert_close(f["edens"], eref, "J path edens")
        jm = O.jmod_signed(r, jref, B)
        big = np.abs(jm) > 1e-9 * np.abs(jm).max()                      # the sign of a vanishing |J| is noise
        assert np.allclose(f["jmod"][big], jm[big], rtol=1e-9, atol=1e-12)
        only = g.fields(r, B, "total", jmod=True)                        # |J| alone (J to scratch)
        assert np.array_equal(only["jmod"], f["jmod"]) and set(only) == {"jmod"}
        full = g.fields(r, B, "total", tens=True, jvec=True)             # tensor path for comparison
        _jscale_close(f["jvec"], full["jvec"], tref, "J path vs tensor path")
    g, o = opensh
    B = np.array([0.2, 0.1, -0.97])
    for sc in ("alpha", "beta", "total", "spindens"):
        tref = o.ctensor(r, sc)
        scale_t = np.maximum(np.abs(o.ctensor(r, "alpha")), np.abs(o.ctensor(r, "beta")))
        _jscale_close(g.fields(r, B, sc, jvec=True)["jvec"], O.jvectors(tref, B), np.maximum(np.abs(tref), scale_t), f"open shell {sc}")
    for kw in (dict(giao=False), dict(diamag=False), dict(paramag=False), dict(screening=False)):
        m, x = cases["c4h4"]["mol"], cases["c4h4"]["xdens"]
        g2 = gb.Gimic(m, x, screening_thrs=1e-8, **kw)
        o2 = O.Oracle.from_files(m, x, screening_thrs=1e-8, **kw)
        tref = o2.ctensor(r[:300], "total")
        _jscale_close(g2.fields(r[:300], B, "total", jvec=True)["jvec"], O.jvectors(tref, B), np.maximum(np.abs(tref), 1e-6), f"J path {kw}")
        g2.close()


def test_jvec_only_path_far_from_origin_and_large(gb):
    """J path on the synthetic ring (atoms 120 bohr from the origin: the per-row tap weights are built from position
    differences) and on a 42-centre flake (tiles with hundreds of active functions)"""
    rng = np.random.default_rng(3)
    B = np.array([0.1, 0.2, 0.97])
    for geometry, natoms, npts in (("ring", 30, 400), ("flake", 42, 300)):
        sh, dens, nbf = fixtures.synthetic_case(natoms, geometry, seed=77, general_p=True)
        flat = fixtures.dens_to_colmajor(dens)
        g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, **sh); o = O.Oracle.from_arrays(dens_a=flat, **sh)
        ctr = sh["coords"][rng.integers(0, natoms, size=npts)]
        r = ctr + rng.normal(scale=2.5, size=(npts, 3))
        tref = o.ctensor(r)
        _jscale_close(g.fields(r, B, "total", jvec=True)["jvec"], O.jvectors(tref, B), tref, f"{geometry} J path")
        g.close()


def test_general_contraction_mol_file(gb, tmp_path):
    """INTGRL blocks with ncf > 1 (general contractions are split into segmented ones, intgrl.f90:172-216) and
    primitive lines that wrap over several records (list-directed reads)"""
    mol = tmp_path / "MOL"
    mol.write_text("""INTGRL        1    0    1    0    0    0    0    0    0
CFOUR
              hand-written general contraction test
2    0            0.10E-08              0    0
9999.00      3.00
8.0    1 2  1  1
O 1      0.000000000000      0.000000000000      0.200000000000
     4   3
    130.7093200000    0.1543289700    0.0000000000    0.0100000000
     23.8088610000    0.5353281400    0.0000000000    0.0200000000
      6.4436083000    0.4446345400   -0.0999672300
    0.3000000000
      1.1695961000    0.0000000000    0.3995128300    0.7001154700
     2   2
      5.0331513000    0.1559162700    0.2000000000
      1.1695961000    0.6076837200    0.9000000000
1.0    1 1  1
H 1      0.000000000000      1.400000000000     -0.900000000000
     2   1
      3.42525091D+00  0.15432897
      0.62391373      0.53532814
""")
    nbf = 3 + 2 * 3 + 1            # O: 3 s + 2 p contractions; H: 1 s
    xd = tmp_path / "XDENS"
    dens = fixtures.synthetic_density(nbf, seed=11)
    fixtures.write_xdens(str(xd), fixtures.dens_to_colmajor(dens))
    g = gb.Gimic(str(mol), str(xd), screening_thrs=1e-8)
    o = O.Oracle.from_files(str(mol), str(xd), screening_thrs=1e-8)
    assert g.nbf == o.nbf == nbf
    rng = np.random.default_rng(2)
    r = rng.uniform(-2, 2, size=(200, 3))
    assert_close(g.jtensors(r), o.ctensor(r), "general contraction")
    g.close()


def test_property_quadrature_vs_oracle(c4h4):
    """get_property (row A16 / N3): shieldings at all nuclei + magnetizability on a weighted point set with per-atom
    point blocks.  The reference's golden (benzene/magnetizability) has no runnable inputs, so this pins GPU == oracle."""
    g, o = c4h4
    xyz = o.atom_coords()
    rng = np.random.default_rng(31)
    counts = rng.integers(300, 900, size=xyz.shape[0])
    r = np.vstack([xyz[a] + rng.normal(scale=1.5, size=(c, 3)) for a, c in enumerate(counts)])
    w = rng.uniform(0.0, 0.05, size=r.shape[0])
    tens = g.jtensors(r)
    assert_close(tens, o.ctensor(r), "tensors on the numgrid-like point set")
    got = g.property(r, w, tens, xyz, counts)
    tot, scont = O.property(r, w, tens, xyz, counts)
    nat = xyz.shape[0]
    tol = lambda a, b: np.abs(a - b).max() <= 1e-10 * np.abs(b).max() + 1e-12
    assert tol(got["sigma"], tot[:nat, 0:3]) and tol(got["sigma_pos"], tot[:nat, 3] / 3) and tol(got["sigma_neg"], tot[:nat, 4] / 3)
    assert tol(got["chi"], tot[nat, 0:3]) and tol(got["chi_pos"], tot[nat, 3] / 3) and tol(got["chi_neg"], tot[nat, 4] / 3)
    ref_contrib = np.diff(scont, axis=1, prepend=0.0)
    assert tol(got["sigma_atoms"], ref_contrib[:nat]) and tol(got["chi_atoms"], ref_contrib[nat])
    assert np.allclose(got["sigma_atoms"].sum(1)[:, 0], got["sigma_iso"], rtol=1e-12)


def test_full_size_nbf10008_properties(gb):
    """BASELINE.json's full basis size (278 centres, nbf = 10 008): oracle parity on a sample the CPU finishes in seconds,
    and size-independent properties on 20 000 points: linearity (UHF total = alpha + beta, spindens = alpha - beta through
    three different operand sets), independence of point order / tiling, grid-slab consistency."""
    sh, da, nbf = fixtures.synthetic_case(278, "flake", seed=1234)
    assert nbf == 10008
    db = fixtures.synthetic_density(nbf, seed=99, general_p=True)
    fa, fb = fixtures.dens_to_colmajor(da), fixtures.dens_to_colmajor(db)
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fa, dens_beta=fb, **sh)
    rng = np.random.default_rng(12)
    lo, hi = sh["coords"].min(0) - 6.0, sh["coords"].max(0) + 6.0
    r = rng.uniform(lo, hi, size=(20000, 3)); r[:, 2] = rng.uniform(-5, 5, size=20000)
    ta, tb, tt, ts = (g.jtensors(r, sc) for sc in ("alpha", "beta", "total", "spindens"))
    scale = np.abs(ta).max() + np.abs(tb).max()
    assert np.abs(tt - (ta + tb)).max() < 1e-12 * scale
    assert np.abs(ts - (ta - tb)).max() < 1e-12 * scale
    p = rng.permutation(20000)
    assert_close(g.jtensors(r[p], "alpha"), ta[p], "point order")
    # regular grid: device-generated slab == explicit points
    from gimic_b200 import synthetic
    origin, basv, pts = synthetic.box_grid(sh["coords"], (32, 32, 8))
    grid = gb.Grid(origin, basv, pts)
    tg = g.jtensors_grid(grid, 1000, 5000, "alpha")
    assert_close(tg, g.jtensors(grid.points()[1000:5000], "alpha"), "grid slab")
    # oracle (dense 7 GEMV per point, 5.6 GB streamed per point): 48 points
    o = O.Oracle.from_arrays(dens_a=fa, dens_b=fb, **sh)
    idx = rng.choice(20000, 48, replace=False)
    assert_close(ta[idx], o.ctensor(r[idx], "alpha"), "alpha vs oracle at nbf=10008")
    assert_close(tt[idx[:16]], o.ctensor(r[idx[:16]], "total"), "total vs oracle at nbf=10008")
    g.close()


# ---- round 2: hot-path basis panels, device-side tile plan, cost-balanced partition, fused field outputs -----------------------
def test_hot_path_basis_panels_vs_oracle(gb, c4h4, opensh):
    """k_basis ITSELF (the panels k_jtensor's TMA copies and epilogue loads consume), not the diagnostic k_basis_dense:
    Phi, dPhi/dr and the exact screening zeros per point, through sort -> tiles -> k_basis -> scatter (bfeval.f90:81-122,295-338)"""
    for (g, o), seed in ((c4h4, 11), (opensh, 12)):       # Turbomole and standard component order
        rng = np.random.default_rng(seed)
        # a dense cluster (many points per tile), scattered points (many small active sets), atoms, a far point
        r = np.vstack([rng.uniform(-1.5, 1.5, size=(300, 3)), rng.uniform(-9, 9, size=(200, 3)), o.atom_coords()[:3], [[40.0, 0, 0]]])
        bf, dr = g.basis_tiles(r)
        assert g.last_tile_info[0] >= 4
        bfd, drd = g.basis(r)                                 # the dense diagnostic kernel: must agree bit for bit where both are non-zero
        for i, p in enumerate(r):
            obf, odr, _, _ = o.calc_basis(p)
            assert_close(bf[i], obf, "bf (hot path)"); assert_close(dr[i], odr, "dr (hot path)")
            assert ((bf[i] == 0) == (obf == 0)).all(), "screening pattern differs"
        # (two kernels, two FMA contractions of the same expressions: equal to rounding, and the same exact zeros)
        assert np.allclose(bf, bfd, rtol=1e-11, atol=1e-300) and np.allclose(dr, drd, rtol=1e-11, atol=1e-300)
        assert ((bf == 0) == (bfd == 0)).all() and ((dr == 0) == (drd == 0)).all()
        assert (bf[-1] == 0).all() and (dr[-1] == 0).all()


def test_hot_path_basis_panels_synthetic_f_shells(gb):
    sh, dens, nbf = fixtures.synthetic_case(14, "flake")
    flat = fixtures.dens_to_colmajor(dens)
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, **sh)
    o = O.Oracle.from_arrays(dens_a=flat, **sh)
    rng = np.random.default_rng(21)
    r = rng.uniform(-9, 9, size=(400, 3)); r[:, 2] *= 0.4
    bf, dr = g.basis_tiles(r)
    for i in range(0, 400, 7):
        obf, odr, _, _ = o.calc_basis(r[i])
        assert_close(bf[i], obf, "bf"); assert_close(dr[i], odr, "dr")
        assert ((bf[i] == 0) == (obf == 0)).all()
    g.close()


def _flake_grid(gb, sh, n):
    origin, basv, pts = fixtures.box_grid(sh["coords"], (n, n, n), margin=6.0, zhalf=6.0)
    return gb.Grid(origin, basv, pts)


def test_partition_union_is_bitwise_the_single_rank_result(gb):
    """every rank tiles the whole grid identically, so the union of the ranks' rows is the one-rank result bit for bit,
    each point owned exactly once; the shares are balanced by cost, not by count (parallel.F90:66-84 splits by count)"""
    sh, dens, nbf = fixtures.synthetic_case(30, "flake")
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fixtures.dens_to_colmajor(dens), **sh)
    grid = _flake_grid(gb, sh, 40)
    n = grid.n
    full = g.jtensors_grid(grid, 0, n)
    B = np.array([0.0, 0.0, 1.0])
    for nranks in (1, 3, 8):
        seen = np.zeros(n, dtype=np.int64)
        got = np.full((n, 9), np.nan)
        costs, counts, flops = [], [], []
        for rk in range(nranks):
            cnt = g.partition(grid, rk, nranks)
            info = g.partition_info()
            assert info["points"] == n and info["owned_points"] == cnt
            res = g.partition_calc(B, "total", tens=True)
            seen[res["index"]] += 1
            got[res["index"]] = res["tens"]
            costs.append(info["cost_owned"]); counts.append(cnt); flops.append(g.stats()["executed_flops"])
            total_cost = info["cost_total"]
        assert (seen == 1).all(), f"nranks={nranks}: points owned {seen.min()}..{seen.max()} times"
        assert np.array_equal(got, full), f"nranks={nranks}: union differs from the single-rank tensors"
        assert sum(costs) == total_cost
        if nranks > 1:
            mean = total_cost / nranks
            # each share is within one tile of the ideal; ~60-170 tiles per rank here (the 256^3 bench grid has 16 000 per rank)
            assert max(costs) <= 1.08 * mean and min(costs) >= 0.92 * mean, costs
            assert max(flops) <= 1.12 * np.mean(flops), flops            # the stats of the launched work, too
            assert max(counts) > 1.15 * min(counts), counts              # equal cost is NOT equal count on a planar molecule
    g.close()


def test_partition_points_fields_and_plan_reuse(gb, c4h4, opensh):
    g, o = c4h4
    rng = np.random.default_rng(31)
    r = rng.uniform(-6, 6, size=(5000, 3))
    B = np.array([0.3, -0.2, 0.9]); B /= np.linalg.norm(B)
    ref = g.fields(r, B, "total", tens=True, jvec=True, jmod=True, acid=True, edens=True)
    parts = {k: np.full_like(v, np.nan) for k, v in ref.items()}
    for rk in range(2):
        g.partition(r, rk, 2)
        a = g.partition_calc(B, "total", tens=True, jvec=True, jmod=True, acid=True, edens=True)
        b = g.partition_calc(B, "total", jvec=True, jmod=True)          # the plan is reused; J path this time
        for k in parts:
            parts[k][a["index"]] = a[k]
        jt = O.jvectors(a["tens"], B)
        assert_close(a["jvec"], jt, "fused jvec vs its own tensor")
        _jscale_close(b["jvec"], a["jvec"], a["tens"], "J path vs tensor path (partitioned)")
        assert np.array_equal(a["index"], b["index"])
    for k in parts:
        assert np.array_equal(parts[k], ref[k]), k
    # open shell: one plan, four spin cases
    gu, ou = opensh
    gu.partition(r[:1500], 0, 1)
    for sc in ("alpha", "beta", "total", "spindens"):
        res = gu.partition_calc(None, sc, tens=True)
        assert_close(res["tens"], ou.ctensor(r[:1500][res["index"]], sc), f"partitioned open-shell {sc}")


def test_partition_on_device_tensors(gb, c4h4):
    import torch
    g, o = c4h4
    rng = np.random.default_rng(32)
    r = rng.uniform(-5, 5, size=(3000, 3))
    rd = torch.from_numpy(r).cuda()
    cnt = g.partition(rd, 1, 3)
    res = g.partition_calc(None, "total", tens=True, device=rd.device)
    assert res["tens"].is_cuda and res["index"].shape[0] == cnt
    idx = res["index"].cpu().numpy()
    assert_close(res["tens"].cpu().numpy(), o.ctensor(r[idx]), "device-resident partition")


def test_thin_point_sets_are_gap_split_on_the_device(gb, c4h4):
    """planar and two-cluster point sets: a Hilbert run that leaves and re-enters the cloud is cut at its gaps (k_tile_split);
    results = oracle, and the active sets stay small (no tile drags in both ends)"""
    g, o = c4h4
    rng = np.random.default_rng(33)
    plane = np.zeros((4000, 3)); plane[:, 0] = rng.uniform(-8, 8, 4000); plane[:, 1] = rng.uniform(-8, 8, 4000); plane[:, 2] = 0.37
    two = np.vstack([rng.normal(size=(700, 3)) * 0.3 + [0, 0, 0], rng.normal(size=(700, 3)) * 0.3 + [25.0, 0, 0]])
    for name, r in (("plane", plane), ("two clusters", two), ("line", np.c_[np.linspace(-30, 30, 999), np.zeros(999), np.zeros(999)])):
        t = g.jtensors(r)
        assert_close(t, o.ctensor(r), name)
        st = g.stats()
        assert st["n_tiles"] >= -(-r.shape[0] // 128)
    # the far cluster sees no basis function: with splitting no tile may carry the near cluster's active set over there
    g.jtensors(two)
    assert g.stats()["sum_nact"] / g.stats()["n_tiles"] <= g.nbf + 8


def test_j_path_plain_tolerance_on_real_densities(c4h4, opensh, cases):
    """J = T.B formed inside the contraction (operands (D, P.B)) at the plain 1e-10 / 1e-12 bound on the reference's own densities"""
    gold = fixtures.golden_npz("c4h4_readgrid.npz")
    g, o = c4h4
    r, B = gold["grid"], gold["magnet"]
    f = g.fields(r, B, "total", jvec=True, jmod=True, edens=True)
    ref, ed = o.ctensor(r, "total", want_edens=True)
    jv = O.jvectors(ref, B)
    assert_close(f["jvec"], jv, "c4h4 J path"); assert_close(f["edens"], ed, "edens (J path)")
    assert_close(np.abs(f["jmod"]), np.abs(O.jmod_signed(r, jv, B)), "|jmod| (J path)")
    gu, ou = opensh
    rng = np.random.default_rng(34)
    ru = rng.uniform(-6, 6, size=(1500, 3))
    Bu = np.array([1.0, 0.0, 0.0])
    for sc in ("alpha", "beta", "total", "spindens"):
        fu = gu.fields(ru, Bu, sc, jvec=True)
        assert_close(fu["jvec"], O.jvectors(ou.ctensor(ru, sc), Bu), f"open-shell J path {sc}")


def test_chunked_host_path_equals_one_call(gb, c4h4, monkeypatch):
    """host buffers larger than one chunk are processed chunk by chunk with the copies overlapped: same numbers as device-resident"""
    import torch
    g, o = c4h4
    rng = np.random.default_rng(35)
    n = (1 << 21) + 70001                      # two chunks
    r = np.empty((n, 3)); r[:] = rng.uniform(-7, 7, size=(n, 3))
    th = g.jtensors(r)
    pick = rng.choice(n, 400, replace=False)
    assert_close(th[pick], o.ctensor(r[pick]), "chunked host path")
    # both chunks filled, no row left untouched
    td = g.jtensors(torch.from_numpy(r).cuda()).cpu().numpy()
    assert np.abs(th - td).max() <= 1e-10 * np.abs(td).max()
    assert not np.isnan(th).any()


def test_nccl_integral_all_reduce_when_two_gpus(gb, cases, tmp_path):
    """integral mode over torch.distributed/NCCL: plane rows split over 2 ranks, one all-reduce of the 7 partial sums
    (integral.f90:157-161 has its collect_sum calls commented out) == the single-GPU sums == the c4h4 golden"""
    import os, subprocess, sys, torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (bench.py --gpus N records the same reduction in its stages.integral_nccl object)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29571", os.path.join(root, "tools", "dist_integral_check.py")], capture_output=True, text=True, timeout=600)
    assert p.returncode == 0 and "OK" in p.stdout, p.stdout[-2000:] + p.stderr[-2000:]


def test_column_sliced_tiles_equal_whole_tiles(gb, monkeypatch):
    """few tiles (a plane, a handful of points): every tile is cut into column slices that run as separate work items and are added by
    k_slice_reduce.  Same numbers as whole tiles (to rounding: the nu sums are split) and as the oracle; both paths (tensor, J)."""
    sh, dens, nbf = fixtures.synthetic_case(14, "flake")
    flat = fixtures.dens_to_colmajor(dens)
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=flat, **sh)
    o = O.Oracle.from_arrays(dens_a=flat, **sh)
    rng = np.random.default_rng(41)
    B = np.array([0.0, 0.6, 0.8])
    for n in (1, 100, 1296, 5000):                 # 1 .. ~45 tiles: 16, 16, ~13 and ~3 slices per tile on 148 SMs
        r = rng.uniform(-7, 7, size=(n, 3)); r[:, 2] *= 0.4
        monkeypatch.setenv("GIMIC_B200_SLICES", "1")
        a = g.fields(r, B, tens=True, jvec=True, jmod=True, acid=True, edens=True)
        ja = g.fields(r, B, jvec=True, jmod=True)
        la = g.stats()["launches"]
        monkeypatch.setenv("GIMIC_B200_SLICES", "0")
        b = g.fields(r, B, tens=True, jvec=True, jmod=True, acid=True, edens=True)
        jb = g.fields(r, B, jvec=True, jmod=True)
        lb = g.stats()["launches"]
        assert la == lb + 2, "the sliced path was not taken"
        ref, ed = o.ctensor(r, want_edens=True)
        assert_close(a["tens"], ref, f"sliced tensors, n={n}"); assert_close(b["tens"], ref, f"whole tiles, n={n}")
        assert_close(a["edens"], ed, "edens")
        _jscale_close(a["jvec"], b["jvec"], ref, "sliced vs whole jvec"); _jscale_close(ja["jvec"], jb["jvec"], ref, "J path sliced vs whole")
        _jscale_close(ja["jvec"], O.jvectors(ref, B), ref, "J path sliced vs oracle")
        assert_close(a["acid"], O.acid_field(ref), "acid (sliced)")
    g.close()


def test_drain_groups_copy_out_the_right_rows(gb, monkeypatch):
    """A partition call with HOST outputs on a range of one panel batch launches the contraction per drain group (a run of
    Hilbert-ordered tiles = of compact output rows) and copies a finished group's rows out while the next group runs.  Same rows,
    bit for bit, as the one-launch path with device outputs and as the ungrouped host path; two ranks too (a rank's first tile is
    not tile 0).  64^3 points at nbf = 540: 2048+ tiles, 4 groups."""
    import torch
    sh, dens, nbf = fixtures.synthetic_case(15, "flake", seed=77)
    g = gb.Gimic.from_arrays(screening_thrs=1e-8, dens_alpha=fixtures.dens_to_colmajor(dens), **sh)
    lo, hi = sh["coords"].min(0) - 6.0, sh["coords"].max(0) + 6.0
    n = 64
    grid = gb.Grid(lo, np.eye(3), [np.linspace(0.0, hi[d] - lo[d], n) for d in range(3)])
    B = np.array([0.2, -0.1, 0.95])
    monkeypatch.setenv("GIMIC_B200_DRAIN_MIN_TILES", "256")
    for rank, world in ((0, 1), (1, 2)):
        monkeypatch.setenv("GIMIC_B200_DRAIN_BATCHES", "1")
        cnt = g.partition(grid, rank, world)
        host = g.partition_calc(B, "total", tens=True, jvec=True, jmod=True)
        launches_grouped = g.stats()["contract_launches"]
        dev = g.partition_calc(B, "total", tens=True, jvec=True, jmod=True, device=torch.device("cuda", 0))
        assert launches_grouped >= 2 and g.stats()["contract_launches"] == 1
        monkeypatch.setenv("GIMIC_B200_DRAIN_BATCHES", "0")
        g.partition(grid, rank, world)
        plain = g.partition_calc(B, "total", tens=True, jvec=True, jmod=True)
        assert g.stats()["contract_launches"] == 1
        for k in ("index", "tens", "jvec", "jmod"):
            d = dev[k].cpu().numpy() if hasattr(dev[k], "cpu") else dev[k]
            assert host[k].shape[0] == cnt and np.array_equal(host[k], d), (rank, k)
            assert np.array_equal(host[k], plain[k]), (rank, k, "grouped vs ungrouped host path")
    g.close()
