"""CPU (numpy) statement of the two algebraic identities k_jtensor relies on, independent of any CUDA code:

  (1) the London/GIAO products  Y_d[p,nu] = sum_mu Phi[p,mu] D[mu,nu] (R_nu - R_mu)_d  obtained WITHOUT extra GEMM planes from
      the running D accumulator at atom boundaries of the K loop (Abel summation), with every atom's
      slot run padded to whole 4-slot K steps exactly like the kernel;
  (2) the J = T.B path: contracting with B before the GEMM (operand sum_b B_b P_b, one tap weight (B x r).(R_A - R_next) per row)
      gives the same J as forming the tensor contribution and contracting afterwards.
"""
import numpy as np


def _setup(seed, natoms=7):
    rng = np.random.default_rng(seed)
    R_atom = rng.uniform(-6, 6, size=(natoms, 3)) + np.array([40.0, -25.0, 10.0])     # far from the origin on purpose
    nfun = rng.integers(1, 11, size=natoms)                                            # active functions per atom (a prefix)
    atom_of = np.repeat(np.arange(natoms), nfun)
    n = atom_of.size
    npts = 16
    r = R_atom.mean(0) + rng.uniform(-2, 2, size=(npts, 3))
    Phi = rng.normal(size=(npts, n))
    D = rng.normal(size=(n, n)); D = 0.5 * (D + D.T)
    P = rng.normal(size=(3, n, n))
    return rng, R_atom, nfun, atom_of, r, Phi, D, P


def _taps(Phi, D, R_atom, nfun, atom_of, centre, weights_for_last):
    """K loop in 4-slot steps over atom runs padded to multiples of 4; returns X0 and the list of (C_A, atom) snapshots"""
    npts, n = Phi.shape
    C = np.zeros((npts, n))
    snaps = []
    f0 = 0
    for a, k in enumerate(nfun):
        kpad = (k + 3) // 4 * 4
        A = np.zeros((npts, kpad)); A[:, :k] = Phi[:, f0:f0 + k]                     # pad slots: Phi = 0
        B = np.zeros((kpad, n)); B[:k] = D[f0:f0 + k]                                  # pad slots gather a valid row (value irrelevant)
        B[k:] = D[f0]
        for s in range(0, kpad, 4):
            C = C + A[:, s:s + 4] @ B[s:s + 4]
        snaps.append((C.copy(), a))
        f0 += k
    return C, snaps


def test_giao_products_from_atom_boundary_taps():
    for seed in range(5):
        rng, R_atom, nfun, atom_of, r, Phi, D, P = _setup(seed)
        Rf = R_atom[atom_of]                                                           # centre of every function
        direct = np.stack([(Phi[:, :, None] * D[None] * (Rf[None, None, :, d] - Rf[None, :, None, d])).sum(1) for d in range(3)], -1)
        c = r.mean(0)                                                                  # "tile centre"
        X0, snaps = _taps(Phi, D, R_atom, nfun, atom_of, c, None)
        Z = np.zeros(X0.shape + (3,))
        for i, (CA, a) in enumerate(snaps):
            delta = R_atom[a] - (R_atom[snaps[i + 1][1]] if i + 1 < len(snaps) else c)  # R_A - R_next, last: R_A - c
            Z += CA[:, :, None] * delta[None, None, :]
        Y = (Rf[None, :, :] - c[None, None, :]) * X0[:, :, None] - Z
        assert np.allclose(X0, Phi @ D, rtol=1e-13, atol=1e-13)
        assert np.abs(Y - direct).max() <= 1e-12 * np.abs(direct).max()


def test_j_path_equals_tensor_then_contract():
    for seed in range(5):
        rng, R_atom, nfun, atom_of, r, Phi, D, P = _setup(seed + 10)
        npts, n = Phi.shape
        dPhi = rng.normal(size=(3, npts, n))                                           # e_m
        Bf = rng.normal(size=3)
        Rf = R_atom[atom_of]
        Y = np.stack([(Phi[:, :, None] * D[None] * (Rf[None, None, :, d] - Rf[None, :, None, d])).sum(1) for d in range(3)], -1)
        X = np.stack([Phi @ P[b] for b in range(3)], -1)                               # X_{1+b}
        rxY = np.cross(r[:, None, :], Y)                                               # (r x Y)_b
        Tp = np.einsum("pnb,mpn->pmb", X + rxY, dPhi)                                  # Tp(m,b)
        J_ref = Tp @ Bf
        # J path: operand sum_b B_b P_b, and w.Y with w = B x r; Y through per-row tap weights
        XB = Phi @ np.einsum("b,bmn->mn", Bf, P)
        w = np.cross(Bf[None, :], r)
        c = r.mean(0)
        X0, snaps = _taps(Phi, D, R_atom, nfun, atom_of, c, None)
        S = np.zeros_like(X0)
        for i, (CA, a) in enumerate(snaps):
            delta = R_atom[a] - (R_atom[snaps[i + 1][1]] if i + 1 < len(snaps) else c)
            S += (w @ delta)[:, None] * CA                                             # one weight per row
        z = XB + ((Rf - c) @ w.T).T * X0 - S
        J = np.einsum("pn,mpn->pm", z, dPhi)
        assert np.abs(J - J_ref).max() <= 1e-12 * np.abs(J_ref).max()
