"""Seeded synthetic workloads: carbon-like centres carrying the def2-TZVP carbon
shell set of the reference's test/c4h4/MOL (5s3p2d1f -> 11 contractions, 20 primitives, 36 cartesian
functions), random symmetric D and antisymmetric (or general) P_b with c4h4-like magnitudes.
Used by bench.py and the tests so the GPU path and the CPU oracle see identical inputs."""
import numpy as np

# def2-TZVP carbon shell set, exponents/coefficients as in the reference's test/c4h4/MOL lines 8-38
C_TZVP = [
    (0, [13575.349682, 2035.233368, 463.22562359, 131.20019598, 42.853015891, 15.584185766],
        [0.0002224581, 0.0017232738, 0.0089255715, 0.0357279845, 0.1107625993, 0.2429562763]),
    (0, [6.2067138508, 2.5764896527], [0.4144026345, 0.2374496866]),
    (0, [0.5769633942], [1.0]),
    (0, [0.2297283136], [1.0]),
    (0, [0.09516444], [1.0]),
    (1, [34.697232244, 7.9582622826, 2.3780826883, 0.8143320818], [0.0053333658, 0.0358641091, 0.1421587333, 0.3427047185]),
    (1, [0.2888754725], [0.4644582243]),
    (1, [0.1005682367], [0.2495578987]),
    (2, [1.097], [1.0]),
    (2, [0.318], [1.0]),
    (3, [0.761], [1.0]),
]
NFUNC_C = sum((l + 1) * (l + 2) // 2 for l, _, _ in C_TZVP)  # 36


def hex_flake(natoms, spacing=2.7, seed=4321, jitter=0.05):
    """compact graphene-like flake: the natoms lattice sites closest to the origin, jittered"""
    a = spacing
    pts = []
    m = int(np.sqrt(natoms)) + 4
    a1 = np.array([np.sqrt(3) * a, 0.0]); a2 = np.array([np.sqrt(3) * a / 2, 1.5 * a])
    for i in range(-m, m + 1):
        for j in range(-m, m + 1):
            o = i * a1 + j * a2
            pts.append(o); pts.append(o + np.array([0.0, a]))
    pts = np.array(pts)
    pts = pts[np.argsort((pts ** 2).sum(1), kind="stable")][:natoms]
    rng = np.random.default_rng(seed)
    xyz = np.zeros((natoms, 3)); xyz[:, :2] = pts
    xyz += rng.uniform(-jitter, jitter, size=xyz.shape)
    return xyz


def ring(natoms, radius=120.0, seed=4321, jitter=0.05):
    ang = 2 * np.pi * np.arange(natoms) / natoms
    xyz = np.stack([radius * np.cos(ang), radius * np.sin(ang), np.zeros(natoms)], 1)
    rng = np.random.default_rng(seed)
    return xyz + rng.uniform(-jitter, jitter, size=xyz.shape)


def synthetic_shells(coords):
    nat = coords.shape[0]
    nctr = np.full(nat, len(C_TZVP), np.int32)
    l = np.array([s[0] for s in C_TZVP] * nat, np.int32)
    npf = np.array([len(s[1]) for s in C_TZVP] * nat, np.int32)
    xp = np.array([x for s in C_TZVP for x in s[1]] * nat)
    cc = np.array([x for s in C_TZVP for x in s[2]] * nat)
    return dict(coords=np.ascontiguousarray(coords), nctr_per_atom=nctr, ctr_l=l, ctr_npf=npf, xp=xp, cc=cc)


def synthetic_density(nbf, seed=1234, general_p=False, dtype=np.float64):
    """D = sym, P_b = antisym (like real data) unless general_p; scaled to c4h4-like magnitudes.
    Returned as dens[4, mu, nu]."""
    rng = np.random.default_rng(seed)
    dens = np.empty((4, nbf, nbf), dtype)
    a = rng.uniform(-1.0, 1.0, size=(nbf, nbf))
    dens[0] = 0.5 * (a + a.T) * 0.8
    dens[0][np.diag_indices(nbf)] = np.abs(np.diag(dens[0])) * 0.2 + 0.02
    for b in range(1, 4):
        a = rng.uniform(-1.0, 1.0, size=(nbf, nbf))
        dens[b] = a * 0.3 if general_p else 0.5 * (a - a.T) * 0.3
    return dens


def synthetic_case(natoms, geometry="flake", seed=1234, general_p=False):
    coords = hex_flake(natoms) if geometry == "flake" else ring(natoms)
    sh = synthetic_shells(coords)
    nbf = natoms * NFUNC_C
    return sh, synthetic_density(nbf, seed, general_p), nbf


def dens_to_colmajor(dens):
    """dens[b, mu, nu] -> the flat column-major (mu fastest) layout of dens.f90 / XDENS, 4 matrices back to back"""
    return np.ascontiguousarray(np.transpose(dens, (0, 2, 1))).reshape(-1)


def box_grid(coords, npts, margin=8.0, zhalf=8.0):
    """even grid over the bounding box of the molecule + margin (z: [-zhalf, zhalf] for flat molecules);
    returns origin, basis vectors, axis points (setup_even_grid, grid.f90:351-373: pts = (i-1)*step)"""
    lo = coords.min(0) - margin
    hi = coords.max(0) + margin
    if hi[2] - lo[2] < 2 * zhalf:
        lo[2], hi[2] = -zhalf, zhalf
    L = hi - lo
    pts = [np.arange(n) * (L[d] / (n - 1)) for d, n in enumerate(npts)]
    return lo, np.eye(3), pts
