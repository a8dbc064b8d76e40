"""ctypes binding to the CPU oracle (oracle/libgimic_oracle.so).

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(gimic_b200/) must never import this module.
"""
import ctypes as C
import os
import subprocess
import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_SO = os.path.join(_ROOT, "oracle", "libgimic_oracle.so")
_lib = None

dp = C.POINTER(C.c_double)
ip = C.POINTER(C.c_int)


def _p(a):
    return None if a is None else a.ctypes.data_as(dp)


def _ip(a):
    return None if a is None else a.ctypes.data_as(ip)


def lib():
    global _lib
    if _lib is None:
        src = os.path.join(_ROOT, "oracle", "gimic_oracle.cpp")
        if (not os.path.exists(_SO)) or os.path.getmtime(_SO) < os.path.getmtime(src):
            subprocess.check_call(["make", "-s", "-C", os.path.join(_ROOT, "oracle")])
        L = C.CDLL(_SO)
        L.go_create_from_files.restype = C.c_void_p
        L.go_create_from_files.argtypes = [C.c_char_p, C.c_char_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_int,
                                           C.c_int, C.c_char_p, C.c_int]
        L.go_create_from_arrays.restype = C.c_void_p
        L.go_create_from_arrays.argtypes = [C.c_int, dp, ip, ip, ip, dp, dp, C.c_int, C.c_int, C.c_double, C.c_int,
                                            C.c_int, C.c_int, dp, dp]
        L.go_destroy.argtypes = [C.c_void_p]
        L.go_next_spherical.argtypes = [C.c_int]
        L.go_c2s.argtypes = [C.c_void_p, C.c_int, dp]
        for f in ("go_nccgto", "go_nbf", "go_natoms", "go_ngto", "go_is_turbomole", "go_nctr", "go_nprim"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        L.go_atom_coords.argtypes = [C.c_void_p, dp]
        L.go_export_shells.argtypes = [C.c_void_p, ip, ip, ip, dp, dp, dp, dp]
        L.go_get_density.argtypes = [C.c_void_p, C.c_int, C.c_int, dp]
        L.go_calc_basis.argtypes = [C.c_void_p, dp, dp, dp, dp, dp]
        L.go_ctensor.restype = C.c_int
        L.go_ctensor.argtypes = [C.c_void_p, C.c_long, dp, C.c_char_p, dp, dp, C.c_int]
        L.go_max_threads.restype = C.c_int
        L.go_acid.restype = C.c_double
        L.go_acid.argtypes = [dp]
        L.go_au2si.restype = C.c_double
        L.go_au2si.argtypes = [C.c_double]
        L.go_jvectors.argtypes = [C.c_long, dp, dp, dp]
        L.go_jmod_signed.argtypes = [C.c_long, dp, dp, dp, dp]
        L.go_acid_field.argtypes = [C.c_long, dp, dp]
        L.go_property.argtypes = [C.c_long, dp, dp, dp, C.c_int, dp, C.c_int, C.POINTER(C.c_long), dp]
        L.go_gauss_points.restype = C.c_int
        L.go_gauss_points.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int, C.c_char_p, dp, dp]
        L.go_grid_file.restype = C.c_void_p
        L.go_grid_file.argtypes = [C.c_long, dp]
        L.go_grid_std.restype = C.c_void_p
        L.go_grid_std.argtypes = [dp, dp, dp, dp, C.c_char_p, C.c_int, C.c_int, ip, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.go_grid_bond.restype = C.c_void_p
        L.go_grid_bond.argtypes = [dp, dp, dp, C.c_double, C.c_int, dp, dp, C.c_int, C.c_double, C.c_int, dp,
                                   C.c_char_p, C.c_int, C.c_int, ip, C.c_int, dp, C.c_int, dp, C.c_int, dp]
        L.go_grid_destroy.argtypes = [C.c_void_p]
        L.go_grid_npts.argtypes = [C.c_void_p, ip]
        L.go_grid_geometry.argtypes = [C.c_void_p, dp, dp, dp, dp, dp, dp]
        L.go_grid_axis.argtypes = [C.c_void_p, C.c_int, dp, dp]
        L.go_grid_points.argtypes = [C.c_void_p, dp]
        L.go_grid_center.argtypes = [C.c_void_p, dp]
        L.go_get_magnet.restype = C.c_int
        L.go_get_magnet.argtypes = [C.c_void_p, C.c_char_p, dp, dp]
        L.go_integrate.restype = C.c_int
        L.go_integrate.argtypes = [C.c_void_p, C.c_void_p, dp, C.c_char_p, C.c_int, dp, C.c_int]
        _lib = L
    return _lib


def _arr(x, dtype=np.float64):
    return np.ascontiguousarray(np.asarray(x, dtype=dtype))


class Grid:
    """grid_t of src/fgimic/grid.f90 as restated by the oracle."""

    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle grid construction failed")
        self.h = handle
        L = lib()
        n = np.zeros(3, dtype=np.int32)
        L.go_grid_npts(self.h, _ip(n))
        self.npts = tuple(int(v) for v in n)
        self.origin = np.zeros(3); self.basv = np.zeros(9); self.lengths = np.zeros(3)
        self.ortho = np.zeros(3); self.center_bond = np.zeros(3)
        rad = np.zeros(1)
        L.go_grid_geometry(self.h, _p(self.origin), _p(self.basv), _p(self.lengths), _p(self.ortho),
                           _p(self.center_bond), _p(rad))
        self.radius = float(rad[0])
        self.basv = self.basv.reshape(3, 3)  # basv[v] = v-th basis vector

    def __del__(self):
        try:
            lib().go_grid_destroy(self.h)
        except Exception:
            pass

    @property
    def n(self):
        return self.npts[0] * self.npts[1] * self.npts[2]

    def points(self):
        r = np.zeros((self.n, 3))
        lib().go_grid_points(self.h, _p(r))
        return r

    def axis(self, d):
        pts = np.zeros(self.npts[d]); wgt = np.zeros(self.npts[d])
        lib().go_grid_axis(self.h, d, _p(pts), _p(wgt))
        return pts, wgt

    def center(self):
        c = np.zeros(3)
        lib().go_grid_center(self.h, _p(c))
        return c

    def magnet(self, axis="", magnet=(0.0, 0.0, 0.0)):
        out = np.zeros(3)
        rc = lib().go_get_magnet(self.h, axis.encode(), _p(_arr(magnet)), _p(out))
        if rc:
            raise RuntimeError(f"get_magnet failed rc={rc}")
        return out


def grid_file(xyz):
    xyz = _arr(xyz)
    return Grid(lib().go_grid_file(xyz.shape[0], _p(xyz)))


def _opt(v, dtype=np.float64):
    return (0, _arr(np.zeros(3), dtype)) if v is None else (1, _arr(v, dtype))


def grid_std(origin, ivec, jvec, lengths, type="even", gauss_order=7, grid_points=None, spacing=None,
             rotation=None, rotation_origin=None):
    hp, gp = _opt(grid_points, np.int32); hs, sp = _opt(spacing)
    hr, rot = _opt(rotation); ho, ro = _opt(rotation_origin)
    return Grid(lib().go_grid_std(_p(_arr(origin)), _p(_arr(ivec)), _p(_arr(jvec)), _p(_arr(lengths)), type.encode(),
                                  gauss_order, hp, _ip(gp), hs, _p(sp), hr, _p(rot), ho, _p(ro)))


def grid_bond(c1, c2, fix, distance, height=None, width=None, up=None, down=None, in_=None, out=None,
              radius=None, magnet=None, type="even", gauss_order=7, grid_points=None, spacing=None,
              rotation=None, rotation_origin=None):
    if height is not None:
        use_hw, hgt, wdt = 1, _arr(height), _arr(width)
    else:
        use_hw, hgt, wdt = 0, _arr([up, down]), _arr([in_, out])
    hp, gp = _opt(grid_points, np.int32); hs, sp = _opt(spacing)
    hr, rot = _opt(rotation); ho, ro = _opt(rotation_origin); hm, mg = _opt(magnet)
    return Grid(lib().go_grid_bond(_p(_arr(c1)), _p(_arr(c2)), _p(_arr(fix)), float(distance), use_hw, _p(hgt), _p(wdt),
                                   0 if radius is None else 1, 0.0 if radius is None else float(radius), hm, _p(mg),
                                   type.encode(), gauss_order, hp, _ip(gp), hs, _p(sp), hr, _p(rot), ho, _p(ro)))


class Oracle:
    def __init__(self, handle):
        if not handle:
            raise RuntimeError("oracle construction failed")
        self.h = handle
        L = lib()
        self.nbf = L.go_nbf(self.h); self.natoms = L.go_natoms(self.h)
        self.nctr = L.go_nctr(self.h); self.nprim = L.go_nprim(self.h)
        self.ngto = L.go_ngto(self.h); self.is_turbomole = bool(L.go_is_turbomole(self.h))

    @classmethod
    def from_files(cls, mol, xdens, uhf=False, screening=True, screening_thrs=1e-8, giao=True, diamag=True,
                   paramag=True, spherical=False):
        err = C.create_string_buffer(512)
        lib().go_next_spherical(int(spherical))
        h = lib().go_create_from_files(str(mol).encode(), str(xdens).encode(), int(uhf), int(screening),
                                       float(screening_thrs), int(giao), int(diamag), int(paramag), err, 512)
        if not h:
            raise RuntimeError("oracle: " + err.value.decode())
        return cls(h)

    @classmethod
    def from_arrays(cls, coords, nctr_per_atom, ctr_l, ctr_npf, xp, cc, dens_a, dens_b=None, turbomole_order=False,
                    screening_thrs=1e-8, giao=True, diamag=True, paramag=True, spherical=False):
        lib().go_next_spherical(int(spherical))
        coords = _arr(coords); nca = _arr(nctr_per_atom, np.int32); cl = _arr(ctr_l, np.int32)
        cn = _arr(ctr_npf, np.int32); xp = _arr(xp); cc = _arr(cc); da = _arr(dens_a)
        db = None if dens_b is None else _arr(dens_b)
        h = lib().go_create_from_arrays(coords.shape[0], _p(coords), _ip(nca), _ip(cl), _ip(cn), _p(xp), _p(cc),
                                        int(turbomole_order), int(dens_b is not None), float(screening_thrs),
                                        int(giao), int(diamag), int(paramag), _p(da), _p(db))
        return cls(h)

    def __del__(self):
        try:
            lib().go_destroy(self.h)
        except Exception:
            pass

    def c2s(self, l):
        """c2s_oper(l)%po of cao2sao.f90 as a (2l+1) x ncart(l) array in this molecule's component order"""
        out = np.zeros((2 * l + 1, (l + 1) * (l + 2) // 2))
        lib().go_c2s(self.h, l, _p(out))
        return out

    def atom_coords(self):
        out = np.zeros((self.natoms, 3))
        lib().go_atom_coords(self.h, _p(out))
        return out

    def export_shells(self):
        nca = np.zeros(self.natoms, np.int32); cl = np.zeros(self.nctr, np.int32); cn = np.zeros(self.nctr, np.int32)
        xp = np.zeros(self.nprim); cc = np.zeros(self.nprim); ncc = np.zeros(self.nprim); thr = np.zeros(self.nctr)
        lib().go_export_shells(self.h, _ip(nca), _ip(cl), _ip(cn), _p(xp), _p(cc), _p(ncc), _p(thr))
        return dict(coords=self.atom_coords(), nctr_per_atom=nca, ctr_l=cl, ctr_npf=cn, xp=xp, cc=cc, ncc=ncc, thrs=thr)

    def density(self, spin=1, which=0):
        """dens.f90 storage after read_dens (UHF-halved, Turbomole-reordered); returned as D[mu, nu]."""
        out = np.zeros(self.nbf * self.nbf)
        lib().go_get_density(self.h, spin, which, _p(out))
        return out.reshape(self.nbf, self.nbf).T.copy()  # column-major (mu fastest) -> [mu, nu]

    def densities(self, spin=1):
        return np.stack([self.density(spin, b) for b in range(4)])

    def calc_basis(self, r):
        n = self.nbf
        bf = np.zeros(n); dr = np.zeros(3 * n); db = np.zeros(3 * n); d2 = np.zeros(9 * n)
        lib().go_calc_basis(self.h, _p(_arr(r)), _p(bf), _p(dr), _p(db), _p(d2))
        return bf, dr.reshape(3, n), db.reshape(3, n), d2.reshape(9, n)

    def ctensor(self, r, spincase="total", nthreads=0, want_edens=False):
        r = _arr(r).reshape(-1, 3)
        n = r.shape[0]
        tens = np.zeros((n, 9)); ed = np.zeros(n) if want_edens else None
        rc = lib().go_ctensor(self.h, n, _p(r), spincase.encode(), _p(tens), _p(ed), nthreads)
        if rc:
            raise RuntimeError(f"oracle ctensor failed rc={rc}")
        return (tens, ed) if want_edens else tens

    def integrate(self, grid, bb, spincase="total", what=0, nthreads=0):
        out = np.zeros(3)
        rc = lib().go_integrate(self.h, grid.h, _p(_arr(bb)), spincase.encode(), what, _p(out), nthreads)
        if rc:
            raise RuntimeError(f"oracle integrate failed rc={rc}")
        return out


def jvectors(tens, b):
    tens = _arr(tens); out = np.zeros((tens.shape[0], 3))
    lib().go_jvectors(tens.shape[0], _p(tens), _p(_arr(b)), _p(out))
    return out


def jmod_signed(r, vec, mag):
    r = _arr(r); vec = _arr(vec); out = np.zeros(r.shape[0])
    lib().go_jmod_signed(r.shape[0], _p(r), _p(vec), _p(_arr(mag)), _p(out))
    return out


def acid_field(tens):
    tens = _arr(tens); out = np.zeros(tens.shape[0])
    lib().go_acid_field(tens.shape[0], _p(tens), _p(out))
    return out


def property(r, w, tens, coords, seg_counts):
    """get_property (jfield.f90:584-929): returns totals[k] = (xx, yy, zz, spos, sneg) and scont[k][s] = cumulative
    ((xx+yy+zz)/3, spos/3, sneg/3) at the end of point block s; k == natoms is the magnetizability"""
    r = _arr(r); w = _arr(w); tens = _arr(tens); coords = _arr(coords)
    seg = np.ascontiguousarray(seg_counts, dtype=np.int64)
    nat, nseg = coords.shape[0], seg.size
    out = np.zeros((nat + 1, 5 + 3 * nseg))
    lib().go_property(r.shape[0], _p(r), _p(w), _p(tens), nat, _p(coords), nseg, seg.ctypes.data_as(C.POINTER(C.c_long)), _p(out))
    return out[:, :5], out[:, 5:].reshape(nat + 1, nseg, 3)


def gauss_points(a, b, npts, order, quadr="gauss"):
    pts = np.zeros(npts); w = np.zeros(npts)
    rc = lib().go_gauss_points(a, b, npts, order, quadr.encode(), _p(pts), _p(w))
    if rc:
        raise RuntimeError(f"gauss_points rc={rc}")
    return pts, w


def au2si(x):
    return lib().go_au2si(float(x))


def max_threads():
    return lib().go_max_threads()
