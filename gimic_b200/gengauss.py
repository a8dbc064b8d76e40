"""Mirror of src/libgimic/gengauss.pyx:5-12: gausspoints(a, b, order, pts, wgts) fills the caller's arrays."""
import numpy as np

from . import _lib


def gausspoints(a, b, order, pts, wgts, quadrature="gauss"):
    assert pts.dtype == np.float64 and wgts.dtype == np.float64 and pts.size == wgts.size
    q = {"gauss": 0, "lobatto": 1}[quadrature]
    _lib.check(_lib.lib().gimic_b200_gauss_points(float(a), float(b), int(pts.size), int(order), q,
                                                  pts.ctypes.data_as(_lib.dp), wgts.ctypes.data_as(_lib.dp)))
