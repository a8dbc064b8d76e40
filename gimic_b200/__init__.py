"""gimic-b200: B200-native (sm_100a) implementation of GIMIC's grid hot path behind a C ABI."""
from ._lib import GimicB200Error, SPINCASES, SO_PATH  # noqa: F401
from .gimic import Gimic, GimicConnector, NotAvailable, Grid, integrate_distributed, slab, c2s_rows, convert_xdens  # noqa: F401
from .gengauss import gausspoints  # noqa: F401

__all__ = ["Gimic", "GimicConnector", "NotAvailable", "Grid", "gausspoints", "integrate_distributed", "slab", "c2s_rows", "convert_xdens", "GimicB200Error", "SPINCASES"]
