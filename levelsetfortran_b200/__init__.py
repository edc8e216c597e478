"""B200-native grid hot path of LevelSetFortran (sign search, WENO5 Gauss-Seidel reinitialisation,
min/max flow) behind a C ABI; this package is the Python mirror of the reference's interface."""
from . import _lib, set_subs, stl, vti  # noqa: F401
from .set_subs import DeviceGrid, ShardedGrid, ReferenceStop, slab_range, minMaxFlow, narrowBand, reinit, signSearch, advectNodes  # noqa: F401

__all__ = ["set_subs", "stl", "vti", "DeviceGrid", "ShardedGrid", "slab_range", "ReferenceStop", "reinit", "narrowBand", "signSearch", "minMaxFlow", "advectNodes"]
