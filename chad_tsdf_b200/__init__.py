"""B200-native implementation of chad::TSDFMap::insert / Submap::finalize."""
from .tsdf_map import TSDFMap  # noqa: F401
from .capi import ChadError  # noqa: F401
