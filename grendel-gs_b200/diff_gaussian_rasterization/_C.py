"""`diff_gaussian_rasterization._C`: the raw helpers the reference calls directly."""
from gs_b200 import _lib
from gs_b200.ops import get_block_XY, get_local2j_ids_bool, get_local2j_ids_bool_adjust_mode6  # noqa: F401

_lib.load()  # fail at import time, loudly, if the native library is missing


def _legacy(name):
    def fn(*a, **k):
        raise NotImplementedError(f"_C.{name}: legacy tile-mask helper of the reference's dead code path "
                                  "(SURVEY.md section 8a rows L1-L4); not provided yet.")
    fn.__name__ = name
    return fn


get_touched_locally = _legacy("get_touched_locally")
get_pixels_compute_locally_and_in_rect = _legacy("get_pixels_compute_locally_and_in_rect")
