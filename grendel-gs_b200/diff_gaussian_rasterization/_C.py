"""`diff_gaussian_rasterization._C`: the raw helpers the reference calls directly."""
from gs_b200 import _lib
from gs_b200.ops import (get_block_XY, get_local2j_ids_bool, get_local2j_ids_bool_adjust_mode6,  # noqa: F401
                         get_pixels_compute_locally_and_in_rect, get_touched_locally)

_lib.load()  # fail at import time, loudly, if the native library is missing
