"""Import shim: the reference imports six gsplat names unconditionally
(/root/reference/gaussian_renderer/__init__.py:18-25).  The alternate `--backend gsplat` is out of scope."""


def _absent(name):
    def fn(*a, **k):
        raise NotImplementedError(f"gsplat.{name}: the gsplat backend is not part of this drop-in; use --backend default")
    fn.__name__ = name
    return fn


rasterization = _absent("rasterization")
fully_fused_projection = _absent("fully_fused_projection")
spherical_harmonics = _absent("spherical_harmonics")
isect_tiles = _absent("isect_tiles")
isect_offset_encode = _absent("isect_offset_encode")
rasterize_to_pixels = _absent("rasterize_to_pixels")
