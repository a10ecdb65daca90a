"""Import-only stand-in for the `plyfile` package (PLY I/O of /root/reference/scene/gaussian_model.py:18 and
scene/dataset_readers.py:32).  It is NOT on the default path: add `<repo>/grendel-gs_b200/shims` to PYTHONPATH only in
environments where the real package is missing and PLY files are never read or written (tests, synthetic benchmarks)."""


class _Missing:
    def __init__(self, *a, **k):
        raise NotImplementedError("plyfile is not installed: PLY import/export is unavailable in this environment")

    @classmethod
    def read(cls, *a, **k):
        raise NotImplementedError("plyfile is not installed")

    @classmethod
    def describe(cls, *a, **k):
        raise NotImplementedError("plyfile is not installed")


class PlyData(_Missing):
    pass


class PlyElement(_Missing):
    pass
