"""ctypes/numpy front-end of the CPU oracle (oracle/gs_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs; the product package never imports this module.
Parity status: unpinned for the rasterizer proper (the reference ships no tests and its CUDA
source is absent), pinned for SH / camera / loss pieces -- see gs_oracle.c header.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))


def build(force=False):
    """Compile oracle/libgso_f32.so and libgso_f64.so with the Makefile next to this file."""
    targets = [os.path.join(_HERE, n) for n in ("libgso_f32.so", "libgso_f64.so")]
    src = os.path.join(_HERE, "gs_oracle.c")
    stale = force or any((not os.path.exists(t)) or os.path.getmtime(t) < os.path.getmtime(src) for t in targets)
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return targets


class Oracle:
    """One precision variant of the oracle.  dtype=np.float32 is the parity oracle (same IEEE
    op sequence as the CUDA kernels in the integer-deciding chain); np.float64 is used for
    finite-difference checks of the hand-written backward."""

    def __init__(self, dtype=np.float32, threads=1):
        build()
        self.dtype = np.dtype(dtype)
        name = "libgso_f32.so" if self.dtype == np.float32 else "libgso_f64.so"
        self.lib = C.CDLL(os.path.join(_HERE, name))
        assert self.lib.gso_real_bytes() == self.dtype.itemsize
        self.lib.gso_render_count.restype = C.c_int64
        self.creal = C.c_float if self.dtype == np.float32 else C.c_double
        self.lib.gso_set_threads(int(threads))

    # -- helpers -------------------------------------------------------------------------
    def _r(self, a):
        return np.ascontiguousarray(a, dtype=self.dtype)

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p)

    def set_threads(self, n):
        self.lib.gso_set_threads(int(n))

    def get_block_xy(self):
        a, b, c = C.c_int(), C.c_int(), C.c_int()
        self.lib.gso_get_block_xy(C.byref(a), C.byref(b), C.byref(c))
        return a.value, b.value, c.value

    def eval_sh(self, deg, sh, dirs):
        sh, dirs = self._r(sh), self._r(dirs)
        out = np.zeros((sh.shape[0], 3), self.dtype)
        self.lib.gso_eval_sh(C.c_int(sh.shape[0]), C.c_int(deg), self._p(sh), self._p(dirs), self._p(out))
        return out

    # -- preprocess ----------------------------------------------------------------------
    def preprocess_forward(self, means3D, scales, rotations, shs, opacities, cam, sh_degree=None, scale_modifier=1.0):
        P = means3D.shape[0]
        D = cam["sh_degree"] if sh_degree is None else sh_degree
        m, s, q, sh, o = map(self._r, (means3D, scales, rotations, shs, opacities))
        V, PM, cp = self._r(cam["viewmatrix"]), self._r(cam["projmatrix"]), self._r(cam["campos"])
        out = dict(
            means2D=np.zeros((P, 2), self.dtype), depths=np.zeros((P,), self.dtype), radii=np.zeros((P,), np.int32),
            cov3D=np.zeros((P, 6), self.dtype), conic_opacity=np.zeros((P, 4), self.dtype),
            rgb=np.zeros((P, 3), self.dtype), clamped=np.zeros((P,), np.uint8))
        self.lib.gso_preprocess_forward(
            C.c_int(P), C.c_int(D), self._p(m), self._p(s), self.creal(scale_modifier), self._p(q), self._p(o),
            self._p(sh), self._p(V), self._p(PM), self._p(cp), C.c_int(cam["image_width"]), C.c_int(cam["image_height"]),
            self.creal(cam["tanfovx"]), self.creal(cam["tanfovy"]), self._p(out["means2D"]), self._p(out["depths"]),
            self._p(out["radii"]), self._p(out["cov3D"]), self._p(out["conic_opacity"]), self._p(out["rgb"]),
            self._p(out["clamped"]))
        return out

    def preprocess_backward(self, means3D, scales, rotations, shs, opacities, cam, radii, clamped, dL_dmeans2D,
                            dL_dconic_opacity, dL_drgb, sh_degree=None, scale_modifier=1.0):
        P = means3D.shape[0]
        D = cam["sh_degree"] if sh_degree is None else sh_degree
        m, s, q, sh, o = map(self._r, (means3D, scales, rotations, shs, opacities))
        V, PM, cp = self._r(cam["viewmatrix"]), self._r(cam["projmatrix"]), self._r(cam["campos"])
        g2, gc, gr = self._r(dL_dmeans2D), self._r(dL_dconic_opacity), self._r(dL_drgb)
        radii = np.ascontiguousarray(radii, np.int32)
        clamped = np.ascontiguousarray(clamped, np.uint8)
        out = dict(means3D=np.zeros((P, 3), self.dtype), scales=np.zeros((P, 3), self.dtype),
                   rotations=np.zeros((P, 4), self.dtype), opacities=np.zeros((P, 1), self.dtype),
                   shs=np.zeros((P, 16, 3), self.dtype))
        self.lib.gso_preprocess_backward(
            C.c_int(P), C.c_int(D), self._p(m), self._p(s), self.creal(scale_modifier), self._p(q), self._p(o),
            self._p(sh), self._p(V), self._p(PM), self._p(cp), C.c_int(cam["image_width"]), C.c_int(cam["image_height"]),
            self.creal(cam["tanfovx"]), self.creal(cam["tanfovy"]), self._p(radii), self._p(clamped), self._p(g2),
            self._p(gc), self._p(gr), self._p(out["means3D"]), self._p(out["scales"]), self._p(out["rotations"]),
            self._p(out["opacities"]), self._p(out["shs"]))
        return out

    # -- distribution helpers ------------------------------------------------------------
    def get_local2j_ids_bool(self, H, W, world_size, means2D, radii, strategy):
        m = self._r(means2D)
        radii = np.ascontiguousarray(radii, np.int32)
        strategy = np.ascontiguousarray(strategy, np.int32)
        out = np.zeros((m.shape[0], world_size), np.uint8)
        self.lib.gso_get_local2j_ids_bool(C.c_int(m.shape[0]), C.c_int(H), C.c_int(W), C.c_int(world_size), self._p(m),
                                          self._p(radii), self._p(strategy), self._p(out))
        return out.astype(bool)

    def get_local2j_ids_bool_rects(self, H, W, world_size, means2D, radii, rects):
        m = self._r(means2D)
        radii = np.ascontiguousarray(radii, np.int32)
        rects = np.ascontiguousarray(rects, np.int32)
        out = np.zeros((m.shape[0], world_size), np.uint8)
        self.lib.gso_get_local2j_ids_bool_rects(C.c_int(m.shape[0]), C.c_int(H), C.c_int(W), C.c_int(world_size),
                                                self._p(m), self._p(radii), self._p(rects), self._p(out))
        return out.astype(bool)

    # -- render --------------------------------------------------------------------------
    def render_forward(self, H, W, means2D, conic_opacity, rgb, depths, radii, compute_locally, bg):
        P = means2D.shape[0]
        m, co, col, bgr = self._r(means2D), self._r(conic_opacity), self._r(rgb), self._r(bg)
        d32 = np.ascontiguousarray(depths, np.float32)
        radii = np.ascontiguousarray(radii, np.int32)
        cl = np.ascontiguousarray(compute_locally, np.uint8).reshape(-1)
        T = ((H + 15) // 16) * ((W + 15) // 16)
        assert cl.size == T
        touched = np.zeros((P,), np.uint32)
        offsets = np.zeros((P,), np.uint32)
        R = self.lib.gso_render_count(C.c_int(P), C.c_int(H), C.c_int(W), self._p(m), self._p(radii), self._p(cl),
                                      self._p(touched), self._p(offsets))
        keys = np.zeros((max(R, 1),), np.uint64)
        ids = np.zeros((max(R, 1),), np.uint32)
        ranges = np.zeros((T, 2), np.uint32)
        self.lib.gso_render_bin(C.c_int(P), C.c_int64(R), C.c_int(H), C.c_int(W), self._p(m), self._p(d32),
                                self._p(radii), self._p(cl), self._p(offsets), self._p(keys), self._p(ids),
                                self._p(ranges))
        color = np.zeros((3, H, W), self.dtype)
        final_T = np.zeros((H, W), self.dtype)
        n_contrib = np.zeros((H, W), np.uint32)
        stats = np.zeros((3,), np.int64)
        self.lib.gso_render_blend_forward(C.c_int(H), C.c_int(W), self._p(m), self._p(co), self._p(col), self._p(bgr),
                                          self._p(cl), self._p(ranges), self._p(ids), self._p(color),
                                          self._p(final_T), self._p(n_contrib), self._p(stats))
        return dict(image=color, final_T=final_T, n_contrib=n_contrib, stats=stats, R=int(R), tiles_touched=touched,
                    offsets=offsets, keys=keys[:R], ids=ids[:R], ranges=ranges, compute_locally=cl)

    def render_backward(self, H, W, means2D, conic_opacity, rgb, bg, fwd, dL_dpix):
        P = means2D.shape[0]
        m, co, col, bgr = self._r(means2D), self._r(conic_opacity), self._r(rgb), self._r(bg)
        g = self._r(dL_dpix)
        ids = np.ascontiguousarray(fwd["ids"] if fwd["R"] > 0 else np.zeros(1, np.uint32), np.uint32)
        fT = self._r(fwd["final_T"])
        out = dict(means2D=np.zeros((P, 2), self.dtype), conic_opacity=np.zeros((P, 4), self.dtype),
                   rgb=np.zeros((P, 3), self.dtype))
        self.lib.gso_render_blend_backward(
            C.c_int(P), C.c_int(H), C.c_int(W), self._p(m), self._p(co), self._p(col), self._p(bgr),
            self._p(fwd["compute_locally"]), self._p(fwd["ranges"]), self._p(ids), self._p(fT),
            self._p(fwd["n_contrib"]), self._p(g), self._p(out["means2D"]), self._p(out["conic_opacity"]),
            self._p(out["rgb"]))
        return out

    # -- loss ----------------------------------------------------------------------------
    def loss(self, img, gt, n_pixels_total, lambda_dssim=0.2):
        """img, gt: (3, rows, W) strip. Returns (Ll1, ssim, dloss/dimg) for
        loss = (1-lambda) Ll1 + lambda (1 - ssim)."""
        x, y = self._r(img), self._r(gt)
        rows, W = x.shape[1], x.shape[2]
        l1, ss = self.creal(0), self.creal(0)
        grad = np.zeros_like(x)
        self.lib.gso_loss(C.c_int(rows), C.c_int(W), C.c_double(float(n_pixels_total)), self.creal(lambda_dssim),
                          self._p(x), self._p(y), C.byref(l1), C.byref(ss), self._p(grad))
        return float(l1.value), float(ss.value), grad

    # -- whole training step on one rank ---------------------------------------------------
    def train_step(self, scene, cam, gt_u8, compute_locally=None, lambda_dssim=0.2, bg=(0.0, 0.0, 0.0)):
        """preprocess -> render -> loss -> backward, the call sequence of
        /root/reference/train_internal.py:139-196 for one camera on one rank (W==1)."""
        H, W = cam["image_height"], cam["image_width"]
        T = ((H + 15) // 16) * ((W + 15) // 16)
        cl = np.ones((T,), np.uint8) if compute_locally is None else compute_locally
        pre = self.preprocess_forward(scene["means3D"], scene["scales"], scene["rotations"], scene["shs"],
                                      scene["opacities"], cam)
        fwd = self.render_forward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], pre["depths"], pre["radii"],
                                  cl, bg)
        gt = np.clip(gt_u8.astype(self.dtype) / self.dtype.type(255.0), 0.0, 1.0)
        l1, ss, dimg = self.loss(fwd["image"], gt, H * W, lambda_dssim)
        rb = self.render_backward(H, W, pre["means2D"], pre["conic_opacity"], pre["rgb"], bg, fwd, dimg)
        pb = self.preprocess_backward(scene["means3D"], scene["scales"], scene["rotations"], scene["shs"],
                                      scene["opacities"], cam, pre["radii"], pre["clamped"], rb["means2D"],
                                      rb["conic_opacity"], rb["rgb"])
        loss = (1.0 - lambda_dssim) * l1 + lambda_dssim * (1.0 - ss)
        return dict(loss=loss, Ll1=l1, ssim=ss, pre=pre, fwd=fwd, render_grads=rb, grads=pb, dL_dimage=dimg)
