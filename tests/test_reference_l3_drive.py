"""The REFERENCE's own step functions executed over the drop-in boundary (authoring container only: needs
/root/reference, which does not exist on the GPU box).

`distributed_preprocess3dgs_and_all2all_final` -> `render_final` -> `batched_loss_computation` -> `loss.backward()` ->
`finish_strategy_final` (/root/reference/gaussian_renderer/__init__.py:878-1037, 1217-1288; loss_distribution.py:2536-2637;
workload_division.py:944-998; train_internal.py:139-196) -- and then the LEGACY sequence `replicated_preprocess3dgs` ->
`render` with a flat `DivisionStrategy` and its `extended_compute_locally` mask (:66-174, :458-507) -- run UNMODIFIED on CPU tensors with W = 1, a real reference
`GaussianModel` holding the parameters and a real `DivisionStrategyFinal`.  There is no GPU here, so the two operator
methods of OUR `diff_gaussian_rasterization.GaussianRasterizer` are replaced by recorders that (a) assert every argument
the reference passes (keyword names, dtypes, shapes, the 12 settings fields, the cuda_args keys) against the contract of
SURVEY.md 8b and (b) answer with the CPU oracle wrapped in an autograd function -- so the loss and the parameter
gradients the REFERENCE code computes through that boundary must equal the oracle's own training step.  That pins the
conventions that only the caller knows: transposed matrices, SH layout (N,16,3), pre-applied activations, bool (TY,TX)
mask, extended_compute_locally = None, the stats_collector keys finish_strategy_final reads.
"""
import os
import subprocess
import sys

import pytest

REF = "/root/reference"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "gaussian_renderer")), reason="reference not present")

CODE = r'''
import math, sys, types
from argparse import Namespace
import numpy as np
import torch
sys.path[:0] = [%(pkg)r, %(root)r, %(shims)r, %(ref)r]

# ---- no GPU in this container: "cuda" placement requests of the reference land on the CPU ------------------------------
def _cpuify(fn):
    def w(*a, **k):
        if str(k.get("device", "")).startswith("cuda"):
            k["device"] = "cpu"
        return fn(*a, **k)
    return w
for name in ("zeros", "ones", "empty", "tensor", "arange", "full"):
    setattr(torch, name, _cpuify(getattr(torch, name)))
torch.cuda.synchronize = lambda *a, **k: None
torch.Tensor.cuda = lambda self, *a, **k: self

import diff_gaussian_rasterization as dgr
import utils.general_utils as utils
import gaussian_renderer as gr
import gaussian_renderer.workload_division as wd
import gaussian_renderer.loss_distribution as ld
from scene.gaussian_model import GaussianModel
from gs_b200 import synthetic as syn
from oracle.oracle import Oracle

W, H, N = 96, 64, 1500
orc = Oracle(np.float32)
bx, by, one = dgr._C.get_block_XY()                      # arguments/__init__.py:254-257
utils.set_block_size(bx, by, one)
utils.set_img_size(H, W)
args = Namespace(bsz=1, log_interval=50, log_folder="/tmp/gs_l3", zhx_debug=False, zhx_time=False, lambda_dssim=0.2,
                 lr_scale_loss=1.0, gaussians_distribution=True, image_distribution=True, local_sampling=False,
                 border_divpos_coeff=1.0, heuristic_decay=0.0, no_heuristics_update=False,
                 adjust_strategy_warmup_iterations=-1, adjust_strategy_warmp_iterations=-1, backend="default")
utils.set_args(args)
utils.set_cur_iter(1)
utils.GLOBAL_RANK, utils.LOCAL_RANK, utils.WORLD_SIZE = 0, 0, 1
class _Group:
    def size(self): return 1
    def rank(self): return 0
utils.DEFAULT_GROUP = utils.DP_GROUP = utils.MP_GROUP = utils.IN_NODE_GROUP = _Group()
class _Timers:
    def start(self, *a, **k): pass
    def stop(self, *a, **k): pass
utils.set_timers(_Timers())
utils.check_initial_gpu_memory_usage = lambda *a, **k: None

cam_d = syn.make_camera(W, H, yaw_deg=3.0)
scene = syn.make_scene(N, W, H, seed=11, radius_px=9.0)
gt = syn.make_gt_image(W, H, seed=5)
camera = types.SimpleNamespace(uid=0, image_height=H, image_width=W, FoVx=cam_d["FoVx"], FoVy=cam_d["FoVy"],
                               world_view_transform=torch.tensor(cam_d["viewmatrix"]),
                               full_proj_transform=torch.tensor(cam_d["projmatrix"]),
                               camera_center=torch.tensor(cam_d["campos"]), original_image=torch.tensor(gt))

# ---- a real reference GaussianModel with the raw parameter layout (scene/gaussian_model.py:219-228) ----------------
pc = GaussianModel(3)
pc.active_sh_degree = 3
op = np.clip(scene["opacities"], 1e-6, 1 - 1e-6)
P = torch.nn.Parameter
pc._xyz = P(torch.tensor(scene["means3D"]))
pc._features_dc = P(torch.tensor(scene["shs"][:, :1].copy()))
pc._features_rest = P(torch.tensor(scene["shs"][:, 1:].copy()))
pc._scaling = P(torch.log(torch.tensor(scene["scales"])))
pc._rotation = P(torch.tensor(scene["rotations"]))
pc._opacity = P(torch.log(torch.tensor(op) / (1 - torch.tensor(op))))

# ---- the operator boundary: recorders + oracle-backed autograd --------------------------------------------------------
calls = []
def cam_of(rs):
    return dict(image_height=rs.image_height, image_width=rs.image_width, tanfovx=rs.tanfovx, tanfovy=rs.tanfovy,
                viewmatrix=rs.viewmatrix.numpy(), projmatrix=rs.projmatrix.numpy(), campos=rs.campos.numpy(),
                sh_degree=rs.sh_degree)

class _Pre(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, scales, rotations, shs, opacities, rs):
        a = [t.detach().numpy() for t in (means3D, scales, rotations, shs, opacities)]
        pre = orc.preprocess_forward(*a, cam_of(rs), scale_modifier=rs.scale_modifier)
        ctx.a, ctx.rs, ctx.pre = a, rs, pre
        outs = (torch.tensor(pre["means2D"]), torch.tensor(pre["rgb"]), torch.tensor(pre["conic_opacity"]),
                torch.tensor(pre["radii"]), torch.tensor(pre["depths"]))
        ctx.mark_non_differentiable(outs[3], outs[4])
        return outs
    @staticmethod
    def backward(ctx, g_m2, g_rgb, g_co, *_):
        z = lambda g, s: np.zeros(s, np.float32) if g is None else g.numpy()
        n = ctx.a[0].shape[0]
        pb = orc.preprocess_backward(*ctx.a, cam_of(ctx.rs), ctx.pre["radii"], ctx.pre["clamped"], z(g_m2, (n, 2)),
                                     z(g_co, (n, 4)), z(g_rgb, (n, 3)))
        return (torch.tensor(pb["means3D"]), torch.tensor(pb["scales"]), torch.tensor(pb["rotations"]),
                torch.tensor(pb["shs"]), torch.tensor(pb["opacities"]), None)

class _Render(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means2D, conic_opacity, rgb, depths, radii, cl, rs):
        a = [t.detach().numpy() for t in (means2D, conic_opacity, rgb, depths, radii)]
        bg = tuple(float(v) for v in rs.bg)
        fwd = orc.render_forward(rs.image_height, rs.image_width, *a, cl.numpy().reshape(-1).astype(np.uint8), bg)
        ctx.a, ctx.rs, ctx.fwd, ctx.bg = a, rs, fwd, bg
        return torch.tensor(fwd["image"])
    @staticmethod
    def backward(ctx, g):
        rb = orc.render_backward(ctx.rs.image_height, ctx.rs.image_width, ctx.a[0], ctx.a[1], ctx.a[2], ctx.bg, ctx.fwd,
                                 g.contiguous().numpy())
        return torch.tensor(rb["means2D"]), torch.tensor(rb["conic_opacity"]), torch.tensor(rb["rgb"]), None, None, None, None

CUDA_ARGS_KEYS = {"mode", "world_size", "global_rank", "local_rank", "mp_world_size", "mp_rank", "log_folder",
                  "log_interval", "iteration", "zhx_debug", "zhx_time", "avoid_pixel_all2all", "stats_collector"}

def check_cuda_args(ca):
    assert set(ca) == CUDA_ARGS_KEYS, sorted(ca)
    for k in ("world_size", "global_rank", "local_rank", "mp_world_size", "mp_rank", "log_interval", "iteration",
              "zhx_debug", "zhx_time"):
        assert isinstance(ca[k], str), (k, type(ca[k]))                    # all str (SURVEY 8b)
    assert ca["mode"] == "train" and isinstance(ca["stats_collector"], dict) and ca["avoid_pixel_all2all"] is False

def preprocess_gaussians(self, *pos, **kw):
    assert not pos and set(kw) == {"means3D", "scales", "rotations", "shs", "opacities", "cuda_args"}, (pos, sorted(kw))
    rs = self.raster_settings
    assert rs._fields == ("image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix",
                          "projmatrix", "sh_degree", "campos", "prefiltered", "debug")
    assert isinstance(rs.image_height, int) and isinstance(rs.image_width, int) and rs.sh_degree == 3
    assert rs.viewmatrix.shape == (4, 4) and rs.projmatrix.shape == (4, 4) and rs.campos.shape == (3,) and rs.bg.shape == (3,)
    assert rs.prefiltered is False and rs.debug is False and rs.scale_modifier == 1.0
    n = kw["means3D"].shape[0]
    for name, shape in (("means3D", (n, 3)), ("scales", (n, 3)), ("rotations", (n, 4)), ("shs", (n, 16, 3)), ("opacities", (n, 1))):
        t = kw[name]
        assert t.dtype == torch.float32 and tuple(t.shape) == shape and t.requires_grad, (name, t.dtype, t.shape)
    # pre-applied activations (scene/gaussian_model.py:109-129): exp'd scales, unit quaternions, sigmoid'd opacity
    assert (kw["scales"] > 0).all() and torch.allclose(kw["rotations"].norm(dim=1), torch.ones(n), atol=1e-5)
    assert ((kw["opacities"] > 0) & (kw["opacities"] < 1)).all()
    check_cuda_args(kw["cuda_args"])
    calls.append("preprocess")
    return _Pre.apply(kw["means3D"], kw["scales"], kw["rotations"], kw["shs"], kw["opacities"], rs)

def render_gaussians(self, *pos, **kw):
    assert not pos and set(kw) == {"means2D", "conic_opacity", "rgb", "depths", "radii", "compute_locally",
                                   "extended_compute_locally", "cuda_args"}, (pos, sorted(kw))
    n = kw["means2D"].shape[0]
    for name, shape, dt in (("means2D", (n, 2), torch.float32), ("conic_opacity", (n, 4), torch.float32), ("rgb", (n, 3), torch.float32),
                            ("depths", (n,), torch.float32), ("radii", (n,), torch.int32)):
        assert tuple(kw[name].shape) == shape and kw[name].dtype == dt, (name, kw[name].shape, kw[name].dtype)
    cl = kw["compute_locally"]
    assert cl.dtype == torch.bool and tuple(cl.shape) == (utils.TILE_Y, utils.TILE_X) and cl.all()
    assert kw["extended_compute_locally"] is None                            # workload_division.py:802-803
    check_cuda_args(kw["cuda_args"])
    calls.append("render")
    img = _Render.apply(kw["means2D"], kw["conic_opacity"], kw["rgb"], kw["depths"], kw["radii"], cl, self.raster_settings)
    sc = kw["cuda_args"]["stats_collector"]                                  # mandatory even at W = 1 (:953-957)
    sc["forward_render_time"], sc["backward_render_time"] = 1.25, 2.5
    z = torch.zeros((), dtype=torch.int64)
    return img, z, z, z

dgr.GaussianRasterizer.preprocess_gaussians = preprocess_gaussians
dgr.GaussianRasterizer.render_gaussians = render_gaussians
assert gr.GaussianRasterizer is dgr.GaussianRasterizer

# ---- the reference's step, unmodified (train_internal.py:139-196) -----------------------------------------------------
dataset = types.SimpleNamespace(cameras=[camera])
history = wd.DivisionStrategyHistoryFinal(dataset, 1, 0)
strategies, gpuid2tasks = wd.start_strategy_final([camera], history)
assert strategies[0].gpu_ids == [0] and list(strategies[0].division_pos) == [0, utils.TILE_Y]
pipe = Namespace(debug=False)
bg = torch.zeros(3)
pkg = gr.distributed_preprocess3dgs_and_all2all_final([camera], pc, pipe, bg, batched_strategies=strategies, mode="train")
imgs, cls = gr.render_final(pkg, strategies)
collectors = [ca["stats_collector"] for ca in pkg["batched_cuda_args"]]
loss_sum, losses = ld.batched_loss_computation(imgs, [camera], cls, strategies, collectors)
loss_sum.backward()
assert calls == ["preprocess", "render"], calls
assert "forward_loss_time" in collectors[0]
assert pkg["batched_locally_preprocessed_mean2D"][0].grad is not None     # retain_grad(): densification reads it (:1050)
wd.finish_strategy_final([camera], history, strategies, collectors)      # KeyError if the collector keys were missing

# ---- the same step by the oracle's own harness ---------------------------------------------------------------------------
ref = orc.train_step(scene, cam_d, gt)
got = float(loss_sum)
assert abs(got - ref["loss"]) <= 2e-6 * abs(ref["loss"]), (got, ref["loss"])
e = ref["grads"]
q, gq = scene["rotations"], e["rotations"]
raw = {"_xyz": e["means3D"], "_features_dc": e["shs"][:, :1], "_features_rest": e["shs"][:, 1:],
       "_scaling": e["scales"] * scene["scales"], "_opacity": e["opacities"] * op * (1 - op),
       "_rotation": gq - q * (q * gq).sum(1, keepdims=True)}
for name, r in raw.items():
    g = getattr(pc, name).grad.numpy()
    rms = float(np.sqrt((r.astype(np.float64) ** 2).mean()))
    err = np.abs(g - r)
    assert (err <= 2e-4 * np.abs(r) + 2e-4 * rms).mean() > 0.999, (name, err.max(), rms)
print("L3-DRIVE-OK loss", got)

# ---- the LEGACY single-camera sequence (gaussian_renderer/__init__.py:66-174 replicated_preprocess3dgs, :458-507 render;
#      workload_division.py:100-199 DivisionStrategy): flat tile-range division, dist_global_strategy in cuda_args, and an
#      extended_compute_locally mask (the local tile range dilated by one tile row + 1) handed to render_gaussians -------
from gaussian_renderer.distribution_config import ImageDistributionConfig
args.image_distribution_config = ImageDistributionConfig("replicated_loss_computation", "DivisionStrategyUniform", False,
                                                         ["backward_render_time"])
for t in (pc._xyz, pc._features_dc, pc._features_rest, pc._scaling, pc._rotation, pc._opacity):
    t.grad = None
LEGACY = {"on": True, "ext": None, "cl": None}
CUDA_ARGS_KEYS.add("dist_global_strategy")
_orig_render = render_gaussians
def render_gaussians_legacy(self, *pos, **kw):
    ext, cl = kw["extended_compute_locally"], kw["compute_locally"]
    assert ext is not None and ext.dtype == torch.bool and tuple(ext.shape) == (utils.TILE_Y, utils.TILE_X)
    assert bool((ext | ~cl).all())                                           # the dilated region covers the local one
    assert isinstance(kw["cuda_args"]["dist_global_strategy"], str)
    LEGACY["ext"], LEGACY["cl"] = ext, cl
    # the drop-in's own argument validation of the legacy mask (gs_b200.ops.render_gaussians) on exactly these arguments:
    # a covering mask gets as far as the device check ("no CPU path"), a mask that does not cover compute_locally is rejected
    from gs_b200 import ops
    a = (kw["means2D"].detach(), kw["conic_opacity"].detach(), kw["rgb"].detach(), kw["depths"], kw["radii"], cl, self.raster_settings)
    try:
        ops.render_gaussians(*a, kw["cuda_args"], extended_compute_locally=ext)
        raise AssertionError("CPU tensors must not pass")
    except ValueError as e:
        assert "CUDA tensor" in str(e), e
    try:
        ops.render_gaussians(*a, kw["cuda_args"], extended_compute_locally=torch.zeros_like(ext))
        raise AssertionError("a mask that does not cover compute_locally must be rejected")
    except ValueError as e:
        assert "cover" in str(e), e
    kw = dict(kw, extended_compute_locally=None)
    return _orig_render(self, **kw)
dgr.GaussianRasterizer.render_gaussians = render_gaussians_legacy
calls.clear()
legacy_strategy = wd.DivisionStrategy(camera, 1, 0, utils.TILE_X, utils.TILE_Y,
                                      torch.ones((utils.TILE_Y, utils.TILE_X)), "DivisionStrategyUniform")
pkg1 = gr.replicated_preprocess3dgs(camera, pc, pipe, bg, strategy=legacy_strategy, mode="train")
img1, cl1 = gr.render(pkg1, legacy_strategy)
assert calls == ["preprocess", "render"], calls
assert LEGACY["ext"] is not None and bool(LEGACY["ext"].all()) and bool(cl1.all())    # W = 1: everything is local
assert torch.allclose(img1, torch.tensor(ref["fwd"]["image"]), atol=2e-6), float((img1 - torch.tensor(ref["fwd"]["image"])).abs().max())
wgt = torch.tensor(np.random.default_rng(2).normal(size=(3, H, W)).astype(np.float32))
(img1 * wgt).sum().backward()
assert pkg1["locally_preprocessed_mean2D"].grad is not None and pc._xyz.grad is not None
rb = orc.render_backward(H, W, ref["pre"]["means2D"], ref["pre"]["conic_opacity"], ref["pre"]["rgb"], (0.0, 0.0, 0.0),
                         ref["fwd"], wgt.numpy())
g2 = pkg1["locally_preprocessed_mean2D"].grad.numpy()
rms = float(np.sqrt((rb["means2D"].astype(np.float64) ** 2).mean()))
assert (np.abs(g2 - rb["means2D"]) <= 2e-4 * np.abs(rb["means2D"]) + 2e-4 * rms).mean() > 0.999
print("L3-LEGACY-OK")
'''


def test_reference_step_functions_run_over_the_dropin_boundary():
    code = CODE % dict(pkg=os.path.join(ROOT, "grendel-gs_b200"), shims=os.path.join(ROOT, "grendel-gs_b200", "shims"),
                       ref=REF, root=ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd="/tmp")
    assert r.returncode == 0, (r.stdout[-2000:], r.stderr[-4000:])
    assert "L3-DRIVE-OK" in r.stdout and "L3-LEGACY-OK" in r.stdout
