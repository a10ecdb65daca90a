"""Diagnostics: per-phase wall clock of Trainer.step (GS_B200_TRACE=1), 1 or N ranks."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "grendel-gs_b200")]
import torch, torch.distributed as dist
from gs_b200 import pipeline, synthetic as syn
os.environ["GS_B200_TRACE"] = "1"
rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local); dev = torch.device("cuda", local)
if world > 1: dist.init_process_group("nccl", device_id=dev)
W, H, N, B = 1920, 1080, 2_000_000, world
scene = syn.make_scene(N, W, H); cams = syn.make_batch_cameras(W, H, B)
gts = [torch.from_numpy(syn.make_gt_image(W, H, seed=1 + k)).pin_memory() for k in range(B)]
tr = pipeline.Trainer(scene, cams, gts, dev, rank, world)
for _ in range(3): tr.step()
tr.trace = {}
n = 10
for _ in range(n): tr.step()
if rank == 0: print("[trace] ms/step per phase (each phase synchronised):", {k: round(v / n, 3) for k, v in tr.trace.items()})
if world > 1: dist.destroy_process_group()
