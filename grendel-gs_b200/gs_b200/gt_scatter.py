"""Ground-truth strips for a step when only ONE rank of the node holds the dataset
(/root/reference/gaussian_renderer/loss_distribution.py:2395-2533, load_camera_from_cpu_to_all_gpu with
--distributed_dataset_storage): the first rank of the node copies the uint8 rows the node needs from its (pinned) host
memory to its GPU and sends every other rank exactly the rows of its strips with one batch of point-to-point
operations; the other ranks post the matching receives.  (Without --distributed_dataset_storage every rank reads its own
strips from its own host copy: pipeline.Trainer's default.)

Strip rows follow get_coverage_y_min/max (loss_distribution.py:2321-2330): tile rows [l, r) -> pixel rows
[16 l, min(16 r, H)).
"""
import torch
import torch.distributed as dist

BLOCK_Y = 16


def coverage(row_l, row_r, image_height):
    return row_l * BLOCK_Y, min(row_r * BLOCK_Y, image_height)


def scatter_gt_strips(gts_host, gpuid2tasks, image_height, device, rank, world, group=None, src=0):
    """gts_host: on rank `src` the list of (3,H,W) uint8 host tensors of the batch (ignored elsewhere);
    gpuid2tasks[gpu] = [(camera index, tile row l, tile row r), ...] (division.start_strategy).
    -> ({camera index: (3, rows, W) uint8 tensor on `device`} for this rank's tasks, bytes copied host -> device here)."""
    mine, ops, h2d = {}, [], 0
    if rank == src:
        on_dev = {}
        for tasks in gpuid2tasks:
            for cam, l, r in tasks:
                if cam not in on_dev:  # the rows any rank of the node needs of this camera, copied once
                    lo = min(coverage(t[1], t[2], image_height)[0] for ts in gpuid2tasks for t in ts if t[0] == cam)
                    hi = max(coverage(t[1], t[2], image_height)[1] for ts in gpuid2tasks for t in ts if t[0] == cam)
                    g = gts_host[cam][:, lo:hi, :]
                    h2d += g.numel()
                    on_dev[cam] = (lo, g.to(device, non_blocking=True))
        keep = []   # sent tensors must outlive the requests
        for dst, tasks in enumerate(gpuid2tasks):
            for cam, l, r in tasks:
                y0, y1 = coverage(l, r, image_height)
                lo, t = on_dev[cam]
                strip = t[:, y0 - lo:y1 - lo, :].contiguous()
                if dst == rank:
                    mine[cam] = strip
                else:
                    keep.append(strip)
                    ops.append(dist.P2POp(dist.isend, strip, dst, group))
    else:
        width = None
        for cam, l, r in gpuid2tasks[rank]:
            y0, y1 = coverage(l, r, image_height)
            if width is None:
                width = _width(gts_host)
            buf = torch.empty((3, y1 - y0, width), dtype=torch.uint8, device=device)
            mine[cam] = buf
            ops.append(dist.P2POp(dist.irecv, buf, src, group))
    if ops:
        for req in dist.batch_isend_irecv(ops):
            req.wait()
    return mine, h2d


def _width(hint):
    """Image width on a receiving rank: it knows the camera geometry, not the pixels."""
    if isinstance(hint, int):
        return hint
    return int(hint[0].shape[2])
