"""Border-pixel exchange between neighbouring tile-row strips.

Restates the row-strip case of /root/reference/gaussian_renderer/loss_distribution.py:601-972
(fast_distributed_loss_computation + _AddRemotePixelsToImage): the SSIM window is 11x11, so a strip's loss needs the
5 rendered rows just above and below it, which live on the neighbouring ranks.  Each rank sends its first / last 5
rows to the previous / next strip owner over NCCL point-to-point (NVLink), pastes what it receives around its strip,
and evaluates the loss with a window widened by those halo rows while summing only its own pixels
(ops.fused_l1_ssim(count_row0, count_row1)).  The sum of the strip losses then equals the full-image loss, and the
backward sends the halo rows' gradients back to the ranks that rendered them.
(The reference's flat tile-range partitions need L-shaped halos; the tile-ROW partition the live trainer uses --
workload_division.py:852-941 -- makes them plain row blocks.  The live path skips the exchange and zero-pads.)
"""
import torch
import torch.distributed as dist

HALF_WINDOW = 5


def _exchange(send_up, send_down, up_rank, down_rank, group):
    """send_up goes to the rank owning the strip above (receives that rank's bottom rows), send_down likewise."""
    ops, recv_up, recv_down = [], None, None
    if up_rank is not None:
        recv_up = torch.empty_like(send_up)
        ops += [dist.P2POp(dist.isend, send_up, up_rank, group), dist.P2POp(dist.irecv, recv_up, up_rank, group)]
    if down_rank is not None:
        recv_down = torch.empty_like(send_down)
        ops += [dist.P2POp(dist.isend, send_down, down_rank, group), dist.P2POp(dist.irecv, recv_down, down_rank, group)]
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()
    return recv_up, recv_down


class _AddRemoteRows(torch.autograd.Function):
    """image (3,H,W) with this rank's strip rows [y0,y1) rendered -> same tensor with rows [y0-5,y0) and [y1,y1+5)
    filled from the neighbouring strips."""

    @staticmethod
    def forward(ctx, image, y0, y1, up_rank, down_rank, group):
        h = HALF_WINDOW
        out = image.clone()
        send_up = image[:, y0:y0 + h].contiguous()
        send_down = image[:, y1 - h:y1].contiguous()
        recv_up, recv_down = _exchange(send_up, send_down, up_rank, down_rank, group)
        if recv_up is not None:
            out[:, y0 - h:y0] = recv_up
        if recv_down is not None:
            out[:, y1:y1 + h] = recv_down
        ctx.meta = (y0, y1, up_rank, down_rank, group)
        return out

    @staticmethod
    def backward(ctx, g):
        y0, y1, up_rank, down_rank, group = ctx.meta
        h = HALF_WINDOW
        g = g.contiguous()
        gin = g.clone()
        # gradients of the rows we received belong to the neighbours; theirs for our border rows come back
        send_up = g[:, y0 - h:y0].contiguous() if up_rank is not None else g[:, :h].contiguous()
        send_down = g[:, y1:y1 + h].contiguous() if down_rank is not None else g[:, :h].contiguous()
        recv_up, recv_down = _exchange(send_up, send_down, up_rank, down_rank, group)
        if up_rank is not None:
            gin[:, y0 - h:y0] = 0
            gin[:, y0:y0 + h] += recv_up
        if down_rank is not None:
            gin[:, y1:y1 + h] = 0
            gin[:, y1 - h:y1] += recv_down
        return gin, None, None, None, None, None


def add_remote_border_rows(image, strategy, image_height, group=None):
    """Returns (image with halo rows, window rows (r0, r1), counted rows (y0, y1)) for this rank's strip of one camera.
    Every rank that renders a strip of the camera must call it (neighbour pairs exchange point-to-point)."""
    y0, y1 = strategy.local_pixel_rows(image_height)
    if y1 - y0 < HALF_WINDOW:
        raise ValueError("strip thinner than the SSIM half window")
    c = strategy.rank
    up = strategy.gpu_ids[c - 1] if c > 0 else None
    down = strategy.gpu_ids[c + 1] if c + 1 < len(strategy.gpu_ids) else None
    out = _AddRemoteRows.apply(image, y0, y1, up, down, group)
    r0 = y0 - HALF_WINDOW if up is not None else y0
    r1 = min(y1 + HALF_WINDOW, image_height) if down is not None else y1
    return out, (r0, r1), (y0, y1)
