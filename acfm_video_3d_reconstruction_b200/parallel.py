"""Data-parallel plumbing of the hot path: one process per GPU, frames x camera hypotheses sharded by CLIP, one
all-reduce of the shared-parameter gradients per step.

The reference's only parallelism is single-process torch.nn.DataParallel around the model and the renderer
(/root/reference/multiframe/main.py:172,184-193): per forward it broadcasts the parameters, scatters the render
inputs and gathers (N,H,W) masks and (N,H,W,20) int64 pix_to_face onto GPU 0.  Here nothing on the data path
crosses GPUs: every (frame, hypothesis) render is independent, the hypothesis softmax is per frame, and the
flow / texture-cycle losses only couple adjacent frames of one clip, so a rank owns whole clips and renders all G
hypotheses of its frames locally.  Only parameters shared by all frames (encoder / heads / handle weights `lbs` /
template `mean_v` / `vert2kp`) need their gradients summed: one bucketed all-reduce over NCCL (NVLink 5 / NVSwitch);
per-frame embeddings (cameras, deforms, hypothesis probabilities) are touched by the owning rank only.
"""
import torch
import torch.distributed as dist


def shard_clips(num_clips, rank, world_size):
    """Contiguous, balanced [lo, hi) range of clip indices owned by `rank` (sizes differ by at most one)."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} / world_size {world_size}")
    base, rem = divmod(num_clips, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_frames(num_clips, frames_per_clip, rank, world_size):
    """Indices (into the flattened B*T frame axis, clip-major as `img.reshape(B*T, ...)`, multiframe/main.py:341-346)
    of the frames owned by `rank`: whole clips, so adjacent frames stay together."""
    lo, hi = shard_clips(num_clips, rank, world_size)
    return torch.arange(lo * frames_per_clip, hi * frames_per_clip)


def multiplex_rows(frame_idx, num_frames_total, G):
    """Rows of a hypothesis-major (G * B*T, ...) tensor (n = g * B*T + bt, multiframe/main.py:578,609) that belong to
    the frames `frame_idx`; the result is again hypothesis-major over the local frames."""
    g = torch.arange(G, device=frame_idx.device)[:, None] * num_frames_total
    return (g + frame_idx[None, :]).reshape(-1)


def allreduce_shared_grads(params, group=None, bucket_bytes=64 << 20):
    """Sum the .grad of the shared parameters over all ranks in place, in flat buckets of at most `bucket_bytes`
    (a handful of large NCCL calls instead of one per tensor; on NVSwitch the cost is launch latency, not links).
    Parameters whose .grad is None on this rank contribute zeros.  Returns the number of collectives issued."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    params = [p for p in params if p.requires_grad]
    calls, bucket, size = 0, [], 0

    def flush():
        nonlocal calls, bucket, size
        if not bucket:
            return
        for p in bucket:
            if p.grad is None:
                p.grad = torch.zeros_like(p)
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for p in bucket:
            n = p.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        calls += 1
        bucket, size = [], 0

    by_dtype = {}
    for p in params:
        by_dtype.setdefault((p.dtype, p.device), []).append(p)
    for plist in by_dtype.values():
        for p in plist:
            nbytes = p.numel() * p.element_size()
            if bucket and size + nbytes > bucket_bytes:
                flush()
            bucket.append(p)
            size += nbytes
        flush()
    return calls


def max_over_ranks(value, device, group=None):
    """Max of a python float over ranks (device-side timings are reported as the slowest rank's)."""
    t = torch.tensor([float(value)], device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t[0])
