"""Frame sharding for multi-GPU rendering.

The forward path shards by frame: every (camera, gaze) frame depends only on the read-only model, so the model is
replicated on each GPU, frames are dealt round-robin to ranks and NO data-path collective exists.  The only
communication is gathering per-rank timings (and optionally images) — `torch.distributed` over NCCL on GPUs, gloo in
the CPU tests.  The backward pass (per-Gaussian gradient reduction) stays single-GPU per BASELINE.json north_star.
"""
import torch
import torch.distributed as dist


def frames_for_rank(n_frames, rank, world):
    """Frame indices rendered by `rank`: rank, rank+world, ... (covers 0..n_frames-1 exactly once over all ranks)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    return list(range(rank, n_frames, world))


def frame_assignment(n_cameras, n_gazes, rank, world):
    """(camera, gaze) pairs of the reference FPS protocol (9 gazes x test cameras, render_compose_gazes_fps.py:26,56)
    owned by `rank`."""
    out = []
    for f in frames_for_rank(n_cameras * n_gazes, rank, world):
        out.append((f % n_cameras, f // n_cameras))
    return out


def gather_timings(local_ms, n_local_frames, device="cpu", group=None):
    """All ranks learn every rank's (elapsed ms, frame count).  Returns a [world, 2] float64 tensor."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    mine = torch.tensor([float(local_ms), float(n_local_frames)], dtype=torch.float64, device=device)
    if world == 1:
        return mine.unsqueeze(0).cpu()
    buf = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(buf, mine, group=group)
    return torch.stack(buf).cpu()


def aggregate_fps(table):
    """Whole-job frames/s = total frames / slowest rank's time (the job ends when the last rank ends)."""
    total = float(table[:, 1].sum())
    slowest = float(table[:, 0].max())
    return total / (slowest * 1e-3) if slowest > 0 else 0.0


def gather_images(image, dst=0, group=None):
    """Rank `dst` receives every rank's rendered frame ([3,H,W] fp32, 24.9 MB at 1080p): one `gather` collective — NCCL over
    NVLink on GPUs (device tensors never touch the host), gloo in the CPU tests.  Returns a [world,3,H,W] tensor on `dst`
    (rank order = frame order of one round of the round-robin deal) and None elsewhere."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return image.unsqueeze(0)
    rank = dist.get_rank(group)
    image = image.contiguous()
    out = torch.empty((world,) + tuple(image.shape), dtype=image.dtype, device=image.device) if rank == dst else None
    dist.gather(image, list(out.unbind(0)) if rank == dst else None, dst=dst, group=group)
    return out
