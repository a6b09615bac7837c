"""Offline multi-GPU plumbing: whole sequences shard across ranks, one process per GPU.

Inside a sequence the path is a strict recurrence (pose_t needs the map of t-1), so there is no collective in the
frame loop.  Two exchanges exist, both outside it:
  scatter_blobs       rank 0 holds one byte blob per rank (a .klg log, or a sequence descriptor) and hands each rank its
                      own: a broadcast size table + one padded uint8 scatter (NCCL over NVLink on GPUs, gloo on CPU)
  gather_trajectories every rank's [frames, 12] (R row-major, t) trajectory back to rank 0, ragged lengths allowed
Tensors live on `device` ("cuda" with the nccl backend, "cpu" with gloo: the CPU tests run this file at world size 2).
"""
import torch
import torch.distributed as dist


def _world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_available() and dist.is_initialized() else (0, 1)


def assign_sequences(n_sequences, world):
    """round-robin partition of sequence ids over ranks: rank r gets r, r + world, ..."""
    return [list(range(r, n_sequences, world)) for r in range(world)]


def scatter_blobs(blobs, device="cpu", src=0):
    """blobs: on rank `src` a list of world_size bytes objects, elsewhere None.  Returns this rank's bytes."""
    rank, world = _world()
    if world == 1:
        return bytes(blobs[0])
    sizes = torch.zeros(world, dtype=torch.int64, device=device)
    if rank == src:
        if len(blobs) != world:
            raise ValueError(f"need one blob per rank ({world}), got {len(blobs)}")
        sizes = torch.tensor([len(b) for b in blobs], dtype=torch.int64, device=device)
    dist.broadcast(sizes, src=src)
    pad = max(int(sizes.max().item()), 1)
    mine = torch.zeros(pad, dtype=torch.uint8, device=device)
    chunks = None
    if rank == src:
        chunks = []
        for b in blobs:
            t = torch.zeros(pad, dtype=torch.uint8)
            if len(b):
                t[:len(b)] = torch.frombuffer(bytearray(b), dtype=torch.uint8)
            chunks.append(t.to(device))
    dist.scatter(mine, chunks, src=src)
    return bytes(mine[:int(sizes[rank].item())].cpu().numpy().tobytes())


def gather_trajectories(traj, dst=0):
    """traj: float32 [n_frames, 12] on this rank's device.  Returns on rank `dst` a list of per-rank [n_r, 12] CPU tensors
    (rank order), elsewhere None."""
    rank, world = _world()
    traj = traj.contiguous()
    if traj.dim() != 2 or traj.shape[1] != 12:
        raise ValueError("trajectory must be [frames, 12]")
    if world == 1:
        return [traj.detach().cpu()]
    n = torch.tensor([traj.shape[0]], dtype=torch.int64, device=traj.device)
    counts = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    pad = max(max(counts), 1)
    mine = torch.zeros(pad, 12, dtype=torch.float32, device=traj.device)
    mine[:traj.shape[0]] = traj
    out = [torch.zeros_like(mine) for _ in range(world)] if rank == dst else None
    dist.gather(mine, out, dst=dst)
    if rank != dst:
        return None
    return [o[:c].cpu() for o, c in zip(out, counts)]


def max_over_ranks(value_ms, device="cpu"):
    """device-timed milliseconds -> the max over ranks (what the bench reports)"""
    rank, world = _world()
    t = torch.tensor([float(value_ms)], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
