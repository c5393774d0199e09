"""Hypothesis sharding across the GPUs of one box.

Hypotheses are independent, so rank r of G refines the contiguous block
[lo, hi) of the B hypotheses with the *global* B kept in the loss-mean divisor; the only
communication of a run is ONE all-gather of the packed per-hypothesis tables at its end, after
which every rank holds the same full tables and computes the same argmin. (When the learning-rate
multipliers are drawn, at construction / `set_batchsize`, rank 0's draw is broadcast once so the job
does not depend on each process's `random` state.)
With `torch.distributed` uninitialised (or world_size 1) everything is a no-op."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_bounds(B, rank, world_size):
    """Contiguous block of ceil(B/G) hypotheses per rank, in hypothesis order (so the
    reference's sequential random.uniform learning-rate draws map identically)."""
    per = (B + world_size - 1) // world_size
    lo = min(rank * per, B)
    return lo, min(lo + per, B)


def shard_range(B):
    rank, ws = world()
    return shard_bounds(B, rank, ws)


def _gather_dim(t_local, B, dim, per, ws):
    """all_gather equally padded shards along `dim`, trim to B."""
    pad = per - t_local.shape[dim]
    if pad > 0:
        shape = list(t_local.shape)
        shape[dim] = pad
        t_local = torch.cat([t_local, t_local.new_zeros(shape)], dim=dim)
    t_local = t_local.contiguous()
    parts = [torch.empty_like(t_local) for _ in range(ws)]
    dist.all_gather(parts, t_local)
    return torch.cat(parts, dim=dim).narrow(dim, 0, B).contiguous()


def gather_hypotheses(B, pose_hist, loss_hist, final):
    """pose_hist [n,Bl,7], loss_hist [n,Bl,K], final [Bl,7] -> global [n,B,7], [n,B,K], [B,7].
    ONE all-gather: the three tables travel packed as [n+1, Bl, 7+K] (the final poses are the extra row)."""
    rank, ws = world()
    if ws == 1:
        return pose_hist, loss_hist, final
    per = (B + ws - 1) // ws
    n, K = pose_hist.shape[0], loss_hist.shape[2]
    last = torch.cat([final, final.new_zeros(final.shape[0], K)], dim=1).unsqueeze(0)
    packed = torch.cat([torch.cat([pose_hist, loss_hist], dim=2), last], dim=0)
    allp = _gather_dim(packed, B, 1, per, ws)
    return allp[:n, :, :7].contiguous(), allp[:n, :, 7:].contiguous(), allp[n, :, :7].contiguous()


def broadcast_from_rank0(t):
    """Make rank 0's tensor (the randomly drawn learning-rate multipliers) the job's, in place: every rank then
    refines its block with the multipliers a single-GPU run with rank 0's `random` state would have drawn."""
    rank, ws = world()
    if ws > 1:
        if t.is_cuda and dist.get_backend() != "nccl":  # e.g. gloo in the CPU tests: stage through the host
            h = t.cpu()
            dist.broadcast(h, src=0)
            t.copy_(h)
        else:
            dist.broadcast(t, src=0)
    return t
