"""Hypothesis sharding across the GPUs of one box.

Hypotheses are independent, so rank r of G refines the contiguous block
[lo, hi) of the B hypotheses with the *global* B kept in the loss-mean divisor; the only
communication of a run is ONE `all_gather_into_tensor` of each rank's flat result buffer at its end, after
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


def flat_sizes(n, Bl, K):
    """Element offsets of [pose_hist n*Bl*7 | loss_hist n*Bl*K | final Bl*7] inside one rank's flat result buffer."""
    a = n * Bl * 7
    b = a + n * Bl * K
    return a, b, b + Bl * 7


def gather_flat(flat, per_rank_elems):
    """ONE `all_gather_into_tensor` of every rank's flat result buffer (padded to `per_rank_elems`) into a preallocated
    [world, per_rank_elems] tensor; world_size 1 returns the buffer itself as [1, elems]. gloo (CPU tests) stages through the host."""
    rank, ws = world()
    if ws == 1:
        return flat.view(1, -1)
    if flat.numel() < per_rank_elems:
        flat = torch.cat([flat, flat.new_zeros(per_rank_elems - flat.numel())])
    out = flat.new_empty(ws, per_rank_elems)
    if flat.is_cuda and dist.get_backend() != "nccl":
        h = out.cpu()
        dist.all_gather_into_tensor(h.view(-1), flat.cpu().contiguous())
        out.copy_(h)
    else:
        dist.all_gather_into_tensor(out.view(-1), flat.contiguous())
    return out


def unpack_flat(host, B, n, K):
    """[world, per_rank_elems] host tensor of flat rank buffers -> pose_hist [n,B,7], loss_hist [n,B,K], final [B,7]."""
    ws = host.shape[0]
    ph, lh, fin = [], [], []
    for r in range(ws):
        lo, hi = shard_bounds(B, r, ws)
        Bl = hi - lo
        if Bl == 0:
            continue
        a, b, c = flat_sizes(n, Bl, K)
        ph.append(host[r, :a].view(n, Bl, 7))
        lh.append(host[r, a:b].view(n, Bl, K))
        fin.append(host[r, b:c].view(Bl, 7))
    if len(ph) == 1:
        return ph[0], lh[0], fin[0]
    return torch.cat(ph, 1), torch.cat(lh, 1), torch.cat(fin, 0)


def final_poses(allf, B, n, K):
    """The [B,7] final poses out of the gathered [world, per_rank_elems] buffer, on its own device."""
    ws = allf.shape[0]
    fin = []
    for r in range(ws):
        lo, hi = shard_bounds(B, r, ws)
        if hi > lo:
            _, b, c = flat_sizes(n, hi - lo, K)
            fin.append(allf[r, b:c].view(hi - lo, 7))
    return fin[0] if len(fin) == 1 else torch.cat(fin, 0)


def gather_hypotheses(B, n, K, flat):
    """One rank's flat result buffer -> the whole job's (pose_hist [n,B,7], loss_hist [n,B,K], final [B,7]) on every rank."""
    rank, ws = world()
    per = (B + ws - 1) // ws
    allf = gather_flat(flat, max(flat_sizes(n, per, K)[2], 1))
    return unpack_flat(allf, B, n, K)


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
