"""Multi-GPU sharding of a batch: independent circuits are split into contiguous slices, one per rank
(one process per GPU), solved without any collective inside the Newton loop, and the results are gathered once
(NCCL over NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md 8e."""
import numpy as np

from ._engine import BatchResult

_FIELDS = ('V', 'Vf', 'converged', 'newton_iters', 'lu_count', 'level_iters', 'level_alpha', 'err')


def shard_bounds(B, rank, world):
    """[lo, hi) of rank's contiguous slice; the first B % world ranks get one extra circuit."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def slice_block(pb, lo, hi):
    from ._flatten import ParamBlock
    out = ParamBlock.__new__(ParamBlock)
    out.topo = pb.topo
    out.B = hi - lo
    for name in ('tones', 'lin_val', 'src_val', 'src_ac', 'src_phase', 'diode_par', 'bjt_par', 'cubic_par'):
        setattr(out, name, np.ascontiguousarray(getattr(pb, name)[lo:hi]))
    return out


def gather_results(plan, local, B, group=None):
    """all_gather of the per-rank BatchResults into one BatchResult of the whole batch (on every rank)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nmax = max(shard_bounds(B, r, world)[1] - shard_bounds(B, r, world)[0] for r in range(world))
    use_cuda = dist.get_backend(group) == 'nccl'
    full = BatchResult(plan, B)
    for name in _FIELDS:
        a = getattr(local, name)
        isc = np.iscomplexobj(a)
        if isc:
            a = a.view(np.float64)
        pad = np.zeros((nmax,) + a.shape[1:], dtype=a.dtype)
        pad[:a.shape[0]] = a
        t = torch.from_numpy(pad)
        if use_cuda:
            t = t.cuda()
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t, group=group)
        dst = getattr(full, name)
        dstv = dst.view(np.float64) if isc else dst
        for r in range(world):
            lo, hi = shard_bounds(B, r, world)
            dstv[lo:hi] = outs[r][:hi - lo].cpu().numpy()
    return full


def run_block_distributed(plan, pb, options, V0=None, group=None, run_fn=None):
    """Solve this rank's slice of the block and gather.  run_fn(plan, pb, options, V0) -> BatchResult."""
    import torch.distributed as dist
    from . import _engine as E
    run_fn = run_fn or E.run_host
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    lo, hi = shard_bounds(pb.B, rank, world)
    v0 = None if V0 is None else np.asarray(V0).reshape(pb.B, -1)[lo:hi]
    local = run_fn(plan, slice_block(pb, lo, hi), options, v0) if hi > lo else BatchResult(plan, 0)
    return gather_results(plan, local, pb.B, group)
