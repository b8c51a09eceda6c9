"""Multi-GPU sharding of a batch: independent circuits are dealt round-robin to the ranks (one process per GPU) --
circuit b goes to rank b % world, so that neighbouring sweep points, which need similar numbers of Newton
iterations, spread over all ranks and the ranks finish together -- solved without any collective inside the Newton
loop, and the results are gathered ONCE: every rank packs its results into one [n][R] fp64 buffer (on the GPU,
straight from the solver's output buffers) and a single all_gather_into_tensor moves it (NCCL over NVLink on the GPU
box, gloo in the CPU tests).  SURVEY.md 8e."""
import numpy as np

from ._engine import BatchResult


def shard_indices(B, rank, world):
    """global indices of the circuits rank solves: b % world == rank"""
    return np.arange(rank, B, world)


def shard_count(B, rank, world):
    return (B - rank + world - 1) // world if B > rank else 0


def slice_block(pb, idx):
    """the ParamBlock of the circuits idx (an index array or a slice)"""
    from ._flatten import ParamBlock
    out = ParamBlock.__new__(ParamBlock)
    out.topo = pb.topo
    for name in ('tones', 'lin_val', 'src_val', 'src_ac', 'src_phase', 'diode_par', 'bjt_par', 'cubic_par'):
        setattr(out, name, np.ascontiguousarray(getattr(pb, name)[idx]))
    out.B = out.tones.shape[0]
    return out


class PackLayout:
    """columns of the packed result row of one circuit (fp64; the int32 reports are exact in fp64)"""

    def __init__(self, plan, levels=True):
        from . import _lib as L
        self.fields = [('V', plan.W), ('Vf', plan.N * (plan.K + 1) * 2), ('err', 1), ('converged', 1),
                       ('newton_iters', 1), ('lu_count', 1)]
        if levels:
            self.fields += [('level_iters', L.MAX_LEVELS), ('level_alpha', L.MAX_LEVELS)]
        self.levels = levels
        self.off = {}
        o = 0
        for name, n in self.fields:
            self.off[name] = (o, o + n)
            o += n
        self.R = o


def pack_host(res, lay):
    """BatchResult (host) -> [n][R] fp64"""
    out = np.zeros((res.B, lay.R))
    for name, _ in lay.fields:
        a = getattr(res, name)
        if np.iscomplexobj(a):
            a = np.ascontiguousarray(a).view(np.float64)
        lo, hi = lay.off[name]
        out[:, lo:hi] = np.asarray(a, dtype=np.float64).reshape(res.B, -1)
    return out


def pack_device(db, lay):
    """DeviceBatch (results in HBM) -> [n][R] fp64 cuda tensor, one fused copy kernel per field"""
    import torch
    out = torch.empty((db.B, lay.R), dtype=torch.float64, device=db.dev)
    for name, _ in lay.fields:
        lo, hi = lay.off[name]
        out[:, lo:hi] = getattr(db, name).reshape(db.B, -1)
    return out


def unpack(plan, rows, lay):
    """[B][R] fp64 (host) -> BatchResult"""
    B = rows.shape[0]
    full = BatchResult(plan, B, levels=lay.levels)
    for name, _ in lay.fields:
        lo, hi = lay.off[name]
        dst = getattr(full, name)
        if np.iscomplexobj(dst):
            dst.view(np.float64).reshape(B, -1)[:] = rows[:, lo:hi]
        else:
            dst.reshape(B, -1)[:] = rows[:, lo:hi]
    return full


def gather_packed(local, B, group=None):
    """local: this rank's [n][R] rows (numpy, or a cuda tensor with the nccl backend).  Returns the [B][R] rows of the
    whole batch in circuit order on the host (every rank), from ONE all_gather_into_tensor."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    nmax = shard_count(B, 0, world)
    use_cuda = dist.get_backend(group) == 'nccl'
    t = torch.from_numpy(np.ascontiguousarray(local)) if isinstance(local, np.ndarray) else local
    if use_cuda and not t.is_cuda:
        t = t.cuda()
    R = t.shape[1]
    if t.shape[0] < nmax:  # ranks with one circuit less: pad to the common row count
        t = torch.cat([t, torch.zeros((nmax - t.shape[0], R), dtype=t.dtype, device=t.device)])
    out = torch.empty((world * nmax, R), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    if use_cuda:
        host = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
        host.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        out = host
    rows = out.numpy().reshape(world, nmax, R)
    b = np.arange(B)
    return rows[b % world, b // world]     # circuit b = row b // world of rank b % world


def gather_unpack_device(plan, local, B, lay, group=None, root=None):
    """NCCL path: local = this rank's [n][R] rows in HBM.  ONE all_gather_into_tensor, the rows put into circuit order
    by one device copy (circuit b = row b // world of rank b % world: a transpose of the [world][nmax] row grid), every
    field split off on the device and copied straight into the page-locked arrays of the BatchResult -- no host-side
    repacking of the gathered results."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    nmax = shard_count(B, 0, world)
    R = lay.R
    t = local
    if t.shape[0] < nmax:  # ranks with one circuit less: pad to the common row count
        t = torch.cat([t, torch.zeros((nmax - t.shape[0], R), dtype=t.dtype, device=t.device)])
    out = torch.empty((world * nmax, R), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    if root is not None and dist.get_rank(group) != root:
        return None                                  # only the root reads the gathered results back to the host
    rows = out.view(world, nmax, R).transpose(0, 1).reshape(world * nmax, R)[:B]      # circuit order
    full = BatchResult(plan, B, pinned=True, levels=lay.levels)
    keep = []
    for name, _ in lay.fields:
        lo, hi = lay.off[name]
        dst = getattr(full, name)
        col = rows[:, lo:hi]
        if dst.dtype == np.int32:
            col = col.to(torch.int32)               # exact: the reports were packed from int32
        col = col.contiguous()
        keep.append(col)
        host = torch.from_numpy(dst.view(np.float64) if np.iscomplexobj(dst) else dst).reshape(B, -1)
        host.copy_(col.reshape(B, -1), non_blocking=True)
    torch.cuda.current_stream().synchronize()
    return full


def run_block_distributed(plan, pb, options, V0=None, group=None, run_fn=None, levels=True, root=None, total=None):
    """Solve this rank's circuits of the block and gather.  run_fn(plan, pb, options, V0) -> BatchResult (host) is
    for tests; the default keeps parameters and results in HBM (DeviceBatch) and packs / gathers on the device.
    root: None = every rank returns the whole batch's results; a rank number = only that rank does (the others return
    None and skip the device-to-host copy of the gathered rows).  total: pb already holds only THIS rank's circuits
    (shard_indices(total, rank, world)) of a batch of `total` circuits."""
    import torch.distributed as dist
    from . import _engine as E
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    B = pb.B if total is None else total
    idx = shard_indices(B, rank, world)
    lay = PackLayout(plan, levels)
    v0 = None if V0 is None else np.asarray(V0, dtype=np.float64).reshape(B, -1)[idx]
    mine = (lambda: pb) if total is not None else (lambda: slice_block(pb, idx))
    if len(idx) == 0:
        local = np.zeros((0, lay.R))
    elif run_fn is not None:
        local = pack_host(run_fn(plan, mine(), options, v0), lay)
    else:
        import torch
        db = E.DeviceBatch(plan, mine(), pinned_upload=True)
        if v0 is not None:
            db.V.copy_(torch.from_numpy(v0), non_blocking=False)
        db.solve(options, have_v0=v0 is not None)
        if v0 is not None and bool((db.converged < 0).any().item()):
            # a direct attempt failed and needs the DC point (:333-335): rerun from the same V0 with it
            db.V.copy_(torch.from_numpy(v0))
            db.solve(options, have_v0=True, with_dc=True)
        local = pack_device(db, lay)
        if dist.get_backend(group) == 'nccl':
            return gather_unpack_device(plan, local, B, lay, group, root)
    rows = gather_packed(local, B, group)
    if root is not None and rank != root:
        return None
    return unpack(plan, rows, lay)
