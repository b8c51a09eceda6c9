"""World-size-2 (and 3) gloo tests of the multi-GPU path on CPU: round-robin sharding of a batch over ranks
(circuit b -> rank b % world) and the single packed end-of-solve gather.  The GPU solve itself is replaced by a deterministic stand-in (there is no CPU
fallback); what is tested is the host logic every rank runs around it."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

import circuits


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_run(plan, pb, options, V0):
    """Stand-in for E.run_host: results are simple functions of the circuit's own parameters."""
    from yalrf_b200._engine import BatchResult
    pb.finalize()
    res = BatchResult(plan, pb.B)
    key = pb.src_val[:, 1, 1]                     # I2.ac of each circuit
    res.V[:] = key[:, None] * np.arange(1, plan.W + 1)[None, :]
    res.Vf[:] = (key[:, None, None] * (1 + 2j))
    res.converged[:] = 1
    res.newton_iters[:] = np.round(key * 1e4).astype(np.int32)
    res.level_iters[:] = 0
    res.level_iters[:, 0] = 3
    res.level_alpha[:] = 0.25
    res.lu_count[:] = 7
    res.err[:] = key
    return res


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import yalrf_b200
    from yalrf_b200 import _dist
    from yalrf_b200._flatten import ParamBlock, Topology, dft_tables

    class FakePlan:                                # plan geometry only; no CUDA
        N, K, S, W = 10, 10, 21, 210

    y = yalrf_b200.YalRF('x')
    h = circuits.diffamp(y)
    topo = Topology(y, 1, (10,))
    pb = ParamBlock(topo, 10e6, B)
    pb.set(h['i2'], 'ac', np.linspace(1e-3, 2e-3, B))
    res = _dist.run_block_distributed(FakePlan, pb, {}, None, run_fn=_fake_run)
    idx = _dist.shard_indices(B, rank, world)
    # the form hb.run_sweep uses: the block already holds this rank's circuits only; results on rank 0 only
    res2 = _dist.run_block_distributed(FakePlan, _dist.slice_block(pb, idx), {}, None, run_fn=_fake_run, root=0, total=B)
    if rank == 0:
        assert np.array_equal(res2.V, res.V) and np.array_equal(res2.Vf, res.Vf) and np.array_equal(res2.err, res.err)
    else:
        assert res2 is None
    if rank == 0:
        np.savez(out, V=res.V, Vf=res.Vf, it=res.newton_iters, conv=res.converged, err=res.err, lev=res.level_iters,
                 alpha=res.level_alpha, lu=res.lu_count)
    assert len(idx) == _dist.shard_count(B, rank, world) and len(idx) in (B // world, B // world + 1)
    assert res.converged.dtype == np.int32 and res.level_iters.dtype == np.int32
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,B', [(2, 16), (2, 7), (3, 10)])
def test_shard_and_gather(tmp_path, world, B):
    out = str(tmp_path / 'g.npz')
    mp.spawn(_worker, args=(world, _free_port(), B, out), nprocs=world, join=True)
    g = np.load(out)
    key = np.linspace(1e-3, 2e-3, B)
    np.testing.assert_allclose(g['V'], key[:, None] * np.arange(1, 211)[None, :], rtol=1e-15)
    np.testing.assert_allclose(g['Vf'], np.broadcast_to(key[:, None, None] * (1 + 2j), (B, 10, 11)), rtol=1e-15)
    assert g['conv'].tolist() == [1] * B
    np.testing.assert_array_equal(g['it'], np.round(key * 1e4).astype(np.int32))
    np.testing.assert_array_equal(g['lev'][:, :2], np.tile([3, 0], (B, 1)))
    np.testing.assert_array_equal(g['alpha'], np.full((B, 64), 0.25))
    np.testing.assert_array_equal(g['lu'], np.full(B, 7))
    np.testing.assert_array_equal(g['err'], key)


def test_shards_cover_everything():
    from yalrf_b200._dist import shard_count, shard_indices
    for B in (0, 1, 5, 64, 4096, 4097):
        for world in (1, 2, 3, 8):
            parts = [shard_indices(B, r, world) for r in range(world)]
            assert sorted(np.concatenate(parts).tolist()) == list(range(B))
            assert [len(p) for p in parts] == [shard_count(B, r, world) for r in range(world)]
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1
            assert shard_count(B, 0, world) == max(len(p) for p in parts)
