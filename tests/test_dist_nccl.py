"""Two ranks, two GPUs, NCCL: hb.run_sweep with hb.distributed = True (circuits dealt round-robin to the ranks, one packed
all_gather_into_tensor over NVLink) must return, on every rank, exactly what one GPU returns for the whole sweep.
Skipped on boxes with a single GPU (the driver's test box); run with `gpurun --gpus 2`."""
import os
import socket

import numpy as np
import pytest

import circuits

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sweep(hb, yb, B):
    y = yb.YalRF('Differential Amplifier')
    h = circuits.diffamp(y)
    vin = np.linspace(100e-6, 50e-3, B)
    vb = np.linspace(1.0, 1.4, B)[::-1].copy()
    return hb.run_sweep(y, [(h['i2'], 'ac', vin / 2), (h['i4'], 'ac', vin / 2), (h['i3'], 'dc', vb), (h['i5'], 'dc', vb)])


def _worker(rank, world, port, B, out):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    import yalrf_b200
    from yalrf_b200.Analyses import MultiToneHarmonicBalance
    hb = MultiToneHarmonicBalance('HB1', 10e6, 10)
    hb.distributed = True
    res = _sweep(hb, yalrf_b200, B)
    np.savez(out % rank, V=res.V, Vf=res.Vf, conv=res.converged, it=res.newton_iters, lev=res.level_iters)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('B', [33])
def test_two_ranks_equal_one_gpu(tmp_path, B):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs two GPUs')
    import torch.multiprocessing as mp
    out = str(tmp_path / 'r%d.npz')
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    import yalrf_b200
    from yalrf_b200.Analyses import MultiToneHarmonicBalance
    hb = MultiToneHarmonicBalance('HB1', 10e6, 10)
    one = _sweep(hb, yalrf_b200, B)
    assert int((one.converged > 0).sum()) == B
    for r in range(2):
        g = np.load(out % r)
        np.testing.assert_array_equal(g['conv'], one.converged)
        np.testing.assert_array_equal(g['it'], one.newton_iters)
        np.testing.assert_array_equal(g['lev'], one.level_iters)
        np.testing.assert_array_equal(g['V'], one.V)          # same kernels on the same inputs: bit-identical
        np.testing.assert_array_equal(g['Vf'], one.Vf)
