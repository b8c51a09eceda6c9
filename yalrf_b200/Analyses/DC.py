"""DC operating point (yalrf/Analyses/DC.py:35-123) computed by the batched GPU kernel k_dc.

Only what harmonic balance needs: node voltages of the HB-reachable devices.  Returns the node voltages
as an (N, 1) array (the reference appends inductor branch currents; they are not used by calc_V0,
MultiToneHarmonicBalance.py:697-704)."""
import ctypes as C

import numpy as np

from .. import _lib as L
from .._engine import get_plan, make_params
from .._flatten import ParamBlock


def dc_solve_batch(plan, pb):
    """Vdc[B][N], info[B][2] = (converged, Newton iterations) for a parameter block."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('yalrf_b200 needs a CUDA device: there is no CPU fallback')
    lib = L.lib()
    pb.finalize()
    dev = torch.device('cuda', torch.cuda.current_device())
    L.check(lib.yhb_set_device(dev.index))
    t = {n: torch.from_numpy(getattr(pb, n)).to(dev) for n in
         ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par', 'cubic_par')}
    p = L.Params()
    p.batch = pb.B
    for n, v in t.items():
        setattr(p, n, v.data_ptr())
    nbytes = lib.yhb_workspace_bytes(plan.handle, pb.B)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    vdc = torch.empty((pb.B, plan.N), dtype=torch.float64, device=dev)
    info = torch.empty((pb.B, 2), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L.check(lib.yhb_dc_solve(plan.handle, C.byref(p), ws.data_ptr(), nbytes, vdc.data_ptr(), info.data_ptr(), st))
    torch.cuda.synchronize()
    return vdc.cpu().numpy(), info.cpu().numpy()


class DC():

    def __init__(self, name):
        self.name = name
        self.x = None
        self.converged = None
        self.iterations = None

    def get_dc_solution(self):
        return self.x

    def run(self, netlist, x0=None, nodeset=None):
        plan, topo = get_plan(netlist, 1.0, 1, match_tones=False)
        pb = ParamBlock(topo, 1.0, 1)
        vdc, info = dc_solve_batch(plan, pb)
        self.x = vdc[0].reshape(-1, 1)
        self.converged = bool(info[0, 0])
        self.iterations = int(info[0, 1])
        return self.x
