"""Plan cache and batch solve calls on top of the C ABI."""
import ctypes as C

import numpy as np

from . import _lib as L
from ._flatten import ParamBlock, Topology, dft_tables, source_kinds

from collections import OrderedDict

_plans = OrderedDict()   # (device, topology, flags) -> Plan, least recently used first
MAX_PLANS = 64           # each plan pins device tables + the staging buffer of its largest batch


def clear_plans():
    """Destroy every cached plan (their device tables and staging buffers are freed)."""
    _plans.clear()


def _current_device():
    """CUDA device the plan will live on (its tables are device memory); -1 without a GPU -- yhb_plan_create then
    fails with 'no CPU fallback'."""
    try:
        import torch
        return torch.cuda.current_device() if torch.cuda.is_available() else -1
    except Exception:
        return -1


class Plan:
    """One compiled topology (yhb_plan) + its grid sizes."""

    def __init__(self, topo, src_kinds, flags, tables=True):
        lib = L.lib()
        self.topo = topo
        self.flags = flags
        d = L.PlanDesc()
        d.num_nodes = topo.N
        d.num_tones = topo.num_tones
        d.K1 = topo.nh[0]
        d.K2 = topo.nh[1] if topo.num_tones == 2 else 0
        keep = []

        def ip(a):
            a = L.i32(a)
            keep.append(a)
            return a.ctypes.data_as(C.POINTER(C.c_int32))

        d.num_lin = len(topo.lin)
        d.lin_kind = ip([k for k, _, _ in topo.lin] or [0])
        d.lin_nodes = ip([n for _, n, _ in topo.lin] or [[0, 0, 0, 0]])
        d.num_src = len(topo.src)
        d.src_kind = ip(src_kinds or [0])
        d.src_nodes = ip([n for n, _ in topo.src] or [[0, 0]])
        d.num_diode = len(topo.diodes)
        d.diode_nodes = ip([n for n, _ in topo.diodes] or [[0, 0]])
        d.num_bjt = len(topo.bjts)
        d.bjt_nodes = ip([n for n, _ in topo.bjts] or [[0, 0, 0]])
        d.num_cubic = len(topo.cubics)
        d.cubic_nodes = ip([n for n, _ in topo.cubics] or [[0, 0]])
        d.nl_kind = ip([k for k, _ in topo.nl_order] or [0])
        d.nl_index = ip([i for _, i in topo.nl_order] or [0])
        if topo.num_tones == 1:
            self.K = topo.nh[0]
        else:
            self.K = (2 * topo.nh[1] + 1) * topo.nh[0] + topo.nh[1]
        self.S = 2 * self.K + 1
        self.N = topo.N
        self.W = self.S * self.N
        self.IDFT, self.DFT = dft_tables(self.K)
        idft, dft = L.f64(self.IDFT), L.f64(self.DFT)
        d.idft = idft.ctypes.data_as(C.POINTER(C.c_double))
        d.dft = dft.ctypes.data_as(C.POINTER(C.c_double))
        d.flags = flags
        h = C.c_void_p()
        L.check(lib.yhb_plan_create(C.byref(d), C.byref(h)))
        self.handle = h

    def core(self):
        """(Nc, Wc, fused): dense core left after peeling the ideal-source rows, and the engine that will run"""
        a, b, c = C.c_int32(), C.c_int32(), C.c_int32()
        L.check(L.lib().yhb_plan_core(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, bool(c.value)

    def band(self):
        """half bandwidth (matrix elements) of the staged engine's band storage, 0 = dense core"""
        return int(L.lib().yhb_plan_band(self.handle))

    def engine(self):
        """'fused' | 'staged-dense' | 'staged-banded' | 'staged-block-thomas': the engine yhb_solve runs for this plan"""
        return L.ENGINES[int(L.lib().yhb_plan_engine(self.handle))]

    def block_sparse(self):
        """levels of the fused engine's block-sparse LU over node blocks (0 = dense Gauss-Jordan of the core)"""
        return int(L.lib().yhb_plan_block_sparse(self.handle, None))

    def flops(self):
        """(reference, dense core, executed) floating-point operations of one Newton-step linear solve"""
        a, b, c = C.c_double(), C.c_double(), C.c_double()
        L.check(L.lib().yhb_plan_flops(self.handle, C.byref(a), C.byref(b), C.byref(c)))
        return a.value, b.value, c.value

    def freqs(self, f1, f2=0.0):
        out = np.zeros(self.K + 1)
        L.check(L.lib().yhb_plan_freqs(self.handle, float(f1), float(f2), out.ctypes.data_as(C.POINTER(C.c_double))))
        return out

    def __del__(self):
        try:
            if self.handle:
                L.lib().yhb_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass


def get_plan(netlist, freq, numharmonics, flags=0, match_tones=True):
    """(plan, topology) for this netlist shape and tone grid; plans are cached by topology."""
    if isinstance(freq, float):
        num_tones, nh = 1, (int(numharmonics),)
    elif isinstance(freq, list) and len(freq) == 2:
        num_tones, nh = 2, (int(numharmonics[0]), int(numharmonics[1]))
    else:  # MultiToneHarmonicBalance.py:247-275 (ints fall through to the error print there)
        raise ValueError('ERROR: only one and two-tones are currently supported (freq must be a float or a list '
                         'of two floats)')
    topo = Topology(netlist, num_tones, nh)
    kinds = source_kinds(topo, freq, match_tones)
    dev = _current_device()
    key = (dev,) + topo.key(kinds, flags)
    plan = _plans.get(key)
    if plan is None:
        if dev >= 0:
            L.check(L.lib().yhb_set_device(dev))
        plan = Plan(topo, kinds, flags)
        _plans[key] = plan
        while len(_plans) > MAX_PLANS:   # Plan.__del__ calls yhb_plan_destroy once the last user drops it
            _plans.popitem(last=False)
    else:
        _plans.move_to_end(key)
    return plan, topo


def _host_array(shape, dtype, pinned, zero=True):
    """host array; page-locked (through torch's caching host allocator) when a GPU is present so that H2D / D2H
    copies run at full speed.  zero=False: the caller overwrites every element (results of a D2H copy)."""
    if pinned:
        import torch
        tdt = {np.float64: torch.float64, np.int32: torch.int32, np.complex128: torch.complex128}[dtype]
        make = torch.zeros if zero else torch.empty
        return make(shape, dtype=tdt, pin_memory=True).numpy()
    return np.zeros(shape, dtype=dtype)


class BatchResult:
    """Results of a batched solve (numpy, host)."""

    def __init__(self, plan, B, pinned=False, levels=True):
        self.B = B
        # every array is written in full by the D2H copies of yhb_run_host / the gather of _dist: no zero fill
        self.V = _host_array((B, plan.W), np.float64, pinned, zero=False)
        self.Vf = _host_array((B, plan.N, plan.K + 1), np.complex128, pinned, zero=False)
        self.converged = _host_array((B,), np.int32, pinned, zero=False)
        self.newton_iters = _host_array((B,), np.int32, pinned, zero=False)
        self.lu_count = _host_array((B,), np.int32, pinned, zero=False)
        # 'NR number of iterations' / 'Alpha level' of every hb_loop call: 768 B per circuit, optional
        self.level_iters = _host_array((B, L.MAX_LEVELS), np.int32, pinned, zero=False) if levels else None
        self.level_alpha = _host_array((B, L.MAX_LEVELS), np.float64, pinned, zero=False) if levels else None
        self.err = _host_array((B,), np.float64, pinned, zero=False)
        self.Inl = None
        self.Ic = None

    def levels(self, b=0):
        if self.level_iters is None:
            raise RuntimeError('per-level reports were not requested (hb.report_levels = False)')
        it = self.level_iters[b]
        return it[it > 0].tolist()


def make_params(pb):
    p = L.Params()
    p.batch = pb.B
    p.tones = L.ptr(pb.tones)
    p.lin_val = L.ptr(pb.lin_val)
    p.src_val = L.ptr(pb.src_val)
    p.diode_par = L.ptr(pb.diode_par)
    p.bjt_par = L.ptr(pb.bjt_par)
    p.cubic_par = L.ptr(pb.cubic_par)
    return p


def make_options(options):
    o = L.Options()
    o.reltol = float(options['reltol'])
    o.abstol = float(options['abstol'])
    o.maxiter = int(options['maxiter'])
    return o


def run_host(plan, pb, options, V0=None, want_inl=False, want_levels=True, want_ic=False):
    """run() for a batch through yhb_run_host: host arrays in, host arrays out (H2D / D2H inside)."""
    pb.finalize()
    res = BatchResult(plan, pb.B, pinned=pb.B >= 256, levels=want_levels)     # page-locked results for large batches
    res.Inl = np.zeros((pb.B, plan.W)) if want_inl else None
    nbjt = len(plan.topo.bjts)
    res.Ic = np.zeros((pb.B, nbjt, plan.S)) if (want_ic and nbjt) else None
    r = L.Result()
    r.V = L.ptr(res.V)
    r.Vf = L.ptr(res.Vf)
    r.converged = L.ptr(res.converged)
    r.newton_iters = L.ptr(res.newton_iters)
    r.lu_count = L.ptr(res.lu_count)
    r.level_iters = L.ptr(res.level_iters)
    r.level_alpha = L.ptr(res.level_alpha)
    r.err = L.ptr(res.err)
    r.Inl = L.ptr(res.Inl)
    r.Ic = L.ptr(res.Ic)
    p = make_params(pb)
    o = make_options(options)
    v0 = None
    if V0 is not None:
        v0 = L.f64(np.asarray(V0).reshape(pb.B, plan.W))
    L.check(L.lib().yhb_run_host(plan.handle, C.byref(p), C.byref(o), L.ptr(v0), C.byref(r)))
    res.core_W = plan.core()[1]
    res.params_bytes = sum(getattr(pb, n).nbytes for n in ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par',
                                                            'cubic_par')) + (0 if v0 is None else v0.nbytes)
    return res


class DeviceBatch:
    """A parameter block resident in HBM (torch owns the buffers) + result / workspace buffers for repeated solves."""

    def __init__(self, plan, pb, pinned_upload=True):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError('yalrf_b200 needs a CUDA device: there is no CPU fallback')
        pb.finalize()
        self.plan, self.B = plan, pb.B
        self.dev = torch.device('cuda', torch.cuda.current_device())
        L.check(L.lib().yhb_set_device(self.dev.index))
        self.t = {n: torch.from_numpy(getattr(pb, n)).to(self.dev) for n in
                  ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par', 'cubic_par')}
        self.params = L.Params()
        self.params.batch = pb.B
        for n, v in self.t.items():
            setattr(self.params, n, v.data_ptr())
        B = pb.B
        f64, i32 = torch.float64, torch.int32
        self.V = torch.zeros((B, plan.W), dtype=f64, device=self.dev)
        self.Vf = torch.zeros((B, plan.N, plan.K + 1, 2), dtype=f64, device=self.dev)
        self.converged = torch.zeros(B, dtype=i32, device=self.dev)
        self.newton_iters = torch.zeros(B, dtype=i32, device=self.dev)
        self.lu_count = torch.zeros(B, dtype=i32, device=self.dev)
        self.level_iters = torch.zeros((B, L.MAX_LEVELS), dtype=i32, device=self.dev)
        self.level_alpha = torch.zeros((B, L.MAX_LEVELS), dtype=f64, device=self.dev)
        self.err = torch.zeros(B, dtype=f64, device=self.dev)
        self.Vdc = torch.zeros((B, plan.N), dtype=f64, device=self.dev)
        self.dcinfo = torch.zeros((B, 2), dtype=i32, device=self.dev)
        self.dc_done = False
        self.ws_bytes = L.lib().yhb_workspace_bytes(plan.handle, B)
        self.ws = torch.empty(self.ws_bytes, dtype=torch.uint8, device=self.dev)
        r = L.Result()
        r.V, r.Vf = self.V.data_ptr(), self.Vf.data_ptr()
        r.converged, r.newton_iters, r.lu_count = (self.converged.data_ptr(), self.newton_iters.data_ptr(),
                                                   self.lu_count.data_ptr())
        r.level_iters, r.level_alpha, r.err = (self.level_iters.data_ptr(), self.level_alpha.data_ptr(),
                                               self.err.data_ptr())
        r.Inl = None
        self.result = r

    def solve(self, options, have_v0=False, stream=None, with_dc=False):
        """DC points + run() for the whole batch on `stream` (default: torch's current stream); no D2H.
        with_dc: also compute the DC points for a warm start (needed when a direct attempt fails, :333-335)."""
        import torch
        lib = L.lib()
        st = (stream or torch.cuda.current_stream()).cuda_stream
        o = make_options(options)
        if not have_v0 and not with_dc:
            # cold start: DC analysis + ladder as one call (the DC kernel runs beside the HB kernel on the fused engine)
            L.check(lib.yhb_solve_cold(self.plan.handle, C.byref(self.params), C.byref(o), self.ws.data_ptr(), self.ws_bytes,
                                       self.Vdc.data_ptr(), self.dcinfo.data_ptr(), C.byref(self.result), st))
            self.dc_done = True
            return
        if with_dc:
            L.check(lib.yhb_dc_solve(self.plan.handle, C.byref(self.params), self.ws.data_ptr(), self.ws_bytes,
                                     self.Vdc.data_ptr(), self.dcinfo.data_ptr(), st))
            self.dc_done = True
        L.check(lib.yhb_solve(self.plan.handle, C.byref(self.params), C.byref(o), self.ws.data_ptr(), self.ws_bytes,
                              1 if have_v0 else 0, self.Vdc.data_ptr() if self.dc_done else None,
                              C.byref(self.result), st))


def to_time(Vf, freqs, t):
    """Time-domain waveforms of phasors on the GPU (yhb_to_time): Vf[B][N][K1] complex, freqs[B][K1] (or [K1]), t[nt]
    -> x[B][N][nt].  The reference does this in Python loops (convert_to_time :125-135, HB_TwoTones.ipynb cell 3)."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError('yalrf_b200 needs a CUDA device: there is no CPU fallback')
    Vf = np.ascontiguousarray(Vf, dtype=np.complex128)
    B, N, K1 = Vf.shape
    f = np.array(np.broadcast_to(np.asarray(freqs, dtype=np.float64), (B, K1)))   # a writable, contiguous copy
    t = np.ascontiguousarray(t, dtype=np.float64)
    dev = torch.device('cuda', torch.cuda.current_device())
    L.check(L.lib().yhb_set_device(dev.index))
    dV = torch.from_numpy(Vf.view(np.float64).reshape(B, N, K1, 2)).to(dev)
    df, dt = torch.from_numpy(f).to(dev), torch.from_numpy(t).to(dev)
    x = torch.empty((B, N, len(t)), dtype=torch.float64, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    L.check(L.lib().yhb_to_time(B, N, K1, dV.data_ptr(), df.data_ptr(), len(t), dt.data_ptr(), x.data_ptr(), st))
    return x.cpu().numpy()
