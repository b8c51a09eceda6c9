"""Netlist -> (topology, per-circuit parameter arrays) for libyhb.

The topology (node indices, device kinds, tone grid) selects a cached plan; the parameter arrays are
rebuilt from the device objects at every run, because the reference re-reads them at every run
(users mutate ``dev.ac``, ``dev.options[...]``, ``hb.freq`` in between).
"""
import numpy as np

from . import _lib as L


class Topology:
    """Hashable description of a netlist shape + tone grid."""

    def __init__(self, netlist, num_tones, nh):
        self.N = netlist.get_num_nodes() - 1
        self.num_tones = num_tones
        self.nh = tuple(nh)
        self.lin, self.src, self.diodes, self.bjts, self.cubics, self.nl_order = [], [], [], [], [], []
        lin_kinds = {'resistor': L.LIN_RESISTOR, 'capacitor': L.LIN_CAPACITOR, 'inductor': L.LIN_INDUCTOR,
                     'gyrator': L.LIN_GYRATOR, 'ihf': L.LIN_IHF}
        for dev in netlist.get_devices():
            k = dev.kind
            if k in lin_kinds:
                nodes = (dev.n1, dev.n2, dev.n3, dev.n4) if k == 'gyrator' else (dev.n1, dev.n2, 0, 0)
                self.lin.append((lin_kinds[k], nodes, dev))
            elif k == 'isource':
                self.src.append(((dev.n1, dev.n2), dev))
            elif k == 'diode':
                self.nl_order.append((0, len(self.diodes)))
                self.diodes.append(((dev.n1, dev.n2), dev))
            elif k == 'bjt':
                self.nl_order.append((1, len(self.bjts)))
                self.bjts.append(((dev.n1, dev.n2, dev.n3), dev))
            elif k == 'cubicnl':
                self.nl_order.append((2, len(self.cubics)))
                self.cubics.append(((dev.n1, dev.n2), dev))
            else:
                raise NotImplementedError('device kind %r is not supported by the HB engine' % k)

    def key(self, src_kinds, flags):
        return (self.N, self.num_tones, self.nh, tuple((k, n) for k, n, _ in self.lin),
                tuple(n for n, _ in self.src), tuple(src_kinds), tuple(n for n, _ in self.diodes),
                tuple(n for n, _ in self.bjts), tuple(n for n, _ in self.cubics), tuple(self.nl_order), flags)


def source_kinds(topo, freq, match_tones=True):
    """calc_Is' tone matching (MultiToneHarmonicBalance.py:652-663) -- exact float equality, as there.
    match_tones=False (DC analysis only): AC sources are open circuits, any frequency is accepted."""
    kinds = []
    f1 = freq if topo.num_tones == 1 else freq[0]
    f2 = None if topo.num_tones == 1 else freq[1]
    for _, dev in topo.src:
        if dev.itype == 'dc':
            kinds.append(L.SRC_DC)
        elif dev.itype == 'ac':
            if not match_tones or dev.freq == f1:
                kinds.append(L.SRC_AC1)
            elif f2 is not None and dev.freq == f2:
                kinds.append(L.SRC_AC2)
            else:
                raise ValueError('AC source %r: freq %r matches no analysis tone %r (the reference leaves fidx '
                                 'unbound here, MultiToneHarmonicBalance.py:655-663)' % (dev.name, dev.freq, freq))
        else:
            kinds.append(L.SRC_DC_ONLY)
    return kinds


def _zeros(shape, pinned, fill=True):
    """float64 host array (zeroed unless the caller overwrites all of it); page-locked for large batches on a GPU box so
    that the H2D copies of yhb_run_host are asynchronous DMA transfers instead of staged pageable copies"""
    if pinned:
        import torch
        make = torch.zeros if fill else torch.empty
        return make(shape, dtype=torch.float64, pin_memory=True).numpy()
    return np.zeros(shape) if fill else np.empty(shape)


def _want_pinned(B):
    if B < 256:
        return False
    import torch
    return torch.cuda.is_available()


class ParamBlock:
    """Parameter arrays of a batch of B circuits sharing one topology."""

    def __init__(self, topo, freq, B=1):
        self.topo = topo
        self.B = B
        pinned = _want_pinned(B)
        # one template row per array (the netlist's current values, re-read at every run), broadcast over the batch
        # with a single contiguous copy each
        f = (freq, 0.0) if topo.num_tones == 1 else (freq[0], freq[1])
        lin = np.zeros((max(len(topo.lin), 1), 2))
        for i, (kind, _, dev) in enumerate(topo.lin):
            if kind == L.LIN_RESISTOR:
                lin[i, 0] = dev.R
            elif kind == L.LIN_CAPACITOR:
                lin[i, 0] = dev.C
            elif kind == L.LIN_INDUCTOR:
                lin[i, 0] = dev.L
            elif kind == L.LIN_GYRATOR:
                lin[i, 0] = dev.G
            else:
                lin[i, 0] = dev.freq
                lin[i, 1] = dev.g
        nsrc = max(len(topo.src), 1)
        src, ac, ph = np.zeros((nsrc, 3)), np.zeros(nsrc), np.zeros(nsrc)
        for i, (_, dev) in enumerate(topo.src):
            src[i, 0] = dev.dc
            ac[i] = dev.ac
            ph[i] = dev.phase
        dio = np.zeros((max(len(topo.diodes), 1), L.DIODE_NPAR))
        for i, (_, dev) in enumerate(topo.diodes):
            dio[i, :] = [float(dev.options[k]) for k in L.DIODE_PARAMS]
        bjt = np.zeros((max(len(topo.bjts), 1), L.BJT_NPAR))
        for i, (_, dev) in enumerate(topo.bjts):
            bjt[i, :-1] = [float(dev.options[k]) for k in L.BJT_PARAMS[:-1]]
            bjt[i, -1] = float(dev.type)
        cub = np.zeros(max(len(topo.cubics), 1))
        for i, (_, dev) in enumerate(topo.cubics):
            cub[i] = dev.alpha

        def batch(row):
            out = _zeros((B,) + row.shape, pinned, fill=False)
            out[:] = row
            return out

        self.tones = batch(np.array(f, dtype=np.float64))
        self.lin_val = batch(lin)
        self.src_val = batch(src)
        self.src_ac = np.tile(ac, (B, 1))
        self.src_phase = np.tile(ph, (B, 1))
        self.diode_par = batch(dio)
        self.bjt_par = batch(bjt)
        self.cubic_par = batch(cub)

    # ---- sweeps: override one field of one device for the whole batch
    def set(self, dev, field, values):
        values = np.asarray(values, dtype=np.float64)
        t = self.topo
        hit = False
        for i, (kind, _, d) in enumerate(t.lin):
            if d is dev and field in ('R', 'C', 'L', 'G', 'freq', 'g'):
                self.lin_val[:, i, 1 if field == 'g' else 0] = values
                hit = True
        for i, (_, d) in enumerate(t.src):
            if d is dev:
                if field == 'dc':
                    self.src_val[:, i, 0] = values
                elif field == 'ac':
                    self.src_ac[:, i] = values
                elif field == 'phase':  # degrees, like the constructor argument (CurrentSource.py:12)
                    self.src_phase[:, i] = np.radians(values)
                else:
                    raise KeyError(field)
                hit = True
        if hasattr(dev, 'options'):
            # users share one options dict between devices (q2.options = q1.options, tests/HB_DiffAmp.py:67-69)
            for i, (_, d) in enumerate(t.diodes):
                if (d.options is dev.options or getattr(d, '_options_of', None) is dev.options) and field in L.DIODE_PARAMS:
                    if field == 'Rs' and getattr(d, '_options_of', None) is not None:
                        raise KeyError('Rs of a diode expanded by hb.diode_rs is a resistor of its own: sweep that instead')
                    self.diode_par[:, i, L.DIODE_PARAMS.index(field)] = values
                    hit = True
            for i, (_, d) in enumerate(t.bjts):
                if d.options is dev.options and field in L.BJT_PARAMS:
                    self.bjt_par[:, i, L.BJT_PARAMS.index(field)] = values
                    hit = True
        for i, (_, d) in enumerate(t.cubics):
            if d is dev and field == 'alpha':
                self.cubic_par[:, i] = values
                hit = True
        if not hit:
            raise KeyError('no parameter %r on device %r' % (field, getattr(dev, 'name', dev)))

    def set_tones(self, f1, f2=None):
        self.tones[:, 0] = f1
        if f2 is not None:
            self.tones[:, 1] = f2

    def finalize(self):
        # iac = ac * exp(j phase)  (MultiToneHarmonicBalance.py:653)
        ph = self.src_phase
        if ph.shape[0] > 1 and np.array_equal(ph, np.broadcast_to(ph[0], ph.shape)):
            iac = self.src_ac * np.exp(1j * ph[0])[None, :]   # one phase per source for the whole batch: the same values
        else:
            iac = self.src_ac * np.exp(1j * ph)
        self.src_val[:, :, 1] = iac.real
        self.src_val[:, :, 2] = iac.imag
        for name in ('tones', 'lin_val', 'src_val', 'diode_par', 'bjt_par', 'cubic_par'):
            setattr(self, name, np.ascontiguousarray(getattr(self, name), dtype=np.float64))
        return self

    @staticmethod
    def stack(blocks):
        out = ParamBlock.__new__(ParamBlock)
        out.topo = blocks[0].topo
        out.B = sum(b.B for b in blocks)
        for name in ('tones', 'lin_val', 'src_val', 'src_ac', 'src_phase', 'diode_par', 'bjt_par', 'cubic_par'):
            setattr(out, name, np.concatenate([getattr(b, name) for b in blocks], axis=0))
        return out


def dft_tables(K):
    """IDFT / DFT exactly as the reference builds them (MultiToneHarmonicBalance.py:294-302), vectorised."""
    S = 2 * K + 1
    s = np.arange(S)[:, None]
    k = np.arange(1, K + 1)[None, :]
    IDFT = np.ones((S, S))
    IDFT[:, 1::2] = np.cos(2 * np.pi * k * s / S)
    IDFT[:, 2::2] = np.sin(2 * np.pi * k * s / S)
    DFT = np.linalg.inv(IDFT)
    return IDFT, DFT
