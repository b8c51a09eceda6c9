"""MultiToneHarmonicBalance with the reference's surface (yalrf/Analyses/MultiToneHarmonicBalance.py:23-728):
``run`` / ``run_oscillator`` / ``print_v`` / ``get_v`` / ``convert_to_time`` / ``plot_v`` and the attributes user
scripts read (``V``, ``Vf``, ``freqs``, ``K``, ``S``, ``N``, ``W``, ``DFT``, ``IDFT``, ``num_tones``).

The Newton loop itself runs on the GPU through libyhb (include/yhb.h); this class only flattens the netlist
into parameter arrays, calls the C ABI and formats results.  ``run_batch`` / ``run_sweep`` are the batched
forms the reference expresses as Python ``for`` loops around ``hb.run`` (tests/HB_DiffAmp.py:113-121).
"""
import numpy as np

from .. import _engine as E
from .. import _lib as L
from .._flatten import ParamBlock

options = dict()
options['reltol'] = 1e-3
options['abstol'] = 1e-6
options['maxiter'] = 100


class _RsZeroOptions:
    """read-through view of a diode's options with Rs = 0 (the series resistance is a device of its own then)"""

    def __init__(self, base):
        self.base = base

    def __getitem__(self, k):
        return 0.0 if k == 'Rs' else self.base[k]

    def keys(self):
        return self.base.keys()


def _expand_diode_rs(netlist):
    """Extension (hb.diode_rs): a copy of the netlist in which every diode with Rs > 0 is a resistor Rs / Area (Diode.init,
    Diode.py:96) from its anode to an internal node -- appended behind the netlist's own nodes, so their indices stay --
    in series with a twin of the diode that reads the same options with Rs = 0."""
    from ..Devices import Diode, Resistor
    out = netlist.copy()
    devs = []
    for dev in out.devices:
        if dev.kind == 'diode' and float(dev.options['Rs']) > 0.:
            ni = out.add_node(dev.name + '.rs')
            devs.append(Resistor(dev.name + '.Rs', dev.n1, ni, float(dev.options['Rs']) / float(dev.options['Area'])))
            twin = Diode(dev.name, ni, dev.n2)
            twin.options = _RsZeroOptions(dev.options)
            twin._options_of = dev.options      # sweeps of the original diode's options reach the twin (ParamBlock.set)
            devs.append(twin)
        else:
            devs.append(dev)
    out.devices = devs
    return out


class MultiToneHarmonicBalance:

    def __init__(self, name, freq=1e9, numharmonics=10):
        self.name = name
        self.freq = freq
        self.numharmonics = numharmonics
        self.options = options.copy()
        # extensions, off by default = reference behaviour
        self.pnp_extension = False    # honour BJT polarity in HB (reference: every BJT is NPN, BJT.py:495-503)
        self.exact_jacobian = False   # re-zero dQdV each iteration (reference accumulates it, :413)
        self.diode_charge = False     # diodes carry their junction / diffusion charge in HB (reference: g and I only, :441-475)
        self.diode_rs = False         # a diode's series resistance Rs counts in DC and HB: an internal node + a resistor Rs / Area
                                      # (the reference's get_mthb_params ignores Rs, Diode.py:395-427); run / run_batch / run_sweep
        self.verbose = False          # print the reference's per-level log lines (:354, :358, :383-384, :597-598)
        self.engine = 'auto'          # 'auto': fused on-chip engine when the circuit fits one SM; 'staged': HBM Jacobian
        self.peel = True              # eliminate ideal-source rows symbolically before the dense factorisation
        self.band = True              # staged engine: band storage for block-banded cores (chains of nodes)
        self.block_thomas = None      # staged engine, block-tridiagonal cores: None = block-by-block solve for large
                                      # blocks (S >= 128), True / False = always / never
        self.block_sparse = None      # fused engine: None = block-sparse LU over node blocks when the core's block
                                      # structure is sparse, True / False = whenever it applies / never (dense solve)
        self.distributed = False      # run_batch / run_sweep: shard the circuits over torch.distributed ranks, gather
        self.gather_root = None       # distributed: None = every rank gets the whole batch's results; r = only rank r
                                      # (the other ranks' run_batch / run_sweep return None)
        self.report_levels = True     # run_batch / run_sweep: copy the per-level reports back (768 B per circuit)
        self.last = None              # BatchResult of the last call

    # ------------------------------------------------------------------ helpers
    def _flags(self):
        return ((L.FLAG_PNP_EXTENSION if self.pnp_extension else 0) | (L.FLAG_EXACT_JACOBIAN if self.exact_jacobian else 0) | (L.FLAG_DIODE_CHARGE if self.diode_charge else 0) |
                (0 if self.peel else L.FLAG_NO_PEEL) | (L.FLAG_STAGED if self.engine == 'staged' else 0) |
                (0 if self.band else L.FLAG_NO_BAND) |
                (L.FLAG_BLOCK_THOMAS if self.block_thomas else 0) |
                (L.FLAG_NO_BLOCK_THOMAS if self.block_thomas is False else 0) |
                (L.FLAG_BLOCK_SPARSE if self.block_sparse else 0) |
                (L.FLAG_NO_BLOCK_SPARSE if self.block_sparse is False else 0))

    def _setup(self, netlist):
        self.netlist = _expand_diode_rs(netlist) if self.diode_rs else netlist.copy()
        self.lin_devs = self.netlist.get_linear_devices()
        self.nonlin_devs = self.netlist.get_nonlinear_devices()
        plan, topo = E.get_plan(self.netlist, self.freq, self.numharmonics, self._flags())
        self._plan = plan
        # :247-279
        if topo.num_tones == 1:
            self.num_tones = 1
            self.f1 = self.freq
        else:
            self.num_tones = 2
            self.f1, self.f2 = self.freq
            self.K1, self.K2 = self.numharmonics
            self.l0 = self.f1 / (2 * self.K2 + 1)
        self.K, self.S, self.N, self.W = plan.K, plan.S, plan.N, plan.W
        self.IDFT, self.DFT = plan.IDFT, plan.DFT
        return plan, topo

    def _freqs(self, plan):
        return plan.freqs(self.f1, self.f2 if self.num_tones == 2 else 0.0)

    def get_node_idx(self, node):
        return self.netlist.get_node_idx(node) - 1

    def _log(self, res, b=0):
        if not self.verbose:
            return
        its = res.level_iters[b]
        n = int(np.count_nonzero(its))
        ladder = 0
        for l in range(n):
            a = res.level_alpha[b, l]
            direct = (l == 0 and self._had_v0)
            if not direct:
                print('Alpha level: {}'.format(a))
                ladder += 1
            print('NR number of iterations: {}'.format(int(its[l])))
        if ladder:
            print('Final convergence: {}'.format(bool(res.converged[b])))

    # ------------------------------------------------------------------ solve
    def run(self, netlist, V0=None):
        """:234-405.  Returns (converged, freqs, Vf, None, None); failure -> (False, None, None, None, None)."""
        plan, topo = self._setup(netlist)
        pb = ParamBlock(topo, self.freq, 1)
        self._had_v0 = V0 is not None
        res = E.run_host(plan, pb, self.options, None if V0 is None else np.asarray(V0, dtype=np.float64),
                         want_ic=True)
        self.last = res
        self._log(res)
        if res.Ic is not None:  # dev.Ic side channel (:512), read by user scripts (tests/ColpittsCB_Analysis.py:96-101)
            for q, (_, dev) in enumerate(topo.bjts):
                dev.Ic = res.Ic[0, q].copy()
        if not res.converged[0]:
            return False, None, None, None, None
        self.freqs = self._freqs(plan)
        if self.num_tones == 2:
            self.freqs = np.abs(self.freqs)  # :389-390
        self.V = res.V[0].reshape(self.W, 1).copy()
        self.Vf = res.Vf[0].copy()
        return True, self.freqs, self.Vf, None, None

    def run_batch(self, netlists, V0=None):
        """Independent solves of netlists sharing one topology (cold starts unless V0[B][W] is given)."""
        plan, topo = self._setup(netlists[0])
        blocks = [ParamBlock(topo, self.freq, 1)]
        key0 = plan
        for nl in netlists[1:]:
            p2, t2 = E.get_plan(nl, self.freq, self.numharmonics, self._flags())
            if p2 is not key0:
                raise ValueError('run_batch: all netlists must share one topology')
            blocks.append(ParamBlock(t2, self.freq, 1))
        pb = ParamBlock.stack(blocks)
        return self._run_block(plan, pb, V0)

    def run_sweep(self, netlist, sweeps, V0=None, freq=None):
        """Batched parameter sweep.  ``sweeps`` = [(device, field, values[B]), ...]: ``field`` is an attribute the
        reference scripts assign between runs (``ac``, ``dc``, ``phase`` [deg], ``R``, ``C``, ``L``, ``alpha``) or a
        key of ``device.options`` (applied to every device sharing that dict).  ``freq`` (optional [B] or [B][2])
        sweeps the analysis tone(s); AC sources at a tone follow it (their kind is fixed in the plan), and so does
        every IdealHarmonicFilter whose ``freq`` equals the nominal tone -- unless the filter's own ``freq`` is
        swept explicitly."""
        plan, topo = self._setup(netlist)
        B = len(np.asarray(sweeps[0][2])) if sweeps else len(np.asarray(freq))
        total = None
        if self.distributed:
            # build the parameter block of THIS rank's circuits only (b % world == rank)
            import torch.distributed as dist
            from .._dist import shard_indices
            idx = shard_indices(B, dist.get_rank(), dist.get_world_size())
            sweeps = [(dev, field, np.asarray(values, dtype=np.float64)[idx]) for dev, field, values in sweeps]
            if freq is not None:
                freq = np.asarray(freq, dtype=np.float64)[idx]
            total, B = B, len(idx)
        pb = ParamBlock(topo, self.freq, B)
        for dev, field, values in sweeps:
            pb.set(dev, field, values)
        if freq is not None:
            f = np.asarray(freq, dtype=np.float64)
            swept = {id(dev) for dev, field, _ in sweeps if field == 'freq'}
            nominal = [self.freq] if topo.num_tones == 1 else list(self.freq)
            cols = [f] if f.ndim == 1 else [f[:, 0], f[:, 1]]
            for dev in [d for k, _, d in topo.lin if k == L.LIN_IHF and id(d) not in swept]:
                for f_nom, col in zip(nominal, cols):
                    if dev.freq == f_nom:  # the filter's exact-equality test (IdealHarmonicFilter.py:34) must keep hitting
                        pb.set(dev, 'freq', col)
            pb.set_tones(*cols)
        return self._run_block(plan, pb, V0, total)

    def _run_block(self, plan, pb, V0, total=None):
        self._had_v0 = V0 is not None
        if self.distributed:
            from .._dist import run_block_distributed
            res = run_block_distributed(plan, pb, self.options, V0, levels=self.report_levels, root=self.gather_root,
                                        total=total)
            if res is None:
                self.last = None
                return None
        else:
            res = E.run_host(plan, pb, self.options, V0, want_levels=self.report_levels)
        res.freqs = self._freqs(plan)
        if self.num_tones == 2:
            res.freqs = np.abs(res.freqs)
        self.last = res
        return res

    # ------------------------------------------------------------------ oscillator (:137-232)
    def run_oscillator(self, netlist, f0, numharmonics, V0, node, useprev=False):
        """Autonomous-oscillator analysis (:137-232), same arguments, prints and return value as the reference.  The
        reference's formulation is kept -- oscillator probe (AC current source + gyrator = voltage source, in series
        with an IdealHarmonicFilter that is a short only at the fundamental), find (fosc, Vosc) with Yosc = probe
        current / probe voltage = 0 (:182-196) -- but the outer iteration that the reference leaves to
        scipy.optimize.fmin (:204-212) runs on the GPU: yhb_oscillator_host, a 2 x 2 Newton iteration with the exact
        Jacobian (csrc/yhb_osc.cuh).  ``hb.osc_method = 'fmin'`` replays the reference's Nelder-Mead instead (every
        objective evaluation one GPU solve), which reproduces its evaluation path."""
        if getattr(self, 'osc_method', 'newton') == 'fmin':
            return self._run_oscillator_fmin(netlist, f0, numharmonics, V0, node, useprev)
        from .. import _osc
        if useprev and getattr(self, 'fosc', None) is not None:
            f0, V0 = self.fosc, self.Vosc      # the reference warm-starts from self.V (:168-169); here: the last point
        res = _osc.solve(self, netlist, float(f0), numharmonics, float(V0), node)
        self.last_osc = res
        self.netlist = res.netlist
        fosc, Vosc = float(res.fosc[0]), float(res.Vosc[0])
        probe_i, probe_z = res.probe
        probe_i.ac, probe_i.freq, probe_z.freq = Vosc, fosc, fosc       # the probe is left at the result (:222-224)
        self.freq = fosc
        self.fosc, self.Vosc = fosc, Vosc
        self.osc_yosc = float(res.yosc[0])
        print('Frequency of oscillation = {} Hz'.format(fosc))
        print('Oscillation amplitude = {} V'.format(Vosc))
        if not (res.converged[0] and res.hb_converged[0] > 0):
            return False, None, None, None, None
        self.freqs = res.freqs[0]
        self.V = res.V[0].reshape(self.W, 1).copy()
        self.Vf = res.Vf[0].copy()
        return True, self.freqs, self.Vf, None, None

    def _run_oscillator_fmin(self, netlist, f0, numharmonics, V0, node, useprev=False):
        """The reference's own outer loop (:152-232): scipy.optimize.fmin on |Yosc| with its arguments; every
        objective evaluation is one warm-started GPU solve."""
        from scipy import optimize
        netlist = netlist.copy()
        self.freq = f0
        self.numharmonics = numharmonics
        Voscprobe = netlist.add_iac(self.name + '.I.Oscprobe', self.name + '.nosc1', 'gnd', ac=V0, freq=f0)
        netlist.add_gyrator(self.name + '.G.Oscprobe', self.name + '.nosc1', self.name + '.nosc2', 'gnd', 'gnd', 1)
        Zoscprobe = netlist.add_idealharmonicfilter(self.name + 'IHF.Oscprobe', self.name + '.nosc2', node, f0)
        self.osc_evals = []

        class SmallEnoughException(Exception):
            pass

        def objFunc(x, info):
            fosc = float(x[0])
            Vosc = float(x[1])
            Voscprobe.ac = Vosc
            Voscprobe.freq = fosc
            Zoscprobe.freq = fosc
            self.freq = fosc
            if info['itercnt'] > 0 or info['useprev']:
                converged, freqs, Vf, _, _ = self.run(netlist, self.V)
            else:
                converged, freqs, Vf, _, _ = self.run(netlist)
            if not converged:
                self.osc_evals.append((fosc, Vosc, 1e2))
                return 1e2
            n1 = self.get_node_idx(info['osc_node'])
            n2 = self.get_node_idx(self.name + '.nosc2')
            Voscx = Vf[n1, 1]
            Iosc = (Vf[n1, 1] - Vf[n2, 1]) * Zoscprobe.g
            Yosc = Iosc / Voscx
            info['itercnt'] += 1
            self.osc_evals.append((fosc, Vosc, abs(Yosc)))
            if self.verbose:
                print('\nIter\tFreq [Hz]\tVosc [V]\tmag(Yosc)')
                print('{}\t{:.8f}\t{:.8f}\t{:.2e}\n'.format(info['itercnt'], fosc, Vosc, abs(Yosc)))
            if abs(Yosc) < 1e-12:
                info['result'] = x
                raise SmallEnoughException()
            return abs(Yosc)

        args = ({'itercnt': 0, 'osc_node': node, 'result': 0, 'useprev': useprev},)
        try:
            xopt = optimize.fmin(func=objFunc, x0=[f0, V0], args=args, xtol=1e-12, ftol=1e-12, maxfun=300,
                                 disp=self.verbose, retall=False, full_output=True)[0]
        except SmallEnoughException:
            xopt = args[0]['result']
        fosc, Vosc = float(xopt[0]), float(xopt[1])
        Voscprobe.ac = Vosc
        Voscprobe.freq = fosc
        Zoscprobe.freq = fosc
        self.freq = fosc
        self.fosc, self.Vosc = fosc, Vosc
        out = self.run(netlist, self.V)
        print('Frequency of oscillation = {} Hz'.format(fosc))
        print('Oscillation amplitude = {} V'.format(Vosc))
        return out

    def run_oscillator_batch(self, netlist, f0, numharmonics, V0, node, sweeps=None, tol=1e-12, maxiter=30):
        """Batched oscillator analysis (start points, parameter sweeps: the reference's loops around run_oscillator,
        tests/HBOscillators.ipynb:3052-3115) as ONE GPU call: yhb_oscillator_host, Newton on (fosc, Vosc) per
        oscillator with the probe current from Kirchhoff's law at `node` and the exact 2 x 2 Jacobian.  f0 / V0:
        arrays of start points; sweeps as in run_sweep.  Returns an object with fosc[B], Vosc[B], converged[B],
        yosc[B] (|Yosc| reached), V, Vf, outer_iterations[B]."""
        from .. import _osc
        res = _osc.solve(self, netlist, f0, numharmonics, V0, node, sweeps=sweeps, tol=tol, maxiter=maxiter)
        self.netlist = res.netlist
        self.last_osc = res
        return res

    # ------------------------------------------------------------------ post-processing (:112-135)
    def print_v(self, node):
        n = self.get_node_idx(node)
        print('Voltage at node: ' + node)
        print('Freq [Hz]\tVmag [V]\tPhase [°]')
        for k in range(0, self.K + 1):
            v = self.Vf[n, k]
            print('{:.2e}\t{:.3e}\t{:.1f}'.format(self.freqs[k], np.abs(v), np.degrees(np.angle(v))))

    def get_v(self, node):
        n = self.get_node_idx(node)
        return self.Vf[n]

    def convert_to_time(self, xf):
        """:125-135 -- single-tone waveform of one node's phasors xf[K+1], 32 K samples over two periods; evaluated on
        the GPU (yhb_to_time).  Returns (t, xt) like the reference (whose t axis is linspace(0, S, S) * 2 / f / S)."""
        S = 32 * self.K
        t = 2 / self.freq / S * np.linspace(0, S, S)
        xf = np.asarray(xf, dtype=np.complex128).reshape(1, 1, -1)
        # sample s sits at phase 2 pi k s / (S / 2): bin 'frequencies' k, 'times' s / (S / 2)
        xt = E.to_time(xf, np.arange(self.K + 1, dtype=np.float64), np.arange(S) / (S / 2))[0, 0]
        return t, xt

    def to_time(self, nodes, t, Vf=None, freqs=None):
        """Quasi-periodic time-domain reconstruction (one or two tones) of node voltages at the times t, on the GPU:
        the loop of tests/HB_TwoTones.ipynb cell 3.  nodes: a node name or a list of names; Vf / freqs default to the
        last run's hb.Vf / hb.freqs.  Returns x[len(nodes)][len(t)] (or x[len(t)] for a single name)."""
        Vf = self.Vf if Vf is None else Vf
        freqs = self.freqs if freqs is None else freqs
        single = isinstance(nodes, str)
        idx = [self.get_node_idx(n) for n in ([nodes] if single else nodes)]
        x = E.to_time(np.asarray(Vf)[idx][None], freqs, t)[0]
        return x[0] if single else x

    def plot_v(self, node):
        """:35-110 -- needs matplotlib, which is an optional dependency here."""
        import matplotlib.pyplot as plt
        n = self.get_node_idx(node)
        plt.figure(figsize=(10, 6))
        plt.title('Frequency-domain $V(j \\omega)$')
        plt.stem(self.freqs, np.abs(self.Vf[n]), markerfmt='r^', linefmt='r')
        plt.xlabel('frequency [Hz]')
        plt.ylabel('V(\'' + node + '\') [V]')
        plt.grid()
        plt.tight_layout()
