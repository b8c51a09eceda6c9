// Autonomous-oscillator solve on the device (run_oscillator, MultiToneHarmonicBalance.py:137-232, with scipy.optimize's
// Nelder-Mead replaced by a damped 2 x 2 Newton iteration on (fosc, Vosc)).
//
// The reference adds an oscillator probe to the netlist -- AC current source (amplitude Vosc) -> unit gyrator
// (= a voltage source at node nosc2) -> IdealHarmonicFilter (1e9 S at the fundamental only) -> probed node -- and
// minimises |Yosc|, Yosc = probe current / probe voltage at the fundamental (:183-196).  Its probe current
// (V(node) - V(nosc2)) * 1e9 S is quantised at one ulp of the phasors (1.5e-7 S); here the probe current is taken from
// Kirchhoff's law at the probed node (every other current into the node: linear branches from the converged
// phasors, the nonlinear part from the engine's Inl output), and the outer iteration solves Yosc(f, A) = 0 with the
// exact 2 x 2 Jacobian: k_fused returns x_p = J^-1 dF/dp for p = f, A from the already assembled exact Jacobian (two
// extra right-hand sides), so that dV/dp = -x_p and, by row `node` of the sensitivity system,
//   d I_other / dp = -g_filter (dV_node/dp - dV_nosc2/dp).
// One inner HB solve per outer iteration and oscillator; the whole batch (start points, parameter sweeps) advances
// together, finished oscillators are skipped by the solver.
#pragma once
#include "yhb_kernels.cuh"

namespace yhb {

struct OscState {
    double f, A;     // accepted point
    double Yr, Yi;   // Yosc at the accepted point
    double sf, sa;   // full Newton step from the accepted point
    double lam;      // damping of the trial being solved
    double ft, At;   // trial point
    int have, done, failed, iters;
    int hb_newton, pad;
    double J[4];     // 2 x 2 Jacobian d(Re Y, Im Y) / d(f, A) at the accepted point (row-major)
    double dbg[8];
};

struct OscArgs {
    OscState *st;
    double *tones, *lin_val, *src_val;  // the batch's parameter arrays (device): rewritten at the probe every iteration
    double *V, *Vacc;                   // [B][W] iterate of the inner solve / V of the accepted point
    const double *Inl, *sens;           // [B][W], [B][2][W]
    const Ctl *ctl;
    int *skip;
    int *live;                          // oscillators still iterating after this update
    int probe_src, probe_filter, node, nosc2;  // nodes in reference numbering
    double tol, max_df_rel, max_da_rel;
    int B;
};

__global__ void k_osc_begin(OscArgs a, const double *f0, const double *A0) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.B) return;
    OscState s;
    s.f = f0[b];
    s.A = A0[b];
    s.Yr = s.Yi = __longlong_as_double(0x7ff0000000000000ll);
    s.sf = s.sa = 0.;
    s.lam = 1.;
    s.ft = s.f;
    s.At = s.A;
    s.have = s.done = s.failed = s.iters = 0;
    s.hb_newton = 0;
    s.pad = 0;
    s.J[0] = s.J[1] = s.J[2] = s.J[3] = 0.;
    a.st[b] = s;
    a.skip[b] = 0;
}

// trial point -> probe parameters; warm start from the accepted point
__global__ void k_osc_set(PlanDev P, OscArgs a) {
    const int b = blockIdx.x;
    if (b >= a.B) return;
    OscState s = a.st[b];
    const bool idle = s.done || s.failed;
    const double ft = (!idle && s.have) ? s.f + s.lam * s.sf : s.f;
    const double At = (!idle && s.have) ? s.A + s.lam * s.sa : s.A;
    if (threadIdx.x == 0) {
        a.st[b].ft = ft;
        a.st[b].At = At;
        a.skip[b] = idle;
        a.tones[2 * b] = ft;                                               // self.freq = fosc        (:164)
        a.lin_val[((size_t)b * P.nlin + a.probe_filter) * 2] = ft;         // Zoscprobe.freq = fosc   (:163)
        a.src_val[((size_t)b * P.nsrc + a.probe_src) * 3 + 1] = At;        // Voscprobe.ac = Vosc     (:161), phase 0
        a.src_val[((size_t)b * P.nsrc + a.probe_src) * 3 + 2] = 0.;
    }
    if (!idle && s.have)
        for (int i = threadIdx.x; i < P.W; i += blockDim.x) a.V[(size_t)b * P.W + i] = a.Vacc[(size_t)b * P.W + i];
}

__device__ __forceinline__ void cmul(double ar, double ai, double br, double bi, double &cr, double &ci) {
    cr = ar * br - ai * bi;
    ci = ar * bi + ai * br;
}

// objective, Jacobian and step control of one oscillator after its inner solve
__global__ void k_osc_update(PlanDev P, OscArgs a) {
    __shared__ int s_accept;
    const int b = blockIdx.x;
    if (b >= a.B) return;
    const int S = P.S, W = P.W;
    const double *V = a.V + (size_t)b * W;
    if (threadIdx.x == 0) {
        OscState s = a.st[b];
        int accept = 0;
        if (!(s.done || s.failed)) {
            const Ctl c = a.ctl[b];
            s.hb_newton += c.newton;
            const double f = s.ft, A = s.At;
            const double *lv = a.lin_val + (size_t)b * P.nlin * 2;
            const double *sv = a.src_val + (size_t)b * P.nsrc * 3;
            const double *Inl = a.Inl + (size_t)b * W;
            const int n = a.node - 1;
            const double vnr = V[n * S + 1], vni = V[n * S + 2];
            // every current into the node but the probe's, at the fundamental: Y' V - Is + Inl  (:580-588)
            double ir = Inl[n * S + 1], ii = Inl[n * S + 2];
            for (int t = P.row_ptr[n]; t < P.row_ptr[n + 1]; ++t) {
                const int p = P.row_pairs[t], m = P.pair_m[p];
                double yr = 0., yi = 0.;
                for (int tt = P.plin_ptr[p]; tt < P.plin_ptr[p + 1]; ++tt) {
                    const int e = P.plin[tt].idx;
                    if (e == a.probe_filter) continue;
                    double re, im;
                    lin_admittance(P, lv, e, 1, fabs(f), re, im);
                    if (P.plin[tt].sgn > 0) { yr += re; yi += im; } else { yr -= re; yi -= im; }
                }
                double pr, pi;
                cmul(yr, yi, V[m * S + 1], V[m * S + 2], pr, pi);
                ir += pr;
                ii += pi;
            }
            for (int t = P.ns_ptr[n]; t < P.ns_ptr[n + 1]; ++t) {
                const int sidx = P.ns[t].idx;
                if (P.src_kind[sidx] != YHB_SRC_AC1) continue;
                const double sg = P.ns[t].sgn > 0 ? 1. : -1.;
                ir -= sg * sv[3 * sidx + 1];
                ii -= sg * sv[3 * sidx + 2];
            }
            // Yosc = Iosc / V(node), Iosc = g (V(node) - V(nosc2)) = -I_other  (:183-185)
            const double vv = vnr * vnr + vni * vni;
            const double yr0 = -(ir * vnr + ii * vni) / vv, yi0 = -(ii * vnr - ir * vni) / vv;
            // sensitivities: dV/df = -x_f / f, dV/dA = -x_a;  dI_other/dp = -g (dV_node/dp - dV_nosc2/dp)
            const double *xf = a.sens + (size_t)b * 2 * W, *xa = xf + W;
            const double g = lv[2 * a.probe_filter + 1];
            const int m2 = a.nosc2 - 1;
            double J[2][2];
            for (int q = 0; q < 2; ++q) {
                const double *x = q == 0 ? xf : xa;
                const double sc = q == 0 ? -1. / f : -1.;
                const double dvr = sc * x[n * S + 1], dvi = sc * x[n * S + 2];
                const double d2r = m2 >= 0 ? sc * x[m2 * S + 1] : 0., d2i = m2 >= 0 ? sc * x[m2 * S + 2] : 0.;
                const double dir = -g * (dvr - d2r), dii = -g * (dvi - d2i);
                // dY = -(dI V - I dV) / V^2
                double t1r, t1i, t2r, t2i, v2r, v2i;
                cmul(dir, dii, vnr, vni, t1r, t1i);
                cmul(ir, ii, dvr, dvi, t2r, t2i);
                cmul(vnr, vni, vnr, vni, v2r, v2i);
                const double nr = -(t1r - t2r), ni = -(t1i - t2i), dd = v2r * v2r + v2i * v2i;
                J[0][q] = (nr * v2r + ni * v2i) / dd;
                J[1][q] = (ni * v2r - nr * v2i) / dd;
            }
            const bool ok = c.converged > 0 && isfinite(yr0) && isfinite(yi0);
            const double ynew = hypot(yr0, yi0), yold = hypot(s.Yr, s.Yi);
            accept = ok && (!s.have || ynew < yold);
            if (!accept) {  // rejected trial: halve the step from the accepted point, or give up
                s.lam *= 0.5;
                if (!s.have || s.lam < 1.0 / 256) s.failed = 1;
            } else {
                s.f = f;
                s.A = A;
                s.Yr = yr0;
                s.Yi = yi0;
                s.have = 1;
                s.iters += 1;
                const double det = J[0][0] * J[1][1] - J[0][1] * J[1][0];
                double nf = -(J[1][1] * yr0 - J[0][1] * yi0) / det;
                double na = -(-J[1][0] * yr0 + J[0][0] * yi0) / det;
                if (!isfinite(nf)) nf = 0.;
                if (!isfinite(na)) na = 0.;
                // trust region
                nf = fmin(fmax(nf, -a.max_df_rel * fabs(f)), a.max_df_rel * fabs(f));
                na = fmin(fmax(na, -a.max_da_rel * fabs(A)), a.max_da_rel * fabs(A));
                s.sf = nf;
                s.sa = na;
                s.lam = 1.;
                s.J[0] = J[0][0]; s.J[1] = J[0][1]; s.J[2] = J[1][0]; s.J[3] = J[1][1];
                s.dbg[0] = xa[n * S + 1]; s.dbg[1] = xa[n * S + 2]; s.dbg[2] = xa[m2 * S + 1]; s.dbg[3] = xa[m2 * S + 2];
                s.dbg[4] = xf[n * S + 1]; s.dbg[5] = xf[n * S + 2]; s.dbg[6] = ir; s.dbg[7] = ii;
                const double ulpf = fabs(f) * 2.220446049250313e-16, ulpa = fabs(A) * 2.220446049250313e-16;
                const bool small = fabs(nf) <= 4 * ulpf && fabs(na) <= 4 * ulpa;
                if (ynew < a.tol || small) s.done = 1;
            }
            if (!(s.done || s.failed)) atomicAdd(a.live, 1);
            a.st[b] = s;
        }
        s_accept = accept;
    }
    __syncthreads();
    if (s_accept)
        for (int i = threadIdx.x; i < W; i += blockDim.x) a.Vacc[(size_t)b * W + i] = V[i];
}

struct OscOut {
    double *fosc, *Vosc, *yosc, *diag;
    int32_t *converged, *outer_iters, *hb_newton_iters;
};

// results; the accepted point's V becomes the iterate the final k_finish reads
__global__ void k_osc_finish(PlanDev P, OscArgs a, OscOut o) {
    const int b = blockIdx.x;
    if (b >= a.B) return;
    const OscState s = a.st[b];
    if (threadIdx.x == 0) {
        if (o.fosc) o.fosc[b] = s.f;
        if (o.Vosc) o.Vosc[b] = s.A;
        if (o.yosc) { o.yosc[2 * b] = s.Yr; o.yosc[2 * b + 1] = s.Yi; }
        if (o.converged) o.converged[b] = s.done && !s.failed;
        if (o.outer_iters) o.outer_iters[b] = s.iters;
        if (o.hb_newton_iters) o.hb_newton_iters[b] = s.hb_newton;
        if (o.diag) {
            double *q = o.diag + 16 * (size_t)b;
            q[0] = s.J[0]; q[1] = s.J[1]; q[2] = s.J[2]; q[3] = s.J[3];
            q[4] = s.sf; q[5] = s.sa; q[6] = s.lam; q[7] = s.failed;
            for (int i = 0; i < 8; ++i) q[8 + i] = s.dbg[i];
        }
        // leave the parameter arrays at the accepted point
        a.tones[2 * b] = s.f;
        a.lin_val[((size_t)b * P.nlin + a.probe_filter) * 2] = s.f;
        a.src_val[((size_t)b * P.nsrc + a.probe_src) * 3 + 1] = s.A;
    }
    if (s.have)
        for (int i = threadIdx.x; i < P.W; i += blockDim.x) a.V[(size_t)b * P.W + i] = a.Vacc[(size_t)b * P.W + i];
}

}  // namespace yhb
