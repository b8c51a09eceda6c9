// Batched DC operating point on the GPU: the cold-start guess of run() (calc_V0,
// yalrf/Analyses/MultiToneHarmonicBalance.py:697-704 -> DC.run, yalrf/Analyses/DC.py:54-123).
// One CTA per circuit runs the whole analysis: Newton with the devices' stateful tanh limiting
// (Diode.py:320-332, BJT.py:438-457, the check_vlimit side effects included), then the
// source-stepping (DC.py:236-279) and gmin-stepping (DC.py:193-234) fall-backs.
// Stamping is done by one thread in netlist order (it fixes the floating-point summation order);
// the linear solves use the CTA-wide elimination of yhb_kernels.cuh.
#pragma once
#include <type_traits>

#include "yhb_kernels.cuh"

namespace yhb {

#ifdef YHB_TIMING
#define YHB_T0_() tmark = clock64()
#define YHB_TACC_(slot) do { unsigned long long n__ = clock64(); atomicAdd(&g_timing[slot], n__ - tmark); tmark = n__; } while (0)
#else
#define YHB_T0_() do {} while (0)
#define YHB_TACC_(slot) do {} while (0)
#endif

constexpr int kDcThreads = 128;  // general engine (k_dc); small circuits run on the one-warp engine (k_dc_warp)

struct DcArgs {
    const double *lin_val, *src_val, *diode_par, *bjt_par, *cubic_par;
    const int *lin_nodes;  // [nlin][4]
    const int *nl_kind;    // [nnl] 0 = diode, 1 = bjt, 2 = cubic, in netlist order
    const int *nl_index;   // [nnl]
    int nnl;
    int M;                 // unknowns: N node voltages + one branch current per inductor
    double *scratch;       // [B][stride] (used when the per-circuit scratch does not fit shared memory)
    size_t stride;
    int use_smem;
    double *Vdc;           // [B][N]
    int *info;             // [B][2] converged, Newton iterations
    // one-warp engine (k_dc_warp): plan-time tables, see build_dc_tables() in yhb_api.cu
    const int *ent_tab;    // [NE + 1] CSR pointers over the NE = M*M + M entries of [A | z] (row-major A, then z),
                           // [NE] destination in the work matrix (row * LD + col) | diagonal-of-a-node flag << 16,
                           // [items] contributions in stamping order: ((t * kDcOut + slot) << 1) | negate
    int ent_ints;
    const int *lane_tab;   // [32][kDcLaneInts] role of every lane in the device evaluation
    // DC analysis beside the harmonic-balance kernel (yhb_solve_cold, YHB_DC_OVERLAP): CTA i takes circuit order[i] (the
    // order in which the HB kernel will ask for them) and publishes ready[b] = 1 once Vdc[b] is in memory
    const int *order;      // [B] or nullptr
    int *ready;            // [B] or nullptr
    int *nready;           // counter of solved circuits (with ready) or nullptr
    int *nstarted;         // counter of circuits whose CTA has begun (with ready; the gate of yhb_solve_cold) or nullptr
    // circuits with bit-identical DC inputs (an amplitude sweep repeats every bias point): leader[b] = first circuit of
    // b's group (k_dc_group); only leaders are solved, k_dc_scatter copies their results to the others
    const int *leader;     // [B] or nullptr
};

constexpr int kDcOut = 12;
constexpr int kDcLaneInts = 8;  // kind (-1 none), half, t, idx, n1, n2, n3, partner lane  // per-device stamp values handed from the parallel evaluation to the ordered stamping

__host__ __device__ inline size_t dc_scratch_doubles(int M, int nd, int nb, int nc) {
    return (size_t)3 * M * M + (size_t)7 * M + 3 * nd + 2 * nb + (size_t)kDcOut * (nd + nb + nc) + 8;
}

// linear stamps in netlist order (DC.py:79-87) into zeroed (A, z); one thread
__device__ inline void dc_stamp_linear(const PlanDev &P, const DcArgs &a, int b, double *A, double *z, int M) {
    const int N = P.N;
    const double *lv = a.lin_val + (size_t)b * P.nlin * 2;
    int branch = N;
    for (int e = 0; e < P.nlin; ++e) {
        const int *nd = a.lin_nodes + 4 * e;
        const int n1 = nd[0], n2 = nd[1], n3 = nd[2], n4 = nd[3];
        const double v = lv[2 * e];
        switch (P.lin_kind[e]) {
            case YHB_LIN_RESISTOR: {  // Resistor.py:30-35
                double g = 1.0 / v;
                double dummy[1];
                (void)dummy;
                if (n1 > 0) A[(size_t)(n1 - 1) * M + n1 - 1] += g;
                if (n2 > 0) A[(size_t)(n2 - 1) * M + n2 - 1] += g;
                if (n1 > 0 && n2 > 0) {
                    A[(size_t)(n1 - 1) * M + n2 - 1] -= g;
                    A[(size_t)(n2 - 1) * M + n1 - 1] -= g;
                }
            } break;
            case YHB_LIN_INDUCTOR: {  // Inductor.py:21-26 (one branch row per inductor)
                int i = branch++;
                if (n1 > 0) { A[(size_t)(n1 - 1) * M + i] = +1.0; A[(size_t)i * M + n1 - 1] = +1.0; }
                if (n2 > 0) { A[(size_t)(n2 - 1) * M + i] = -1.0; A[(size_t)i * M + n2 - 1] = -1.0; }
            } break;
            case YHB_LIN_GYRATOR: {  // Gyrator.py:22-30 (uses G at DC)
#define YHB_G(r, cc, s) do { if ((r) > 0 && (cc) > 0) A[(size_t)((r) - 1) * M + (cc) - 1] += (s) * v; } while (0)
                YHB_G(n1, n2, +1.);
                YHB_G(n1, n3, -1.);
                YHB_G(n2, n1, -1.);
                YHB_G(n2, n4, +1.);
                YHB_G(n3, n1, +1.);
                YHB_G(n3, n4, -1.);
                YHB_G(n4, n2, -1.);
                YHB_G(n4, n3, +1.);
#undef YHB_G
            } break;
            default:  // capacitor (Capacitor.py:30-31), ihf (IdealHarmonicFilter.py:23-24): open at DC
                break;
        }
    }
    const double *sv = a.src_val + (size_t)b * P.nsrc * 3;
    for (int s = 0; s < P.nsrc; ++s) {  // CurrentSource.py:23-26
        if (P.src_kind[s] != YHB_SRC_DC && P.src_kind[s] != YHB_SRC_DC_ONLY) continue;
        int n1 = P.src_nodes[2 * s], n2 = P.src_nodes[2 * s + 1];
        if (n1 > 0) z[n1 - 1] = z[n1 - 1] + sv[3 * s];
        if (n2 > 0) z[n2 - 1] = z[n2 - 1] - sv[3 * s];
    }
    for (int i = 0; i < N; ++i) A[(size_t)i * M + i] += 1e-12;  // DC.py:86-87
}

struct DcCtx {
    const PlanDev *P;
    const DcArgs *a;
    int b, M;
    double *A, *Anl, *Aw, *z, *zs, *znl, *zw, *xk, *x, *xprev, *x0;
    double *dst, *bst, *dout;
    double *s_row, *s_val;
    int *s_idx;
    int *s_flag;
    int newton;
};

__device__ __forceinline__ double dc_v(const double *x, int n) { return n > 0 ? x[n - 1] : 0.; }

// Diode.limit_diode_voltage, Diode.py:320-332
__device__ inline double dc_limit(double &old, double v, double lim) {
    v = old + lim * tanh((v - old) / lim);
    old = v;
    return v;
}

__device__ inline void dc_stamp2(double *A, double *z, int M, int n1, int n2, double g, double I) {
    if (n1 > 0) A[(size_t)(n1 - 1) * M + (n1 - 1)] += g;
    if (n2 > 0) A[(size_t)(n2 - 1) * M + (n2 - 1)] += g;
    if (n1 > 0 && n2 > 0) {
        A[(size_t)(n1 - 1) * M + (n2 - 1)] -= g;
        A[(size_t)(n2 - 1) * M + (n1 - 1)] -= g;
    }
    if (n1 > 0) z[n1 - 1] = z[n1 - 1] - I;
    if (n2 > 0) z[n2 - 1] = z[n2 - 1] + I;
}

// Device evaluation at xk, one thread per device (the limiter state of a device is private to it); the stamp
// values go to dout, and dc_apply_stamps adds them into (Anl, znl) in netlist order.
__device__ inline void dc_eval_devices(DcCtx &c, const double *xk) {
    const PlanDev &P = *c.P;
    const int b = c.b;
    for (int t = threadIdx.x; t < c.a->nnl; t += blockDim.x) {
        const int kind = c.a->nl_kind[t], idx = c.a->nl_index[t];
        double *o = c.dout + (size_t)kDcOut * t;
        if (kind == 0) {  // Diode.calc_dc + add_dc_stamps, Diode.py:98-116, :238-294
            const double *p = c.a->diode_par + ((size_t)b * P.nd + idx) * YHB_DIODE_NPAR;
            double *st = c.dst + 3 * idx;  // Vdold, It, gt
            const int n1 = P.diode_nodes[2 * idx], n2 = P.diode_nodes[2 * idx + 1];
            const double A_ = p[YHB_DIODE_AREA];
            const double Is = p[YHB_DIODE_IS] * A_, Isr = p[YHB_DIODE_ISR] * A_, Ikf = p[YHB_DIODE_IKF] * A_;
            const double Ibv = p[YHB_DIODE_IBV] * A_, Rs = p[YHB_DIODE_RS] / A_;
            const double N = p[YHB_DIODE_N], Nr = p[YHB_DIODE_NR], Bv = p[YHB_DIODE_BV];
            const double Vt = YHB_K_BOLTZ * p[YHB_DIODE_TEMP] / YHB_Q_ELEM;
            double Vtot = dc_v(xk, n1) - dc_v(xk, n2);
            double Id = st[1] + st[2] * Vtot;
            double Vd = Vtot - Id * Rs;
            Vd = dc_limit(st[0], Vd, 10. * N * Vt);
            double Idf = Is * (exp_lim(Vd / (N * Vt)) - 1.);
            double gdf = Is / (N * Vt) * exp_lim(Vd / (N * Vt));
            double Idr = Isr * (exp_lim(Vd / (Nr * Vt)) - 1.);
            double gdr = Isr / (Nr * Vt) * exp_lim(Vd / (Nr * Vt));
            if (isfinite(Ikf)) {
                Idf = Idf * sqrt(Ikf / (Ikf + Idf));
                gdf = gdf * (1. - 0.5 * Idf / (Ikf + Idf)) * sqrt(Ikf / (Ikf + Idf));
            }
            double Ibr = 0., gbr = 0.;
            if (isfinite(Bv)) {
                Ibr = Ibv * exp_lim(-Bv / Vt) * (1. - exp_lim(-Vd / Vt));
                gbr = Ibv / Vt * exp_lim(-(Bv + Vd) / Vt);
            }
            Id = Idf + Idr + Ibr + (Vd * 1e-12);
            double gd = gdf + gdr + gbr + (1e-12);
            st[1] = (Id - gd * Vd) / (1. + gd * Rs);
            st[2] = gd / (1. + gd * Rs);
            o[0] = st[2];
            o[1] = st[1];
        } else if (kind == 1) {  // BJT.calc_dc + add_dc_stamps, BJT.py:134-167, :319-416 (honours self.type)
            const double *p = c.a->bjt_par + ((size_t)b * P.nb + idx) * YHB_BJT_NPAR;
            double *st = c.bst + 2 * idx;  // Vbeold, Vbcold
            const int B = P.bjt_nodes[3 * idx], C = P.bjt_nodes[3 * idx + 1], E = P.bjt_nodes[3 * idx + 2];
            const double A_ = p[YHB_BJT_AREA], type = p[YHB_BJT_TYPE];
            const double Vt = YHB_K_BOLTZ * p[YHB_BJT_TEMP] / YHB_Q_ELEM;
            const double Is = p[YHB_BJT_IS] * A_, Ise = p[YHB_BJT_ISE] * A_, Isc = p[YHB_BJT_ISC] * A_;
            const double Ikf = p[YHB_BJT_IKF] * A_, Ikr = p[YHB_BJT_IKR] * A_;
            const double Nf = p[YHB_BJT_NF], Nr = p[YHB_BJT_NR], Ne = p[YHB_BJT_NE], Nc = p[YHB_BJT_NC];
            const double Vaf = p[YHB_BJT_VAF], Var = p[YHB_BJT_VAR], Bf = p[YHB_BJT_BF], Br = p[YHB_BJT_BR];
            const double gmin = 1e-12;
            double Vb = dc_v(xk, B), Vc = dc_v(xk, C), Ve = dc_v(xk, E);
            double Vbe = (Vb - Ve) * type, Vbc = (Vb - Vc) * type;
            Vbe = dc_limit(st[0], Vbe, 10. * Nf * Vt);
            Vbc = dc_limit(st[1], Vbc, 10. * Nr * Vt);
            double If = Is * (exp_lim(Vbe / (Nf * Vt)) - 1.);
            double Ibei = If / Bf;
            double Iben = Ise * (exp_lim(Vbe / (Ne * Vt)) - 1.);
            double Ibe = Ibei + Iben + gmin * Vbe;
            double gbei = Is / (Nf * Vt * Bf) * exp_lim(Vbe / (Nf * Vt));
            double gben = Ise / (Ne * Vt) * exp_lim(Vbe / (Ne * Vt));
            double gpi = gbei + gben + gmin;
            double Ir = Is * (exp_lim(Vbc / (Nr * Vt)) - 1.);
            double Ibci = Ir / Br;
            double Ibcn = Isc * (exp_lim(Vbc / (Nc * Vt)) - 1.);
            double Ibc = Ibci + Ibcn + gmin * Vbc;
            double gbci = Is / (Nr * Vt * Br) * exp_lim(Vbc / (Nr * Vt));
            double gbcn = Isc / (Nc * Vt) * exp_lim(Vbc / (Nc * Vt));
            double gmu = gbci + gbcn + gmin;
            double Q1 = 1. / (1. - (Vbc / Vaf) - (Vbe / Var));
            double Q2 = (If / Ikf) + (Ir / Ikr);
            double sq = sqrt(1. + 4. * Q2);
            double Qb = (Q1 / 2.) * (1. + sq);
            double It = (If - Ir) / Qb;
            double gif = gbei * Bf, gir = gbci * Br;
            double dQb_dVbe = Q1 * ((Qb / Var) + (gif / (Ikf * sq)));
            double dQb_dVbc = Q1 * ((Qb / Vaf) + (gir / (Ikr * sq)));
            double gmf = (1. / Qb) * (+gif - It * dQb_dVbe);
            double gmr = (1. / Qb) * (-gir - It * dQb_dVbc);
            double Ibeeq = Ibe - gpi * Vbe;
            double Ibceq = Ibc - gmu * Vbc;
            double Iceeq = It - gmf * Vbe - gmr * Vbc;
            o[0] = gmu + gpi;
            o[1] = -gmu;
            o[2] = -gpi;
            o[3] = -gmu + gmf + gmr;
            o[4] = gmu - gmr;
            o[5] = -gmf;
            o[6] = -gpi - gmf - gmr;
            o[7] = gmr;
            o[8] = gpi + gmf;
            o[9] = (-Ibeeq - Ibceq) * type;
            o[10] = (+Ibceq - Iceeq) * type;
            o[11] = (+Ibeeq + Iceeq) * type;
            (void)B; (void)C; (void)E;
        } else {  // CubicNonLinearity.add_dc_stamps, CubicNonLinearity.py:20-34 (g = 3 alpha V^2 here)
            const double alpha = c.a->cubic_par[(size_t)b * P.nc + idx];
            const int n1 = P.cubic_nodes[2 * idx], n2 = P.cubic_nodes[2 * idx + 1];
            double V = dc_v(xk, n1) - dc_v(xk, n2);
            double g = 3 * alpha * V * V;
            double I = alpha * V * V * V;
            double Ieq = I - g * V;
            o[0] = g;
            o[1] = Ieq;
        }
    }
}

// stamps of every device into (Anl, znl) in netlist order (fixes the summation order), thread 0 only
__device__ inline void dc_apply_stamps(DcCtx &c) {
    const PlanDev &P = *c.P;
    const int M = c.M;
    double *A = c.Anl, *z = c.znl;
    for (int t = 0; t < c.a->nnl; ++t) {
        const int kind = c.a->nl_kind[t], idx = c.a->nl_index[t];
        const double *o = c.dout + (size_t)kDcOut * t;
        if (kind == 0) {
            dc_stamp2(A, z, M, P.diode_nodes[2 * idx], P.diode_nodes[2 * idx + 1], o[0], o[1]);
        } else if (kind == 2) {
            dc_stamp2(A, z, M, P.cubic_nodes[2 * idx], P.cubic_nodes[2 * idx + 1], o[0], o[1]);
        } else {
            const int B = P.bjt_nodes[3 * idx], C = P.bjt_nodes[3 * idx + 1], E = P.bjt_nodes[3 * idx + 2];
#define YHB_A(r, cc, v) do { if ((r) > 0 && (cc) > 0) A[(size_t)((r) - 1) * M + ((cc) - 1)] += (v); } while (0)
            YHB_A(B, B, o[0]);
            YHB_A(B, C, o[1]);
            YHB_A(B, E, o[2]);
            YHB_A(C, B, o[3]);
            YHB_A(C, C, o[4]);
            YHB_A(C, E, o[5]);
            YHB_A(E, B, o[6]);
            YHB_A(E, C, o[7]);
            YHB_A(E, E, o[8]);
#undef YHB_A
            if (B > 0) z[B - 1] = z[B - 1] + o[9];
            if (C > 0) z[C - 1] = z[C - 1] + o[10];
            if (E > 0) z[E - 1] = z[E - 1] + o[11];
        }
    }
}

// Diode.check_vlimit (Diode.py:296-313) / BJT.check_vlimit (BJT.py:418-436) on the NEW solution; both
// advance the limiter state.  One thread per device; every device is visited (no early exit), as in DC.py:176-179.
// Returns this thread's partial verdict (combine with __syncthreads_and).
__device__ inline bool dc_check_vlimit(DcCtx &c, const double *x, double vabstol) {
    const PlanDev &P = *c.P;
    bool ok = true;
    for (int t = threadIdx.x; t < c.a->nnl; t += blockDim.x) {
        const int kind = c.a->nl_kind[t], idx = c.a->nl_index[t];
        if (kind == 0) {
            const double *p = c.a->diode_par + ((size_t)c.b * P.nd + idx) * YHB_DIODE_NPAR;
            double *st = c.dst + 3 * idx;
            const double Vt = YHB_K_BOLTZ * p[YHB_DIODE_TEMP] / YHB_Q_ELEM;
            const double Rs = p[YHB_DIODE_RS] / p[YHB_DIODE_AREA];
            double Vtot = dc_v(x, P.diode_nodes[2 * idx]) - dc_v(x, P.diode_nodes[2 * idx + 1]);
            double Id = st[1] + st[2] * Vtot;
            double Vd = Vtot - Id * Rs;
            double Vdlim = dc_limit(st[0], Vd, 10. * p[YHB_DIODE_N] * Vt);
            if (Vd - Vdlim > vabstol) ok = false;
        } else if (kind == 1) {
            const double *p = c.a->bjt_par + ((size_t)c.b * P.nb + idx) * YHB_BJT_NPAR;
            double *st = c.bst + 2 * idx;
            const double Vt = YHB_K_BOLTZ * p[YHB_BJT_TEMP] / YHB_Q_ELEM;
            const double type = p[YHB_BJT_TYPE];
            double Vb = dc_v(x, P.bjt_nodes[3 * idx]), Vc = dc_v(x, P.bjt_nodes[3 * idx + 1]);
            double Ve = dc_v(x, P.bjt_nodes[3 * idx + 2]);
            double Vbe = (Vb - Ve) * type, Vbc = (Vb - Vc) * type;
            double Vbelim = dc_limit(st[0], Vbe, 10. * p[YHB_BJT_NF] * Vt);
            double Vbclim = dc_limit(st[1], Vbc, 10. * p[YHB_BJT_NR] * Vt);
            if ((Vbe - Vbelim > vabstol) || (Vbc - Vbclim > vabstol)) ok = false;
        }
    }
    return ok;
}

// DC.solve_dc_nonlinear (DC.py:125-191) on (A + extra_g on node diagonals, zscale * z) from xstart.
// Result in c.x.  Uniform return value.
__device__ inline bool dc_newton(DcCtx &c, double extra_g, double zscale, const double *xstart) {
    const PlanDev &P = *c.P;
    const int M = c.M, N = P.N, tid = threadIdx.x, nt = blockDim.x;
    const double reltol = 1e-3, vabstol = 1e-6, iabstol = 1e-12;
    const int maxiter = 150;
    for (int i = tid; i < M; i += nt) c.xk[i] = xstart[i];
    __syncthreads();
    int k = 0;
    bool converged = false;
    while (!converged && k < maxiter) {
        YHB_T0();
        for (int i = tid; i < M * M; i += nt) c.Anl[i] = 0.;
        for (int i = tid; i < M; i += nt) c.znl[i] = 0.;
        __syncthreads();
        dc_eval_devices(c, c.xk);
        __syncthreads();
        if (tid == 0) YHB_TACC(24);
        if (tid == 0) dc_apply_stamps(c);
        __syncthreads();
        if (tid == 0) YHB_TACC(25);
        for (int i = tid; i < M * M; i += nt) {
            int r = i / M, cc = i - r * M;
            double a = c.A[i];
            if (extra_g != 0. && r == cc && r < N) a = a + extra_g;
            c.Aw[i] = a + c.Anl[i];
        }
        for (int i = tid; i < M; i += nt) c.zw[i] = zscale * c.z[i] + c.znl[i];
        __syncthreads();
        if (tid == 0) YHB_TACC(26);
        block_ge_solve(c.Aw, M, c.zw, c.x, M, c.s_row, c.s_val, c.s_idx);
        c.newton += 1;
        if (tid == 0) YHB_TACC(27);
        int nan = 0;
        for (int i = tid; i < M; i += nt) nan |= isnan(c.x[i]);
        if (__syncthreads_or(nan)) return false;  // Solver.py:14
        bool okc = true;
        for (int i = tid; i < M; i += nt)
            if (fabs(c.x[i] - c.xk[i]) > reltol * fabs(c.xk[i]) + (i < N ? vabstol : iabstol)) okc = false;
        bool vlim = dc_check_vlimit(c, c.x, vabstol);
        converged = __syncthreads_and(okc && vlim) != 0;
        if (!converged) {
            for (int i = tid; i < M; i += nt) c.xk[i] = c.x[i];
            k = k + 1;
        }
        __syncthreads();
        if (tid == 0) YHB_TACC(28);
    }
    return converged;
}

__global__ void __launch_bounds__(kDcThreads) k_dc(PlanDev P, DcArgs a, int batch) {
    extern __shared__ double s_row[];
    __shared__ double s_val[32];
    __shared__ int s_idx[33];
    __shared__ int s_flag;
    if ((int)blockIdx.x >= batch) return;
    const int b = a.order ? a.order[blockIdx.x] : blockIdx.x;
    if (a.leader && a.leader[b] != b) return;
    const int M = a.M, N = P.N, tid = threadIdx.x, nt = blockDim.x;
    DcCtx c;
    c.P = &P;
    c.a = &a;
    c.b = b;
    c.M = M;
    double *sc = a.use_smem ? s_row + (M + 2) : a.scratch + (size_t)b * a.stride;
    c.A = sc;
    c.Anl = c.A + (size_t)M * M;
    c.Aw = c.Anl + (size_t)M * M;
    c.z = c.Aw + (size_t)M * M;
    c.znl = c.z + M;
    c.zw = c.znl + M;
    c.xk = c.zw + M;
    c.x = c.xk + M;
    c.xprev = c.x + M;
    c.x0 = c.xprev + M;
    c.dst = c.x0 + M;
    c.bst = c.dst + 3 * P.nd;
    c.dout = c.bst + 2 * P.nb;
    c.s_row = s_row;
    c.s_val = s_val;
    c.s_idx = s_idx;
    c.s_flag = &s_flag;
    c.newton = 0;

    for (int i = tid; i < M * M; i += nt) c.A[i] = 0.;
    for (int i = tid; i < M; i += nt) { c.z[i] = 0.; c.x0[i] = 0.; c.x[i] = 0.; }
    for (int i = tid; i < 3 * P.nd + 2 * P.nb; i += nt) c.dst[i] = 0.;  // Diode.init / BJT.init
    __syncthreads();
    if (tid == 0) dc_stamp_linear(P, a, b, c.A, c.z, M);
    __syncthreads();

    bool solved = false;
    if (a.nnl == 0) {  // DC.py:90-95
        for (int i = tid; i < M * M; i += nt) c.Aw[i] = c.A[i];
        for (int i = tid; i < M; i += nt) c.zw[i] = c.z[i];
        __syncthreads();
        block_ge_solve(c.Aw, M, c.zw, c.x, M, s_row, s_val, s_idx);
        int nan = 0;
        for (int i = tid; i < M; i += nt) nan |= isnan(c.x[i]);
        solved = !__syncthreads_or(nan);
    }
    if (!solved) solved = dc_newton(c, 0., 1., c.x0);
    if (!solved) {  // source stepping, DC.py:236-279
        double alpha = 1e-3, delta = 1e-3;
        for (int i = tid; i < M; i += nt) { c.xprev[i] = c.x0[i]; c.x[i] = c.x0[i]; }
        __syncthreads();
        bool conv = false;
        int k = 0;
        while (!conv && k < 50) {
            for (int i = tid; i < M; i += nt) c.zw[i] = c.x[i];  // start vector (zw is rebuilt inside)
            __syncthreads();
            for (int i = tid; i < M; i += nt) c.x0[i] = c.zw[i];
            __syncthreads();
            bool ok = dc_newton(c, 0., alpha, c.x0);
            if (ok) {
                for (int i = tid; i < M; i += nt) c.xprev[i] = c.x[i];
                if (alpha >= 1.0) conv = true;
                else {
                    alpha = alpha + delta;
                    delta = 2 * delta;
                    if (alpha >= 1.0) alpha = 1.0;
                }
            } else {
                for (int i = tid; i < M; i += nt) c.x[i] = c.xprev[i];
                alpha = alpha - delta / 2;
                delta = delta / 2;
                if (alpha < 1e-3) break;
            }
            __syncthreads();
            k = k + 1;
        }
        solved = conv;
        if (!solved) {  // gmin stepping, DC.py:193-234
            double gmin = 0.01, rate = 0.05;
            for (int i = tid; i < M; i += nt) { c.xprev[i] = 0.; c.x[i] = 0.; }
            __syncthreads();
            conv = false;
            for (int guard = 0; guard < 10000 && !conv; ++guard) {
                for (int i = tid; i < M; i += nt) c.x0[i] = c.x[i];
                __syncthreads();
                bool ok = dc_newton(c, gmin, 1., c.x0);
                if (ok) {
                    for (int i = tid; i < M; i += nt) c.xprev[i] = c.x[i];
                    if (gmin < 1e-12) conv = true;
                    else {
                        gmin = gmin * rate;
                        rate = rate / 4;
                    }
                } else {
                    for (int i = tid; i < M; i += nt) c.x[i] = c.xprev[i];
                    gmin = gmin / rate;
                    rate = 2 * rate;
                    if (gmin > 0.01) break;
                }
                __syncthreads();
            }
            solved = conv;
        }
    }
    __syncthreads();
    for (int n = tid; n < N; n += nt) a.Vdc[(size_t)b * N + n] = c.x[n];
    if (tid == 0 && a.info) {
        a.info[2 * b] = solved ? 1 : 0;
        a.info[2 * b + 1] = c.newton;
    }
    if (a.ready) {
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicExch(a.ready + b, 1);
    }
}


// --------------------------------------------------------------------------------------------- //
// One-warp engine for small circuits (M <= MMAX <= 24 unknowns, at most 32 evaluation lanes): the same analysis
// with the same arithmetic in the same order, organised for latency instead of generality --
//  * all vectors (x, xk, xprev, x0) live one element per lane, the limiter states in the registers of the lane
//    that owns the junction;
//  * a BJT is evaluated by two lanes (B-E junction / B-C junction: limiter, exponentials, junction currents and
//    conductances), which exchange their results by shuffle, share Qb and It, and each finish one of gmf / gmr;
//  * the ordered stamping is done per matrix ENTRY: entry e of [A | z] adds its contributions in netlist order
//    (plan-time CSR), so the summation order of dc_apply_stamps is kept while the entries run in parallel;
//  * the linear solve keeps every row in the registers of one lane: pivot search by redux on the bit pattern
//    (ties: lowest current row position, the idamax rule), the reciprocal of every candidate computed while the
//    vote is in flight, pivot row broadcast by shuffle, implicit row interchange (a position per lane).
// --------------------------------------------------------------------------------------------- //
template <int I, int N, class F>
__device__ __forceinline__ void static_for(F &&f) {
    if constexpr (I < N) {
        f(std::integral_constant<int, I>{});
        static_for<I + 1, N>(f);
    }
}

__device__ __forceinline__ double shfl_d(double v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ double dcw_gather(double x, int n) {  // dc_v: node voltage, 0 for ground
    const double v = shfl_d(x, n > 0 ? n - 1 : 0);
    return n > 0 ? v : 0.;
}

template <int MMAX>
struct DcWarp {
    // role of this lane
    int kind, half, t, n1, n2, n3, partner;
    double st0, st1, st2;  // limiter state (junction voltage of the previous evaluation); diode: It, gt
    // hoisted constants.  BJT junction lane: lim, nvt1 (Nf|Nr Vt), nvt2 (Ne|Nc Vt), Is, Is2 (Ise|Isc), beta (Bf|Br),
    // g10 = Is / (N1 Vt beta), g20 = Is2 / (N2 Vt), Vx (Var|Vaf), Ikx (Ikf|Ikr), type; shared: Vaf, Var, Ikf, Ikr
    double lim, nvt1, nvt2, Is, Is2, beta, g10, g20, Vx, Ikx, type, Vaf, Var, Ikf, Ikr;
    // diode lane: lim, nvt1 = N Vt, nvt2 = Nr Vt, Is, Is2 = Isr, g10 = Is / (N Vt), g20 = Isr / (Nr Vt),
    // Ikx = Ikf, Vx = Bv, beta = Ibv, Vaf = Vt, Var = Rs;  cubic lane: Is = alpha
    int M, N, LD, nnl, lane;
    double *A, *z, *dout, *Aw;
    const int *ent_ptr;  // shared-memory copy of DcArgs::ent_tab
    int newton;
    unsigned long long tmark;

    __device__ __forceinline__ void eval(double xk) {
        const double Va = dcw_gather(xk, n1), Vb_ = dcw_gather(xk, n2), Vc_ = dcw_gather(xk, n3);
        double *o = dout + (size_t)kDcOut * t;
        const double gmin = 1e-12;
        // every shuffle below is executed by the whole warp, outside the role branches
        double Vj = 0., Ij = 0., Ibj = 0., gj = 0., gij = 0.;
        if (kind == 1) {  // BJT junction lane: n1 = B, n2 = C, n3 = E (BJT.py:319-416)
            Vj = (Va - (half ? Vb_ : Vc_)) * type;
            Vj = dc_limit(st0, Vj, lim);
            const double e1 = exp_lim(Vj / nvt1), e2 = exp_lim(Vj / nvt2);
            Ij = Is * (e1 - 1.);                       // If | Ir
            const double Ibi = Ij / beta;
            const double Ibn = Is2 * (e2 - 1.);
            Ibj = Ibi + Ibn + gmin * Vj;               // Ibe | Ibc
            const double gbi = g10 * e1;
            const double gbn = g20 * e2;
            gj = gbi + gbn + gmin;                     // gpi | gmu
            gij = gbi * beta;                          // gif | gir
        } else if (kind == 0) {  // Diode.calc_dc + add_dc_stamps, Diode.py:98-116, :238-294
            const double Vt = Vaf, Rs = Var, Ikf_ = Ikx, Bv = Vx, Ibv = beta;
            const double Vtot = Va - Vb_;
            double Id = st1 + st2 * Vtot;
            double Vd = Vtot - Id * Rs;
            Vd = dc_limit(st0, Vd, lim);
            const double ef = exp_lim(Vd / nvt1), er = exp_lim(Vd / nvt2);
            double Idf = Is * (ef - 1.);
            double gdf = g10 * ef;
            const double Idr = Is2 * (er - 1.);
            const double gdr = g20 * er;
            if (isfinite(Ikf_)) {
                Idf = Idf * sqrt(Ikf_ / (Ikf_ + Idf));
                gdf = gdf * (1. - 0.5 * Idf / (Ikf_ + Idf)) * sqrt(Ikf_ / (Ikf_ + Idf));
            }
            double Ibr = 0., gbr = 0.;
            if (isfinite(Bv)) {
                Ibr = Ibv * exp_lim(-Bv / Vt) * (1. - exp_lim(-Vd / Vt));
                gbr = Ibv / Vt * exp_lim(-(Bv + Vd) / Vt);
            }
            Id = Idf + Idr + Ibr + (Vd * 1e-12);
            const double gd = gdf + gdr + gbr + (1e-12);
            st1 = (Id - gd * Vd) / (1. + gd * Rs);
            st2 = gd / (1. + gd * Rs);
            o[0] = st2;
            o[1] = st1;
        } else if (kind == 2) {  // CubicNonLinearity.add_dc_stamps, CubicNonLinearity.py:20-34
            const double alpha = Is;
            const double V = Va - Vb_;
            const double g = 3 * alpha * V * V;
            const double I = alpha * V * V * V;
            o[0] = g;
            o[1] = I - g * V;
        }
        const double Vo = shfl_d(Vj, partner), Io = shfl_d(Ij, partner);
        double gm = 0., Ieq = 0., It = 0.;
        const double Vbe = half ? Vo : Vj, Vbc = half ? Vj : Vo;
        if (kind == 1) {
            const double If = half ? Io : Ij, Ir = half ? Ij : Io;
            const double Q1 = 1. / (1. - (Vbc / Vaf) - (Vbe / Var));
            const double Q2 = (If / Ikf) + (Ir / Ikr);
            const double sq = sqrt(1. + 4. * Q2);
            const double Qb = (Q1 / 2.) * (1. + sq);
            It = (If - Ir) / Qb;
            const double dQ = Q1 * ((Qb / Vx) + (gij / (Ikx * sq)));  // dQb_dVbe | dQb_dVbc
            gm = (1. / Qb) * ((half ? -gij : gij) - It * dQ);          // gmf | gmr
            Ieq = Ibj - gj * Vj;                                       // Ibeeq | Ibceq
        }
        const double go = shfl_d(gj, partner), gmo = shfl_d(gm, partner), Ieqo = shfl_d(Ieq, partner);
        if (kind == 1) {
            const double gpi = half ? go : gj, gmu = half ? gj : go;
            const double gmf = half ? gmo : gm, gmr = half ? gm : gmo;
            const double Ibeeq = half ? Ieqo : Ieq, Ibceq = half ? Ieq : Ieqo;
            const double Iceeq = It - gmf * Vbe - gmr * Vbc;
            if (!half) {
                o[0] = gmu + gpi;
                o[1] = -gmu;
                o[2] = -gpi;
                o[3] = -gmu + gmf + gmr;
                o[4] = gmu - gmr;
                o[5] = -gmf;
            } else {
                o[6] = -gpi - gmf - gmr;
                o[7] = gmr;
                o[8] = gpi + gmf;
                o[9] = (-Ibeeq - Ibceq) * type;
                o[10] = (+Ibceq - Iceeq) * type;
                o[11] = (+Ibeeq + Iceeq) * type;
            }
        }
    }

    // check_vlimit on the new solution (advances the limiter state); this lane's verdict
    __device__ __forceinline__ bool vlimit(double x, double vabstol) {
        const double Va = dcw_gather(x, n1), Vb_ = dcw_gather(x, n2), Vc_ = dcw_gather(x, n3);
        if (kind == 1) {
            const double Vj = (Va - (half ? Vb_ : Vc_)) * type;
            const double Vl = dc_limit(st0, Vj, lim);
            return !(Vj - Vl > vabstol);
        }
        if (kind == 0) {
            const double Vtot = Va - Vb_;
            const double Id = st1 + st2 * Vtot;
            const double Vd = Vtot - Id * Var;
            const double Vl = dc_limit(st0, Vd, lim);
            return !(Vd - Vl > vabstol);
        }
        return true;
    }

    // [A + extra_g | zscale z] + ordered nonlinear stamps -> Aw (shared), then the solve; returns x (element = lane)
    __device__ __forceinline__ double assemble_solve(double extra_g, double zscale) {
        const int NE = M * M + M;
        const int *ent_dst = ent_ptr + NE + 1, *ent_item = ent_dst + NE;
        for (int e = lane; e < NE; e += 32) {
            const int p0 = ent_ptr[e], p1 = ent_ptr[e + 1], dst = ent_dst[e];
            double base = A[e];  // z follows A
            if (e < M * M) {
                if (extra_g != 0. && (dst >> 16)) base = base + extra_g;
            } else {
                base = zscale * base;
            }
            double acc = 0.;
            for (int p = p0; p < p1; ++p) {
                const int it = ent_item[p];
                const double v = dout[it >> 1];
                acc = acc + ((it & 1) ? -v : v);
            }
            Aw[dst & 0xffff] = base + acc;
        }
        __syncwarp();
        if (lane == 0) YHB_TACC_(25);
        double a[MMAX], r = 0.;
#pragma unroll
        for (int j = 0; j < MMAX; ++j) a[j] = (lane < M && j < M) ? Aw[lane * LD + j] : 0.;
        if (lane < M) r = Aw[lane * LD + M];
        __syncwarp();
        int pos = lane;
        int piv[MMAX];
        static_for<0, MMAX>([&](auto kc) {
            constexpr int k = decltype(kc)::value;
            piv[k] = 0;
            if (k < M) {
                const bool cand = lane < M && pos >= k;
                // every candidate inverts its own entry while the vote is in flight (the division's fast path: no
                // branch to the special cases that the structural zeros of the matrix would take; zero pivot -> NaN)
                const double rc = rcp_fast(cand ? a[k] : 1.);
                const unsigned long long bits = cand ? (unsigned long long)__double_as_longlong(fabs(a[k])) : 0ull;
                const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
                const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
                const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
                int pp = __reduce_min_sync(0xffffffffu, (cand && hi == mhi && lo == mlo) ? pos : 0x7fffffff);
                if (pp == 0x7fffffff) pp = k;
                const int pl = __ffs(__ballot_sync(0xffffffffu, lane < M && pos == pp)) - 1;
                piv[k] = pl;
                if (pos == k) pos = pp;  // implicit interchange of positions k and pp
                if (lane == pl) pos = k;
                const double rinv = shfl_d(rc, pl);
                const double l = -(a[k] * rinv);
                const bool live = lane < M && pos > k;
#pragma unroll
                for (int j = k + 1; j < MMAX; ++j) {
                    const double pj = shfl_d(a[j], pl);
                    if (live) a[j] = fma(l, pj, a[j]);
                }
                const double pr = shfl_d(r, pl);
                if (live) r = fma(l, pr, r);
            }
        });
        double x = 0.;
        static_for<0, MMAX>([&](auto jc) {
            constexpr int j = MMAX - 1 - decltype(jc)::value;
            if (j < M) {
                const bool own = lane == piv[j];  // the other lanes divide 1 by 1: no special-case path
                const double xj = shfl_d((own ? r : 1.) / (own ? a[j] : 1.), piv[j]);
                if (lane < M && pos < j) r = fma(-a[j], xj, r);
                if (lane == j) x = xj;
            }
        });
        return x;
    }

    // DC.solve_dc_nonlinear (DC.py:125-191); x in / out per lane.  Uniform return value.
    __device__ __forceinline__ bool newton_loop(double extra_g, double zscale, double xstart, double &x) {
        const double reltol = 1e-3, vabstol = 1e-6, iabstol = 1e-12;
        const int maxiter = 150;
        double xk = xstart;
        int k = 0;
        bool converged = false;
        while (!converged && k < maxiter) {
            YHB_T0_();
            eval(xk);
            __syncwarp();
            if (lane == 0) YHB_TACC_(24);
            x = assemble_solve(extra_g, zscale);
            newton += 1;
            if (lane == 0) YHB_TACC_(27);
            if (__any_sync(0xffffffffu, lane < M && isnan(x))) return false;  // Solver.py:14
            bool okc = true;
            if (lane < M && fabs(x - xk) > reltol * fabs(xk) + (lane < N ? vabstol : iabstol)) okc = false;
            const bool vl = vlimit(x, vabstol);
            converged = __all_sync(0xffffffffu, okc && vl);
            if (!converged) {
                xk = x;
                k = k + 1;
            }
            if (lane == 0) YHB_TACC_(28);
        }
        return converged;
    }
};

template <int MMAX>
#ifndef YHB_DC_WARP_MINBLOCKS
#define YHB_DC_WARP_MINBLOCKS 16  // 128 registers: six of these warps fit beside two CTAs of the fused HB kernel
#endif
__global__ void __launch_bounds__(32, YHB_DC_WARP_MINBLOCKS) k_dc_warp(PlanDev P, DcArgs a, int batch) {
    extern __shared__ double smw[];
    const int M = a.M, N = P.N, lane = threadIdx.x;
    if ((int)blockIdx.x >= batch) return;
    const int b = a.order ? a.order[blockIdx.x] : blockIdx.x;
    if (a.leader && a.leader[b] != b) return;
    if (a.nstarted && lane == 0) atomicAdd(a.nstarted, 1);
    int *tab = reinterpret_cast<int *>(smw + (size_t)M * M + M + (size_t)kDcOut * max(a.nnl, 1) + (size_t)M * ((M + 1) | 1));
    for (int i = lane; i < a.ent_ints; i += 32) tab[i] = a.ent_tab[i];
    DcWarp<MMAX> w;
    w.M = M;
    w.N = N;
    w.LD = (M + 1) | 1;
    w.nnl = a.nnl;
    w.lane = lane;
    w.A = smw;
    w.z = w.A + M * M;
    w.dout = w.z + M;
    w.Aw = w.dout + (size_t)kDcOut * max(a.nnl, 1);
    w.ent_ptr = tab;
    w.newton = 0;
    // role of this lane; hoisted constants with the operations of calc_dc
    const int *lt = a.lane_tab + kDcLaneInts * lane;
    w.kind = lt[0]; w.half = lt[1]; w.t = lt[2];
    const int idx = lt[3];
    w.n1 = lt[4]; w.n2 = lt[5]; w.n3 = lt[6]; w.partner = lt[7];
    w.st0 = w.st1 = w.st2 = 0.;  // Diode.init / BJT.init
    w.lim = w.nvt1 = w.nvt2 = w.Is = w.Is2 = w.beta = w.g10 = w.g20 = w.Vx = w.Ikx = w.type = 1.;
    w.Vaf = w.Var = w.Ikf = w.Ikr = 1.;
    if (w.kind == 1) {
        const double *p = a.bjt_par + ((size_t)b * P.nb + idx) * YHB_BJT_NPAR;
        const double A_ = p[YHB_BJT_AREA];
        const double Vt = YHB_K_BOLTZ * p[YHB_BJT_TEMP] / YHB_Q_ELEM;
        const double N1 = w.half ? p[YHB_BJT_NR] : p[YHB_BJT_NF], N2 = w.half ? p[YHB_BJT_NC] : p[YHB_BJT_NE];
        w.type = p[YHB_BJT_TYPE];
        w.Is = p[YHB_BJT_IS] * A_;
        w.Is2 = (w.half ? p[YHB_BJT_ISC] : p[YHB_BJT_ISE]) * A_;
        w.beta = w.half ? p[YHB_BJT_BR] : p[YHB_BJT_BF];
        w.Ikf = p[YHB_BJT_IKF] * A_;
        w.Ikr = p[YHB_BJT_IKR] * A_;
        w.Vaf = p[YHB_BJT_VAF];
        w.Var = p[YHB_BJT_VAR];
        w.Vx = w.half ? w.Vaf : w.Var;
        w.Ikx = w.half ? w.Ikr : w.Ikf;
        w.lim = 10. * N1 * Vt;
        w.nvt1 = N1 * Vt;
        w.nvt2 = N2 * Vt;
        w.g10 = w.Is / (N1 * Vt * w.beta);
        w.g20 = w.Is2 / (N2 * Vt);
    } else if (w.kind == 0) {
        const double *p = a.diode_par + ((size_t)b * P.nd + idx) * YHB_DIODE_NPAR;
        const double A_ = p[YHB_DIODE_AREA];
        const double Vt = YHB_K_BOLTZ * p[YHB_DIODE_TEMP] / YHB_Q_ELEM;
        const double Nn = p[YHB_DIODE_N], Nr = p[YHB_DIODE_NR];
        w.Is = p[YHB_DIODE_IS] * A_;
        w.Is2 = p[YHB_DIODE_ISR] * A_;
        w.Ikx = p[YHB_DIODE_IKF] * A_;
        w.beta = p[YHB_DIODE_IBV] * A_;
        w.Var = p[YHB_DIODE_RS] / A_;
        w.Vx = p[YHB_DIODE_BV];
        w.Vaf = Vt;
        w.lim = 10. * Nn * Vt;
        w.nvt1 = Nn * Vt;
        w.nvt2 = Nr * Vt;
        w.g10 = w.Is / (Nn * Vt);
        w.g20 = w.Is2 / (Nr * Vt);
    } else if (w.kind == 2) {
        w.Is = a.cubic_par[(size_t)b * P.nc + idx];
    }

    for (int i = lane; i < M * M + M; i += 32) w.A[i] = 0.;  // A and z are contiguous
    __syncwarp();
    if (lane == 0) dc_stamp_linear(P, a, b, w.A, w.z, M);
    __syncwarp();

    double x = 0., x0 = 0., xprev = 0.;
    bool solved = false;
    if (a.nnl == 0) {  // DC.py:90-95
        x = w.assemble_solve(0., 1.);
        solved = !__any_sync(0xffffffffu, lane < M && isnan(x));
    }
    // DC.run: plain Newton, then source stepping (DC.py:236-279), then gmin stepping (DC.py:193-234); one call
    // site of the Newton loop, driven by a small state machine
    int phase = solved ? 3 : 0, k = 0;
    double alpha = 1e-3, delta = 1e-3, gmin = 0.01, rate = 0.05;
    while (phase < 3) {
        const double eg = phase == 2 ? gmin : 0.;
        const double zs = phase == 1 ? alpha : 1.;
        if (phase != 0) x0 = x;
        const bool ok = w.newton_loop(eg, zs, x0, x);
        if (phase == 0) {
            if (ok) { solved = true; phase = 3; }
            else { phase = 1; xprev = x0; x = x0; k = 0; }
        } else if (phase == 1) {
            bool stop = false;
            if (ok) {
                xprev = x;
                if (alpha >= 1.0) { solved = true; phase = 3; }
                else {
                    alpha = alpha + delta;
                    delta = 2 * delta;
                    if (alpha >= 1.0) alpha = 1.0;
                }
            } else {
                x = xprev;
                alpha = alpha - delta / 2;
                delta = delta / 2;
                if (alpha < 1e-3) stop = true;
            }
            k = k + 1;
            if (phase == 1 && (stop || k >= 50)) { phase = 2; xprev = 0.; x = 0.; k = 0; }
        } else {
            if (ok) {
                xprev = x;
                if (gmin < 1e-12) { solved = true; phase = 3; }
                else {
                    gmin = gmin * rate;
                    rate = rate / 4;
                }
            } else {
                x = xprev;
                gmin = gmin / rate;
                rate = 2 * rate;
                if (gmin > 0.01) phase = 3;
            }
            k = k + 1;
            if (k >= 10000) phase = 3;
        }
    }
    if (lane < N) a.Vdc[(size_t)b * N + lane] = x;
    if (lane == 0 && a.info) {
        a.info[2 * b] = solved ? 1 : 0;
        a.info[2 * b + 1] = w.newton;
    }
    if (a.ready) {
        __threadfence();
        __syncwarp();
        if (lane == 0) {
            atomicExch(a.ready + b, 1);
            if (a.nready) atomicAdd(a.nready, 1);
        }
    }
}

// ---- circuits with identical DC inputs ------------------------------------------------------------------------------
// The DC analysis reads the linear element values, the DC source values and the device parameters -- not the tones and
// not the AC amplitudes / phases -- so the points of an amplitude or frequency sweep that share a bias point are ONE DC
// problem.  word i of circuit b's DC inputs (a superset of what the DC kernels read), 0 <= i < dc_words():
__device__ __forceinline__ int dc_words(const PlanDev &P) {
    return P.nlin + P.nsrc + P.nd * YHB_DIODE_NPAR + P.nb * YHB_BJT_NPAR + P.nc;
}
__device__ __forceinline__ unsigned long long dc_word(const PlanDev &P, const DcArgs &a, int b, int i) {
    double v;
    if (i < P.nlin) v = a.lin_val[((size_t)b * P.nlin + i) * 2];
    else if ((i -= P.nlin) < P.nsrc) v = a.src_val[((size_t)b * P.nsrc + i) * 3];
    else if ((i -= P.nsrc) < P.nd * YHB_DIODE_NPAR) v = a.diode_par[(size_t)b * P.nd * YHB_DIODE_NPAR + i];
    else if ((i -= P.nd * YHB_DIODE_NPAR) < P.nb * YHB_BJT_NPAR) v = a.bjt_par[(size_t)b * P.nb * YHB_BJT_NPAR + i];
    else v = a.cubic_par[(size_t)b * P.nc + (i - P.nb * YHB_BJT_NPAR)];
    return (unsigned long long)__double_as_longlong(v);
}

// one warp per circuit: order-independent 64-bit hash of the DC inputs (sum of mixed (word, position) pairs)
__global__ void __launch_bounds__(128) k_dc_hash(PlanDev P, DcArgs a, int batch, unsigned long long *hash) {
    const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (b >= batch) return;
    const int nw = dc_words(P);
    unsigned long long h = 0;
    for (int i = lane; i < nw; i += 32) {
        unsigned long long x = dc_word(P, a, b, i) + 0x9e3779b97f4a7c15ull * (unsigned long long)(i + 1);
        x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull;
        x ^= x >> 27; x *= 0x94d049bb133111ebull;
        x ^= x >> 31;
        h += x;
    }
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xffffffffu, h, o);
    if (lane == 0) hash[b] = h;
}

// leader[b] = the first circuit of b's chunk (kDcGroupChunk circuits) whose DC inputs equal b's bit for bit; kDcGroupLanes
// lanes share the scan of one circuit's chunk (strided), the smallest match wins
constexpr int kDcGroupChunk = 8192;
constexpr int kDcGroupLanes = 16;
__global__ void __launch_bounds__(256) k_dc_group(PlanDev P, DcArgs a, int batch, const unsigned long long *hash, int *leader,
                                                  int *nlead) {
    const int g = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = g / kDcGroupLanes, sub = g % kDcGroupLanes;
    const bool live = b < batch;
    const int bb = live ? b : batch - 1;
    const int c0 = bb / kDcGroupChunk * kDcGroupChunk;
    const unsigned long long hb = hash[bb];
    const int nw = dc_words(P);
    int lead = bb;
    for (int j = c0 + sub; j < bb; j += kDcGroupLanes) {
        if (hash[j] != hb) continue;
        bool same = true;
        for (int i = 0; i < nw && same; ++i) same = dc_word(P, a, j, i) == dc_word(P, a, bb, i);
        if (same) { lead = j; break; }
    }
#pragma unroll
    for (int o = kDcGroupLanes / 2; o > 0; o >>= 1) lead = min(lead, __shfl_xor_sync(0xffffffffu, lead, o));
    if (live && sub == 0) {
        leader[b] = lead;
        if (nlead && lead == b) atomicAdd(nlead, 1);
    }
}

// Gate between the DC kernel (side stream) and the HB kernel of an overlapped cold solve: opens when every DC problem's
// CTA has begun, i.e. is resident or done -- HB CTAs, which wait for operating points, can then never keep a DC CTA off
// the device -- and the DC kernel's long tail (its slowest circuits) runs beside the HB kernel.  More DC problems than
// fit the device at once: the gate opens when the last wave has started.  Measured on the DiffAmp sweep (4096 circuits;
// 64 / 256 / 1024 / 4096 distinct bias points): 15.8 / 16.9 / 17.1 / 34.6 ms against 17.1 / 19.4 / 19.0 / 42.0 ms with
// the DC kernel in front of the HB kernel.
__global__ void k_dc_gate(const int *nlead, const int *nstarted) {
    const long long t0 = clock64();
    const int n = *reinterpret_cast<const volatile int *>(nlead);
    while (*reinterpret_cast<const volatile int *>(nstarted) < n) {
        __nanosleep(500);
        if (clock64() - t0 > 4000000000ll) break;  // seconds: the DC kernel never ran; the HB kernel has its own time-out
    }
}

// results of the leaders to the other circuits of their groups
__global__ void __launch_bounds__(256) k_dc_scatter(int N, int batch, const int *leader, double *Vdc, int *info) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    const int l = leader[b];
    if (l == b) return;
    for (int n = 0; n < N; ++n) Vdc[(size_t)b * N + n] = Vdc[(size_t)l * N + n];
    if (info) {
        info[2 * b] = info[2 * l];
        info[2 * b + 1] = info[2 * l + 1];
    }
}

}  // namespace yhb
