// Batched DC operating point on the GPU: the cold-start guess of run() (calc_V0,
// yalrf/Analyses/MultiToneHarmonicBalance.py:697-704 -> DC.run, yalrf/Analyses/DC.py:54-123).
// One CTA per circuit runs the whole analysis: Newton with the devices' stateful tanh limiting
// (Diode.py:320-332, BJT.py:438-457, the check_vlimit side effects included), then the
// source-stepping (DC.py:236-279) and gmin-stepping (DC.py:193-234) fall-backs.
// Stamping is done by one thread in netlist order (it fixes the floating-point summation order);
// the linear solves use the CTA-wide elimination of yhb_kernels.cuh.
#pragma once
#include "yhb_kernels.cuh"

namespace yhb {

constexpr int kDcThreads = 128;  // upper bound; tiny circuits run one warp per circuit

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
};

constexpr int kDcOut = 12;  // per-device stamp values handed from the parallel evaluation to the ordered stamping

__host__ __device__ inline size_t dc_scratch_doubles(int M, int nd, int nb, int nc) {
    return (size_t)3 * M * M + (size_t)7 * M + 3 * nd + 2 * nb + (size_t)kDcOut * (nd + nb + nc) + 8;
}

// Gaussian elimination with partial pivoting on [A | rhs] by ONE warp, n <= 32: rows live on lanes during the
// elimination (each lane walks its own row), columns live on lanes during the row interchange.  Same pivot rule
// and arithmetic as block_ge_solve (dgetf2-style reciprocal scaling); only __syncwarp, no block barrier.
__device__ inline void warp_ge_solve(double *A, int lda, double *rhs, double *x, int n) {
    const int lane = threadIdx.x & 31;
    for (int k = 0; k < n; ++k) {
        const double v = (lane >= k && lane < n) ? fabs(A[(size_t)lane * lda + k]) : -1.;
        const unsigned long long bits = v < 0. ? 0ull : (unsigned long long)__double_as_longlong(v);
        const bool valid = lane >= k && lane < n;
        const unsigned hi = (unsigned)(bits >> 32), lo = (unsigned)bits;
        const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
        const unsigned mlo = __reduce_max_sync(0xffffffffu, hi == mhi ? lo : 0u);
        int p = __reduce_min_sync(0xffffffffu, (valid && hi == mhi && lo == mlo) ? lane : 0x7fffffff);
        if (p == 0x7fffffff) p = k;
        if (p != k) {
            for (int j = k + lane; j <= n; j += 32) {
                double *pk = (j < n) ? &A[(size_t)k * lda + j] : &rhs[k];
                double *pp = (j < n) ? &A[(size_t)p * lda + j] : &rhs[p];
                const double t = *pk;
                *pk = *pp;
                *pp = t;
            }
        }
        __syncwarp();
        const double rinv = 1. / A[(size_t)k * lda + k];
        if (lane > k && lane < n) {
            double *ai = A + (size_t)lane * lda;
            const double *ak = A + (size_t)k * lda;
            const double l = -(ai[k] * rinv);
            for (int j = k + 1; j < n; ++j) ai[j] = fma(l, ak[j], ai[j]);
            rhs[lane] = fma(l, rhs[k], rhs[lane]);
        }
        __syncwarp();
    }
    for (int j = n - 1; j >= 0; --j) {
        const double xj = rhs[j] / A[(size_t)j * lda + j];
        __syncwarp();
        if (lane < j) rhs[lane] = fma(-A[(size_t)lane * lda + j], xj, rhs[lane]);
        if (lane == 0) x[j] = xj;
        __syncwarp();
    }
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
        for (int i = tid; i < M * M; i += nt) c.Anl[i] = 0.;
        for (int i = tid; i < M; i += nt) c.znl[i] = 0.;
        __syncthreads();
        dc_eval_devices(c, c.xk);
        __syncthreads();
        if (tid == 0) dc_apply_stamps(c);
        __syncthreads();
        for (int i = tid; i < M * M; i += nt) {
            int r = i / M, cc = i - r * M;
            double a = c.A[i];
            if (extra_g != 0. && r == cc && r < N) a = a + extra_g;
            c.Aw[i] = a + c.Anl[i];
        }
        for (int i = tid; i < M; i += nt) c.zw[i] = zscale * c.z[i] + c.znl[i];
        __syncthreads();
        if (nt == 32) {
            warp_ge_solve(c.Aw, M, c.zw, c.x, M);
            __syncthreads();
        } else {
            block_ge_solve(c.Aw, M, c.zw, c.x, M, c.s_row, c.s_val, c.s_idx);
        }
        c.newton += 1;
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
    }
    return converged;
}

__global__ void __launch_bounds__(kDcThreads) k_dc(PlanDev P, DcArgs a, int batch) {
    extern __shared__ double s_row[];
    __shared__ double s_val[32];
    __shared__ int s_idx[33];
    __shared__ int s_flag;
    const int b = blockIdx.x;
    if (b >= batch) return;
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
    if (tid == 0) {  // linear stamps in netlist order, DC.py:79-87
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
                    if (n1 > 0) c.A[(size_t)(n1 - 1) * M + n1 - 1] += g;
                    if (n2 > 0) c.A[(size_t)(n2 - 1) * M + n2 - 1] += g;
                    if (n1 > 0 && n2 > 0) {
                        c.A[(size_t)(n1 - 1) * M + n2 - 1] -= g;
                        c.A[(size_t)(n2 - 1) * M + n1 - 1] -= g;
                    }
                } break;
                case YHB_LIN_INDUCTOR: {  // Inductor.py:21-26 (one branch row per inductor)
                    int i = branch++;
                    if (n1 > 0) { c.A[(size_t)(n1 - 1) * M + i] = +1.0; c.A[(size_t)i * M + n1 - 1] = +1.0; }
                    if (n2 > 0) { c.A[(size_t)(n2 - 1) * M + i] = -1.0; c.A[(size_t)i * M + n2 - 1] = -1.0; }
                } break;
                case YHB_LIN_GYRATOR: {  // Gyrator.py:22-30 (uses G at DC)
#define YHB_G(r, cc, s) do { if ((r) > 0 && (cc) > 0) c.A[(size_t)((r) - 1) * M + (cc) - 1] += (s) * v; } while (0)
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
            if (n1 > 0) c.z[n1 - 1] = c.z[n1 - 1] + sv[3 * s];
            if (n2 > 0) c.z[n2 - 1] = c.z[n2 - 1] - sv[3 * s];
        }
        for (int i = 0; i < N; ++i) c.A[(size_t)i * M + i] += 1e-12;  // DC.py:86-87
    }
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
}

}  // namespace yhb
