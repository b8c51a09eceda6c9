// Nonlinear device models evaluated per time sample on the GPU.
//
// Follows Diode.get_mthb_params (yalrf/Devices/Diode.py:395-427), BJT.get_hb_params
// (yalrf/Devices/BJT.py:462-582) and CubicNonLinearity.get_mthb_params
// (yalrf/Devices/CubicNonLinearity.py:36-41) operation for operation: this translation unit is
// compiled with -fmad=false so that every multiply/add rounds separately like the reference's
// Python scalar arithmetic; only exp/tanh/pow differ (by <= 2 ulp, CUDA libdevice vs libm).
// Sample-invariant sub-expressions are hoisted into the *Derived structs (same operations, same
// order, evaluated once per circuit and device instead of once per sample).
#pragma once
#include <math.h>

#include "../../include/yhb.h"

namespace yhb {

// scipy.constants.k / .e (CODATA exact), Diode.py:2, BJT.py:1
#define YHB_K_BOLTZ 1.380649e-23
#define YHB_Q_ELEM 1.602176634e-19

// Diode.py:434-435 / BJT.py:588-589
__device__ __forceinline__ double exp_lim(double x) {
    const double e200 = 7.225973768125749e+86;  // np.exp(200.)
    return (x < 200.) ? exp(x) : e200 + e200 * (x - 200.);
}

struct PnJunction {  // BJT.py:591-608 with sample-invariant factors hoisted
    double Cj, Vj, Mj, fcvj;
    double c2, den;         // Cj / (1-Fc)^Mj ; Vj (1-Fc)
    double q1;              // Cj Vj / (1-Mj)
    double one_m;           // 1 - Mj
    double X0, X1, X2, cjvj, fc, fc2;
};

__device__ inline void pn_derive(double Cj, double Vj, double Mj, double Fc, PnJunction &j) {
    j.Cj = Cj;
    j.Vj = Vj;
    j.Mj = Mj;
    j.fc = Fc;
    j.fcvj = Fc * Vj;
    j.c2 = Cj / pow((1. - Fc), Mj);
    j.den = Vj * (1. - Fc);
    j.q1 = Cj * Vj / (1. - Mj);
    j.one_m = 1. - Mj;
    j.X0 = (1. - pow((1. - Fc), (1. - Mj))) / (1. - Mj);
    j.X1 = (1. - Fc * (1. + Mj)) / pow((1. - Fc), (1. + Mj));
    j.X2 = Mj / (2. * pow((1. - Fc), (1. + Mj)));
    j.cjvj = Cj * Vj;
    j.fc2 = Fc * Fc;
}

// (C, Q) of one depletion junction.  Cj == 0 short-circuits to exact zeros (0 * finite).
__device__ inline void pn_eval(const PnJunction &j, double V, double &C, double &Q) {
    if (j.Cj == 0.) {
        C = 0.;
        Q = 0.;
        return;
    }
    if (V <= j.fcvj) {
        double base = 1. - V / j.Vj;
        double pm = pow(base, -j.Mj);  // base^(1-Mj) = base * base^(-Mj): one pow instead of two (<= 1 ulp apart)
        C = j.Cj * pm;
        Q = j.q1 * (1. - base * pm);
    } else {
        C = j.c2 * (1. + j.Mj * (V - j.fcvj) / j.den);
        double r = V / j.Vj;
        double X = j.X0 + j.X1 * (r - j.fc) + j.X2 * (r * r - j.fc2);
        Q = j.cjvj * X;
    }
}

struct DiodeDerived {
    double lim;         // 10 N Vt
    double nvt, nrvt;   // N Vt, Nr Vt
    double Is, Isr;     // raw (Diode.py:396-398)
    double gdf0, gdr0;  // Is/(N Vt), Isr/(Nr Vt)
    double Ikf, Bv, Ibv, vt;
    double ibr0, gbr0;  // Ibv*exp_lim(-Bv/Vt), Ibv/Vt
    int has_ikf, has_bv;
    PnJunction j;       // depletion charge (YHB_FLAG_DIODE_CHARGE; Cj0 scaled by Area like Diode.init does, Diode.py:95)
    double Tt, Cp;
};

__device__ inline void diode_derive(const double *p, DiodeDerived &d) {
    double Vt = YHB_K_BOLTZ * p[YHB_DIODE_TEMP] / YHB_Q_ELEM;
    double N = p[YHB_DIODE_N], Nr = p[YHB_DIODE_NR];
    d.vt = Vt;
    d.lim = 10. * N * Vt;
    d.nvt = N * Vt;
    d.nrvt = Nr * Vt;
    d.Is = p[YHB_DIODE_IS];
    d.Isr = p[YHB_DIODE_ISR];
    d.gdf0 = d.Is / (N * Vt);
    d.gdr0 = d.Isr / (Nr * Vt);
    d.Ikf = p[YHB_DIODE_IKF];
    d.Bv = p[YHB_DIODE_BV];
    d.Ibv = p[YHB_DIODE_IBV];
    d.has_ikf = isfinite(d.Ikf);
    d.has_bv = isfinite(d.Bv);
    d.ibr0 = d.has_bv ? d.Ibv * exp_lim(-d.Bv / Vt) : 0.;
    d.gbr0 = d.Ibv / Vt;
    pn_derive(p[YHB_DIODE_CJ0] * p[YHB_DIODE_AREA], p[YHB_DIODE_VJ], p[YHB_DIODE_M], p[YHB_DIODE_FC], d.j);
    d.Tt = p[YHB_DIODE_TT];
    d.Cp = p[YHB_DIODE_CP];
}

// returns (gd, Id); Vd / Vdold are the unlimited junction voltages of this and the previous iterate
__device__ inline void diode_eval(const DiodeDerived &d, double Vd, double Vdold, double &gd, double &Id, double *Vlim = nullptr) {
    Vd = Vdold + d.lim * tanh((Vd - Vdold) / d.lim);  // Diode.py:407
    if (Vlim) *Vlim = Vd;
    double ef = exp_lim(Vd / d.nvt);
    double er = exp_lim(Vd / d.nrvt);
    double Idf = d.Is * (ef - 1.);
    double Idr = d.Isr * (er - 1.);
    double gdf = d.gdf0 * ef;
    double gdr = d.gdr0 * er;
    if (d.has_ikf) {  // Diode.py:414-416 (gdf uses the already-scaled Idf, as written)
        Idf = Idf * sqrt(d.Ikf / (d.Ikf + Idf));
        gdf = gdf * (1. - 0.5 * Idf / (d.Ikf + Idf)) * sqrt(d.Ikf / (d.Ikf + Idf));
    }
    double Ibr = 0., gbr = 0.;
    if (d.has_bv) {  // Diode.py:420-422
        Ibr = d.ibr0 * (1. - exp_lim(-Vd / d.vt));
        gbr = d.gbr0 * exp_lim(-(d.Bv + Vd) / d.vt);
    }
    Id = Idf + Idr + Ibr;
    gd = gdf + gdr + gbr;
}

// CubicNonLinearity.py:36-41 (g = 2 alpha V^2 as written in the reference)
__device__ __forceinline__ void cubic_eval(double alpha, double V, double &g, double &I) {
    g = 2 * alpha * V * V;
    I = alpha * V * V * V;
}

// Extension (YHB_FLAG_DIODE_CHARGE): the diode's charge and capacitance as the reference's own operating-point code
// writes them (Diode.calc_oppoint, Diode.py:207-235): Cd = Cp + Tt gd + Cj, Qd = Cp Vd + Tt Id + Qj, with the depletion
// pair (Cj, Qj) of Diode.py:219-228 (the shape of BJT.py:591-608).
// Vd: the LIMITED junction voltage diode_eval worked with (Vd = Vdold + lim tanh(...)), gd / Id its results
__device__ inline void diode_charge_eval(const DiodeDerived &d, double Vd, double gd, double Id, double &Cd, double &Qd) {
    double Cj, Qj;
    pn_eval(d.j, Vd, Cj, Qj);
    Cd = d.Cp + d.Tt * gd + Cj;
    Qd = d.Cp * Vd + d.Tt * Id + Qj;
}

struct BjtDerived {
    double sgn;  // +1, or -1 for a PNP under YHB_FLAG_PNP_EXTENSION
    double limf, limr, nfvt, nrvt, nevt, ncvt;
    double Is, Ise, Isc, Bf, Br, Ikf, Ikr, Vaf, Var;
    double gbei0, gben0, gbci0, gbcn0;
    PnJunction je, jc;
    double Cjs, Vjs, Mjs, qs1;  // substrate (BJT.py:549-554)
    double Xcjc, Tf, Xtf, Itf, Tr, vtf144, tfx2, tfxv;
};

__device__ inline void bjt_derive(const double *p, int flags, BjtDerived &d) {
    double A = p[YHB_BJT_AREA];
    double Vt = YHB_K_BOLTZ * p[YHB_BJT_TEMP] / YHB_Q_ELEM;
    double Nf = p[YHB_BJT_NF], Nr = p[YHB_BJT_NR], Ne = p[YHB_BJT_NE], Nc = p[YHB_BJT_NC];
    d.sgn = ((flags & YHB_FLAG_PNP_EXTENSION) && p[YHB_BJT_TYPE] < 0.) ? -1. : 1.;
    d.Is = p[YHB_BJT_IS] * A;  // BJT.init, BJT.py:117-127
    d.Ise = p[YHB_BJT_ISE] * A;
    d.Isc = p[YHB_BJT_ISC] * A;
    d.Ikf = p[YHB_BJT_IKF] * A;
    d.Ikr = p[YHB_BJT_IKR] * A;
    d.Itf = p[YHB_BJT_ITF] * A;
    d.Bf = p[YHB_BJT_BF];
    d.Br = p[YHB_BJT_BR];
    d.Vaf = p[YHB_BJT_VAF];
    d.Var = p[YHB_BJT_VAR];
    d.limf = 10. * Nf * Vt;
    d.limr = 10. * Nr * Vt;
    d.nfvt = Nf * Vt;
    d.nrvt = Nr * Vt;
    d.nevt = Ne * Vt;
    d.ncvt = Nc * Vt;
    d.gbei0 = d.Is / (Nf * Vt * d.Bf);
    d.gben0 = d.Ise / (Ne * Vt);
    d.gbci0 = d.Is / (Nr * Vt * d.Br);
    d.gbcn0 = d.Isc / (Nc * Vt);
    double Fc = p[YHB_BJT_FC];
    pn_derive(p[YHB_BJT_CJE] * A, p[YHB_BJT_VJE], p[YHB_BJT_MJE], Fc, d.je);
    pn_derive(p[YHB_BJT_CJC] * A, p[YHB_BJT_VJC], p[YHB_BJT_MJC], Fc, d.jc);
    d.Cjs = p[YHB_BJT_CJS] * A;
    d.Vjs = p[YHB_BJT_VJS];
    d.Mjs = p[YHB_BJT_MJS];
    d.qs1 = d.Cjs * d.Vjs / (1 - d.Mjs);
    d.Xcjc = p[YHB_BJT_XCJC];
    d.Tf = p[YHB_BJT_TF];
    d.Xtf = p[YHB_BJT_XTF];
    d.Tr = p[YHB_BJT_TR];
    d.vtf144 = 1.44 * p[YHB_BJT_VTF];
    d.tfx2 = d.Tf * d.Xtf * 2;
    d.tfxv = d.Tf * d.Xtf / (1.44 * p[YHB_BJT_VTF]);
}

// out[14] in the reference's return order (BJT.py:582):
// Ib, Ic, Ie, Qbe, Qbc, Qsc, gmu, gpi, gmf, gmr, Cbc, Cbe, Cbebc, Csc
__device__ inline void bjt_eval(const BjtDerived &d, double Vb, double Vc, double Ve, double Vbold,
                                double Vcold, double Veold, double *out) {
    const double gmin = 1e-12, cmin = 1e-18;
    if (d.sgn < 0.) {  // PNP extension: mirror terminal voltages
        Vb = -Vb; Vc = -Vc; Ve = -Ve; Vbold = -Vbold; Vcold = -Vcold; Veold = -Veold;
    }
    double Vbe = Vb - Ve, Vbc = Vb - Vc, Vsc = 0. - Vc;  // Vs forced to 0, MultiToneHarmonicBalance.py:504
    double Vbeold = Vbold - Veold, Vbcold = Vbold - Vcold;
    Vbe = Vbeold + d.limf * tanh((Vbe - Vbeold) / d.limf);  // BJT.py:502-503
    Vbc = Vbcold + d.limr * tanh((Vbc - Vbcold) / d.limr);

    double ef = exp_lim(Vbe / d.nfvt);
    double ee = (d.Ise != 0.) ? exp_lim(Vbe / d.nevt) : 1.;  // Ise == 0: 0 * (e - 1) = 0 either way
    double er = exp_lim(Vbc / d.nrvt);
    double ec = (d.Isc != 0.) ? exp_lim(Vbc / d.ncvt) : 1.;

    double If = d.Is * (ef - 1.);
    double Ibei = If / d.Bf;
    double Iben = d.Ise * (ee - 1.);
    double Ibe = Ibei + Iben + gmin * Vbe;
    double Ir = d.Is * (er - 1.);
    double Ibci = Ir / d.Br;
    double Ibcn = d.Isc * (ec - 1.);
    double Ibc = Ibci + Ibcn + gmin * Vbc;

    double Q1 = 1. / (1. - (Vbc / d.Vaf) - (Vbe / d.Var));
    double Q2 = (If / d.Ikf) + (Ir / d.Ikr);
    double sq = sqrt(1. + 4. * Q2);
    double Qb = (Q1 / 2.) * (1. + sq);
    double It = (If - Ir) / Qb;

    double gbei = d.gbei0 * ef;
    double gben = d.gben0 * ee;
    double gpi = gbei + gben + gmin;
    double gbci = d.gbci0 * er;
    double gbcn = d.gbcn0 * ec;
    double gmu = gbci + gbcn + gmin;
    double gif = gbei * d.Bf;
    double gir = gbci * d.Br;

    double dQb_dVbe = Q1 * ((Qb / d.Var) + (gif / (d.Ikf * sq)));
    double dQb_dVbc = Q1 * ((Qb / d.Vaf) + (gir / (d.Ikr * sq)));
    double gmf = (1. / Qb) * (+gif - It * dQb_dVbe);
    double gmr = (1. / Qb) * (-gir - It * dQb_dVbc);

    double Cbedep, Qbe, Cbcdep, Qbc;
    pn_eval(d.je, Vbe, Cbedep, Qbe);
    pn_eval(d.jc, Vbc, Cbcdep, Qbc);

    double Cscdep = 0., Qsc = 0.;
    if (d.Cjs != 0.) {  // BJT.py:549-554
        if (Vsc <= 0) {
            double base = 1. - (Vsc / d.Vjs);
            Cscdep = d.Cjs * pow(base, -d.Mjs);
            Qsc = d.qs1 * (1 - pow(base, 1. - d.Mjs));
        } else {
            Cscdep = d.Cjs * (1. + d.Mjs * Vsc / d.Vjs);
            Qsc = d.Cjs * Vsc / (1 + (d.Mjs * Vsc) / (2. * d.Vjs));
        }
    }

    // transit time, BJT.py:556-558
    // Tf == 0 (no transit time): every use of etf below is multiplied by an exact zero, and exp_lim is finite and
    // positive, so the exponential is skipped without changing any result (0 * finite = 0, NaN from rat survives)
    double etf = (d.Tf != 0.) ? exp_lim(Vbc / d.vtf144) : 1.;
    double rat = If / (If + d.Itf);  // 0/0 = NaN at If = Itf = 0, as in the reference
    double Tff = d.Tf * (1. + d.Xtf * (rat * rat) * etf);
    double ipi = If + d.Itf;
    double dTff_dVbe = (d.tfx2 * gif * If * d.Itf / (ipi * ipi * ipi)) * etf;
    double dTff_dVbc = d.tfxv * (rat * rat) * etf;

    double Cbcidep = d.Xcjc * Cbcdep;
    double Cbcdiff = d.Tr * gir;
    double Cbediff = (1. / Qb) * (If * dTff_dVbe + Tff * (gif - (If / Qb) * dQb_dVbe));
    double Cbebc = (If / Qb) * (dTff_dVbc - (Tff / Qb) * dQb_dVbc) + cmin;

    double Ib = Ibe + Ibc;
    double Ic = It - Ibc;
    double Ie = Ib + Ic;

    double Cbc = Cbcidep + Cbcdiff + cmin;
    Qbc = Qbc + (Ir * d.Tr + cmin * Vbc);
    double Cbe = Cbedep + Cbediff + cmin;
    Qbe = Qbe + (If * Tff / Qb + cmin * Vbe);
    double Csc = Cscdep + cmin;
    Qsc = Qsc + cmin * Vsc;

    double s = d.sgn;  // currents and charges flip for the mirrored device; small-signal terms do not
    out[0] = s * Ib; out[1] = s * Ic; out[2] = s * Ie; out[3] = s * Qbe; out[4] = s * Qbc; out[5] = s * Qsc;
    out[6] = gmu; out[7] = gpi; out[8] = gmf; out[9] = gmr;
    out[10] = Cbc; out[11] = Cbe; out[12] = Cbebc; out[13] = Csc;
}

}  // namespace yhb
