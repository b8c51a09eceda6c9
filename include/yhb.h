/*
 * yhb.h -- C ABI of the B200-native harmonic-balance engine (libyhb.so).
 *
 * The reference (victorpreuss/YalRF) is pure Python and has no FFI: its boundary for this path is
 * the object surface MultiToneHarmonicBalance.run / hb_loop / run_oscillator
 * (yalrf/Analyses/MultiToneHarmonicBalance.py:137-631) fed by Netlist.add_* (yalrf/Netlist.py:31-694).
 * The entry points below are what a ctypes binding inside that class would call instead of its own
 * numpy loop; each one cites the reference code it replaces.  Plain pointers and sizes only.
 *
 * Conventions
 *   - All arrays are fp64 / int32, circuit-major ("[B][...]" = B circuits of one batch, contiguous).
 *   - Node indices follow the reference: 0 = 'gnd', 1..N = Netlist.add_node order (Netlist.py:703-725).
 *   - The unknown vector V is the reference's layout (MultiToneHarmonicBalance.py:351-352, 398-400):
 *     W = N*S reals, node-major, row n*S = DC, n*S+2k-1 = Re, n*S+2k = Im of harmonic k, S = 2K+1.
 *   - Device entry points take DEVICE pointers and a cudaStream_t (passed as void*); they never
 *     synchronise the device unless stated.  *_host entry points take HOST pointers and do their own
 *     copies.  Return value: 0 = ok, negative = error (text via yhb_last_error()).  Non-convergence of
 *     a circuit is data (converged[b] = 0), not an error -- as in the reference (:386-387).
 *   - There is no CPU fallback: every entry point that computes needs a CUDA device.
 */
#ifndef YHB_H
#define YHB_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YHB_VERSION 210

/* linear element kinds: the classes that implement add_mthb_stamps (MultiToneHarmonicBalance.py:686-695) */
enum {
    YHB_LIN_RESISTOR = 0,  /* Resistor.py:43-65;   val[0] = R                                  */
    YHB_LIN_CAPACITOR = 1, /* Capacitor.py:62-85;  val[0] = C   (1e-12 S at DC)                */
    YHB_LIN_INDUCTOR = 2,  /* Inductor.py:64-87;   val[0] = L   (1e6 S at DC, no branch row)   */
    YHB_LIN_GYRATOR = 3,   /* Gyrator.py:38-73;    val[0] = G   (HB stamps hard-code 1; DC uses G) */
    YHB_LIN_IHF = 4        /* IdealHarmonicFilter.py:33-57; val[0] = freq, val[1] = g          */
};

/* independent current sources, consumed by calc_Is (MultiToneHarmonicBalance.py:647-684) */
enum {
    YHB_SRC_DC = 0,  /* itype == 'dc': val = (dc, -, -)                                             */
    YHB_SRC_AC1 = 1, /* itype == 'ac', freq == f1: val = (-, ac*cos(phase), ac*sin(phase))          */
    YHB_SRC_AC2 = 2, /* itype == 'ac', freq == f2 (two-tone only)                                   */
    YHB_SRC_DC_ONLY = 3 /* itype == 'both' (Netlist.add_isource): stamped by the DC analysis
                           (CurrentSource.py:23-26), ignored by calc_Is (:652, :675)                  */
};

/* raw diode model parameters, Diode.options (Diode.py:6-36); get_mthb_params uses the raw values */
enum {
    YHB_DIODE_TEMP = 0, YHB_DIODE_IS, YHB_DIODE_N, YHB_DIODE_ISR, YHB_DIODE_NR, YHB_DIODE_IKF,
    YHB_DIODE_BV, YHB_DIODE_IBV, YHB_DIODE_RS, YHB_DIODE_AREA,
    YHB_DIODE_CJ0, YHB_DIODE_VJ, YHB_DIODE_M, YHB_DIODE_FC, YHB_DIODE_TT, YHB_DIODE_CP, YHB_DIODE_NPAR
};
/* CJ0 .. CP: junction / diffusion charge (Diode.py:207-235), used by the reference's AC and transient analyses only; in
 * harmonic balance they count with YHB_FLAG_DIODE_CHARGE */

/* raw BJT model parameters, BJT.options (BJT.py:5-52) + polarity (BJT.py:71) */
enum {
    YHB_BJT_TEMP = 0, YHB_BJT_IS, YHB_BJT_NF, YHB_BJT_NR, YHB_BJT_IKF, YHB_BJT_IKR, YHB_BJT_VAF,
    YHB_BJT_VAR, YHB_BJT_ISE, YHB_BJT_NE, YHB_BJT_ISC, YHB_BJT_NC, YHB_BJT_BF, YHB_BJT_BR,
    YHB_BJT_CJE, YHB_BJT_VJE, YHB_BJT_MJE, YHB_BJT_CJC, YHB_BJT_VJC, YHB_BJT_MJC, YHB_BJT_XCJC,
    YHB_BJT_CJS, YHB_BJT_VJS, YHB_BJT_MJS, YHB_BJT_FC, YHB_BJT_TF, YHB_BJT_XTF, YHB_BJT_VTF,
    YHB_BJT_ITF, YHB_BJT_TR, YHB_BJT_AREA, YHB_BJT_TYPE /* +1 npn, -1 pnp */, YHB_BJT_NPAR
};

#define YHB_FLAG_PNP_EXTENSION 1  /* honour BJT polarity in HB (the reference ignores it, BJT.py:495-503) */
#define YHB_FLAG_EXACT_JACOBIAN 2 /* re-zero dQdV every iteration (the reference accumulates it, :413/:438) */
#define YHB_FLAG_DIODE_CHARGE 512 /* extension: diodes carry charge in HB, Qd = Cp Vd + Tt Id + Qj(Vd), Cd = Cp + Tt gd + Cj(Vd)
                                     (the reference's MTHB stamps diode conductance and current only, :441-475) */

#define YHB_FLAG_NO_PEEL 4        /* factor the full W x W system (no structural peeling of ideal-source rows) */
#define YHB_FLAG_NO_BAND 16       /* staged engine: dense core storage even when the core is block banded over nodes */
#define YHB_FLAG_STAGED 8         /* force the staged engine (Jacobian in HBM) even when the circuit fits one SM */
#define YHB_FLAG_BLOCK_THOMAS 32  /* staged engine: solve a block-TRIdiagonal core block by block (S x S node blocks,
                                     pivoting inside the diagonal blocks) even when it is small; chosen automatically
                                     for large blocks (config 5: ladders at many harmonics), where the one-CTA banded
                                     elimination is out of reach */
#define YHB_FLAG_NO_BLOCK_THOMAS 64
#define YHB_FLAG_NO_BLOCK_SPARSE 128 /* fused engine: dense Gauss-Jordan of the core even when its node blocks are sparse */
#define YHB_FLAG_BLOCK_SPARSE 256    /* fused engine: block-sparse LU of the core whenever it applies (also for dense block structures) */

#define YHB_MAX_LEVELS 64 /* hb_loop calls recorded per solve: 1 direct attempt + <=50 continuation levels */

typedef struct yhb_plan yhb_plan; /* opaque: topology + device tables of one netlist shape.  The tables are immutable, but a
                                     plan also owns mutable scratch (the pinned counters of the launch loop, the staging
                                     buffer and stream of the *_host entry points): calls on ONE plan serialise on an
                                     internal mutex.  Use one plan per host thread / stream for concurrency.  A plan
                                     belongs to the CUDA device that was current when it was created. */

/* Topology of one netlist shape (what Netlist + run()'s setup :234-302 determine).  Host pointers. */
typedef struct {
    int32_t num_nodes; /* N: non-ground nodes */
    int32_t num_tones; /* 1: K = K1 (:247-251); 2: box truncation K = (2 K2 + 1) K1 + K2 (:252-273) */
    int32_t K1, K2;
    int32_t num_lin;
    const int32_t *lin_kind;  /* [num_lin]    */
    const int32_t *lin_nodes; /* [num_lin][4] */
    int32_t num_src;
    const int32_t *src_kind;  /* [num_src]    */
    const int32_t *src_nodes; /* [num_src][2] */
    int32_t num_diode;
    const int32_t *diode_nodes; /* [num_diode][2]  anode, cathode */
    int32_t num_bjt;
    const int32_t *bjt_nodes; /* [num_bjt][3]  B, C, E (substrate is forced to 0 V, :504) */
    int32_t num_cubic;
    const int32_t *cubic_nodes; /* [num_cubic][2] */
    /* optional netlist order of the nonlinear devices (get_nonlinear_devices, Netlist.py:791-797), which fixes
     * the floating-point summation order of shared stamps: kind 0 = diode, 1 = bjt, 2 = cubic, index into the
     * tables above; [num_diode + num_bjt + num_cubic] each.  NULL = diodes, cubics, bjts. */
    const int32_t *nl_kind;
    const int32_t *nl_index;
    const double *idft; /* optional [S][S] row-major table as built at :294-300; NULL = analytic */
    const double *dft;  /* optional [S][S] = inv(idft) as at :302; NULL = analytic inverse       */
    int32_t flags;
} yhb_plan_desc;

/* Per-circuit parameters, re-read at every run (users mutate device objects between runs). */
typedef struct {
    int32_t batch;
    const double *tones;     /* [B][2]  f1, f2 (f2 ignored for one tone) */
    const double *lin_val;   /* [B][num_lin][2]   */
    const double *src_val;   /* [B][num_src][3]   dc, ac_re, ac_im */
    const double *diode_par; /* [B][num_diode][YHB_DIODE_NPAR] */
    const double *bjt_par;   /* [B][num_bjt][YHB_BJT_NPAR]     */
    const double *cubic_par; /* [B][num_cubic]   alpha */
} yhb_params;

/* MultiToneHarmonicBalance.options (:18-21) */
typedef struct {
    double reltol;
    double abstol;
    int32_t maxiter;
    int32_t reserved;
} yhb_options;

/* Results of a solve.  Any pointer may be NULL to skip that output.
 * struct_size MUST be set to sizeof(yhb_result) by the caller: the library reads only that many bytes and treats
 * the fields beyond them as NULL, so a binding written against an older header keeps working (and a binding that
 * forgets the field fails loudly with "yhb_result.struct_size not set" instead of reading past the struct). */
typedef struct {
    size_t struct_size;
    double *V;             /* [B][W]  in: start vector when have_v0, out: hb.V (:393)                  */
    double *Vf;            /* [B][N][K+1][2]  complex phasors hb.Vf (:396-403)                         */
    int32_t *converged;    /* [B]                                                                     */
    int32_t *newton_iters; /* [B]  Newton iterations over all hb_loop calls                            */
    int32_t *lu_count;     /* [B]  LU factorisations                                                   */
    int32_t *level_iters;  /* [B][YHB_MAX_LEVELS]  'NR number of iterations' of each hb_loop call, 0-terminated */
    double *level_alpha;   /* [B][YHB_MAX_LEVELS]  'Alpha level' of each call (1.0 for the direct attempt) */
    double *err;           /* [B]  'HB total error' sum|F| of the last hb_loop call (:597)             */
    double *Inl;           /* [B][W]  nonlinear current spectrum fft(it) + Omega fft(qt) (:584) at the returned V   */
    double *Ic;            /* [B][num_bjt][S]  collector-current samples of the last evaluation: the reference's
                              dev.Ic side channel (:512), read by tests/ColpittsCB_Analysis.py:96-101               */
} yhb_result;

const char *yhb_last_error(void);
int yhb_version(void);
/* CUDA device used by the calling thread for all later calls (cudaSetDevice) */
int yhb_set_device(int32_t device);

/* Per-kernel device timing (CUDA events on the launching stream around every launch of the solve loop).
 * yhb_profile_enable(1) switches it on for the calling process; yhb_profile_read resolves the recorded events
 * (synchronises them) and returns, per kernel kind, launches and total milliseconds since the last reset.
 * Kinds: 0 prepare, 1 dc, 2 init, 3 eval, 4 assemble, 5 lu, 6 finish, 7 fused. */
#define YHB_PROF_KINDS 8
int yhb_profile_enable(int32_t on);
int yhb_profile_reset(void);
int yhb_profile_read(int64_t *launches /* [YHB_PROF_KINDS] */, double *ms /* [YHB_PROF_KINDS] */);

/* Measures the FP64 FMA throughput of the current device (independent DFMA chains on every SM) and returns it
 * in TFLOP/s: the roofline denominator of the LU stage (MEASURED_PEAKS.json has no fp64 figure). */
int yhb_fp64_peak(double *tflops, void *stream);

/* Debug builds only (csrc/build.sh -DYHB_TIMING): clock64() cycles of the fused engine's phases summed over all
 * CTAs since the last reset.  Slots: 0 evaluation, 1 peeled rows / right-hand side, 2 core assembly, 3 core solve,
 * 4 back-substitution of the peeled rows + update, 5 ladder bookkeeping; 8 panel factorisation (owner warp),
 * 10-12 barrier wait (owner / next owner / other warps), 13-14 trailing update (next owner / other warps).
 * The production build returns -1. */
int yhb_debug_timing(unsigned long long *out64 /* [64] */, int32_t reset);

/* plan = the setup part of run() (:234-302): grid sizes, DFT/IDFT tables, stamp tables */
int yhb_plan_create(const yhb_plan_desc *desc, yhb_plan **out);
void yhb_plan_destroy(yhb_plan *plan);
int yhb_plan_dims(const yhb_plan *plan, int32_t *N, int32_t *K, int32_t *S, int32_t *W);
/* size of the dense core left after peeling ideal-source rows / columns (Wc = Nc * S) and whether the fused
 * on-chip engine will be used for this plan */
int yhb_plan_core(const yhb_plan *plan, int32_t *Nc, int32_t *Wc, int32_t *fused);
/* half bandwidth kl = ku (in matrix elements) of the band storage the staged engine uses for a block-banded core
 * (chains of nodes, BASELINE config 5), 0 = dense core */
int yhb_plan_band(const yhb_plan *plan);
/* which engine yhb_solve runs for this plan: 0 fused (on chip), 1 staged dense, 2 staged banded, 3 staged block-Thomas */
int yhb_plan_engine(const yhb_plan *plan);
/* fused engine: number of levels of the block-sparse LU over node blocks the plan will run (the nodes of a level are
 * factored concurrently), 0 = dense Gauss-Jordan of the whole core; *doubles (optional) = size of the block-row
 * storage in doubles */
int yhb_plan_block_sparse(const yhb_plan *plan, int32_t *doubles);
/* floating-point operations of ONE Newton-step linear solve: *lu_reference = the W x W dgetrf + dgetrs the reference
 * runs (:620-621, SURVEY 8d: 2/3 W^3 + 2 W^2), *lu_core_dense = the same for the Wc x Wc core left after peeling,
 * *lu_executed = what the engine chosen for this plan really executes (block-sparse LU over node blocks, tensor-core
 * Gauss-Jordan, banded or block-Thomas elimination).  Any pointer may be NULL. */
int yhb_plan_flops(const yhb_plan *plan, double *lu_reference, double *lu_core_dense, double *lu_executed);
/* real frequency of each bin of circuit tones (f1, f2), as self.freqs (:251, :271): host helper */
int yhb_plan_freqs(const yhb_plan *plan, double f1, double f2, double *freqs /* [K+1] */);

/* bytes of device scratch the solve entry points need for a batch of this size */
size_t yhb_workspace_bytes(const yhb_plan *plan, int32_t batch);

/* calc_V0 (:697-704) = DC.run (DC.py:54-123): batched DC operating points, Vdc[B][N] node voltages.
 * dc_info[B][2] = (converged, Newton iterations).  Circuits with bit-identical DC inputs are solved once (see
 * yhb_solve_cold); every circuit of such a group reports the group's results. */
int yhb_dc_solve(const yhb_plan *plan, const yhb_params *p, void *ws, size_t ws_bytes, double *Vdc,
                 int32_t *dc_info, void *stream);

/* run() (:234-405): optional direct attempt from V0, then the source-stepping ladder (:338-384) from
 * the DC point; every hb_loop (:407-631) runs on the GPU.  have_v0 != 0: res->V holds the start.
 * Vdc[B][N]: DC node voltages (from yhb_dc_solve), used for cold starts and after a failed direct attempt; with
 * have_v0 it may be NULL, and a circuit whose direct attempt fails then reports converged = -1 ("needs the DC
 * point": solve again from the same V0 with Vdc). */
int yhb_solve(const yhb_plan *plan, const yhb_params *p, const yhb_options *opt, void *ws,
              size_t ws_bytes, int32_t have_v0, const double *Vdc, const yhb_result *res, void *stream);

/* Cold run() as ONE call: calc_V0 (:321-324 -> DC.run) and the source-stepping ladder (:338-384).  Circuits of the batch
 * whose DC inputs are bit-identical (linear element values, DC source values, device parameters: the points of an
 * amplitude or frequency sweep that share a bias point) share ONE DC analysis.  On the fused (on-chip) engine the DC
 * kernel -- one warp per distinct DC problem, as long as its slowest circuit -- runs BESIDE the harmonic-balance kernel:
 * on a high-priority side stream of the plan, forked from and joined into `stream` by events; a small gate kernel on
 * `stream` holds the HB kernel back until every DC problem's CTA has begun (so that waiting HB CTAs can never keep DC
 * work off the device); the HB kernel takes the circuits strongest drive first (its last wave of circuits is the
 * cheapest), passes over circuits whose operating point has not arrived yet and serves them last.  The staged engines
 * run DC and HB one after the other.  Vdc[B][N] / dc_info[B][2] (optional outputs) as in yhb_dc_solve.  A circuit whose
 * operating point never arrived (the DC kernel did not run within seconds) reports converged = -2.
 * The staged engines' solve loops synchronise `stream` once per Newton round (they size the next launches from the
 * number of active circuits); the fused engine does not synchronise. */
int yhb_solve_cold(const yhb_plan *plan, const yhb_params *p, const yhb_options *opt, void *ws, size_t ws_bytes,
                   double *Vdc, int32_t *dc_info, const yhb_result *res, void *stream);

/* one hb_loop (:407-631) at source-stepping factor alpha[B] from res->V */
int yhb_newton(const yhb_plan *plan, const yhb_params *p, const yhb_options *opt, void *ws,
               size_t ws_bytes, const double *alpha, const yhb_result *res, void *stream);

/* Whole run() for a batch with HOST buffers: H2D of the parameters, DC points, solve, D2H of the results,
 * synchronous.  V0 (host, [B][W]) may be NULL for cold starts.  This is the end-to-end entry. */
int yhb_run_host(const yhb_plan *plan, const yhb_params *host_params, const yhb_options *opt,
                 const double *V0, const yhb_result *host_res);

/* ---- autonomous oscillators: run_oscillator (:137-232) ----------------------------------------------------------
 * The netlist handed to yhb_plan_create already contains the reference's oscillator probe (:146-150): AC current
 * source (amplitude = Vosc) -> unit gyrator -> IdealHarmonicFilter -> probed node.  The GPU owns the outer iteration
 * that replaces scipy.optimize.fmin (:204-212): a damped 2 x 2 Newton iteration on (fosc, Vosc) solving
 * Yosc = Iosc / V(node) = 0 (:183-196), the probe current taken from Kirchhoff's law at the node and the Jacobian from
 * two extra right-hand sides of the already factored HB Jacobian; one inner HB solve per outer iteration, the whole
 * batch (start points, parameter sweeps) advances together.  Fused (on-chip) engine, one tone. */
typedef struct {
    size_t struct_size;    /* sizeof(yhb_osc_options) */
    int32_t probe_src;     /* index (source table) of the probe current source: its ac amplitude is Vosc          */
    int32_t probe_filter;  /* index (linear table) of the probe IdealHarmonicFilter: its freq follows fosc        */
    int32_t node;          /* probed node, reference numbering                                                    */
    int32_t max_outer;     /* outer iterations, 0 = 30                                                            */
    double tol;            /* |Yosc| target in S, 0 = 1e-12 (the reference's early exit, :191)                    */
    double max_df_rel;     /* trust region of one outer step, relative: 0 = 0.02 in frequency ...                 */
    double max_dv_rel;     /* ... 0 = 0.3 in amplitude                                                            */
} yhb_osc_options;

typedef struct {
    size_t struct_size;       /* sizeof(yhb_osc_result) */
    double *fosc;             /* [B]  in: start frequency f0, out: oscillation frequency (also hb.freq, :226)      */
    double *Vosc;             /* [B]  in: start amplitude V0, out: oscillation amplitude                          */
    double *yosc;             /* [B][2]  Re, Im of Yosc at the returned point (may be NULL)                       */
    int32_t *converged;       /* [B]  (may be NULL)                                                               */
    int32_t *outer_iters;     /* [B]  accepted outer iterations (may be NULL)                                     */
    int32_t *hb_newton_iters; /* [B]  Newton iterations of all inner HB solves (may be NULL)                      */
    double *diag;             /* [B][8]  diagnostics at the returned point (may be NULL): the 2 x 2 Jacobian
                                 d(Re Yosc, Im Yosc) / d(f, V) row-major, the next Newton step (df, dV), its damping,
                                 1.0 if the outer iteration gave up                                               */
} yhb_osc_result;

/* Device pointers.  p's tones / lin_val (probe filter) / src_val (probe source) are REWRITTEN every outer iteration and
 * hold the returned point afterwards.  res = the HB result at the returned point (as the final self.run of :226-227).
 * Synchronises the stream once per outer iteration (it reads back how many oscillators are still iterating). */
int yhb_solve_oscillator(const yhb_plan *plan, const yhb_params *p, const yhb_options *opt,
                         const yhb_osc_options *osc, void *ws, size_t ws_bytes, const yhb_result *res,
                         const yhb_osc_result *osc_res, void *stream);
/* the same with HOST buffers (H2D / D2H inside, synchronous); the host parameter arrays are not modified */
int yhb_oscillator_host(const yhb_plan *plan, const yhb_params *host_params, const yhb_options *opt,
                        const yhb_osc_options *osc, const yhb_result *host_res, const yhb_osc_result *host_osc_res);

/* ---- post-processing: convert_to_time (:125-135) and the quasi-periodic two-tone reconstruction the reference's
 * notebooks do in a Python loop (tests/HB_TwoTones.ipynb cell 3).  Device pointers.
 *   x[b][n][s] = Re Vf[b][n][0] + sum_{k >= 1} ( Re Vf[b][n][k] cos(2 pi f[b][k] t[s]) - Im Vf[b][n][k] sin(2 pi f[b][k] t[s]) )
 * Vf[B][N][K1][2] phasors (hb.Vf), freqs[B][K1] bin frequencies (hb.freqs), t[nt] sample times, x[B][N][nt]. */
int yhb_to_time(int32_t batch, int32_t N, int32_t K1, const double *Vf, const double *freqs, int32_t nt,
                const double *t, double *x, void *stream);

/* ---- stage entry points (parity tests / profiling); device pointers -------------------------------- */
/* ifft (:706-713): vt[B][W] = IDFT applied per node */
int yhb_stage_ifft(const yhb_plan *plan, int32_t batch, const double *V, double *vt, void *stream);
/* device sweep (:441-570) on time samples: per-sample model outputs.
 * diode_out[B][num_diode][2][S] = (gd, Id); cubic_out[B][num_cubic][2][S] = (g, I);
 * bjt_out[B][num_bjt][14][S] in get_hb_params order (BJT.py:582). vtprev = previous iterate's samples. */
int yhb_stage_device_eval(const yhb_plan *plan, const yhb_params *p, const double *vt,
                          const double *vtprev, double *diode_out, double *bjt_out, double *cubic_out,
                          void *stream);
/* residual + Jacobian (:572-588) for iterate V with limiter reference vtprev and source factor alpha:
 * J[B][W][W] row-major, F[B][W].  cacc (may be NULL) = accumulated capacitance waveforms
 * [B][num_bjt][4][S] in/out (the reference's never-reset dQdV, :413). */
int yhb_stage_assemble(const yhb_plan *plan, const yhb_params *p, void *ws, size_t ws_bytes,
                       const double *V, const double *vtprev, const double *alpha, double *cacc,
                       double *J, double *F, void *stream);
/* the fused engine's register-tiled Gauss-Jordan solver on its own: x[B][n] = A^-1 b, A[B][n][n] row-major,
 * 1 <= n <= 127 */
int yhb_stage_gj(int32_t batch, int32_t n, const double *A, const double *b, double *x, void *stream);
/* the staged engines' panel Gauss-Jordan solver (16 pivot columns per panel, DMMA trailing updates; the core solve of
 * large dense cores and of the block rows of the block-Thomas engine) on its own: X[B][n][nrhs] = A^-1 Bm,
 * A[B][n][n] and Bm[B][n][nrhs] row-major.  Synchronises the stream (it owns temporary device memory). */
int yhb_stage_gjp(int32_t batch, int32_t n, int32_t nrhs, const double *A, const double *Bm, double *X, void *stream);
/* LU with partial pivoting + solve (:620-621): x[B][W] = J^-1 F; J is overwritten by its factors */
int yhb_stage_lu(int32_t batch, int32_t W, double *J, const double *F, double *x, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* YHB_H */
