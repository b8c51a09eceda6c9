// Internal data structures of libyhb: the compiled plan (topology + stamp tables) and the
// per-batch workspace carved out of the caller's scratch buffer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/yhb.h"

namespace yhb {

// waveform slots produced by the device sweep, per circuit: [nwave][S]
//   diode d : base + 0 = gd, + 1 = Id
//   cubic c : base + 0 = g,  + 1 = I
//   bjt   q : base + 0..13 in BJT.get_hb_params return order (BJT.py:582)
constexpr int kBjtWaves = 14;
constexpr int kBjtCapBase = 10;  // Cbc, Cbe, Cbebc, Csc

// A signed reference to a waveform slot (sign = +1 / -1), used by all stamp lists.
struct Term {
    int32_t idx;
    int32_t sgn;
};

// Device-side view of the plan (all pointers are device pointers).
struct PlanDev {
    int N, K, S, W, ntones, K2;
    int nlin, nsrc, nd, nb, nc, nwave;
    int flags;
    int npair;    // structurally nonzero node pairs (J blocks)
    int npair_nl; // pairs with a nonlinear (dense) part; these come first in the pair list
    const double *idft;  // [S][S]  x_s = sum_c idft[s][c] X_c   (:294-300)
    const double *dft;   // [S][S]  X_r = sum_s dft[r][s] x_s    (:302)
    const int *bin_k1;   // [K+1] tone multipliers of every bin: f = k1 f1 + k2 f2
    const int *bin_k2;
    const int *lin_kind;   // [nlin]
    const int *src_kind;   // [nsrc]
    const int *src_nodes;  // [nsrc][2]
    const int *diode_nodes, *bjt_nodes, *cubic_nodes;  // reference numbering, 0 = gnd
    const int *dev_wave_base;                          // [nd + nc + nb] first waveform slot of each device
    // J block structure
    const int *pair_n, *pair_m;  // [npair] 0-based nodes (row = KCL node, col = voltage node)
    const int *pair_of;          // [N][N] -> pair index or -1
    const int *row_ptr;          // [N+1] CSR over pairs sorted by (n, m)
    const int *row_pairs;        // [npair]
    // linear admittance terms per pair: element index, sign (Term.idx = linear element)
    const int *plin_ptr;  // [npair+1]
    const Term *plin;
    // conductance / capacitance waveform terms per nonlinear pair (Term.idx = waveform slot)
    const int *pg_ptr;  // [npair_nl+1]
    const Term *pg;
    const int *pc_ptr;  // [npair_nl+1]  (Term.idx = slot in the accumulated-capacitance array: q*4 + j)
    const Term *pc;
    // time-domain current / charge terms per node (Term.idx = waveform slot)
    const int *ni_ptr;  // [N+1]
    const Term *ni;
    const int *nq_ptr;  // [N+1]
    const Term *nq;
    // independent-source terms per node (Term.idx = source index)
    const int *ns_ptr;  // [N+1]
    const Term *ns;
    // ---- dense core after structural peeling of ideal-source rows / columns (see peel() in yhb_api.cu)
    // A-steps (row singletons, unit gyrator pivot): dV[col] = sgn * (F[row] - sum_others J[row,m] dV[m]), in order.
    // B-steps (column singletons): the same formula, evaluated in REVERSE order after the core solve.
    // Core: J[core_rows, core_cols] y = F[core_rows] - sum_{A-known m} J[row, m] dV[m].
    int Nc, Wc, nA, nB;
    int nAlev, nBlev;  // levels of mutually independent A / B steps (a_lev / b_lev: [levels + 1] step ranges)
    int band;  // > 0: the core is block banded over nodes; kl = ku = band (in matrix elements), band storage
    int btd;   // the core is block TRIdiagonal over nodes and is solved block by block (block-Thomas engine): the
               // Jacobian is stored as Nc x 3 column-major S x S blocks (sub-diagonal, diagonal, super-diagonal)
    const int *core_rows, *core_cols;  // [Nc] node indices
    const int *core_col_of;            // [N] -> position among core_cols or -1
    const int *a_row, *a_col, *a_sgn;  // [nA]
    const int *a_ptr, *a_pair, *a_m;   // other entries of each A row: CSR [nA+1], pair index, column node
    const int *b_row, *b_col, *b_sgn;  // [nB], in the order the steps are processed
    const int *a_lev, *b_lev;
    const int *b_ptr, *b_pair, *b_m;
    const int *r_ptr, *r_pair, *r_m;   // per core row: entries in A-known columns, CSR [Nc+1]
    // ---- block-sparse LU of the core over node blocks (bs_ok): the core is stored as Nc block ROWS, block row k =
    // S rows of [ D_k | blocks (k, j) of the nodes eliminated after k | rhs_k | blocks of nodes eliminated before k ]
    // (structurally nonzero blocks and fill only), and eliminated level by level of the elimination tree: the
    // nodes of a level are independent and are factored concurrently, one warp each (bs_solve in yhb_kernels.cuh)
    int ncblk;                // core node blocks that hold a pair
    const int *cblk;          // [ncblk][3] row node, column node (core positions), pair
    int bs_ok, bs_nlev, bs_total, bs_zero;
    const int *bs_order;      // [Nc] elimination order (core positions)
    const int *bs_pos;        // [Nc] position of a node in the order
    const int *bs_level_ptr;  // [bs_nlev + 1] levels over bs_order
    const int *bs_rowbase;    // [Nc] offset of block row k in the storage (doubles)
    const int *bs_ld;         // [Nc] leading dimension of block row k (odd)
    const int *bs_below_ptr;  // [Nc + 1] CSR over bs_below
    const int *bs_below;      // block rows i (eliminated after k) that hold a block (i, k)
    const int *bs_lev_uniform;  // [bs_nlev] all nodes of the level have the same later blocks and the same rows below:
                                // their Schur updates hit the same target tiles and are applied in one pass
    const int *bs_nlater;     // [Nc] blocks (k, j) with j eliminated after k
    const int *bs_nblk;       // [Nc] blocks in block row k (diagonal included); rhs column = S * bs_nblk[k]... see bs_rhsoff
    const int *bs_rhsoff;     // [Nc] column of rhs_k in block row k
    const int *bs_coloff;     // [Nc][Nc] first column of block (k, j) in block row k, or -1
    const int *bs_blkof;      // [Nc][Nc] node j of the slot-th block of block row k (slot 0 = k itself)
    // every table above points into one device blob; its first blob_hot_bytes hold what the Newton loop reads
    const char *blob;
    int blob_hot_bytes;
};

// pointer fields of PlanDev that lie in the hot prefix of the blob (re-based onto a shared-memory copy by k_fused)
#define YHB_HOT_TABLES(X)                                                                                      \
    X(idft) X(dft) X(diode_nodes) X(bjt_nodes) X(cubic_nodes) X(dev_wave_base) X(pair_n) X(pair_m) X(pair_of) \
    X(row_ptr) X(row_pairs) X(pg_ptr) X(pg) X(pc_ptr) X(pc) X(ni_ptr) X(ni) X(nq_ptr) X(nq) X(core_rows)      \
    X(core_cols) X(core_col_of) X(a_row) X(a_col) X(a_sgn) X(a_ptr) X(a_pair) X(a_m) X(b_row) X(b_col)        \
    X(b_sgn) X(b_ptr) X(b_pair) X(b_m) X(a_lev) X(b_lev) X(r_ptr) X(r_pair) X(r_m) X(bs_order) X(bs_pos) X(bs_level_ptr) X(bs_rowbase)      \
    X(bs_ld) X(bs_nlater) X(bs_nblk) X(bs_rhsoff) X(bs_coloff) X(bs_blkof) X(bs_below_ptr) X(bs_below) X(bs_lev_uniform) X(cblk)

struct Plan {
    PlanDev d;                  // device view
    std::vector<void *> allocs; // device allocations owned by the plan
    std::vector<int> bin_k1, bin_k2;
    // host copies used by the DC solver setup
    std::vector<int> lin_kind, lin_nodes, src_kind, src_nodes, diode_nodes, bjt_nodes, cubic_nodes;
    const int *lin_nodes_dev;   // [nlin][4]
    int num_inductors;
    // cached buffers of yhb_run_host
    void *host_ws;
    size_t host_ws_bytes;
};

// Per-circuit Newton / continuation state (run() :330-387 and hb_loop :423-426 locals).
struct Ctl {
    int mode;       // 0 = direct attempt from V0, 1 = source-stepping ladder, 2 = done
    int converged;  // final answer
    int itercnt;    // Newton iterations of the current hb_loop
    int first;      // next evaluation is the first of an hb_loop (no limiting, reset dQdV)
    int level;      // continuation loop counter (:345)
    int nlevels;    // hb_loop calls so far
    int have_dv;    // a NaN-free Newton step exists (:624-627)
    int newton, lus;
    int fixed_alpha;  // yhb_newton: single hb_loop, no ladder
    double alpha, inc, dec;
    double err;
};

// Workspace layout for a batch (device pointers into the caller's scratch).
struct Workspace {
    int batch;
    Ctl *ctl;            // [B]
    double *V;           // [B][W]   current iterate (aliases res->V when given)
    double *Vlevel;      // [B][W]   Vprev of the ladder (:339, :365)
    double *dV;          // [B][W]   last NaN-free Newton step
    double *vtprev;      // [B][W]   previous iterate's time samples (limiter reference)
    double *F;           // [B][W]
    double *cacc;        // [B][nb][4][S] accumulated capacitance waveforms (never-reset dQdV)
    double *spec;        // [B][npair_nl][2][S]  DFT of the pair conductance / capacitance waveforms
    double *ylin;        // [B][npair][K+1][2]
    double *omega;       // [B][K+1]
    int *negbin;         // [B][K+1] real frequency of the bin is negative (two-tone, :725)
    int stage_first;     // stage entry points: treat the evaluation as the first of an hb_loop
    int dc_valid;        // dcV holds DC operating points (needed for cold starts and after a failed direct attempt)
    double *Is;          // [B][W]  unscaled source vector (:647-684)
    double *scratch;     // [B][scratch_stride] evaluation scratch when it does not fit shared memory
    size_t scratch_stride;
    double *J;           // staged engine: core Jacobians [B][j_stride] (stage entry point: caller's full [B][W][W])
    size_t j_stride;
    double *Inl;         // [B][W] optional output: nonlinear current spectrum of the last evaluation (or nullptr)
    double *Ic;          // [B][nb][S] optional output: collector-current samples of the last evaluation (:512)
    // oscillator solve (yhb_solve_oscillator): circuits to leave alone, sensitivity output, d(linear admittance)/df
    const int *skip;     // [B] or nullptr: circuit b is not solved (its oscillator is finished)
    double *sens;        // [B][2][W] or nullptr: J^-1 dF/df and J^-1 dF/dA at the converged point (exact Jacobian)
    double *dylin;       // [B][npair][K+1] or nullptr: f * d Im(ylin) / df  (capacitors +, inductors -)
    int sens_src_n1, sens_src_n2;  // nodes of the probe current source (reference numbering): dF/dA = -dIs/dA
    // buffers of the oscillator solve (carved for every batch; small next to the Newton state)
    void *osc_state;     // [B] OscState
    double *osc_Vacc;    // [B][W]
    double *osc_Inl;     // [B][W]
    double *osc_sens;    // [B][2][W]
    double *osc_dylin;   // [B][npair][K+1]
    int *osc_skip;       // [B]
    int *osc_live;       // [1]
    double *pairwav;     // [B][npair_nl][2][S] pair conductance / capacitance waveforms (staged engine)
    double *xstep;       // [B][x_stride]  staged engine: Newton step vector + scratch of newton_pre / newton_post
    size_t x_stride;
    double *rhs;         // [B][Wc+1]
    // panel Gauss-Jordan solver of the staged engines (yhb_dense.cuh), per circuit; gj_n = pivots per system
    // (Wc for a dense core, S for the block rows of the block-Thomas engine), 0 = not used by this plan
    int gj_n, gj_lpad;
    double *gj_L;        // [B][16][gj_lpad] multipliers of the current panel
    double *gj_U;        // [B][16][gj_ld]   snapshot of the pivot rows of the current panel
    double *gj_T;        // [B][16][16]
    int gj_ld;           // row stride of a system's working matrix
    double *gj_piv;      // [B][gj_n]
    int *gj_rowof;       // [B][gj_n]
    int *gj_used;        // [B][gj_n]
    int *gj_pp;          // [B][16]
    int *gj_fail;        // [B]
    int *list[2];        // [B] active-circuit lists (ping-pong)
    int *count;          // [4] device counters: active[0], active[1], done, spare
    int32_t *level_iters;  // [B][YHB_MAX_LEVELS]
    double *level_alpha;   // [B][YHB_MAX_LEVELS]
    double *dcV;         // [B][N] DC node voltages
    int *dcinfo;         // [B][2]
    double *dcscratch;   // DC solver scratch
    size_t dc_stride;
    // cold solves with the DC analysis overlapped (yhb_solve_cold): the HB kernel takes the circuits in order[] (most
    // work first: sorted by the drive level key[]), waits for dc_ready[b] and builds its start vector from dcV itself
    double *key;         // [B] sum of the AC source amplitudes (k_prepare)
    int *order;          // [B] queue position -> circuit
    int *dc_ready;       // [B + 8] dc_ready[l] = 1: the DC point of circuit l (a group leader) is in dcV; then the counters of
                         // an overlapped cold solve (see the ticket logic of k_fused)
    unsigned long long *dc_hash;  // [B] hash of every circuit's DC inputs (k_dc_hash)
    int *dc_leader;      // [B] first circuit with the same DC inputs (k_dc_group)
    int use_order;       // the fused kernel pops circuits through order[]
    int dc_wait;         // cold start beside the DC kernel: the fused kernel takes the circuits whose operating point (their
                         // group leader's) has arrived, defers the others to list[0], and initialises V from dcV itself
};

void set_error(const std::string &msg);

}  // namespace yhb
