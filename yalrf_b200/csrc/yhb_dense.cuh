// Dense-block kernels of the staged engines: Gauss-Jordan elimination with partial pivoting of HBM / L2 resident
// systems [A | carried columns] by PANELS of 16 pivot columns, for a batch of systems at once.  This is the core
// solve (:620-621) of
//   * large dense cores            (BASELINE config 3 at [8,8]: Wc = 289, one system per circuit, carried = rhs), and of
//   * the block-Thomas engine      (config 5: per block row one system [D_i | U_i | b_i] of S = 1089 pivots with
//                                   S + 1 carried columns, then the Schur complement of the next block row),
// which round 1 ran as cuBLAS / cuSOLVER calls from a host loop.  Per panel two launches:
//   k_gjp_panel   one CTA per system factors the panel's 16 columns of all n rows, rows in registers (up to 1152
//                 rows; a shared-memory variant takes more): per column a vote -- warp arg-max by redux on the bit
//                 pattern of |a| (first row of maximal |a| among the rows not yet used: the idamax rule of dgetrf),
//                 cross-warp through shared memory, re-reduced by every warp -- the pivot row published by its owner,
//                 multipliers l = -(a * (1 / pivot)) as dgetf2 scales them, in-panel update of the thread's own rows.
//                 Rows are never moved (implicit pivoting) and rows above the pivot are eliminated too (Gauss-Jordan:
//                 no triangular solves, full-height rank-16 updates).  Its tail snapshots the 16 pivot rows for the
//                 trailing columns into a separate buffer U.
//   k_gjp_update  CTAs of 128 rows x 32 columns complete their columns of U (row s is still subject to the panel's
//                 earlier pivots) and apply A += L (n x 16) . U (16 x cols) with FP64 tensor-core instructions
//                 (mma.sync.m8n8k4.f64, DMMA): the trailing-update contraction, the only place the tensor pipe is
//                 used, as BASELINE.json asks.
// k_gjp_finish scales the carried columns by the pivots and un-permutes them; k_btd_schur is the DMMA GEMM
// D_{i+1} -= L_{i+1} U'_i (shared-memory tiles, cp.async double buffering); k_btd_back the back substitution.
#pragma once
#include "yhb_internal.cuh"

namespace yhb {

constexpr int kGjpNB = 16;  // pivot columns per panel
// first carried column of a system of n pivots: the pivot columns are padded to whole panels, so that the trailing
// update of a panel (columns >= 16 (j + 1)) covers every carried column also for the last, partial panel
__host__ __device__ inline int gjp_npc(int n) { return (n + kGjpNB - 1) & ~(kGjpNB - 1); }

__device__ __forceinline__ void dmma884(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(d0), "+d"(d1)
                 : "d"(a), "d"(b));
}

// A batch of systems under elimination.  System index: pos = blockIdx -> s = list ? list[pos] : pos (active circuits).
struct GjpBatch {
    double *M;        // working matrices, system s at M + s * m_stride (+ m_off), [n][ld] row-major:
                      // [A (n columns) | pad to gjp_npc(n) | carried columns], ncols = gjp_npc(n) + carried
    size_t m_stride, m_off;
    double *L;        // multipliers of the current panel, system s at L + s * kGjpNB * lpad: [16][lpad]
    double *U;        // snapshot of the pivot rows of the current panel, system s at U + s * kGjpNB * ld: [16][ld]
    double *T;        // [16][16] per system: T[s][s'] = l_{s'}[p_s], s' < s
    double *piv;      // [n] per system (stride n): pivot value of every column
    int *rowof;       // [n] per system: row that eliminated column c (-1: no admissible pivot)
    int *used;        // [n] per system: row has been a pivot row
    int *pp;          // [16] per system: pivot rows of the current panel (-1: none)
    int *fail;        // per system: a column had no admissible pivot (all NaN): the solution is reported as NaN
    const int *list;  // optional indirection (staged engine's active list) and its length
    const int *count;
    int n, ld, ncols, lpad;
    // Two systems per circuit (the block-Thomas engine eliminates its chain from both ends): launch positions
    // [0, npos) are the circuits' first systems (matrix at m_off), [npos, 2 npos) their second ones (matrix at m_off2);
    // the scratch buffers (L, U, T, piv, rowof, used, pp, fail) of the second systems sit vs_stride systems further on.
    int nchain, npos, vs_stride;
    size_t m_off2;
};

struct GjpSys {
    int b;      // circuit (-1: nothing to do at this launch position)
    int vs;     // index of the system's scratch buffers
    int chain;  // 0 / 1
};

__device__ __forceinline__ GjpSys gjp_locate(const GjpBatch &g, int pos) {
    int chain = 0;
    if (g.nchain > 1 && pos >= g.npos) {
        chain = 1;
        pos -= g.npos;
    }
    if (g.count && pos >= *g.count) return GjpSys{-1, -1, 0};
    const int b = g.list ? g.list[pos] : pos;
    return GjpSys{b, b + chain * g.vs_stride, chain};
}

__device__ __forceinline__ double *gjp_matrix(const GjpBatch &g, const GjpSys &gs) {
    return g.M + (size_t)gs.b * g.m_stride + (gs.chain ? g.m_off2 : g.m_off);
}

// warp-wide arg-max of (key_hi, key_lo) with ties to the lowest row; key_hi == 0 means "no candidate"
__device__ __forceinline__ void gjp_warp_best(unsigned &hi, unsigned &lo, int &row) {
    const unsigned mhi = __reduce_max_sync(0xffffffffu, hi);
    const bool c1 = hi == mhi && mhi != 0u;
    const unsigned mlo = __reduce_max_sync(0xffffffffu, c1 ? lo : 0u);
    const bool c2 = c1 && lo == mlo;
    const unsigned r = __reduce_min_sync(0xffffffffu, c2 ? (unsigned)row : 0x7fffffffu);
    hi = mhi;
    lo = mlo;
    row = mhi != 0u ? (int)r : -1;
}

// Tail of a panel factorisation, by the panel's CTA after its last column: a SNAPSHOT of the 16 pivot rows for the
// columns right of the panel, U[s][c] = M[p_s][c] (the update kernel's CTAs split the rows among themselves, and all of
// them read the pivot rows that some of them overwrite), and the 16 x 16 table T[s][s'] = l_{s'}[p_s], s' < s: what
// pivot row s still owes to the panel's earlier pivots.  The update kernel completes its columns of U with it.
__device__ __forceinline__ void gjp_panel_tail(const GjpBatch &g, int sys, int j, const double *M, const double *Lg, int *pps) {
    const int tid = threadIdx.x, nt = blockDim.x;
    const int ld = g.ld, lpad = g.lpad;
    __syncthreads();  // the panel's global writes (pp, multipliers) are visible to the whole CTA
    if (tid < kGjpNB) pps[tid] = g.pp[sys * kGjpNB + tid];
    __syncthreads();
    double *T = g.T + (size_t)sys * kGjpNB * kGjpNB;
    for (int e = tid; e < kGjpNB * kGjpNB; e += nt) {
        const int s = e >> 4, s1 = e & 15;
        T[e] = (s1 < s && pps[s] >= 0 && pps[s1] >= 0) ? Lg[(size_t)s1 * lpad + pps[s]] : 0.;
    }
    double *U = g.U + (size_t)sys * kGjpNB * ld;
    // one column pair per thread and trip, the 16 loads of a trip in flight together
    const int c0 = (j + 1) * kGjpNB;
    const double *rowp[kGjpNB];
#pragma unroll
    for (int s = 0; s < kGjpNB; ++s) rowp[s] = M + (size_t)(pps[s] >= 0 ? pps[s] : 0) * ld;
    for (int c = c0 + 2 * tid; c < g.ncols; c += 2 * nt) {  // c0 and ld are even: 16-byte accesses, the pair stays inside the row
        double2 v[kGjpNB];
#pragma unroll
        for (int s = 0; s < kGjpNB; ++s) v[s] = *reinterpret_cast<const double2 *>(rowp[s] + c);
#pragma unroll
        for (int s = 0; s < kGjpNB; ++s)
            *reinterpret_cast<double2 *>(U + (size_t)s * ld + c) = pps[s] >= 0 ? v[s] : make_double2(0., 0.);
    }
}

// Panel j (columns 16 j .. 16 j + 15 of A) of every system, rows in REGISTERS: thread t owns rows t + k blockDim.x,
// k < RPT, and keeps their 16 panel entries in registers (the column loop is fully unrolled).  Per column two
// barriers: the vote (warp arg-max by redux on the bit pattern of |a|, cross-warp through shared memory, re-reduced by
// every warp) and the pivot row, which its owner publishes to shared memory.  n <= 576 RPT.
template <int RPT>
__global__ void __launch_bounds__(576, 1) k_gjp_panel(GjpBatch g, int j) {
    __shared__ unsigned s_hi[32], s_lo[32];
    __shared__ int s_row[32];
    __shared__ double urow[2][kGjpNB];
    __shared__ int pps[kGjpNB];
    const GjpSys gs = gjp_locate(g, blockIdx.x);
    if (gs.b < 0) return;
    const int sys = gs.vs;  // scratch buffers
    const int n = g.n, ld = g.ld, lpad = g.lpad;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    double *M = gjp_matrix(g, gs);
    double *Lg = g.L + (size_t)sys * kGjpNB * lpad;
    int *used = g.used + (size_t)sys * n;
    const int col0 = j * kGjpNB;
    const int nb = min(kGjpNB, n - col0);
    double a[RPT][kGjpNB];
    unsigned um = 0;  // bit k: row tid + k nt has been a pivot row (or does not exist)
#pragma unroll
    for (int k = 0; k < RPT; ++k) {
        const int r = tid + k * nt;
        if (r < n) {
            const double2 *src = reinterpret_cast<const double2 *>(M + (size_t)r * ld + col0);  // 16-byte aligned: ld, col0 even
#pragma unroll
            for (int c = 0; c < kGjpNB; c += 2) {
                const double2 v = src[c >> 1];
                a[k][c] = c < nb ? v.x : 0.;
                a[k][c + 1] = c + 1 < nb ? v.y : 0.;
            }
            if (j > 0) um |= (used[r] ? 1u : 0u) << k;
            else used[r] = 0;  // a new system: the flags of the previous one are stale
        } else {
#pragma unroll
            for (int c = 0; c < kGjpNB; ++c) a[k][c] = 0.;
            um |= 1u << k;
        }
    }
    if (j == 0 && tid == 0) g.fail[sys] = 0;
#pragma unroll
    for (int s = 0; s < kGjpNB; ++s) {
        if (s >= nb) {  // past the last column: empty steps keep the update kernel uniform
#pragma unroll
            for (int k = 0; k < RPT; ++k)
                if (tid + k * nt < n) Lg[(size_t)s * lpad + tid + k * nt] = 0.;
            if (tid == 0) g.pp[sys * kGjpNB + s] = -1;
            continue;
        }
        // arg-max of |a| over the rows not yet used; NaN entries never win (comparison false)
        double best = -1.;
        int bi = -1;
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const double v = fabs(a[k][s]);
            if (!((um >> k) & 1u) && v > best) { best = v; bi = tid + k * nt; }
        }
        const unsigned long long bits = (unsigned long long)__double_as_longlong(best);
        unsigned hi = bi >= 0 ? (unsigned)(bits >> 32) + 1u : 0u, lo = (unsigned)bits;
        int row = bi;
        gjp_warp_best(hi, lo, row);
        if (lane == 0) { s_hi[warp] = hi; s_lo[warp] = lo; s_row[warp] = row; }
        __syncthreads();
        hi = lane < nwarp ? s_hi[lane] : 0u;
        lo = lane < nwarp ? s_lo[lane] : 0u;
        row = lane < nwarp ? s_row[lane] : -1;
        gjp_warp_best(hi, lo, row);
        const int p = row;  // identical in every thread
        const int par = s & 1;
        // the owner publishes the pivot row (its panel entries right of the pivot) and records the pivot
#pragma unroll
        for (int k = 0; k < RPT; ++k)
            if (p >= 0 && tid + k * nt == p) {
#pragma unroll
                for (int c = s; c < kGjpNB; ++c) urow[par][c] = a[k][c];
                um |= 1u << k;
                used[p] = 1;
                g.pp[sys * kGjpNB + s] = p;
                g.rowof[(size_t)sys * n + col0 + s] = p;
                g.piv[(size_t)sys * n + col0 + s] = a[k][s];
            }
        if (p < 0 && tid == 0) {  // no admissible pivot: the column is skipped, the system is reported as NaN
            g.pp[sys * kGjpNB + s] = -1;
            g.rowof[(size_t)sys * n + col0 + s] = -1;
            g.piv[(size_t)sys * n + col0 + s] = 0.;
            g.fail[sys] = 1;
        }
        __syncthreads();
        if (p < 0) {
#pragma unroll
            for (int k = 0; k < RPT; ++k)
                if (tid + k * nt < n) Lg[(size_t)s * lpad + tid + k * nt] = 0.;
            continue;
        }
        const double rinv = 1. / urow[par][s];
#pragma unroll
        for (int k = 0; k < RPT; ++k) {
            const int r = tid + k * nt;
            if (r >= n) continue;
            const double l = r == p ? 0. : -(a[k][s] * rinv);  // Gauss-Jordan: rows above the pivot are eliminated too
            Lg[(size_t)s * lpad + r] = l;
#pragma unroll
            for (int c = s + 1; c < kGjpNB; ++c) a[k][c] = fma(l, urow[par][c], a[k][c]);
        }
    }
    gjp_panel_tail(g, sys, j, M, Lg, pps);
}

// The same panel factorisation with the panel in shared memory (column-major, 16 * (n | 1) doubles): any n up to
// ~14 000 rows, one barrier per column, about 4x the instructions of the register version.
__global__ void __launch_bounds__(1024, 1) k_gjp_panel_smem(GjpBatch g, int j) {
    extern __shared__ double P[];  // [16][npad] column-major panel
    __shared__ unsigned s_hi[2][32], s_lo[2][32];
    __shared__ int s_row[2][32];
    __shared__ int pps[kGjpNB];
    const GjpSys gs = gjp_locate(g, blockIdx.x);
    if (gs.b < 0) return;
    const int sys = gs.vs;  // scratch buffers
    const int n = g.n, ld = g.ld, npad = n | 1, lpad = g.lpad;
    const int tid = threadIdx.x, nt = blockDim.x, lane = tid & 31, warp = tid >> 5, nwarp = nt >> 5;
    double *M = gjp_matrix(g, gs);
    double *Lg = g.L + (size_t)sys * kGjpNB * lpad;
    int *used = g.used + (size_t)sys * n;
    const int col0 = j * kGjpNB;
    const int nb = min(kGjpNB, n - col0);
    for (int idx = tid; idx < n * kGjpNB; idx += nt) {
        const int r = idx >> 4, c = idx & 15;
        P[c * npad + r] = c < nb ? M[(size_t)r * ld + col0 + c] : 0.;
    }
    unsigned um = 0;  // bit k: this thread's k-th row (tid + k nt) has been a pivot row
    if (j > 0) {
        int k = 0;
        for (int r = tid; r < n; r += nt, ++k) um |= (used[r] ? 1u : 0u) << k;
    } else {
        for (int r = tid; r < n; r += nt) used[r] = 0;  // a new system: the flags of the previous one are stale
    }
    if (j == 0 && tid == 0) g.fail[sys] = 0;
    __syncthreads();
    for (int s = 0; s < kGjpNB; ++s) {
        double *Ps = P + s * npad;
        if (s >= nb) {
            for (int r = tid; r < n; r += nt) Lg[(size_t)s * lpad + r] = 0.;
            if (tid == 0) g.pp[sys * kGjpNB + s] = -1;
            continue;
        }
        double best = -1.;
        int bi = -1;
        {
            int k = 0;
            for (int r = tid; r < n; r += nt, ++k) {
                const double v = fabs(Ps[r]);
                if (!((um >> k) & 1u) && v > best) { best = v; bi = r; }
            }
        }
        const unsigned long long bits = (unsigned long long)__double_as_longlong(best);
        unsigned hi = bi >= 0 ? (unsigned)(bits >> 32) + 1u : 0u, lo = (unsigned)bits;
        int row = bi;
        gjp_warp_best(hi, lo, row);
        const int par = s & 1;
        if (lane == 0) { s_hi[par][warp] = hi; s_lo[par][warp] = lo; s_row[par][warp] = row; }
        __syncthreads();
        hi = lane < nwarp ? s_hi[par][lane] : 0u;
        lo = lane < nwarp ? s_lo[par][lane] : 0u;
        row = lane < nwarp ? s_row[par][lane] : -1;
        gjp_warp_best(hi, lo, row);
        const int p = row;  // identical in every thread
        if (tid == 0) {
            g.pp[sys * kGjpNB + s] = p;
            g.rowof[(size_t)sys * n + col0 + s] = p;
            g.piv[(size_t)sys * n + col0 + s] = p >= 0 ? Ps[p] : 0.;
            if (p < 0) g.fail[sys] = 1;
        }
        if (p < 0) {
            for (int r = tid; r < n; r += nt) Lg[(size_t)s * lpad + r] = 0.;
            continue;
        }
        const double rinv = 1. / Ps[p];
        int k = 0;
        for (int r = tid; r < n; r += nt, ++k) {
            if (r == p) {
                um |= 1u << k;
                used[r] = 1;
                Lg[(size_t)s * lpad + r] = 0.;
                continue;
            }
            const double l = -(Ps[r] * rinv);
            Lg[(size_t)s * lpad + r] = l;
            for (int s2 = s + 1; s2 < nb; ++s2) {
                double *q = P + s2 * npad;
                q[r] = fma(l, q[p], q[r]);
            }
        }
        // the next step's barrier orders these writes (own rows only) before any other thread reads row p' of them
    }
    gjp_panel_tail(g, sys, j, M, Lg, pps);
}

// Trailing update of panel j on the tensor pipe:  M += L (n x 16) . U (16 x columns), U = the completed pivot rows.
// One CTA = 128 rows x 32 columns (grid: column chunks right of the panel, row chunks, systems); one warp = 16 rows:
// 2 x 4 tiles of 8 x 8, every operand loaded once, all global loads issued before the first barrier.  Warp 0 first
// completes the CTA's 32 columns of the pivot-row snapshot, U[s] += sum_{s' < s} T[s][s'] U[s'] (sequential in s).
constexpr int kGjpUpdCols = 32, kGjpUpdRows = 128;
__global__ void __launch_bounds__(256) k_gjp_update(GjpBatch g, int j) {
    __shared__ double Ts[kGjpNB][kGjpNB];
    __shared__ double Us[kGjpNB][kGjpUpdCols + 4];
    const GjpSys gs = gjp_locate(g, blockIdx.z);
    if (gs.b < 0) return;
    const int sys = gs.vs;  // scratch buffers
    const int n = g.n, ld = g.ld, lpad = g.lpad;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int gq = lane >> 2, t = lane & 3;
    const int c0 = (j + 1) * kGjpNB + kGjpUpdCols * blockIdx.x;
    const int r0 = kGjpUpdRows * blockIdx.y + 16 * warp;
    double *M = gjp_matrix(g, gs);
    const double *Lg = g.L + (size_t)sys * kGjpNB * lpad;
    const double *U = g.U + (size_t)sys * kGjpNB * ld;
    Ts[tid >> 4][tid & 15] = g.T[(size_t)sys * kGjpNB * kGjpNB + tid];
    for (int e = tid; e < kGjpNB * kGjpUpdCols; e += 256) {
        const int s = e >> 5, c = e & 31;
        Us[s][c] = c0 + c < g.ncols ? U[(size_t)s * ld + c0 + c] : 0.;
    }
    // this warp's operands: multipliers and C tiles (issued now, consumed after the barriers)
    double af[2][4];
    double2 cv[2][4];
    bool rok[2], cok[4];
#pragma unroll
    for (int ct = 0; ct < 4; ++ct) cok[ct] = c0 + 8 * ct + 2 * t < g.ncols;  // ld is even: a pair never leaves the row
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int r = r0 + 8 * i + gq;
        rok[i] = r < n;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) af[i][kk] = rok[i] ? Lg[(size_t)(4 * kk + t) * lpad + r] : 0.;
        const double *crow = M + (size_t)(rok[i] ? r : 0) * ld + c0 + 2 * t;
#pragma unroll
        for (int ct = 0; ct < 4; ++ct)
            cv[i][ct] = (rok[i] && cok[ct]) ? *reinterpret_cast<const double2 *>(crow + 8 * ct) : make_double2(0., 0.);
    }
    __syncthreads();
    if (warp == 0) {
        double u[kGjpNB];
#pragma unroll
        for (int s = 0; s < kGjpNB; ++s) u[s] = Us[s][lane];
#pragma unroll
        for (int s = 1; s < kGjpNB; ++s) {
            double a0 = u[s], a1 = 0.;
#pragma unroll
            for (int s1 = 0; s1 < kGjpNB; s1 += 2) {
                if (s1 < s) a0 = fma(Ts[s][s1], u[s1], a0);
                if (s1 + 1 < s) a1 = fma(Ts[s][s1 + 1], u[s1 + 1], a1);
            }
            u[s] = a0 + a1;
            Us[s][lane] = u[s];
        }
    }
    __syncthreads();
    double bf[4][4];
#pragma unroll
    for (int ct = 0; ct < 4; ++ct)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) bf[ct][kk] = Us[4 * kk + t][8 * ct + gq];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int ct = 0; ct < 4; ++ct)
#pragma unroll
            for (int kk = 0; kk < 4; ++kk) dmma884(cv[i][ct].x, cv[i][ct].y, af[i][kk], bf[ct][kk]);
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        double *crow = M + (size_t)(rok[i] ? r0 + 8 * i + gq : 0) * ld + c0 + 2 * t;
#pragma unroll
        for (int ct = 0; ct < 4; ++ct)
            if (rok[i] && cok[ct]) *reinterpret_cast<double2 *>(crow + 8 * ct) = cv[i][ct];
    }
}

// Carried columns -> solution: X[c][k] = M[rowof[c]][gjp_npc(n) + k] / piv[c].  out0 (optional): columns 0 .. nout0 - 1 of the
// carried block, [n][ld0] row-major (pad columns zeroed); out1 (optional): carried column col1 as a vector [n].
struct GjpFinish {
    double *out0;
    size_t out0_stride, out0_off, out0_off2;  // per circuit; _off2: the circuits' second systems (GjpBatch::nchain)
    int nout0, ld0;
    double *out1;
    size_t out1_stride, out1_off, out1_off2;
    int col1;
};

__global__ void __launch_bounds__(256) k_gjp_finish(GjpBatch g, GjpFinish f) {
    const GjpSys gs = gjp_locate(g, blockIdx.y);
    if (gs.b < 0) return;
    const int sys = gs.vs;  // scratch buffers
    const int n = g.n, ld = g.ld;
    const double *M = gjp_matrix(g, gs);
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    const int fail = g.fail[sys];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int c = blockIdx.x * 8 + warp; c < n; c += gridDim.x * 8) {
        const int r = g.rowof[(size_t)sys * n + c];
        const double pv = g.piv[(size_t)sys * n + c];
        const bool bad = fail || r < 0;
        const double *src = M + (size_t)(bad ? 0 : r) * ld + gjp_npc(n);
        if (f.out0) {
            double *dst = f.out0 + (size_t)gs.b * f.out0_stride + (gs.chain ? f.out0_off2 : f.out0_off) + (size_t)c * f.ld0;
            for (int k = lane; k < f.ld0; k += 32) dst[k] = k < f.nout0 ? (bad ? nan : src[k] / pv) : 0.;
        }
        if (f.out1 && lane == 0)
            f.out1[(size_t)gs.b * f.out1_stride + (gs.chain ? f.out1_off2 : f.out1_off) + c] = bad ? nan : src[f.col1] / pv;
    }
}

// ---------------------------------------------------------------------------------------------------------------- //
// Block-Thomas engine: storage of one circuit's block-tridiagonal core (Nc block rows of S x S node blocks).  The chain is
// eliminated from BOTH ends towards the middle block row mid = Nc / 2 (two block rows in flight per circuit: half the
// serial depth).  Upper rows r <= mid:
//   Mr [S][ldm]  = [ D_r | pad to npc = gjp_npc(S) | U_r (block (r, r+1)) | rhs_r | pad ]
//   Lr [S][lds]  = sub-diagonal block (r, r-1) (pad column zero); after block row r is eliminated this storage holds
//                  U'_r = D_r^-1 U_r
// Lower rows r > mid are stored mirrored -- Mr = [ D_r | pad | L_r | rhs_r | pad ], side slot = U_r, later L'_r = D_r^-1 L_r --
// so that the same kernels run both chains: eliminating x_r from its inner neighbour's row is
//   [D_q | rhs_q] -= (side slot of q) . [X'_r | b'_r],   q = r + 1 (upper chain) or r - 1 (lower chain).
// The middle row takes one such update from each side (the one from below reads U_mid out of Mr's carried block).
// ---------------------------------------------------------------------------------------------------------------- //
__host__ __device__ inline int btd_ldm(int S) { return (gjp_npc(S) + S + 2) & ~1; }
__host__ __device__ inline int btd_lds(int S) { return (S + 2) & ~1; }
__host__ __device__ inline size_t btd_row_doubles(int S) { return (size_t)S * btd_ldm(S) + (size_t)S * btd_lds(S); }
__host__ __device__ inline int btd_mid(int Nc) { return Nc / 2; }

// C[S x S] (leading S columns of Mr) -= A ([S][lda]) . B (X'_prev, [S][lds]);  rhs column -= A . b'_prev.
// 64 x 64 tiles, k slabs of 16 staged in shared memory by cp.async (double buffered), 8 warps of 16 x 32 outputs each.
// Launch positions [0, npos) work on (r[0], rprev[0]), positions [npos, 2 npos) on (r[1], rprev[1]) (nchain = 2).
constexpr int kSchurBM = 64, kSchurBN = 64, kSchurBK = 16;
struct SchurArgs {
    double *J;        // circuit b at J + b * j_stride
    size_t j_stride;
    double *rhs;      // [B][Wc + 1]: b' of the eliminated block rows / rhs of the others
    size_t rhs_stride;
    const int *list, *count;
    int S;
    int nchain, npos;
    int r[2];         // block row being updated (Schur) / substituted (back)
    int rprev[2];     // the eliminated neighbour whose X' and b' (Schur) / the solved neighbour whose x (back) is used
    int a_in_m[2];    // Schur: the coupling block A is the carried block of Mr ([S][ldm] at column npc) instead of the side slot
};

__device__ __forceinline__ int schur_locate(const SchurArgs &a, int pos, int &chain) {
    chain = 0;
    if (a.nchain > 1 && pos >= a.npos) {
        chain = 1;
        pos -= a.npos;
    }
    if (a.count && pos >= *a.count) return -1;
    return a.list ? a.list[pos] : pos;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz));
}

__global__ void __launch_bounds__(256) k_btd_schur(SchurArgs a) {
    __shared__ __align__(16) double As[2][kSchurBM][kSchurBK + 4];
    __shared__ __align__(16) double Bs[2][kSchurBK][kSchurBN + 4];
    int chain;
    const int b = schur_locate(a, blockIdx.z, chain);
    if (b < 0) return;
    const int S = a.S, ldm = btd_ldm(S), lds = btd_lds(S);
    const int ar = a.r[chain], arp = a.rprev[chain];
    double *Jb = a.J + (size_t)b * a.j_stride;
    double *Mr = Jb + (size_t)ar * btd_row_doubles(S);
    const bool inm = a.a_in_m[chain] != 0;
    const double *A = inm ? Mr + gjp_npc(S) : Mr + (size_t)S * ldm;                    // coupling block of row r
    const int lda = inm ? ldm : lds;
    const double *Bm = Jb + (size_t)arp * btd_row_doubles(S) + (size_t)S * ldm;         // X'_prev
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int r0 = blockIdx.y * kSchurBM;
    const int ntx = (S + kSchurBN - 1) / kSchurBN;
    if ((int)blockIdx.x == ntx) {
        // right-hand side: rhs_r[row] -= sum_k A[row][k] b'_{r-1}[k], one warp per row
        const double *x = a.rhs + (size_t)b * a.rhs_stride + (size_t)arp * S;
        for (int rr = warp; rr < kSchurBM; rr += 8) {
            const int row = r0 + rr;
            if (row >= S) break;
            const double *arow = A + (size_t)row * lda;
            double acc = 0.;
            for (int k = lane; k < S; k += 32) acc = fma(arow[k], x[k], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
            if (lane == 0) Mr[(size_t)row * ldm + gjp_npc(S) + S] -= acc;
        }
        return;
    }
    const int c0 = blockIdx.x * kSchurBN;
    const int nk = (S + kSchurBK - 1) / kSchurBK;
    auto stage = [&](int buf, int kt) {
        const int k0 = kt * kSchurBK;
        // A tile: 64 rows x 16 k = 512 chunks of 2 doubles... 16-byte chunks: 8 per row
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int ch = tid + 256 * i;
            const int rr = ch >> 3, kc = (ch & 7) * 2;
            const bool ok = r0 + rr < S && k0 + kc < lds;  // (a chunk may reach one column past S: times a zero-filled row of B)
            cp_async16(&As[buf][rr][kc], A + (size_t)(ok ? r0 + rr : 0) * lda + (ok ? k0 + kc : 0), ok);
        }
        // B tile: 16 k x 64 columns: 32 chunks per row
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const int ch = tid + 256 * i;
            const int kk = ch >> 5, cc = (ch & 31) * 2;
            const bool ok = k0 + kk < S && c0 + cc < lds;
            cp_async16(&Bs[buf][kk][cc], Bm + (size_t)(ok ? k0 + kk : 0) * lds + (ok ? c0 + cc : 0), ok);
        }
        asm volatile("cp.async.commit_group;");
    };
    const int gq = lane >> 2, t = lane & 3;
    const int wr = (warp >> 1) * 16, wc = (warp & 1) * 32;
    double c[2][4][2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) c[i][jj][0] = c[i][jj][1] = 0.;
    stage(0, 0);
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) {
            stage(buf ^ 1, kt + 1);
            asm volatile("cp.async.wait_group 1;");
        } else {
            asm volatile("cp.async.wait_group 0;");
        }
        __syncthreads();
        // columns k >= S of A's last slab hold the pad column (zero) or zero fill; rows k >= S of B are zero filled
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            double af[2], bfr[4];
#pragma unroll
            for (int i = 0; i < 2; ++i) af[i] = -As[buf][wr + 8 * i + gq][4 * kk + t];
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) bfr[jj] = Bs[buf][4 * kk + t][wc + 8 * jj + gq];
#pragma unroll
            for (int i = 0; i < 2; ++i)
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) dmma884(c[i][jj][0], c[i][jj][1], af[i], bfr[jj]);
        }
        __syncthreads();
    }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        const int row = r0 + wr + 8 * i + gq;
        if (row >= S) continue;
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
            const int col = c0 + wc + 8 * jj + 2 * t;
            double *dst = Mr + (size_t)row * ldm + col;
            if (col < S) dst[0] += c[i][jj][0];
            if (col + 1 < S) dst[1] += c[i][jj][1];
        }
    }
}

// back substitution of the block-Thomas chains: x_r = b'_r - X'_r x_q, q = rprev (the neighbour nearer the middle), in place
// in rhs.  grid (row chunks, circuits x chains)
__global__ void __launch_bounds__(256) k_btd_back(SchurArgs a) {
    int chain;
    const int b = schur_locate(a, blockIdx.y, chain);
    if (b < 0) return;
    const int S = a.S, ldm = btd_ldm(S), lds = btd_lds(S);
    const double *U = a.J + (size_t)b * a.j_stride + (size_t)a.r[chain] * btd_row_doubles(S) + (size_t)S * ldm;  // X'_r
    double *x = a.rhs + (size_t)b * a.rhs_stride + (size_t)a.r[chain] * S;
    const double *xn = a.rhs + (size_t)b * a.rhs_stride + (size_t)a.rprev[chain] * S;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int row = blockIdx.x * 8 + warp; row < S; row += gridDim.x * 8) {
        const double *ur = U + (size_t)row * lds;
        double acc = 0.;
        for (int k = lane; k < S; k += 32) acc = fma(ur[k], xn[k], acc);
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
        if (lane == 0) x[row] -= acc;
    }
}

}  // namespace yhb
