// libyhb.so -- C ABI (include/yhb.h) over the staged harmonic-balance kernels.
// Host side: plan compilation (stamp tables), workspace carving, and the launch loop that plays
// run() / hb_loop (yalrf/Analyses/MultiToneHarmonicBalance.py:234-631) for a whole batch.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "yhb_dc.cuh"
#include "yhb_dense.cuh"
#include "yhb_kernels.cuh"
#include "yhb_osc.cuh"

namespace yhb {

static thread_local std::string g_err;
void set_error(const std::string &msg) { g_err = msg; }

#define YHB_CUDA(expr)                                                                         \
    do {                                                                                       \
        cudaError_t e_ = (expr);                                                               \
        if (e_ != cudaSuccess) {                                                               \
            set_error(std::string(#expr) + ": " + cudaGetErrorString(e_));                     \
            return -2;                                                                         \
        }                                                                                      \
    } while (0)

static inline size_t align_up(size_t x, size_t a = 256) { return (x + a - 1) / a * a; }

struct FusedChoice {
    bool ok;
    int NT;       // 8 x 8 tiles per dimension of the core system (0 = scalar elimination in shared memory)
    size_t smem;  // dynamic shared memory bytes
    int hot;      // bit 0: plan tables staged in shared memory; bit 1: per-circuit ylin / omega cached in shared memory
};

constexpr size_t kMaxEvalSmem = 200 * 1024;

// instantiated tile counts of the fused engine (a core of Wc unknowns needs Wc + 1 <= 8 NT)
#define YHB_FOR_NT(X) X(3) X(6) X(8) X(11) X(16)

// block-Thomas engine (block Gauss-Jordan of [D_i | U_i | b_i] + Schur complements), multiply-adds x 2 per core solve:
// per block row S pivots x S rows x (columns right of the pivot in D_i + S of U_i + 1), then D_{i+1} -= L U' (S^3)
static double btd_flops(const PlanDev &d) {
    const double S = d.S, Nc = d.Nc;
    return Nc * (2.0 * S * S * (0.5 * S + S + 1)) + (Nc - 1) * (2.0 * S * S * S + 2.0 * S * S) + (Nc - 1) * 2.0 * S * S;
}

// which fused-engine variant runs this plan (if any)
static FusedChoice fused_choice(const PlanDev &d) {
    constexpr size_t lim = 227 * 1024 - 2048;  // dynamic part: the 227 KB per-CTA limit includes the static __shared__
    constexpr size_t lim2 = 113 * 1024 - 1024;  // ... of a CTA that shares the SM with a second one
    // the linear cache is appended when it does not cost residency (2 CTAs/SM stay 2) and fits
    auto with_cache = [&](size_t b, int hot) {
        const size_t c = (fused_lin_cache_doubles(d) + 2) * sizeof(double);
        const bool fits = b <= lim2 ? b + c <= lim2 : b + c <= lim;
        return fits ? FusedChoice{true, 0, b + c, hot | 2} : FusedChoice{true, 0, b, hot};
    };
    FusedChoice f{false, 0, 0, 0};
    if (d.flags & YHB_FLAG_STAGED) return f;
    const int need = std::max(1, (d.Wc + 1 + 7) / 8);
    if (d.bs_ok && !getenv("YHB_NO_BS_COMPACT")) {
        // block-sparse plans: the compact variant (core in block-row storage, 256 threads) when three CTAs fit one SM
        constexpr size_t lim3 = (228 * 1024) / 3 - 1024 - 1024;  // per-CTA reserve and the static __shared__
        const size_t b = fused_smem_doubles<kFusedBsNT>(d, true) * sizeof(double);
        if (getenv("YHB_DEBUG"))
            fprintf(stderr, "yhb: compact block-sparse variant needs %zu B of %zu (block rows %d doubles, %d levels, hot tables %d B, "
                    "%d nonlinear pairs)\n", b, lim3, d.bs_total, d.bs_nlev, d.blob_hot_bytes, d.npair_nl);
        if (b <= lim3) {
            const size_t c = (fused_lin_cache_doubles(d) + 2) * sizeof(double);
            return b + c <= lim3 ? FusedChoice{true, kFusedBsNT, b + c, 3} : FusedChoice{true, kFusedBsNT, b, 1};
        }
        // too large for three CTAs per SM: the compact variant with TWO CTAs per SM (192 threads each) instead of the
        // dense-core variant's 32 NT threads, most of which wait at barriers while the block-sparse solve runs
        // (ActiveLoad DiffAmp, 27 nonlinear pairs, 99 KB: 178 k -> 188 k solves/s against k_fused<11> on 352 threads)
        if (b <= lim2 && !getenv("YHB_NO_BS_COMPACT2")) {
            const size_t c = (fused_lin_cache_doubles(d) + 2) * sizeof(double);
            return b + c <= lim2 ? FusedChoice{true, kFusedBsNT, b + c, 3} : FusedChoice{true, kFusedBsNT, b, 1};
        }
    }
    for (int hot = 1; hot >= 0; --hot) {
#define YHB_CASE(t)                                                                 \
    if (need <= t) {                                                                \
        size_t b = fused_smem_doubles<t>(d, hot != 0) * sizeof(double);             \
        if (b <= lim) { FusedChoice c = with_cache(b, hot); c.NT = t; return c; }   \
    }
        YHB_FOR_NT(YHB_CASE)
#undef YHB_CASE
        size_t b = fused_smem_doubles<0>(d, hot != 0) * sizeof(double);
        if (b <= lim) return with_cache(b, hot);
    }
    return f;
}

// ---- optional per-kernel timing ---------------------------------------------------------------------
enum { PK_PREPARE = 0, PK_DC, PK_INIT, PK_EVAL, PK_ASSEMBLE, PK_LU, PK_FINISH, PK_FUSED };
struct Prof {
    bool on = false;
    std::mutex mu;
    long long launches[YHB_PROF_KINDS] = {0};
    double ms[YHB_PROF_KINDS] = {0};
    struct Span { int kind; cudaEvent_t a, b; };
    std::vector<Span> open;
    std::vector<cudaEvent_t> pool;
    cudaEvent_t get() {
        if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
        cudaEvent_t e;
        cudaEventCreate(&e);
        return e;
    }
};
static Prof g_prof;

struct ProfScope {  // records an event pair around the launches made while it is alive
    int kind;
    cudaStream_t st;
    cudaEvent_t a = nullptr, b = nullptr;
    ProfScope(int k, cudaStream_t s) : kind(k), st(s) {
        if (!g_prof.on) return;
        std::lock_guard<std::mutex> lk(g_prof.mu);
        a = g_prof.get();
        b = g_prof.get();
        cudaEventRecord(a, st);
    }
    ~ProfScope() {
        if (!a) return;
        cudaEventRecord(b, st);
        std::lock_guard<std::mutex> lk(g_prof.mu);
        g_prof.open.push_back({kind, a, b});
    }
};

template <typename T>
static const T *upload(Plan *pl, const std::vector<T> &v) {
    void *d = nullptr;
    size_t bytes = std::max<size_t>(v.size(), 1) * sizeof(T);
    if (cudaMalloc(&d, bytes) != cudaSuccess) return nullptr;
    if (!v.empty()) cudaMemcpy(d, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice);
    pl->allocs.push_back(d);
    return reinterpret_cast<const T *>(d);
}

struct TermList {
    std::vector<std::vector<Term>> rows;
    void add(int row, int idx, int sgn) {
        if ((int)rows.size() <= row) rows.resize(row + 1);
        rows[row].push_back(Term{idx, sgn});
    }
    void flatten(int nrows, std::vector<int> &ptr, std::vector<Term> &out) const {
        ptr.assign(nrows + 1, 0);
        out.clear();
        for (int r = 0; r < nrows; ++r) {
            if (r < (int)rows.size()) out.insert(out.end(), rows[r].begin(), rows[r].end());
            ptr[r + 1] = (int)out.size();
        }
    }
};

// Caller's yhb_result -> a full-size copy: only struct_size bytes are read, later fields are NULL.
static int load_result(const yhb_result *in, yhb_result &out) {
    memset(&out, 0, sizeof(out));
    if (!in) { set_error("null result"); return -1; }
    const size_t n = in->struct_size;
    if (n < sizeof(size_t) + sizeof(void *) || n > 4096 || (n % sizeof(void *)) != 0) {
        set_error("yhb_result.struct_size not set (must be sizeof(yhb_result) of the header the caller was built with)");
        return -1;
    }
    memcpy(&out, in, std::min(n, sizeof(out)));
    out.struct_size = sizeof(out);
    return 0;
}

}  // namespace yhb

using namespace yhb;

struct yhb_plan : public yhb::Plan {
    std::vector<int> nl_kind, nl_index;
    const int *nl_kind_dev, *nl_index_dev;
    // one-warp DC engine tables (build_dc_tables); dc_warp = the circuit qualifies
    const int *dc_ent_dev, *dc_lane_dev;
    int dc_warp, dc_ent_ints;
    int *pinned_counts;
    cudaStream_t own_stream;
    // yhb_run_host buffers
    void *hbuf;
    size_t hbuf_bytes;
    std::mutex mu;
    std::vector<int> bs_host;  // block-sparse structure per core node: nlater, block rows below, 0, 0 (yhb_plan_flops)
    // yhb_solve_cold: high-priority side stream of the DC kernel and the fork / join events
    cudaStream_t dc_stream;
    cudaEvent_t ev_fork, ev_join;
};

extern "C" const char *yhb_last_error(void) { return g_err.c_str(); }
extern "C" int yhb_version(void) { return YHB_VERSION; }
extern "C" int yhb_set_device(int32_t device) {
    YHB_CUDA(cudaSetDevice(device));
    return 0;
}

extern "C" int yhb_profile_enable(int32_t on) {
    g_prof.on = on != 0;
    return 0;
}

static void prof_resolve() {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (auto &s : g_prof.open) {
        cudaEventSynchronize(s.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, s.a, s.b) == cudaSuccess) {
            g_prof.ms[s.kind] += ms;
            g_prof.launches[s.kind] += 1;
        }
        g_prof.pool.push_back(s.a);
        g_prof.pool.push_back(s.b);
    }
    g_prof.open.clear();
}

extern "C" int yhb_profile_reset(void) {
    prof_resolve();
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (int i = 0; i < YHB_PROF_KINDS; ++i) { g_prof.launches[i] = 0; g_prof.ms[i] = 0.; }
    return 0;
}

extern "C" int yhb_profile_read(int64_t *launches, double *ms) {
    prof_resolve();
    std::lock_guard<std::mutex> lk(g_prof.mu);
    for (int i = 0; i < YHB_PROF_KINDS; ++i) {
        if (launches) launches[i] = g_prof.launches[i];
        if (ms) ms[i] = g_prof.ms[i];
    }
    return 0;
}

// Plan-time tables of the one-warp DC engine (k_dc_warp): the role of every lane in the device evaluation (a BJT
// takes two lanes, one per junction) and, per entry of [A | z], its contributions in the order dc_apply_stamps
// adds them (netlist order, then stamp order within the device).
static bool build_dc_tables(yhb_plan *pl) {
    const PlanDev &d = pl->d;
    const int M = d.N + pl->num_inductors;
    const int nnl = (int)pl->nl_kind.size();
    int lanes = 0;
    for (int t = 0; t < nnl; ++t) lanes += pl->nl_kind[t] == 1 ? 2 : 1;
    pl->dc_warp = M <= 24 && lanes <= 32;
    pl->dc_ent_dev = pl->dc_lane_dev = nullptr;
    pl->dc_ent_ints = 0;
    if (!pl->dc_warp) return true;
    std::vector<int> lane_tab(32 * kDcLaneInts, 0);
    for (int l = 0; l < 32; ++l) { lane_tab[l * kDcLaneInts] = -1; lane_tab[l * kDcLaneInts + 7] = l; }
    std::vector<std::vector<int>> ent((size_t)M * M + M);
    auto addA = [&](int r, int c, int t, int slot, int neg) {
        if (r > 0 && c > 0) ent[(size_t)(r - 1) * M + (c - 1)].push_back(((t * kDcOut + slot) << 1) | neg);
    };
    auto addz = [&](int r, int t, int slot, int neg) {
        if (r > 0) ent[(size_t)M * M + (r - 1)].push_back(((t * kDcOut + slot) << 1) | neg);
    };
    int l = 0;
    for (int t = 0; t < nnl; ++t) {
        const int kind = pl->nl_kind[t], idx = pl->nl_index[t];
        if (kind == 1) {
            const int B = pl->bjt_nodes[3 * idx], C = pl->bjt_nodes[3 * idx + 1], E = pl->bjt_nodes[3 * idx + 2];
            for (int h = 0; h < 2; ++h, ++l) {
                int *q = &lane_tab[l * kDcLaneInts];
                q[0] = 1; q[1] = h; q[2] = t; q[3] = idx; q[4] = B; q[5] = C; q[6] = E; q[7] = h ? l - 1 : l + 1;
            }
            const int rc[3] = {B, C, E};
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 3; ++j) addA(rc[i], rc[j], t, 3 * i + j, 0);
            for (int i = 0; i < 3; ++i) addz(rc[i], t, 9 + i, 0);
        } else {
            const int *nodes = kind == 0 ? &pl->diode_nodes[2 * idx] : &pl->cubic_nodes[2 * idx];
            const int n1 = nodes[0], n2 = nodes[1];
            int *q = &lane_tab[l * kDcLaneInts];
            q[0] = kind; q[1] = 0; q[2] = t; q[3] = idx; q[4] = n1; q[5] = n2; q[6] = 0; q[7] = l;
            ++l;
            addA(n1, n1, t, 0, 0);  // dc_stamp2
            addA(n2, n2, t, 0, 0);
            if (n1 > 0 && n2 > 0) {
                addA(n1, n2, t, 0, 1);
                addA(n2, n1, t, 0, 1);
            }
            addz(n1, t, 1, 1);
            addz(n2, t, 1, 0);
        }
    }
    const int NE = (int)ent.size(), LD = (M + 1) | 1;
    std::vector<int> tab(NE + 1, 0), dst(NE), items;
    for (int e = 0; e < NE; ++e) {
        items.insert(items.end(), ent[e].begin(), ent[e].end());
        tab[e + 1] = (int)items.size();
        if (e < M * M) {
            const int r = e / M, c = e % M;
            dst[e] = (r * LD + c) | ((r == c && r < d.N) ? 1 << 16 : 0);
        } else {
            dst[e] = (e - M * M) * LD + M;
        }
    }
    tab.insert(tab.end(), dst.begin(), dst.end());
    tab.insert(tab.end(), items.begin(), items.end());
    pl->dc_ent_ints = (int)tab.size();
    pl->dc_ent_dev = upload(pl, tab);
    pl->dc_lane_dev = upload(pl, lane_tab);
    return pl->dc_ent_dev && pl->dc_lane_dev;
}

extern "C" int yhb_plan_create(const yhb_plan_desc *desc, yhb_plan **out) {
    if (!desc || !out) { set_error("null argument"); return -1; }
    if (desc->num_nodes < 1 || (desc->num_tones != 1 && desc->num_tones != 2) || desc->K1 < 1 ||
        (desc->num_tones == 2 && desc->K2 < 1)) {
        set_error("bad plan descriptor: need N >= 1, one or two tones, K >= 1");
        return -1;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        set_error("no CUDA device: libyhb has no CPU fallback");
        return -3;
    }
    yhb_plan *pl = new (std::nothrow) yhb_plan();
    if (!pl) { set_error("out of memory"); return -1; }
    PlanDev &d = pl->d;
    const int N = desc->num_nodes;
    d.N = N;
    d.ntones = desc->num_tones;
    d.K2 = desc->num_tones == 2 ? desc->K2 : 0;
    if (desc->num_tones == 1) {  // :247-251
        d.K = desc->K1;
        for (int k = 0; k <= d.K; ++k) { pl->bin_k1.push_back(k); pl->bin_k2.push_back(0); }
    } else {  // :252-273
        d.K = (2 * desc->K2 + 1) * desc->K1 + desc->K2;
        for (int k1 = 0; k1 <= desc->K1; ++k1)
            for (int k2 = -desc->K2; k2 <= desc->K2; ++k2) {
                if (k1 == 0 && k2 < 0) continue;
                pl->bin_k1.push_back(k1);
                pl->bin_k2.push_back(k2);
            }
    }
    d.S = 2 * d.K + 1;
    d.W = d.S * N;
    d.nlin = desc->num_lin;
    d.nsrc = desc->num_src;
    d.nd = desc->num_diode;
    d.nb = desc->num_bjt;
    d.nc = desc->num_cubic;
    d.flags = desc->flags;
    const int S = d.S, K = d.K;

    auto check_nodes = [&](const int32_t *a, int cnt) {
        for (int i = 0; i < cnt; ++i)
            if (a[i] < 0 || a[i] > N) return false;
        return true;
    };
    if (!check_nodes(desc->lin_nodes, 4 * d.nlin) || !check_nodes(desc->src_nodes, 2 * d.nsrc) ||
        !check_nodes(desc->diode_nodes, 2 * d.nd) || !check_nodes(desc->bjt_nodes, 3 * d.nb) ||
        !check_nodes(desc->cubic_nodes, 2 * d.nc)) {
        set_error("node index out of range");
        delete pl;
        return -1;
    }
    auto check_range = [&](const int32_t *a, int cnt, int lo, int hi) {
        for (int i = 0; i < cnt; ++i)
            if (a[i] < lo || a[i] > hi) return false;
        return true;
    };
    if ((d.nlin && !desc->lin_kind) || (d.nsrc && !desc->src_kind) ||
        !check_range(desc->lin_kind, d.nlin, YHB_LIN_RESISTOR, YHB_LIN_IHF) ||
        !check_range(desc->src_kind, d.nsrc, YHB_SRC_DC, YHB_SRC_DC_ONLY)) {
        set_error("linear element / source kind out of range");
        delete pl;
        return -1;
    }
    if (desc->nl_kind && desc->nl_index) {
        const int cnt[3] = {d.nd, d.nb, d.nc};
        std::vector<char> seen[3];
        for (int k = 0; k < 3; ++k) seen[k].assign(cnt[k], 0);
        bool okk = true;
        for (int t = 0; okk && t < d.nd + d.nb + d.nc; ++t) {
            const int k = desc->nl_kind[t], i = desc->nl_index[t];
            okk = k >= 0 && k <= 2 && i >= 0 && i < cnt[k] && !seen[k][i];
            if (okk) seen[k][i] = 1;
        }
        if (!okk) {
            set_error("nl_kind / nl_index must list every nonlinear device exactly once");
            delete pl;
            return -1;
        }
    }
    pl->lin_kind.assign(desc->lin_kind, desc->lin_kind + d.nlin);
    pl->lin_nodes.assign(desc->lin_nodes, desc->lin_nodes + 4 * d.nlin);
    pl->src_kind.assign(desc->src_kind, desc->src_kind + d.nsrc);
    pl->src_nodes.assign(desc->src_nodes, desc->src_nodes + 2 * d.nsrc);
    pl->diode_nodes.assign(desc->diode_nodes, desc->diode_nodes + 2 * d.nd);
    pl->bjt_nodes.assign(desc->bjt_nodes, desc->bjt_nodes + 3 * d.nb);
    pl->cubic_nodes.assign(desc->cubic_nodes, desc->cubic_nodes + 2 * d.nc);
    pl->num_inductors = 0;
    for (int e = 0; e < d.nlin; ++e) pl->num_inductors += pl->lin_kind[e] == YHB_LIN_INDUCTOR;

    // nonlinear devices in netlist order (kind 0 diode, 1 bjt, 2 cubic)
    const int nnl = d.nd + d.nb + d.nc;
    if (desc->nl_kind && desc->nl_index) {
        pl->nl_kind.assign(desc->nl_kind, desc->nl_kind + nnl);
        pl->nl_index.assign(desc->nl_index, desc->nl_index + nnl);
    } else {
        for (int i = 0; i < d.nd; ++i) { pl->nl_kind.push_back(0); pl->nl_index.push_back(i); }
        for (int i = 0; i < d.nc; ++i) { pl->nl_kind.push_back(2); pl->nl_index.push_back(i); }
        for (int i = 0; i < d.nb; ++i) { pl->nl_kind.push_back(1); pl->nl_index.push_back(i); }
    }

    // waveform slots: device order inside the evaluation kernel is diodes, cubics, bjts
    std::vector<int> wave_base;
    int nw = 0;
    const bool dq = (desc->flags & YHB_FLAG_DIODE_CHARGE) != 0;  // diodes: g, I and, with the charge extension, c, q
    for (int i = 0; i < d.nd; ++i) { wave_base.push_back(nw); nw += dq ? 4 : 2; }
    for (int i = 0; i < d.nc; ++i) { wave_base.push_back(nw); nw += 2; }
    for (int i = 0; i < d.nb; ++i) { wave_base.push_back(nw); nw += kBjtWaves; }
    d.nwave = nw;

    // ---- J block structure and stamp lists (:459-475 diode/cubic, :524-570 BJT), netlist order
    std::map<std::pair<int, int>, int> pair_id;
    std::vector<std::pair<int, int>> pairs;
    auto get_pair = [&](int n, int m) {  // reference node numbers, both > 0
        auto key = std::make_pair(n - 1, m - 1);
        auto it = pair_id.find(key);
        if (it != pair_id.end()) return it->second;
        int id = (int)pairs.size();
        pair_id[key] = id;
        pairs.push_back(key);
        return id;
    };
    TermList pg, pc, ni, nq, ns, plin;
    auto g_stamp = [&](int n, int m, int wave, int sgn) { if (n > 0 && m > 0) pg.add(get_pair(n, m), wave, sgn); };
    auto c_stamp = [&](int n, int m, int slot, int sgn) { if (n > 0 && m > 0) pc.add(get_pair(n, m), slot, sgn); };
    for (int t = 0; t < nnl; ++t) {
        int kind = pl->nl_kind[t], i = pl->nl_index[t];
        if (kind == 0 || kind == 2) {
            const int *nd = kind == 0 ? &pl->diode_nodes[2 * i] : &pl->cubic_nodes[2 * i];
            int wb = wave_base[kind == 0 ? i : d.nd + i];
            int n1 = nd[0], n2 = nd[1];
            g_stamp(n1, n1, wb, +1);
            if (n1 > 0) ni.add(n1 - 1, wb + 1, +1);
            g_stamp(n2, n2, wb, +1);
            if (n2 > 0) ni.add(n2 - 1, wb + 1, -1);
            g_stamp(n1, n2, wb, -1);
            g_stamp(n2, n1, wb, -1);
            if (kind == 0 && dq) {  // charge extension: capacitance slot 4 nb + i (the BJT slots are 4 q + j), charge wave wb + 3
                const int slot = 4 * d.nb + i;
                c_stamp(n1, n1, slot, +1);
                c_stamp(n2, n2, slot, +1);
                c_stamp(n1, n2, slot, -1);
                c_stamp(n2, n1, slot, -1);
                if (n1 > 0) nq.add(n1 - 1, wb + 3, +1);
                if (n2 > 0) nq.add(n2 - 1, wb + 3, -1);
            }
        } else {
            const int *nd = &pl->bjt_nodes[3 * i];
            int wb = wave_base[d.nd + d.nc + i];
            int B = nd[0], C = nd[1], E = nd[2];
            const int Ib = wb, Ic = wb + 1, Ie = wb + 2, Qbe = wb + 3, Qbc = wb + 4, Qsc = wb + 5;
            const int gmu = wb + 6, gpi = wb + 7, gmf = wb + 8, gmr = wb + 9;
            const int Cbc = 4 * i, Cbe = 4 * i + 1, Cbebc = 4 * i + 2, Csc = 4 * i + 3;
            // :524-530
            g_stamp(B, B, gmu, +1); g_stamp(B, B, gpi, +1);
            c_stamp(B, B, Cbc, +1); c_stamp(B, B, Cbe, +1); c_stamp(B, B, Cbebc, +1);
            if (B > 0) { ni.add(B - 1, Ib, +1); nq.add(B - 1, Qbe, +1); nq.add(B - 1, Qbc, +1); }
            // :532-538
            g_stamp(B, C, gmu, -1);
            g_stamp(C, B, gmu, -1); g_stamp(C, B, gmf, +1); g_stamp(C, B, gmr, +1);
            c_stamp(B, C, Cbc, -1); c_stamp(B, C, Cbebc, -1);
            c_stamp(C, B, Cbc, -1);
            // :540-546
            g_stamp(B, E, gpi, -1);
            g_stamp(E, B, gpi, -1); g_stamp(E, B, gmf, -1); g_stamp(E, B, gmr, -1);
            c_stamp(B, E, Cbe, -1);
            c_stamp(E, B, Cbe, -1); c_stamp(E, B, Cbebc, -1);
            // :548-554
            g_stamp(C, C, gmu, +1); g_stamp(C, C, gmr, -1);
            c_stamp(C, C, Cbc, +1); c_stamp(C, C, Csc, +1);
            if (C > 0) { ni.add(C - 1, Ic, +1); nq.add(C - 1, Qbc, -1); nq.add(C - 1, Qsc, +1); }
            // :556-562
            g_stamp(C, E, gmf, -1);
            g_stamp(E, C, gmr, +1);
            c_stamp(E, C, Cbebc, +1);
            // :564-570
            g_stamp(E, E, gpi, +1); g_stamp(E, E, gmf, +1);
            c_stamp(E, E, Cbe, +1);
            if (E > 0) { ni.add(E - 1, Ie, -1); nq.add(E - 1, Qbe, -1); }
        }
    }
    d.npair_nl = (int)pairs.size();
    // linear stamps (add_mthb_stamps of each linear class), netlist order
    auto l_stamp = [&](int n, int m, int e, int sgn) { if (n > 0 && m > 0) plin.add(get_pair(n, m), e, sgn); };
    for (int e = 0; e < d.nlin; ++e) {
        const int *nd = &pl->lin_nodes[4 * e];
        if (pl->lin_kind[e] == YHB_LIN_GYRATOR) {  // Gyrator.py:45-73
            l_stamp(nd[0], nd[1], e, +1); l_stamp(nd[1], nd[0], e, -1);
            l_stamp(nd[0], nd[2], e, -1); l_stamp(nd[2], nd[0], e, +1);
            l_stamp(nd[1], nd[3], e, +1); l_stamp(nd[3], nd[1], e, -1);
            l_stamp(nd[2], nd[3], e, -1); l_stamp(nd[3], nd[2], e, +1);
        } else {
            l_stamp(nd[0], nd[0], e, +1);
            l_stamp(nd[1], nd[1], e, +1);
            l_stamp(nd[0], nd[1], e, -1);
            l_stamp(nd[1], nd[0], e, -1);
        }
    }
    d.npair = (int)pairs.size();
    for (int s = 0; s < d.nsrc; ++s) {
        int n1 = pl->src_nodes[2 * s], n2 = pl->src_nodes[2 * s + 1];
        if (n1 > 0) ns.add(n1 - 1, s, +1);
        if (n2 > 0) ns.add(n2 - 1, s, -1);
    }

    // ---- structural peeling of ideal-source rows / columns (gyrator pivots are exactly +-1: no rounding, no
    // growth).  Reference netlists model every voltage source as current source + gyrator
    // (tests/HB_Diode.py:12-13), which leaves rows with a single unknown and unknowns that appear in a single
    // row; eliminating them symbolically shrinks the dense system LAPACK would factor (:620) from W to Wc.
    std::vector<int> unit_sign(d.npair, 0);
    for (int p = 0; p < d.npair; ++p) {
        if (p < d.npair_nl) continue;
        const auto &row = p < (int)plin.rows.size() ? plin.rows[p] : std::vector<Term>();
        if (row.size() == 1 && pl->lin_kind[row[0].idx] == YHB_LIN_GYRATOR) unit_sign[p] = row[0].sgn;
    }
    std::vector<std::vector<int>> row_cols(N), col_rows(N);  // pairs by row / by column
    for (int p = 0; p < d.npair; ++p) {
        row_cols[pairs[p].first].push_back(p);
        col_rows[pairs[p].second].push_back(p);
    }
    std::vector<char> act_row(N, 1), act_col(N, 1), a_known(N, 0);
    std::vector<int> a_row, a_col, a_sgn, b_row, b_col, b_sgn;
    if (!(desc->flags & YHB_FLAG_NO_PEEL)) {
        bool changed = true;
        while (changed) {
            changed = false;
            for (int n = 0; n < N; ++n) {  // row singletons
                if (!act_row[n]) continue;
                int cnt = 0, last = -1;
                bool linear_row = true;  // an ideal-source row carries no nonlinear block
                for (int p : row_cols[n]) {
                    if (p < d.npair_nl) linear_row = false;
                    if (!a_known[pairs[p].second]) { ++cnt; last = p; }
                }
                if (linear_row && cnt == 1 && unit_sign[last] != 0 && act_col[pairs[last].second]) {
                    a_row.push_back(n);
                    a_col.push_back(pairs[last].second);
                    a_sgn.push_back(unit_sign[last]);
                    act_row[n] = 0;
                    act_col[pairs[last].second] = 0;
                    a_known[pairs[last].second] = 1;
                    changed = true;
                }
            }
            for (int m = 0; m < N; ++m) {  // column singletons
                if (!act_col[m]) continue;
                int cnt = 0, last = -1;
                bool linear_col = true;  // the recovered unknown (a source current) feeds no nonlinear block
                for (int p : col_rows[m]) {
                    if (p < d.npair_nl) linear_col = false;
                    if (act_row[pairs[p].first]) { ++cnt; last = p; }
                }
                if (linear_col && cnt == 1 && unit_sign[last] != 0) {
                    b_row.push_back(pairs[last].first);
                    b_col.push_back(m);
                    b_sgn.push_back(unit_sign[last]);
                    act_row[pairs[last].first] = 0;
                    act_col[m] = 0;
                    changed = true;
                }
            }
        }
    }
    std::vector<int> core_rows, core_cols, core_col_of(N, -1);
    for (int n = 0; n < N; ++n) if (act_row[n]) core_rows.push_back(n);
    for (int n = 0; n < N; ++n) if (act_col[n]) { core_col_of[n] = (int)core_cols.size(); core_cols.push_back(n); }
    d.Nc = (int)core_rows.size();
    d.Wc = d.Nc * S;
    // block bandwidth of the core in its row / column orderings (chains of nodes give block-tridiagonal cores)
    {
        int bwn = 0;
        for (int r = 0; r < d.Nc; ++r)
            for (int p : row_cols[core_rows[r]]) {
                int cpos = core_col_of[pairs[p].second];
                if (cpos >= 0) bwn = std::max(bwn, std::abs(cpos - r));
            }
        const int kl = (bwn + 1) * S - 1;
        // worth band storage when the band (3 kl + 1 per row) is well below the dense row
        d.band = (!(desc->flags & YHB_FLAG_NO_BAND) && d.Nc > 0 && (long)3 * kl + 1 < (long)d.Wc / 2) ? kl : 0;
        // block tridiagonal with blocks too large for the one-CTA banded elimination (or on request): block-Thomas
        d.btd = (bwn <= 1 && d.Nc >= 3 && !(desc->flags & YHB_FLAG_NO_BLOCK_THOMAS) &&
                 ((desc->flags & YHB_FLAG_BLOCK_THOMAS) || S >= 128)) ? 1 : 0;
    }
    // ---- block-sparse LU structure of the core (PlanDev::bs_*): ordering by levels of independent minimum-degree
    // nodes, fill, block-row storage map
    std::vector<int> bs_order, bs_pos, bs_level_ptr(1, 0), bs_rowbase, bs_ld, bs_nlater, bs_nblk, bs_rhsoff, bs_coloff,
        bs_blkof, bs_below_ptr(1, 0), bs_below, bs_lev_uniform;
    d.bs_ok = 0;
    d.bs_nlev = 0;
    d.bs_total = 0;
    d.bs_zero = 1;
    {
        const int Nc = d.Nc;
        bool same = Nc >= 2 && Nc <= 16 && S <= 32 && !(desc->flags & YHB_FLAG_NO_BLOCK_SPARSE);
        for (int r = 0; same && r < Nc; ++r) same = core_rows[r] == core_cols[r];
        // pivoting stays inside the diagonal blocks, so every core node needs a conductance of its own on the
        // diagonal: a nonlinear device or a resistor (a node held only by gyrators, capacitors or inductors has a
        // zero or DC-singular diagonal block, e.g. the source rows of an unpeeled system: dense solve)
        for (int r = 0; same && r < Nc; ++r) {
            auto it = pair_id.find(std::make_pair(core_rows[r], core_rows[r]));
            bool own = false;
            if (it != pair_id.end()) {
                const int p = it->second;
                own = p < d.npair_nl;
                if (p < (int)plin.rows.size())
                    for (const Term &t : plin.rows[p]) own = own || pl->lin_kind[t.idx] == YHB_LIN_RESISTOR;
            }
            same = own;
        }
        if (same) {
            std::vector<std::vector<char>> nz(Nc, std::vector<char>(Nc, 0));
            for (int r = 0; r < Nc; ++r) {
                nz[r][r] = 1;
                for (int p : row_cols[core_rows[r]]) {
                    int cpos = core_col_of[pairs[p].second];
                    if (cpos >= 0) nz[r][cpos] = 1;
                }
            }
            int nnz0 = 0;
            for (int r = 0; r < Nc; ++r)
                for (int c = 0; c < Nc; ++c) nnz0 += nz[r][c];
            std::vector<char> gone(Nc, 0);
            bs_pos.assign(Nc, -1);
            while ((int)bs_order.size() < Nc) {
                std::vector<int> deg(Nc, 0);
                int mind = 1 << 30;
                for (int k = 0; k < Nc; ++k) {
                    if (gone[k]) continue;
                    for (int j = 0; j < Nc; ++j)
                        if (j != k && !gone[j] && (nz[k][j] || nz[j][k])) deg[k]++;
                    mind = std::min(mind, deg[k]);
                }
                std::vector<int> cand;
                for (int k = 0; k < Nc; ++k)
                    if (!gone[k] && deg[k] <= mind + 1) cand.push_back(k);
                std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) { return deg[a] < deg[b]; });
                std::vector<int> level;
                for (int k : cand) {
                    bool indep = true;
                    for (int t : level) indep = indep && !nz[k][t] && !nz[t][k];
                    if (indep) level.push_back(k);
                }
                for (int k : level) {
                    bs_pos[k] = (int)bs_order.size();
                    bs_order.push_back(k);
                    gone[k] = 1;
                }
                for (int k : level)  // fill among the remaining neighbours of k
                    for (int i = 0; i < Nc; ++i)
                        if (!gone[i] && nz[i][k])
                            for (int j = 0; j < Nc; ++j)
                                if (!gone[j] && nz[k][j]) nz[i][j] = 1;
                bs_level_ptr.push_back((int)bs_order.size());
            }
            int nnz1 = 0;
            for (int r = 0; r < Nc; ++r)
                for (int c = 0; c < Nc; ++c) nnz1 += nz[r][c];
            bs_rowbase.assign(Nc, 0);
            bs_ld.assign(Nc, 0);
            bs_nlater.assign(Nc, 0);
            bs_nblk.assign(Nc, 0);
            bs_rhsoff.assign(Nc, 0);
            bs_coloff.assign((size_t)Nc * Nc, -1);
            bs_blkof.assign((size_t)Nc * Nc, -1);
            size_t total = 0;
            for (int k = 0; k < Nc; ++k) {
                std::vector<int> later, earlier;
                for (int j = 0; j < Nc; ++j) {
                    if (j == k || !nz[k][j]) continue;
                    (bs_pos[j] > bs_pos[k] ? later : earlier).push_back(j);
                }
                int slot = 0, col = 0;
                auto put = [&](int j) {
                    bs_coloff[(size_t)k * Nc + j] = col;
                    bs_blkof[(size_t)k * Nc + slot] = j;
                    ++slot;
                    col += S;
                };
                put(k);
                for (int j : later) put(j);
                bs_rhsoff[k] = col;
                col += 1;
                for (int j : earlier) put(j);
                bs_nlater[k] = (int)later.size();
                bs_nblk[k] = slot;
                bs_ld[k] = col | 1;  // odd: conflict-free column walks
                bs_rowbase[k] = (int)total;
                total += (size_t)S * bs_ld[k];
            }
            for (int k = 0; k < Nc; ++k) {  // block rows that still hold a block (i, k) when k is eliminated
                for (int i = 0; i < Nc; ++i)
                    if (bs_pos[i] > bs_pos[k] && nz[i][k]) bs_below.push_back(i);
                bs_below_ptr.push_back((int)bs_below.size());
            }
            for (size_t lev = 0; lev + 1 < bs_level_ptr.size(); ++lev) {
                bool uni = true;
                const int k0 = bs_order[bs_level_ptr[lev]];
                for (int t = bs_level_ptr[lev] + 1; t < bs_level_ptr[lev + 1] && uni; ++t) {
                    const int k = bs_order[t];
                    uni = bs_nlater[k] == bs_nlater[k0] &&
                          bs_below_ptr[k + 1] - bs_below_ptr[k] == bs_below_ptr[k0 + 1] - bs_below_ptr[k0];
                    for (int sl = 0; uni && sl < bs_nlater[k0]; ++sl)
                        uni = bs_blkof[(size_t)k * Nc + 1 + sl] == bs_blkof[(size_t)k0 * Nc + 1 + sl];
                    for (int q = 0; uni && q < bs_below_ptr[k0 + 1] - bs_below_ptr[k0]; ++q)
                        uni = bs_below[bs_below_ptr[k] + q] == bs_below[bs_below_ptr[k0] + q];
                }
                bs_lev_uniform.push_back(uni && bs_level_ptr[lev + 1] - bs_level_ptr[lev] > 1 ? 1 : 0);
            }
            const bool sparse = nnz1 < Nc * Nc;
            // never more than the dense matrix of the fused engine
            const bool fits = total <= (size_t)(d.Wc + 1) * (d.Wc + 1);
            if (fits && (sparse || (desc->flags & YHB_FLAG_BLOCK_SPARSE))) {
                d.bs_ok = 1;
                d.bs_nlev = (int)bs_level_ptr.size() - 1;
                d.bs_total = (int)total;
            }
            d.bs_zero = nnz1 != nnz0;  // fill blocks hold no pair: only then the storage needs zeroing before assembly
        }
    }
    if (d.bs_ok)
        for (int k = 0; k < d.Nc; ++k) {
            pl->bs_host.push_back(bs_nlater[k]);
            pl->bs_host.push_back(bs_below_ptr[k + 1] - bs_below_ptr[k]);
            pl->bs_host.push_back(0);
            pl->bs_host.push_back(0);
        }
    d.nA = (int)a_row.size();
    d.nB = (int)b_row.size();
    // Steps that do not read each other's result run side by side: an A step reads the columns fixed by earlier A steps
    // among its row's other entries; a B step (processed in reverse order of discovery) reads the columns recovered by
    // B steps processed before it.  Steps are stable-sorted by level; the B lists are stored in processing order.
    std::vector<int> a_lev(1, 0), b_lev(1, 0);
    {
        auto by_level = [&](std::vector<int> &row, std::vector<int> &col, std::vector<int> &sgn, std::vector<int> &lev_ptr) {
            const int n = (int)row.size();
            std::vector<int> made(N, -1), level(n, 0), perm(n);
            for (int i = 0; i < n; ++i) {
                int lv = 0;
                for (int p : row_cols[row[i]]) {
                    const int m = pairs[p].second;
                    if (m != col[i] && made[m] >= 0) lv = std::max(lv, level[made[m]] + 1);
                }
                level[i] = lv;
                made[col[i]] = i;
                perm[i] = i;
            }
            std::stable_sort(perm.begin(), perm.end(), [&](int x, int y) { return level[x] < level[y]; });
            std::vector<int> r2(n), c2(n), s2(n);
            for (int i = 0; i < n; ++i) { r2[i] = row[perm[i]]; c2[i] = col[perm[i]]; s2[i] = sgn[perm[i]]; }
            row = r2; col = c2; sgn = s2;
            for (int i = 0; i < n; ++i)
                if (i + 1 == n || level[perm[i + 1]] != level[perm[i]]) lev_ptr.push_back(i + 1);
        };
        by_level(a_row, a_col, a_sgn, a_lev);
        std::reverse(b_row.begin(), b_row.end());
        std::reverse(b_col.begin(), b_col.end());
        std::reverse(b_sgn.begin(), b_sgn.end());
        by_level(b_row, b_col, b_sgn, b_lev);
    }
    d.nAlev = (int)a_lev.size() - 1;
    d.nBlev = (int)b_lev.size() - 1;
    std::vector<int> a_ptr(1, 0), a_pair, a_m, b_ptr(1, 0), b_pair, b_m, r_ptr(1, 0), r_pair, r_m;
    for (int a = 0; a < d.nA; ++a) {  // the other entries of an A row (all A-known by construction)
        for (int p : row_cols[a_row[a]])
            if (pairs[p].second != a_col[a]) { a_pair.push_back(p); a_m.push_back(pairs[p].second); }
        a_ptr.push_back((int)a_pair.size());
    }
    for (int b = 0; b < d.nB; ++b) {
        for (int p : row_cols[b_row[b]])
            if (pairs[p].second != b_col[b]) { b_pair.push_back(p); b_m.push_back(pairs[p].second); }
        b_ptr.push_back((int)b_pair.size());
    }
    for (int r = 0; r < d.Nc; ++r) {  // entries of a core row in A-known columns move to the right-hand side
        for (int p : row_cols[core_rows[r]])
            if (a_known[pairs[p].second]) { r_pair.push_back(p); r_m.push_back(pairs[p].second); }
        r_ptr.push_back((int)r_pair.size());
    }

    std::vector<int> pair_n, pair_m, pair_of((size_t)N * N, -1);
    for (int p = 0; p < d.npair; ++p) {
        pair_n.push_back(pairs[p].first);
        pair_m.push_back(pairs[p].second);
        pair_of[(size_t)pairs[p].first * N + pairs[p].second] = p;
    }
    std::vector<int> row_ptr(N + 1, 0), row_pairs;
    for (int n = 0; n < N; ++n) {
        for (int m = 0; m < N; ++m)
            if (pair_of[(size_t)n * N + m] >= 0) row_pairs.push_back(pair_of[(size_t)n * N + m]);
        row_ptr[n + 1] = (int)row_pairs.size();
    }
    // DFT tables: caller's (reference-identical numpy tables) or analytic (:294-302)
    std::vector<double> idft((size_t)S * S), dft((size_t)S * S);
    if (desc->idft && desc->dft) {
        std::copy(desc->idft, desc->idft + (size_t)S * S, idft.begin());
        std::copy(desc->dft, desc->dft + (size_t)S * S, dft.begin());
    } else {
        const double pi = 3.141592653589793;
        for (int s = 0; s < S; ++s) {
            idft[(size_t)s * S] = 1.0;
            dft[s] = 1.0 / S;
            for (int k = 1; k <= K; ++k) {
                double ang = 2 * pi * (double)(((long long)k * s) % S) / S;
                idft[(size_t)s * S + 2 * k - 1] = std::cos(ang);
                idft[(size_t)s * S + 2 * k] = std::sin(ang);
                dft[(size_t)(2 * k - 1) * S + s] = 2.0 / S * std::cos(ang);
                dft[(size_t)(2 * k) * S + s] = 2.0 / S * std::sin(ang);
            }
        }
    }

    // All plan tables live in ONE device blob; the tables the Newton loop reads ("hot": transforms, stamp lists,
    // peeling lists) come first so that the fused engine can stage that prefix in shared memory (k_fused).
    bool ok = true;
    std::vector<char> blob;
    std::vector<std::pair<const void **, size_t>> fix;
    auto stage = [&](const void **field, const void *data, size_t bytes) {
        size_t off = align_up(blob.size(), 16);
        blob.resize(off + std::max<size_t>(bytes, 16));
        if (bytes) memcpy(blob.data() + off, data, bytes);
        fix.push_back({field, off});
    };
#define UP(field, vec) stage(reinterpret_cast<const void **>(&d.field), (vec).data(), (vec).size() * sizeof((vec)[0]))
    UP(idft, idft);
    UP(dft, dft);
    UP(diode_nodes, pl->diode_nodes);
    UP(bjt_nodes, pl->bjt_nodes);
    UP(cubic_nodes, pl->cubic_nodes);
    UP(dev_wave_base, wave_base);
    UP(pair_n, pair_n);
    UP(pair_m, pair_m);
    UP(pair_of, pair_of);
    UP(row_ptr, row_ptr);
    UP(row_pairs, row_pairs);
    std::vector<int> ptr_pg, ptr_pc, ptr_ni, ptr_nq, ptr_ns, ptr_plin;
    std::vector<Term> t_pg, t_pc, t_ni, t_nq, t_ns, t_plin;
    pg.flatten(d.npair_nl, ptr_pg, t_pg); UP(pg_ptr, ptr_pg); UP(pg, t_pg);
    pc.flatten(d.npair_nl, ptr_pc, t_pc); UP(pc_ptr, ptr_pc); UP(pc, t_pc);
    ni.flatten(N, ptr_ni, t_ni); UP(ni_ptr, ptr_ni); UP(ni, t_ni);
    nq.flatten(N, ptr_nq, t_nq); UP(nq_ptr, ptr_nq); UP(nq, t_nq);
    // core node blocks that hold a pair: (row node, column node, pair), the work list of the core assembly
    std::vector<int> cblk;
    for (int r = 0; r < d.Nc; ++r)
        for (int c = 0; c < d.Nc; ++c) {
            auto it = pair_id.find(std::make_pair(core_rows[r], core_cols[c]));
            if (it != pair_id.end()) { cblk.push_back(r); cblk.push_back(c); cblk.push_back(it->second); }
        }
    d.ncblk = (int)cblk.size() / 3;
    UP(cblk, cblk);
    UP(core_rows, core_rows); UP(core_cols, core_cols); UP(core_col_of, core_col_of);
    UP(a_row, a_row); UP(a_col, a_col); UP(a_sgn, a_sgn); UP(a_ptr, a_ptr); UP(a_pair, a_pair); UP(a_m, a_m);
    UP(b_row, b_row); UP(b_col, b_col); UP(b_sgn, b_sgn); UP(b_ptr, b_ptr); UP(b_pair, b_pair); UP(b_m, b_m);
    UP(a_lev, a_lev); UP(b_lev, b_lev);
    UP(r_ptr, r_ptr); UP(r_pair, r_pair); UP(r_m, r_m);
    UP(bs_order, bs_order); UP(bs_pos, bs_pos); UP(bs_level_ptr, bs_level_ptr); UP(bs_rowbase, bs_rowbase);
    UP(bs_ld, bs_ld); UP(bs_nlater, bs_nlater); UP(bs_nblk, bs_nblk); UP(bs_rhsoff, bs_rhsoff);
    UP(bs_coloff, bs_coloff); UP(bs_blkof, bs_blkof); UP(bs_below_ptr, bs_below_ptr); UP(bs_below, bs_below);
    UP(bs_lev_uniform, bs_lev_uniform);
    d.blob_hot_bytes = (int)align_up(blob.size(), 16);
    // k_prepare / k_dc only
    UP(bin_k1, pl->bin_k1);
    UP(bin_k2, pl->bin_k2);
    UP(lin_kind, pl->lin_kind);
    UP(src_kind, pl->src_kind);
    UP(src_nodes, pl->src_nodes);
    plin.flatten(d.npair, ptr_plin, t_plin); UP(plin_ptr, ptr_plin); UP(plin, t_plin);
    ns.flatten(N, ptr_ns, t_ns); UP(ns_ptr, ptr_ns); UP(ns, t_ns);
#undef UP
    {
        void *dblob = nullptr;
        ok = cudaMalloc(&dblob, align_up(blob.size(), 256)) == cudaSuccess;
        if (ok) {
            pl->allocs.push_back(dblob);
            ok = cudaMemcpy(dblob, blob.data(), blob.size(), cudaMemcpyHostToDevice) == cudaSuccess;
            d.blob = reinterpret_cast<const char *>(dblob);
            for (auto &f : fix) *f.first = d.blob + f.second;
        }
    }
    pl->lin_nodes_dev = upload(pl, pl->lin_nodes);
    ok = ok && build_dc_tables(pl);
    pl->nl_kind_dev = upload(pl, pl->nl_kind);
    pl->nl_index_dev = upload(pl, pl->nl_index);
    ok = ok && pl->lin_nodes_dev && pl->nl_kind_dev && pl->nl_index_dev;
    pl->pinned_counts = nullptr;
    pl->own_stream = nullptr;
    pl->dc_stream = nullptr;
    pl->ev_fork = pl->ev_join = nullptr;
    pl->hbuf = nullptr;
    pl->hbuf_bytes = 0;
    if (ok) ok = cudaHostAlloc((void **)&pl->pinned_counts, 4 * sizeof(int), cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        set_error(std::string("plan upload failed: ") + cudaGetErrorString(cudaGetLastError()));
        yhb_plan_destroy(pl);
        return -2;
    }
    *out = pl;
    return 0;
}

extern "C" void yhb_plan_destroy(yhb_plan *pl) {
    if (!pl) return;
    for (void *p : pl->allocs) cudaFree(p);
    if (pl->pinned_counts) cudaFreeHost(pl->pinned_counts);
    if (pl->hbuf) cudaFree(pl->hbuf);
    if (pl->own_stream) cudaStreamDestroy(pl->own_stream);
    if (pl->dc_stream) cudaStreamDestroy(pl->dc_stream);
    if (pl->ev_fork) cudaEventDestroy(pl->ev_fork);
    if (pl->ev_join) cudaEventDestroy(pl->ev_join);
    delete pl;
}

extern "C" int yhb_plan_dims(const yhb_plan *pl, int32_t *N, int32_t *K, int32_t *S, int32_t *W) {
    if (!pl) { set_error("null plan"); return -1; }
    if (N) *N = pl->d.N;
    if (K) *K = pl->d.K;
    if (S) *S = pl->d.S;
    if (W) *W = pl->d.W;
    return 0;
}

extern "C" int yhb_plan_band(const yhb_plan *pl) { return pl ? pl->d.band : -1; }

extern "C" int yhb_plan_engine(const yhb_plan *pl) {
    if (!pl) { set_error("null plan"); return -1; }
    if (fused_choice(pl->d).ok) return 0;
    if (pl->d.btd) return 3;
    return pl->d.band > 0 ? 2 : 1;
}

extern "C" int yhb_plan_block_sparse(const yhb_plan *pl, int32_t *doubles) {
    if (!pl) { set_error("null plan"); return -1; }
    const bool on = pl->d.bs_ok && fused_choice(pl->d).ok;
    if (doubles) *doubles = on ? pl->d.bs_total : 0;
    return on ? pl->d.bs_nlev : 0;
}

namespace { int gjp_pivots(const PlanDev &d); }
// Floating-point operations of ONE Newton-step linear solve: what the reference hands to LAPACK, a dense
// factorisation of the peeled core, and what the engine chosen for this plan really executes (multiply-adds x 2 of the
// factorisation + substitution loops; transforms, assembly and device evaluation are not counted).
extern "C" int yhb_plan_flops(const yhb_plan *pl, double *lu_reference, double *lu_core_dense, double *lu_executed) {
    if (!pl) { set_error("null plan"); return -1; }
    const PlanDev &d = pl->d;
    const double W = d.W, Wc = d.Wc, S = d.S;
    if (lu_reference) *lu_reference = (2.0 / 3.0) * W * W * W + 2.0 * W * W;  // dgetrf + dgetrs of W x W
    if (lu_core_dense) *lu_core_dense = (2.0 / 3.0) * Wc * Wc * Wc + 2.0 * Wc * Wc;
    if (!lu_executed) return 0;
    const FusedChoice fc = fused_choice(d);
    double ex = 0.;
    if (fc.ok && d.bs_ok && pl->bs_host.size() == (size_t)4 * d.Nc) {
        // bs_factor_row: S pivots, S - 1 other rows, columns right of the pivot up to nact; bs_schur: S multiply-adds
        // per updated element; bs_back: S per later block and row
        for (int k = 0; k < d.Nc; ++k) {
            const double nlater = pl->bs_host[4 * k], nbelow = pl->bs_host[4 * k + 1];
            const double nact = S * (1 + nlater) + 1;
            for (int c = 0; c < d.S; ++c) ex += 2.0 * (S - 1) * (nact - c - 1);
            ex += (nact - S) * S;                                // rows scaled to unit pivots
            ex += 2.0 * nbelow * S * (S * nlater + 1) * S;       // Schur updates of the block rows below
            ex += 2.0 * nlater * S * S;                          // back substitution
        }
    } else if (fc.ok && fc.NT > 0) {
        // gj_solve_mma: per 4-column panel, every tile column right of (and including) the owner gets a rank-4 update
        // over all NT row tiles (8 x 8 x 4 multiply-adds each) + the panel factorisation in registers
        const int NT = fc.NT, npanel = (d.Wc + 3) / 4;
        for (int kp = 0; kp < npanel; ++kp) ex += 2.0 * (NT - (kp >> 1)) * NT * 256.0 + 2.0 * 8 * NT * 6;
    } else if (d.btd) {
        ex = btd_flops(d);
    } else if (gjp_pivots(d) > 0) {
        ex = Wc * Wc * Wc + 2.0 * Wc * Wc;  // panel Gauss-Jordan: full-height rank-16 updates of the columns right of the panel
    } else if (d.band > 0) {
        const double kl = d.band;
        ex = Wc * (2.0 * kl * (2 * kl + 1) + 4.0 * kl);
    } else {
        ex = (2.0 / 3.0) * Wc * Wc * Wc + 2.0 * Wc * Wc;
    }
    *lu_executed = ex;
    return 0;
}

extern "C" int yhb_plan_core(const yhb_plan *pl, int32_t *Nc, int32_t *Wc, int32_t *fused) {
    if (!pl) { set_error("null plan"); return -1; }
    if (Nc) *Nc = pl->d.Nc;
    if (Wc) *Wc = pl->d.Wc;
    if (fused) *fused = fused_choice(pl->d).ok;
    return 0;
}

extern "C" int yhb_plan_freqs(const yhb_plan *pl, double f1, double f2, double *freqs) {
    if (!pl || !freqs) { set_error("null argument"); return -1; }
    for (int k = 0; k <= pl->d.K; ++k)
        freqs[k] = pl->d.ntones == 1 ? f1 * (double)k : (double)pl->bin_k1[k] * f1 + (double)pl->bin_k2[k] * f2;
    return 0;
}

// ---- workspace ---------------------------------------------------------------------------------
namespace {

struct Carver {
    char *base;
    size_t off;
    template <typename T>
    T *take(size_t count) {
        off = align_up(off);
        T *p = base ? reinterpret_cast<T *>(base + off) : nullptr;
        off += count * sizeof(T);
        return p;
    }
};


// pivots per system of the panel Gauss-Jordan solver for this plan's staged engine: S for the block rows of the
// block-Thomas engine, Wc for a dense core large enough to be worth the per-panel launches, 0 = one-CTA elimination
constexpr int kGjpMinCore = 128;
int gjp_pivots(const PlanDev &d) {
    if (d.btd) return d.S;
    if (d.band > 0 || d.Wc < kGjpMinCore) return 0;
    return d.Wc;
}
int gjp_dense_ld(int Wc) { return (gjp_npc(Wc) + 2) & ~1; }  // [J | pad to whole panels | rhs | pad]: even

struct Layout {
    Workspace ws;
    DiodeDerived *dd;
    BjtDerived *bd;
    size_t bytes;
    bool eval_smem;
    bool fused;
    FusedChoice fc;
};

Layout carve(const yhb_plan *pl, int B, void *base) {
    const PlanDev &d = pl->d;
    Layout L;
    Carver c{reinterpret_cast<char *>(base), 0};
    Workspace &ws = L.ws;
    const size_t W = d.W, S = d.S, K1 = d.K + 1;
    ws.batch = B;
    ws.stage_first = 1;
    ws.ctl = c.take<Ctl>(B);
    ws.V = c.take<double>(B * W);
    ws.Vlevel = c.take<double>(B * W);
    ws.dV = c.take<double>(B * W);
    ws.vtprev = c.take<double>(B * W);
    ws.F = c.take<double>(B * W);
    ws.cacc = c.take<double>((size_t)B * std::max(d.nb, 1) * 4 * S);
    ws.spec = c.take<double>((size_t)B * std::max(d.npair_nl, 1) * 2 * S);
    ws.ylin = c.take<double>((size_t)B * d.npair * K1 * 2);
    ws.omega = c.take<double>(B * K1);
    ws.negbin = c.take<int>(B * K1);
    ws.Is = c.take<double>(B * W);
    const size_t esd = eval_scratch_doubles(d);
    L.eval_smem = esd * sizeof(double) <= kMaxEvalSmem;
    ws.scratch_stride = L.eval_smem ? 0 : esd;
    ws.scratch = c.take<double>(L.eval_smem ? 1 : (size_t)B * esd);
    L.fc = fused_choice(d);
    L.fused = L.fc.ok;
    const int gjn = L.fused ? 0 : gjp_pivots(d);
    ws.j_stride = L.fused ? 0
                  : d.btd ? (size_t)d.Nc * btd_row_doubles(d.S)
                          : (d.band > 0 ? (size_t)d.Wc * band_ld(d.band)
                                        : (size_t)d.Wc * (gjn ? gjp_dense_ld(d.Wc) : d.Wc));
    ws.J = c.take<double>(L.fused ? 1 : (size_t)B * ws.j_stride);
    ws.x_stride = W + newton_scratch_doubles(d);
    ws.xstep = c.take<double>(L.fused ? 1 : B * ws.x_stride);
    ws.rhs = c.take<double>(L.fused ? 1 : (size_t)B * (d.Wc + 1));
    ws.pairwav = c.take<double>((size_t)B * std::max(d.npair_nl, 1) * 2 * S);
    ws.gj_n = gjn;
    ws.gj_lpad = (gjn + 7) & ~7;
    const size_t Bs = (size_t)B * (d.btd ? 2 : 1);  // the block-Thomas engine runs two systems per circuit (both chain ends)
    ws.gj_L = c.take<double>(gjn ? Bs * kGjpNB * ws.gj_lpad : 1);
    ws.gj_ld = !gjn ? 0 : (d.btd ? btd_ldm(d.S) : gjp_dense_ld(d.Wc));
    ws.gj_U = c.take<double>(gjn ? Bs * kGjpNB * ws.gj_ld : 1);
    ws.gj_T = c.take<double>(Bs * kGjpNB * kGjpNB);
    ws.gj_piv = c.take<double>(gjn ? Bs * gjn : 1);
    ws.gj_rowof = c.take<int>(gjn ? Bs * gjn : 1);
    ws.gj_used = c.take<int>(gjn ? Bs * gjn : 1);
    ws.gj_pp = c.take<int>(Bs * kGjpNB);
    ws.gj_fail = c.take<int>(Bs);
    ws.Inl = nullptr;
    ws.Ic = nullptr;
    ws.skip = nullptr;
    ws.sens = nullptr;
    ws.dylin = nullptr;
    ws.sens_src_n1 = ws.sens_src_n2 = 0;
    ws.osc_state = c.take<OscState>(B);
    ws.osc_Vacc = c.take<double>(B * W);
    ws.osc_Inl = c.take<double>(B * W);
    ws.osc_sens = c.take<double>(2 * B * W);
    ws.osc_dylin = c.take<double>((size_t)B * d.npair * K1);
    ws.osc_skip = c.take<int>(B);
    ws.osc_live = c.take<int>(4);
    ws.list[0] = c.take<int>(B);
    ws.list[1] = c.take<int>(B);
    ws.count = c.take<int>(4);
    ws.level_iters = c.take<int32_t>((size_t)B * YHB_MAX_LEVELS);
    ws.level_alpha = c.take<double>((size_t)B * YHB_MAX_LEVELS);
    ws.dcV = c.take<double>((size_t)B * d.N);
    ws.dcinfo = c.take<int>(2 * (size_t)B);
    const int M = d.N + pl->num_inductors;
    ws.dc_stride = dc_scratch_doubles(M, d.nd, d.nb, d.nc);
    ws.dcscratch = c.take<double>((size_t)B * ws.dc_stride);
    ws.key = c.take<double>(B);
    ws.order = c.take<int>(B);
    ws.dc_ready = c.take<int>((size_t)B + 8);  // + the counters of an overlapped cold solve
    ws.dc_hash = c.take<unsigned long long>(B);
    ws.dc_leader = c.take<int>(B);
    ws.use_order = 0;
    ws.dc_wait = 0;
    L.dd = c.take<DiodeDerived>((size_t)B * std::max(d.nd, 1));
    L.bd = c.take<BjtDerived>((size_t)B * std::max(d.nb, 1));
    L.bytes = align_up(c.off);
    return L;
}

int check_params(const yhb_plan *pl, const yhb_params *p) {
    if (!pl || !p) { set_error("null argument"); return -1; }
    if (p->batch < 1) { set_error("batch must be >= 1"); return -1; }
    const PlanDev &d = pl->d;
    if (!p->tones || (d.nlin && !p->lin_val) || (d.nsrc && !p->src_val) || (d.nd && !p->diode_par) ||
        (d.nb && !p->bjt_par) || (d.nc && !p->cubic_par)) {
        set_error("missing parameter array");
        return -1;
    }
    return 0;
}

int launch_prepare(const yhb_plan *pl, const yhb_params *p, const Layout &L, cudaStream_t st) {
    PrepArgs a;
    a.tones = p->tones;
    a.lin_val = p->lin_val;
    a.src_val = p->src_val;
    a.diode_par = p->diode_par;
    a.bjt_par = p->bjt_par;
    a.omega = L.ws.omega;
    a.ylin = L.ws.ylin;
    a.Is = L.ws.Is;
    a.negbin = L.ws.negbin;
    a.dd = L.dd;
    a.bd = L.bd;
    a.dylin = L.ws.dylin;
    a.key = L.ws.use_order ? L.ws.key : nullptr;
    ProfScope ps(PK_PREPARE, st);
    k_prepare<<<p->batch, 128, 0, st>>>(pl->d, a, p->batch);
    if (L.ws.use_order)
        k_order<<<(int)(((size_t)p->batch * kOrderLanes + 255) / 256), 256, 0, st>>>(L.ws.key, L.ws.order, p->batch);
    YHB_CUDA(cudaGetLastError());
    return 0;
}

EvalArgs eval_args(const yhb_plan *pl, const yhb_params *p, const yhb_options *opt, const Layout &L) {
    EvalArgs e;
    memset(&e, 0, sizeof(e));
    e.cubic_par = p->cubic_par;
    e.dd = L.dd;
    e.bd = L.bd;
    e.reltol = opt ? opt->reltol : 1e-3;
    e.abstol = opt ? opt->abstol : 1e-6;
    e.maxiter = opt ? opt->maxiter : 100;
    e.use_smem = L.eval_smem;
    return e;
}

int eval_smem_bytes(const yhb_plan *pl, const Layout &L) {
    return L.eval_smem ? (int)(eval_scratch_doubles(pl->d) * sizeof(double)) : 0;
}

// the attribute is per device: set it on every call (cheap), like the fused launcher does
int ensure_eval_attr(int bytes) {
    if (bytes > 48 * 1024)
        YHB_CUDA(cudaFuncSetAttribute(k_eval<kEvalThreads>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxEvalSmem));
    return 0;
}

int launch_finish(const yhb_plan *pl, Layout &L, const yhb_result *res, cudaStream_t st) {
    FinishArgs fa;
    fa.Vf = res->Vf;
    fa.converged = res->converged;
    fa.newton_iters = res->newton_iters;
    fa.lu_count = res->lu_count;
    fa.level_iters = res->level_iters;
    fa.level_alpha = res->level_alpha;
    fa.err = res->err;
    {
        ProfScope ps(PK_FINISH, st);
        k_finish<<<L.ws.batch, 128, 0, st>>>(pl->d, L.ws, fa);
    }
    YHB_CUDA(cudaGetLastError());
    return 0;
}

// Panel Gauss-Jordan of nsys systems (yhb_dense.cuh): per panel of 16 pivot columns one factorisation launch and
// one DMMA trailing-update launch over the remaining columns.
int gjp_run(const GjpBatch &g, int nsys, cudaStream_t st) {
    const int npanel = (g.n + kGjpNB - 1) / kGjpNB;
    const int rpt = g.n <= 576 ? 1 : (g.n <= 1152 ? 2 : 0);  // rows per thread of the register panel kernel
    const size_t smem = (size_t)kGjpNB * (g.n | 1) * sizeof(double);
    if (rpt == 0) {
        if (smem > 227 * 1024 - 4096) { set_error("panel Gauss-Jordan: more than 14 000 pivots per system"); return -4; }
        YHB_CUDA(cudaFuncSetAttribute(k_gjp_panel_smem, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    const int nrow_chunks = (g.n + kGjpUpdRows - 1) / kGjpUpdRows;
    for (int j = 0; j < npanel; ++j) {
        if (rpt == 1) k_gjp_panel<1><<<nsys, (g.n + 31) & ~31, 0, st>>>(g, j);
        else if (rpt == 2) k_gjp_panel<2><<<nsys, ((g.n + 1) / 2 + 31) & ~31, 0, st>>>(g, j);
        else k_gjp_panel_smem<<<nsys, 1024, smem, st>>>(g, j);
        const int rem = g.ncols - (j + 1) * kGjpNB;
        if (rem <= 0) continue;
        k_gjp_update<<<dim3((rem + kGjpUpdCols - 1) / kGjpUpdCols, nrow_chunks, nsys), 256, 0, st>>>(g, j);
    }
    YHB_CUDA(cudaGetLastError());
    return 0;
}

GjpBatch gjp_batch(const Workspace &ws, const int *list) {
    GjpBatch g;
    memset(&g, 0, sizeof(g));
    g.M = ws.J;
    g.m_stride = ws.j_stride;
    g.L = ws.gj_L;
    g.U = ws.gj_U;
    g.T = ws.gj_T;
    g.piv = ws.gj_piv;
    g.rowof = ws.gj_rowof;
    g.used = ws.gj_used;
    g.pp = ws.gj_pp;
    g.fail = ws.gj_fail;
    g.list = list;
    g.count = nullptr;  // the host knows the number of queued circuits: grids are sized exactly
    g.n = ws.gj_n;
    g.lpad = ws.gj_lpad;
    return g;
}

// Block-Thomas elimination of the queued circuits' block-tridiagonal cores (blocks as k_assemble_btd laid them out;
// right-hand sides in the block rows, solution into ws.rhs), all circuits concurrently and every chain from BOTH ends
// towards the middle block row mid (two block rows per circuit in every launch: half the serial depth):
//   elimination step j:    upper chain r = j, lower chain r = Nc - 1 - j (inner neighbour q = r + 1 / r - 1):
//                          [D_r | rhs_r] -= side(r) [X'_prev | b'_prev]                  (k_btd_schur: DMMA GEMM + GEMV)
//                          Gauss-Jordan of [D_r | carried(r) | rhs_r], pivoting inside D_r (gjp_run)
//                          X'_r = D_r^-1 carried(r) -> side storage,  b'_r = D_r^-1 rhs_r -> ws.rhs   (k_gjp_finish)
//   middle row:            one update from each side, then x_mid = D_mid^-1 rhs_mid
//   back substitution:     x_r = b'_r - X'_r x_q outwards from the middle, both chains per launch     (k_btd_back)
int block_thomas_solve(const PlanDev &d, const Workspace &ws, const int *list, int nsolve, cudaStream_t st) {
    const int S = d.S, Nc = d.Nc, mid = btd_mid(Nc);
    const int nup = mid, nlo = Nc - 1 - mid;  // rows of the upper / lower chain
    const size_t rowd = btd_row_doubles(S);
    GjpBatch g = gjp_batch(ws, list);
    g.ld = btd_ldm(S);
    g.ncols = gjp_npc(S) + S + 1;
    g.npos = nsolve;
    g.vs_stride = ws.batch;
    SchurArgs sa;
    memset(&sa, 0, sizeof(sa));
    sa.J = ws.J;
    sa.j_stride = ws.j_stride;
    sa.rhs = ws.rhs;
    sa.rhs_stride = (size_t)d.Wc + 1;
    sa.list = list;
    sa.count = nullptr;
    sa.S = S;
    sa.npos = nsolve;
    const int ntx = (S + kSchurBN - 1) / kSchurBN, nty = (S + kSchurBM - 1) / kSchurBM;
    const int nfin = std::min((S + 7) / 8, 148);
    // one Gauss-Jordan pass over block rows r0 (and r1 if >= 0); next: the carried block becomes X' in the side storage
    auto eliminate = [&](int r0, int r1, bool want_x) -> int {
        g.nchain = r1 >= 0 ? 2 : 1;
        g.m_off = (size_t)r0 * rowd;
        g.m_off2 = r1 >= 0 ? (size_t)r1 * rowd : 0;
        if (int rc = gjp_run(g, g.nchain * nsolve, st)) return rc;
        GjpFinish f;
        memset(&f, 0, sizeof(f));
        if (want_x) {
            f.out0 = ws.J;
            f.out0_stride = ws.j_stride;
            f.out0_off = g.m_off + (size_t)S * btd_ldm(S);
            f.out0_off2 = g.m_off2 + (size_t)S * btd_ldm(S);
            f.nout0 = S;
            f.ld0 = btd_lds(S);
        }
        f.out1 = ws.rhs;
        f.out1_stride = (size_t)d.Wc + 1;
        f.out1_off = (size_t)r0 * S;
        f.out1_off2 = r1 >= 0 ? (size_t)r1 * S : 0;
        f.col1 = S;
        k_gjp_finish<<<dim3(nfin, g.nchain * nsolve), 256, 0, st>>>(g, f);
        return 0;
    };
    const int steps = std::max(nup, nlo);
    for (int j = 0; j < steps; ++j) {
        const int ru = j < nup ? j : -1, rl = j < nlo ? Nc - 1 - j : -1;
        const int r0 = ru >= 0 ? ru : rl, r1 = (ru >= 0 && rl >= 0) ? rl : -1;
        if (j > 0) {
            sa.nchain = r1 >= 0 ? 2 : 1;
            sa.r[0] = r0;
            sa.rprev[0] = r0 == ru ? r0 - 1 : r0 + 1;
            sa.a_in_m[0] = 0;
            sa.r[1] = r1;
            sa.rprev[1] = r1 + 1;
            sa.a_in_m[1] = 0;
            k_btd_schur<<<dim3(ntx + 1, nty, sa.nchain * nsolve), 256, 0, st>>>(sa);
        }
        if (int rc = eliminate(r0, r1, true)) return rc;
    }
    // the middle row: the update from above, then the one from below (U_mid is the carried block of its Mr), the solve
    sa.nchain = 1;
    sa.r[0] = mid;
    if (nup > 0) {
        sa.rprev[0] = mid - 1;
        sa.a_in_m[0] = 0;
        k_btd_schur<<<dim3(ntx + 1, nty, nsolve), 256, 0, st>>>(sa);
    }
    if (nlo > 0) {
        sa.rprev[0] = mid + 1;
        sa.a_in_m[0] = 1;
        k_btd_schur<<<dim3(ntx + 1, nty, nsolve), 256, 0, st>>>(sa);
    }
    if (int rc = eliminate(mid, -1, false)) return rc;
    for (int dist = 1; dist <= steps; ++dist) {  // outwards: row mid - dist from mid - dist + 1, row mid + dist from mid + dist - 1
        const int ru = mid - dist >= 0 ? mid - dist : -1, rl = mid + dist < Nc ? mid + dist : -1;
        const int r0 = ru >= 0 ? ru : rl, r1 = (ru >= 0 && rl >= 0) ? rl : -1;
        sa.nchain = r1 >= 0 ? 2 : 1;
        sa.r[0] = r0;
        sa.rprev[0] = r0 == ru ? r0 + 1 : r0 - 1;
        sa.r[1] = r1;
        sa.rprev[1] = r1 - 1;
        k_btd_back<<<dim3(nfin, sa.nchain * nsolve), 256, 0, st>>>(sa);
    }
    YHB_CUDA(cudaGetLastError());
    return 0;
}

// calc_V0 (:697-704) = DC.run (DC.py:54-123) for the batch.  order / ready: see DcArgs (overlap with the HB kernel).
size_t dc_warp_smem(const yhb_plan *pl) {
    const PlanDev &d = pl->d;
    const int M = d.N + pl->num_inductors, LD = (M + 1) | 1, nnl = d.nd + d.nb + d.nc;
    return ((size_t)M * M + M + (size_t)kDcOut * std::max(nnl, 1) + (size_t)M * LD) * sizeof(double) +
           (size_t)pl->dc_ent_ints * sizeof(int);
}

bool dc_one_warp(const yhb_plan *pl) {
    const char *dc_env = getenv("YHB_DC_ENGINE");  // read per call: tests flip it within one process
    return pl->dc_warp && !(dc_env && !strcmp(dc_env, "block"));
}

// DcArgs with the fields both the grouping kernels and the DC kernels read
DcArgs dc_inputs(const yhb_params *p) {
    DcArgs a;
    memset(&a, 0, sizeof(a));
    a.lin_val = p->lin_val;
    a.src_val = p->src_val;
    a.diode_par = p->diode_par;
    a.bjt_par = p->bjt_par;
    a.cubic_par = p->cubic_par;
    return a;
}

// Groups of circuits with bit-identical DC inputs -> ws.dc_leader (see k_dc_group).  force: also for a single circuit
// (the HB kernel of an overlapped cold solve reads the leaders).  Returns the leader array, or nullptr when grouping is off.
const int *dc_groups(const yhb_plan *pl, const yhb_params *p, const Layout &L, cudaStream_t st, bool force) {
    static const bool no_groups = getenv("YHB_NO_DC_GROUPS") != nullptr;  // A/B runs
    if (!force && (no_groups || p->batch < 2)) return nullptr;
    const DcArgs a = dc_inputs(p);
    ProfScope ps(PK_DC, st);
    k_dc_hash<<<(p->batch + 3) / 4, 128, 0, st>>>(pl->d, a, p->batch, L.ws.dc_hash);
    k_dc_group<<<(int)(((size_t)p->batch * kDcGroupLanes + 255) / 256), 256, 0, st>>>(pl->d, a, p->batch, L.ws.dc_hash, L.ws.dc_leader,
                                                       force ? L.ws.dc_ready + p->batch + 3 : nullptr);
    return L.ws.dc_leader;
}

// leader: groups from dc_groups() (only the leaders are solved, the others get copies) or nullptr
int dc_launch(const yhb_plan *pl, const yhb_params *p, const Layout &L, double *Vdc, int32_t *dc_info, const int *order,
              int *ready, cudaStream_t st, const int *leader) {
    const PlanDev &d = pl->d;
    DcArgs a = dc_inputs(p);
    a.lin_nodes = pl->lin_nodes_dev;
    a.nl_kind = pl->nl_kind_dev;
    a.nl_index = pl->nl_index_dev;
    a.nnl = d.nd + d.nb + d.nc;
    a.M = d.N + pl->num_inductors;
    a.scratch = L.ws.dcscratch;
    a.stride = L.ws.dc_stride;
    a.Vdc = Vdc ? Vdc : L.ws.dcV;
    a.info = dc_info ? dc_info : L.ws.dcinfo;
    a.ent_tab = pl->dc_ent_dev;
    a.ent_ints = pl->dc_ent_ints;
    a.lane_tab = pl->dc_lane_dev;
    a.order = order;
    a.ready = ready;
    a.nready = ready ? L.ws.dc_ready + p->batch + 4 : nullptr;
    a.nstarted = ready ? L.ws.dc_ready + p->batch + 5 : nullptr;
    a.leader = leader;
    a.use_smem = 0;
    const bool groups = leader != nullptr;
    // YHB_DC_ENGINE=block forces the general engine (A/B runs, parity tests of the one-warp engine against it)
    if (dc_one_warp(pl)) {
        const size_t smw = dc_warp_smem(pl);
        ProfScope ps(PK_DC, st);
        if (a.M <= 12) k_dc_warp<12><<<p->batch, 32, smw, st>>>(d, a, p->batch);
        else k_dc_warp<24><<<p->batch, 32, smw, st>>>(d, a, p->batch);
        if (groups) k_dc_scatter<<<(p->batch + 255) / 256, 256, 0, st>>>(d.N, p->batch, a.leader, a.Vdc, a.info);
        YHB_CUDA(cudaGetLastError());
        return 0;
    }
    size_t dsm = (a.M + 2) * sizeof(double);
    a.use_smem = (dsm + a.stride * sizeof(double)) <= 40 * 1024;  // small circuits: the whole DC analysis on chip
    if (a.use_smem) dsm += a.stride * sizeof(double);
    {
        ProfScope ps(PK_DC, st);
        k_dc<<<p->batch, kDcThreads, dsm, st>>>(d, a, p->batch);
        if (groups) k_dc_scatter<<<(p->batch + 255) / 256, 256, 0, st>>>(d.N, p->batch, a.leader, a.Vdc, a.info);
    }
    YHB_CUDA(cudaGetLastError());
    return 0;
}

// occupancy / attributes / launch of the fused kernel variant NT
#define YHB_FUSED_CASES(X) X(0) X(1) YHB_FOR_NT(X)
int fused_query(int NT, int fsmem, int *per_sm, cudaFuncAttributes *fa) {
#define YHB_Q(t)                                                                                                  \
    case t:                                                                                                       \
        YHB_CUDA(cudaFuncSetAttribute(k_fused<t>, cudaFuncAttributeMaxDynamicSharedMemorySize, fsmem));            \
        YHB_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, k_fused<t>, fused_threads(t), fsmem));      \
        YHB_CUDA(cudaFuncGetAttributes(fa, k_fused<t>));                                                           \
        return 0;
    switch (NT) {
        YHB_FUSED_CASES(YHB_Q)
        default: set_error("no fused kernel for this core size"); return -4;
    }
#undef YHB_Q
}
int fused_launch(int NT, int grid, int fsmem, cudaStream_t st, const PlanDev &d, const Workspace &ws, const EvalArgs &ea,
                 int hot) {
#define YHB_L(t) \
    case t: k_fused<t><<<grid, fused_threads(t), fsmem, st>>>(d, ws, ea, hot); break;
    switch (NT) {
        YHB_FUSED_CASES(YHB_L)
        default: set_error("no fused kernel for this core size"); return -4;
    }
#undef YHB_L
    YHB_CUDA(cudaGetLastError());
    return 0;
}

// DC analysis inside a cold solve (yhb_solve_cold): outputs of the DC kernel
struct DcOverlap {
    double *Vdc;
    int32_t *info;
};

// the Newton / continuation launch loop shared by yhb_solve and yhb_newton
// dco (fused engine, cold start only): the DC analysis runs on the plan's high-priority side stream BESIDE the HB
// kernel, which takes the circuits in the DC kernel's order and waits for each operating point (ws.dc_wait)
int run_loop(const yhb_plan *cpl, const yhb_params *p, const yhb_options *opt, Layout &L, int have_v0,
             const double *alpha_fixed, const yhb_result *res, cudaStream_t st, const DcOverlap *dco = nullptr) {
    yhb_plan *pl = const_cast<yhb_plan *>(cpl);  // side stream and events are created on first use
    const PlanDev &d = pl->d;
    const int B = p->batch;
    Workspace &ws = L.ws;
    // overlap: the DC kernel runs BESIDE the HB kernel
    int fused_per_sm = 0, sms = 0;
    bool overlap = false;
    if (L.fused) {
        int dev = 0;
        YHB_CUDA(cudaGetDevice(&dev));
        YHB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        cudaFuncAttributes fa;
        if (int rc = fused_query(L.fc.NT, (int)L.fc.smem, &fused_per_sm, &fa)) return rc;
        fused_per_sm = std::max(1, fused_per_sm);
        // Cold solves: circuits with identical DC inputs share one DC analysis (an amplitude sweep repeats every bias point),
        // and the DC kernel -- as long as its slowest circuit, one warp per distinct problem -- runs BESIDE the HB kernel:
        // side stream, a ready flag per distinct problem, and a gate kernel in front of the HB kernel that opens when every
        // DC CTA has begun (resident or done), so that the waiting HB CTAs can never keep DC work off the device.
        // YHB_NO_DC_OVERLAP=1: DC first, on the same stream (A/B runs).
        static const bool no_overlap = getenv("YHB_NO_DC_OVERLAP") != nullptr;
        overlap = dco && !no_overlap && dc_one_warp(pl);
    }
    const int *leader = nullptr;
    if (dco) {
        if (!L.fused || have_v0) { set_error("internal: the fused cold solve needs the fused engine and a cold start"); return -4; }
        ws.use_order = 1;
        ws.dc_valid = 1;
        if (dco->Vdc) ws.dcV = dco->Vdc;
        if (overlap) {
            if (!pl->dc_stream) {
                int lo = 0, hi = 0;
                YHB_CUDA(cudaDeviceGetStreamPriorityRange(&lo, &hi));
                YHB_CUDA(cudaStreamCreateWithPriority(&pl->dc_stream, cudaStreamNonBlocking, hi));
                YHB_CUDA(cudaEventCreateWithFlags(&pl->ev_fork, cudaEventDisableTiming));
                YHB_CUDA(cudaEventCreateWithFlags(&pl->ev_join, cudaEventDisableTiming));
            }
            ws.dc_wait = 1;
            YHB_CUDA(cudaMemsetAsync(ws.dc_ready, 0, sizeof(int) * ((size_t)B + 8), st));
            YHB_CUDA(cudaMemsetAsync(ws.list[0], 0, sizeof(int) * (size_t)B, st));  // the HB kernel's deferred list
        }
        leader = dc_groups(pl, p, L, st, overlap);  // before the fork: the HB kernel reads the leaders too
    }
    if (int rc = launch_prepare(pl, p, L, st)) return rc;
    if (dco && !overlap)  // DC first, on the same stream
        if (int rc = dc_launch(pl, p, L, ws.dcV, dco->info, nullptr, nullptr, st, leader)) return rc;
    {
        ProfScope ps(PK_INIT, st);
        k_init<<<B, 128, 0, st>>>(d, ws, have_v0, alpha_fixed);
    }
    YHB_CUDA(cudaGetLastError());
    if (overlap) {  // fork: the DC kernel on the high-priority stream, in the HB kernel's work order; then the gate
        YHB_CUDA(cudaEventRecord(pl->ev_fork, st));
        YHB_CUDA(cudaStreamWaitEvent(pl->dc_stream, pl->ev_fork, 0));
        if (int rc = dc_launch(pl, p, L, ws.dcV, dco->info, ws.order, ws.dc_ready, pl->dc_stream, leader)) return rc;
        YHB_CUDA(cudaEventRecord(pl->ev_join, pl->dc_stream));
        k_dc_gate<<<1, 1, 0, st>>>(ws.dc_ready + B + 3, ws.dc_ready + B + 5);
        YHB_CUDA(cudaGetLastError());
    }
    EvalArgs ea = eval_args(pl, p, opt, L);
    if (L.fused) {
        // persistent CTAs pulling circuits from a queue: grid = resident CTAs per SM x SMs
        const int grid = std::min(B, fused_per_sm * sms);
        if (getenv("YHB_DEBUG"))
            fprintf(stderr, "yhb: k_fused<%d> threads %d smem %d B, %d CTAs/SM, grid %d, block-sparse %d, DC beside it %d\n", L.fc.NT,
                    fused_threads(L.fc.NT), (int)L.fc.smem, fused_per_sm, grid, d.bs_ok, (int)overlap);
        {
            ProfScope ps(PK_FUSED, st);
            if (int rc = fused_launch(L.fc.NT, grid, (int)L.fc.smem, st, d, ws, ea, L.fc.hot)) return rc;
        }
        if (overlap) YHB_CUDA(cudaStreamWaitEvent(st, pl->ev_join, 0));  // join: Vdc / dc_info are complete for the caller
        return launch_finish(pl, L, res, st);
    }
    const int smem = eval_smem_bytes(pl, L);
    if (int rc = ensure_eval_attr(smem)) return rc;
    int cur = 0, nactive = B;
    int *hc = pl->pinned_counts;
    long rounds = 0;
    while (nactive > 0) {
        YHB_CUDA(cudaMemsetAsync(ws.count + (cur ^ 1), 0, sizeof(int), st));
        ea.cur = cur;
        {
            ProfScope ps(PK_EVAL, st);
            // scratch in HBM = a large circuit: all the threads a CTA can have (scratch in shared memory: 256)
            if (L.eval_smem) k_eval<kEvalThreads><<<nactive, kEvalThreads, smem, st>>>(d, ws, ea);
            else k_eval<kEvalThreadsLarge><<<nactive, kEvalThreadsLarge, 0, st>>>(d, ws, ea);
        }
        YHB_CUDA(cudaGetLastError());
        YHB_CUDA(cudaMemcpyAsync(hc, ws.count, 4 * sizeof(int), cudaMemcpyDeviceToHost, st));
        YHB_CUDA(cudaStreamSynchronize(st));
        const int nsolve = hc[cur ^ 1];
        if (nsolve > 0 && ws.gj_n > 0) {
            // panel Gauss-Jordan (dense core) / block-Thomas engine: peeled rows + rhs, assembly, elimination of all
            // queued circuits concurrently, back-substitution of the peeled rows + update
            const size_t rowbuf = 4;
            {
                ProfScope ps(PK_LU, st);
                k_solve_core<<<nsolve, 1024, rowbuf * sizeof(double), st>>>(d, ws, ea, cur ^ 1, 1);
            }
            if (d.btd) {
                dim3 ga(nsolve, 3 * d.Nc, (d.S * d.S + kBtdChunk - 1) / kBtdChunk);
                ProfScope ps(PK_ASSEMBLE, st);
                k_assemble_btd<<<ga, 256, 0, st>>>(d, ws, ea, cur ^ 1);
            } else {
                dim3 ga(nsolve, (d.Wc + kAsmRows - 1) / kAsmRows);
                ProfScope ps(PK_ASSEMBLE, st);
                k_assemble_core<<<ga, 256, 0, st>>>(d, ws, ea, cur ^ 1, gjp_dense_ld(d.Wc));
            }
            YHB_CUDA(cudaGetLastError());
            {
                ProfScope ps(PK_LU, st);
                if (d.btd) {
                    if (int rc = block_thomas_solve(d, ws, ws.list[cur ^ 1], nsolve, st)) return rc;
                } else {
                    GjpBatch g = gjp_batch(ws, ws.list[cur ^ 1]);
                    g.ld = gjp_dense_ld(d.Wc);
                    g.ncols = gjp_npc(d.Wc) + 1;
                    if (int rc = gjp_run(g, nsolve, st)) return rc;
                    GjpFinish f;
                    memset(&f, 0, sizeof(f));
                    f.out1 = ws.rhs;
                    f.out1_stride = (size_t)d.Wc + 1;
                    f.col1 = 0;
                    k_gjp_finish<<<dim3(std::min((d.Wc + 7) / 8, 148), nsolve), 256, 0, st>>>(g, f);
                }
                k_solve_core<<<nsolve, 1024, rowbuf * sizeof(double), st>>>(d, ws, ea, cur ^ 1, 2);
            }
            YHB_CUDA(cudaGetLastError());
        } else if (nsolve > 0) {
            if (d.Wc > 0) {
                dim3 ga(nsolve, (d.Wc + kAsmRows - 1) / kAsmRows);
                ProfScope ps(PK_ASSEMBLE, st);
                k_assemble_core<<<ga, 256, 0, st>>>(d, ws, ea, cur ^ 1, d.Wc);
            }
            {
                ProfScope ps(PK_LU, st);
                // large cores: the HBM-resident elimination is latency bound, more warps per circuit hide it
                const int lu_threads = d.Wc >= 160 ? 1024 : kLuThreads;
                const size_t rowbuf = d.band > 0 ? std::min<size_t>(d.Wc, 2 * (size_t)d.band + 1) + 3 : (size_t)d.Wc + 2;
                k_solve_core<<<nsolve, lu_threads, rowbuf * sizeof(double), st>>>(d, ws, ea, cur ^ 1, 0);
            }
            YHB_CUDA(cudaGetLastError());
        }
        nactive = nsolve;
        cur ^= 1;
        if (++rounds > 100000) { set_error("Newton launch loop did not terminate"); return -4; }
    }
    return launch_finish(pl, L, res, st);
}

}  // namespace

extern "C" size_t yhb_workspace_bytes(const yhb_plan *pl, int32_t batch) {
    if (!pl || batch < 1) return 0;
    return carve(pl, batch, nullptr).bytes;
}

extern "C" int yhb_dc_solve(const yhb_plan *pl, const yhb_params *p, void *wsbuf, size_t ws_bytes, double *Vdc,
                            int32_t *dc_info, void *stream) {
    if (int rc = check_params(pl, p)) return rc;
    Layout L = carve(pl, p->batch, wsbuf);
    if (!wsbuf || ws_bytes < L.bytes) { set_error("workspace too small"); return -1; }
    return dc_launch(pl, p, L, Vdc, dc_info, nullptr, nullptr, (cudaStream_t)stream,
                     dc_groups(pl, p, L, (cudaStream_t)stream, false));
}

// lock = false: the caller (yhb_run_host) already holds pl->mu
static int solve_common(const yhb_plan *pl, const yhb_params *p, const yhb_options *opt, void *wsbuf,
                        size_t ws_bytes, int have_v0, const double *Vdc, const double *alpha,
                        const yhb_result *res_in, void *stream, bool lock, const DcOverlap *dco = nullptr) {
    if (int rc = check_params(pl, p)) return rc;
    yhb_result R;
    if (int rc = load_result(res_in, R)) return rc;
    Layout L = carve(pl, p->batch, wsbuf);
    if (!wsbuf || ws_bytes < L.bytes) { set_error("workspace too small"); return -1; }
    if (have_v0 && !R.V) { set_error("have_v0 set but res->V is null"); return -1; }
    if (!have_v0 && !Vdc && !dco) { set_error("cold start needs Vdc"); return -1; }
    if (R.V) L.ws.V = R.V;
    L.ws.Inl = R.Inl;
    L.ws.Ic = R.Ic;
    if (Vdc) L.ws.dcV = const_cast<double *>(Vdc);
    L.ws.dc_valid = Vdc != nullptr;
    std::unique_lock<std::mutex> lk(const_cast<yhb_plan *>(pl)->mu, std::defer_lock);  // pinned counters are per plan
    if (lock) lk.lock();
    return run_loop(pl, p, opt, L, have_v0, alpha, &R, (cudaStream_t)stream, dco);
}

// calc_V0 + run() for cold starts (:321-324, :338-384) as ONE call.  Fused engine: the DC kernel runs beside the HB
// kernel (side stream, per-circuit ready flags) instead of in front of it; staged engines: DC, then the solve.
static int solve_cold(const yhb_plan *pl, const yhb_params *p, const yhb_options *opt, void *wsbuf, size_t ws_bytes,
                      double *Vdc, int32_t *dc_info, const yhb_result *res, void *stream, bool lock) {
    if (int rc = check_params(pl, p)) return rc;
    Layout L = carve(pl, p->batch, wsbuf);
    if (!wsbuf || ws_bytes < L.bytes) { set_error("workspace too small"); return -1; }
    if (L.fused) {
        DcOverlap dco{Vdc ? Vdc : L.ws.dcV, dc_info ? dc_info : L.ws.dcinfo};
        return solve_common(pl, p, opt, wsbuf, ws_bytes, 0, nullptr, nullptr, res, stream, lock, &dco);
    }
    if (int rc = dc_launch(pl, p, L, Vdc, dc_info, nullptr, nullptr, (cudaStream_t)stream,
                           dc_groups(pl, p, L, (cudaStream_t)stream, false)))
        return rc;
    return solve_common(pl, p, opt, wsbuf, ws_bytes, 0, Vdc ? Vdc : L.ws.dcV, nullptr, res, stream, lock);
}

extern "C" int yhb_solve_cold(const yhb_plan *pl, const yhb_params *p, const yhb_options *opt, void *wsbuf,
                              size_t ws_bytes, double *Vdc, int32_t *dc_info, const yhb_result *res, void *stream) {
    return solve_cold(pl, p, opt, wsbuf, ws_bytes, Vdc, dc_info, res, stream, true);
}

extern "C" int yhb_solve(const yhb_plan *pl, const yhb_params *p, const yhb_options *opt, void *wsbuf,
                         size_t ws_bytes, int32_t have_v0, const double *Vdc, const yhb_result *res, void *stream) {
    return solve_common(pl, p, opt, wsbuf, ws_bytes, have_v0, Vdc, nullptr, res, stream, true);
}

extern "C" int yhb_newton(const yhb_plan *pl, const yhb_params *p, const yhb_options *opt, void *wsbuf,
                          size_t ws_bytes, const double *alpha, const yhb_result *res, void *stream) {
    if (!alpha) { set_error("null alpha"); return -1; }
    return solve_common(pl, p, opt, wsbuf, ws_bytes, 1, nullptr, alpha, res, stream, true);
}

// ---- end-to-end entry with host buffers -----------------------------------------------------------
// Device staging of a *_host call: parameters + results + workspace in ONE allocation cached on the plan (grown on
// demand).  The caller holds pl->mu for as long as it uses the buffer and the plan's stream.
namespace {
struct HostStage {
    yhb_params dp;
    yhb_result dr;
    void *wsb;
    size_t wsbytes;
};

int stage_host(yhb_plan *pl, size_t B, const yhb_result &want, HostStage &hs) {
    const PlanDev &d = pl->d;
    const size_t W = d.W, K1 = d.K + 1;
    if (!pl->own_stream) YHB_CUDA(cudaStreamCreateWithFlags(&pl->own_stream, cudaStreamNonBlocking));
    hs.wsbytes = yhb_workspace_bytes(pl, (int)B);
    auto plan_bytes = [&](Carver &cv) {
        yhb_params &dp = hs.dp;
        yhb_result &dr = hs.dr;
        memset(&dr, 0, sizeof(dr));
        dr.struct_size = sizeof(dr);
        dp.batch = (int)B;
        dp.tones = cv.take<double>(B * 2);
        dp.lin_val = cv.take<double>(B * std::max(d.nlin, 1) * 2);
        dp.src_val = cv.take<double>(B * std::max(d.nsrc, 1) * 3);
        dp.diode_par = cv.take<double>(B * std::max(d.nd, 1) * YHB_DIODE_NPAR);
        dp.bjt_par = cv.take<double>(B * std::max(d.nb, 1) * YHB_BJT_NPAR);
        dp.cubic_par = cv.take<double>(B * std::max(d.nc, 1));
        dr.V = cv.take<double>(B * W);
        dr.Vf = cv.take<double>(B * d.N * K1 * 2);
        dr.converged = cv.take<int32_t>(B);
        dr.newton_iters = cv.take<int32_t>(B);
        dr.lu_count = cv.take<int32_t>(B);
        // the per-level reports are 768 B per circuit: staged (and copied back) only when the caller asks for them
        dr.level_iters = want.level_iters ? cv.take<int32_t>(B * YHB_MAX_LEVELS) : nullptr;
        dr.level_alpha = want.level_alpha ? cv.take<double>(B * YHB_MAX_LEVELS) : nullptr;
        dr.err = cv.take<double>(B);
        dr.Inl = want.Inl ? cv.take<double>(B * W) : nullptr;
        dr.Ic = want.Ic ? cv.take<double>(B * std::max(d.nb, 1) * d.S) : nullptr;
        hs.wsb = cv.take<char>(hs.wsbytes);
    };
    Carver c{nullptr, 0};
    plan_bytes(c);
    const size_t need = align_up(c.off);
    if (need > pl->hbuf_bytes) {
        if (pl->hbuf) cudaFree(pl->hbuf);
        pl->hbuf = nullptr;
        pl->hbuf_bytes = 0;
        YHB_CUDA(cudaMalloc(&pl->hbuf, need));
        pl->hbuf_bytes = need;
    }
    Carver c2{reinterpret_cast<char *>(pl->hbuf), 0};
    plan_bytes(c2);
    return 0;
}

int h2d_params(const yhb_plan *pl, const yhb_params *hp, const yhb_params &dp, cudaStream_t st) {
    const PlanDev &d = pl->d;
    const size_t B = hp->batch;
#define H2D(dst, src, cnt)                                                                                   \
    do {                                                                                                     \
        if ((cnt) > 0 && (src))                                                                              \
            YHB_CUDA(cudaMemcpyAsync((void *)(dst), (src), (cnt) * sizeof(*(src)), cudaMemcpyHostToDevice, st)); \
    } while (0)
    H2D(dp.tones, hp->tones, B * 2);
    H2D(dp.lin_val, hp->lin_val, B * d.nlin * 2);
    H2D(dp.src_val, hp->src_val, B * d.nsrc * 3);
    H2D(dp.diode_par, hp->diode_par, B * d.nd * YHB_DIODE_NPAR);
    H2D(dp.bjt_par, hp->bjt_par, B * d.nb * YHB_BJT_NPAR);
    H2D(dp.cubic_par, hp->cubic_par, B * d.nc);
#undef H2D
    return 0;
}

int d2h_result(const yhb_plan *pl, size_t B, const yhb_result &hr, const yhb_result &dr, cudaStream_t st) {
    const PlanDev &d = pl->d;
    const size_t W = d.W, K1 = d.K + 1;
#define D2H(dst, src, cnt)                                                                              \
    do {                                                                                                \
        if ((dst) && (src)) YHB_CUDA(cudaMemcpyAsync((dst), (src), (cnt) * sizeof(*(dst)), cudaMemcpyDeviceToHost, st)); \
    } while (0)
    D2H(hr.V, dr.V, B * W);
    D2H(hr.Vf, dr.Vf, B * d.N * K1 * 2);
    D2H(hr.converged, dr.converged, B);
    D2H(hr.newton_iters, dr.newton_iters, B);
    D2H(hr.lu_count, dr.lu_count, B);
    D2H(hr.level_iters, dr.level_iters, B * YHB_MAX_LEVELS);
    D2H(hr.level_alpha, dr.level_alpha, B * YHB_MAX_LEVELS);
    D2H(hr.err, dr.err, B);
    D2H(hr.Inl, dr.Inl, B * W);
    D2H(hr.Ic, dr.Ic, B * d.nb * d.S);
#undef D2H
    return 0;
}
}  // namespace

extern "C" int yhb_run_host(const yhb_plan *cpl, const yhb_params *hp, const yhb_options *opt, const double *V0,
                            const yhb_result *hr_in) {
    yhb_plan *pl = const_cast<yhb_plan *>(cpl);
    if (int rc = check_params(pl, hp)) return rc;
    yhb_result hr;
    if (int rc = load_result(hr_in, hr)) return rc;
    // the staging buffer, the stream and the pinned counters belong to the plan: one host call at a time per plan
    std::lock_guard<std::mutex> lk(pl->mu);
    const PlanDev &d = pl->d;
    const size_t B = hp->batch, W = d.W;
    HostStage hs;
    if (int rc = stage_host(pl, B, hr, hs)) return rc;
    cudaStream_t st = pl->own_stream;
    yhb_params &dp = hs.dp;
    yhb_result &dr = hs.dr;
    void *wsb = hs.wsb;
    const size_t wsbytes = hs.wsbytes;
    if (int rc = h2d_params(pl, hp, dp, st)) return rc;
    if (V0) YHB_CUDA(cudaMemcpyAsync(dr.V, V0, B * W * sizeof(double), cudaMemcpyHostToDevice, st));
    Layout L = carve(pl, (int)B, wsb);
    if (!V0) {  // cold start: DC operating points (calc_V0, :321-322) beside / before the ladder
        if (int rc = solve_cold(pl, &dp, opt, wsb, wsbytes, nullptr, nullptr, &dr, st, false)) return rc;
    } else {
        // warm start: the DC analysis is only needed if a direct attempt fails (:333-335); such circuits come
        // back with converged = -1 and the batch is re-run from the same V0 with the DC points available
        if (int rc = solve_common(pl, &dp, opt, wsb, wsbytes, 1, nullptr, nullptr, &dr, st, false)) return rc;
        std::vector<int32_t> conv(B);
        YHB_CUDA(cudaMemcpyAsync(conv.data(), dr.converged, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
        YHB_CUDA(cudaStreamSynchronize(st));
        bool need_dc = false;
        for (size_t i = 0; i < B; ++i) need_dc = need_dc || conv[i] < 0;
        if (need_dc) {
            YHB_CUDA(cudaMemcpyAsync(dr.V, V0, B * W * sizeof(double), cudaMemcpyHostToDevice, st));
            if (int rc = yhb_dc_solve(pl, &dp, wsb, wsbytes, nullptr, nullptr, st)) return rc;
            if (int rc = solve_common(pl, &dp, opt, wsb, wsbytes, 1, L.ws.dcV, nullptr, &dr, st, false)) return rc;
        }
    }
    if (int rc = d2h_result(pl, B, hr, dr, st)) return rc;
    YHB_CUDA(cudaStreamSynchronize(st));
    return 0;
}

// ---- autonomous oscillators (run_oscillator, :137-232) ------------------------------------------------
namespace {
int load_osc(const yhb_osc_options *in, yhb_osc_options &o, const yhb_osc_result *rin, yhb_osc_result &r) {
    memset(&o, 0, sizeof(o));
    memset(&r, 0, sizeof(r));
    if (!in || !rin) { set_error("null oscillator options / result"); return -1; }
    if (in->struct_size < 4 * sizeof(int32_t) + sizeof(size_t) || in->struct_size > 4096 ||
        rin->struct_size < sizeof(size_t) + 2 * sizeof(void *) || rin->struct_size > 4096) {
        set_error("yhb_osc_options / yhb_osc_result: struct_size not set");
        return -1;
    }
    memcpy(&o, in, std::min((size_t)in->struct_size, sizeof(o)));
    memcpy(&r, rin, std::min((size_t)rin->struct_size, sizeof(r)));
    if (o.max_outer <= 0) o.max_outer = 30;
    if (!(o.tol > 0)) o.tol = 1e-12;          // the reference stops at |Yosc| < 1e-12 (:191)
    if (!(o.max_df_rel > 0)) o.max_df_rel = 0.02;
    if (!(o.max_dv_rel > 0)) o.max_dv_rel = 0.3;
    return 0;
}

// device part of the oscillator solve; the caller holds pl->mu.  dp's tones / lin_val / src_val are rewritten.
int osc_run(const yhb_plan *pl, const yhb_params *dp, const yhb_options *opt, const yhb_osc_options &oo, void *wsbuf,
            size_t ws_bytes, const yhb_result &R, const yhb_osc_result &dor, cudaStream_t st, int32_t *iterations) {
    const PlanDev &d = pl->d;
    const int B = dp->batch;
    if (d.ntones != 1) { set_error("oscillator solve: single-tone analyses only (as run_oscillator, :137)"); return -1; }
    if (oo.probe_src < 0 || oo.probe_src >= d.nsrc || oo.probe_filter < 0 || oo.probe_filter >= d.nlin ||
        oo.node < 1 || oo.node > d.N || pl->lin_kind[oo.probe_filter] != YHB_LIN_IHF ||
        pl->src_kind[oo.probe_src] != YHB_SRC_AC1) {
        set_error("oscillator solve: probe_src must be an AC source at the tone, probe_filter an IdealHarmonicFilter, "
                  "node a non-ground node");
        return -1;
    }
    const int fn1 = pl->lin_nodes[4 * oo.probe_filter], fn2 = pl->lin_nodes[4 * oo.probe_filter + 1];
    if (fn1 != oo.node && fn2 != oo.node) { set_error("oscillator solve: the probe filter is not connected to node"); return -1; }
    if (!dor.fosc || !dor.Vosc) { set_error("oscillator solve: fosc / Vosc (start point in, result out) are required"); return -1; }
    Layout L = carve(pl, B, wsbuf);
    if (!wsbuf || ws_bytes < L.bytes) { set_error("workspace too small"); return -1; }
    if (!L.fused) { set_error("oscillator solve: the circuit does not fit the on-chip engine"); return -4; }
    Workspace &ws = L.ws;
    if (R.V) ws.V = R.V;
    ws.Inl = ws.osc_Inl;
    ws.Ic = R.Ic;
    ws.sens = ws.osc_sens;
    ws.dylin = ws.osc_dylin;
    ws.skip = ws.osc_skip;
    ws.sens_src_n1 = pl->src_nodes[2 * oo.probe_src];
    ws.sens_src_n2 = pl->src_nodes[2 * oo.probe_src + 1];
    ws.dc_valid = 1;
    OscArgs a;
    a.st = reinterpret_cast<OscState *>(ws.osc_state);
    a.tones = const_cast<double *>(dp->tones);
    a.lin_val = const_cast<double *>(dp->lin_val);
    a.src_val = const_cast<double *>(dp->src_val);
    a.V = ws.V;
    a.Vacc = ws.osc_Vacc;
    a.Inl = ws.osc_Inl;
    a.sens = ws.osc_sens;
    a.ctl = ws.ctl;
    a.skip = ws.osc_skip;
    a.live = ws.osc_live;
    a.probe_src = oo.probe_src;
    a.probe_filter = oo.probe_filter;
    a.node = oo.node;
    a.nosc2 = fn1 == oo.node ? fn2 : fn1;
    a.tol = oo.tol;
    a.max_df_rel = oo.max_df_rel;
    a.max_da_rel = oo.max_dv_rel;
    a.B = B;
    k_osc_begin<<<(B + 127) / 128, 128, 0, st>>>(a, dor.fosc, dor.Vosc);
    YHB_CUDA(cudaGetLastError());
    // calc_V0: the DC point does not depend on the probe (AC source, filter: 1e-12 S at DC)
    if (int rc = yhb_dc_solve(pl, dp, wsbuf, ws_bytes, ws.dcV, ws.dcinfo, st)) return rc;
    int *hc = pl->pinned_counts;
    int it = 0;
    for (; it < oo.max_outer; ++it) {
        k_osc_set<<<B, 128, 0, st>>>(d, a);
        YHB_CUDA(cudaMemsetAsync(ws.osc_live, 0, sizeof(int), st));
        // first pass: cold start of every oscillator (run(netlist), :171); afterwards warm starts from the accepted
        // point (run(netlist, self.V), :166), the DC point at hand for a failed direct attempt (:333-335)
        if (int rc = run_loop(pl, dp, opt, L, it > 0, nullptr, &R, st)) return rc;
        k_osc_update<<<B, 128, 0, st>>>(d, a);
        YHB_CUDA(cudaGetLastError());
        YHB_CUDA(cudaMemcpyAsync(hc, ws.osc_live, sizeof(int), cudaMemcpyDeviceToHost, st));
        YHB_CUDA(cudaStreamSynchronize(st));
        if (hc[0] == 0) { ++it; break; }
    }
    if (iterations) *iterations = it;
    OscOut o;
    o.fosc = dor.fosc;
    o.Vosc = dor.Vosc;
    o.yosc = dor.yosc;
    o.converged = dor.converged;
    o.outer_iters = dor.outer_iters;
    o.hb_newton_iters = dor.hb_newton_iters;
    o.diag = dor.diag;
    k_osc_finish<<<B, 128, 0, st>>>(d, a, o);
    YHB_CUDA(cudaGetLastError());
    ws.skip = nullptr;  // phasors of the accepted points for every oscillator
    return launch_finish(pl, L, &R, st);
}
}  // namespace

extern "C" int yhb_solve_oscillator(const yhb_plan *pl, const yhb_params *p, const yhb_options *opt,
                                    const yhb_osc_options *osc, void *wsbuf, size_t ws_bytes, const yhb_result *res,
                                    const yhb_osc_result *ores, void *stream) {
    if (int rc = check_params(pl, p)) return rc;
    yhb_result R;
    if (int rc = load_result(res, R)) return rc;
    yhb_osc_options oo;
    yhb_osc_result dor;
    if (int rc = load_osc(osc, oo, ores, dor)) return rc;
    std::lock_guard<std::mutex> lk(const_cast<yhb_plan *>(pl)->mu);
    return osc_run(pl, p, opt, oo, wsbuf, ws_bytes, R, dor, (cudaStream_t)stream, nullptr);
}

extern "C" int yhb_oscillator_host(const yhb_plan *cpl, const yhb_params *hp, const yhb_options *opt,
                                   const yhb_osc_options *osc, const yhb_result *hr_in, const yhb_osc_result *hor_in) {
    yhb_plan *pl = const_cast<yhb_plan *>(cpl);
    if (int rc = check_params(pl, hp)) return rc;
    yhb_result hr;
    if (int rc = load_result(hr_in, hr)) return rc;
    yhb_osc_options oo;
    yhb_osc_result hor;
    if (int rc = load_osc(osc, oo, hor_in, hor)) return rc;
    if (!hor.fosc || !hor.Vosc) { set_error("oscillator solve: fosc / Vosc (start point in, result out) are required"); return -1; }
    std::lock_guard<std::mutex> lk(pl->mu);
    const size_t B = hp->batch;
    HostStage hs;
    if (int rc = stage_host(pl, B, hr, hs)) return rc;
    cudaStream_t st = pl->own_stream;
    if (int rc = h2d_params(pl, hp, hs.dp, st)) return rc;
    // oscillator results: one small device block [fosc | Vosc | yosc (2) | converged, outer, newton (ints)]
    double *dblk = nullptr;
    YHB_CUDA(cudaMalloc(&dblk, B * (20 * sizeof(double) + 4 * sizeof(int32_t))));
    yhb_osc_result dor;
    memset(&dor, 0, sizeof(dor));
    dor.struct_size = sizeof(dor);
    dor.fosc = dblk;
    dor.Vosc = dblk + B;
    dor.yosc = dblk + 2 * B;
    dor.diag = dblk + 4 * B;
    int32_t *iblk = reinterpret_cast<int32_t *>(dblk + 20 * B);
    dor.converged = iblk;
    dor.outer_iters = iblk + B;
    dor.hb_newton_iters = iblk + 2 * B;
    int rc = 0;
    auto cp = [&](void *dst, const void *src, size_t bytes, cudaMemcpyKind k) {
        if (rc == 0 && dst && src && cudaMemcpyAsync(dst, src, bytes, k, st) != cudaSuccess) {
            set_error("oscillator solve: copy failed");
            rc = -2;
        }
    };
    cp(dor.fosc, hor.fosc, B * sizeof(double), cudaMemcpyHostToDevice);
    cp(dor.Vosc, hor.Vosc, B * sizeof(double), cudaMemcpyHostToDevice);
    if (rc == 0) rc = osc_run(pl, &hs.dp, opt, oo, hs.wsb, hs.wsbytes, hs.dr, dor, st, nullptr);
    if (rc == 0) rc = d2h_result(pl, B, hr, hs.dr, st);
    cp(hor.fosc, dor.fosc, B * sizeof(double), cudaMemcpyDeviceToHost);
    cp(hor.Vosc, dor.Vosc, B * sizeof(double), cudaMemcpyDeviceToHost);
    cp(hor.yosc, dor.yosc, 2 * B * sizeof(double), cudaMemcpyDeviceToHost);
    cp(hor.diag, dor.diag, 16 * B * sizeof(double), cudaMemcpyDeviceToHost);
    cp(hor.converged, dor.converged, B * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cp(hor.outer_iters, dor.outer_iters, B * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cp(hor.hb_newton_iters, dor.hb_newton_iters, B * sizeof(int32_t), cudaMemcpyDeviceToHost);
    cudaStreamSynchronize(st);
    cudaFree(dblk);
    return rc;
}

// ---- stage entry points -----------------------------------------------------------------------------
namespace yhb {

__global__ void k_stage_ifft(PlanDev P, int batch, const double *V, double *vt) {
    const int b = blockIdx.x;
    if (b >= batch) return;
    for (int i = threadIdx.x; i < P.W; i += blockDim.x) {
        int n = i / P.S, s = i - n * P.S;
        const double *row = P.idft + (size_t)s * P.S;
        const double *x = V + (size_t)b * P.W + n * P.S;
        double acc = 0.;
        for (int c = 0; c < P.S; ++c) acc = fma(row[c], x[c], acc);
        vt[(size_t)b * P.W + i] = acc;
    }
}

// Stand-alone device sweep (:441-570): HBM in (vt, vtprev samples + raw model parameters), HBM out (model outputs
// per sample).  Persistent CTAs walk chunks of kDevChunk circuits: the hoisted model constants of the chunk are
// derived into shared memory by one thread per (circuit, device), then one thread per (circuit, device, sample)
// evaluates; loads and stores are contiguous along the sample index.
template <int kDevChunk, int MINB>
__global__ void __launch_bounds__(256, MINB) k_stage_device_eval(PlanDev P, int batch, const double *diode_par,
                                                              const double *bjt_par, const double *cubic_par,
                                                              const double *vt, const double *vtprev, double *diode_out,
                                                              double *bjt_out, double *cubic_out) {
    extern __shared__ double sm[];
    DiodeDerived *dd = reinterpret_cast<DiodeDerived *>(sm);
    BjtDerived *bd = reinterpret_cast<BjtDerived *>(dd + (size_t)kDevChunk * P.nd);
    const int S = P.S, W = P.W;
    const int ndev = P.nd + P.nc + P.nb;
    const int per = ndev * S;
    for (int c0 = blockIdx.x * kDevChunk; c0 < batch; c0 += gridDim.x * kDevChunk) {
        const int nc = min(kDevChunk, batch - c0);
        __syncthreads();
        for (int i = threadIdx.x; i < nc * P.nd; i += blockDim.x)
            diode_derive(diode_par + ((size_t)c0 * P.nd + i) * YHB_DIODE_NPAR, dd[i]);
        for (int i = threadIdx.x; i < nc * P.nb; i += blockDim.x)
            bjt_derive(bjt_par + ((size_t)c0 * P.nb + i) * YHB_BJT_NPAR, P.flags, bd[i]);
        __syncthreads();
        for (int it = threadIdx.x; it < nc * per; it += blockDim.x) {
            const int cl = it / per, r = it - cl * per;
            const int dev = r / S, s = r - dev * S;
            const size_t b = (size_t)c0 + cl;
            const double *v = vt + b * W, *vo = vtprev + b * W;
            if (dev < P.nd + P.nc) {
                const int *nodes = (dev < P.nd) ? P.diode_nodes + 2 * dev : P.cubic_nodes + 2 * (dev - P.nd);
                int n1 = nodes[0] - 1, n2 = nodes[1] - 1;
                double vd = (n1 >= 0 ? v[n1 * S + s] : 0.) - (n2 >= 0 ? v[n2 * S + s] : 0.);
                double vdold = (n1 >= 0 ? vo[n1 * S + s] : 0.) - (n2 >= 0 ? vo[n2 * S + s] : 0.);
                double g, I;
                double *o;
                if (dev < P.nd) {
                    diode_eval(dd[cl * P.nd + dev], vd, vdold, g, I);
                    o = diode_out + (b * P.nd + dev) * 2 * S + s;
                } else {
                    cubic_eval(cubic_par[b * P.nc + dev - P.nd], vd, g, I);
                    o = cubic_out + (b * P.nc + dev - P.nd) * 2 * S + s;
                }
                o[0] = g;
                o[S] = I;
            } else {
                const int q = dev - P.nd - P.nc;
                const int *nodes = P.bjt_nodes + 3 * q;
                int B = nodes[0] - 1, C = nodes[1] - 1, E = nodes[2] - 1;
                double out[kBjtWaves];
                bjt_eval(bd[cl * P.nb + q], B >= 0 ? v[B * S + s] : 0., C >= 0 ? v[C * S + s] : 0.,
                         E >= 0 ? v[E * S + s] : 0., B >= 0 ? vo[B * S + s] : 0., C >= 0 ? vo[C * S + s] : 0.,
                         E >= 0 ? vo[E * S + s] : 0., out);
                double *o = bjt_out + (b * P.nb + q) * kBjtWaves * S + s;
#pragma unroll
                for (int j = 0; j < kBjtWaves; ++j) o[(size_t)j * S] = out[j];
            }
        }
    }
}

}  // namespace yhb

extern "C" int yhb_stage_ifft(const yhb_plan *pl, int32_t batch, const double *V, double *vt, void *stream) {
    if (!pl || !V || !vt || batch < 1) { set_error("bad argument"); return -1; }
    k_stage_ifft<<<batch, 256, 0, (cudaStream_t)stream>>>(pl->d, batch, V, vt);
    YHB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int yhb_stage_device_eval(const yhb_plan *pl, const yhb_params *p, const double *vt, const double *vtprev,
                                     double *diode_out, double *bjt_out, double *cubic_out, void *stream) {
    if (int rc = check_params(pl, p)) return rc;
    const PlanDev &d = pl->d;
    if (!vt || !vtprev || (d.nd && !diode_out) || (d.nb && !bjt_out) || (d.nc && !cubic_out)) {
        set_error("missing array");
        return -1;
    }
    int dev = 0, sms = 0;
    YHB_CUDA(cudaGetDevice(&dev));
    YHB_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    // variant = chunk size (circuits per work item) and resident CTAs per SM; YHB_DEVSTAGE picks one for A/B runs
    static const int variant = getenv("YHB_DEVSTAGE") ? atoi(getenv("YHB_DEVSTAGE")) : 0;
    auto launch = [&](auto kern, int chunk, int minb) -> int {
        size_t smem = (size_t)chunk * (d.nd * sizeof(DiodeDerived) + d.nb * sizeof(BjtDerived));
        if (smem > 200 * 1024) { set_error("too many devices for the stage kernel"); return -1; }
        if (smem > 48 * 1024)
            YHB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int chunks = (p->batch + chunk - 1) / chunk;
        const int grid = std::min(chunks, minb * sms);  // persistent CTAs
        kern<<<grid, 256, smem, (cudaStream_t)stream>>>(d, p->batch, p->diode_par, p->bjt_par, p->cubic_par, vt, vtprev,
                                                        diode_out, bjt_out, cubic_out);
        YHB_CUDA(cudaGetLastError());
        return 0;
    };
    switch (variant) {  // measured on the 16 384-circuit DiffAmp shape: 4 CTAs/SM x chunks of 8 is the fastest
        case 1: return launch(k_stage_device_eval<32, 2>, 32, 2);
        case 2: return launch(k_stage_device_eval<8, 3>, 8, 3);
        case 3: return launch(k_stage_device_eval<16, 3>, 16, 3);
        default: return launch(k_stage_device_eval<8, 4>, 8, 4);
    }
}

extern "C" int yhb_stage_assemble(const yhb_plan *pl, const yhb_params *p, void *wsbuf, size_t ws_bytes, const double *V,
                                  const double *vtprev, const double *alpha, double *cacc, double *J, double *F,
                                  void *stream) {
    if (int rc = check_params(pl, p)) return rc;
    if (!V || !J || !F) { set_error("missing array"); return -1; }
    Layout L = carve(pl, p->batch, wsbuf);
    if (!wsbuf || ws_bytes < L.bytes) { set_error("workspace too small"); return -1; }
    const PlanDev &d = pl->d;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t n = (size_t)p->batch * d.W;
    if (int rc = launch_prepare(pl, p, L, st)) return rc;
    L.ws.V = const_cast<double *>(V);
    L.ws.F = F;
    L.ws.J = J;
    L.ws.cacc = cacc;
    L.ws.stage_first = vtprev == nullptr;
    if (vtprev) YHB_CUDA(cudaMemcpyAsync(L.ws.vtprev, vtprev, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    EvalArgs ea = eval_args(pl, p, nullptr, L);
    ea.stage_only = 1;
    ea.stage_alpha = alpha;
    const int smem = eval_smem_bytes(pl, L);
    if (int rc = ensure_eval_attr(smem)) return rc;
    k_eval<kEvalThreads><<<p->batch, kEvalThreads, smem, st>>>(d, L.ws, ea);
    dim3 ga(p->batch, (d.W + kAsmRows - 1) / kAsmRows);
    k_assemble<<<ga, 256, 0, st>>>(d, L.ws, ea);
    YHB_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int yhb_stage_gj(int32_t batch, int32_t n, const double *A, const double *F, double *x, void *stream) {
    if (batch < 1 || n < 1 || n > 127 || !A || !F || !x) { set_error("bad argument (1 <= n <= 127)"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int need = (n + 1 + 7) / 8;
    bool done = false;
#define YHB_CASE(t)                                                                                         \
    if (!done && need <= t) {                                                                               \
        size_t smem = (GjBufs<t>::kMatrixDoubles + (sizeof(GjBufs<t>) + 7) / 8 + (size_t)(n + 1)) * sizeof(double); \
        YHB_CUDA(cudaFuncSetAttribute(k_gj_stage<t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_gj_stage<t><<<batch, fused_threads(t), smem, st>>>(batch, n, A, F, x);                            \
        done = true;                                                                                        \
    }
    YHB_FOR_NT(YHB_CASE)
#undef YHB_CASE
    YHB_CUDA(cudaGetLastError());
    return 0;
}

// the panel Gauss-Jordan solver of the staged engines on its own: X[B][n][nrhs] = A^-1 Bm, A[B][n][n] row-major
namespace yhb {
__global__ void k_gjp_pack(int n, int nrhs, int ld, const double *A, const double *Bm, double *M) {
    const size_t b = blockIdx.y;
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n * ld; e += gridDim.x * blockDim.x) {
        const int r = e / ld, c = e - r * ld;
        double v = 0.;
        const int npc = gjp_npc(n);
        if (c < n) v = A[(b * n + r) * n + c];
        else if (c >= npc && c < npc + nrhs) v = Bm[(b * n + r) * nrhs + (c - npc)];
        M[b * (size_t)n * ld + e] = v;
    }
}
}  // namespace yhb

extern "C" int yhb_stage_gjp(int32_t batch, int32_t n, int32_t nrhs, const double *A, const double *Bm, double *X,
                             void *stream) {
    if (batch < 1 || n < 1 || nrhs < 1 || !A || !Bm || !X) { set_error("bad argument"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    const int npc = gjp_npc(n), ld = (npc + nrhs + 2) & ~1, lpad = (n + 7) & ~7;
    Carver c{nullptr, 0};
    auto carve_all = [&](Carver &cv, GjpBatch &g) {
        g.M = cv.take<double>((size_t)batch * n * ld);
        g.L = cv.take<double>((size_t)batch * kGjpNB * lpad);
        g.U = cv.take<double>((size_t)batch * kGjpNB * ld);
        g.T = cv.take<double>((size_t)batch * kGjpNB * kGjpNB);
        g.piv = cv.take<double>((size_t)batch * n);
        g.rowof = cv.take<int>((size_t)batch * n);
        g.used = cv.take<int>((size_t)batch * n);
        g.pp = cv.take<int>((size_t)batch * kGjpNB);
        g.fail = cv.take<int>(batch);
    };
    GjpBatch g;
    memset(&g, 0, sizeof(g));
    carve_all(c, g);
    char *buf = nullptr;
    YHB_CUDA(cudaMalloc(&buf, align_up(c.off)));
    Carver c2{buf, 0};
    carve_all(c2, g);
    g.m_stride = (size_t)n * ld;
    g.n = n;
    g.ld = ld;
    g.ncols = npc + nrhs;
    g.lpad = lpad;
    k_gjp_pack<<<dim3(std::max(1, std::min(256, n * ld / 256)), batch), 256, 0, st>>>(n, nrhs, ld, A, Bm, g.M);
    int rc = gjp_run(g, batch, st);
    if (rc == 0) {
        GjpFinish f;
        memset(&f, 0, sizeof(f));
        f.out0 = X;
        f.out0_stride = (size_t)n * nrhs;
        f.nout0 = nrhs;
        f.ld0 = nrhs;
        k_gjp_finish<<<dim3(std::min((n + 7) / 8, 148), batch), 256, 0, st>>>(g, f);
        if (cudaGetLastError() != cudaSuccess) { set_error("k_gjp_finish launch failed"); rc = -2; }
    }
    cudaStreamSynchronize(st);
    cudaFree(buf);
    return rc;
}

extern "C" int yhb_stage_lu(int32_t batch, int32_t W, double *J, const double *F, double *x, void *stream) {
    if (batch < 1 || W < 1 || !J || !F || !x) { set_error("bad argument"); return -1; }
    k_lu_stage<<<batch, kLuThreads, (W + 1) * sizeof(double), (cudaStream_t)stream>>>(batch, W, J, F, x);
    YHB_CUDA(cudaGetLastError());
    return 0;
}

// ---- post-processing: phasors -> time-domain waveforms (convert_to_time :125-135, tests/HB_TwoTones.ipynb cell 3) ----
namespace yhb {
__global__ void __launch_bounds__(256) k_to_time(int N, int K1, const double *Vf, const double *freqs, int nt,
                                                 const double *t, double *x) {
    extern __shared__ double sh[];  // phasors of this (circuit, node): re[K1], im[K1], then f[K1]
    const int bn = blockIdx.y, b = bn / N;
    double *re = sh, *im = sh + K1, *f = sh + 2 * K1;
    for (int k = threadIdx.x; k < K1; k += blockDim.x) {
        re[k] = Vf[2 * ((size_t)bn * K1 + k)];
        im[k] = Vf[2 * ((size_t)bn * K1 + k) + 1];
        f[k] = freqs[(size_t)b * K1 + k];
    }
    __syncthreads();
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= nt) return;
    const double ts = t[s];
    double acc = re[0];
    for (int k = 1; k < K1; ++k) {  // the reference's loop, term by term in its order
        double sn, cs;
        sincos(kTwoPi * f[k] * ts, &sn, &cs);
        acc = acc + re[k] * cs - im[k] * sn;
    }
    x[(size_t)bn * nt + s] = acc;
}
}  // namespace yhb

extern "C" int yhb_to_time(int32_t batch, int32_t N, int32_t K1, const double *Vf, const double *freqs, int32_t nt,
                           const double *t, double *x, void *stream) {
    if (batch < 1 || N < 1 || K1 < 1 || nt < 1 || !Vf || !freqs || !t || !x) { set_error("bad argument"); return -1; }
    const size_t smem = 3 * (size_t)K1 * sizeof(double);
    if (smem > 200 * 1024) { set_error("yhb_to_time: too many bins"); return -1; }
    if (smem > 48 * 1024) YHB_CUDA(cudaFuncSetAttribute(k_to_time, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_to_time<<<dim3((nt + 255) / 256, batch * N), 256, smem, (cudaStream_t)stream>>>(N, K1, Vf, freqs, nt, t, x);
    YHB_CUDA(cudaGetLastError());
    return 0;
}

// ---- FP64 pipe peak (roofline denominator of the LU stage) -----------------------------------------
namespace yhb {
__global__ void __launch_bounds__(256) k_fp64_peak(double *out, int iters) {
    double a0 = threadIdx.x * 1e-9, a1 = a0 + 1., a2 = a0 + 2., a3 = a0 + 3., a4 = a0 + 4., a5 = a0 + 5., a6 = a0 + 6.,
           a7 = a0 + 7.;
    const double m = 0.999999, c = 1e-7;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
        a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}
}  // namespace yhb

extern "C" int yhb_fp64_peak(double *tflops, void *stream) {
    if (!tflops) { set_error("null argument"); return -1; }
    cudaStream_t st = (cudaStream_t)stream;
    cudaDeviceProp prop;
    int dev = 0;
    YHB_CUDA(cudaGetDevice(&dev));
    YHB_CUDA(cudaGetDeviceProperties(&prop, dev));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 15;
    double *buf = nullptr;
    YHB_CUDA(cudaMalloc(&buf, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(a, st);
        k_fp64_peak<<<blocks, threads, 0, st>>>(buf, iters);
        cudaEventRecord(b, st);
        cudaEventSynchronize(b);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, a, b);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(a);
    cudaEventDestroy(b);
    cudaFree(buf);
    YHB_CUDA(cudaGetLastError());
    *tflops = 2.0 * 8 * (double)iters * blocks * threads / (best * 1e-3) / 1e12;
    return 0;
}

// ---- phase timing of the fused engine (debug builds only: -DYHB_TIMING) -------------------------------
extern "C" int yhb_debug_timing(unsigned long long *out64, int32_t reset) {
#ifdef YHB_TIMING
    if (out64) YHB_CUDA(cudaMemcpyFromSymbol(out64, yhb::g_timing, sizeof(unsigned long long) * 64));
    if (reset) {
        unsigned long long z[64] = {0};
        YHB_CUDA(cudaMemcpyToSymbol(yhb::g_timing, z, sizeof(z)));
    }
    return 0;
#else
    (void)out64; (void)reset;
    set_error("libyhb was built without -DYHB_TIMING");
    return -1;
#endif
}
