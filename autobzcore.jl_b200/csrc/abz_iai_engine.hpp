// Host-side IAI engine: nested adaptive Gauss-Kronrod (7,15) with the reference's control flow,
// batched level-synchronously for the device.  Pure C++ (no CUDA) so that the same state machine can
// be driven by the device backend (abz_api.cu) and, in tests/, by a CPU backend.
//
// Restates do_solve(::FourierIntegrand, lims, ::NestedQuad) + init_nest (src/fourier.jl:432-510) on
// top of QuadGK.jl's do_quadgk / adapt / refine / evalrule and DataStructures.jl's binary heap with
// Base.Reverse on Segment.E (call site of IteratedIntegration.auxquadgk: src/algorithms.jl:236-237).
// Every 1-D adaptive integral is an independent state machine whose trajectory depends only on its
// own integrand values and tolerance; evaluating all live panels of all live integrals in one device
// batch per round leaves every accept/refine decision - and hence EvalCounter's numevals
// (src/fourier.jl:516-523) - unchanged with respect to the sequential recursion.
//
// Multi-rank: the 15 nodes of every panel of the OUTERMOST integral are dealt round-robin to the ranks;
// each rank runs the inner integrals of its own nodes, and when its local work is exhausted all ranks meet in ONE small
// sum-allreduce per outer refinement step (values of the outstanding outer panels, zeros for nodes a rank does not own,
// plus evaluation counts), after which every rank makes the identical accept/refine decision.  x + 0 is exact, so the
// result is bit-identical to the single-rank solve.
//
// With leaf_tasks = true the innermost (level-0) integrals are not run by this state machine: each is
// handed to the backend as ONE task (slot, a, b, atol) that returns (I, E, numevals) - the device runs
// the whole 1-D adaptive loop with one warp per task (abz_iai.cuh, iai_leaf_kernel).
//
// Look-ahead on the outermost integral (speculate = true, children of the outermost panels handed to the backend as whole
// tasks): QuadGK refines one panel at a time, so a round would carry the 2 x 15 tasks of ONE bisection and the device would idle between
// rounds.  When the outermost integral bisects its worst panel, the engine also starts the bisections QuadGK is most likely to ask for
// next - the four quarters of that panel (its halves usually stay the worst ones while a feature is being resolved) and the panel that
// is next in the heap - and parks the results in a cache keyed by (a, b).  When QuadGK's own order reaches such a panel the halves are
// taken from the cache (or claimed while still in flight) instead of being evaluated; what is never reached is dropped and its
// evaluations are not counted.  Every accept / refine decision is taken on exactly the values and in exactly the order of the
// sequential algorithm, so the integral, the error estimate and numevals are unchanged - only the number of rounds drops
// (C3, eta = 1e-4: 80 -> 32 rounds, 105 -> 58 ms; profiles/r02_c3_lookahead_timing.log).
#pragma once
#include <cmath>
#include <cstdint>
#include <deque>
#include <string>
#include <vector>

#if defined(__CUDACC__)
#define ABZ_HD __host__ __device__
#else
#define ABZ_HD
#endif

namespace abz_iai {

struct cplx { double re, im; };

// round-to-nearest products/sums that the compiler must not contract into FMAs, so that host and
// device combine panels with bit-identical arithmetic (QuadGK.evalrule's operation order)
ABZ_HD inline double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
ABZ_HD inline double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
ABZ_HD inline cplx cadd_rn(cplx a, cplx b) { return cplx{add_rn(a.re, b.re), add_rn(a.im, b.im)}; }
ABZ_HD inline cplx csub_rn(cplx a, cplx b) { return cplx{add_rn(a.re, -b.re), add_rn(a.im, -b.im)}; }
ABZ_HD inline cplx cscale_rn(cplx a, double w) { return cplx{mul_rn(a.re, w), mul_rn(a.im, w)}; }

// QUADPACK qk15 abscissae (x <= 0 half, as QuadGK.kronrod(7) orders them), Kronrod and Gauss weights
#define ABZ_GK_X0 (-0.991455371120812639206854697526329)
#define ABZ_GK_X1 (-0.949107912342758524526189684047851)
#define ABZ_GK_X2 (-0.864864423359769072789712788640926)
#define ABZ_GK_X3 (-0.741531185599394439863864773280788)
#define ABZ_GK_X4 (-0.586087235467691130294144838258730)
#define ABZ_GK_X5 (-0.405845151377397166906606412076961)
#define ABZ_GK_X6 (-0.207784955007898467600689403773245)
#define ABZ_GK_W0 0.022935322010529224963732008058970
#define ABZ_GK_W1 0.063092092629978553290700663189204
#define ABZ_GK_W2 0.104790010322250183839876322541518
#define ABZ_GK_W3 0.140653259715525918745189590510238
#define ABZ_GK_W4 0.169004726639267902826583426598550
#define ABZ_GK_W5 0.190350578064785409913256402421014
#define ABZ_GK_W6 0.204432940075298892414161999234649
#define ABZ_GK_W7 0.209482141084727828012999174891714
#define ABZ_GK_G0 0.129484966168869693270611432679082
#define ABZ_GK_G1 0.279705391489276667901467771423780
#define ABZ_GK_G2 0.381830050505118944950369775488975
#define ABZ_GK_G3 0.417959183673469387755102040816327

// offsets (1 + x) and (1 - x) in QuadGK.evalrule's evaluation order:
// (x2+,x2-),(x1+,x1-),(x4+,x4-),(x3+,x3-),(x6+,x6-),(x5+,x5-), centre, (x7+,x7-)
ABZ_HD inline double gk_off(int j) {
    switch (j) {
        case 0: return 1.0 + ABZ_GK_X1;
        case 1: return 1.0 - ABZ_GK_X1;
        case 2: return 1.0 + ABZ_GK_X0;
        case 3: return 1.0 - ABZ_GK_X0;
        case 4: return 1.0 + ABZ_GK_X3;
        case 5: return 1.0 - ABZ_GK_X3;
        case 6: return 1.0 + ABZ_GK_X2;
        case 7: return 1.0 - ABZ_GK_X2;
        case 8: return 1.0 + ABZ_GK_X5;
        case 9: return 1.0 - ABZ_GK_X5;
        case 10: return 1.0 + ABZ_GK_X4;
        case 11: return 1.0 - ABZ_GK_X4;
        case 12: return 1.0;
        case 13: return 1.0 + ABZ_GK_X6;
        default: return 1.0 - ABZ_GK_X6;
    }
}
// node j of evalrule on [a, b]
ABZ_HD inline double gk_node(double a, double b, int j) {
    double s = mul_rn(0.5, add_rn(b, -a));
    return add_rn(a, mul_rn(gk_off(j), s));
}

// QuadGK.evalrule: 15 values in gk_node order -> Kronrod estimate I and (Kronrod - Gauss) difference D;
// the panel error is |D|
ABZ_HD inline void gk_combine(double a, double b, const cplx* f, cplx* I, cplx* D) {
    double s = mul_rn(0.5, add_rn(b, -a));
    cplx fg = cadd_rn(f[0], f[1]), fk = cadd_rn(f[2], f[3]);
    cplx Ig = cscale_rn(fg, ABZ_GK_G0);
    cplx Ik = cadd_rn(cscale_rn(fg, ABZ_GK_W1), cscale_rn(fk, ABZ_GK_W0));
    fg = cadd_rn(f[4], f[5]); fk = cadd_rn(f[6], f[7]);
    Ig = cadd_rn(Ig, cscale_rn(fg, ABZ_GK_G1));
    Ik = cadd_rn(Ik, cadd_rn(cscale_rn(fg, ABZ_GK_W3), cscale_rn(fk, ABZ_GK_W2)));
    fg = cadd_rn(f[8], f[9]); fk = cadd_rn(f[10], f[11]);
    Ig = cadd_rn(Ig, cscale_rn(fg, ABZ_GK_G2));
    Ik = cadd_rn(Ik, cadd_rn(cscale_rn(fg, ABZ_GK_W5), cscale_rn(fk, ABZ_GK_W4)));
    Ig = cadd_rn(Ig, cscale_rn(f[12], ABZ_GK_G3));
    Ik = cadd_rn(Ik, cadd_rn(cscale_rn(f[12], ABZ_GK_W7), cscale_rn(cadd_rn(f[13], f[14]), ABZ_GK_W6)));
    cplx Iks = cscale_rn(Ik, s), Igs = cscale_rn(Ig, s);
    *I = Iks;
    *D = csub_rn(Iks, Igs);
}

// value of the user integrand from the device's per-node quantity y (tr G or tr H):
// vkind 0: y;  1: -Im(y)/pi (aps_example/aps_example.jl:30);  2: a*y + b (test/fourier.jl:41)
ABZ_HD inline cplx post_value(int vkind, cplx y, cplx la, cplx lb) {
    if (vkind == 0) return y;
    if (vkind == 1) return cplx{-y.im / 3.14159265358979323846, 0.0};
    cplx t{add_rn(mul_rn(la.re, y.re), -mul_rn(la.im, y.im)), add_rn(mul_rn(la.re, y.im), mul_rn(la.im, y.re))};
    return cadd_rn(t, lb);
}

// ---- iterated limits (IteratedIntegration.CubicLimits / TetrahedralLimits / any AbstractIteratedLimits) ----------------------
// kind 2 = general limits served by the caller (the reference's `segments` + `fixandeliminate`, e.g. the polyhedral IBZ of
// ext/SymmetryReduceBZExt.jl:33-58): fn(dim, x_fixed, segs, maxseg, user) writes the ascending breakpoints of the variable of
// `dim` (1-based, dim = ndim outermost) given the outer variables fixed so far (x_fixed[0] = outermost) and returns their number.
typedef int32_t (*limits_fn)(int32_t dim, const double* x_fixed, double* segs, int32_t maxseg, void* user);
constexpr int MAXSEG = 64;
struct Limits {
    int kind = 0, nd = 0;   // kind 0: x_d in [a_d, b_d];  1: 0 <= x_d <= a_d s, then s <- x_d / a_d;  2: fn
    double a[3] = {0, 0, 0}, b[3] = {0, 0, 0}, s = 1.0;
    limits_fn fn = nullptr; void* user = nullptr;
    double fixed[3] = {0, 0, 0}; int nfixed = 0;
    // breakpoints of the outermost remaining variable (limit_iterate / segments); false: the callback failed
    bool segments(std::vector<double>& out) const {
        out.clear();
        if (kind == 0) { out.push_back(a[nd - 1]); out.push_back(b[nd - 1]); return true; }
        if (kind == 1) { out.push_back(0.0); out.push_back(a[nd - 1] * s); return true; }
        double buf[MAXSEG];
        const int n = fn ? fn(nd, fixed, buf, MAXSEG, user) : -1;
        if (n < 2 || n > MAXSEG) return false;
        for (int i = 0; i < n; i++) { if (i > 0 && !(buf[i] >= buf[i - 1])) return false; out.push_back(buf[i]); }
        return true;
    }
    Limits fix(double x) const {
        Limits r = *this;
        r.nd = nd - 1;
        if (kind == 1) r.s = x / a[nd - 1];
        if (kind == 2 && nfixed < 3) r.fixed[r.nfixed++] = x;
        return r;
    }
};

struct Seg { double E, a, b; cplx I; };

// DataStructures.jl BinaryHeap with Base.Reverse: lt(x, y) = y.E < x.E
inline bool seg_lt_rev(const Seg& x, const Seg& y) { return y.E < x.E; }
inline void heap_percolate_down(std::vector<Seg>& xs, size_t i, Seg x) {   // 1-based i
    const size_t n = xs.size();
    for (;;) {
        size_t l = 2 * i;
        if (l > n) break;
        size_t r = l + 1;
        size_t j = (r > n || seg_lt_rev(xs[l - 1], xs[r - 1])) ? l : r;
        if (!seg_lt_rev(xs[j - 1], x)) break;
        xs[i - 1] = xs[j - 1];
        i = j;
    }
    xs[i - 1] = x;
}
inline void heap_percolate_up(std::vector<Seg>& xs, size_t i, Seg x) {
    for (;;) {
        size_t j = i / 2;
        if (j < 1) break;
        if (!seg_lt_rev(x, xs[j - 1])) break;
        xs[i - 1] = xs[j - 1];
        i = j;
    }
    xs[i - 1] = x;
}
inline Seg heap_pop(std::vector<Seg>& xs) {
    Seg x = xs[0];
    Seg y = xs.back();
    xs.pop_back();
    if (!xs.empty()) heap_percolate_down(xs, 1, y);
    return x;
}
inline void heap_push(std::vector<Seg>& xs, const Seg& x) {
    xs.push_back(x);
    heap_percolate_up(xs, xs.size(), x);
}

// ---- one round of device work ----------------------------------------------------------------------
struct Round {
    // contractions queued by the panels started in the previous round (workspace_contract!, src/fourier.jl:478)
    std::vector<double> c3_x; std::vector<int64_t> c3_slot;
    std::vector<double> c2_x; std::vector<int64_t> c2_parent, c2_slot;
    // innermost panels: evaluate the 15 nodes of [a, b] on the series in level-1 slot `slot` and combine
    std::vector<double> seg_a, seg_b; std::vector<int64_t> seg_slot;
    std::vector<cplx> seg_I, seg_D;                       // outputs
    // leaf tasks: whole innermost adaptive integrals
    std::vector<double> task_a, task_b, task_atol; std::vector<int64_t> task_slot;
    std::vector<cplx> task_I; std::vector<double> task_E; std::vector<int64_t> task_ne;   // outputs
    // middle tasks (3-d solves): whole level-1 adaptive integrals over x2 on the series in level-2 slot `slot`
    std::vector<double> mid_a, mid_b, mid_atol; std::vector<int64_t> mid_slot;
    std::vector<cplx> mid_I; std::vector<double> mid_E; std::vector<int64_t> mid_ne;       // outputs (mid_ne: innermost evaluations)
    void clear_inputs() {
        c3_x.clear(); c3_slot.clear(); c2_x.clear(); c2_parent.clear(); c2_slot.clear();
        seg_a.clear(); seg_b.clear(); seg_slot.clear();
        task_a.clear(); task_b.clear(); task_atol.clear(); task_slot.clear();
        mid_a.clear(); mid_b.clear(); mid_atol.clear(); mid_slot.clear();
    }
};

enum { IAI_OK = 0, IAI_E_NAN = -4, IAI_E_ARENA = -2, IAI_E_STALL = -7, IAI_E_LIMITS = -8 };

// Backend concept:  int lanes()                     number of rounds that may be in flight at once (>= 1);
//                   int submit(int lane, Round&)    starts the queued inputs of one round (may return before the work is done);
//                   int wait(int lane, Round&)      completes it and fills the outputs; both return 0 or an error code;
//                   int exchange(double* buf, size_t n)  in-place sum over ranks (only called when nranks > 1).
// Lanes: the children of the outermost integral's panels are dealt to the lanes round-robin and everything below a child
// stays in its lane, so every lane is a sequence of rounds of its own (on the device: its own stream), and the lanes overlap.
// Every 1-D integral still sees exactly its own evaluations in its own order: results and numevals do not depend on the
// number of lanes.
template <class Backend>
class Engine {
public:
    // mid_tasks: in 3-d solves over CubicLimits / TetrahedralLimits the middle integrals (one per node of the outermost panels) are
    // handed to the backend whole as well (abz_iai.cuh, iai_mid_kernel): a round is then one refinement step of the OUTERMOST integral
    Engine(Backend& be, int ndim, const Limits& lims, double atol, double rtol, int64_t maxevals, int64_t cap2,
           int64_t cap1, bool leaf_tasks, int rank = 0, int nranks = 1, bool mid_tasks = false, bool speculate = false)
        : be_(be), ndim_(ndim), lims_(lims), atol_(atol), rtol_(rtol), maxevals_(maxevals),
          leaf_tasks_(leaf_tasks && ndim >= 2), mid_tasks_(mid_tasks && leaf_tasks && ndim == 3 && lims.kind != 2),
          rank_(rank), nranks_(ndim >= 2 ? nranks : 1) {
        // look-ahead needs every child of an outermost panel to be one backend task (so that its evaluations can be kept apart)
        // (several ranks: look-ahead panels are shared panels like the regular ones - their node values meet in the same allreduce, every
        // rank parks / claims / drops them identically, and a failure on one rank reaches all of them through that exchange)
        spec_on_ = speculate && lims.kind != 2 && ((ndim == 3 && mid_tasks_) || (ndim == 2 && leaf_tasks_));
        for (int64_t i = cap2 - 1; i >= 0; i--) free2_.push_back(i);
        for (int64_t i = cap1 - 1; i >= 0; i--) free1_.push_back(i);
        L_ = be.lanes() < 1 ? 1 : be.lanes();
        q_seg_.resize(L_); q_task_.resize(L_); fl_seg_.resize(L_); fl_task_.resize(L_);
        q_mid_.resize(L_); fl_mid_.resize(L_);
        cur_.resize(L_); next_.resize(L_); inflight_.assign(L_, 0);
    }

    int64_t numevals = 0, rounds = 0, exchanges = 0;   // numevals: all ranks' evaluations once the solve has finished
    int64_t spec_started = 0, spec_used = 0;           // look-ahead: half-panels started ahead of QuadGK's order / later consumed
    int spec_depth = 1;                                // panels of the heap bisected ahead per refinement of the outermost integral
    int spec_policy = 3;                               // 1: the panel(s) next in the heap, 2: the quarters of the panel being bisected, 3: both
    cplx result{0, 0};
    double result_err = 0;
    std::string error;

    int run() {
        std::vector<double> segs;
        if (!lims_.segments(segs)) return limits_error();
        int root = new_integral(ndim_ - 1, lims_, atol_, -1, -1, -1, -1, 0);
        int rc = start_initial(root, segs);
        if (rc) return rc;
        int local_rc = IAI_OK;
        while (!done_) {
            // start a round in every idle lane that has work queued
            for (int g = 0; g < L_ && !local_rc; g++) {
                if (inflight_[g] || (q_seg_[g].empty() && q_task_[g].empty() && q_mid_[g].empty())) continue;
                rounds++;
                cur_[g].clear_inputs();
                std::swap(cur_[g], next_[g]);           // cur_ = inputs queued so far, next_ = empty
                fl_seg_[g].clear(); fl_task_[g].clear(); fl_mid_[g].clear();
                fl_seg_[g].swap(q_seg_[g]); fl_task_[g].swap(q_task_[g]); fl_mid_[g].swap(q_mid_[g]);
                rc = be_.submit(g, cur_[g]);
                if (rc) { local_rc = rc; break; }
                inflight_[g] = 1;
                fifo_.push_back(g);
            }
            if (local_rc || fifo_.empty()) {
                drain();
                // local work exhausted: single rank = stalled; several ranks = meet the others (one allreduce)
                if (nranks_ == 1) { if (local_rc) return local_rc; error = "IAI engine stalled"; return IAI_E_STALL; }
                rc = exchange_step(local_rc);
                if (rc) return rc;
                continue;
            }
            const int g = fifo_.front();
            fifo_.pop_front();
            rc = be_.wait(g, cur_[g]);
            inflight_[g] = 0;
            if (rc) { local_rc = rc; continue; }
            const std::vector<Item>& segs = fl_seg_[g];
            const std::vector<Item>& tasks = fl_task_[g];
            const Round& R = cur_[g];
            numevals += 15 * (int64_t)segs.size();
            for (size_t i = 0; i < segs.size() && !rc; i++) {
                cplx D = R.seg_D[i];
                double E = std::hypot(D.re, D.im);
                rc = segment_done(segs[i].q, segs[i].pend, R.seg_I[i], E);
            }
            for (size_t i = 0; i < tasks.size() && !rc; i++)
                rc = task_done(tasks[i], R.task_I[i], R.task_E[i], R.task_ne[i]);
            const std::vector<Item>& mids = fl_mid_[g];
            for (size_t i = 0; i < mids.size() && !rc; i++)
                rc = task_done(mids[i], R.mid_I[i], R.mid_E[i], R.mid_ne[i]);
            if (rc) local_rc = rc;
        }
        drain();
        if (nranks_ > 1) {   // total evaluation count over the ranks (EvalCounter semantics of the whole solve)
            double ne = (double)numevals;
            rc = be_.exchange(&ne, 1);
            if (rc) return rc;
            numevals = (int64_t)ne;
        }
        return IAI_OK;
    }

private:
    struct Pend { double a, b; cplx vals[15]; int remaining, tag; bool shared; int spec; };   // spec: look-ahead cache entry or -1
    // a half-panel of the outermost integral evaluated ahead of QuadGK's order
    struct Spec { double a, b; bool live, done, bad; int claim_tag; Seg seg; int64_t ne; };
    struct Integral {
        int level; Limits lims; double atol; int64_t slot; int pq, ppend, pi;   // parent integral / panel / node
        int lane;                                                              // the lane its rounds run in
        std::vector<Seg> heap; cplx I; double E; int64_t numevals; Seg popped, s1, s2; bool has1, has2;
        int init_remaining;                                                    // initial segments still being evaluated
    };
    struct Item { int q, pend, i; int64_t slot; };

    Backend& be_;
    int ndim_; Limits lims_; double atol_, rtol_; int64_t maxevals_; bool leaf_tasks_, mid_tasks_;
    int rank_, nranks_; int64_t spawn_counter_ = 0;
    std::vector<std::pair<int, int>> shared_pends_;    // outstanding (integral, panel) of the outermost integral, creation order
    std::deque<Integral> ints_; std::vector<int> free_int_;
    std::deque<Pend> pends_; std::vector<int> free_pend_;
    std::vector<int64_t> free2_, free1_;
    int L_ = 1; int64_t lane_counter_ = 0;
    std::vector<std::vector<Item>> q_seg_, q_task_, fl_seg_, fl_task_, q_mid_, fl_mid_;   // per lane: queued / in flight
    std::vector<Round> cur_, next_;
    std::vector<char> inflight_;
    std::deque<int> fifo_;                                               // lanes in flight, oldest first
    bool done_ = false;
    bool spec_on_ = false;
    std::vector<Spec> spec_;

    // complete (and discard) whatever is still in flight: the backend's buffers must be quiescent before we return
    void drain() {
        while (!fifo_.empty()) { const int g = fifo_.front(); fifo_.pop_front(); be_.wait(g, cur_[g]); inflight_[g] = 0; }
    }

    int new_integral(int level, const Limits& lims, double atol, int64_t slot, int pq, int ppend, int pi, int lane) {
        int id;
        if (!free_int_.empty()) { id = free_int_.back(); free_int_.pop_back(); }
        else { id = (int)ints_.size(); ints_.emplace_back(); }
        Integral& q = ints_[id];
        q.level = level; q.lims = lims; q.atol = atol; q.slot = slot; q.pq = pq; q.ppend = ppend; q.pi = pi; q.lane = lane;
        q.heap.clear(); q.I = cplx{0, 0}; q.E = 0; q.numevals = 0; q.has1 = q.has2 = false;
        return id;
    }
    int new_pend(double a, double b, int tag) {
        int id;
        if (!free_pend_.empty()) { id = free_pend_.back(); free_pend_.pop_back(); }
        else { id = (int)pends_.size(); pends_.emplace_back(); }
        Pend& p = pends_[id];
        p.a = a; p.b = b; p.tag = tag; p.remaining = 15; p.shared = false; p.spec = -1;
        for (int i = 0; i < 15; i++) p.vals[i] = cplx{0.0, 0.0};
        return id;
    }
    int alloc_slot(int level, int64_t* slot) {
        std::vector<int64_t>& fr = (level == 2) ? free2_ : free1_;
        if (fr.empty()) { error = "IAI arena exhausted (too many live panels)"; return IAI_E_ARENA; }
        *slot = fr.back(); fr.pop_back();
        return IAI_OK;
    }
    void free_slot(int level, int64_t slot) { ((level == 2) ? free2_ : free1_).push_back(slot); }
    int limits_error() {
        error = "IAI: the limits callback failed (it must return at least 2 ascending breakpoints, at most 64)";
        return IAI_E_LIMITS;
    }
    // a whole-integral task of the backend finished: (I, E, evaluations).  Evaluations of a look-ahead panel stay with its cache entry.
    int task_done(const Item& it, cplx I, double E, int64_t ne) {
        const int sp = pends_[it.pend].spec;
        if (sp >= 0) spec_[sp].ne += ne; else numevals += ne;
        if (!std::isfinite(E)) {
            if (sp < 0) return nan_error(pends_[it.pend]);
            spec_[sp].bad = true;                     // reported only if QuadGK's own order reaches this panel
        }
        return child_done(it.q, it.pend, it.i, it.slot, I);
    }
    int find_spec(double a, double b) const {
        for (size_t k = 0; k < spec_.size(); k++) if (spec_[k].live && spec_[k].a == a && spec_[k].b == b) return (int)k;
        return -1;
    }
    int new_spec(double a, double b) {
        size_t k = 0;
        while (k < spec_.size() && spec_[k].live) k++;
        if (k == spec_.size()) spec_.emplace_back();
        spec_[k] = Spec{a, b, true, false, false, 0, Seg{0.0, a, b, cplx{0.0, 0.0}}, 0};
        return (int)k;
    }

    // do_quadgk's first pass: evalrule on every initial segment [segs[k], segs[k+1]]; panel k carries tag -k
    int start_initial(int qi, const std::vector<double>& segs) {
        const int ns = (int)segs.size() - 1;
        ints_[qi].init_remaining = ns;
        ints_[qi].heap.assign((size_t)ns, Seg{0.0, 0.0, 0.0, cplx{0.0, 0.0}});
        for (int k = 0; k < ns; k++) {
            int rc = start_segment(qi, segs[k], segs[k + 1], -k);
            if (rc) return rc;
        }
        return IAI_OK;
    }
    int nan_error(const Pend& p) {
        error = "integrand produced NaN/Inf in the interval (" + std::to_string(p.a) + ", " + std::to_string(p.b) + ")";
        return IAI_E_NAN;
    }

    // evalrule on [a, b] of integral q: innermost -> queue the panel; outer -> spawn 15 child integrals
    int start_segment(int qi, double a, double b, int tag, int spec = -1) {
        int pend = new_pend(a, b, tag);
        pends_[pend].spec = spec;
        const int level = ints_[qi].level;
        if (level == 0) {
            const int g = ints_[qi].lane;
            q_seg_[g].push_back(Item{qi, pend, 0, 0});
            next_[g].seg_a.push_back(a); next_[g].seg_b.push_back(b); next_[g].seg_slot.push_back(ints_[qi].slot);
            return IAI_OK;
        }
        const bool shared = (nranks_ > 1 && level == ndim_ - 1);
        if (shared) { pends_[pend].shared = true; shared_pends_.push_back({qi, pend}); }
        for (int i = 0; i < 15; i++) {
            if (shared && (spawn_counter_++ % nranks_) != rank_) continue;   // another rank owns this node
            const double x = gk_node(a, b, i);
            const Limits clims = ints_[qi].lims.fix(x);
            std::vector<double> csegs;
            if (!clims.segments(csegs)) return limits_error();
            const double ca = csegs.front(), cb = csegs.back();
            const double len = cb - ca;                       // len = segs[end] - segs[1] (src/fourier.jl:476)
            int64_t slot;
            int rc = alloc_slot(level, &slot);
            if (rc) return rc;
            // children of the outermost integral are dealt to the lanes; below that a child stays in its parent's lane
            const int g = (level == ndim_ - 1) ? (int)(lane_counter_++ % L_) : ints_[qi].lane;
            Round& nx = next_[g];
            if (level == 2) { nx.c3_x.push_back(x); nx.c3_slot.push_back(slot); }
            else { nx.c2_x.push_back(x); nx.c2_parent.push_back(ints_[qi].slot); nx.c2_slot.push_back(slot); }
            const double catol = ints_[qi].atol / len;        // inner abstol = abstol/len (src/fourier.jl:479-480)
            if (level == 2 && mid_tasks_ && csegs.size() == 2) {
                q_mid_[g].push_back(Item{qi, pend, i, slot});
                nx.mid_a.push_back(ca); nx.mid_b.push_back(cb); nx.mid_atol.push_back(catol); nx.mid_slot.push_back(slot);
                continue;
            }
            if (level == 1 && leaf_tasks_ && csegs.size() == 2) {
                q_task_[g].push_back(Item{qi, pend, i, slot});
                nx.task_a.push_back(ca); nx.task_b.push_back(cb); nx.task_atol.push_back(catol);
                nx.task_slot.push_back(slot);
                continue;
            }
            int child = new_integral(level - 1, clims, catol, slot, qi, pend, i, g);
            rc = start_initial(child, csegs);
            if (rc) return rc;
        }
        return IAI_OK;
    }

    // a child integral (node i of panel `pend` of integral qi) finished with value v
    int child_done(int qi, int pend, int i, int64_t slot, cplx v) {
        free_slot(ints_[qi].level, slot);
        Pend& p = pends_[pend];
        p.vals[i] = v;
        if (p.shared) return IAI_OK;           // combined in exchange_step once every rank has delivered its nodes
        if (--p.remaining > 0) return IAI_OK;
        cplx I, D;
        gk_combine(p.a, p.b, p.vals, &I, &D);
        return segment_done(qi, pend, I, std::hypot(D.re, D.im));
    }

    // all ranks: sum the outstanding outermost panels' node values (zeros for foreign nodes), then decide identically
    int exchange_step(int local_rc) {
        exchanges++;
        std::vector<std::pair<int, int>> sp;
        sp.swap(shared_pends_);
        std::vector<double> buf(30 * sp.size() + 1, 0.0);
        for (size_t k = 0; k < sp.size(); k++)
            for (int i = 0; i < 15; i++) { buf[30 * k + 2 * i] = pends_[sp[k].second].vals[i].re; buf[30 * k + 2 * i + 1] = pends_[sp[k].second].vals[i].im; }
        buf[30 * sp.size()] = local_rc ? 1.0 : 0.0;
        int rc = be_.exchange(buf.data(), buf.size());
        if (rc) return rc;
        if (buf[30 * sp.size()] != 0.0) {
            if (local_rc) return local_rc;
            error = "IAI: another rank reported an error (singular matrix or NaN/Inf in the integrand)";
            return IAI_E_NAN;
        }
        if (sp.empty()) { error = "IAI engine stalled"; return IAI_E_STALL; }
        for (size_t k = 0; k < sp.size(); k++) {
            Pend& p = pends_[sp[k].second];
            for (int i = 0; i < 15; i++) p.vals[i] = cplx{buf[30 * k + 2 * i], buf[30 * k + 2 * i + 1]};
            cplx I, D;
            gk_combine(p.a, p.b, p.vals, &I, &D);
            rc = segment_done(sp[k].first, sp[k].second, I, std::hypot(D.re, D.im));
            if (rc) return rc;
        }
        return IAI_OK;
    }

    int finish(int qi) {
        Integral& q = ints_[qi];
        cplx Iv = q.heap[0].I; double Ev = q.heap[0].E;
        for (size_t k = 1; k < q.heap.size(); k++) { Iv = cplx{Iv.re + q.heap[k].I.re, Iv.im + q.heap[k].I.im}; Ev += q.heap[k].E; }
        if (q.pq < 0) { result = Iv; result_err = Ev; done_ = true; return IAI_OK; }
        const int pq = q.pq, ppend = q.ppend, pi = q.pi; const int64_t slot = q.slot;
        free_int_.push_back(qi);
        return child_done(pq, ppend, pi, slot, Iv);
    }

    int refine(int qi) {
        Integral& q = ints_[qi];
        Seg s = heap_pop(q.heap);
        q.popped = s;
        const double mid = (s.a + s.b) / 2;
        q.has1 = q.has2 = false;
        if (!(spec_on_ && q.pq < 0)) {
            int rc = start_segment(qi, s.a, mid, 1);
            if (rc) return rc;
            return start_segment(qi, mid, s.b, 2);
        }
        // outermost integral with look-ahead: take the halves from the cache where they were started ahead of time
        int ntake = 0; int take_tag[2]; Seg take_seg[2];
        for (int t = 1; t <= 2; t++) {
            const double a = (t == 1) ? s.a : mid, b = (t == 1) ? mid : s.b;
            const int e = find_spec(a, b);
            if (e < 0) { int rc = start_segment(qi, a, b, t); if (rc) return rc; continue; }
            spec_used++;
            if (!spec_[e].done) { spec_[e].claim_tag = t; continue; }            // still in flight: delivered when it completes
            spec_[e].live = false;
            numevals += spec_[e].ne;
            if (spec_[e].bad || !std::isfinite(spec_[e].seg.E)) { Pend p{}; p.a = a; p.b = b; return nan_error(p); }
            take_tag[ntake] = t; take_seg[ntake] = spec_[e].seg; ntake++;
        }
        // start the bisection of the panel that is next in the heap (first of the top three whose halves are not cached yet), as long as
        // the arena keeps room for the next regular bisection (30 slots) on top of this one (30)
        const std::vector<int64_t>& fr = (q.level == 2) ? free2_ : free1_;
        if (spec_policy & 2) {
            // the four quarters of the panel being bisected: if one of its halves is QuadGK's next choice (the usual case while a
            // feature is being resolved) its bisection is already there
            const double m1 = (s.a + mid) / 2, m2 = (mid + s.b) / 2;
            const double qa[4] = {s.a, m1, mid, m2}, qb[4] = {m1, mid, m2, s.b};
            for (int h = 0; h < 2 && fr.size() >= 60; h++) {
                if (find_spec(qa[2 * h], qb[2 * h]) >= 0 || find_spec(qa[2 * h + 1], qb[2 * h + 1]) >= 0) continue;
                for (int k = 2 * h; k < 2 * h + 2; k++) {
                    int e = new_spec(qa[k], qb[k]);
                    int rc = start_segment(qi, qa[k], qb[k], 3, e);
                    if (rc) return rc;
                }
                spec_started += 2;
            }
        }
        if (spec_policy & 1) {
            size_t cand[3] = {0, 1, 2};
            if (q.heap.size() > 2 && seg_lt_rev(q.heap[2], q.heap[1])) { cand[1] = 2; cand[2] = 1; }
            int started = 0;
            for (size_t c = 0; c < 3 && cand[c] < q.heap.size() && started < spec_depth && fr.size() >= 60; c++) {
                const Seg t = q.heap[cand[c]];
                const double tm = (t.a + t.b) / 2;
                if (find_spec(t.a, tm) >= 0 || find_spec(tm, t.b) >= 0) continue;
                int e1 = new_spec(t.a, tm);
                int rc = start_segment(qi, t.a, tm, 3, e1);
                if (rc) return rc;
                int e2 = new_spec(tm, t.b);
                rc = start_segment(qi, tm, t.b, 3, e2);
                if (rc) return rc;
                spec_started += 2;
                started++;
            }
        }
        for (int k = 0; k < ntake; k++) { int rc = accept_half(qi, take_tag[k], take_seg[k]); if (rc) return rc; }
        return IAI_OK;
    }

    int segment_done(int qi, int pend, cplx Is, double Es) {
        const Pend p = pends_[pend];
        free_pend_.push_back(pend);
        const Seg seg{Es, p.a, p.b, Is};
        if (p.spec >= 0) {
            // look-ahead panel: park the result, or hand it over if QuadGK's order has reached it in the meantime
            Spec& e = spec_[p.spec];
            e.done = true; e.seg = seg;
            if (e.claim_tag == 0) return IAI_OK;
            e.live = false;
            numevals += e.ne;
            if (e.bad || !std::isfinite(Es)) return nan_error(p);
            return accept_half(qi, e.claim_tag, seg);
        }
        if (!std::isfinite(Es)) return nan_error(p);
        Integral& q = ints_[qi];
        if (p.tag <= 0) {
            // do_quadgk: I and E are left folds over the initial segments in their order; no subdivision when already converged
            // (finish() sums the vector in that same order), else heapify! (DataStructures: percolate_down from the last parent)
            q.heap[(size_t)(-p.tag)] = seg;
            if (--q.init_remaining > 0) return IAI_OK;
            q.I = q.heap[0].I; q.E = q.heap[0].E;
            for (size_t k = 1; k < q.heap.size(); k++) { q.I = cplx{q.I.re + q.heap[k].I.re, q.I.im + q.heap[k].I.im}; q.E += q.heap[k].E; }
            q.numevals = 15 * (int64_t)q.heap.size();
            if (q.numevals >= maxevals_ || q.E <= q.atol || q.E <= rtol_ * std::hypot(q.I.re, q.I.im)) return finish(qi);
            for (size_t i = q.heap.size() / 2; i >= 1; i--) heap_percolate_down(q.heap, i, q.heap[i - 1]);
            return refine(qi);
        }
        return accept_half(qi, p.tag, seg);
    }

    // one half (tag 1: left, 2: right) of the bisected panel of integral qi is known; both known -> QuadGK's refine bookkeeping
    int accept_half(int qi, int tag, const Seg& seg) {
        Integral& q = ints_[qi];
        if (tag == 1) { q.s1 = seg; q.has1 = true; } else { q.s2 = seg; q.has2 = true; }
        if (!(q.has1 && q.has2)) return IAI_OK;
        const Seg& s = q.popped;
        q.I = cplx{(q.I.re - s.I.re) + q.s1.I.re + q.s2.I.re, (q.I.im - s.I.im) + q.s1.I.im + q.s2.I.im};
        q.E = (q.E - s.E) + q.s1.E + q.s2.E;
        q.numevals += 30;
        heap_push(q.heap, q.s1);
        heap_push(q.heap, q.s2);
        if (q.E > q.atol && q.E > rtol_ * std::hypot(q.I.re, q.I.im) && q.numevals < maxevals_) return refine(qi);
        return finish(qi);
    }
};

}  // namespace abz_iai
