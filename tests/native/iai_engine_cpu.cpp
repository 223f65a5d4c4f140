// TEST DOUBLE (tests/ only): drives the product's host-side IAI engine (csrc/abz_iai_engine.hpp) with a CPU
// backend whose arithmetic is the oracle (liborc.so), so the C++ control flow - level-synchronous batching,
// slot management, leaf-task plumbing, numevals accounting - is checked on a box without a GPU against the
// oracle's sequential recursion (orc_iai).  Never linked into libautobz_cuda.so.
#include <complex.h>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

#include "abz_iai_engine.hpp"

extern "C" {
void orc_contract(const double* C, long rows, int M, int lo, double period, double x, double* out);
int orc_resolvent_trace_batch(const double* H, int n, long nk, int nw, const double* z, const double* sigma, double* out);
}

using namespace abz_iai;

struct CpuBackend {
    const double* coeffs; int n, ndim; int M[3], lo[3]; double period[3];
    int fkind, vkind; double z[2]; const double* sigma; cplx la, lb; double rtol; int64_t maxevals;
    std::map<int64_t, std::vector<double>> L2, L1;
    long launches = 0;
    int (*xfn)(double*, long, void*) = nullptr;
    int exchange(double* buf, size_t n) { return xfn ? xfn(buf, (long)n, nullptr) : -6; }

    cplx point(const std::vector<double>* c1, double x) {
        const long nn = (long)n * n;
        std::vector<double> H(2 * nn);
        orc_contract(c1 ? c1->data() : coeffs, nn, M[0], lo[0], period[0], x, H.data());
        cplx y{0, 0};
        if (fkind == 1) { for (int d = 0; d < n; d++) { y.re += H[2 * (d * (n + 1))]; y.im += H[2 * (d * (n + 1)) + 1]; } }
        else { double out[2]; orc_resolvent_trace_batch(H.data(), n, 1, 1, z, sigma, out); y = cplx{out[0], out[1]}; }
        return post_value(vkind, y, la, lb);
    }
    void panel(const std::vector<double>* c1, double a, double b, cplx* I, cplx* D) {
        cplx f[15];
        for (int j = 0; j < 15; j++) f[j] = point(c1, gk_node(a, b, j));
        gk_combine(a, b, f, I, D);
    }
    static double cabs_(cplx v) { return std::hypot(v.re, v.im); }

    void leaf_task(const std::vector<double>* c1, double a, double b, double atol, cplx* Iout, double* Eout, int64_t* neout) {
        std::vector<Seg> heap;
        cplx I, D;
        panel(c1, a, b, &I, &D);
        double E = cabs_(D);
        heap.push_back(Seg{E, a, b, I});
        int64_t ne = 15;
        bool go = std::isfinite(E) && !(ne >= maxevals || E <= atol || E <= rtol * cabs_(I));
        while (go) {
            Seg s = heap_pop(heap);
            double mid = (s.a + s.b) / 2;
            cplx I1, D1, I2, D2;
            panel(c1, s.a, mid, &I1, &D1);
            panel(c1, mid, s.b, &I2, &D2);
            double E1 = cabs_(D1), E2 = cabs_(D2);
            I = cplx{(I.re - s.I.re) + I1.re + I2.re, (I.im - s.I.im) + I1.im + I2.im};
            E = (E - s.E) + E1 + E2;
            ne += 30;
            heap_push(heap, Seg{E1, s.a, mid, I1});
            heap_push(heap, Seg{E2, mid, s.b, I2});
            if (!(std::isfinite(E1) && std::isfinite(E2))) { E = NAN; break; }
            go = (E > atol && E > rtol * cabs_(I) && ne < maxevals);
        }
        cplx Iv = heap[0].I; double Ev = heap[0].E;
        for (size_t k = 1; k < heap.size(); k++) { Iv.re += heap[k].I.re; Iv.im += heap[k].I.im; Ev += heap[k].E; }
        *Iout = Iv; *Eout = std::isfinite(E) ? Ev : NAN; *neout = ne;
    }

    // limits of the innermost variable below a middle node x2 (CubicLimits / TetrahedralLimits), as iai_mid_kernel derives them
    int lkind = 0; double lima[3] = {0, 0, 0}, limb[3] = {0, 0, 0};
    void mid_task(const std::vector<double>& c2, double a, double b, double atol, cplx* Iout, double* Eout, int64_t* ne_leaves) {
        const long rows = (long)n * n * M[0];
        int64_t nel = 0;
        bool bad = false;
        auto panel_mid = [&](double pa, double pb, cplx* I, cplx* D) {
            cplx f[15];
            for (int j = 0; j < 15; j++) {
                const double x2 = gk_node(pa, pb, j);
                std::vector<double> c1(2 * rows);
                orc_contract(c2.data(), rows, M[1], lo[1], period[1], x2, c1.data());
                double ca, cb;
                if (lkind == 0) { ca = lima[0]; cb = limb[0]; } else { ca = 0.0; cb = lima[0] * (x2 / lima[1]); }
                double E; int64_t ne;
                leaf_task(&c1, ca, cb, atol / (cb - ca), &f[j], &E, &ne);
                nel += ne;
                if (!std::isfinite(E)) bad = true;
            }
            gk_combine(pa, pb, f, I, D);
        };
        std::vector<Seg> heap;
        cplx I, D;
        panel_mid(a, b, &I, &D);
        double E = cabs_(D);
        heap.push_back(Seg{E, a, b, I});
        int64_t ne = 15;
        bool go = std::isfinite(E) && !bad && !(ne >= maxevals || E <= atol || E <= rtol * cabs_(I));
        while (go) {
            Seg s = heap_pop(heap);
            double mid = (s.a + s.b) / 2;
            cplx I1, D1, I2, D2;
            panel_mid(s.a, mid, &I1, &D1);
            panel_mid(mid, s.b, &I2, &D2);
            double E1 = cabs_(D1), E2 = cabs_(D2);
            I = cplx{(I.re - s.I.re) + I1.re + I2.re, (I.im - s.I.im) + I1.im + I2.im};
            E = (E - s.E) + E1 + E2;
            ne += 30;
            heap_push(heap, Seg{E1, s.a, mid, I1});
            heap_push(heap, Seg{E2, mid, s.b, I2});
            if (!(std::isfinite(E1) && std::isfinite(E2)) || bad) { E = NAN; break; }
            go = (E > atol && E > rtol * cabs_(I) && ne < maxevals);
        }
        cplx Iv = heap[0].I; double Ev = heap[0].E;
        for (size_t k = 1; k < heap.size(); k++) { Iv.re += heap[k].I.re; Iv.im += heap[k].I.im; Ev += heap[k].E; }
        *Iout = Iv; *Eout = (std::isfinite(E) && !bad) ? Ev : NAN; *ne_leaves = nel;
    }

    // lanes: IAI_CPU_LANES rounds in flight; the work of a round is done in wait(), so an engine that read a round's outputs
    // before waiting for it would see stale data
    int nlanes = 1;
    int lanes() const { return nlanes; }
    int submit(int, Round&) { return 0; }
    int wait(int, Round& R) { return run_round(R); }

    int run_round(Round& R) {
        const long nn = (long)n * n;
        for (size_t i = 0; i < R.c3_x.size(); i++) {
            long rows = nn * M[0] * M[1];
            std::vector<double> out(2 * rows);
            orc_contract(coeffs, rows, M[2], lo[2], period[2], R.c3_x[i], out.data());
            L2[R.c3_slot[i]] = std::move(out);
        }
        for (size_t i = 0; i < R.c2_x.size(); i++) {
            long rows = nn * M[0];
            std::vector<double> out(2 * rows);
            const double* src = (ndim == 2) ? coeffs : L2.at(R.c2_parent[i]).data();
            orc_contract(src, rows, M[1], lo[1], period[1], R.c2_x[i], out.data());
            L1[R.c2_slot[i]] = std::move(out);
        }
        const size_t ns = R.seg_a.size(), nt = R.task_a.size();
        R.seg_I.resize(ns); R.seg_D.resize(ns);
        for (size_t i = 0; i < ns; i++)
            panel(ndim >= 2 ? &L1.at(R.seg_slot[i]) : nullptr, R.seg_a[i], R.seg_b[i], &R.seg_I[i], &R.seg_D[i]);
        R.task_I.resize(nt); R.task_E.resize(nt); R.task_ne.resize(nt);
        for (size_t t = 0; t < nt; t++)      // whole innermost adaptive integral (what iai_leaf_kernel does per warp)
            leaf_task(&L1.at(R.task_slot[t]), R.task_a[t], R.task_b[t], R.task_atol[t], &R.task_I[t], &R.task_E[t], &R.task_ne[t]);
        const size_t nm = R.mid_a.size();
        R.mid_I.resize(nm); R.mid_E.resize(nm); R.mid_ne.resize(nm);
        for (size_t t = 0; t < nm; t++)      // whole middle adaptive integral (what iai_mid_kernel does per CTA)
            mid_task(L2.at(R.mid_slot[t]), R.mid_a[t], R.mid_b[t], R.mid_atol[t], &R.mid_I[t], &R.mid_E[t], &R.mid_ne[t]);
        launches++;
        return 0;
    }
};

extern "C" int iai_cpu_solve(const double* coeffs, int n, int ndim, const int* M, const int* lo, const double* period, int lkind,
                             const double* la, const double* lb, int fkind, int vkind, const double* z, const double* sigma,
                             const double* lin, double atol, double rtol, long maxevals, int leaf_tasks, long cap2, long cap1,
                             int rank, int nranks, int (*xfn)(double*, long, void*), double* out, long* stats) {
    CpuBackend be;
    be.coeffs = coeffs; be.n = n; be.ndim = ndim;
    for (int d = 0; d < 3; d++) { be.M[d] = d < ndim ? M[d] : 1; be.lo[d] = d < ndim ? lo[d] : 0; be.period[d] = d < ndim ? period[d] : 1.0; }
    be.fkind = fkind; be.vkind = vkind; be.z[0] = z ? z[0] : 0; be.z[1] = z ? z[1] : 0; be.sigma = sigma;
    be.la = cplx{lin ? lin[0] : 1.0, lin ? lin[1] : 0.0}; be.lb = cplx{lin ? lin[2] : 0.0, lin ? lin[3] : 0.0};
    be.rtol = rtol; be.maxevals = maxevals; be.xfn = xfn;
    if (const char* e = getenv("IAI_CPU_LANES")) be.nlanes = atoi(e) > 0 ? atoi(e) : 1;
    Limits lims; lims.kind = lkind; lims.nd = ndim; lims.s = 1.0;
    for (int d = 0; d < ndim; d++) { lims.a[d] = la[d]; lims.b[d] = lb ? lb[d] : 0.0; }
    be.lkind = lkind;
    for (int d = 0; d < ndim; d++) { be.lima[d] = lims.a[d]; be.limb[d] = lims.b[d]; }
    // leaf_tasks & 3: 0 host-driven panels, 1 innermost integrals as tasks, 2 middle integrals as tasks too; | 4: look-ahead on the
    // outermost integral (stats[4], stats[5] = half-panels started ahead / consumed)
    const int mode = leaf_tasks & 3;
    Engine<CpuBackend> eng(be, ndim, lims, atol, rtol, maxevals, cap2, cap1, mode != 0, rank, nranks, mode == 2, (leaf_tasks & 4) != 0);
    if (leaf_tasks & 8) eng.spec_depth = 2;              // | 8: two panels ahead
    if (leaf_tasks & 48) eng.spec_policy = (leaf_tasks >> 4) & 3;   // | 16: heap-top policy, | 32: quarters, | 48: both
    int rc = eng.run();
    stats[0] = eng.numevals; stats[1] = eng.rounds; stats[2] = be.launches; stats[3] = eng.exchanges;
    stats[4] = eng.spec_started; stats[5] = eng.spec_used;
    if (rc) return rc;
    out[0] = eng.result.re; out[1] = eng.result.im; out[2] = eng.result_err;
    return 0;
}

// the same engine over general iterated limits served by a callback (abz_iai_solve_general's control flow on the CPU backend)
extern "C" int iai_cpu_solve_general(const double* coeffs, int n, int ndim, const int* M, const int* lo, const double* period,
                                     limits_fn lfn, void* luser, int fkind, int vkind, const double* z, const double* sigma,
                                     const double* lin, double atol, double rtol, long maxevals, int leaf_tasks, long cap2, long cap1,
                                     int rank, int nranks, int (*xfn)(double*, long, void*), double* out, long* stats) {
    CpuBackend be;
    be.coeffs = coeffs; be.n = n; be.ndim = ndim;
    for (int d = 0; d < 3; d++) { be.M[d] = d < ndim ? M[d] : 1; be.lo[d] = d < ndim ? lo[d] : 0; be.period[d] = d < ndim ? period[d] : 1.0; }
    be.fkind = fkind; be.vkind = vkind; be.z[0] = z ? z[0] : 0; be.z[1] = z ? z[1] : 0; be.sigma = sigma;
    be.la = cplx{lin ? lin[0] : 1.0, lin ? lin[1] : 0.0}; be.lb = cplx{lin ? lin[2] : 0.0, lin ? lin[3] : 0.0};
    be.rtol = rtol; be.maxevals = maxevals; be.xfn = xfn;
    if (const char* e = getenv("IAI_CPU_LANES")) be.nlanes = atoi(e) > 0 ? atoi(e) : 1;
    Limits lims; lims.kind = 2; lims.nd = ndim; lims.s = 1.0; lims.fn = lfn; lims.user = luser;
    Engine<CpuBackend> eng(be, ndim, lims, atol, rtol, maxevals, cap2, cap1, leaf_tasks != 0, rank, nranks);
    int rc = eng.run();
    stats[0] = eng.numevals; stats[1] = eng.rounds; stats[2] = be.launches; stats[3] = eng.exchanges;
    if (rc) return rc;
    out[0] = eng.result.re; out[1] = eng.result.im; out[2] = eng.result_err;
    return 0;
}
