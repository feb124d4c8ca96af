// TEST-ONLY host build of the kernel's per-walker engine (rv_core.cuh / rv_loglik.cuh, thread-per-walker
// mapping PL == P, no shuffles).  It lets the CPU test-suite exercise the exact device source for logic
// errors where no GPU is present.  It is NOT part of the product library and is never loaded by it.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../rvel_mcmc_b200/csrc/rv_loglik.cuh"
#include "../../rvel_mcmc_b200/csrc/rv_core_g.cuh"
#include "../../rvel_mcmc_b200/csrc/rv_model.h"

namespace {
struct HostFetch {
    unsigned long long* ctr;
    template <class G> long long operator()(const G&) { return (long long)((*ctr)++); }
};
struct HostAll { bool operator()(bool f) const { return f; } };

template <class WK, int NCOORD>
void run_with(const rv::LoglikArgs& a) {
    WK w;
    w.grp.init(0);
    std::vector<double> hist(WK::LANE_DOUBLES, 0.0);
    w.hist.p = hist.data(); w.hist.stride = 1;
    HostFetch f{a.item_counter};
    HostAll all;
    rv::run_items(w, a, a.ot, a.orv, a.oerr, f, all, true);
}

template <int P, int D>
void run(const rv::LoglikArgs& a) {
    if (a.model->dense_output) run_with<rv::WalkerG<P, D, P, 4>, P * D>(a);
    else run_with<rv::WalkerG<P, D, P, 0>, P * D>(a);
}
}  // namespace

static int g_monotone = 0, g_dense = 0, g_reverse = 0;
extern "C" void mirror_set_reverse_order(int v) { g_reverse = v; }   // items taken in reversed walker order (LoglikArgs::order)
extern "C" void mirror_set_monotone(int v) { g_monotone = v; }
extern "C" void mirror_set_dense(int v) { g_dense = v; }

extern "C" int mirror_loglik(int P, const double* fixed, int nvars, const int* fp, const int* fe, double hill, int dims,
                             const double* tf, const double* rvf, const double* ef, int nf,
                             const double* tb, const double* rvb, const double* eb, int nb, double npoints,
                             const double* theta, long long W, double* logp, int* status,
                             const double* times, int nt, double* rv_out, unsigned long long* counters) {
    rv::Model m;
    int rc = rv::build_model(&m, P, fixed, nvars, fp, fe, hill, dims);
    if (rc) return rc;
    m.monotone_backward = g_monotone;
    m.dense_output = g_dense;
    std::vector<double> ot(nf + nb), orv(nf + nb), oerr(nf + nb);
    for (int i = 0; i < nf; i++) { ot[i] = tf[i]; orv[i] = rvf[i]; oerr[i] = ef[i]; }
    for (int i = 0; i < nb; i++) { ot[nf + i] = tb[i]; orv[nf + i] = rvb[i]; oerr[nf + i] = eb[i]; }
    std::vector<double> part(2 * W, 0.0);
    std::vector<int> pst(2 * W, -1);
    unsigned long long ctr = 0, work[2] = {0, 0};
    rv::LoglikArgs a;
    memset(&a, 0, sizeof a);
    a.model = &m; a.theta = theta; a.W = W;
    a.ot = ot.data(); a.orv = orv.data(); a.oerr = oerr.data(); a.nf = nf; a.nb = nb;
    a.times = times; a.nt = nt; a.rv_out = rv_out;
    a.part_chi2 = part.data(); a.part_status = pst.data();
    a.item_counter = &ctr; a.work_counters = work;
    std::vector<int> order((size_t)W);
    for (long long w = 0; w < W; w++) order[(size_t)w] = (int)(W - 1 - w);
    if (g_reverse && !times) a.order = order.data();
    const int key = P * 10 + m.D;
    switch (key) {
        case 12: run<1, 2>(a); break;
        case 13: run<1, 3>(a); break;
        case 22: run<2, 2>(a); break;
        case 23: run<2, 3>(a); break;
        case 32: run<3, 2>(a); break;
        case 33: run<3, 3>(a); break;
        case 42: run<4, 2>(a); break;
        case 43: run<4, 3>(a); break;
        case 52: run<5, 2>(a); break;
        case 53: run<5, 3>(a); break;
        default: return -9;
    }
    if (times) {
        for (long long w = 0; w < W; w++) status[w] = pst[w];
    } else {
        for (long long w = 0; w < W; w++) {
            const int sb = pst[w], sf = pst[W + w];
            const int s = sf != 0 ? sf : sb;
            status[w] = s;
            logp[w] = (s == 0) ? -((part[w] + part[W + w]) / npoints) : -INFINITY;
        }
    }
    if (counters) { counters[0] = work[0]; counters[1] = work[1]; }
    return 0;
}
