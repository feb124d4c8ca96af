// TEST-ONLY sequential emulation of the variational CTA algorithm (rv_var.cuh): the executor runs every
// "thread" of the block in turn between barriers, so the CPU test-suite exercises the exact device source.
// NOT part of the product library and never loaded by it.
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "../../rvel_mcmc_b200/csrc/rv_var.cuh"
#include "../../rvel_mcmc_b200/csrc/rv_var2.cuh"
#include "../../rvel_mcmc_b200/csrc/rv_model.h"

namespace {
template <int P, int D>
struct HostVarExec {
    std::vector<rv::VarThread<P, D>> th;
    double cur[2] = {0.0, 0.0};
    template <class F> void each(F&& f) { for (auto& t : th) f(t); }
    void sync() {}
    void stage_max(const rv::VarThread<P, D>&, double a, double b) { if (a > cur[0]) cur[0] = a; if (b > cur[1]) cur[1] = b; }
    void read_max(double& a, double& b) { a = cur[0]; b = cur[1]; cur[0] = cur[1] = 0.0; }
    long long fetch(unsigned long long* ctr) { return (long long)((*ctr)++); }
    void add_work(unsigned long long* wc, unsigned long long nf, unsigned long long na) { if (wc) { wc[0] += nf; wc[1] += na; } }
};

int g_var_nt_cap = 0;       // > 0: threads of one emulated CTA; models that need more run in chunks of second-order pairs

template <int P, int D>
int run_launch(const rv::VarArgs& a, const rv::VarLayout& L, int NT) {
    std::vector<double> sm((size_t)L.total, 0.0);
    HostVarExec<P, D> ex;
    ex.th.resize(NT);
    for (int t = 0; t < NT; t++) rv::var_assign(ex.th[t], t, L);
    rv::var_run_items<P, D>(ex, a, L, sm.data());
    return 0;
}

template <int P, int D>
int run(const rv::VarArgs& a, int nv) {
    int NT = 64;
    rv::VarLayout L = rv::var_layout(P, D, nv, NT);
    while (L.need > NT) { NT += 32; L = rv::var_layout(P, D, nv, NT); }
    if (g_var_nt_cap <= 0 || L.need <= g_var_nt_cap) return run_launch<P, D>(a, L, NT);
    // the launcher's chunked schedule (launch_var_chunked in rv_var_kernels.cu)
    NT = g_var_nt_cap;
    const int n2 = nv * (nv + 1) / 2, chunk = rv::var_chunk_pairs(P, nv, NT);
    if (chunk < 1) return -30;
    for (int lo = 0; lo < n2; lo += chunk) {
        *a.item_counter = 0;
        const int n = (n2 - lo < chunk) ? n2 - lo : chunk;
        run_launch<P, D>(a, rv::var_layout(P, D, nv, NT, lo, n), NT);
    }
    return 0;
}
// the set-per-lane layout (rv_var2.cuh): barriers and named-barrier signals are no-ops in a sequential run because
// every phase is completed for all threads before the next one starts
template <int P, int D>
struct HostVar2Exec {
    std::vector<rv::Var2Thread<P, D>> th;
    double cur[2] = {0.0, 0.0};
    template <class F> void each(F&& f) { for (auto& t : th) f(t); }
    void sync() {}
    void producer_sync() {}
    void signal(int) {}
    void wait(int) {}
    void stage_max(const rv::Var2Thread<P, D>&, double a, double b) { if (a > cur[0]) cur[0] = a; if (b > cur[1]) cur[1] = b; }
    void read_max(double& a, double& b) { a = cur[0]; b = cur[1]; cur[0] = cur[1] = 0.0; }
    long long fetch(unsigned long long* ctr) { return (long long)((*ctr)++); }
    void add_work(unsigned long long* wc, unsigned long long nf, unsigned long long na) { if (wc) { wc[0] += nf; wc[1] += na; } }
};

template <int P, int D>
int run2(const rv::VarArgs& a, int nv) {
    const rv::Var2Layout L = rv::var2_layout(P, D, nv);
    std::vector<double> sm((size_t)L.total, 0.0);
    HostVar2Exec<P, D> ex;
    ex.th.resize(L.NT);
    for (int t = 0; t < L.NT; t++) rv::var2_assign(ex.th[t], t, L);
    std::vector<double> hist(rv::var2_hist_doubles(P, D, nv), 0.0);
    rv::var2_run_items<P, D>(ex, a, L, sm.data(), hist.data());
    return 0;
}
}  // namespace

static int g_var_layout = 0;      // 0: thread per (set, planet) (rv_var.cuh); 2: lane per set (rv_var2.cuh)
extern "C" void mirror_set_var_layout(int v) { g_var_layout = v; }
extern "C" void mirror_set_var_nt_cap(int v) { g_var_nt_cap = v; }

extern "C" int mirror_loglik_d_dd(int P, const double* fixed, int nvars, const int* fp, const int* fe, double hill, int dims,
                                  const double* tf, const double* rvf, const double* ef, int nf,
                                  const double* tb, const double* rvb, const double* eb, int nb, double npoints,
                                  const double* theta, long long W, double* logp, double* grad, double* hess, int* status,
                                  unsigned long long* counters) {
    rv::Model m;
    int rc = rv::build_model(&m, P, fixed, nvars, fp, fe, hill, dims);
    if (rc) return rc;
    std::vector<double> ot(nf + nb), orv(nf + nb), oerr(nf + nb);
    for (int i = 0; i < nf; i++) { ot[i] = tf[i]; orv[i] = rvf[i]; oerr[i] = ef[i]; }
    for (int i = 0; i < nb; i++) { ot[nf + i] = tb[i]; orv[nf + i] = rvb[i]; oerr[nf + i] = eb[i]; }
    const int nsets = rv::var_nsets(nvars);
    std::vector<double> part((size_t)2 * W * nsets, 0.0);
    std::vector<int> pst(2 * W, -1);
    unsigned long long ctr = 0, work[2] = {0, 0};
    rv::VarArgs a;
    memset(&a, 0, sizeof a);
    a.model = &m; a.theta = theta; a.W = W;
    a.ot = ot.data(); a.orv = orv.data(); a.oerr = oerr.data(); a.nf = nf; a.nb = nb; a.npoints = npoints; a.check_prior = m.check_prior;
    a.part = part.data(); a.part_status = pst.data(); a.item_counter = &ctr; a.work_counters = work;
    const int key = P * 10 + m.D + (g_var_layout == 2 ? 100 : 0);
    switch (key) {
        case 112: run2<1, 2>(a, nvars); break;
        case 113: run2<1, 3>(a, nvars); break;
        case 122: run2<2, 2>(a, nvars); break;
        case 123: run2<2, 3>(a, nvars); break;
        case 12: run<1, 2>(a, nvars); break;
        case 13: run<1, 3>(a, nvars); break;
        case 22: run<2, 2>(a, nvars); break;
        case 23: run<2, 3>(a, nvars); break;
        case 32: run<3, 2>(a, nvars); break;
        case 33: run<3, 3>(a, nvars); break;
        case 42: run<4, 2>(a, nvars); break;
        case 43: run<4, 3>(a, nvars); break;
        case 52: run<5, 2>(a, nvars); break;
        case 53: run<5, 3>(a, nvars); break;
        default: return -9;
    }
    for (long long w = 0; w < W; w++) {
        const int sb = pst[w], sf = pst[W + w];
        const int s = sf != 0 ? sf : sb;
        status[w] = s;
        if (s != 0) { logp[w] = -INFINITY; continue; }
        const double* pb = &part[(size_t)w * nsets];
        const double* pf = &part[(size_t)(W + w) * nsets];
        logp[w] = -(pb[0] + pf[0]);
        for (int v = 0; v < nvars; v++) grad[w * nvars + v] = -(pb[1 + v] + pf[1 + v]);
        int k = 0;
        for (int x = 0; x < nvars; x++)
            for (int y = 0; y <= x; y++, k++) {
                const double v = -(pb[1 + nvars + k] + pf[1 + nvars + k]);
                hess[((size_t)w * nvars + x) * nvars + y] = v;
                hess[((size_t)w * nvars + y) * nvars + x] = v;
            }
    }
    if (counters) { counters[0] = work[0]; counters[1] = work[1]; }
    return 0;
}

// ---- per-chain SMALA arithmetic (rv_smala.cuh) ------------------------------------------------------------
#include "../../rvel_mcmc_b200/csrc/rv_smala.cuh"

extern "C" int mirror_smala_propose(int n, const double* th, const double* g, const double* H, int cur_status, double eps,
                                    double alpha, unsigned long long seed, unsigned long long id, unsigned step,
                                    double* prop, double* q_fwd) {
    std::vector<double> scr((size_t)5 * n * n, 0.0);
    double q = 0.0;
    const int r = rv::smala_propose_one(n, th, g, H, cur_status, eps, alpha, seed, id, step, prop, q, scr.data());
    *q_fwd = q;
    return r;
}

extern "C" int mirror_smala_accept(int n, const double* th, double logp, const double* prop, double p_logp,
                                   const double* p_grad, const double* p_hess, int p_status, int geo_status, double q_fwd,
                                   double eps, double alpha, unsigned long long seed, unsigned long long id, unsigned step,
                                   int* flag) {
    std::vector<double> scr((size_t)5 * n * n, 0.0);
    return rv::smala_accept_one(n, th, logp, prop, p_logp, p_grad, p_hess, p_status, geo_status, q_fwd, eps, alpha, seed, id,
                                step, flag, scr.data());
}
