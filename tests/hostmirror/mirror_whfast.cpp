// TEST-ONLY host build of the optional WHFast engine (rv_whfast.cuh).  NOT part of the product library.
#include <math.h>
#include <string.h>
#include <vector>
#include "../../rvel_mcmc_b200/csrc/rv_whfast.cuh"
#include "../../rvel_mcmc_b200/csrc/rv_model.h"

extern "C" int mirror_whfast(int P, const double* fixed, int nvars, const int* fp, const int* fe, double hill, int dims,
                             double dt0, const double* tf, const double* rvf, const double* ef, int nf,
                             const double* tb, const double* rvb, const double* eb, int nb, double npoints,
                             const double* theta, long long W, double* logp, int* status,
                             const double* times, int nt, double* rv_out, unsigned long long* counters) {
    rv::Model m;
    int rc = rv::build_model(&m, P, fixed, nvars, fp, fe, hill, dims);
    if (rc) return rc;
    m.integrator = 1; m.dt0 = dt0;
    std::vector<double> ot(nf + nb), orv(nf + nb), oerr(nf + nb);
    for (int i = 0; i < nf; i++) { ot[i] = tf[i]; orv[i] = rvf[i]; oerr[i] = ef[i]; }
    for (int i = 0; i < nb; i++) { ot[nf + i] = tb[i]; orv[nf + i] = rvb[i]; oerr[nf + i] = eb[i]; }
    std::vector<double> part(2 * W, 0.0);
    std::vector<int> pst(2 * W, -1);
    unsigned long long work[2] = {0, 0};
    rv::WhArgs a;
    memset(&a, 0, sizeof a);
    a.model = &m; a.theta = theta; a.W = W;
    a.ot = ot.data(); a.orv = orv.data(); a.oerr = oerr.data(); a.nf = nf; a.nb = nb;
    a.times = times; a.nt = nt; a.rv_out = rv_out;
    a.part_chi2 = part.data(); a.part_status = pst.data(); a.work_counters = work;
    const long long n_items = times ? W : 2 * W;
    for (long long it = 0; it < n_items; it++) {
        switch (P * 10 + m.D) {
            case 12: rv::whfast_item<1, 2>(a, it); break;
            case 13: rv::whfast_item<1, 3>(a, it); break;
            case 22: rv::whfast_item<2, 2>(a, it); break;
            case 23: rv::whfast_item<2, 3>(a, it); break;
            case 32: rv::whfast_item<3, 2>(a, it); break;
            case 33: rv::whfast_item<3, 3>(a, it); break;
            case 42: rv::whfast_item<4, 2>(a, it); break;
            case 43: rv::whfast_item<4, 3>(a, it); break;
            case 52: rv::whfast_item<5, 2>(a, it); break;
            case 53: rv::whfast_item<5, 3>(a, it); break;
            default: return -9;
        }
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
