/*
 * rv_samplers.c -- CPU ORACLE of the propose/accept rules (test infrastructure, NOT product code).
 *
 * Restates, on top of rv_oracle.c's likelihood:
 *   mcmc.py:89-121   Mh.generate_proposal / Mh.step
 *   mcmc.py:57-65    Ensemble.step -> emcee 2.2.1 EnsembleSampler._propose_stretch (emcee is an un-vendored
 *                    dependency, script.sh:9 pins 2.2.1; algorithm: Goodman & Weare 2010 stretch move, a=2,
 *                    two half-ensembles)
 * PARITY: accept/reject SEQUENCES are unpinned in the reference (no seeded chain is stored, emcee draws
 * from its own unseeded RandomState), so the device samplers define a counter-based RNG contract
 * (Philox-4x32-10 keyed by seed / walker id / step / stream) and this file restates that contract
 * independently; the GPU tests require identical decisions from identical streams.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int orc_get_logp(int P, const double *elems, double hill_factor, const double *tf, const double *rvf, const double *ef, int nf,
                 const double *tb, const double *rvb, const double *eb, int nb, double npoints, double *logp, long *counters,
                 int *leg_status);

/* Philox-4x32-10 (Salmon, Moraes, Dror & Shaw 2011) */
void orc_philox(uint64_t seed, uint64_t id, uint32_t step, uint32_t stream, uint32_t out[4]) {
    uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    uint32_t c0 = (uint32_t)id, c1 = (uint32_t)(id >> 32), c2 = step, c3 = stream;
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}
static double u53(uint32_t hi, uint32_t lo) {
    return ((double)(hi >> 5) * 67108864.0 + (double)(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0);
}
#define RNG_ACCEPT 1u
#define RNG_STRETCH_Z 0x10u
#define RNG_STRETCH_J 0x20u
#define RNG_NORMAL 0x100u

typedef struct {
    int P, nvars; const double *fixed; const int *fp, *fe; double hill;
    const double *tf, *rvf, *ef; int nf; const double *tb, *rvb, *eb; int nb; double npoints;
} prob_t;

static double eval_lnprob(const prob_t *q, const double *theta, int *status) {
    double elems[8 * 7], lp;
    memcpy(elems, q->fixed, sizeof(double) * (size_t)(q->P * 7));
    for (int v = 0; v < q->nvars; v++) elems[q->fp[v] * 7 + q->fe[v]] = theta[v];
    int st = orc_get_logp(q->P, elems, q->hill, q->tf, q->rvf, q->ef, q->nf, q->tb, q->rvb, q->eb, q->nb, q->npoints, &lp, NULL, NULL);
    if (status) *status = st;
    return st == 0 ? lp : -INFINITY;
}

/* W independent MH chains.  accepted[nsteps][W] (uint8), chain[nsteps][W][nvars] optional. */
int orc_mh_run(int P, const double *fixed, int nvars, const int *fp, const int *fe, double hill,
               const double *tf, const double *rvf, const double *ef, int nf,
               const double *tb, const double *rvb, const double *eb, int nb, double npoints,
               double *theta, double *logp, const double *scales, double step_size, uint64_t seed, uint64_t first_id,
               uint32_t first_step, int nsteps, long W, double *chain, unsigned char *accepted, int nthreads) {
    prob_t q = {P, nvars, fixed, fp, fe, hill, tf, rvf, ef, nf, tb, rvb, eb, nb, npoints};
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (long w = 0; w < W; w++) {
        double *th = theta + w * nvars;
        double prop[64];
        double lp = eval_lnprob(&q, th, NULL);
        for (int k = 0; k < nsteps; k++) {
            const uint32_t step = first_step + (uint32_t)k;
            for (int j = 0; 2 * j < nvars; j++) {
                uint32_t r[4];
                orc_philox(seed, first_id + (uint64_t)w, step, RNG_NORMAL + (uint32_t)j, r);
                const double u1 = u53(r[0], r[1]), u2 = u53(r[2], r[3]);
                const double rad = sqrt(-2.0 * log(u1));
                const double z0 = rad * cos(2.0 * M_PI * u2), z1 = rad * sin(2.0 * M_PI * u2);
                prop[2 * j] = th[2 * j] + step_size * scales[2 * j] * z0;
                if (2 * j + 1 < nvars) prop[2 * j + 1] = th[2 * j + 1] + step_size * scales[2 * j + 1] * z1;
            }
            int st;
            const double lpp = eval_lnprob(&q, prop, &st);
            int acc = 0;
            if (st == 0) {
                uint32_t r[4];
                orc_philox(seed, first_id + (uint64_t)w, step, RNG_ACCEPT, r);
                acc = exp(lpp - lp) > u53(r[0], r[1]);
            }
            if (acc) { memcpy(th, prop, sizeof(double) * (size_t)nvars); lp = lpp; }
            if (accepted) accepted[(size_t)k * W + w] = (unsigned char)acc;
            if (chain) memcpy(chain + ((size_t)k * W + w) * nvars, th, sizeof(double) * (size_t)nvars);
        }
        logp[w] = lp;
    }
    return 0;
}

/* one half-step of the emcee-2.2.1 stretch move with the counter-based draws: S[nS] (ids id0..) updated in place
 * against the complementary half C[nC]; lnp[nS] in/out; accepted[nS] optional */
static void stretch_half(const prob_t *pq, double *S, long nS, uint64_t id0, const double *C, long nC, double *lnp,
                         double a, uint64_t seed, uint32_t step, uint32_t half, unsigned char *accepted, int nthreads) {
    const int nvars = pq->nvars;
    double *q = (double *)malloc(sizeof(double) * (size_t)(nS * nvars + 1));
    double *qlp = (double *)malloc(sizeof(double) * (size_t)nS);
    double *zz = (double *)malloc(sizeof(double) * (size_t)nS);
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
    for (long i = 0; i < nS; i++) {
        uint32_t r[4], rj[4];
        orc_philox(seed, id0 + (uint64_t)i, step, RNG_STRETCH_Z + half, r);
        orc_philox(seed, id0 + (uint64_t)i, step, RNG_STRETCH_J + half, rj);
        /* fixed sequence of correctly-rounded operations -- the RNG contract includes the proposal arithmetic, so that
         * walker positions are bit-identical functions of the accept/reject history on every implementation */
        const double t = fma(a - 1.0, u53(r[0], r[1]), 1.0);
        const double tt = t * t;
        zz[i] = tt / a;
        const long j = (long)(((uint64_t)rj[0] * (uint64_t)nC) >> 32);
        for (int v = 0; v < nvars; v++) {
            const double c = C[j * nvars + v];
            const double cs = c - S[i * nvars + v];
            q[i * nvars + v] = fma(-zz[i], cs, c);
        }
        qlp[i] = eval_lnprob(pq, q + i * nvars, NULL);
    }
    for (long i = 0; i < nS; i++) {
        uint32_t r[4];
        orc_philox(seed, id0 + (uint64_t)i, step, RNG_STRETCH_Z + half, r);
        const double lnpdiff = (double)(nvars - 1) * log(zz[i]) + qlp[i] - lnp[i];
        const int acc = lnpdiff > log(u53(r[2], r[3]));
        if (acc) { memcpy(S + i * nvars, q + i * nvars, sizeof(double) * (size_t)nvars); lnp[i] = qlp[i]; }
        if (accepted) accepted[i] = (unsigned char)acc;
    }
    free(q); free(qlp); free(zz);
}

int orc_stretch_half(int P, const double *fixed, int nvars, const int *fp, const int *fe, double hill,
                     const double *tf, const double *rvf, const double *ef, int nf,
                     const double *tb, const double *rvb, const double *eb, int nb, double npoints,
                     double *S, long nS, uint64_t id0, const double *C, long nC, double *lnp, double a, uint64_t seed,
                     uint32_t step, uint32_t half, unsigned char *accepted, int nthreads) {
    prob_t pq = {P, nvars, fixed, fp, fe, hill, tf, rvf, ef, nf, tb, rvb, eb, nb, npoints};
    stretch_half(&pq, S, nS, id0, C, nC, lnp, a, seed, step, half, accepted, nthreads < 1 ? 1 : nthreads);
    return 0;
}

/* full ensemble: theta[W][nvars], lnp[W] in/out; first half = walkers [0,W/2), second half = the rest */
int orc_stretch_run(int P, const double *fixed, int nvars, const int *fp, const int *fe, double hill,
                    const double *tf, const double *rvf, const double *ef, int nf,
                    const double *tb, const double *rvb, const double *eb, int nb, double npoints,
                    double *theta, double *lnp, int have_lnp, double a, uint64_t seed, uint32_t first_step, int nsteps, long W,
                    double *chain, unsigned char *accepted, int nthreads) {
    prob_t pq = {P, nvars, fixed, fp, fe, hill, tf, rvf, ef, nf, tb, rvb, eb, nb, npoints};
    if (nthreads < 1) nthreads = 1;
    const long h = W / 2;
    if (!have_lnp) {
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
        for (long w = 0; w < W; w++) lnp[w] = eval_lnprob(&pq, theta + w * nvars, NULL);
    }
    for (int k = 0; k < nsteps; k++) {
        const uint32_t step = first_step + (uint32_t)k;
        for (uint32_t half = 0; half < 2; half++) {
            double *S = theta + (half == 0 ? 0 : h * nvars);
            const double *C = theta + (half == 0 ? h * nvars : 0);
            const long id0 = half == 0 ? 0 : h;
            stretch_half(&pq, S, h, (uint64_t)id0, C, h, lnp + id0, a, seed, step, half,
                         accepted ? accepted + (size_t)k * W + id0 : NULL, nthreads);
        }
        if (chain) memcpy(chain + (size_t)k * W * nvars, theta, sizeof(double) * (size_t)(W * nvars));
    }
    return 0;
}
