/*
 * rv_whfast.c -- CPU ORACLE of the optional WHFast variant (test infrastructure, NOT product code).
 *
 * PARITY UNPINNED: the reference never selects WHFast (no `sim.integrator =` anywhere), so no
 * reference output exists for this path.  It is part of the north_star ("plus an optional WHFast fixed-step
 * variant"), so this file restates the published algorithm -- Wisdom & Holman 1991 in Jacobi coordinates as
 * implemented by rebound's WHFast (Rein & Tamayo 2015): per step Kepler drift dt/2 of every Jacobi body by the
 * universal-variable solver (Stumpff functions, Newton), interaction kick dt (direct gravity without the
 * star--planet-1 pair, transformed to Jacobi accelerations, plus G eta r'/r'^3 for the outer bodies), Kepler
 * drift dt/2, synchronised every step (safe_mode = 1), no symplectic correctors -- on top of the same set-up
 * (state.py:36-47), encounter test and chi^2 (state.py:89-110) as the IAS15 path.  The last step before every
 * epoch is shortened to land on it (rebound's exact_finish_time = 1).  Each leg is swept monotonically with
 * dt = +-dt0 (rebound would take the whole first backward hop as ONE step if dt kept the wrong sign).
 * Its checks are: agreement with the IAS15 oracle at O(dt^2), time reversibility, Kepler-solver identities.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

int orc_initial_conditions(int P, const double *elems, double *out_com, double *out_rel);
int orc_prior_hard(int P, const double *elems);

#define WH_OK 0
#define WH_PRIOR 1
#define WH_ENCOUNTER 3
#define WH_NONFINITE 8
#define NB 8

typedef struct { double m[NB], x[NB][3], v[NB][3]; int N; } sys_t;     /* index 0 = star (inertial) */
typedef struct { double x[NB][3], v[NB][3]; } jac_t;                    /* Jacobi; index 0 = centre of mass */

/* Stumpff functions c0..c3 of z (Danby 1992: series for |z| < 0.1 after quartering, then the doubling formulae) */
static void stumpff(double z, double c[4]) {
    int n = 0;
    while (fabs(z) > 0.1) { z *= 0.25; n++; }
    c[3] = (1. - z / 20. * (1. - z / 42. * (1. - z / 72. * (1. - z / 110. * (1. - z / 156. * (1. - z / 210.)))))) / 6.;
    c[2] = (1. - z / 12. * (1. - z / 30. * (1. - z / 56. * (1. - z / 90. * (1. - z / 132. * (1. - z / 182.)))))) / 2.;
    c[1] = 1. - z * c[3];
    c[0] = 1. - z * c[2];
    for (; n > 0; n--) {
        c[3] = (c[2] + c[0] * c[3]) * 0.25;
        c[2] = c[1] * c[1] * 0.5;
        c[1] = c[0] * c[1];
        c[0] = 2. * c[0] * c[0] - 1.;
    }
}

/* advance a two-body orbit (relative position x, velocity v, gravitational parameter M) by dt. returns 0 or WH_NONFINITE */
int orc_kepler_step(double M, double dt, double *x, double *v) {
    const double r0 = sqrt(x[0] * x[0] + x[1] * x[1] + x[2] * x[2]);
    const double v2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2];
    const double eta0 = x[0] * v[0] + x[1] * v[1] + x[2] * v[2];
    const double beta = 2. * M / r0 - v2;
    const double zeta0 = M - beta * r0;
    double X = dt / r0 * (1. - dt * eta0 * 0.5 / (r0 * r0));
    double c[4], G1 = 0, G2 = 0, G3 = 0, r = r0;
    int conv = 0;
    for (int it = 0; it < 30; it++) {
        stumpff(beta * X * X, c);
        G1 = X * c[1]; G2 = X * X * c[2]; G3 = X * X * X * c[3];
        r = r0 + eta0 * G1 + zeta0 * G2;
        const double F = r0 * X + eta0 * G2 + zeta0 * G3 - dt;
        const double dX = -F / r;
        X += dX;
        if (fabs(dX) <= 1e-15 * fabs(X) || dX == 0.0) { conv = 1; break; }
    }
    if (!conv) {    /* F is monotone in X (F' = r > 0): bisect */
        double lo = 0, hi = dt / r0;
        for (int k = 0; k < 200; k++) {
            stumpff(beta * hi * hi, c);
            const double F = r0 * hi + eta0 * hi * hi * c[2] + zeta0 * hi * hi * hi * c[3] - dt;
            if ((dt > 0 && F > 0) || (dt < 0 && F < 0)) break;
            hi *= 2;
        }
        for (int k = 0; k < 200; k++) {
            X = 0.5 * (lo + hi);
            stumpff(beta * X * X, c);
            const double F = r0 * X + eta0 * X * X * c[2] + zeta0 * X * X * X * c[3] - dt;
            if ((F > 0) == (dt > 0)) hi = X; else lo = X;
        }
    }
    stumpff(beta * X * X, c);
    G1 = X * c[1]; G2 = X * X * c[2]; G3 = X * X * X * c[3];
    r = r0 + eta0 * G1 + zeta0 * G2;
    if (!isfinite(r) || r == 0.0) return WH_NONFINITE;
    const double f = -M * G2 / r0, g = dt - M * G3, fd = -M * G1 / (r0 * r), gd = -M * G2 / r;   /* f-1, g, f', g'-1 */
    for (int d = 0; d < 3; d++) {
        const double nx = x[d] + f * x[d] + g * v[d];
        const double nv = v[d] + fd * x[d] + gd * v[d];
        x[d] = nx; v[d] = nv;
    }
    return WH_OK;
}

static void to_jacobi(const sys_t *s, jac_t *j) {
    double eta = s->m[0], sx[3], sv[3];
    for (int d = 0; d < 3; d++) { sx[d] = s->m[0] * s->x[0][d]; sv[d] = s->m[0] * s->v[0][d]; }
    for (int i = 1; i < s->N; i++) {
        for (int d = 0; d < 3; d++) { j->x[i][d] = s->x[i][d] - sx[d] / eta; j->v[i][d] = s->v[i][d] - sv[d] / eta; }
        for (int d = 0; d < 3; d++) { sx[d] += s->m[i] * s->x[i][d]; sv[d] += s->m[i] * s->v[i][d]; }
        eta += s->m[i];
    }
    for (int d = 0; d < 3; d++) { j->x[0][d] = sx[d] / eta; j->v[0][d] = sv[d] / eta; }
}

static void from_jacobi(sys_t *s, const jac_t *j) {
    double eta = 0;
    for (int i = 0; i < s->N; i++) eta += s->m[i];
    double sx[3], sv[3];
    for (int d = 0; d < 3; d++) { sx[d] = j->x[0][d] * eta; sv[d] = j->v[0][d] * eta; }   /* mass-weighted sum of bodies 0..i */
    for (int i = s->N - 1; i >= 1; i--) {
        /* x_i = x'_i + R_{i-1};  eta_i R_i = eta_{i-1} R_{i-1} + m_i x_i  =>  R_{i-1} = (S_i - m_i x'_i)/eta_i */
        for (int d = 0; d < 3; d++) {
            const double Rx = (sx[d] - s->m[i] * j->x[i][d]) / eta, Rv = (sv[d] - s->m[i] * j->v[i][d]) / eta;
            s->x[i][d] = j->x[i][d] + Rx; s->v[i][d] = j->v[i][d] + Rv;
            sx[d] -= s->m[i] * s->x[i][d]; sv[d] -= s->m[i] * s->v[i][d];
        }
        eta -= s->m[i];
    }
    for (int d = 0; d < 3; d++) { s->x[0][d] = sx[d] / s->m[0]; s->v[0][d] = sv[d] / s->m[0]; }
}

/* interaction kick on the Jacobi velocities */
static void kick(const sys_t *s, jac_t *j, double dt) {
    double a[NB][3];
    memset(a, 0, sizeof a);
    for (int i = 0; i < s->N; i++)
        for (int k = i + 1; k < s->N; k++) {
            if (i == 0 && k == 1) continue;            /* solved exactly by the Kepler drift of Jacobi body 1 */
            double dx[3], r2 = 0;
            for (int d = 0; d < 3; d++) { dx[d] = s->x[i][d] - s->x[k][d]; r2 += dx[d] * dx[d]; }
            const double r3i = 1. / (r2 * sqrt(r2));
            for (int d = 0; d < 3; d++) { a[i][d] -= s->m[k] * r3i * dx[d]; a[k][d] += s->m[i] * r3i * dx[d]; }
        }
    double eta = s->m[0], sa[3];
    for (int d = 0; d < 3; d++) sa[d] = s->m[0] * a[0][d];
    for (int i = 1; i < s->N; i++) {
        double aj[3];
        for (int d = 0; d < 3; d++) aj[d] = a[i][d] - sa[d] / eta;
        for (int d = 0; d < 3; d++) sa[d] += s->m[i] * a[i][d];
        eta += s->m[i];
        if (i > 1) {
            double r2 = 0;
            for (int d = 0; d < 3; d++) r2 += j->x[i][d] * j->x[i][d];
            const double k3 = eta / (r2 * sqrt(r2));
            for (int d = 0; d < 3; d++) aj[d] += k3 * j->x[i][d];
        }
        for (int d = 0; d < 3; d++) j->v[i][d] += dt * aj[d];
    }
}

static int wh_step(sys_t *s, double dt) {
    jac_t j;
    to_jacobi(s, &j);
    double eta = s->m[0];
    for (int i = 1; i < s->N; i++) { eta += s->m[i]; if (orc_kepler_step(eta, 0.5 * dt, j.x[i], j.v[i])) return WH_NONFINITE; }
    for (int d = 0; d < 3; d++) j.x[0][d] += 0.5 * dt * j.v[0][d];
    from_jacobi(s, &j);
    kick(s, &j, dt);
    eta = s->m[0];
    for (int i = 1; i < s->N; i++) { eta += s->m[i]; if (orc_kepler_step(eta, 0.5 * dt, j.x[i], j.v[i])) return WH_NONFINITE; }
    for (int d = 0; d < 3; d++) j.x[0][d] += 0.5 * dt * j.v[0][d];
    from_jacobi(s, &j);
    return WH_OK;
}

static int encounter(const sys_t *s, double min2) {
    if (min2 == 0.0) return 0;
    for (int i = 0; i < s->N; i++)
        for (int k = 0; k < i; k++) {
            double r2 = 0;
            for (int d = 0; d < 3; d++) { const double dx = s->x[i][d] - s->x[k][d]; r2 += dx * dx; }
            if (r2 < min2) return 1;
        }
    return 0;
}

static void setup(sys_t *s, int P, const double *elems, double hill_factor, double *min2) {
    double ic[NB * 7];
    orc_initial_conditions(P, elems, ic, NULL);
    s->N = P + 1;
    double hmax = 0;
    for (int i = 0; i <= P; i++) {
        s->m[i] = ic[i * 7];
        for (int d = 0; d < 3; d++) { s->x[i][d] = ic[i * 7 + 1 + d]; s->v[i][d] = ic[i * 7 + 4 + d]; }
    }
    for (int i = 0; i < P; i++) {
        const double rh = elems[i * 7 + 1] * pow(elems[i * 7 + 0] / 3.0, 1.0 / 3.0);
        if (rh > hmax) hmax = rh;
    }
    const double emd = hill_factor * hmax;
    *min2 = emd * emd;
}

/* integrate to tmax with steps of dt (sign given), the last one shortened (exact_finish_time = 1) */
static int integrate(sys_t *s, double *t, double dt, double tmax, double min2, long *nsteps) {
    if (encounter(s, min2)) return WH_ENCOUNTER;
    const double sgn = dt >= 0 ? 1.0 : -1.0;
    while (*t != tmax) {
        double h = dt;
        if ((*t + dt) * sgn >= tmax * sgn) h = tmax - *t;
        if (wh_step(s, h)) return WH_NONFINITE;
        if ((*t + dt) * sgn >= tmax * sgn) *t = tmax; else *t += h;
        if (nsteps) (*nsteps)++;
        if (encounter(s, min2)) return WH_ENCOUNTER;
        if (*nsteps > 100000000L) return WH_NONFINITE;
    }
    return WH_OK;
}

/* star vx at times[nt], visited in the given order; dt0 > 0; the step sign follows the direction of every hop */
int orc_whfast_get_rv(int P, const double *elems, double hill_factor, double dt0, const double *times, int nt, double *rv,
                      long *counters) {
    sys_t s;
    double min2, t = 0;
    long ns = 0;
    setup(&s, P, elems, hill_factor, &min2);
    for (int i = 0; i < nt; i++) {
        const double dt = times[i] >= t ? dt0 : -dt0;
        const int r = integrate(&s, &t, dt, times[i], min2, &ns);
        if (r) { if (counters) counters[1] += ns; return r; }
        rv[i] = s.v[0][0];
        if (!isfinite(rv[i])) return WH_NONFINITE;
    }
    if (counters) counters[1] += ns;
    return WH_OK;
}

/* logp with both legs swept monotonically: forward = tf ascending, backward = tb descending (tb stored ascending) */
int orc_whfast_get_logp(int P, const double *elems, double hill_factor, double dt0,
                        const double *tf, const double *rvf, const double *ef, int nf,
                        const double *tb, const double *rvb, const double *eb, int nb, double npoints, double *logp,
                        long *counters) {
    *logp = -INFINITY;
    if (orc_prior_hard(P, elems)) return WH_PRIOR;
    double *buf = (double *)malloc(sizeof(double) * (size_t)(nf + 2 * nb + 2));
    double *rf = buf, *rb = buf + nf, *tr = rb + nb;
    int st = orc_whfast_get_rv(P, elems, hill_factor, dt0, tf, nf, rf, counters);
    if (st == 0) {
        for (int i = 0; i < nb; i++) tr[i] = tb[nb - 1 - i];
        st = orc_whfast_get_rv(P, elems, hill_factor, dt0, tr, nb, rb, counters);
    }
    if (st == 0) {
        double cf = 0, cb = 0;
        for (int i = 0; i < nf; i++) cf += ((rf[i] - rvf[i]) * (rf[i] - rvf[i])) / (ef[i] * ef[i]);
        for (int i = 0; i < nb; i++) { const int k = nb - 1 - i; cb += ((rb[i] - rvb[k]) * (rb[i] - rvb[k])) / (eb[k] * eb[k]); }
        *logp = -((cb + cf) / npoints);
    }
    free(buf);
    return st;
}

int orc_whfast_logp_batch(int P, const double *fixed, int nvars, const int *fp, const int *fe, double hill_factor, double dt0,
                          const double *tf, const double *rvf, const double *ef, int nf,
                          const double *tb, const double *rvb, const double *eb, int nb, double npoints,
                          const double *theta, long W, double *logp, int *status, long *counters, int nthreads) {
    long c1 = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads) reduction(+ : c1)
    for (long w = 0; w < W; w++) {
        double elems[NB * 7];
        long c[3] = {0, 0, 0};
        memcpy(elems, fixed, sizeof(double) * (size_t)(P * 7));
        for (int v = 0; v < nvars; v++) elems[fp[v] * 7 + fe[v]] = theta[w * nvars + v];
        status[w] = orc_whfast_get_logp(P, elems, hill_factor, dt0, tf, rvf, ef, nf, tb, rvb, eb, nb, npoints, &logp[w], c);
        c1 += c[1];
    }
    if (counters) counters[1] += c1;
    return 0;
}
