// rv_whfast.cuh -- optional fixed-step variant of the RV log-likelihood: Wisdom-Holman in Jacobi coordinates as in
// rebound's WHFast (Rein & Tamayo 2015; safe_mode = 1, no correctors), one thread per (walker, leg).
//
// The reference never selects WHFast (parity unpinned); the BASELINE north_star asks for it as an option.
// Same set-up (state.py:36-47), encounter test, chi^2 and prior as the IAS15 path; per step
//   Kepler drift dt/2 (universal variables, Stumpff functions, Newton)  ->  interaction kick dt  ->  Kepler drift dt/2,
// the last step before every epoch shortened to land on it (exact_finish_time = 1).  Each leg is swept monotonically
// with dt = +-dt0 (forward: obs.tf ascending; backward: obs.tb descending).
#pragma once
#include "rv_core.cuh"

namespace rv {

struct WhArgs {
    const Model* model;
    const double* theta;
    long long W;
    const double *ot, *orv, *oerr;
    int nf, nb;
    const double* times;   // RV-curve mode: star vx at times[nt], visited in the given order
    int nt;
    double* rv_out;
    double* part_chi2;     // [2W]: item w = backward leg, item W+w = forward leg (as the IAS15 path)
    int* part_status;
    unsigned long long* work_counters;   // [1] += steps
};

// Stumpff functions c0..c3 (Danby): series for |z| <= 0.1 after quartering, doubling formulae back
RV_HD void stumpff(double z, double (&c)[4]) {
    int n = 0;
    while (fabs(z) > 0.1) { z *= 0.25; n++; }
    c[3] = (1. - z * (1. / 20.) * (1. - z * (1. / 42.) * (1. - z * (1. / 72.) * (1. - z * (1. / 110.) * (1. - z * (1. / 156.) * (1. - z * (1. / 210.))))))) * (1. / 6.);
    c[2] = (1. - z * (1. / 12.) * (1. - z * (1. / 30.) * (1. - z * (1. / 56.) * (1. - z * (1. / 90.) * (1. - z * (1. / 132.) * (1. - z * (1. / 182.))))))) * 0.5;
    c[1] = 1. - z * c[3];
    c[0] = 1. - z * c[2];
    for (; n > 0; n--) {
        c[3] = (c[2] + c[0] * c[3]) * 0.25;
        c[2] = c[1] * c[1] * 0.5;
        c[1] = c[0] * c[1];
        c[0] = 2. * c[0] * c[0] - 1.;
    }
}

// two-body advance of (x, v) about a centre of gravitational parameter M by dt; false on failure
template <int D>
RV_HD bool kepler_step(double M, double dt, double (&x)[D], double (&v)[D]) {
    double r02 = 0.0, v2 = 0.0, eta0 = 0.0;
#pragma unroll
    for (int d = 0; d < D; d++) { r02 = fma(x[d], x[d], r02); v2 = fma(v[d], v[d], v2); eta0 = fma(x[d], v[d], eta0); }
    const double r0 = sqrt(r02);
    const double beta = 2. * M / r0 - v2;
    const double zeta0 = M - beta * r0;
    double X = dt / r0 * (1. - dt * eta0 * 0.5 / r02);
    double c[4], G1 = 0.0, G2 = 0.0, G3 = 0.0, r = r0;
    bool conv = false;
    for (int it = 0; it < 30; it++) {
        stumpff(beta * X * X, c);
        G1 = X * c[1]; G2 = X * X * c[2]; G3 = X * X * X * c[3];
        r = r0 + eta0 * G1 + zeta0 * G2;
        const double F = r0 * X + eta0 * G2 + zeta0 * G3 - dt;
        const double dX = -F / r;
        X += dX;
        if (fabs(dX) <= 1e-15 * fabs(X) || dX == 0.0) { conv = true; break; }
    }
    if (!conv) {    // F is monotone in X (dF/dX = r > 0): bisect
        double lo = 0.0, hi = dt / r0;
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
    if (!isfinite(r) || r == 0.0) return false;
    const double f = -M * G2 / r0, g = dt - M * G3, fd = -M * G1 / (r0 * r), gd = -M * G2 / r;   // f-1, g, f', g'-1
#pragma unroll
    for (int d = 0; d < D; d++) {
        const double nx = x[d] + f * x[d] + g * v[d];
        const double nv = v[d] + fd * x[d] + gd * v[d];
        x[d] = nx; v[d] = nv;
    }
    return true;
}

// One system: star (index 0) + P planets, inertial barycentric state.
template <int P, int D>
struct WhSystem {
    static constexpr int N = P + 1;
    double m[N], x[N][D], v[N][D];
    double min2;
    unsigned long long n_steps;

    RV_HD bool encounter() const {
        if (min2 == 0.0) return false;
        bool hit = false;
#pragma unroll
        for (int i = 0; i < N; i++)
#pragma unroll
            for (int k = 0; k < i; k++) {
                double r2 = 0.0;
#pragma unroll
                for (int d = 0; d < D; d++) { const double dx = x[i][d] - x[k][d]; r2 = fma(dx, dx, r2); }
                hit = hit || (r2 < min2);
            }
        return hit;
    }

    // drift(dt/2) kick(dt) drift(dt/2), synchronised
    RV_HD bool step(double dt) {
        double jx[N][D], jv[N][D];
        // inertial -> Jacobi
        {
            double eta = m[0], sx[D], sv[D];
#pragma unroll
            for (int d = 0; d < D; d++) { sx[d] = m[0] * x[0][d]; sv[d] = m[0] * v[0][d]; }
#pragma unroll
            for (int i = 1; i < N; i++) {
#pragma unroll
                for (int d = 0; d < D; d++) { jx[i][d] = x[i][d] - sx[d] / eta; jv[i][d] = v[i][d] - sv[d] / eta; }
#pragma unroll
                for (int d = 0; d < D; d++) { sx[d] += m[i] * x[i][d]; sv[d] += m[i] * v[i][d]; }
                eta += m[i];
            }
#pragma unroll
            for (int d = 0; d < D; d++) { jx[0][d] = sx[d] / eta; jv[0][d] = sv[d] / eta; }
        }
        bool ok = true;
        auto drift = [&](double h) {
            double eta = m[0];
#pragma unroll
            for (int i = 1; i < N; i++) { eta += m[i]; ok = kepler_step<D>(eta, h, jx[i], jv[i]) && ok; }
#pragma unroll
            for (int d = 0; d < D; d++) jx[0][d] += h * jv[0][d];
        };
        auto to_inertial = [&]() {
            double eta = 0.0;
#pragma unroll
            for (int i = 0; i < N; i++) eta += m[i];
            double sx[D], sv[D];
#pragma unroll
            for (int d = 0; d < D; d++) { sx[d] = jx[0][d] * eta; sv[d] = jv[0][d] * eta; }
#pragma unroll
            for (int i = N - 1; i >= 1; i--) {
#pragma unroll
                for (int d = 0; d < D; d++) {
                    const double Rx = (sx[d] - m[i] * jx[i][d]) / eta, Rv = (sv[d] - m[i] * jv[i][d]) / eta;
                    x[i][d] = jx[i][d] + Rx; v[i][d] = jv[i][d] + Rv;
                    sx[d] -= m[i] * x[i][d]; sv[d] -= m[i] * v[i][d];
                }
                eta -= m[i];
            }
#pragma unroll
            for (int d = 0; d < D; d++) { x[0][d] = sx[d] / m[0]; v[0][d] = sv[d] / m[0]; }
        };
        drift(0.5 * dt);
        to_inertial();
        // interaction kick: direct gravity without the star--planet-1 pair, Jacobi accelerations, + G eta r'/r'^3 (i > 1)
        {
            double a[N][D];
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int d = 0; d < D; d++) a[i][d] = 0.0;
#pragma unroll
            for (int i = 0; i < N; i++)
#pragma unroll
                for (int k = i + 1; k < N; k++) {
                    if (i == 0 && k == 1) continue;
                    double dx[D], r2 = 0.0;
#pragma unroll
                    for (int d = 0; d < D; d++) { dx[d] = x[i][d] - x[k][d]; r2 = fma(dx[d], dx[d], r2); }
                    const double r3i = 1. / (r2 * sqrt(r2));
#pragma unroll
                    for (int d = 0; d < D; d++) { a[i][d] -= m[k] * r3i * dx[d]; a[k][d] += m[i] * r3i * dx[d]; }
                }
            double eta = m[0], sa[D];
#pragma unroll
            for (int d = 0; d < D; d++) sa[d] = m[0] * a[0][d];
#pragma unroll
            for (int i = 1; i < N; i++) {
                double aj[D];
#pragma unroll
                for (int d = 0; d < D; d++) aj[d] = a[i][d] - sa[d] / eta;
#pragma unroll
                for (int d = 0; d < D; d++) sa[d] += m[i] * a[i][d];
                eta += m[i];
                if (i > 1) {
                    double r2 = 0.0;
#pragma unroll
                    for (int d = 0; d < D; d++) r2 = fma(jx[i][d], jx[i][d], r2);
                    const double k3 = eta / (r2 * sqrt(r2));
#pragma unroll
                    for (int d = 0; d < D; d++) aj[d] += k3 * jx[i][d];
                }
#pragma unroll
                for (int d = 0; d < D; d++) jv[i][d] += dt * aj[d];
            }
        }
        drift(0.5 * dt);
        to_inertial();
        n_steps++;
        return ok;
    }

    // sim.integrate(tmax) with steps of dt (sign given), the last one shortened; ST_OK / ST_ENCOUNTER / ST_NONFINITE
    RV_HD int integrate(double& t, double dt, double tmax, int max_steps) {
        if (encounter()) return ST_ENCOUNTER;
        const double sgn = dt >= 0.0 ? 1.0 : -1.0;
        while (t != tmax) {
            const bool last = (t + dt) * sgn >= tmax * sgn;
            const double h = last ? tmax - t : dt;
            if (!step(h)) return ST_NONFINITE;
            t = last ? tmax : t + h;
            if (encounter()) return ST_ENCOUNTER;
            if (n_steps > (unsigned long long)max_steps) return ST_NONFINITE;
        }
        return ST_OK;
    }

    // setup_sim (state.py:36-47); ST_OK or ST_PRIOR
    RV_HD int setup(const Model* __restrict__ md, const double* __restrict__ theta, bool check_prior) {
        double el[P][NELEM];
        bool bad = false;
        const double m0 = md->m_star;
#pragma unroll
        for (int i = 0; i < P; i++) {
#pragma unroll
            for (int k = 0; k < NELEM; k++) {
                const int s = md->src[i * NELEM + k];
                el[i][k] = (s >= 0) ? theta[s] : md->fixed[i * NELEM + k];
            }
            bad = bad || prior_hard(el[i]);
        }
        if (check_prior && bad) return ST_PRIOR;
        double xr[P][3], vr[P][3], hill = 0.0, mtot = m0, cx[3] = {0, 0, 0}, cv[3] = {0, 0, 0};
        m[0] = m0;
#pragma unroll
        for (int i = 0; i < P; i++) {
            pal_to_cart(el[i], m0, xr[i], vr[i]);
            m[i + 1] = el[i][EL_M];
            mtot += el[i][EL_M];
            const double rh = el[i][EL_A] * pow(el[i][EL_M] / (3.0 * m0), 1.0 / 3.0);
            if (rh > hill) hill = rh;
#pragma unroll
            for (int d = 0; d < 3; d++) { cx[d] += el[i][EL_M] * xr[i][d]; cv[d] += el[i][EL_M] * vr[i][d]; }
        }
#pragma unroll
        for (int d = 0; d < D; d++) {
            x[0][d] = -cx[d] / mtot; v[0][d] = -cv[d] / mtot;
#pragma unroll
            for (int i = 0; i < P; i++) { x[i + 1][d] = xr[i][d] - cx[d] / mtot; v[i + 1][d] = vr[i][d] - cv[d] / mtot; }
        }
        const double emd = md->hill_factor * hill;
        min2 = emd * emd;
        n_steps = 0;
        return ST_OK;
    }
};

// one work item: a leg of a walker's likelihood, or (curve mode) a walker's RV curve
template <int P, int D>
RV_HD void whfast_item(const WhArgs& a, long long item) {
    const Model* __restrict__ md = a.model;
    const bool curve = a.times != nullptr;
    const bool backward = !curve && item < a.W;
    const long long wi = curve ? item : (backward ? item : item - a.W);
    WhSystem<P, D> s;
    int st = s.setup(md, a.theta + wi * md->nvars, !curve);
    double chi2 = 0.0, t = 0.0;
    if (st == ST_OK) {
        const int n = curve ? a.nt : (backward ? a.nb : a.nf);
        const int base = backward ? a.nf : 0;
        for (int ii = 0; ii < n && st == ST_OK; ii++) {
            const int ie = backward ? base + n - 1 - ii : base + ii;
            const double tmax = curve ? a.times[ii] : a.ot[ie];
            const double dt = tmax >= t ? md->dt0 : -md->dt0;
            st = s.integrate(t, dt, tmax, md->max_attempts);
            if (st != ST_OK) break;
            const double vx = s.v[0][0];
            if (!isfinite(vx)) { st = ST_NONFINITE; break; }
            if (curve) {
                a.rv_out[wi * a.nt + ii] = vx;
            } else {
                const double r = vx - a.orv[ie], er = a.oerr[ie];
                chi2 += (r * r) / (er * er);
            }
        }
    }
    a.part_status[item] = st;
    if (!curve) a.part_chi2[item] = chi2;
#if defined(__CUDA_ARCH__)
    if (a.work_counters) atomicAdd(&a.work_counters[1], s.n_steps);
#else
    if (a.work_counters) a.work_counters[1] += s.n_steps;
#endif
}

}  // namespace rv
