/*
 * rv_oracle.c -- CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the arithmetic that rvel-mcmc's hot path executes
 * on the CPU.  The reference (Python 2, /root/reference) delegates every
 * floating-point operation of that path to the third-party `rebound` C library
 * (un-vendored, un-pinned; API era v2.17-v3.3, 2016-17: state.py:235-245 uses
 * add_variation(order=2, first_order=, first_order_2=)).  `rebound` is not
 * installable here, so this file restates its published algorithms
 * (Rein & Spiegel 2015 IAS15; Rein & Tamayo 2016 variational equations;
 * Pal 2009 elements) and anchors parity on the reference's own call sites:
 *
 *   state.py:36-47    setup_sim        -> orc_setup()
 *   state.py:61-73    get_rv           -> orc_get_rv()
 *   state.py:89-98    get_chi2         -> orc_get_logp()
 *   state.py:103-110  get_logp         -> orc_get_logp()
 *   state.py:218-248  setup_sim_vars   -> orc_setup_vars()
 *   state.py:253-294  get_chi2_d_dd / get_logp_d_dd -> orc_get_logp_d_dd()
 *   state.py:299-315  priorHard        -> orc_prior_hard()
 *
 * PARITY PIN: tests/test_oracle_kat.py checks this file against every golden
 * value the reference holds for the path (16-digit initial
 * conditions, logp=-2.41616612321 on HD155358.vels, two 1000-point REBOUND RV
 * curves, three Encounter parameter vectors).  Gradient/Hessian values are
 * NOT pinned by any reference output ("parity unpinned" for derivatives);
 * they are cross-checked against finite differences of the pinned likelihood.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library.
 *
 * Build: see oracle/Makefile  (gcc -O2 -ffp-contract=off -fopenmp -shared).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_OK          0
#define ORC_PRIOR       1
#define ORC_ENCOUNTER   3   /* REBOUND's REB_EXIT_ENCOUNTER */
#define ORC_NONFINITE   8

/* internal run states (REBOUND: REB_RUNNING=-1, REB_RUNNING_LAST_STEP=-3) */
#define ST_RUNNING     (-1)
#define ST_LAST_STEP   (-3)
#define ST_SUCCESS       0

/* element slots of one planet */
enum { EL_M = 0, EL_A, EL_H, EL_K, EL_L, EL_IX, EL_IY, EL_N };

typedef struct { double m, x, y, z, vx, vy, vz, ax, ay, az; } part_t;
typedef struct { int order, index, ia, ib; } varcfg_t;

/* Gauss-Radau spacings (Everhart 1985; the table IAS15 is built on) */
static const double H[8] = {
    0.0, 0.0562625605369221464656521910318, 0.180240691736892364987579942780,
    0.352624717113169637373907769648, 0.547153626330555383001448554766,
    0.734210177215410531523210605558, 0.885320946839095768090359771030,
    0.977520613561287501891174488626};
/* RR[n][i] = h_n - h_i ; CC[j][k]: b_k += dg_j * CC[j][k] (k<j) ; DD[k][j]: g_j = sum_k b_k DD[k][j] */
static double RR[8][8], CC[7][7], DD[7][7];
static int tables_ready = 0;

static void build_tables(void) {
    if (tables_ready) return;
    long double c[7][7], d[7][7];
    memset(c, 0, sizeof c); memset(d, 0, sizeof d);
    for (int n = 1; n < 8; n++)
        for (int i = 0; i < n; i++) RR[n][i] = (double)((long double)H[n] - (long double)H[i]);
    /* c[j][k] = coeff of s^k in prod_{i=1..j} (s - h_i); d = inverse map (Everhart eqs. 7-8) */
    for (int j = 0; j < 7; j++) { c[j][j] = 1.0L; d[j][j] = 1.0L; }
    for (int j = 1; j < 7; j++) {
        c[j][0] = -(long double)H[j] * c[j - 1][0];
        d[j][0] = (long double)H[1] * d[j - 1][0];
        for (int k = 1; k < j; k++) {
            c[j][k] = c[j - 1][k - 1] - (long double)H[j] * c[j - 1][k];
            d[j][k] = d[j - 1][k - 1] + (long double)H[k + 1] * d[j - 1][k];
        }
    }
    for (int j = 0; j < 7; j++)
        for (int k = 0; k < 7; k++) { CC[j][k] = (double)c[j][k]; DD[j][k] = (double)d[j][k]; }
    tables_ready = 1;
}

typedef struct {
    int N, N_var, nvc;           /* N counts real + variational particles */
    varcfg_t *vc;
    part_t *p;
    double t, dt, dt_last_done, G;
    int status;
    double exit_min_distance;
    double epsilon, min_dt;
    int var_in_norm;             /* 1: 2017-era norms over all particles; 0: real particles only */
    int N3;
    double *x0, *v0, *a0, *at, *csx, *csv;
    double *b[7], *g[7], *e[7], *br[7], *er[7];
    double *store;
    long n_force, n_steps, n_reject;
} sim_t;

static void sim_init(sim_t *s, int Nmax, int nvc_max) {
    build_tables();
    memset(s, 0, sizeof *s);
    s->p = (part_t *)calloc((size_t)Nmax, sizeof(part_t));
    s->vc = (varcfg_t *)calloc((size_t)(nvc_max > 0 ? nvc_max : 1), sizeof(varcfg_t));
    s->N3 = 3 * Nmax;
    size_t per = (size_t)s->N3;
    s->store = (double *)calloc(per * (6 + 35), sizeof(double));
    double *q = s->store;
    s->x0 = q; q += per; s->v0 = q; q += per; s->a0 = q; q += per;
    s->at = q; q += per; s->csx = q; q += per; s->csv = q; q += per;
    for (int k = 0; k < 7; k++) { s->b[k] = q; q += per; }
    for (int k = 0; k < 7; k++) { s->g[k] = q; q += per; }
    for (int k = 0; k < 7; k++) { s->e[k] = q; q += per; }
    for (int k = 0; k < 7; k++) { s->br[k] = q; q += per; }
    for (int k = 0; k < 7; k++) { s->er[k] = q; q += per; }
    /* rebound.Simulation() defaults (state.py:37): G=1, t=0, dt=1e-3, IAS15, eps=1e-9 */
    s->G = 1.0; s->t = 0.0; s->dt = 0.001; s->dt_last_done = 0.0;
    s->epsilon = 1e-9; s->min_dt = 0.0; s->status = ST_RUNNING;
    s->var_in_norm = 0;
}
static void sim_free(sim_t *s) { free(s->p); free(s->vc); free(s->store); }

/* ------------------------------------------------------------------ */
/* Pal (2009) elements -> cartesian, relative to a primary at rest at the origin
 * with mass Mp (rebound: reb_tools_pal_to_particle; reached from
 * state.py:41 `sim.add(primary=sim.particles[0], **planet)`).
 * Written over a 2nd-order jet type so the same code yields the variational
 * initial conditions (vary(p,e) / vary(p,e1,e2), state.py:236,245).       */
typedef struct { double v, d1, d2, d12; } jet;   /* value, d/da, d/db, d2/dadb */
static inline jet J(double v) { jet r = {v, 0, 0, 0}; return r; }
static inline jet jadd(jet a, jet b) { jet r = {a.v + b.v, a.d1 + b.d1, a.d2 + b.d2, a.d12 + b.d12}; return r; }
static inline jet jsub(jet a, jet b) { jet r = {a.v - b.v, a.d1 - b.d1, a.d2 - b.d2, a.d12 - b.d12}; return r; }
static inline jet jneg(jet a) { jet r = {-a.v, -a.d1, -a.d2, -a.d12}; return r; }
static inline jet jmul(jet a, jet b) {
    jet r = {a.v * b.v, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2,
             a.d12 * b.v + a.d1 * b.d2 + a.d2 * b.d1 + a.v * b.d12};
    return r;
}
static inline jet jscale(double s, jet a) { jet r = {s * a.v, s * a.d1, s * a.d2, s * a.d12}; return r; }
/* f(a) given f, f', f'' at a.v */
static inline jet jchain(jet a, double f, double f1, double f2) {
    jet r = {f, f1 * a.d1, f1 * a.d2, f1 * a.d12 + f2 * a.d1 * a.d2};
    return r;
}
static inline jet jinv(jet a) { double i = 1.0 / a.v; return jchain(a, i, -i * i, 2 * i * i * i); }
static inline jet jdiv(jet a, jet b) { return jmul(a, jinv(b)); }
static inline jet jsqrt(jet a) { double s = sqrt(a.v); return jchain(a, s, 0.5 / s, -0.25 / (s * a.v)); }
static inline jet jsin(jet a) { double s = sin(a.v), c = cos(a.v); return jchain(a, s, c, -s); }
static inline jet jcos(jet a) { double s = sin(a.v), c = cos(a.v); return jchain(a, c, -s, -c); }

/* Kepler's equation in Pal form:  p - k sin(l+p) + h cos(l+p) = 0. */
static double solve_kepler_pal(double h, double k, double l) {
    double e2 = h * h + k * k, p;
    if (e2 < 0.09) {
        p = 0.0;
    } else { /* start from an eccentric-anomaly guess (Danby) */
        double e = sqrt(e2), w = atan2(h, k), M = l - w;
        M = fmod(M, 2 * M_PI); if (M > M_PI) M -= 2 * M_PI; if (M < -M_PI) M += 2 * M_PI;
        double E = M + 0.85 * e * (sin(M) >= 0 ? 1.0 : -1.0);
        p = E - M;
    }
    for (int it = 0; it < 60; it++) {
        double s = sin(l + p), c = cos(l + p);
        double f = p - k * s + h * c, f1 = 1.0 - k * c - h * s;
        double dp = -f / f1;
        p += dp;
        if (fabs(dp) < 1e-16 * (1.0 + fabs(p))) break;
    }
    return p;
}

typedef struct { jet m, x, y, z, vx, vy, vz; } jpart;

static jpart pal_to_particle_jet(double G, double Mprim, jet m, jet a, jet l, jet k, jet h, jet ix, jet iy) {
    /* p is an implicit function of (h,k,l): differentiate f(p;h,k,l)=0 via the jet Newton step. */
    double pv = solve_kepler_pal(h.v, k.v, l.v);
    jet p = J(pv);
    /* two jet-Newton corrections starting from the converged value give exact 1st and 2nd derivatives */
    for (int it = 0; it < 3; it++) {
        jet F = jadd(l, p);
        jet s = jsin(F), c = jcos(F);
        jet f = jadd(jsub(p, jmul(k, s)), jmul(h, c));
        jet f1 = jsub(jsub(J(1.0), jmul(k, c)), jmul(h, s));
        p = jsub(p, jdiv(f, f1));
        p.v = pv; /* value is already converged; keep it bit-stable */
    }
    jet F = jadd(l, p);
    jet slp = jsin(F), clp = jcos(F);
    jet q = jadd(jmul(k, clp), jmul(h, slp));
    jet one = J(1.0), two = J(2.0);
    jet lp = jsub(one, jsqrt(jsub(jsub(one, jmul(h, h)), jmul(k, k))));
    jet p2l = jdiv(p, jsub(two, lp));
    jet xi = jmul(a, jsub(jadd(clp, jmul(p2l, h)), k));
    jet eta = jmul(a, jsub(jsub(slp, jmul(p2l, k)), h));
    jet izarg = jsub(jsub(J(4.0), jmul(ix, ix)), jmul(iy, iy));
    if (izarg.v < 0) izarg = jneg(izarg);
    jet iz = jsqrt(izarg);
    jet W = jsub(jmul(eta, ix), jmul(xi, iy));
    jpart o;
    o.m = m;
    o.x = jadd(xi, jscale(0.5, jmul(iy, W)));
    o.y = jsub(eta, jscale(0.5, jmul(ix, W)));
    o.z = jscale(0.5, jmul(iz, W));
    jet an = jsqrt(jdiv(jscale(G, jadd(m, J(Mprim))), a));
    jet q2l = jdiv(q, jsub(two, lp));
    jet pref = jdiv(an, jsub(one, q));
    jet dxi = jmul(pref, jadd(jneg(slp), jmul(q2l, h)));
    jet deta = jmul(pref, jsub(clp, jmul(q2l, k)));
    jet dW = jsub(jmul(deta, ix), jmul(dxi, iy));
    o.vx = jadd(dxi, jscale(0.5, jmul(iy, dW)));
    o.vy = jsub(deta, jscale(0.5, jmul(ix, dW)));
    o.vz = jscale(0.5, jmul(iz, dW));
    return o;
}

/* Barycentric jets of star + planets: setup_sim (state.py:36-47) = star m=1 at origin,
 * planets relative to the star, move_to_com.  elems: [P][7] (m,a,h,k,l,ix,iy).
 * Directions: (pa,ea) gets d1=1, (pb,eb) gets d2=1 (pass pa<0 for none).  */
static void barycentric_jets(int P, const double *elems, int pa, int ea, int pb, int eb, jpart *out /*[P+1]*/) {
    jpart *bod = out;
    memset(bod, 0, sizeof(jpart) * (size_t)(P + 1));
    bod[0].m = J(1.0);
    for (int i = 0; i < P; i++) {
        jet el[EL_N];
        for (int e = 0; e < EL_N; e++) {
            el[e] = J(elems[i * EL_N + e]);
            if (i == pa && e == ea) el[e].d1 = 1.0;
            if (i == pb && e == eb) el[e].d2 = 1.0;
        }
        bod[i + 1] = pal_to_particle_jet(1.0, 1.0, el[EL_M], el[EL_A], el[EL_L], el[EL_K], el[EL_H], el[EL_IX], el[EL_IY]);
    }
    /* move_to_com: x_i -= sum m_j x_j / sum m_j */
    jet mt = J(0), cx = J(0), cy = J(0), cz = J(0), cvx = J(0), cvy = J(0), cvz = J(0);
    for (int i = 0; i <= P; i++) {
        mt = jadd(mt, bod[i].m);
        cx = jadd(cx, jmul(bod[i].m, bod[i].x)); cy = jadd(cy, jmul(bod[i].m, bod[i].y)); cz = jadd(cz, jmul(bod[i].m, bod[i].z));
        cvx = jadd(cvx, jmul(bod[i].m, bod[i].vx)); cvy = jadd(cvy, jmul(bod[i].m, bod[i].vy)); cvz = jadd(cvz, jmul(bod[i].m, bod[i].vz));
    }
    jet im = jinv(mt);
    cx = jmul(cx, im); cy = jmul(cy, im); cz = jmul(cz, im);
    cvx = jmul(cvx, im); cvy = jmul(cvy, im); cvz = jmul(cvz, im);
    for (int i = 0; i <= P; i++) {
        bod[i].x = jsub(bod[i].x, cx); bod[i].y = jsub(bod[i].y, cy); bod[i].z = jsub(bod[i].z, cz);
        bod[i].vx = jsub(bod[i].vx, cvx); bod[i].vy = jsub(bod[i].vy, cvy); bod[i].vz = jsub(bod[i].vz, cvz);
    }
}

/* priorHard (state.py:299-315) */
int orc_prior_hard(int P, const double *elems) {
    for (int i = 0; i < P; i++) {
        const double *el = elems + i * EL_N;
        if (el[EL_A] <= 0.02) return 1;
        if (el[EL_M] <= 5e-6) return 1;
        if (el[EL_H] * el[EL_H] + el[EL_K] * el[EL_K] >= 1.0) return 1;
        if (el[EL_IX] * el[EL_IX] + el[EL_IY] * el[EL_IY] >= 4.0) return 1;
    }
    return 0;
}

/* KAT-1 helper: cartesian state before and after move_to_com. out: [(P+1)][7] = m,x,y,z,vx,vy,vz */
int orc_initial_conditions(int P, const double *elems, double *out_com, double *out_rel) {
    jpart bod[8];
    if (P > 7) return -1;
    if (out_rel) {
        memset(out_rel, 0, sizeof(double) * 7);
        out_rel[0] = 1.0;
        for (int i = 0; i < P; i++) {
            const double *el = elems + i * EL_N;
            jpart q = pal_to_particle_jet(1.0, 1.0, J(el[EL_M]), J(el[EL_A]), J(el[EL_L]), J(el[EL_K]), J(el[EL_H]), J(el[EL_IX]), J(el[EL_IY]));
            double *o = out_rel + (i + 1) * 7;
            o[0] = q.m.v; o[1] = q.x.v; o[2] = q.y.v; o[3] = q.z.v; o[4] = q.vx.v; o[5] = q.vy.v; o[6] = q.vz.v;
        }
    }
    barycentric_jets(P, elems, -1, 0, -1, 0, bod);
    for (int i = 0; i <= P; i++) {
        double *o = out_com + i * 7;
        o[0] = bod[i].m.v; o[1] = bod[i].x.v; o[2] = bod[i].y.v; o[3] = bod[i].z.v;
        o[4] = bod[i].vx.v; o[5] = bod[i].vy.v; o[6] = bod[i].vz.v;
    }
    return 0;
}

/* ------------------------------------------------------------------ */
/* gravity: REB_GRAVITY_BASIC on the real particles + variational equations */
static void gravity(sim_t *s) {
    const int Nr = s->N - s->N_var;
    part_t *p = s->p;
    const double G = s->G;
    for (int i = 0; i < Nr; i++) { p[i].ax = 0; p[i].ay = 0; p[i].az = 0; }
    for (int i = 0; i < Nr; i++)
        for (int j = 0; j < Nr; j++) {
            if (i == j) continue;
            const double dx = p[i].x - p[j].x, dy = p[i].y - p[j].y, dz = p[i].z - p[j].z;
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            const double pref = -G / (r * r * r) * p[j].m;
            p[i].ax += pref * dx; p[i].ay += pref * dy; p[i].az += pref * dz;
        }
    /* a_i = -sum_j G m_j f(d), f(d)=d/r^3, d=r_i-r_j.
       Df[u]   = u/r^3 - 3 d (d.u)/r^5
       D2f[u,w]= -3[u(d.w) + w(d.u) + d(u.w)]/r^5 + 15 d (d.u)(d.w)/r^7             */
    for (int v = 0; v < s->nvc; v++) {
        const varcfg_t vc = s->vc[v];
        part_t *q = p + vc.index;
        for (int i = 0; i < Nr; i++) { q[i].ax = 0; q[i].ay = 0; q[i].az = 0; }
        for (int i = 0; i < Nr; i++)
            for (int j = i + 1; j < Nr; j++) {
                const double dx = p[i].x - p[j].x, dy = p[i].y - p[j].y, dz = p[i].z - p[j].z;
                const double r2 = dx * dx + dy * dy + dz * dz;
                const double r = sqrt(r2);
                const double r3i = 1.0 / (r2 * r), r5i = r3i / r2;
                const double Gmi = G * p[i].m, Gmj = G * p[j].m;
                /* linear operator on this set's own displacement */
                const double ux = q[i].x - q[j].x, uy = q[i].y - q[j].y, uz = q[i].z - q[j].z;
                const double du = dx * ux + dy * uy + dz * uz;
                double fx = ux * r3i - 3.0 * dx * du * r5i;
                double fy = uy * r3i - 3.0 * dy * du * r5i;
                double fz = uz * r3i - 3.0 * dz * du * r5i;
                /* accumulate "Df-like" vector A so that a_i -= Gm_j * A + dGm_j f ; a_j += Gm_i * A + dGm_i f */
                double mix = q[j].m * G * r3i, mjx = q[i].m * G * r3i; /* dGm_j/r^3 , dGm_i/r^3 */
                double aix = Gmj * fx + mix * dx, aiy = Gmj * fy + mix * dy, aiz = Gmj * fz + mix * dz;
                double ajx = Gmi * fx + mjx * dx, ajy = Gmi * fy + mjx * dy, ajz = Gmi * fz + mjx * dz;
                if (vc.order == 2) {
                    const part_t *qa = p + vc.ia, *qb = p + vc.ib;
                    const double r7i = r5i / r2;
                    const double ax_ = qa[i].x - qa[j].x, ay_ = qa[i].y - qa[j].y, az_ = qa[i].z - qa[j].z;
                    const double bx_ = qb[i].x - qb[j].x, by_ = qb[i].y - qb[j].y, bz_ = qb[i].z - qb[j].z;
                    const double da = dx * ax_ + dy * ay_ + dz * az_;
                    const double db = dx * bx_ + dy * by_ + dz * bz_;
                    const double ab = ax_ * bx_ + ay_ * by_ + az_ * bz_;
                    /* D2f[a,b] */
                    const double c5 = -3.0 * r5i, c7 = 15.0 * da * db * r7i;
                    const double sx = c5 * (ax_ * db + bx_ * da + dx * ab) + c7 * dx;
                    const double sy = c5 * (ay_ * db + by_ * da + dy * ab) + c7 * dy;
                    const double sz = c5 * (az_ * db + bz_ * da + dz * ab) + c7 * dz;
                    /* Df[a], Df[b] for the mass cross terms */
                    const double fax = ax_ * r3i - 3.0 * dx * da * r5i, fay = ay_ * r3i - 3.0 * dy * da * r5i, faz = az_ * r3i - 3.0 * dz * da * r5i;
                    const double fbx = bx_ * r3i - 3.0 * dx * db * r5i, fby = by_ * r3i - 3.0 * dy * db * r5i, fbz = bz_ * r3i - 3.0 * dz * db * r5i;
                    const double Gamj = G * qa[j].m, Gbmj = G * qb[j].m, Gami = G * qa[i].m, Gbmi = G * qb[i].m;
                    aix += Gmj * sx + Gamj * fbx + Gbmj * fax; aiy += Gmj * sy + Gamj * fby + Gbmj * fay; aiz += Gmj * sz + Gamj * fbz + Gbmj * faz;
                    ajx += Gmi * sx + Gami * fbx + Gbmi * fax; ajy += Gmi * sy + Gami * fby + Gbmi * fay; ajz += Gmi * sz + Gami * fbz + Gbmi * faz;
                }
                q[i].ax -= aix; q[i].ay -= aiy; q[i].az -= aiz;
                q[j].ax += ajx; q[j].ay += ajy; q[j].az += ajz;
            }
    }
    s->n_force++;
}

/* reb_run_heartbeat's exit_min_distance test (state.py:46 sets the distance) */
static void encounter_check(sim_t *s) {
    if (s->exit_min_distance == 0.0) return;
    const double min2 = s->exit_min_distance * s->exit_min_distance;
    const int Nr = s->N - s->N_var;
    for (int i = 0; i < Nr; i++)
        for (int j = 0; j < i; j++) {
            const double x = s->p[i].x - s->p[j].x, y = s->p[i].y - s->p[j].y, z = s->p[i].z - s->p[j].z;
            if (x * x + y * y + z * z < min2) s->status = ORC_ENCOUNTER;
        }
}

static void predict_next_step(double ratio, int N3, double *const *_e, double *const *_b, double **e, double **b) {
    if (ratio > 20.0) {
        for (int j = 0; j < 7; j++) for (int k = 0; k < N3; k++) { e[j][k] = 0; b[j][k] = 0; }
        return;
    }
    const double q1 = ratio, q2 = q1 * q1, q3 = q1 * q2, q4 = q2 * q2, q5 = q2 * q3, q6 = q3 * q3, q7 = q3 * q4;
    for (int k = 0; k < N3; k++) {
        const double be0 = _b[0][k] - _e[0][k], be1 = _b[1][k] - _e[1][k], be2 = _b[2][k] - _e[2][k], be3 = _b[3][k] - _e[3][k];
        const double be4 = _b[4][k] - _e[4][k], be5 = _b[5][k] - _e[5][k], be6 = _b[6][k] - _e[6][k];
        e[0][k] = q1 * (_b[6][k] * 7.0 + _b[5][k] * 6.0 + _b[4][k] * 5.0 + _b[3][k] * 4.0 + _b[2][k] * 3.0 + _b[1][k] * 2.0 + _b[0][k]);
        e[1][k] = q2 * (_b[6][k] * 21.0 + _b[5][k] * 15.0 + _b[4][k] * 10.0 + _b[3][k] * 6.0 + _b[2][k] * 3.0 + _b[1][k]);
        e[2][k] = q3 * (_b[6][k] * 35.0 + _b[5][k] * 20.0 + _b[4][k] * 10.0 + _b[3][k] * 4.0 + _b[2][k]);
        e[3][k] = q4 * (_b[6][k] * 35.0 + _b[5][k] * 15.0 + _b[4][k] * 5.0 + _b[3][k]);
        e[4][k] = q5 * (_b[6][k] * 21.0 + _b[5][k] * 6.0 + _b[4][k]);
        e[5][k] = q6 * (_b[6][k] * 7.0 + _b[5][k]);
        e[6][k] = q7 * _b[6][k];
        b[0][k] = e[0][k] + be0; b[1][k] = e[1][k] + be1; b[2][k] = e[2][k] + be2; b[3][k] = e[3][k] + be3;
        b[4][k] = e[4][k] + be4; b[5][k] = e[5][k] + be5; b[6][k] = e[6][k] + be6;
    }
}

static int is_normal(double x) { return isnormal(x); }

/* One IAS15 step attempt.  Returns 1 if accepted, 0 if rejected (caller retries). */
static int ias15_attempt(sim_t *s) {
    const int N = s->N, N3 = 3 * N;
    part_t *p = s->p;
    double **b = s->b, **g = s->g, **e = s->e;
    double *x0 = s->x0, *v0 = s->v0, *a0 = s->a0, *at = s->at, *csx = s->csx, *csv = s->csv;
    for (int k = 0; k < N; k++) {
        x0[3 * k] = p[k].x; x0[3 * k + 1] = p[k].y; x0[3 * k + 2] = p[k].z;
        v0[3 * k] = p[k].vx; v0[3 * k + 1] = p[k].vy; v0[3 * k + 2] = p[k].vz;
        a0[3 * k] = p[k].ax; a0[3 * k + 1] = p[k].ay; a0[3 * k + 2] = p[k].az;
    }
    /* g from b (Everhart eq. 7), same association order as rebound's tabulated form: b6*d + b5*d + ... + b_j */
    for (int k = 0; k < N3; k++)
        for (int j = 0; j < 7; j++) {
            double t2 = 0.0;
            for (int m = 6; m > j; m--) t2 += b[m][k] * DD[m][j];
            g[j][k] = t2 + b[j][k];
        }
    const double t_beginning = s->t;
    double pc_err = 1e300, pc_err_last = 2.0;
    int iterations = 0;
    while (1) {
        if (pc_err < 1e-16) break;
        if (iterations > 2 && pc_err_last <= pc_err) break;
        if (iterations >= 12) break;
        pc_err_last = pc_err;
        pc_err = 0.0;
        iterations++;
        for (int n = 1; n < 8; n++) {
            double sc[9];
            sc[0] = s->dt * H[n];
            sc[1] = sc[0] * sc[0] / 2.0;
            sc[2] = sc[1] * H[n] / 3.0;
            sc[3] = sc[2] * H[n] / 2.0;
            sc[4] = 3.0 * sc[3] * H[n] / 5.0;
            sc[5] = 2.0 * sc[4] * H[n] / 3.0;
            sc[6] = 5.0 * sc[5] * H[n] / 7.0;
            sc[7] = 3.0 * sc[6] * H[n] / 4.0;
            sc[8] = 7.0 * sc[7] * H[n] / 9.0;
            s->t = t_beginning + sc[0];
            for (int i = 0; i < N; i++)
                for (int c = 0; c < 3; c++) {
                    const int k = 3 * i + c;
                    double xk = -csx[k] + (sc[8] * b[6][k] + sc[7] * b[5][k] + sc[6] * b[4][k] + sc[5] * b[3][k] + sc[4] * b[2][k] + sc[3] * b[1][k] + sc[2] * b[0][k] + sc[1] * a0[k] + sc[0] * v0[k]);
                    double val = xk + x0[k];
                    if (c == 0) p[i].x = val; else if (c == 1) p[i].y = val; else p[i].z = val;
                }
            gravity(s);
            for (int k = 0; k < N; k++) { at[3 * k] = p[k].ax; at[3 * k + 1] = p[k].ay; at[3 * k + 2] = p[k].az; }
            double maxak = 0.0, maxb6 = 0.0;
            for (int k = 0; k < N3; k++) {
                double tmp = g[n - 1][k];
                double gk = at[k] - a0[k];
                double val = gk / RR[n][0];
                for (int i = 1; i < n; i++) val = (val - g[i - 1][k]) / RR[n][i];
                g[n - 1][k] = val;
                tmp = val - tmp;
                for (int i = 0; i < n - 1; i++) b[i][k] += tmp * CC[n - 1][i];
                b[n - 1][k] += tmp;
                if (n == 7) {
                    const double ak = fabs(at[k]);
                    if (is_normal(ak) && ak > maxak) maxak = ak;
                    const double b6 = fabs(tmp);
                    if (is_normal(b6) && b6 > maxb6) maxb6 = b6;
                }
            }
            if (n == 7) pc_err = maxb6 / maxak;
        }
    }
    s->t = t_beginning;
    const double safety = 0.25;
    const double dt_done = s->dt;
    double dt_new;
    {
        double maxak = 0.0, maxb6 = 0.0;
        const int Nn = s->var_in_norm ? N : N - s->N_var;
        for (int i = 0; i < Nn; i++) {
            const double v2 = p[i].vx * p[i].vx + p[i].vy * p[i].vy + p[i].vz * p[i].vz;
            const double x2 = p[i].x * p[i].x + p[i].y * p[i].y + p[i].z * p[i].z;
            if (fabs(v2 * s->dt * s->dt / x2) < 1e-16) continue;
            for (int k = 3 * i; k < 3 * (i + 1); k++) {
                const double ak = fabs(at[k]);
                if (is_normal(ak) && ak > maxak) maxak = ak;
                const double b6 = fabs(b[6][k]);
                if (is_normal(b6) && b6 > maxb6) maxb6 = b6;
            }
        }
        const double err = maxb6 / maxak;
        if (is_normal(err)) dt_new = pow(s->epsilon / err, 1.0 / 7.0) * dt_done;
        else dt_new = dt_done / safety;
        if (fabs(dt_new) < s->min_dt) dt_new = copysign(s->min_dt, dt_new);
        if (fabs(dt_new / dt_done) < safety) {
            for (int k = 0; k < N; k++) {
                p[k].x = x0[3 * k]; p[k].y = x0[3 * k + 1]; p[k].z = x0[3 * k + 2];
                p[k].vx = v0[3 * k]; p[k].vy = v0[3 * k + 1]; p[k].vz = v0[3 * k + 2];
                p[k].ax = a0[3 * k]; p[k].ay = a0[3 * k + 1]; p[k].az = a0[3 * k + 2];
            }
            s->dt = dt_new;
            if (s->dt_last_done != 0.0) {
                const double ratio = s->dt / s->dt_last_done;
                predict_next_step(ratio, N3, s->er, s->br, e, b);
            }
            s->n_reject++;
            return 0;
        }
        if (fabs(dt_new / dt_done) > 1.0)
            if (dt_new / dt_done > 1.0 / safety) dt_new = dt_done / safety;
        s->dt = dt_new;
    }
    const double dt2 = dt_done * dt_done;
    for (int k = 0; k < N3; k++) {
        {
            double a = x0[k];
            csx[k] += (b[6][k] / 72. + b[5][k] / 56. + b[4][k] / 42. + b[3][k] / 30. + b[2][k] / 20. + b[1][k] / 12. + b[0][k] / 6. + a0[k] / 2.) * dt2 + v0[k] * dt_done;
            x0[k] = a + csx[k];
            csx[k] += a - x0[k];
        }
        {
            double a = v0[k];
            csv[k] += (b[6][k] / 8. + b[5][k] / 7. + b[4][k] / 6. + b[3][k] / 5. + b[2][k] / 4. + b[1][k] / 3. + b[0][k] / 2. + a0[k]) * dt_done;
            v0[k] = a + csv[k];
            csv[k] += a - v0[k];
        }
    }
    s->t += dt_done;
    s->dt_last_done = dt_done;
    for (int k = 0; k < N; k++) {
        p[k].x = x0[3 * k]; p[k].y = x0[3 * k + 1]; p[k].z = x0[3 * k + 2];
        p[k].vx = v0[3 * k]; p[k].vy = v0[3 * k + 1]; p[k].vz = v0[3 * k + 2];
    }
    for (int j = 0; j < 7; j++) {
        memcpy(s->er[j], e[j], sizeof(double) * (size_t)N3);
        memcpy(s->br[j], b[j], sizeof(double) * (size_t)N3);
    }
    predict_next_step(s->dt / dt_done, N3, s->er, s->br, e, b);
    return 1;
}

/* reb_step: force at the start of the step, then retry attempts until one is accepted */
static void sim_step(sim_t *s) {
    gravity(s);
    for (;;) {
        s->n_steps++;
        if (ias15_attempt(s)) break;
        if (s->n_steps > 50000000L || !isfinite(s->dt) || s->dt == 0.0) { s->status = ORC_NONFINITE; break; }
    }
}

/* reb_check_exit with exact_finish_time=1 */
static int check_exit(sim_t *s, double tmax, double *last_full_dt) {
    const double sgn = copysign(1.0, s->dt);
    if (s->status >= 0) return s->status;
    if ((s->t + s->dt) * sgn >= tmax * sgn) {
        if (s->t == tmax) {
            s->status = ST_SUCCESS;
        } else if (s->status == ST_LAST_STEP) {
            double tscale = 1e-12 * fabs(tmax);
            if (tscale < 1e-200) tscale = 1e-12;
            if (fabs(s->t - tmax) < tscale) s->status = ST_SUCCESS;
            else s->dt = tmax - s->t;
        } else {
            s->status = ST_LAST_STEP;
            if (s->dt_last_done != 0.0) *last_full_dt = s->dt_last_done;
            s->dt = tmax - s->t;
        }
    } else if (s->status == ST_LAST_STEP) {
        s->status = ST_RUNNING;
    }
    return s->status;
}

/* sim.integrate(tmax) (state.py:71).  Returns 0 or ORC_ENCOUNTER / ORC_NONFINITE. */
static int sim_integrate(sim_t *s, double tmax) {
    double last_full_dt = s->dt;
    s->dt_last_done = 0.0;
    s->status = ST_RUNNING;
    encounter_check(s);
    while (check_exit(s, tmax, &last_full_dt) < 0) {
        sim_step(s);
        if (s->status == ORC_NONFINITE) break;
        encounter_check(s);
    }
    s->dt = last_full_dt;
    return s->status;
}

/* setup_sim (state.py:36-47) */
static void orc_setup(sim_t *s, int P, const double *elems, double hill_factor) {
    jpart bod[8];
    barycentric_jets(P, elems, -1, 0, -1, 0, bod);
    s->N = P + 1; s->N_var = 0; s->nvc = 0;
    double hmax = 0.0;
    for (int i = 0; i <= P; i++) {
        part_t *q = &s->p[i];
        q->m = bod[i].m.v; q->x = bod[i].x.v; q->y = bod[i].y.v; q->z = bod[i].z.v;
        q->vx = bod[i].vx.v; q->vy = bod[i].vy.v; q->vz = bod[i].vz.v;
        q->ax = q->ay = q->az = 0;
    }
    for (int i = 0; i < P; i++) {
        const double *el = elems + i * EL_N;
        double r = el[EL_A] * pow(el[EL_M] / (3.0 * 1.0), 1.0 / 3.0);
        if (r > hmax) hmax = r;
    }
    s->exit_min_distance = hill_factor * hmax;
}

/* get_rv (state.py:61-73).  counters (optional): [force evals, step attempts, rejections] */
int orc_get_rv(int P, const double *elems, double hill_factor, const double *times, int nt, double *rv, long *counters) {
    sim_t s;
    sim_init(&s, P + 1, 0);
    orc_setup(&s, P, elems, hill_factor);
    int st = ORC_OK;
    for (int i = 0; i < nt; i++) {
        int r = sim_integrate(&s, times[i]);
        if (r != 0) { st = r; break; }
        rv[i] = s.p[0].vx;
        if (!isfinite(rv[i])) { st = ORC_NONFINITE; break; }
    }
    if (counters) { counters[0] += s.n_force; counters[1] += s.n_steps; counters[2] += s.n_reject; }
    sim_free(&s);
    return st;
}

/* get_logp (state.py:103-110) = -get_chi2 (state.py:89-98); prior first.
 * leg_status[0]: forward leg, [1]: backward leg (the reference raises from whichever leg fails first;
 * forward runs first).                                                    */
int orc_get_logp(int P, const double *elems, double hill_factor,
                 const double *tf, const double *rvf, const double *ef, int nf,
                 const double *tb, const double *rvb, const double *eb, int nb,
                 double npoints, double *logp, long *counters, int *leg_status) {
    if (leg_status) { leg_status[0] = 0; leg_status[1] = 0; }
    if (orc_prior_hard(P, elems)) { *logp = -INFINITY; return ORC_PRIOR; }
    double *rf = (double *)malloc(sizeof(double) * (size_t)(nf + nb + 2));
    double *rb = rf + nf + 1;
    int st = orc_get_rv(P, elems, hill_factor, tf, nf, rf, counters);
    if (leg_status) leg_status[0] = st;
    if (st == 0) {
        st = orc_get_rv(P, elems, hill_factor, tb, nb, rb, counters);
        if (leg_status) leg_status[1] = st;
    }
    if (st != 0) { *logp = -INFINITY; free(rf); return st; }
    double chi2f = 0.0, chi2b = 0.0;
    for (int i = 0; i < nf; i++) chi2f += ((rf[i] - rvf[i]) * (rf[i] - rvf[i])) / (ef[i] * ef[i]);
    for (int i = 0; i < nb; i++) chi2b += ((rb[i] - rvb[i]) * (rb[i] - rvb[i])) / (eb[i] * eb[i]);
    *logp = -((chi2b + chi2f) / npoints);
    free(rf);
    return ORC_OK;
}

/* ------------------------------------------------------------------ */
/* setup_sim_vars (state.py:229-248): real set + nvars 1st-order sets + nvars(nvars+1)/2 2nd-order sets
 * (vindex1 >= vindex2, row-major in vindex1), every set an exact derivative of the barycentric ICs. */
static void orc_setup_vars(sim_t *s, int P, const double *elems, double hill_factor,
                           int nvars, const int *fp, const int *fe) {
    const int Nr = P + 1;
    orc_setup(s, P, elems, hill_factor);
    const int nsets2 = nvars * (nvars + 1) / 2;
    s->N = Nr * (1 + nvars + nsets2);
    s->N_var = s->N - Nr;
    s->nvc = nvars + nsets2;
    jpart bod[8];
    for (int v = 0; v < nvars; v++) {
        barycentric_jets(P, elems, fp[v], fe[v], -1, 0, bod);
        s->vc[v].order = 1; s->vc[v].index = Nr * (1 + v); s->vc[v].ia = s->vc[v].ib = 0;
        for (int i = 0; i < Nr; i++) {
            part_t *q = &s->p[Nr * (1 + v) + i];
            q->m = bod[i].m.d1; q->x = bod[i].x.d1; q->y = bod[i].y.d1; q->z = bod[i].z.d1;
            q->vx = bod[i].vx.d1; q->vy = bod[i].vy.d1; q->vz = bod[i].vz.d1; q->ax = q->ay = q->az = 0;
        }
    }
    int v2 = 0;
    for (int v1i = 0; v1i < nvars; v1i++)
        for (int v2i = 0; v2i <= v1i; v2i++) {
            barycentric_jets(P, elems, fp[v1i], fe[v1i], fp[v2i], fe[v2i], bod);
            varcfg_t *c = &s->vc[nvars + v2];
            c->order = 2; c->index = Nr * (1 + nvars + v2); c->ia = Nr * (1 + v1i); c->ib = Nr * (1 + v2i);
            for (int i = 0; i < Nr; i++) {
                part_t *q = &s->p[c->index + i];
                q->m = bod[i].m.d12; q->x = bod[i].x.d12; q->y = bod[i].y.d12; q->z = bod[i].z.d12;
                q->vx = bod[i].vx.d12; q->vy = bod[i].vy.d12; q->vz = bod[i].vz.d12; q->ax = q->ay = q->az = 0;
            }
            v2++;
        }
}

/* one leg of get_chi2_d_dd (state.py:262-284): epochs visited in the given order */
static int var_leg(sim_t *s, int nvars, const double *t, const double *rv, const double *er, int n, int reversed,
                   double fac, double *chi2, double *d, double *dd) {
    const int Nr = s->N - s->N_var;
    for (int ii = 0; ii < n; ii++) {
        const int i = reversed ? n - 1 - ii : ii;
        int r = sim_integrate(s, t[i]);
        if (r != 0) return r;
        const double vx = s->p[0].vx, res = vx - rv[i];
        *chi2 += res * res * 1. / (er[i] * er[i] * fac);
        int v2 = 0;
        for (int a = 0; a < nvars; a++) {
            const double da = s->p[Nr * (1 + a)].vx;
            d[a] += 2. * da * res * 1. / (er[i] * er[i] * fac);
            for (int b = 0; b <= a; b++) {
                const double db = s->p[Nr * (1 + b)].vx;
                const double dab = s->p[Nr * (1 + nvars + v2)].vx;
                dd[a * nvars + b] += 2. * dab * res * 1. / (er[i] * er[i] * fac) + 2. * da * db * 1. / (er[i] * er[i] * fac);
                dd[b * nvars + a] = dd[a * nvars + b];
                v2++;
            }
        }
    }
    return 0;
}

/* get_logp_d_dd (state.py:290-294).  fp/fe: free parameter -> (planet index 0-based, element slot).
 * grad[nvars], hess[nvars*nvars] row-major.  var_in_norm: see sim_t.  NB the reference does NOT apply
 * priorHard here (state.py:290); callers do (mcmc.py:171).                                       */
int orc_get_logp_d_dd(int P, const double *elems, double hill_factor, int nvars, const int *fp, const int *fe,
                      const double *tf, const double *rvf, const double *ef, int nf,
                      const double *tb, const double *rvb, const double *eb, int nb,
                      double npoints, int var_in_norm, double *logp, double *grad, double *hess, long *counters) {
    const int Nr = P + 1, nsets = 1 + nvars + nvars * (nvars + 1) / 2;
    double chi2f = 0, chi2b = 0;
    double *df = (double *)calloc((size_t)(2 * nvars + 2 * nvars * nvars), sizeof(double));
    double *db = df + nvars, *ddf = db + nvars, *ddb = ddf + nvars * nvars;
    int st;
    {
        sim_t s;
        sim_init(&s, Nr * nsets, nsets);
        s.var_in_norm = var_in_norm;
        orc_setup_vars(&s, P, elems, hill_factor, nvars, fp, fe);
        st = var_leg(&s, nvars, tf, rvf, ef, nf, 0, npoints, &chi2f, df, ddf);
        if (counters) { counters[0] += s.n_force; counters[1] += s.n_steps; counters[2] += s.n_reject; }
        sim_free(&s);
    }
    if (st == 0) {
        sim_t s;
        sim_init(&s, Nr * nsets, nsets);
        s.var_in_norm = var_in_norm;
        orc_setup_vars(&s, P, elems, hill_factor, nvars, fp, fe);
        st = var_leg(&s, nvars, tb, rvb, eb, nb, 1, npoints, &chi2b, db, ddb);
        if (counters) { counters[0] += s.n_force; counters[1] += s.n_steps; counters[2] += s.n_reject; }
        sim_free(&s);
    }
    if (st == 0) {
        *logp = -(chi2b + chi2f);
        for (int a = 0; a < nvars; a++) grad[a] = -(db[a] + df[a]);
        for (int a = 0; a < nvars * nvars; a++) hess[a] = -(ddb[a] + ddf[a]);
    } else {
        *logp = -INFINITY;
    }
    free(df);
    return st;
}

/* Variational initial conditions for tests: out[nsets][P+1][7] (m,x,y,z,vx,vy,vz) */
int orc_var_initial_conditions(int P, const double *elems, int nvars, const int *fp, const int *fe, double *out) {
    const int Nr = P + 1, nsets = 1 + nvars + nvars * (nvars + 1) / 2;
    sim_t s;
    sim_init(&s, Nr * nsets, nsets);
    orc_setup_vars(&s, P, elems, 0.0, nvars, fp, fe);
    for (int i = 0; i < s.N; i++) {
        double *o = out + 7 * i;
        o[0] = s.p[i].m; o[1] = s.p[i].x; o[2] = s.p[i].y; o[3] = s.p[i].z; o[4] = s.p[i].vx; o[5] = s.p[i].vy; o[6] = s.p[i].vz;
    }
    sim_free(&s);
    return 0;
}

/* ------------------------------------------------------------------ */
/* Batched helpers (theta -> elements through the model's free map), OpenMP over walkers.
 * model: fixed[P][7] defaults, free map (fp[v], fe[v]) overwritten from theta[w][v].       */
static void theta_to_elems(int P, const double *fixed, int nvars, const int *fp, const int *fe, const double *theta, double *elems) {
    memcpy(elems, fixed, sizeof(double) * (size_t)(P * EL_N));
    for (int v = 0; v < nvars; v++) elems[fp[v] * EL_N + fe[v]] = theta[v];
}

int orc_logp_batch(int P, const double *fixed, int nvars, const int *fp, const int *fe, double hill_factor,
                   const double *tf, const double *rvf, const double *ef, int nf,
                   const double *tb, const double *rvb, const double *eb, int nb, double npoints,
                   const double *theta, long W, double *logp, int *status, long *counters, int nthreads) {
    long c0 = 0, c1 = 0, c2 = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads) reduction(+ : c0, c1, c2)
    for (long w = 0; w < W; w++) {
        double elems[8 * EL_N];
        long c[3] = {0, 0, 0};
        theta_to_elems(P, fixed, nvars, fp, fe, theta + w * nvars, elems);
        status[w] = orc_get_logp(P, elems, hill_factor, tf, rvf, ef, nf, tb, rvb, eb, nb, npoints, &logp[w], c, NULL);
        c0 += c[0]; c1 += c[1]; c2 += c[2];
    }
    if (counters) { counters[0] += c0; counters[1] += c1; counters[2] += c2; }
    return 0;
}

int orc_logp_d_dd_batch(int P, const double *fixed, int nvars, const int *fp, const int *fe, double hill_factor,
                        const double *tf, const double *rvf, const double *ef, int nf,
                        const double *tb, const double *rvb, const double *eb, int nb, double npoints, int var_in_norm,
                        const double *theta, long W, double *logp, double *grad, double *hess, int *status,
                        long *counters, int nthreads) {
    long c0 = 0, c1 = 0, c2 = 0;
    if (nthreads < 1) nthreads = 1;
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads) reduction(+ : c0, c1, c2)
    for (long w = 0; w < W; w++) {
        double elems[8 * EL_N];
        long c[3] = {0, 0, 0};
        theta_to_elems(P, fixed, nvars, fp, fe, theta + w * nvars, elems);
        if (orc_prior_hard(P, elems)) { status[w] = ORC_PRIOR; logp[w] = -INFINITY; continue; }
        status[w] = orc_get_logp_d_dd(P, elems, hill_factor, nvars, fp, fe, tf, rvf, ef, nf, tb, rvb, eb, nb, npoints,
                                      var_in_norm, &logp[w], grad + w * nvars, hess + w * nvars * nvars, c);
        c0 += c[0]; c1 += c[1]; c2 += c[2];
    }
    if (counters) { counters[0] += c0; counters[1] += c1; counters[2] += c2; }
    return 0;
}
