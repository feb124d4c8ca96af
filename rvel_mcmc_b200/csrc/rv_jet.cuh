// rv_jet.cuh -- second-order jets (value, d/da, d/db, d2/dadb) for the variational initial conditions.
//
// The reference builds them with rebound's vary(pindex, vname[, vname2]) + move_to_com()
// (state.py:229-248); the net effect is "every variational particle is the exact derivative of the
// barycentric initial conditions with respect to the free parameters" (SURVEY App. A.5), which is what
// pushing jets through Pal->cartesian + COM shift computes directly.
#pragma once
#include "rv_core.cuh"

namespace rv {

struct Jet { double v, d1, d2, d12; };

RV_HD Jet J(double v) { return Jet{v, 0.0, 0.0, 0.0}; }
RV_HD Jet operator+(Jet a, Jet b) { return Jet{a.v + b.v, a.d1 + b.d1, a.d2 + b.d2, a.d12 + b.d12}; }
RV_HD Jet operator-(Jet a, Jet b) { return Jet{a.v - b.v, a.d1 - b.d1, a.d2 - b.d2, a.d12 - b.d12}; }
RV_HD Jet operator-(Jet a) { return Jet{-a.v, -a.d1, -a.d2, -a.d12}; }
RV_HD Jet operator*(Jet a, Jet b) {
    return Jet{a.v * b.v, a.d1 * b.v + a.v * b.d1, a.d2 * b.v + a.v * b.d2,
               a.d12 * b.v + a.d1 * b.d2 + a.d2 * b.d1 + a.v * b.d12};
}
RV_HD Jet operator*(double s, Jet a) { return Jet{s * a.v, s * a.d1, s * a.d2, s * a.d12}; }
RV_HD Jet chain(Jet a, double f, double f1, double f2) {
    return Jet{f, f1 * a.d1, f1 * a.d2, f1 * a.d12 + f2 * a.d1 * a.d2};
}
RV_HD Jet jinv(Jet a) { const double i = 1.0 / a.v; return chain(a, i, -i * i, 2.0 * i * i * i); }
RV_HD Jet operator/(Jet a, Jet b) { return a * jinv(b); }
RV_HD Jet jsqrt(Jet a) { const double s = sqrt(a.v); return chain(a, s, 0.5 / s, -0.25 / (s * a.v)); }
RV_HD void jsincos(Jet a, Jet& s, Jet& c) {
    double sv, cv;
    sincos(a.v, &sv, &cv);
    s = chain(a, sv, cv, -sv);
    c = chain(a, cv, -sv, -cv);
}

struct JState { Jet m, x[3], v[3]; };

// Pal elements (as jets) -> cartesian relative to a primary of mass Mp at rest at the origin.
RV_HD JState pal_to_cart_jet(const Jet* el, double Mp) {
    const Jet m = el[EL_M], a = el[EL_A], h = el[EL_H], k = el[EL_K], l = el[EL_L], ix = el[EL_IX], iy = el[EL_IY];
    double slv, clv, pv;
    kepler_pal(h.v, k.v, l.v, slv, clv, pv);
    Jet p = J(pv);
    // p(h,k,l) is implicit: two jet-Newton corrections from the converged value give exact 1st/2nd derivatives
    for (int it = 0; it < 3; it++) {
        Jet s, c;
        jsincos(l + p, s, c);
        const Jet f = p - k * s + h * c;
        const Jet f1 = J(1.0) - k * c - h * s;
        p = p - f / f1;
        p.v = pv;
    }
    Jet slp, clp;
    jsincos(l + p, slp, clp);
    const Jet q = k * clp + h * slp;
    const Jet one = J(1.0), two = J(2.0);
    const Jet lp = one - jsqrt(one - h * h - k * k);
    const Jet p2l = p / (two - lp);
    const Jet xi = a * (clp + p2l * h - k);
    const Jet eta = a * (slp - p2l * k - h);
    Jet izarg = J(4.0) - ix * ix - iy * iy;
    if (izarg.v < 0.0) izarg = -izarg;
    const Jet iz = jsqrt(izarg);
    const Jet W = eta * ix - xi * iy;
    JState o;
    o.m = m;
    o.x[0] = xi + 0.5 * (iy * W);
    o.x[1] = eta - 0.5 * (ix * W);
    o.x[2] = 0.5 * (iz * W);
    const Jet an = jsqrt((m + J(Mp)) / a);
    const Jet q2l = q / (two - lp);
    const Jet pref = an / (one - q);
    const Jet dxi = pref * (-slp + q2l * h);
    const Jet deta = pref * (clp - q2l * k);
    const Jet dW = deta * ix - dxi * iy;
    o.v[0] = dxi + 0.5 * (iy * dW);
    o.v[1] = deta - 0.5 * (ix * dW);
    o.v[2] = 0.5 * (iz * dW);
    return o;
}

}  // namespace rv
