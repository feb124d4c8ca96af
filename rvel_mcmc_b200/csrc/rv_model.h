// rv_model.h -- host-side construction of rv::Model from the reference's parameter schema
// (state.py:8-31: per-planet element dicts, free-variable list = keys minus ignored ones).
#pragma once
#include <string.h>
#include "rv_core.cuh"

namespace rv {

// fixed: [P][7] element values (m,a,h,k,l,ix,iy; absent keys = 0, as rebound's sim.add defaults);
// free_planet/free_elem: the theta vector's slots in order.  dims: 0 = auto (2 when every ix,iy is
// pinned to zero, else 3).  Returns 0 or a negative error code.
inline int build_model(Model* m, int P, const double* fixed, int nvars, const int* free_planet,
                       const int* free_elem, double hill_factor, int dims) {
    if (P < 1 || P > MAXP) return -2;
    if (nvars < 0 || nvars > P * NELEM) return -3;
    memset(m, 0, sizeof(*m));
    m->P = P;
    m->nvars = nvars;
    for (int i = 0; i < MAXP * NELEM; i++) m->src[i] = -1;
    for (int i = 0; i < P * NELEM; i++) m->fixed[i] = fixed[i];
    bool inclined = false;
    for (int v = 0; v < nvars; v++) {
        const int p = free_planet[v], e = free_elem[v];
        if (p < 0 || p >= P || e < 0 || e >= NELEM) return -4;
        if (m->src[p * NELEM + e] >= 0) return -5;  // duplicate slot
        m->src[p * NELEM + e] = v;
        m->free_planet[v] = p;
        m->free_elem[v] = e;
        if (e == EL_IX || e == EL_IY) inclined = true;
    }
    for (int p = 0; p < P; p++)
        if (fixed[p * NELEM + EL_IX] != 0.0 || fixed[p * NELEM + EL_IY] != 0.0) inclined = true;
    if (dims == 0) dims = inclined ? 3 : 2;
    if (dims != 2 && dims != 3) return -6;
    if (dims == 2 && inclined) return -7;
    m->D = dims;
    m->hill_factor = hill_factor;
    m->dt0 = 1e-3;       // rebound.Simulation() defaults (state.py:37)
    m->epsilon = 1e-9;
    m->m_star = 1.0;     // state.py:38
    m->max_attempts = 1 << 20;
    m->check_prior = 1;
    m->monotone_backward = 0;
    m->dense_output = 0;
    m->integrator = 0;
    return 0;
}

}  // namespace rv
