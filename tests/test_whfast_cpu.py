"""Optional WHFast variant on the CPU.  PARITY UNPINNED (the reference never selects WHFast): the checks are
the integrator's own invariants, agreement of the two independent implementations (oracle/rv_whfast.c and the device
engine rv_whfast.cuh compiled for the host), and convergence to the KAT-pinned IAS15 results as dt -> 0."""
import ctypes as C

import numpy as np

import rvtest as T

Z2 = np.zeros((2, 7))


def test_kepler_solver_invariants():
    o = T.oracle()
    rng = np.random.RandomState(1)
    for _ in range(50):
        x = rng.normal(size=3); v = 0.7 * rng.normal(size=3)
        M = 1.0 + rng.uniform(0, 0.01)
        if 0.5 * v @ v - M / np.linalg.norm(x) > -0.05:      # keep bound orbits
            continue
        x0, v0 = x.copy(), v.copy()
        dt = rng.uniform(0.01, 3.0)
        assert o.orc_kepler_step(C.c_double(M), C.c_double(dt), T.vp(x), T.vp(v)) == 0
        E = lambda a, b: 0.5 * b @ b - M / np.linalg.norm(a)
        assert abs(E(x, v) - E(x0, v0)) < 2e-12 * abs(E(x0, v0))
        assert np.abs(np.cross(x, v) - np.cross(x0, v0)).max() < 2e-12
        assert o.orc_kepler_step(C.c_double(M), C.c_double(-dt), T.vp(x), T.vp(v)) == 0
        assert np.abs(x - x0).max() < 1e-11 and np.abs(v - v0).max() < 1e-11


def test_whfast_converges_to_ias15_rv_curve():
    E3 = T.elems_from_planets(T.KAT3_PLANETS)                       # the KAT-3 system (mcmc_benchmark_smala.py:32)
    tt = np.linspace(0, 20, 50)
    st, r_ias = T.orc_rv(E3, 0.0, tt)
    errs = []
    for dt in (0.02, 0.01, 0.005):
        st, rv, n = T.orc_whfast_rv(E3, 0.0, dt, tt)
        assert st == 0
        errs.append(np.abs(rv - r_ias).max() / np.abs(r_ias).max())
    assert errs[0] < 1e-4 and errs[2] < errs[1] < errs[0] and errs[2] < 0.2 * errs[0]


def test_device_engine_matches_oracle_whfast():
    obs = T.load_vels("HD155358.vels")
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 16, 2)
    theta[0] = T.HD_SOL
    dt = 2 * np.pi * 0.65773033 ** 1.5 / 20                           # P_inner / 20 (BASELINE configs[4])
    lo, so, co = T.orc_whfast_logp_batch(Z2, T.FP10, T.FE10, 2.0, dt, obs, theta)
    lm, sm, cm = T.mirror_whfast(Z2, T.FP10, T.FE10, 2.0, dt, obs, theta)
    assert np.array_equal(so, sm) and (so == 0).all()
    assert co[1] == cm[1]                                            # same number of steps
    assert np.abs(lm - lo).max() < 1e-9
    # and it approximates the KAT-2-pinned IAS15 likelihood at O(dt^2)
    li, si, _ = T.orc_logp_batch(Z2, T.FP10, T.FE10, 2.0, obs, theta)
    assert np.abs(lm - li).max() < 0.05 * np.abs(li).max()
    lo2, _, _ = T.orc_whfast_logp_batch(Z2, T.FP10, T.FE10, 2.0, dt / 4, obs, theta)
    assert np.abs(lo2 - li).max() < 0.3 * np.abs(lo - li).max()


def test_device_engine_statuses_and_other_shapes():
    obs = T.load_vels("HD155358.vels")
    dt = 0.05
    theta = np.array([T.HD_SOL, T.KAT5[1][0], T.HD_SOL])
    theta[2][3] = 1e-6
    lo, so, _ = T.orc_whfast_logp_batch(Z2, T.FP10, T.FE10, 2.0, dt, obs, theta)
    lm, sm, _ = T.mirror_whfast(Z2, T.FP10, T.FE10, 2.0, dt, obs, theta)
    assert list(so) == [0, 3, 1] and list(sm) == [0, 3, 1]
    # three planets, inclined, RV curve mode visiting times in the given (non-monotone) order
    planets = [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0, "ix": 0.05, "iy": 0.02},
               {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1},
               {"m": 1.0e-3, "a": 0.59, "h": 0.0, "k": 0.03, "l": 0.7}]
    E = T.elems_from_planets(planets)
    times = np.array([-3.0, -1.0, 0.5, 0.5, 4.0, 2.0])
    st, rv_o, n = T.orc_whfast_rv(E, 1.0, 0.01, times)
    rv_m, sm, _ = T.mirror_whfast(E, [], [], 1.0, 0.01, None, np.zeros((1, 0)), times=times)
    assert st == 0 and sm[0] == 0
    assert np.abs(rv_m[0] - rv_o).max() < 1e-12


def test_device_engine_four_and_five_planets():
    # the WHFast engine instantiated for four / five planets (Jacobi chain of any length) against its oracle
    for npl in (4, 5):
        obs, fixed, fp, fe, center, sc = T.many_planet_problem(npl, nper=10, tmax=8.0)
        theta = T.gaussian_ball(center, sc, 3, 9, width=1e-3)
        dt = 2 * np.pi * 0.2275 ** 1.5 / 20
        lo, so, co = T.orc_whfast_logp_batch(fixed, fp, fe, 1.0, dt, obs, theta)
        lm, sm, cm = T.mirror_whfast(fixed, fp, fe, 1.0, dt, obs, theta)
        assert np.array_equal(so, sm) and (so == 0).all()
        assert co[1] == cm[1]
        assert np.abs(lm - lo).max() < 1e-9 * max(1.0, np.abs(lo).max())
