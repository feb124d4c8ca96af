"""The kernel's per-walker engine (rv_core.cuh / rv_loglik.cuh) compiled for the host (test-only mirror,
thread-per-walker mapping) against the independent CPU oracle.  Catches logic errors in the device source
without a GPU; the GPU parity tests proper are in test_gpu_loglik.py."""
import numpy as np

import rvtest as T


def _hd(nw, seed):
    obs = T.load_vels("HD155358.vels")
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, nw, seed)
    return obs, theta


def test_mirror_kat2_and_counts():
    obs = T.load_vels("HD155358.vels")
    logp, st, cnt = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, np.array([T.HD_SOL]))
    assert st[0] == 0
    assert abs(logp[0] - T.KAT2_LOGP) < 5e-11
    assert cnt[1] == 1913           # same step-attempt sequence as the oracle / SURVEY B.6
    # coplanar specialisation == general 3-D path
    logp3, st3, _ = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, np.array([T.HD_SOL]), dims=3)
    assert abs(logp3[0] - logp[0]) < 1e-11


def test_mirror_matches_oracle_on_walker_ball():
    obs, theta = _hd(24, 7)
    lo, so, _ = T.orc_logp_batch(np.zeros((2, 7)), T.FP10, T.FE10, 1.0, obs, theta)
    lm, sm, _ = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 1.0, obs, theta)
    assert np.array_equal(so, sm)
    ok = so == 0
    assert ok.sum() > 0
    assert np.abs(lo[ok] - lm[ok]).max() < 1e-6          # north_star: logp within 1e-6 absolute
    assert np.all(np.isneginf(lm[~ok]))


def test_mirror_encounters_and_prior():
    obs = T.load_vels("HD155358.vels")
    vecs = np.array([v for v, _ in T.KAT5] + [T.HD_SOL])
    vecs[-1, 3] = 1e-6                 # m <= 5e-6 -> prior
    lm, sm, _ = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, vecs)
    assert list(sm) == [3, 3, 3, 1]
    assert np.all(np.isneginf(lm))


def test_mirror_rv_curve_kat3():
    obs = T.load_vels("TEST_2-1_COMPACT.vels")
    tg, rg = T.load_rvcurve("rvcurve_ben_2-1.txt")
    E = T.elems_from_planets(T.KAT3_PLANETS)
    times = np.linspace(obs.tb[0], obs.tf[-1], 1000)
    rv, st = T.mirror_loglik(E, [], [], 1.0, obs, np.zeros((1, 0)), times=times)
    assert st[0] == 0
    assert np.abs(rv[0] - rg).max() / np.abs(rg).max() < 1e-9     # north_star: RVs within 1e-9 relative


def test_mirror_one_and_three_planets():
    rng = np.random.RandomState(3)
    obs = T.Obs()
    obs.tf = np.append([0], np.sort(rng.uniform(0, 6.0, 20))); obs.tb = np.sort(rng.uniform(0, -6.0, 20))
    obs.rvf = 1e-4 * rng.normal(size=21); obs.rvb = 1e-4 * rng.normal(size=20)
    obs.errorf = np.full(21, 3e-4); obs.errorb = np.full(20, 3e-4); obs.Npoints = 40
    for planets in ([{"a": 0.35, "m": 0.001965}],
                    [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0},
                     {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1},
                     {"m": 1.1e-3, "a": 0.59, "h": 0.01, "k": 0.03, "l": 0.4}],
                    [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0, "ix": 0.05, "iy": -0.02},
                     {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1, "ix": -0.03, "iy": 0.04}]):
        E = T.elems_from_planets(planets)
        so, lo = T.orc_logp(E, 1.0, obs)
        lm, sm, _ = T.mirror_loglik(E, [], [], 1.0, obs, np.zeros((1, 0)))
        assert so == sm[0] == 0
        assert abs(lo - lm[0]) < 1e-9 * max(1.0, abs(lo))


def test_mirror_monotone_backward_option():
    # state.py:91 order (0 -> tb[0] -> forward hops) vs one monotone sweep (state.py:273 order): same logp, fewer steps
    obs, theta = _hd(6, 9)
    l0, s0, c0 = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, theta)
    T.mirror().mirror_set_monotone(1)
    try:
        l1, s1, c1 = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, theta)
    finally:
        T.mirror().mirror_set_monotone(0)
    assert np.array_equal(s0, s1) and (s0 == 0).all()
    assert np.abs(l1 - l0).max() < 1e-9
    assert c1[1] < 0.75 * c0[1]           # 1244 -> 637 backward steps


def _massive_system():
    # planets as heavy as the star: sum(m_p)/m_star >= 1, so the STAR can set the IAS15 error norms (rebound takes the
    # maxima over all particles; the kernels derive the star from the planets and switch its terms on for this case)
    planets = [{"m": 0.6, "a": 1.0, "h": 0.05, "k": 0.0, "l": 0.3}, {"m": 0.7, "a": 4.0, "h": 0.0, "k": 0.1, "l": 2.0}]
    E = T.elems_from_planets(planets)
    rng = np.random.RandomState(1)
    obs = T.Obs()
    obs.tf = np.concatenate([[0.0], np.sort(rng.uniform(0, 20, 15))]); obs.tb = np.sort(rng.uniform(-20, 0, 15))
    obs.rvf = 0.1 * rng.normal(size=16); obs.rvb = 0.1 * rng.normal(size=15)
    obs.errorf = np.full(16, 0.05); obs.errorb = np.full(15, 0.05); obs.Npoints = 30
    theta = np.array([[1.0, 4.0], [1.02, 3.9], [0.97, 4.2]])
    return E, [0, 1], [1, 1], obs, theta


def test_mirror_massive_planets_star_in_norm():
    E, fp, fe, obs, theta = _massive_system()
    lo, so, co = T.orc_logp_batch(E, fp, fe, 0.0, obs, theta)
    lm, sm, cm = T.mirror_loglik(E, fp, fe, 0.0, obs, theta)
    assert np.array_equal(so, sm) and (so == 0).all()
    assert np.abs(lm - lo).max() < 1e-9 * np.abs(lo).max()
    assert abs(co[1] - cm[1]) <= 3
    lo2, go, ho, so2, _ = T.orc_logp_d_dd_batch(E, fp, fe, 0.0, obs, theta)
    lm2, gm, hm, sm2, _ = T.mirror_loglik_d_dd(E, fp, fe, 0.0, obs, theta)
    assert (sm2 == 0).all() and np.abs(lm2 - lo2).max() < 1e-9 * np.abs(lo2).max()
    assert np.abs(gm - go).max() < 1e-8 * np.abs(go).max() and np.abs(hm - ho).max() < 1e-8 * np.abs(ho).max()


def test_mirror_dense_output_option():
    # natural IAS15 steps + RVs read from the step's acceleration polynomial instead of a truncated step per epoch:
    # same likelihood (north_star: 1e-6 absolute), step count independent of the number of epochs
    obs, theta = _hd(6, 9)
    lo, so, _ = T.orc_logp_batch(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, theta)
    l0, s0, c0 = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, theta)
    T.mirror().mirror_set_dense(1)
    try:
        l1, s1, c1 = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, theta)
        rng = np.random.RandomState(2)
        o = T.Obs()
        o.tf = np.concatenate([[0.0], np.sort(rng.uniform(0, 30, 500))]); o.tb = np.sort(rng.uniform(-30, 0, 500))
        o.rvf = 1e-4 * rng.normal(size=501); o.rvb = 1e-4 * rng.normal(size=500)
        o.errorf = np.full(501, 2e-4); o.errorb = np.full(500, 2e-4); o.Npoints = 1000
        l2, s2, c2 = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, o, theta[:2])
        # encounter vectors still flag, whatever the epoch handling (KAT-5)
        l3, s3, _ = T.mirror_loglik(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, np.array([k[0] for k in T.KAT5]))
    finally:
        T.mirror().mirror_set_dense(0)
    lo2, so2, co2 = T.orc_logp_batch(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, o, theta[:2])
    assert np.array_equal(s1, so) and (s1 == 0).all() and np.abs(l1 - lo).max() < 1e-9
    assert c1[1] < 0.7 * c0[1]
    assert (s2 == 0).all() and np.abs(l2 - lo2).max() < 1e-9 * np.abs(lo2).max() and c2[1] < 0.6 * co2[1]
    assert list(s3) == [3, 3, 3]


FIVE_PLANETS = [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0},
                {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1},
                {"m": 1.1e-3, "a": 0.59, "h": 0.01, "k": 0.03, "l": 0.4},
                {"m": 0.6e-3, "a": 0.95, "h": -0.02, "k": 0.01, "l": 2.9},
                {"m": 0.3e-3, "a": 1.55, "h": 0.03, "k": -0.02, "l": -2.2}]


def test_mirror_four_and_five_planets():
    # the reference's schema is open in the number of planets (state.py:8-31); the plain engine is built for up to five
    rng = np.random.RandomState(5)
    obs = T.Obs()
    obs.tf = np.append([0], np.sort(rng.uniform(0, 5.0, 12))); obs.tb = np.sort(rng.uniform(0, -5.0, 12))
    obs.rvf = 1e-4 * rng.normal(size=13); obs.rvb = 1e-4 * rng.normal(size=12)
    obs.errorf = np.full(13, 3e-4); obs.errorb = np.full(12, 3e-4); obs.Npoints = 24
    for npl, incl in ((4, False), (5, False), (4, True), (5, True)):
        planets = [dict(p) for p in FIVE_PLANETS[:npl]]
        if incl:
            for i, p in enumerate(planets):
                p["ix"] = 0.02 * (i + 1); p["iy"] = -0.015 * (i - 1)
        E = T.elems_from_planets(planets)
        # free parameters: a, m of every planet, drawn in a small ball
        fp = [i for i in range(npl) for _ in range(2)]; fe = [1, 0] * npl
        center = np.array([[p["a"], p["m"]] for p in planets]).reshape(-1)
        theta = center[None, :] * (1.0 + 1e-3 * rng.normal(size=(3, 2 * npl)))
        lo, so, co = T.orc_logp_batch(E, fp, fe, 1.0, obs, theta)
        lm, sm, cm = T.mirror_loglik(E, fp, fe, 1.0, obs, theta)
        assert np.array_equal(so, sm) and (so == 0).all()
        assert np.abs(lm - lo).max() < 1e-9 * np.abs(lo).max()
        assert abs(co[1] - cm[1]) <= 3


def test_mirror_item_order_does_not_show_in_the_results():
    # LoglikArgs::order (the cost-ordered schedule of the device launch): items are taken in another order, results are stored
    # by walker -- identical output, prior violations and Encounters included
    obs, theta = _hd(5, 13)
    theta = np.vstack([theta, np.array([k[0] for k in T.KAT5])])
    theta[1, 3] = 1e-6
    E = np.zeros((2, 7))
    l0, s0, c0 = T.mirror_loglik(E, T.FP10, T.FE10, 2.0, obs, theta)
    T.mirror().mirror_set_reverse_order(1)
    try:
        l1, s1, c1 = T.mirror_loglik(E, T.FP10, T.FE10, 2.0, obs, theta)
    finally:
        T.mirror().mirror_set_reverse_order(0)
    assert np.array_equal(s0, s1) and np.array_equal(l0, l1) and list(c0) == list(c1)
    assert list(s0[[1, 5, 6, 7]]) == [1, 3, 3, 3]
