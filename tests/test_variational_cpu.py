"""Variational path on the CPU: the oracle's State.get_logp_d_dd against finite differences of the KAT-pinned
likelihood (the only pin available for derivatives), and the sequential emulation of the CUDA block
algorithm (rv_var.cuh) against the oracle."""
import numpy as np
import pytest

import rvtest as T

Z2 = np.zeros((2, 7))
# SURVEY App. B.9 probe values at the KAT-2 point (secondary: no reference output pins them)
B9_GRAD = [307.60027893, 1.0868579045, -0.35879657906, 731.41573369, -0.0046994336853, 101.53628861,
           -3.1126508437, 2.2590588034, -1813.9663833, -2.9139319471]
B9_HDIAG = [-3.9431084423e5, -44.078831086, -153.33342758, -8.3249339369e7, -69.404149840, -1.9274820008e4,
            -20.179175988, -69.102545988, -5.6647101946e7, -29.002875034]


def _hd():
    return T.load_vels("HD155358.vels")


@pytest.fixture(params=[0, 2], ids=["thread-per-set-planet", "lane-per-set"])
def var_layout(request):
    """Both CTA algorithms run through the same emulation tests: rv_var.cuh (one thread per (set, planet); three-planet
    systems) and rv_var2.cuh (one lane per set, producer warp; one- and two-planet systems)."""
    T.mirror().mirror_set_var_layout(request.param)
    yield request.param
    T.mirror().mirror_set_var_layout(0)


def _var_logp(obs, theta):
    """logp on the variational path's step sequence (forward + monotone backward), value only."""
    lo, _, _, so, _ = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, np.atleast_2d(theta))
    assert (so == 0).all()
    return lo


def test_oracle_derivatives_match_probe_values_and_kat2():
    obs = _hd()
    lo, go, ho, so, cnt = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, np.array([T.HD_SOL]))
    assert so[0] == 0
    assert abs(lo[0] - T.KAT2_LOGP) < 5e-11          # same 12 digits as the plain path (KAT-2)
    assert np.allclose(go[0], B9_GRAD, rtol=2e-9)
    assert np.allclose(np.diag(ho[0]), B9_HDIAG, rtol=2e-9)
    assert np.allclose(ho[0], ho[0].T)
    assert abs(ho[0][0][1] - 607.97975395) < 1e-6 and abs(ho[0][3][8] / 3.9365822418e6 - 1) < 2e-9


def test_oracle_gradient_and_hessian_match_finite_differences():
    obs = _hd()
    th0 = np.array(T.HD_SOL)
    _, g0, h0, _, _ = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, th0[None, :])
    g0, h0 = g0[0], h0[0]
    scale = np.array(T.HD_SCALE_VEC)
    # gradient: 4th-order central differences of the likelihood
    for i in range(10):
        h = 2e-2 * scale[i]
        pts = np.array([th0 + k * h * np.eye(10)[i] for k in (-2, -1, 1, 2)])
        f = _var_logp(obs, pts)
        fd = (f[0] - 8 * f[1] + 8 * f[2] - f[3]) / (12 * h)
        assert abs(fd - g0[i]) <= 1e-6 * abs(g0[i]) + 2e-7 * np.abs(g0).max() * scale[i] / scale.max(), (i, fd, g0[i])
    # Hessian columns: central differences of the variational gradient
    for i in (0, 3, 9):
        h = 1e-2 * scale[i]
        pts = np.array([th0 + k * h * np.eye(10)[i] for k in (-2, -1, 1, 2)])
        _, g, _, so, _ = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, pts)
        assert (so == 0).all()
        fd = (g[0] - 8 * g[1] + 8 * g[2] - g[3]) / (12 * h)
        err = np.abs(fd - h0[:, i]) / (np.abs(h0[:, i]) + 1e-3 * np.abs(h0[:, i]).max())
        assert err.max() < 1e-6, (i, err)


def test_norm_choice_changes_derivatives_below_tolerance():
    obs = _hd()
    a = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, np.array([T.HD_SOL]), var_in_norm=0)
    b = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, np.array([T.HD_SOL]), var_in_norm=1)
    assert np.allclose(a[1], b[1], rtol=1e-8) and np.allclose(a[2], b[2], rtol=1e-8, atol=1e-8 * np.abs(a[2]).max())
    assert b[4][1] > a[4][1]      # the 2017-era norm takes more steps


def test_block_emulation_matches_oracle_hd155358(var_layout):
    obs = _hd()
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 3, 5)
    theta[0] = T.HD_SOL
    lo, go, ho, so, co = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, theta)
    lm, gm, hm, sm, cm = T.mirror_loglik_d_dd(Z2, T.FP10, T.FE10, 2.0, obs, theta)
    assert np.array_equal(so, sm) and (so == 0).all()
    assert co[1] == cm[1]                       # identical IAS15 step sequence
    assert np.abs(lm - lo).max() < 1e-10
    assert np.abs(gm / go - 1).max() < 1e-7
    assert (np.abs(hm - ho) / (np.abs(ho) + 1e-6 * np.abs(ho).max())).max() < 1e-7


def test_block_emulation_prior_and_encounter_status(var_layout):
    obs = _hd()
    theta = np.array([T.HD_SOL, T.KAT5[1][0], T.HD_SOL])
    theta[2][3] = 1e-6                              # m <= 5e-6: hard prior
    lo, go, ho, so, _ = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, theta)
    lm, gm, hm, sm, _ = T.mirror_loglik_d_dd(Z2, T.FP10, T.FE10, 2.0, obs, theta)
    assert list(so) == [0, 3, 1] and list(sm) == [0, 3, 1]
    assert np.isneginf(lm[1]) and np.isneginf(lm[2])


@pytest.mark.parametrize("planets,free", [
    ([{"a": 0.35, "m": 0.001965}], [("a",)]),
    ([{"a": 0.2275, "h": 0.0, "k": 0.0, "m": 0.001965}], [("a", "h", "k")]),
    ([{"m": 1e-3, "a": 0.3, "h": 0.02, "k": -0.03, "l": 0.4, "ix": 0.1, "iy": -0.05}], [("a", "ix", "m", "l", "iy")]),
    ([{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0}, {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1},
      {"m": 1e-3, "a": 0.59, "h": 0.0, "k": 0.03, "l": 0.7}], [("a", "m"), ("h", "l"), ("a", "k", "m")]),
    ([{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0, "ix": 0.05, "iy": 0.02},
      {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1, "ix": -0.03, "iy": 0.0}],
     [("a", "ix", "m", "l"), ("h", "k", "m", "iy")]),
    ([{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0}, {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1}],
     [("m",), ("m",)]),
])
def test_block_emulation_other_shapes(planets, free, var_layout):
    if var_layout == 2 and len(planets) > 2:
        pytest.skip("the lane-per-set layout carries one or two planets")
    E = T.elems_from_planets(planets)
    fp, fe, th = [], [], []
    for i, keys in enumerate(free):
        for k in keys:
            fp.append(i); fe.append(T.ELEMS.index(k)); th.append(planets[i][k])
    rng = np.random.RandomState(3)
    obs = T.Obs()
    obs.tf = np.concatenate([[0.0], np.sort(rng.uniform(0, 4.0, 12))])
    obs.tb = np.sort(rng.uniform(-4.0, 0, 12))
    obs.rvf = 1e-4 * rng.normal(size=13); obs.rvb = 1e-4 * rng.normal(size=12)
    obs.errorf = np.full(13, 2e-4); obs.errorb = np.full(12, 2e-4)
    obs.Npoints = 24
    theta = np.array([th])
    lo, go, ho, so, co = T.orc_logp_d_dd_batch(E, fp, fe, 1.0, obs, theta)
    lm, gm, hm, sm, cm = T.mirror_loglik_d_dd(E, fp, fe, 1.0, obs, theta)
    assert so[0] == 0 and sm[0] == 0 and abs(co[1] - cm[1]) <= 2     # step sequences may split on a borderline reject
    assert abs(lm[0] - lo[0]) < 1e-9 * max(1.0, abs(lo[0]))
    assert np.abs(gm - go).max() <= 1e-7 * np.abs(go).max()
    assert np.abs(hm - ho).max() <= 1e-7 * np.abs(ho).max()


def test_chunked_second_order_sets_equal_one_launch():
    """Models whose (set, planet) threads do not fit one CTA run their second-order pairs in several launches
    (launch_var_chunked): emulated here with a 64-thread cap on HD155358 (66 sets x 2 planets -> four launches); value and
    gradient identical, Hessian to rounding (the launches' predictor-corrector iteration counts may differ)."""
    obs = _hd()
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 2, 5)
    lib = T.mirror()
    lib.mirror_set_var_layout(0)
    l0, g0, h0, s0, c0 = T.mirror_loglik_d_dd(Z2, T.FP10, T.FE10, 2.0, obs, theta)
    lib.mirror_set_var_nt_cap(64)
    try:
        l1, g1, h1, s1, c1 = T.mirror_loglik_d_dd(Z2, T.FP10, T.FE10, 2.0, obs, theta)
    finally:
        lib.mirror_set_var_nt_cap(0)
    assert np.array_equal(s0, s1) and (s0 == 0).all()
    assert c1[1] == 4 * c0[1]                   # four launches, same step sequence in each
    assert np.array_equal(l0, l1) and np.array_equal(g0, g1)
    assert np.abs(h1 - h0).max() <= 1e-12 * np.abs(h0).max()


def test_four_planets_in_chunks_match_oracle():
    """Four planets, 20 free parameters (231 sets x 4 planets = 924 threads against 448 per CTA): three launches."""
    obs, fixed, fp, fe, center, sc = T.many_planet_problem(4, nper=6, tmax=6.0)
    theta = center[None, :]
    lo, go, ho, so, _ = T.orc_logp_d_dd_batch(fixed, fp, fe, 2.0, obs, theta)
    lib = T.mirror()
    lib.mirror_set_var_layout(0)
    lib.mirror_set_var_nt_cap(448)
    try:
        lm, gm, hm, sm, _ = T.mirror_loglik_d_dd(fixed, fp, fe, 2.0, obs, theta)
    finally:
        lib.mirror_set_var_nt_cap(0)
    assert so[0] == 0 and sm[0] == 0
    assert abs(lm[0] - lo[0]) < 1e-9
    assert np.abs(gm[0] - go[0]).max() <= 1e-8 * np.abs(go[0]).max()
    assert np.abs(hm[0] - ho[0]).max() <= 1e-8 * np.abs(ho[0]).max()
