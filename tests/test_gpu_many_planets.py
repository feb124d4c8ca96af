"""Four and five planets (the reference's schema is open in the number of planets, state.py:8-31): the plain likelihood, MH and
the stretch move against the oracle at the same parity bar as the BASELINE configs; value + gradient + Hessian in chunks of
second-order sets."""
import numpy as np
import pytest

import rvtest as T
import parity_horizon as PH

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from rvel_mcmc_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def _abi_model(ctx, fixed, fp, fe, hill):
    from rvel_mcmc_b200 import _abi
    return _abi.ModelHandle(ctx, fixed, fp, fe, hill)


@pytest.mark.parametrize("name", ["four", "five"])
def test_wide_ball_matches_oracle(ctx, name):
    obs, fixed, fp, fe, hill, center, sc = PH.problem(name)
    oh, m = PH._handles(ctx, obs, fixed, fp, fe, hill)
    # tight ball (all walkers integrate) + wide ball (Encounters, prior violations, large chi^2)
    theta = np.vstack([T.gaussian_ball(center, sc, 192, 3, width=1e-3), T.gaussian_ball(center, sc, 320, 4, width=0.25)])
    theta[0] = center
    lg, sg = m.loglik(oh, theta)
    lo, so, _ = T.orc_logp_batch(fixed, fp, fe, hill, obs, theta)
    assert np.array_equal(sg, so)
    assert (so[:192] == 0).all() and (so == 3).sum() > 0 and (so == 1).sum() > 0
    ok = so == 0
    assert np.abs(lg[ok] - lo[ok]).max() < 1e-6 * np.maximum(1.0, np.abs(lo[ok])).max()
    assert np.abs(lg[:192] - lo[:192]).max() < 1e-6
    # batch composition does not matter (lane groups of 4 / 5 leave idle lanes at the end of a warp)
    l2, s2 = m.loglik(oh, theta[5:48])
    assert np.array_equal(l2, lg[5:48]) and np.array_equal(s2, sg[5:48])
    # RV curves (get_rv, state.py:61-73) on the same engine
    m0 = _abi_model(ctx, fixed, fp, fe, 0.0)
    for times in (obs.tf, obs.tb):
        rv_g, st_g = m0.rv_curve(theta[:6], times)
        assert (st_g == 0).all()
        for w in range(6):
            E = np.zeros_like(fixed)
            for q, (p, e) in enumerate(zip(fp, fe)):
                E[p, e] = theta[w, q]
            so1, ref = T.orc_rv(E, 0.0, times)
            assert so1 == 0
            assert np.abs(rv_g[w] - ref).max() <= 1e-9 * np.abs(ref).max(), w
    m0.close()
    m.close(); oh.close()


@pytest.mark.parametrize("name", ["four", "five"])
def test_samplers_identical_decisions(ctx, name):
    obs, fixed, fp, fe, hill, center, sc = PH.problem(name)
    r = PH.mh_horizon(ctx, name, 32, 200, sc, 1e-3, seed=11)
    assert r["first_divergent_step"] is None and r["mismatched_decisions"] == 0, r
    assert 0.05 < r["accept_rate"] < 0.95, r
    assert r["max_abs_theta_diff"] < 1e-9 and r["max_abs_logp_diff"] < 1e-6
    r = PH.stretch_horizon(ctx, name, 64, 40, seed=5)
    assert r["first_divergent_step"] is None and r["mismatched_decisions"] == 0, r


@pytest.mark.parametrize("name,nw", [("four", 3), ("five", 2)])
def test_value_gradient_hessian_match_oracle(ctx, name, nw):
    """State.get_logp_d_dd (state.py:290-294) for four / five planets, 20 / 25 free parameters: 231 / 351 variational sets run
    in chunks of second-order pairs (launch_var_chunked; three launches for four planets, six for five); value, gradient and
    Hessian against the oracle at the bar of the BASELINE configs (1e-6 relative), statuses equal, SMALA runs on the model."""
    obs, fixed, fp, fe, hill, center, sc = PH.problem(name)
    oh, m = PH._handles(ctx, obs, fixed, fp, fe, hill)
    theta = T.gaussian_ball(center, sc, nw, 3, width=1e-3)
    theta[0] = center
    lg, gg, hg, sg = m.loglik_d_dd(oh, theta)
    lo, go, ho, so, _ = T.orc_logp_d_dd_batch(fixed, fp, fe, hill, obs, theta)
    assert np.array_equal(sg, so) and (so == 0).all()
    assert np.abs(lg - lo).max() < 1e-6
    for w in range(nw):
        assert np.abs(gg[w] - go[w]).max() <= 1e-6 * np.abs(go[w]).max(), w
        assert np.abs(hg[w] - ho[w]).max() <= 1e-6 * np.abs(ho[w]).max(), w
    # the plain kernel agrees on the value
    lp, sp = m.loglik(oh, theta)
    assert np.abs(lp - lg).max() < 1e-8
    # an out-of-prior walker and an encounter are reported per walker, not as an error
    bad = theta[:2].copy()
    bad[0, list(zip(fp, fe)).index((0, T.ELEMS.index("m")))] = 1e-7
    lb, gb, hb, sb = m.loglik_d_dd(oh, bad)
    assert sb[0] == 1 and np.isneginf(lb[0]) and not gb[0].any() and sb[1] == 0
    m.close(); oh.close()


def test_four_planets_inclined_in_chunks(ctx):
    """Four planets, two of them inclined (D = 3 kernels, 320 threads per block): {a, m, l} of every planet and (ix, iy) of the
    first = 14 free parameters, 120 sets x 4 planets = 480 threads -> two launches."""
    obs, _, _, _, center, _ = T.many_planet_problem(4)
    planets = [dict(p) for p in T.FIVE_PLANETS[:4]]
    planets[0]["ix"], planets[0]["iy"] = 0.03, -0.02
    planets[2]["ix"], planets[2]["iy"] = -0.01, 0.04
    E = T.elems_from_planets(planets)
    fp, fe, th = [], [], []
    for i in range(4):
        for k in ("a", "m", "l") + (("ix", "iy") if i == 0 else ()):
            fp.append(i); fe.append(T.ELEMS.index(k)); th.append(planets[i][k])
    oh, m = PH._handles(ctx, obs, E, fp, fe, 2.0)
    rng = np.random.RandomState(8)
    theta = np.array([th]) * (1 + 1e-4 * rng.normal(size=(3, len(th))))
    lg, gg, hg, sg = m.loglik_d_dd(oh, theta)
    lo, go, ho, so, _ = T.orc_logp_d_dd_batch(E, fp, fe, 2.0, obs, theta)
    assert np.array_equal(sg, so) and (so == 0).all()
    assert np.abs(lg - lo).max() < 1e-6 * max(1.0, np.abs(lo).max())
    for w in range(3):
        assert np.abs(gg[w] - go[w]).max() <= 1e-6 * np.abs(go[w]).max(), w
        assert np.abs(hg[w] - ho[w]).max() <= 1e-6 * np.abs(ho[w]).max(), w
    m.close(); oh.close()
