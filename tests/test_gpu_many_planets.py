"""Four and five planets (the reference's schema is open in the number of planets, state.py:8-31): the plain likelihood, MH and
the stretch move against the oracle at the same parity bar as the BASELINE configs; the variational path refuses (error -30)."""
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


def test_variational_path_refuses_more_than_three_planets(ctx):
    from rvel_mcmc_b200 import _abi
    obs, fixed, fp, fe, hill, center, sc = PH.problem("four")
    oh, m = PH._handles(ctx, obs, fixed, fp, fe, hill)
    with pytest.raises(_abi.RvGpuError, match=r"\(-30\)"):
        m.loglik_d_dd(oh, center[None, :])
    m.close(); oh.close()
