"""BASELINE configs[3] ("C4"): synthetic three-planet near-resonant system, 15 free parameters -- the same parity bar as
the two-planet headline problem: wide walker balls with Encounters and prior violations (status equal, logp 1e-6),
value / gradient / Hessian vs the oracle (1e-6 relative), and identical MH accept/reject decisions."""
import numpy as np
import pytest

import parity_horizon as PH
import rvtest as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from rvel_mcmc_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def test_c4_wide_ball_with_encounters_matches_oracle(ctx):
    obs, fixed, fp, fe, hill, center, sc = PH.problem("c4")
    oh, m = PH._handles(ctx, obs, fixed, fp, fe, hill)
    theta = np.concatenate([T.gaussian_ball(center, sc, 384, 3, width=0.1), T.gaussian_ball(center, sc, 384, 4, width=1e-3)])
    theta[0, 3] = 4e-6                      # priorHard: m <= 5e-6 (state.py:305)
    theta[1, 5] = 0.015                     # a <= 0.02
    theta[2, 11] = 0.8; theta[2, 12] = 0.7  # h^2 + k^2 >= 1
    lg, sg = m.loglik(oh, theta)
    lo, so, _ = T.orc_logp_batch(fixed, fp, fe, hill, obs, theta, nthreads=16)
    assert list(sg[:3]) == [1, 1, 1]
    assert (so == 3).sum() > 50 and (so == 0).sum() > 500
    mism = np.nonzero(sg != so)[0]
    assert len(mism) <= 1, (mism, sg[mism], so[mism])          # an encounter exactly at the threshold may flip on rounding
    ok = (so == 0) & (sg == 0)
    assert np.all(np.abs(lg[ok] - lo[ok]) < 1e-6 * np.maximum(1.0, np.abs(lo[ok])))
    near = ok & (np.arange(len(theta)) >= 384)
    assert np.abs(lg[near] - lo[near]).max() < 1e-6            # absolute, around the solution
    assert np.all(np.isneginf(lg[sg != 0]))
    # thread-per-walker mapping: same answers
    m.set_option("mapping", 1)
    l1, s1 = m.loglik(oh, theta)
    assert np.array_equal(s1, sg) and np.abs(l1[ok] - lg[ok]).max() < 1e-6 * np.abs(lg[ok]).max()


def test_c4_value_gradient_hessian_ball_matches_oracle(ctx):
    obs, fixed, fp, fe, hill, center, sc = PH.problem("c4")
    oh, m = PH._handles(ctx, obs, fixed, fp, fe, hill)
    theta = np.concatenate([T.gaussian_ball(center, sc, 24, 5, width=1e-3), T.gaussian_ball(center, sc, 24, 6, width=0.05)])
    theta[0, 3] = 4e-6
    lg, gg, hg, sg = m.loglik_d_dd(oh, theta)
    lo, go, ho, so, _ = T.orc_logp_d_dd_batch(fixed, fp, fe, hill, obs, theta, nthreads=16)
    assert sg[0] == 1 and (sg == so).mean() >= 0.95
    assert (so == 3).sum() >= 1
    ok = (so == 0) & (sg == 0)
    assert ok.sum() >= 30
    assert np.all(np.abs(lg[ok] - lo[ok]) < 1e-6 * np.maximum(1.0, np.abs(lo[ok])))
    for w in np.where(ok)[0]:
        assert np.abs(gg[w] - go[w]).max() <= 1e-6 * np.abs(go[w]).max(), w
        assert np.abs(hg[w] - ho[w]).max() <= 1e-6 * np.abs(ho[w]).max(), w
        assert np.array_equal(hg[w], hg[w].T)
    bad = sg != 0
    assert np.isneginf(lg[bad]).all() and (gg[bad] == 0).all() and (hg[bad] == 0).all()


def test_c4_mh_identical_decisions(ctx):
    # tuned proposal: scales = conditional widths 1/sqrt(-H_ii) at the truth, step 0.6 (acceptance ~0.3);
    # 32 chains x 1000 steps = 32 000 decisions
    obs, fixed, fp, fe, hill, center, sc = PH.problem("c4")
    oh, m = PH._handles(ctx, obs, fixed, fp, fe, hill)
    _, _, hg, sg = m.loglik_d_dd(oh, center[None, :])
    assert sg[0] == 0
    scales = 1.0 / np.sqrt(-np.diag(hg[0]))
    r = PH.mh_horizon(ctx, "c4", 32, 1000, scales, 0.6, seed=7)
    from test_gpu_parity_horizon import _record
    _record(r)
    assert r["first_divergent_step"] is None and r["mismatched_decisions"] == 0, r
    assert 0.15 < r["accept_rate"] < 0.5, r
    assert r["max_abs_theta_diff"] < 1e-9 and r["max_abs_logp_diff"] < 1e-6
    # the reference's own scales (mcmc_benchmark_mh.py:52-53: step 1e-2 x {m 1e-3, a 0.3, h .5, k .5, l pi/2}) from the truth:
    # Encounters and huge chi^2 -- decisions still identical
    r = PH.mh_horizon(ctx, "c4", 64, 50, sc, 1e-2, seed=3)
    assert r["first_divergent_step"] is None, r
