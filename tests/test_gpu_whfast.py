"""GPU parity of the optional WHFast variant (model option integrator = 1) against its CPU oracle, through the C ABI.
PARITY UNPINNED with respect to the reference (no WHFast call site)."""
import numpy as np
import pytest

import rvtest as T

pytestmark = pytest.mark.gpu
Z2 = np.zeros((2, 7))


@pytest.fixture(scope="module")
def ctx():
    from rvel_mcmc_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def _handles(ctx, obs, fixed, fp, fe, hill, dt):
    from rvel_mcmc_b200 import _abi
    oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints) if obs is not None else None
    m = _abi.ModelHandle(ctx, fixed, fp, fe, hill)
    m.set_option("dt0", dt)
    m.set_option("integrator", 1)
    return oh, m


def test_whfast_loglik_matches_oracle_and_ias15(ctx):
    obs = T.load_vels("HD155358.vels")
    dt = 2 * np.pi * 0.65773033 ** 1.5 / 20
    oh, m = _handles(ctx, obs, Z2, T.FP10, T.FE10, 2.0, dt)
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 2048, 4)
    theta[0] = T.HD_SOL; theta[1] = T.KAT5[1][0]; theta[2] = T.HD_SOL; theta[2][3] = 1e-6
    lg, sg = m.loglik(oh, theta)
    lo, so, _ = T.orc_whfast_logp_batch(Z2, T.FP10, T.FE10, 2.0, dt, obs, theta, nthreads=16)
    assert np.array_equal(sg, so) and sg[1] == 3 and sg[2] == 1
    ok = so == 0
    assert np.abs(lg[ok] - lo[ok]).max() < 1e-8
    m.set_option("integrator", 0)
    li, si = m.loglik(oh, theta)
    assert np.abs(lg[ok] - li[ok]).max() < 0.05 * np.abs(li[ok]).max()      # O(dt^2) from the KAT-pinned IAS15 value
    with pytest.raises(Exception):
        m.set_option("integrator", 1)
        m.loglik_d_dd(oh, theta[:1])


def test_whfast_rv_curve_and_mh_chain(ctx):
    from test_samplers_cpu import _small_problem
    E3 = T.elems_from_planets(T.KAT3_PLANETS)
    times = np.array([-3.0, -1.0, 0.5, 0.5, 4.0, 2.0])
    _, m = _handles(ctx, None, E3, [], [], 1.0, 0.01)
    rv, st = m.rv_curve(np.zeros((1, 0)), times)
    so, rv_o, _ = T.orc_whfast_rv(E3, 1.0, 0.01, times)
    assert st[0] == 0 and so == 0 and np.abs(rv[0] - rv_o).max() < 1e-12
    # the samplers run on whichever integrator the model selects
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0, 0.02)
    r = m.mh_run(oh, np.tile(center, (64, 1)), [3e-4, 0.01, 0.01], 5.0, 200, seed=3, thin=4)
    assert 0.1 < r["n_accept"].mean() / 200 < 0.9
    assert abs(r["chain"][10:, :, 0].mean() - 0.35) < 2e-3


def test_state_api_whfast():
    import os
    from rvel_mcmc_b200 import observations, state
    obs = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
    s = state.State(T.planets_from_vec(T.HD_SOL)); s.hillRadiusFactor = 2.
    li = s.get_logp(obs)
    w = s.deepcopy(); w.hillRadiusFactor = 2.; w.integrator = "whfast"; w.dt = 0.05
    lw = w.get_logp(obs)
    assert abs(li - T.KAT2_LOGP) < 5e-11 and abs(lw - li) < 0.02 * abs(li) and lw != li
    assert w.deepcopy().integrator == "whfast"
