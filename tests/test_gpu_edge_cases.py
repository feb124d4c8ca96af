"""Edge cases of every entry point through the C ABI: empty and single-walker batches, ragged / duplicate epochs, a leg
with no epochs, thinning and optional outputs, walkers that start outside the prior, invalid arguments (error codes, no
crash), maximum epoch count, large walker counts."""
import numpy as np
import pytest

import rvtest as T
from test_samplers_cpu import _small_problem

pytestmark = pytest.mark.gpu
Z2 = np.zeros((2, 7))


@pytest.fixture(scope="module")
def ctx():
    from rvel_mcmc_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def _oh(ctx, obs):
    from rvel_mcmc_b200 import _abi
    return _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)


def _m(ctx, fixed, fp, fe, hill):
    from rvel_mcmc_b200 import _abi
    return _abi.ModelHandle(ctx, fixed, fp, fe, hill)


def test_variational_empty_single_ragged_and_no_free_parameter(ctx):
    obs = T.load_vels("HD155358.vels")
    oh = _oh(ctx, obs)
    m = _m(ctx, Z2, T.FP10, T.FE10, 2.0)
    r = m.loglik_d_dd(oh, np.zeros((0, 10)))
    assert r[0].shape == (0,) and r[1].shape == (0, 10) and r[2].shape == (0, 10, 10)
    # ragged legs: one forward epoch only; a backward leg with duplicate epochs and t = 0
    o2 = T.Obs()
    o2.tf = obs.tf[:1]; o2.rvf = obs.rvf[:1]; o2.errorf = obs.errorf[:1]
    o2.tb = np.array([-3.0, -1.5, -1.5, 0.0]); o2.rvb = np.zeros(4); o2.errorb = np.full(4, 1e-4); o2.Npoints = 5
    lg, gg, hg, sg = m.loglik_d_dd(_oh(ctx, o2), np.array([T.HD_SOL]))
    lo, go, ho, so, _ = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, o2, np.array([T.HD_SOL]))
    assert sg[0] == so[0] == 0 and abs(lg[0] - lo[0]) < 1e-6 * abs(lo[0])
    assert np.abs(gg - go).max() <= 1e-6 * np.abs(go).max() and np.abs(hg - ho).max() <= 1e-6 * np.abs(ho).max()
    # a model with every element pinned has nothing to differentiate: value only
    E = T.elems_from_planets(T.planets_from_vec(T.HD_SOL))
    m0 = _m(ctx, E, [], [], 2.0)
    lg, gg, hg, sg = m0.loglik_d_dd(oh, np.zeros((3, 0)))
    assert (sg == 0).all() and np.abs(lg - T.KAT2_LOGP).max() < 5e-11 and gg.shape == (3, 0)


def test_obs_validation_and_limits(ctx):
    from rvel_mcmc_b200 import _abi
    with pytest.raises(_abi.RvGpuError):
        _abi.ObsHandle(ctx, np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), np.zeros(0), 1)     # no epochs at all
    with pytest.raises(ValueError):
        _abi.ObsHandle(ctx, np.zeros(3), np.zeros(2), np.zeros(3), np.zeros(0), np.zeros(0), np.zeros(0), 1)
    # many epochs (the observation arrays no longer fit in shared memory next to the walker state: read from HBM),
    # one planet, forward leg only (empty backward leg)
    n = 20000
    o = T.Obs()
    o.tf = np.linspace(0, 4.0, n); o.rvf = np.zeros(n); o.errorf = np.full(n, 1e-3)
    o.tb = np.zeros(0); o.rvb = np.zeros(0); o.errorb = np.zeros(0); o.Npoints = n
    E = T.elems_from_planets([{"a": 0.35, "m": 0.001965}])
    m = _m(ctx, E, [0], [1], 1.0)
    lg, sg = m.loglik(_oh(ctx, o), np.array([[0.35], [0.36]]))
    so, lo = T.orc_logp(E, 1.0, o)
    assert (sg == 0).all() and abs(lg[0] - lo) < 1e-9 * abs(lo)


def test_model_validation(ctx):
    from rvel_mcmc_b200 import _abi
    for bad in (dict(fp=[0, 0], fe=[1, 1]),             # duplicate slot
                dict(fp=[2], fe=[1]),                   # planet out of range
                dict(fp=[0], fe=[9])):                  # element out of range
        with pytest.raises(_abi.RvGpuError):
            _abi.ModelHandle(ctx, Z2, bad["fp"], bad["fe"], 1.0)
    with pytest.raises(_abi.RvGpuError):
        _abi.ModelHandle(ctx, np.zeros((_abi.MAX_PLANETS + 1, 7)), [], [], 1.0)       # more than RV_MAX_PLANETS planets
    m = _m(ctx, Z2, T.FP10, T.FE10, 2.0)
    with pytest.raises(_abi.RvGpuError):
        m.set_option("no_such_option", 1)
    with pytest.raises(_abi.RvGpuError):
        m.set_option("integrator", 7)


def test_samplers_options_and_degenerate_inputs(ctx):
    from rvel_mcmc_b200 import _abi
    obs, E, fp, fe, center = _small_problem()
    oh = _oh(ctx, obs)
    m = _m(ctx, E, fp, fe, 1.0)
    sc = [3e-4, 0.01, 0.01]
    # zero steps: state unchanged, logp evaluated
    r = m.mh_run(oh, np.tile(center, (4, 1)), sc, 5.0, 0)
    assert np.array_equal(r["theta"], np.tile(center, (4, 1))) and np.isfinite(r["logp"]).all() and r["chain"].shape[0] == 0
    # thinning keeps every thin-th state of the unthinned chain; no-chain mode agrees on the final state
    a = m.mh_run(oh, np.tile(center, (8, 1)), sc, 5.0, 60, seed=5, thin=1)
    b = m.mh_run(oh, np.tile(center, (8, 1)), sc, 5.0, 60, seed=5, thin=7)
    c = m.mh_run(oh, np.tile(center, (8, 1)), sc, 5.0, 60, seed=5, record_chain=False)
    assert np.array_equal(b["chain"], a["chain"][6::7][:8]) and np.array_equal(c["theta"], a["theta"]) and c["chain"] is None
    # restarting from a saved state with first_step continues the same chain
    h1 = m.mh_run(oh, np.tile(center, (8, 1)), sc, 5.0, 25, seed=5)
    h2 = m.mh_run(oh, h1["theta"], sc, 5.0, 35, seed=5, first_step=25, logp=h1["logp"])
    assert np.array_equal(h2["theta"], a["theta"])
    # walkers that start outside the hard prior never move and keep logp = -inf
    th = np.tile(center, (4, 1)); th[2, 0] = 0.01
    r = m.mh_run(oh, th, [0.0, 0.0, 0.0], 5.0, 10)
    assert np.isneginf(r["logp"][2]) and r["n_accept"][2] == 0
    # stretch: odd walker counts are refused (emcee asserts the same); two walkers are allowed (degenerate but legal here)
    with pytest.raises(_abi.RvGpuError):
        m.stretch_run(oh, np.tile(center, (5, 1)), 2)
    r = m.stretch_run(oh, T.gaussian_ball(center, sc, 2, 1, width=1.0), 5, seed=1)
    assert r["chain"].shape == (5, 2, 3)
    # SMALA refuses a model without free parameters and WHFast
    m0 = _m(ctx, E, [], [], 1.0)
    with pytest.raises(_abi.RvGpuError):
        m0.smala_run(oh, np.zeros((2, 0)), 1.2, 0.14, 3)
    with pytest.raises(_abi.RvGpuError):
        m.alsmala_run(oh, np.tile(center, (2, 1)), 1.2, 0.14, -1.0, 3)


def test_large_batch_properties(ctx):
    # 2^20 single-planet walkers: finite, deterministic, permutation-equivariant
    obs, E, fp, fe, center = _small_problem()
    oh = _oh(ctx, obs)
    m = _m(ctx, E, fp, fe, 1.0)
    W = 1 << 20
    theta = T.gaussian_ball(center, [3e-4, 0.01, 0.01], W, 3, width=1.0)
    l1, s1 = m.loglik(oh, theta)
    assert (s1 == 0).mean() > 0.99 and np.isfinite(l1[s1 == 0]).all()
    perm = np.random.RandomState(0).permutation(W)
    l2, s2 = m.loglik(oh, theta[perm])
    assert np.array_equal(l2, l1[perm]) and np.array_equal(s2, s1[perm])


def test_massive_planets_star_in_norm(ctx):
    from test_hostmirror import _massive_system
    E, fp, fe, obs, theta = _massive_system()
    oh = _oh(ctx, obs)
    m = _m(ctx, E, fp, fe, 0.0)
    lg, sg = m.loglik(oh, theta)
    lo, so, _ = T.orc_logp_batch(E, fp, fe, 0.0, obs, theta)
    assert np.array_equal(sg, so) and np.abs(lg - lo).max() < 1e-9 * np.abs(lo).max()
    lg2, gg, hg, sg2 = m.loglik_d_dd(oh, theta)
    lo2, go, ho, so2, _ = T.orc_logp_d_dd_batch(E, fp, fe, 0.0, obs, theta)
    assert (sg2 == 0).all() and np.abs(gg - go).max() < 1e-6 * np.abs(go).max() and np.abs(hg - ho).max() < 1e-6 * np.abs(ho).max()


def test_non_finite_parameters_are_reported_not_hung(ctx):
    obs = T.load_vels("HD155358.vels")
    oh = _oh(ctx, obs)
    m = _m(ctx, Z2, T.FP10, T.FE10, 2.0)
    theta = np.tile(np.array(T.HD_SOL), (6, 1))
    theta[0, 0] = np.nan; theta[1, 4] = np.inf; theta[2, 3] = np.nan; theta[3, 1] = np.nan; theta[4, 5] = -np.inf
    lg, sg = m.loglik(oh, theta)
    assert sg[5] == 0 and (sg[:5] != 0).all() and np.isneginf(lg[:5]).all()
    lg, gg, hg, sg = m.loglik_d_dd(oh, theta)
    assert sg[5] == 0 and (sg[:5] != 0).all() and np.isneginf(lg[:5]).all()
    m.set_option("integrator", 1); m.set_option("dt0", 0.1)
    lg, sg = m.loglik(oh, theta)
    assert sg[5] == 0 and (sg[:5] != 0).all()


def test_prior_test_is_per_call_not_sticky_on_shared_models():
    """State.get_logp_d_dd integrates outside the hard prior (state.py:290-294); that must not switch the prior test off
    for later callers of the same cached device model (the samplers always reject priorHard violations, mcmc.py:171)."""
    import os
    from rvel_mcmc_b200 import observations, state, _abi
    obs = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
    st = state.State(T.planets_from_vec(T.HD_SOL)); st.hillRadiusFactor = 2.
    ctx = _abi.default_context()
    m, oh = st._model(ctx), obs._handle(ctx)
    bad = np.array(T.HD_SOL); bad[3] = 4e-6                        # m <= 5e-6
    out = state.State(T.planets_from_vec(bad)); out.hillRadiusFactor = 2.
    assert out._model(ctx) is m                                    # same schema -> same cached device model
    lp, d, dd = out.get_logp_d_dd(obs)                             # no prior test on this path
    assert np.isfinite(lp) and np.isfinite(d).all()
    theta = np.array([T.HD_SOL, bad])
    lg, gg, hg, sg = m.loglik_d_dd(oh, theta)
    assert list(sg) == [0, 1] and np.isneginf(lg[1]) and (gg[1] == 0).all()
    r = m.smala_run(oh, theta, 0.025, 1.4, 3, seed=1)
    assert r["status"][1] == 1 and r["n_accept"][1] == 0           # a chain started outside the prior never moves
    assert np.array_equal(r["theta"][1], bad)
    lg2, _, _, sg2 = m.loglik_d_dd(oh, theta, check_prior=False)
    assert list(sg2) == [0, 0] and abs(lg2[1] - lp) < 1e-12


def test_observation_edits_reach_the_device():
    """The reference reads obs on every call; the device copy is keyed on content, so edits are never served stale."""
    import os
    from rvel_mcmc_b200 import observations, state, _abi
    obs = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
    st = state.State(T.planets_from_vec(T.HD_SOL)); st.hillRadiusFactor = 2.
    l0 = st.get_logp(obs)
    assert abs(l0 - T.KAT2_LOGP) < 5e-11
    h0 = obs._handle(_abi.default_context())
    assert obs._handle(_abi.default_context()) is h0               # unchanged data: same device copy
    obs.rvf[3] += 5e-4                                             # in-place edit, same length
    st.logp = None
    l1 = st.get_logp(obs)
    assert abs(l1 - l0) > 1e-6
    obs.rvf = obs.rvf.copy(); obs.rvf[3] -= 5e-4                   # replaced array
    st.logp = None
    assert abs(st.get_logp(obs) - l0) < 1e-12
    obs.Npoints = 50
    st.logp = None
    assert abs(st.get_logp(obs) - 2 * l0) < 1e-9
    # handles do not outlive their context
    c2 = _abi.Context(0)
    h2 = obs._handle(c2)
    c2.close()
    assert h2.h is None


def test_chain_walkers_records_a_prefix_of_the_ensemble():
    """Context option chain_walkers: chain rows hold walkers [0, n) only; everything else (final state, acceptance counts,
    the recorded values themselves) is unchanged."""
    from rvel_mcmc_b200 import _abi
    ctx = _abi.Context(0)
    try:
        obs = T.load_vels("HD155358.vels")
        oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
        m = _abi.ModelHandle(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
        theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 64, 9)
        full = m.stretch_run(oh, theta, 6, seed=3, thin=2)
        part = m.stretch_run(oh, theta, 6, seed=3, thin=2, chain_walkers=10)
        assert part["chain"].shape == (3, 10, 10) and part["chain_lnp"].shape == (3, 10)
        assert np.array_equal(part["chain"], full["chain"][:, :10]) and np.array_equal(part["chain_lnp"], full["chain_lnp"][:, :10])
        assert np.array_equal(part["theta"], full["theta"]) and np.array_equal(part["n_accept"], full["n_accept"])
        sc = np.array(T.HD_SCALE_VEC)
        fm = m.mh_run(oh, theta, sc, 0.1, 5, seed=4)
        pm = m.mh_run(oh, theta, sc, 0.1, 5, seed=4, chain_walkers=7)
        assert pm["chain"].shape == (5, 7, 10) and np.array_equal(pm["chain"], fm["chain"][:, :7])
        assert np.array_equal(pm["chain_logp"], fm["chain_logp"][:, :7]) and np.array_equal(pm["theta"], fm["theta"])
        fs = m.smala_run(oh, theta[:16], 0.025, 1.4, 3, seed=5)
        ps = m.smala_run(oh, theta[:16], 0.025, 1.4, 3, seed=5, chain_walkers=4)
        assert ps["chain"].shape == (3, 4, 10) and np.array_equal(ps["chain"], fs["chain"][:, :4])
        assert np.array_equal(ps["theta"], fs["theta"])
        # the option is per call: the next call without it records every walker again
        again = m.stretch_run(oh, theta, 2, seed=3)
        assert again["chain"].shape == (2, 64, 10)
    finally:
        ctx.close()
