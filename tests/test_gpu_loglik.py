"""GPU parity tests proper: the CUDA path, called through the C ABI (ctypes -> librvgpu.so), against the
CPU oracle and the reference's golden values.  Tolerances are north_star's: RVs within 1e-9 relative,
log-likelihood within 1e-6 absolute; statuses (prior / Encounter) must match exactly."""
import numpy as np
import pytest

import rvtest as T

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    from rvel_mcmc_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def _obs_handle(ctx, obs):
    from rvel_mcmc_b200 import _abi
    return _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)


def _model(ctx, fixed, fp, fe, hill, dims=0, mapping=0):
    from rvel_mcmc_b200 import _abi
    m = _abi.ModelHandle(ctx, fixed, fp, fe, hill, dims)
    if mapping:
        m.set_option("mapping", mapping)
    return m


@pytest.mark.parametrize("mapping", [0, 1])
def test_kat2_kat5_kat6_through_abi(ctx, mapping):
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0, mapping=mapping)
    theta = np.array([T.HD_SOL, T.KAT6_VEC] + [v for v, _ in T.KAT5])
    logp, st = m.loglik(oh, theta)
    assert list(st) == [0, 0, 3, 3, 3]
    assert abs(logp[0] - T.KAT2_LOGP) < 5e-11          # (Ex)HD155358.ipynb:149
    assert abs(logp[1] - T.KAT6_LOGP) < 5e-6
    assert np.all(np.isneginf(logp[2:]))


@pytest.mark.parametrize("fn,planets", [("rvcurve_ben_2-1.txt", T.KAT3_PLANETS), ("rvcurve_ben_3-1.txt", T.KAT4_PLANETS)])
def test_kat3_kat4_rv_curves_through_abi(ctx, fn, planets):
    obs = T.load_vels("TEST_2-1_COMPACT.vels")
    tg, rg = T.load_rvcurve(fn)
    times = np.linspace(obs.tb[0], obs.tf[-1], 1000)
    m = _model(ctx, T.elems_from_planets(planets), [], [], 1.0)
    rv, st = m.rv_curve(np.zeros((1, 0)), times)
    assert st[0] == 0
    assert np.abs(rv[0] - rg).max() / np.abs(rg).max() < 1e-9


@pytest.mark.parametrize("mapping", [0, 1])
def test_walker_ball_matches_oracle(ctx, mapping):
    # HD155358 shape (config 2): Gaussian ball of walkers around the published solution, both hill factors
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 512, 11)
    theta[5, 3] = 1e-6            # prior violation (m)
    theta[6, 0] = 0.01            # prior violation (a)
    theta[7, 1] = 0.9; theta[7, 2] = 0.6   # h^2+k^2 >= 1
    for hill in (1.0, 2.0):
        m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, hill, mapping=mapping)
        lg, sg = m.loglik(oh, theta)
        lo, so, _ = T.orc_logp_batch(np.zeros((2, 7)), T.FP10, T.FE10, hill, obs, theta)
        assert np.array_equal(sg, so)
        assert list(sg[5:8]) == [1, 1, 1]
        ok = so == 0
        assert ok.sum() > 100
        assert np.abs(lg[ok] - lo[ok]).max() < 1e-6
        assert np.all(np.isneginf(lg[~ok]))


def test_wide_ball_with_encounters_matches_oracle(ctx):
    # a wide ball provokes Encounter storms like the reference's emcee runs (HD155358.ipynb:123-149)
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 768, 5, width=10.0)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    lg, sg = m.loglik(oh, theta)
    lo, so, _ = T.orc_logp_batch(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, theta)
    assert (so == 3).sum() > 5 and (so == 0).sum() > 50
    mism = np.nonzero(sg != so)[0]
    # an encounter exactly at the threshold may flip on rounding; allow at most 1 in 768
    assert len(mism) <= 1, (mism, sg[mism], so[mism])
    ok = (so == 0) & (sg == 0)
    assert np.abs(lg[ok] - lo[ok]).max() < 1e-6 * np.maximum(1.0, np.abs(lo[ok])).max()


def test_results_do_not_depend_on_batch_composition(ctx):
    # dynamic work distribution must not change a single bit of any walker's result
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 300, 3, width=8.0)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    l1, s1 = m.loglik(oh, theta)
    perm = np.random.RandomState(0).permutation(300)
    l2, s2 = m.loglik(oh, theta[perm])
    assert np.array_equal(s1[perm], s2)
    assert np.array_equal(l1[perm], l2)
    l3, s3 = m.loglik(oh, theta[:7])
    assert np.array_equal(l3, l1[:7]) and np.array_equal(s3, s1[:7])


def test_one_three_planets_and_inclined(ctx):
    rng = np.random.RandomState(3)
    obs = T.Obs()
    obs.tf = np.append([0], np.sort(rng.uniform(0, 6.0, 20))); obs.tb = np.sort(rng.uniform(0, -6.0, 20))
    obs.rvf = 1e-4 * rng.normal(size=21); obs.rvb = 1e-4 * rng.normal(size=20)
    obs.errorf = np.full(21, 3e-4); obs.errorb = np.full(20, 3e-4); obs.Npoints = 40
    oh = _obs_handle(ctx, obs)
    for planets in ([{"a": 0.35, "m": 0.001965}],
                    [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0},
                     {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1},
                     {"m": 1.1e-3, "a": 0.59, "h": 0.01, "k": 0.03, "l": 0.4}],
                    [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0, "ix": 0.05, "iy": -0.02},
                     {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1, "ix": -0.03, "iy": 0.04}],
                    [{"m": 0.92e-3, "a": 0.2275, "h": -0.06, "k": 0.015, "l": -1.0, "ix": 0.05, "iy": -0.02},
                     {"m": 1.95e-3, "a": 0.3665, "h": 0.02, "k": 0.0, "l": 2.1, "ix": -0.03, "iy": 0.04},
                     {"m": 1.1e-3, "a": 0.59, "h": 0.01, "k": 0.03, "l": 0.4, "ix": 0.01, "iy": 0.01}]):
        E = T.elems_from_planets(planets)
        so, lo = T.orc_logp(E, 1.0, obs)
        m = _model(ctx, E, [], [], 1.0)
        lg, sg = m.loglik(oh, np.zeros((3, 0)))
        assert so == 0 and list(sg) == [0, 0, 0]
        assert np.abs(lg - lo).max() < 1e-9 * max(1.0, abs(lo))
        # D=3 engine on a coplanar system == D=2 engine
        if len(planets[0]) <= 5:
            m3 = _model(ctx, E, [], [], 1.0, dims=3)
            l3, s3 = m3.loglik(oh, np.zeros((1, 0)))
            assert abs(l3[0] - lg[0]) < 1e-10 * max(1.0, abs(lo))


def test_edge_cases(ctx):
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    lg, sg = m.loglik(oh, np.zeros((0, 10)))          # empty batch
    assert lg.shape == (0,) and sg.shape == (0,)
    lg, sg = m.loglik(oh, np.array([T.HD_SOL]))       # single walker
    assert sg[0] == 0
    # ragged legs: forward leg only one epoch (t=0), long backward leg
    o2 = T.Obs()
    o2.tf = obs.tf[:1]; o2.rvf = obs.rvf[:1]; o2.errorf = obs.errorf[:1]
    o2.tb = obs.tb; o2.rvb = obs.rvb; o2.errorb = obs.errorb; o2.Npoints = 100
    l2, s2 = m.loglik(_obs_handle(ctx, o2), np.array([T.HD_SOL]))
    so, lo = T.orc_logp(T.elems_from_planets(T.planets_from_vec(T.HD_SOL)), 2.0, o2)
    assert s2[0] == so == 0 and abs(l2[0] - lo) < 1e-9
    # duplicate epochs and an epoch equal to t=0 in the middle of a leg
    o3 = T.Obs()
    o3.tf = np.array([0.0, 0.5, 0.5, 1.25]); o3.rvf = np.zeros(4); o3.errorf = np.full(4, 1e-4)
    o3.tb = np.array([-2.0, -1.0, 0.0]); o3.rvb = np.zeros(3); o3.errorb = np.full(3, 1e-4); o3.Npoints = 7
    l3, s3 = m.loglik(_obs_handle(ctx, o3), np.array([T.HD_SOL]))
    so, lo = T.orc_logp(T.elems_from_planets(T.planets_from_vec(T.HD_SOL)), 2.0, o3)
    assert s3[0] == so == 0 and abs(l3[0] - lo) < 1e-6 * abs(lo)


def test_state_api_on_gpu():
    # the reference-facing classes: State.get_logp / get_rv / get_chi2 + Encounter exception
    import os
    from rvel_mcmc_b200 import observations, state, Encounter
    obs = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
    s = state.State(T.planets_from_vec(T.HD_SOL))
    s.hillRadiusFactor = 2.
    assert abs(s.get_logp(obs) - T.KAT2_LOGP) < 5e-11
    assert abs(-s.get_chi2(obs) - T.KAT2_LOGP) < 5e-11
    rv = s.get_rv(obs.tb)
    assert abs(rv[-1] - (-0.00041883056816320016)) < 1e-13       # tb[-1] = 0: star vx at t=0 (KAT-1), after 0 -> tb[0] -> 0
    bad = state.State(T.planets_from_vec(T.KAT5[1][0]))
    with pytest.raises(Encounter):
        bad.get_logp(obs)
    with pytest.raises(Encounter):
        bad.get_rv(obs.tf)


def test_million_walker_properties(ctx):
    # full-size batch (BASELINE configs[4] lower range): size-independent properties
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    base = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 4096, 21)
    theta = np.tile(base, (32, 1))                      # 131072 walkers, 32 copies of each vector
    lg, sg = m.loglik(oh, theta)
    lg = lg.reshape(32, 4096); sg = sg.reshape(32, 4096)
    assert np.all(sg == sg[0]) and np.all(lg == lg[0])   # bit-identical replicas
    lo, so, _ = T.orc_logp_batch(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, base[:64])
    assert np.array_equal(so, sg[0, :64])
    assert np.abs(lo - lg[0, :64])[so == 0].max() < 1e-6


def test_monotone_backward_option_matches_default(ctx):
    # model option: backward leg swept once from 0 to the most negative epoch (state.py:273 order) instead of
    # state.py:91's stored order -- same likelihood within the north_star tolerance, about 2/3 of the work
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 4096, 77)
    theta[0] = T.HD_SOL
    ctx.count_work(True); ctx.work_counters(reset=True)
    l0, s0 = m.loglik(oh, theta)
    c0 = ctx.work_counters(reset=True)
    m.set_option("monotone_backward", 1)
    l1, s1 = m.loglik(oh, theta)
    c1 = ctx.work_counters(reset=True)
    ctx.count_work(False)
    assert np.array_equal(s0, s1) and (s0 == 0).all()
    assert np.abs(l1 - l0).max() < 1e-8 and abs(l1[0] - T.KAT2_LOGP) < 5e-11
    assert c1[1] < 0.75 * c0[1]


def test_kat1_initial_conditions_on_gpu(ctx):
    # (Ex)HD155358.ipynb:84-95 -- the 16-digit barycentric particles rebound printed, from the device's setup code
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    parts, st = m.initial_conditions(np.array([T.HD_SOL]))
    com = parts[0]
    assert st[0] == 0 and com[0, 0] == 1.0
    np.testing.assert_allclose(com[0, [4, 5]], [-0.00041883056816320016, 0.00014019875566797076], rtol=4e-15, atol=0)
    np.testing.assert_allclose(com[0, [1, 2]], [-0.00015729068590102283, -0.00035766924337062825], rtol=4e-15, atol=0)
    np.testing.assert_allclose(com[1, [4, 5]], [1.345239134152752, -0.31103808924058135], rtol=0, atol=6e-16)
    np.testing.assert_allclose(com[1, [1, 2]], [-0.10055834617969424, -0.5753744383506549], rtol=0, atol=6e-16)
    np.testing.assert_allclose(com[2, [4, 5]], [-0.9277725732029665, 0.16229778378922738], rtol=0, atol=6e-16)
    np.testing.assert_allclose(com[2, [1, 2]], [0.2964757596787923, 1.0432799562638122], rtol=0, atol=6e-16)
    assert np.all(com[:, [3, 6]] == 0.0)
    # and the reference-facing view: State.setup_sim().particles[0].vx is the RV at t = 0
    from rvel_mcmc_b200 import state
    s = state.State(T.planets_from_vec(T.HD_SOL)); s.hillRadiusFactor = 2.
    sim = s.setup_sim()
    assert sim.N == 3 and abs(sim.particles[0].vx - (-0.00041883056816320016)) < 1e-18
    assert abs(sim.particles[0].vx - s.get_rv([0.0])[0]) < 1e-18
    assert abs(sim.exit_min_distance - 2. * 1.04404207 * (0.00083037971 / 3.) ** (1. / 3.)) < 1e-15


def test_rv_of_walker_ball_matches_oracle_to_1e9(ctx):
    # north_star tolerance for the observable itself: RVs within 1e-9 relative of rebound (here: of the KAT-pinned oracle),
    # both epoch orders the reference uses (obs.tf forward; obs.tb = first hop to the most negative epoch, then forward)
    obs = T.load_vels("HD155358.vels")
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 0.0)
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 48, 21, width=0.3)
    for times in (obs.tf, obs.tb):
        rv, st = m.rv_curve(theta, times)
        assert (st == 0).all()
        for w in range(len(theta)):
            so, ro = T.orc_rv(T.elems_from_planets(T.planets_from_vec(theta[w])), 0.0, times)
            assert so == 0
            assert np.abs(rv[w] - ro).max() <= 1e-9 * np.abs(ro).max(), w


def test_dense_output_option_matches_default(ctx):
    # model option: one continuous integration per leg, RVs from the step polynomial (not the default: the default keeps
    # rebound's exact-finish-time step sequence)
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 4096, 78)
    theta[0] = T.HD_SOL
    for k in range(3):
        theta[1 + k] = T.KAT5[k][0]
    ctx.count_work(True); ctx.work_counters(reset=True)
    l0, s0 = m.loglik(oh, theta)
    c0 = ctx.work_counters(reset=True)
    m.set_option("dense_output", 1)
    l1, s1 = m.loglik(oh, theta)
    c1 = ctx.work_counters(reset=True)
    ctx.count_work(False)
    assert np.array_equal(s0, s1) and list(s1[1:4]) == [3, 3, 3]
    ok = s0 == 0
    assert np.abs(l1[ok] - l0[ok]).max() < 1e-8 and abs(l1[0] - T.KAT2_LOGP) < 5e-11
    assert c1[1] < 0.7 * c0[1] and c1[0] < 0.7 * c0[0]
    for mapping in (1,):                      # thread-per-walker mapping keeps its own per-planet dense data
        m2 = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0, mapping=mapping)
        m2.set_option("dense_output", 1)
        l2, s2 = m2.loglik(oh, theta[:256])
        assert np.array_equal(s2, s1[:256]) and np.abs(l2[ok[:256]] - l1[:256][ok[:256]]).max() < 1e-8


def test_cost_ordered_schedule_changes_nothing_but_the_order(ctx):
    # batches of >= 4096 walkers are scheduled most-expensive-first (model option cost_order, default on): the order in which
    # the kernel takes its items must never show in the results -- wide ball with prior violations and Encounters, the
    # equilibrated posterior ensemble, three planets, and the samplers built on the same launch
    import os
    obs = T.load_vels("HD155358.vels")
    oh = _obs_handle(ctx, obs)
    m = _model(ctx, np.zeros((2, 7)), T.FP10, T.FE10, 2.0)
    ens = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "hd155358_equilibrated_ensemble.npy"))
    theta = np.vstack([T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 6000, 31, width=12.0), ens[:6000]])
    theta[17, 3] = 1e-6; theta[18, 0] = 0.01; theta[19, 1] = np.nan
    m.set_option("cost_order", 0)
    l0, s0 = m.loglik(oh, theta)
    m.set_option("cost_order", 1)
    l1, s1 = m.loglik(oh, theta)
    assert np.array_equal(s0, s1) and np.array_equal(l0, l1)
    assert (s0 == 3).sum() > 0 and (s0 == 1).sum() > 0 and (s0 == 0).sum() > 6000
    # the same walkers in a small batch (below the ordering threshold) and against the oracle
    l2, s2 = m.loglik(oh, theta[6000:6256])
    assert np.array_equal(l2, l1[6000:6256]) and np.array_equal(s2, s1[6000:6256])
    lo, so, _ = T.orc_logp_batch(np.zeros((2, 7)), T.FP10, T.FE10, 2.0, obs, theta[5990:6010])
    assert np.array_equal(so, s1[5990:6010])
    ok = so == 0
    assert np.abs(lo[ok] - l1[5990:6010][ok]).max() < 1e-6
    # MH chains: 4096 chains in one call (ordered) == the same chains in two shards of 2048 (not ordered)
    sc = np.array(T.HD_SCALE_VEC)
    r = m.mh_run(oh, ens[:4096], sc, 0.02, 3, seed=9, record_chain=False, record_accepts=True)
    ra = m.mh_run(oh, ens[:2048], sc, 0.02, 3, seed=9, record_chain=False, record_accepts=True)
    rb = m.mh_run(oh, ens[2048:4096], sc, 0.02, 3, seed=9, first_chain_id=2048, record_chain=False, record_accepts=True)
    assert np.array_equal(r["theta"], np.vstack([ra["theta"], rb["theta"]]))
    assert np.array_equal(r["accepted"], np.hstack([ra["accepted"], rb["accepted"]]))
