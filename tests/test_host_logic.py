"""Host-side mirror of the reference API: parameter ordering, schema, observation parsing (no GPU)."""
import os

import numpy as np
import pytest

import rvtest as T
from rvel_mcmc_b200 import observations, state, mcmc, driver


def test_param_order_is_python2_dict_order():
    # {m,a,h,k,l} -> a,h,k,m,l ; labels as in (Ex)HD155358.ipynb
    s = state.State([{"m": 1e-3, "a": 1.0, "h": 0.1, "k": 0.2, "l": 0.3}, {"m": 2e-3, "a": 2.0, "h": 0.4, "k": 0.5, "l": 0.6}])
    assert s.get_keys() == ['$a_0$', '$h_0$', '$k_0$', '$m_0$', '$l_0$', '$a_1$', '$h_1$', '$k_1$', '$m_1$', '$l_1$']
    assert np.allclose(s.get_params(), [1.0, 0.1, 0.2, 1e-3, 0.3, 2.0, 0.4, 0.5, 2e-3, 0.6])
    assert s.Nvars == 10
    assert state.State([{"a": 0.35, "m": 0.001}]).get_rawkeys() == ["a", "m"]
    s7 = state.State([{"m": 1e-3, "a": 1., "h": 0., "k": 0., "l": 0., "ix": 0.1, "iy": 0.2}])
    assert s7.get_rawkeys() == ["a", "ix", "h", "k", "m", "l", "iy"]


def test_set_shift_get_roundtrip_and_cache_invalidation():
    s = state.State([{"m": 1e-3, "a": 1.0, "h": 0.1, "k": 0.2, "l": 0.3}])
    s.logp = -1.0
    v = np.array([1.1, 0.0, 0.1, 2e-3, 0.5])
    s.set_params(v)
    assert s.logp is None and np.allclose(s.get_params(), v)
    s.logp = -2.0
    s.shift_params(np.ones(5) * 0.01)
    assert s.logp is None and np.allclose(s.get_params(), v + 0.01)
    with pytest.raises(AttributeError):
        s.set_params([1.0])


def test_ignore_vars_and_ignore_params():
    planets = [{"m": 1e-3, "a": 1.0, "h": 0.1, "k": 0.2, "l": 0.3}]
    s = state.State([dict(planets[0])], ignore_vars=["m", "l"])
    assert s.get_rawkeys() == ["a", "h", "k"] and s.Nvars == 3
    s = state.State([dict(planets[0])], ignore_vars='m')          # substring semantics, state.py:26
    assert s.get_rawkeys() == ["a", "h", "k", "l"]
    s = state.State([dict(planets[0])], ignore_params=[["h", "k"]])
    assert s.get_rawkeys() == ["a", "m", "l"]
    fp, fe = s._free_slots()
    assert fp == [0, 0, 0] and fe == [1, 0, 4]
    fixed = s._fixed_matrix()
    assert fixed[0, 2] == 0.1 and fixed[0, 3] == 0.2


def test_prior_hard_and_deepcopy():
    s = state.State([{"m": 1e-3, "a": 1.0, "h": 0.1, "k": 0.2, "l": 0.3}])
    assert not s.priorHard()
    for k, v in (("a", 0.02), ("m", 5e-6), ("h", 0.99)):
        c = s.deepcopy(); c.planets[0][k] = v
        assert c.priorHard() == (k != "h" or 0.99 ** 2 + 0.2 ** 2 >= 1.0)
    bad = s.deepcopy(); bad.planets[0]["a"] = 0.01
    assert bad.get_logp(None) == -np.inf           # prior short-circuits before any engine call (state.py:104)
    s.hillRadiusFactor = 2.0
    assert s.deepcopy().hillRadiusFactor == 1.0    # as the reference (state.py:212)
    c = s.deepcopy(); c.planets[0]["a"] = 5.0
    assert s.planets[0]["a"] == 1.0


def test_observation_fromfile_matches_reference_layout():
    obs = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
    assert len(obs.tb) == 61 and len(obs.tf) == 61 and obs.Npoints == 100
    assert obs.tb[-1] == 0.0 and np.all(np.diff(obs.tb) > 0) and np.all(np.diff(obs.tf) > 0)
    assert abs(obs.tb[0] - (-32.49)) < 0.01 and abs(obs.tf[-1] - 31.54) < 0.01
    o2 = T.load_vels("HD155358.vels")
    for a in ("tf", "tb", "rvf", "rvb", "errorf", "errorb"):
        assert np.array_equal(getattr(obs, a), getattr(o2, a))
    assert np.array_equal(obs.t, np.concatenate((obs.tb, obs.tf)))
    obs2 = observations.Observation_FromFile(os.path.join(T.GOLDEN, "TEST_2-1_COMPACT.vels"), Npoints=100)
    assert len(obs2.tb) == 60 and len(obs2.tf) == 59


def test_stretch_sampler_on_gaussian_target():
    # the emcee-2.2.1 stretch move restated in mcmc.StretchSampler samples a known target correctly
    def lnp(X):
        return -0.5 * np.sum((np.atleast_2d(X) / np.array([1.0, 3.0])) ** 2, axis=1)
    smp = mcmc.StretchSampler(20, 2, lnp, seed=5)
    p = np.random.RandomState(1).normal(size=(20, 2))
    lp = None
    chain = []
    for _ in range(1500):
        p, lp, _ = smp.run_mcmc(p, 1, lnprob0=lp)
        chain.append(p.copy())
    c = np.concatenate(chain[300:])
    assert abs(c[:, 0].std() - 1.0) < 0.1 and abs(c[:, 1].std() - 3.0) < 0.3
    with pytest.raises(AssertionError):
        mcmc.StretchSampler(7, 2, lnp)
    with pytest.raises(AssertionError):
        mcmc.StretchSampler(2, 2, lnp)


def test_autocorrelation_helpers():
    rng = np.random.RandomState(0)
    x = np.zeros(4000)
    for i in range(1, 4000):
        x[i] = 0.9 * x[i - 1] + rng.normal()
    tau = driver.ac_time(x)
    assert 4 <= tau <= 10            # 0.9^k < 0.5 at k = 7
    assert abs(driver.auto_correlation(x)[0] - 1.0) < 1e-12


def test_vels_write_read_round_trip(tmp_path):
    # on-disk format either side of the path (observations.py:52-69, driver.py:215-222)
    import os
    from rvel_mcmc_b200 import observations
    o = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
    assert len(o.tb) == 61 and len(o.tf) == 61 and o.tb[-1] == 0.0 and o.Npoints == 100
    assert np.array_equal(o.t, np.concatenate((o.tb, o.tf))) and len(o.rv) == 122 and len(o.err) == 122
    f = str(tmp_path / "copy.vels")
    observations.write_vels(f, o)
    o2 = observations.Observation_FromFile(f, Npoints=100)
    for name in ("tf", "tb", "rvf", "rvb", "errorf", "errorb"):
        assert np.allclose(getattr(o, name), getattr(o2, name), rtol=1e-8, atol=1e-12), name


def test_driver_persistence_formats(tmp_path, monkeypatch):
    # driver.py:46-54, 429-448: '<name>_<md5>.npy' dumps, 'log<name>' lines, 'aux_<md5>' notes
    from rvel_mcmc_b200 import driver, state
    monkeypatch.chdir(tmp_path)
    st = state.State([{"a": 0.35, "m": 0.001965}], ignore_vars=["m"])
    h = driver._run_id(st, "label")
    chain = np.arange(12.0).reshape(4, 3)
    driver.save_data(chain, "chain", h)
    assert np.array_equal(driver.load_data("chain", h), chain)
    driver.writing_to_log(chain[:2], "_t", True)
    driver.writing_to_log("START", "_t", True)
    driver.writing_to_log(chain, "_t", False)
    lines = open("log_t").read().split("\n")
    assert lines[0].split() == ["0.0", "1.0", "2.0", "3.0", "4.0", "5.0"] and lines[1].strip() == "START" and len(lines) == 3
    driver.save_aux_mh(h, st, "label", 100, {"a": 3e-4}, 5)
    txt = open("aux_" + h.hexdigest()).read()
    assert txt.startswith("initial = [{") and "label, Niter, Scale, Stepsize = 'label', 100" in txt


def test_ac_times_and_efficacy_definitions():
    # driver.py:343-382 (AC time = first lag with normalised autocorrelation < 0.5) and :412-414 (efficacy)
    from datetime import datetime, timedelta
    from rvel_mcmc_b200 import driver
    rng = np.random.RandomState(0)
    x = np.zeros((4000, 2))
    for i in range(1, len(x)):
        x[i, 0] = 0.9 * x[i - 1, 0] + rng.normal()          # AC(lag) = 0.9^lag -> first lag below 0.5 is 7
        x[i, 1] = rng.normal()
    b = driver.McmcBundle(None, x, np.zeros(len(x)), [], None, len(x), None, trimmedchain=x[500:])
    act = driver.plot_ACTimes(b, (1, 1))
    assert 5 <= act[0] <= 9 and act[1] == 1 and b.mcmc_actimes is act
    be = driver.McmcBundle(None, np.concatenate([x, x]), None, [], None, 2 * len(x), None, is_emcee=True, Nwalkers=2)
    assert np.allclose(driver.ac_times(be), [driver.ac_time(x[:, 0]), 1.0])
    t0 = datetime(2020, 1, 1)
    stamps = [t0, t0 + timedelta(seconds=3), t0 + timedelta(seconds=13)]
    assert abs(driver.efficacy(1000, [2.0, 5.0], stamps) - 1000 / (10.0 * 5.0)) < 1e-12      # from the SECOND stamp


def test_fast_keyword_selects_dense_output_without_touching_the_callers_state():
    """mcmc.Mh / Ensemble take fast=True (not in the reference): the sampler's own State copy evaluates the plain likelihood
    with the dense-output option; the caller's State and the default behaviour are unchanged."""
    from rvel_mcmc_b200 import mcmc, state
    s = state.State([{"a": 0.2275, "h": 0., "k": 0., "m": 0.001965}], ignore_vars=["m"])
    obs = object()
    assert mcmc.Mh(s, obs).state.dense_output is False
    mh = mcmc.Mh(s, obs, fast=True)
    assert mh.state.dense_output is True and s.dense_output is False
    assert mh.generate_proposal().dense_output is True           # proposals inherit it (State.deepcopy)
    np.random.seed(1)
    ens = mcmc.Ensemble(s, obs, {'a': 3e-4, 'h': 0.01, 'k': 0.01}, nwalkers=8, fast=True)
    assert ens.state.dense_output is True and len(ens.states) == 8
