"""GPU parity of the device SMALA sampler (rv_smala_run) against the numpy/scipy oracle of mcmc.py:126-187 driven by the
same counter-based random numbers: identical accept/reject decisions, states equal to rounding, posterior agreement
with MH (the reference's own cross-sampler validation)."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import rvtest as T
from test_samplers_cpu import _small_problem

sys.path.insert(0, os.path.join(T.ROOT, "oracle"))
import smala_oracle as S  # noqa: E402

pytestmark = pytest.mark.gpu
Z2 = np.zeros((2, 7))


@pytest.fixture(scope="module")
def ctx():
    from rvel_mcmc_b200 import _abi
    c = _abi.Context(0)
    yield c
    c.close()


def _handles(ctx, obs, fixed, fp, fe, hill):
    from rvel_mcmc_b200 import _abi
    oh = _abi.ObsHandle(ctx, obs.tf, obs.rvf, obs.errorf, obs.tb, obs.rvb, obs.errorb, obs.Npoints)
    return oh, _abi.ModelHandle(ctx, fixed, fp, fe, hill)


def _oracle_chains(obs, E, fp, fe, hill, theta0, eps, alpha, seed, nsteps, first_id=0):
    P = E.shape[0]

    def evaluate(theta):
        lo, go, ho, so, _ = T.orc_logp_d_dd_batch(E, fp, fe, hill, obs, np.atleast_2d(theta), nthreads=1)
        return int(so[0]), float(lo[0]), go[0], ho[0]

    def prior(theta):
        el = np.ascontiguousarray(E.copy().reshape(-1))
        for v in range(len(fp)):
            el[fp[v] * 7 + fe[v]] = theta[v]
        return bool(T.oracle().orc_prior_hard(P, T.vp(el)))

    def one(w):
        return S.smala_chain(T.oracle(), evaluate, prior, theta0[w], eps, alpha, seed, first_id + w, 0, nsteps)

    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one, range(len(theta0))))
    return (np.stack([r[0] for r in res], axis=1), np.stack([r[1] for r in res], axis=1),
            np.array([r[2] for r in res]))


def test_smala_small_problem_identical_decisions_10k(ctx):
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0)
    W, nsteps, eps, alpha = 8, 1300, 1.2, 0.14                # (Ex)Full Test + Usage Example.ipynb run_smala(.., 1.2, 0.14)
    theta0 = T.gaussian_ball(center, [3e-4, 0.01, 0.01], W, 4, width=0.3)
    r = m.smala_run(oh, theta0, eps, alpha, nsteps, seed=31, record_accepts=True)
    chain_o, acc_o, lp_o = _oracle_chains(obs, E, fp, fe, 1.0, theta0, eps, alpha, 31, nsteps)
    assert (r["status"] == 0).all()
    assert np.array_equal(r["accepted"], acc_o)               # 10 400 decisions
    assert 0.3 < acc_o.mean() < 0.95
    assert np.abs(r["chain"] - chain_o).max() < 1e-8
    assert np.abs(r["logp"] - lp_o).max() < 1e-6


def test_smala_hd155358_identical_decisions_and_sharding_invariance(ctx):
    obs = T.load_vels("HD155358.vels")
    oh, m = _handles(ctx, obs, Z2, T.FP10, T.FE10, 2.0)
    W, nsteps, eps, alpha = 8, 12, 0.025, 1.4                 # (Ex)HD155358.ipynb:640
    theta0 = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 9, width=1e-2)
    r = m.smala_run(oh, theta0, eps, alpha, nsteps, seed=5, record_accepts=True)
    chain_o, acc_o, lp_o = _oracle_chains(obs, Z2, T.FP10, T.FE10, 2.0, theta0, eps, alpha, 5, nsteps)
    assert np.array_equal(r["accepted"], acc_o)
    assert np.abs(r["chain"] - chain_o).max() < 1e-8
    # chains keyed by global id: the second half run alone gives the same chains
    r2 = m.smala_run(oh, theta0[4:], eps, alpha, nsteps, seed=5, first_chain_id=4, record_accepts=True)
    assert np.array_equal(r2["accepted"], r["accepted"][:, 4:]) and np.array_equal(r2["chain"], r["chain"][:, 4:])


def test_smala_rejects_prior_and_flags_bad_start(ctx):
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0)
    theta0 = np.tile(center, (4, 1))
    theta0[1, 0] = 0.01                                       # a <= 0.02: start state outside the hard prior
    r = m.smala_run(oh, theta0, 1.2, 0.14, 20, seed=3)
    assert r["status"][1] != 0 and r["n_accept"][1] == 0 and np.array_equal(r["theta"][1], theta0[1])
    assert (r["status"][[0, 2, 3]] == 0).all() and (r["n_accept"][[0, 2, 3]] > 0).all()


def test_smala_posterior_agrees_with_mh(ctx):
    from rvel_mcmc_b200.samplers import ess
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0)
    mh = m.mh_run(oh, np.tile(center, (256, 1)), [3e-4, 0.01, 0.01], 5.0, 1500, seed=1, thin=5)
    sm = m.smala_run(oh, np.tile(center, (256, 1)), 1.2, 0.14, 600, seed=2, thin=2)
    a = mh["chain"][100:].reshape(-1, 3); b = sm["chain"][50:].reshape(-1, 3)
    ne_a, _ = ess(mh["chain"][100:]); ne_b, tau_b = ess(sm["chain"][50:])
    assert tau_b < 4.0                                        # SMALA decorrelates in ~1 step here (AC 1/1/1 in the notebook)
    for i in range(3):
        mcse = np.sqrt(a[:, i].var() / max(ne_a, 10) + b[:, i].var() / max(ne_b, 10))
        assert abs(a[:, i].mean() - b[:, i].mean()) < 5 * mcse + 1e-12, i
        assert abs(a[:, i].std() / b[:, i].std() - 1.0) < 0.15


def test_alsmala_identical_decisions_and_schedule(ctx):
    # Alsmala + run_alsmala's schedule (mcmc.py:191-234, driver.py:171-200): full steps early, cheap MALA steps later
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0)
    W, nsteps, eps, alpha, bern_a = 8, 400, 1.2, 0.14, 3.0
    theta0 = T.gaussian_ball(center, [3e-4, 0.01, 0.01], W, 4, width=0.3)
    r = m.alsmala_run(oh, theta0, eps, alpha, bern_a, nsteps, seed=41, record_accepts=True)

    def evaluate(theta):
        lo, go, ho, so, _ = T.orc_logp_d_dd_batch(E, fp, fe, 1.0, obs, np.atleast_2d(theta), nthreads=1)
        return int(so[0]), float(lo[0]), go[0], ho[0]

    def evaluate_plain(theta):
        lo, so, _ = T.orc_logp_batch(E, fp, fe, 1.0, obs, np.atleast_2d(theta), nthreads=1)
        return int(so[0]), float(lo[0])

    def prior(theta):
        el = np.ascontiguousarray(E.copy().reshape(-1))
        for v in range(len(fp)):
            el[fp[v] * 7 + fe[v]] = theta[v]
        return bool(T.oracle().orc_prior_hard(1, T.vp(el)))

    def one(w):
        return S.alsmala_chain(T.oracle(), evaluate, evaluate_plain, prior, theta0[w], eps, alpha, bern_a, 0, 41, w, 0, nsteps)

    with ThreadPoolExecutor(8) as ex:
        res = list(ex.map(one, range(W)))
    acc_o = np.stack([x[1] for x in res], axis=1); chain_o = np.stack([x[0] for x in res], axis=1)
    assert np.array_equal(r["full_step"], res[0][2])
    assert 0.15 < r["full_step"].mean() < 0.6 and r["full_step"][:20].mean() > r["full_step"][-100:].mean()
    assert np.array_equal(r["accepted"], acc_o)
    assert np.abs(r["chain"] - chain_o).max() < 1e-8
    # same posterior as SMALA (reference's cross-sampler check), far fewer variational evaluations
    al = m.alsmala_run(oh, np.tile(center, (256, 1)), eps, alpha, bern_a, 600, seed=2, thin=2)
    sm = m.smala_run(oh, np.tile(center, (256, 1)), eps, alpha, 600, seed=3, thin=2)
    a = al["chain"][100:].reshape(-1, 3); b = sm["chain"][100:].reshape(-1, 3)
    for i in range(3):
        assert abs(a[:, i].mean() - b[:, i].mean()) < 0.1 * b[:, i].std()
        assert abs(a[:, i].std() / b[:, i].std() - 1.0) < 0.15
