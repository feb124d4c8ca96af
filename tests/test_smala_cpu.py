"""SMALA on the CPU: the device sampler's per-chain arithmetic (rv_smala.cuh, compiled for the host) against the
numpy/scipy oracle of mcmc.py:126-187 (oracle/smala_oracle.py), and a short chain end to end."""
import ctypes as C
import os
import sys

import numpy as np

import rvtest as T
from test_samplers_cpu import _small_problem

sys.path.insert(0, os.path.join(T.ROOT, "oracle"))
import smala_oracle as S  # noqa: E402

Z2 = np.zeros((2, 7))


def mirror_propose(th, g, H, eps, alpha, seed, cid, step, cur_status=0):
    n = len(th)
    prop = np.zeros(n)
    qf = C.c_double()
    r = T.mirror().mirror_smala_propose(n, T.vp(np.ascontiguousarray(th)), T.vp(np.ascontiguousarray(g)),
                                        T.vp(np.ascontiguousarray(H)), cur_status, C.c_double(eps), C.c_double(alpha),
                                        C.c_ulonglong(seed), C.c_ulonglong(cid), C.c_uint(step), T.vp(prop), C.byref(qf))
    return r, prop, qf.value


def mirror_accept(th, logp, prop, p_logp, p_grad, p_hess, p_status, geo, qf, eps, alpha, seed, cid, step):
    flag = C.c_int(0)
    a = T.mirror().mirror_smala_accept(len(th), T.vp(np.ascontiguousarray(th)), C.c_double(logp),
                                       T.vp(np.ascontiguousarray(prop)), C.c_double(p_logp),
                                       T.vp(np.ascontiguousarray(p_grad)), T.vp(np.ascontiguousarray(p_hess)), p_status, geo,
                                       C.c_double(qf), C.c_double(eps), C.c_double(alpha), C.c_ulonglong(seed),
                                       C.c_ulonglong(cid), C.c_uint(step), C.byref(flag))
    return a, flag.value


def test_softabs_proposal_and_density_match_numpy_at_hd155358():
    obs = T.load_vels("HD155358.vels")
    th = np.array(T.HD_SOL)
    lo, go, ho, so, _ = T.orc_logp_d_dd_batch(Z2, T.FP10, T.FE10, 2.0, obs, th[None, :])
    eps, alpha = 0.025, 1.4                                   # (Ex)HD155358.ipynb:640
    mu, Ginv = S.proposal_mean_cov(th, go[0], ho[0], eps, alpha)
    L = np.linalg.cholesky(Ginv)
    for step in range(5):
        new = mu + eps * L @ S.normals(T.oracle(), 7, 3, step, 10)
        q = S.stats.multivariate_normal.logpdf(new, mean=mu, cov=eps ** 2 * Ginv)
        r, prop, qf = mirror_propose(th, go[0], ho[0], eps, alpha, 7, 3, step)
        assert r == 0
        assert np.abs((prop - new) / (new - th)).max() < 1e-8
        assert abs(qf - q) < 1e-8


def test_not_spd_is_flagged_not_fatal():
    n = 3
    th = np.array([0.35, 0.02, 0.01])
    H = np.diag([-1e3, 0.0, -5.0])                 # a zero eigenvalue: lam/tanh(alpha lam) = nan (mcmc.py:137)
    r, prop, qf = mirror_propose(th, np.ones(n), H, 1.2, 0.14, 1, 0, 0)
    assert r == 9 and np.array_equal(prop, th)
    r, prop, qf = mirror_propose(th, np.ones(n), np.diag([-1e3, np.nan, -5.0]), 1.2, 0.14, 1, 0, 0)
    assert r == 9
    a, flag = mirror_accept(th, -1.0, th, -1.0, np.ones(n), H, 0, 0, 0.0, 1.2, 0.14, 1, 0, 0)
    assert a == 0 and flag == 9
    r, prop, qf = mirror_propose(th, np.ones(n), np.diag([-1e3, -2.0, -5.0]), 1.2, 0.14, 1, 0, 0, cur_status=3)
    assert r == 9                                  # a start state that did not evaluate cannot propose


def test_short_chain_matches_oracle_decisions():
    obs, E, fp, fe, center = _small_problem()
    eps, alpha = 1.2, 0.14                         # (Ex)Full Test + Usage Example.ipynb: run_smala(..., 1.2, 0.14)

    def evaluate(theta):
        lo, go, ho, so, _ = T.orc_logp_d_dd_batch(E, fp, fe, 1.0, obs, np.atleast_2d(theta), nthreads=1)
        return int(so[0]), float(lo[0]), go[0], ho[0]

    def evaluate_mirror(theta):
        lo, go, ho, so, _ = T.mirror_loglik_d_dd(E, fp, fe, 1.0, obs, np.atleast_2d(theta))
        return int(so[0]), float(lo[0]), go[0], ho[0]

    def prior(theta):
        el = E.copy().reshape(-1)
        for v in range(len(fp)):
            el[fp[v] * 7 + fe[v]] = theta[v]
        return bool(T.oracle().orc_prior_hard(1, T.vp(el)))

    nsteps, seed, cid = 120, 99, 4
    chain_o, acc_o, _ = S.smala_chain(T.oracle(), evaluate, prior, center, eps, alpha, seed, cid, 0, nsteps)
    # the device algorithm, run on the host: mirror evaluation + mirror propose / accept
    th = center.copy()
    st, logp, g, H = evaluate_mirror(th)
    acc_m = np.zeros(nsteps, dtype=np.uint8)
    for k in range(nsteps):
        geo, prop, qf = mirror_propose(th, g, H, eps, alpha, seed, cid, k)
        st2, lp2, g2, H2 = evaluate_mirror(prop)
        a, flag = mirror_accept(th, logp, prop, lp2, g2, H2, st2, geo, qf, eps, alpha, seed, cid, k)
        if a:
            th, logp, g, H = prop, lp2, g2, H2
        acc_m[k] = a
        assert np.abs(th - chain_o[k]).max() < 1e-7 * np.abs(center).max(), k
    assert np.array_equal(acc_m, acc_o)
    assert 0.3 < acc_o.mean() < 0.95
