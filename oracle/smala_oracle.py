"""CPU ORACLE of Smala.step (test infrastructure, NOT product code) -- numpy/scipy restatement of mcmc.py:126-187.

    softabs               mcmc.py:135-139   lam, Q = eig(-H); lam~ = lam/tanh(alpha lam); G = Q diag(lam~) Q^T
    generate_proposal     mcmc.py:144-153   Ginv = inv(G); L = cholesky(Ginv); mu = theta + eps^2 Ginv g/2; theta* = mu + eps L z
    transitionProbability mcmc.py:158-162   scipy.stats.multivariate_normal.logpdf(to, mean=mu(from), cov=eps^2 Ginv(from))
    step                  mcmc.py:167-187   priorHard -> reject; Encounter -> reject; accept iff exp(...) > u

The likelihood, gradient and Hessian come from oracle/rv_oracle.c (orc_get_logp_d_dd).  PARITY: no seeded SMALA chain is
stored in the reference ("parity unpinned" for accept/reject sequences); the random numbers follow the device samplers'
counter-based contract (Philox-4x32-10 keyed by seed / chain id / step / stream; z_j by Box-Muller) restated here through
orc_philox.  Only tests/ may import this module.
"""
import ctypes as C

import numpy as np
from scipy import stats

RNG_ACCEPT, RNG_NORMAL = 1, 0x100


def _u53(hi, lo):
    return (float(hi >> 5) * 67108864.0 + float(lo >> 6) + 0.5) * (1.0 / 9007199254740992.0)


def philox(lib, seed, cid, step, stream):
    out = (C.c_uint32 * 4)()
    lib.orc_philox(C.c_uint64(seed), C.c_uint64(cid), C.c_uint32(step), C.c_uint32(stream), out)
    return [int(x) for x in out]


def normals(lib, seed, cid, step, n):
    z = np.zeros(n + 1)
    for j in range((n + 1) // 2):
        r = philox(lib, seed, cid, step, RNG_NORMAL + j)
        u1, u2 = _u53(r[0], r[1]), _u53(r[2], r[3])
        rad = np.sqrt(-2.0 * np.log(u1))
        z[2 * j] = rad * np.cos(2.0 * np.pi * u2)
        z[2 * j + 1] = rad * np.sin(2.0 * np.pi * u2)
    return z[:n]


def softabs(hessians, alpha):
    lam, Q = np.linalg.eig(-hessians)
    lam_twig = lam * 1. / np.tanh(alpha * lam)
    return np.dot(Q, np.dot(np.diag(lam_twig), Q.T))


def proposal_mean_cov(theta, logp_d, logp_dd, eps, alpha):
    Ginv = np.linalg.inv(softabs(logp_dd, alpha))
    mu = theta + (eps) ** 2 * np.dot(Ginv, logp_d) / 2.
    return mu, Ginv


def smala_chain(lib, evaluate, prior_hard, theta0, eps, alpha, seed, chain_id, first_step, nsteps):
    """evaluate(theta) -> (status, logp, grad, hess); prior_hard(theta) -> bool.
    Returns (chain[nsteps][n], accepted[nsteps], logp_final)."""
    theta = np.array(theta0, dtype=float)
    n = len(theta)
    st, logp, g, H = evaluate(theta)
    assert st == 0, "start state must evaluate"
    chain = np.zeros((nsteps, n))
    accepted = np.zeros(nsteps, dtype=np.uint8)
    for k in range(nsteps):
        step = first_step + k
        mu, Ginv = proposal_mean_cov(theta, g, H, eps, alpha)
        Ginvsqrt = np.linalg.cholesky(Ginv)
        new = mu + eps * np.dot(Ginvsqrt, normals(lib, seed, chain_id, step, n))
        acc = False
        if not prior_hard(new):
            q_ts_t = stats.multivariate_normal.logpdf(new, mean=mu, cov=eps ** 2 * Ginv)
            st2, logp2, g2, H2 = evaluate(new)
            if st2 == 0:
                mu2, Ginv2 = proposal_mean_cov(new, g2, H2, eps, alpha)
                q_t_ts = stats.multivariate_normal.logpdf(theta, mean=mu2, cov=eps ** 2 * Ginv2)
                r = philox(lib, seed, chain_id, step, RNG_ACCEPT)
                acc = np.exp(logp2 - logp + q_t_ts - q_ts_t) > _u53(r[0], r[1])
        if acc:
            theta, logp, g, H = new, logp2, g2, H2
        chain[k] = theta
        accepted[k] = 1 if acc else 0
    return chain, accepted, logp


RNG_SCHEDULE = 2
SCHEDULE_ID = 0xFFFFFFFFFFFFFFFF


def alsmala_chain(lib, evaluate, evaluate_plain, prior_hard, theta0, eps, alpha, bern_a, niter_total, seed, chain_id,
                  first_step, nsteps):
    """Alsmala (mcmc.py:191-234) under run_alsmala's schedule (driver.py:171-200): iteration i is a full SMALA step with
    probability exp(-bern_a*i/Niter), else step_mala, whose proposal and BOTH transition densities use the stale gradient /
    Hessian carried by the state (mcmc.py:195-212) and which evaluates only the plain likelihood (evaluate_plain(theta) ->
    (status, logp)).  Schedule draw: one per iteration, Philox(seed, SCHEDULE_ID, i, RNG_SCHEDULE), shared by all chains.
    Returns (chain, accepted, full_step, logp_final)."""
    theta = np.array(theta0, dtype=float)
    n = len(theta)
    st, logp, g, H = evaluate(theta)
    assert st == 0
    chain = np.zeros((nsteps, n)); accepted = np.zeros(nsteps, dtype=np.uint8); full = np.zeros(nsteps, dtype=np.uint8)
    for k in range(nsteps):
        step = first_step + k
        r = philox(lib, seed, SCHEDULE_ID, step, RNG_SCHEDULE)
        do_full = np.exp(-bern_a * step / float(niter_total if niter_total > 0 else nsteps)) > _u53(r[0], r[1])
        full[k] = 1 if do_full else 0
        mu, Ginv = proposal_mean_cov(theta, g, H, eps, alpha)
        new = mu + eps * np.dot(np.linalg.cholesky(Ginv), normals(lib, seed, chain_id, step, n))
        acc = False
        if not prior_hard(new):
            q_ts_t = stats.multivariate_normal.logpdf(new, mean=mu, cov=eps ** 2 * Ginv)
            if do_full:
                st2, logp2, g2, H2 = evaluate(new)
            else:
                st2, logp2 = evaluate_plain(new)
                g2, H2 = g, H                                   # prop.logp_d = logp_d; prop.logp_dd = logp_dd (mcmc.py:205-206)
            if st2 == 0:
                mu2, Ginv2 = proposal_mean_cov(new, g2, H2, eps, alpha)
                q_t_ts = stats.multivariate_normal.logpdf(theta, mean=mu2, cov=eps ** 2 * Ginv2)
                ra = philox(lib, seed, chain_id, step, RNG_ACCEPT)
                acc = np.exp(logp2 - logp + q_t_ts - q_ts_t) > _u53(ra[0], ra[1])
        if acc:
            theta, logp, g, H = new, logp2, g2, H2
        chain[k] = theta; accepted[k] = 1 if acc else 0
    return chain, accepted, full, logp
