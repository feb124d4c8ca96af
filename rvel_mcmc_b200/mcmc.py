"""Samplers -- Python-3 mirror of the reference's mcmc.py (Mcmc, Mh, Smala, Alsmala, Ensemble, lnprob).

``Mh`` / ``Smala`` / ``Alsmala`` keep the reference's one-chain, numpy-RNG semantics step for step; each
likelihood evaluation is one call into the CUDA engine.  ``Ensemble`` no longer needs emcee: the affine
stretch move of emcee 2.2.1 (the version the reference ran, script.sh:9) is restated in
``StretchSampler`` and evaluates each half-ensemble as ONE batched kernel call.
Fused many-chain device samplers live in ``rvel_mcmc_b200.samplers``.
"""
from datetime import datetime

import numpy as np

from . import _abi
from ._abi import Encounter


class Mcmc(object):
    def __init__(self, initial_state, obs):
        self.state = initial_state.deepcopy()
        self.obs = obs

    def step(self):
        return True

    def step_force(self):
        tries = 1
        while self.step() == False:  # noqa: E712  (reference idiom, mcmc.py:21)
            tries += 1
            pass
        return tries


def lnprob(x, e):
    """Static lnprob handed to the ensemble sampler (mcmc.py:28-35): -inf on any failure."""
    e.state.set_params(x)
    try:
        logp = e.state.get_logp(e.obs)
    except Exception:
        print("Collision! {t}".format(t=datetime.utcnow()))
        return -np.inf
    return logp


def lnprob_batch(X, e):
    """lnprob for a whole set of walkers in one kernel launch; same values as map(lnprob, X)."""
    X = np.atleast_2d(np.asarray(X, dtype=np.float64))
    ctx = _abi.default_context()
    model = e.state._model(ctx)
    logp, status = model.loglik(e.obs._handle(ctx), X)
    logp = np.where(status == _abi.RV_OK, logp, -np.inf)
    e.totalErrorCount += int(np.sum((status != _abi.RV_OK) & (status != _abi.RV_PRIOR)))
    return logp


class StretchSampler(object):
    """Affine-invariant stretch move, restating emcee 2.2.1's EnsembleSampler (Goodman & Weare 2010):
    two half-ensembles; for each walker of a half draw zz = ((a-1)u+1)^2/a, a partner j from the other half,
    propose q = c_j - zz (c_j - s), accept iff (dim-1) ln zz + lnp(q) - lnp(s) > ln u'.  The RNG draws
    (rand(Ns); randint(Nc, size=Ns); rand(Ns)) come from the sampler's own RandomState, as in emcee."""

    def __init__(self, nwalkers, dim, lnprob_batch_fn, args=(), a=2.0, live_dangerously=False, seed=None):
        self.k = nwalkers
        self.dim = dim
        self.a = a
        self.lnprob_batch_fn = lnprob_batch_fn
        self.args = args
        if self.k % 2 != 0:
            raise AssertionError("The number of walkers must be even.")
        if not live_dangerously and self.k < 2 * self.dim:
            raise AssertionError("The number of walkers needs to be more than twice the dimension of your "
                                 "parameter space... unless you're crazy!")
        self._random = np.random.mtrand.RandomState(seed)
        self.naccepted = np.zeros(self.k)
        self.iterations = 0

    @property
    def random_state(self):
        return self._random.get_state()

    def _get_lnprob(self, p):
        lp = np.asarray(self.lnprob_batch_fn(p, *self.args), dtype=np.float64)
        if np.any(np.isnan(lp)):
            raise ValueError("lnprob returned NaN.")
        return lp

    def _propose_stretch(self, p0, p1, lnprob0):
        s = np.atleast_2d(p0)
        Ns = len(s)
        c = np.atleast_2d(p1)
        Nc = len(c)
        zz = ((self.a - 1.) * self._random.rand(Ns) + 1) ** 2. / self.a
        rint = self._random.randint(Nc, size=(Ns,))
        q = c[rint] - zz[:, np.newaxis] * (c[rint] - s)
        newlnprob = self._get_lnprob(q)
        lnpdiff = (self.dim - 1.) * np.log(zz) + newlnprob - lnprob0
        accept = (lnpdiff > np.log(self._random.rand(len(lnpdiff))))
        return q, newlnprob, accept

    def run_mcmc(self, pos0, N, rstate0=None, lnprob0=None):
        if rstate0 is not None:
            self._random.set_state(rstate0)
        p = np.array(pos0, dtype=np.float64)
        if p.shape != (self.k, self.dim):
            raise ValueError("pos0 must have shape (nwalkers, dim)")
        lnprob = lnprob0
        if lnprob is None:
            lnprob = self._get_lnprob(p)
        lnprob = np.array(lnprob, dtype=np.float64)
        halfk = int(self.k / 2)
        first, second = slice(halfk), slice(halfk, self.k)
        for _ in range(int(N)):
            self.iterations += 1
            for S0, S1 in [(first, second), (second, first)]:
                q, newlnp, acc = self._propose_stretch(p[S0], p[S1], lnprob[S0])
                if np.any(acc):
                    lnprob[S0][acc] = newlnp[acc]
                    p[S0][acc] = q[acc]
                    self.naccepted[S0][acc] += 1
        return p, lnprob, self.random_state


class Ensemble(Mcmc):
    """emcee-style affine sampler coupled with the CUDA engine (mcmc.py:40-75)."""

    def __init__(self, initial_state, obs, scales, nwalkers=10, live_dangerously=False):
        super(Ensemble, self).__init__(initial_state, obs)
        self.set_scales(scales)
        self.nwalkers = nwalkers
        self.states = [self.state.get_params() for i in range(nwalkers)]
        self.previous_states = [self.state.get_params() for i in range(nwalkers)]
        self.lnprob = None
        self.totalErrorCount = 0
        for i, s in enumerate(self.states):
            shift = 0.1e-2 * self.scales * np.random.normal(size=self.state.Nvars)
            self.states[i] += shift
        self.sampler = StretchSampler(nwalkers, self.state.Nvars, lnprob_batch, args=[self],
                                      live_dangerously=live_dangerously)

    def step(self):
        self.previous_states = self.states
        self.states, self.lnprob, rstate = self.sampler.run_mcmc(self.states, 1, lnprob0=self.lnprob)
        for i in range(len(self.states)):
            for j in range(len(self.states[0])):
                if self.previous_states[i][j] != self.states[i][j]:
                    return True
        else:
            return False

    def set_scales(self, scales):
        self.scales = np.ones(self.state.Nvars)
        keys = self.state.get_rawkeys()
        for i, k in enumerate(keys):
            if k in scales:
                self.scales[i] = scales[k]


class Mh(Mcmc):
    """Metropolis-Hastings (mcmc.py:80-121)."""

    def __init__(self, initial_state, obs):
        super(Mh, self).__init__(initial_state, obs)
        self.step_size = 3e-5
        self.scales = np.ones(self.state.Nvars)

    def generate_proposal(self):
        prop = self.state.deepcopy()
        shift = self.step_size * self.scales * np.random.normal(size=self.state.Nvars)
        prop.shift_params(shift)
        return prop

    def set_scales(self, scales):
        self.scales = np.ones(self.state.Nvars)
        keys = self.state.get_rawkeys()
        for i, k in enumerate(keys):
            if k in scales:
                self.scales[i] = scales[k]

    def step(self):
        while True:
            try:
                logp = self.state.get_logp(self.obs)
                proposal = self.generate_proposal()
                if proposal.priorHard():
                    return False
                logp_proposal = proposal.get_logp(self.obs)
                if np.exp(logp_proposal - logp) > np.random.uniform():
                    self.state = proposal
                    return True
                return False
            except Encounter:
                print("Collision! {t}".format(t=datetime.utcnow()))
                return False


class Smala(Mcmc):
    """Simplified manifold MALA with the SoftAbs metric (mcmc.py:126-187)."""

    def __init__(self, initial_state, obs, eps, alp):
        super(Smala, self).__init__(initial_state, obs)
        self.epsilon = eps
        self.alpha = alp

    def softabs(self, hessians):
        lam, Q = np.linalg.eig(-hessians)
        lam_twig = lam * 1. / np.tanh(self.alpha * lam)
        H_twig = np.dot(Q, np.dot(np.diag(lam_twig), Q.T))
        return H_twig

    def generate_proposal(self):
        logp, logp_d, logp_dd = self.state.get_logp_d_dd(self.obs)
        Ginv = np.linalg.inv(self.softabs(logp_dd))
        Ginvsqrt = np.linalg.cholesky(Ginv)
        mu = self.state.get_params() + (self.epsilon) ** 2 * np.dot(Ginv, logp_d) / 2.
        newparams = mu + self.epsilon * np.dot(Ginvsqrt, np.random.normal(0., 1., self.state.Nvars))
        prop = self.state.deepcopy()
        prop.set_params(newparams)
        return prop

    def transitionProbability(self, state_from, state_to):
        from scipy import stats
        logp, logp_d, logp_dd = state_from.get_logp_d_dd(self.obs)
        Ginv = np.linalg.inv(self.softabs(logp_dd))
        mu = state_from.get_params() + (self.epsilon) ** 2 * np.dot(Ginv, logp_d) / 2.
        return stats.multivariate_normal.logpdf(state_to.get_params(), mean=mu, cov=(self.epsilon) ** 2 * Ginv)

    def step(self):
        while True:
            try:
                stateStar = self.generate_proposal()
                if stateStar.priorHard():
                    return False
                q_ts_t = self.transitionProbability(self.state, stateStar)
                q_t_ts = self.transitionProbability(stateStar, self.state)
                break
            except Encounter:
                print("Collision! {t}".format(t=datetime.utcnow()))
                return False
            except np.linalg.LinAlgError:
                print("np.linalg.linalg.LinAlgErrorhas occured, investigate later...")
                print(stateStar.get_params())
                print(self.state.get_params())
                raise SystemExit(1)     # the reference calls quit() (mcmc.py:183)
        if np.exp(stateStar.logp - self.state.logp + q_t_ts - q_ts_t) > np.random.uniform():
            self.state = stateStar
            return True
        return False


class Alsmala(Smala):
    """SMALA alternating with MALA steps that reuse stale derivatives (mcmc.py:191-234)."""

    def __init__(self, initial_state, obs, eps, alp):
        super(Alsmala, self).__init__(initial_state, obs, eps, alp)

    def generate_proposal_mala(self):
        logp, logp_d, logp_dd = self.state.get_logp(self.obs), self.state.logp_d, self.state.logp_dd
        Ginv = np.linalg.inv(self.softabs(logp_dd))
        Ginvsqrt = np.linalg.cholesky(Ginv)
        mu = self.state.get_params() + (self.epsilon) ** 2 * np.dot(Ginv, logp_d) / 2.
        newparams = mu + self.epsilon * np.dot(Ginvsqrt, np.random.normal(0., 1., self.state.Nvars))
        prop = self.state.deepcopy()
        prop.set_params(newparams)
        prop.logp_d = logp_d
        prop.logp_dd = logp_dd
        return prop

    def transitionProbability_mala(self, state_from, state_to):
        from scipy import stats
        logp, logp_d, logp_dd = state_from.get_logp(self.obs), state_from.logp_d, state_from.logp_dd
        Ginv = np.linalg.inv(self.softabs(logp_dd))
        mu = state_from.get_params() + (self.epsilon) ** 2 * np.dot(Ginv, logp_d) / 2.
        return stats.multivariate_normal.logpdf(state_to.get_params(), mean=mu, cov=(self.epsilon) ** 2 * Ginv)

    def step_mala(self):
        while True:
            try:
                stateStar = self.generate_proposal_mala()
                if stateStar.priorHard():
                    return False
                q_ts_t = self.transitionProbability_mala(self.state, stateStar)
                q_t_ts = self.transitionProbability_mala(stateStar, self.state)
                break
            except Encounter:
                print("Collision! {t}".format(t=datetime.utcnow()))
                return False
            except np.linalg.LinAlgError:
                print("np.linalg.linalg.LinAlgErrorhas occured, investigate later...")
                print(stateStar.get_params())
                print(self.state.get_params())
                raise SystemExit(1)
        if np.exp(stateStar.logp - self.state.logp + q_t_ts - q_ts_t) > np.random.uniform():
            self.state = stateStar
            return True
        return False
