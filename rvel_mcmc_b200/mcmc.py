"""Samplers -- Python-3 mirror of the reference's mcmc.py (Mcmc, Mh, Smala, Alsmala, Ensemble, lnprob).

``Mh`` / ``Smala`` / ``Alsmala`` keep the reference's one-chain, numpy-RNG semantics step for step; each
likelihood evaluation is one call into the CUDA engine.  ``Ensemble`` no longer needs emcee: the affine
stretch move of emcee 2.2.1 (the version the reference ran, script.sh:9) is restated in
``StretchSampler`` and evaluates each half-ensemble as ONE batched kernel call.
Fused many-chain device samplers live in ``rvel_mcmc_b200.samplers``.
"""
from datetime import datetime, timezone

import numpy as np

from . import _abi
from ._abi import Encounter

def _utcnow():
    """Naive UTC timestamp, what the reference's datetime.utcnow() returned (driver.py:64)."""
    return datetime.now(timezone.utc).replace(tzinfo=None)



class Mcmc(object):
    def __init__(self, initial_state, obs, fast=False):
        """fast (not in the reference, mcmc.py:12-15): evaluate the plain likelihood with the dense-output option -- one
        continuous IAS15 integration per leg with natural steps, velocities at the epochs read from the step's own
        polynomial instead of rebound's truncated step per epoch.  Same logp to ~1e-12 (measured 1e-12 on the HD155358
        ball, profiles/), a step count that no longer grows with the number of epochs, and ~3x lower latency per
        evaluation -- which is what bounds a single chain / a small ensemble on the GPU.  The default (False) keeps the
        reference's step sequence decision by decision."""
        self.state = initial_state.deepcopy()
        if fast:
            self.state.dense_output = True
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
        print("Collision! {t}".format(t=_utcnow()))
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


def _scale_vector(state, scales):
    """Per-parameter scale vector in get_params() order; parameters without an entry keep 1 (mcmc.py:69-75, 98-104)."""
    return np.array([float(scales.get(k, 1.0)) for k in state.get_rawkeys()], dtype=np.float64)


def _collision():
    print("Collision! {t}".format(t=_utcnow()))
    return False


class Ensemble(Mcmc):
    """emcee-style affine sampler coupled with the CUDA engine (mcmc.py:40-75): the walkers start in a ball
    theta + 1e-3*scales*N(0,1) (numpy's global RNG, one normal(size=Nvars) per walker), and each step() is one
    stretch-move sweep over both half-ensembles, every half evaluated as ONE batched kernel call."""

    def __init__(self, initial_state, obs, scales, nwalkers=10, live_dangerously=False, fast=False):
        Mcmc.__init__(self, initial_state, obs, fast=fast)
        self.set_scales(scales)
        self.nwalkers = nwalkers
        centre = self.state.get_params()
        self.previous_states = [centre.copy() for _ in range(nwalkers)]
        self.states = [centre + 0.1e-2 * self.scales * np.random.normal(size=self.state.Nvars) for _ in range(nwalkers)]
        self.lnprob = None
        self.totalErrorCount = 0
        self.sampler = StretchSampler(nwalkers, self.state.Nvars, lnprob_batch, args=[self],
                                      live_dangerously=live_dangerously)

    def step(self):
        """One ensemble sweep; True iff any walker moved (mcmc.py:57-65)."""
        before = np.asarray(self.states, dtype=np.float64)
        self.previous_states = self.states
        self.states, self.lnprob, _ = self.sampler.run_mcmc(self.states, 1, lnprob0=self.lnprob)
        return bool(np.any(before != np.asarray(self.states)))

    def set_scales(self, scales):
        self.scales = _scale_vector(self.state, scales)


class Mh(Mcmc):
    """Metropolis-Hastings with an axis-aligned Gaussian proposal (mcmc.py:80-121).  Draw order per step, as in the
    reference: normal(size=Nvars) for the proposal, then -- only when the proposal passes the hard prior and
    integrates without an encounter -- one uniform() for the accept test."""

    def __init__(self, initial_state, obs, fast=False):
        Mcmc.__init__(self, initial_state, obs, fast=fast)
        self.step_size = 3e-5
        self.scales = np.ones(self.state.Nvars)

    def set_scales(self, scales):
        self.scales = _scale_vector(self.state, scales)

    def generate_proposal(self):
        candidate = self.state.deepcopy()
        candidate.shift_params(self.step_size * self.scales * np.random.normal(size=self.state.Nvars))
        return candidate

    def step(self):
        try:
            current = self.state.get_logp(self.obs)
            candidate = self.generate_proposal()
            if candidate.priorHard():
                return False
            if np.exp(candidate.get_logp(self.obs) - current) > np.random.uniform():
                self.state = candidate
                return True
            return False
        except Encounter:
            return _collision()


class Smala(Mcmc):
    """Simplified manifold MALA with the SoftAbs metric (mcmc.py:126-187).

    At a state with log-posterior derivatives (g, H): G = softabs(H), proposal N(mu, eps^2 G^-1) with
    mu = theta + eps^2 G^-1 g / 2.  One value+gradient+Hessian evaluation (State.get_logp_d_dd) per step."""

    def __init__(self, initial_state, obs, eps, alp):
        Mcmc.__init__(self, initial_state, obs)
        self.epsilon = eps
        self.alpha = alp

    def softabs(self, hessians):
        """Q diag(lam / tanh(alpha lam)) Q^T of -H (mcmc.py:135-139)."""
        lam, Q = np.linalg.eig(-hessians)
        return np.dot(Q, np.dot(np.diag(lam * 1. / np.tanh(self.alpha * lam)), Q.T))

    def _kernel(self, state, derivs):
        """Mean and inverse metric of the proposal launched from `state`; derivs() -> (logp, g, H)."""
        _, grad, hess = derivs(state)
        Ginv = np.linalg.inv(self.softabs(hess))
        return state.get_params() + (self.epsilon) ** 2 * np.dot(Ginv, grad) / 2., Ginv

    def _fresh(self, state):
        return state.get_logp_d_dd(self.obs)

    def _draw(self, derivs, carry=False):
        mu, Ginv = self._kernel(self.state, derivs)
        root = np.linalg.cholesky(Ginv)
        candidate = self.state.deepcopy()
        candidate.set_params(mu + self.epsilon * np.dot(root, np.random.normal(0., 1., self.state.Nvars)))
        if carry:          # Alsmala's MALA step hands the stale derivatives on (mcmc.py:205-206)
            candidate.logp_d, candidate.logp_dd = self.state.logp_d, self.state.logp_dd
        return candidate

    def _density(self, state_from, state_to, derivs):
        from scipy import stats
        mu, Ginv = self._kernel(state_from, derivs)
        return stats.multivariate_normal.logpdf(state_to.get_params(), mean=mu, cov=(self.epsilon) ** 2 * Ginv)

    def generate_proposal(self):
        return self._draw(self._fresh)

    def transitionProbability(self, state_from, state_to):
        return self._density(state_from, state_to, self._fresh)

    def _transition(self, draw, density):
        """Shared accept logic of step / step_mala (mcmc.py:167-187, 214-234)."""
        star = None
        try:
            star = draw()
            if star.priorHard():
                return False
            forward = density(self.state, star)
            backward = density(star, self.state)
        except Encounter:
            return _collision()
        except np.linalg.LinAlgError:
            print("np.linalg.linalg.LinAlgErrorhas occured, investigate later...")
            if star is not None:
                print(star.get_params())
            print(self.state.get_params())
            raise SystemExit(1)     # the reference calls quit() here (mcmc.py:183)
        if np.exp(star.logp - self.state.logp + backward - forward) > np.random.uniform():
            self.state = star
            return True
        return False

    def step(self):
        return self._transition(self.generate_proposal, self.transitionProbability)


class Alsmala(Smala):
    """SMALA alternating with MALA steps that reuse the gradient and Hessian of the last full step (mcmc.py:191-234):
    a MALA step costs one plain likelihood evaluation instead of a variational one."""

    def __init__(self, initial_state, obs, eps, alp):
        Smala.__init__(self, initial_state, obs, eps, alp)

    def _stale(self, state):
        return state.get_logp(self.obs), state.logp_d, state.logp_dd

    def generate_proposal_mala(self):
        return self._draw(self._stale, carry=True)

    def transitionProbability_mala(self, state_from, state_to):
        return self._density(state_from, state_to, self._stale)

    def step_mala(self):
        return self._transition(self.generate_proposal_mala, self.transitionProbability_mala)
