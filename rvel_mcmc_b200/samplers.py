"""Fused many-walker samplers and their multi-GPU sharding.

* MH / SMALA chains are independent: chain c lives on rank c // (W/G); no collective during sampling
  (RNG streams are keyed by the GLOBAL chain id, so results do not depend on G).
* The affine stretch move needs the complementary half-ensemble: every rank owns a contiguous slice of each
  half, updates it in place, and the updated half is all-gathered (NCCL over NVLink on B200; gloo in the CPU
  tests) before the other half moves.  Nothing else crosses ranks.
"""
import numpy as np


class ShardedStretch(object):
    """Affine stretch ensemble sharded over `world` ranks.

    theta: torch tensor [W][nvars] (the FULL ensemble, replicated on every rank); lnp_local: [W/world] values of
    the walkers this rank owns, ordered (first-half slice, second-half slice).  half_step_fn(S_view, id0_S, C_view,
    lnp_view, step, half) must update its S slice and lnp slice in place (rv_stretch_half_dev on the GPU).
    """

    def __init__(self, half_step_fn, W, nvars, rank=0, world=1, dist=None):
        if W % 2 or (W // 2) % world:
            raise ValueError("walkers (%d) must split evenly into two halves of %d-rank slices" % (W, world))
        self.fn = half_step_fn
        self.W, self.nvars, self.rank, self.world, self.dist = W, nvars, rank, world, dist
        self.h = W // 2
        self.n_loc = self.h // world        # walkers this rank owns in each half

    def owned(self, half):
        lo = half * self.h + self.rank * self.n_loc
        return lo, lo + self.n_loc

    def step(self, theta, lnp_local, step):
        for half in (0, 1):
            lo, hi = self.owned(half)
            S = theta[lo:hi]
            Cfull = theta[self.h:] if half == 0 else theta[:self.h]
            self.fn(S, lo, Cfull, lnp_local[half * self.n_loc:(half + 1) * self.n_loc], step, half)
            if self.world > 1:
                full_half = theta[half * self.h:(half + 1) * self.h]
                self.dist.all_gather_into_tensor(full_half, S.clone())

    def gather_lnp(self, lnp_local, torch):
        """Full lnp[W] on every rank, in ensemble order."""
        if self.world == 1:
            return lnp_local.clone()
        out = torch.empty(self.W, dtype=lnp_local.dtype, device=lnp_local.device)
        for half in (0, 1):
            self.dist.all_gather_into_tensor(out[half * self.h:(half + 1) * self.h],
                                             lnp_local[half * self.n_loc:(half + 1) * self.n_loc].contiguous())
        return out


def chain_shard(W, rank, world):
    """Contiguous block of independent chains owned by `rank` (MH / SMALA)."""
    base, rem = divmod(W, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def integrated_autocorr_time(x, c=5.0):
    """Sokal's windowed integrated autocorrelation time of a 1-D series."""
    x = np.asarray(x, dtype=np.float64)
    n = len(x)
    if n < 8:
        return float("nan")
    y = x - x.mean()
    f = np.fft.rfft(y, 2 * n)
    acf = np.fft.irfft(f * np.conjugate(f))[:n]
    if acf[0] <= 0:
        return float("nan")
    acf /= acf[0]
    tau = 2.0 * np.cumsum(acf) - 1.0
    for m in range(1, n):
        if m >= c * tau[m]:
            return float(tau[m])
    return float(tau[-1])


def ensemble_autocorr_time(x, c=5.0):
    """Integrated autocorrelation time of x[steps][walkers]: the autocovariance is averaged over the walkers before Sokal's
    window is applied (Goodman & Weare 2010; what emcee's integrated_time does) -- far less noisy than per-walker estimates
    when the chains are only a few tens of autocorrelation times long."""
    x = np.asarray(x, dtype=np.float64)
    n = x.shape[0]
    if n < 8:
        return float("nan")
    y = x - x.mean(axis=0, keepdims=True)
    f = np.fft.rfft(y, 2 * n, axis=0)
    acf = np.fft.irfft(f * np.conjugate(f), axis=0)[:n].mean(axis=1)
    if not acf[0] > 0:
        return float("nan")
    acf /= acf[0]
    tau = 2.0 * np.cumsum(acf) - 1.0
    for m in range(1, n):
        if m >= c * tau[m]:
            return float(tau[m])
    return float(tau[-1])


def ess(chain, max_walkers=256):
    """Effective sample size of chain[steps][walkers][nvars] (or [steps][nvars]): total samples / max_i tau_i,
    tau_i the integrated autocorrelation time of parameter i from the walker-averaged autocovariance."""
    chain = np.asarray(chain)
    if chain.ndim == 2:
        chain = chain[:, None, :]
    n, w, d = chain.shape
    taus = [ensemble_autocorr_time(chain[:, :min(w, max_walkers), i]) for i in range(d)]
    tau = float(np.nanmax(taus))
    return n * w / max(tau, 1.0), tau


def stretch_run_sharded(model, obs_handle, theta0, nsteps, a=2.0, seed=0, thin=1, device=None, dist=None,
                        record_chain=True):
    """Affine stretch ensemble of W walkers sharded over the ranks of `dist` (torch.distributed, NCCL on GPUs).

    Every rank holds the full position array in HBM, owns W/(2*world) walkers of each half, moves them with
    rv_stretch_half_dev and all-gathers the updated half (the only exchange).  With world == 1 (or
    dist None) this is the single-GPU ensemble.  Returns dict(theta[W][nvars], lnp[W], chain, chain_lnp, n_accept[W])
    as numpy arrays, identical on every rank and -- because the random numbers are keyed by the ensemble index --
    identical for every world size.
    """
    import torch
    world = dist.get_world_size() if dist is not None else 1
    rank = dist.get_rank() if dist is not None else 0
    theta0 = np.ascontiguousarray(theta0, dtype=np.float64)
    W, nv = theta0.shape
    dev = device if device is not None else torch.device("cuda", model.ctx.device)
    theta = torch.from_numpy(theta0).to(dev)
    stream = torch.cuda.current_stream(dev).cuda_stream

    def half_fn(S, id0, Cfull, lnp_view, step, half):
        model.stretch_half_dev(obs_handle, S.data_ptr(), S.shape[0], id0, Cfull.data_ptr(), Cfull.shape[0],
                               lnp_view.data_ptr(), a, seed, step, half, d_n_accept=nacc_view[half].data_ptr(),
                               stream=stream)

    sh = ShardedStretch(half_fn, W, nv, rank, world, dist)
    n_loc = sh.n_loc
    # lnprob0 of the owned walkers (emcee evaluates the whole ensemble on the first call, mcmc.py:57-59)
    lnp_local = torch.empty(2 * n_loc, dtype=torch.float64, device=dev)
    st_local = torch.empty(2 * n_loc, dtype=torch.int32, device=dev)
    nacc = torch.zeros(2 * n_loc, dtype=torch.int64, device=dev)
    nacc_view = [nacc[:n_loc], nacc[n_loc:]]
    for half in (0, 1):
        lo, hi = sh.owned(half)
        model.loglik_dev(obs_handle, theta[lo:hi].data_ptr(), n_loc, lnp_local[half * n_loc:].data_ptr(),
                         st_local[half * n_loc:].data_ptr(), stream)
    lnp_local[st_local != 0] = float("-inf")
    rows = nsteps // thin if record_chain else 0
    chain = torch.empty((rows, W, nv), dtype=torch.float64, device=dev) if rows else None
    chain_lnp = torch.empty((rows, W), dtype=torch.float64, device=dev) if rows else None
    row = 0
    for k in range(nsteps):
        sh.step(theta, lnp_local, k)
        if rows and (k + 1) % thin == 0:
            chain[row] = theta
            chain_lnp[row] = sh.gather_lnp(lnp_local, torch)
            row += 1
    lnp = sh.gather_lnp(lnp_local, torch)
    nacc_full = sh.gather_lnp(nacc.to(torch.float64), torch)
    torch.cuda.synchronize(dev)
    return dict(theta=theta.cpu().numpy(), lnp=lnp.cpu().numpy(),
                chain=chain.cpu().numpy() if rows else None, chain_lnp=chain_lnp.cpu().numpy() if rows else None,
                n_accept=nacc_full.cpu().numpy().astype(np.uint64))
