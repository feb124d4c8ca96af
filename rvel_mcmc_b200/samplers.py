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
                self.dist.all_gather_into_tensor(full_half, S)

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


def ess(chain):
    """Effective sample size of chain[steps][walkers][nvars] (or [steps][nvars]): total samples / max_i tau_i,
    tau_i the integrated autocorrelation time of parameter i averaged over walkers."""
    chain = np.asarray(chain)
    if chain.ndim == 2:
        chain = chain[:, None, :]
    n, w, d = chain.shape
    taus = []
    for i in range(d):
        t = [integrated_autocorr_time(chain[:, k, i]) for k in range(min(w, 64))]
        taus.append(np.nanmean(t))
    tau = float(np.nanmax(taus))
    return n * w / max(tau, 1.0), tau
