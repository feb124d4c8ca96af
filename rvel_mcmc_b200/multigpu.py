"""Several GPUs from ONE Python process (the reference is a single process; `script.sh` only backgrounds 8 unrelated
jobs).  A DeviceGroup owns one rv_ctx per visible GPU and shards a batch of parameter vectors / independent chains over
them by contiguous blocks, one host thread per GPU (ctypes releases the GIL during the library call).  Nothing crosses
GPUs: MH / SMALA chains and likelihood batches are independent, and the random streams are keyed by the GLOBAL chain id,
so the results are bit-identical to a single-GPU run.  The affine stretch ensemble needs one exchange per half-step:
DeviceGroup.stretch_run keeps a full copy of the positions on every GPU and copies each GPU's freshly updated slice to
the others with cudaMemcpyPeer (NVLink peer-to-peer DMA); under torchrun, samplers.stretch_run_sharded does the same with
an NCCL all-gather."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np

from . import _abi
from .samplers import chain_shard


class DeviceGroup(object):
    def __init__(self, devices=None):
        if devices is None:
            import ctypes as C
            lib = _abi.load()
            devices = []
            d = 0
            while True:          # probe devices until rv_ctx_create refuses
                h = C.c_void_p()
                if lib.rv_ctx_create(d, C.byref(h)) != 0:
                    break
                lib.rv_ctx_destroy(h)
                devices.append(d)
                d += 1
            if not devices:
                raise _abi.RvGpuError("no usable GPU: %s" % lib.rv_last_error(None).decode())
        self.devices = list(devices)
        self.ctxs = [_abi.Context(d) for d in self.devices]
        self.pool = ThreadPoolExecutor(len(self.ctxs))

    def __len__(self):
        return len(self.ctxs)

    def close(self):
        self.pool.shutdown()
        for c in self.ctxs:
            c.close()

    def _handles(self, state, obs):
        return [(state._model(c), _obs_handle(obs, c)) for c in self.ctxs]

    def _map(self, W, fn):
        """fn(rank, lo, hi) on every GPU's block of [0, W); returns the list of results in rank order."""
        shards = [chain_shard(W, r, len(self.ctxs)) for r in range(len(self.ctxs))]
        futs = [self.pool.submit(fn, r, lo, hi) for r, (lo, hi) in enumerate(shards) if hi > lo]
        return [f.result() for f in futs]

    # ---- State.get_logp / get_logp_d_dd for a batch -------------------------------------------------------------
    def loglik(self, state, obs, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        hs = self._handles(state, obs)
        parts = self._map(len(theta), lambda r, lo, hi: hs[r][0].loglik(hs[r][1], theta[lo:hi]))
        return np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts])

    def loglik_d_dd(self, state, obs, theta):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        hs = self._handles(state, obs)
        parts = self._map(len(theta), lambda r, lo, hi: hs[r][0].loglik_d_dd(hs[r][1], theta[lo:hi]))
        return tuple(np.concatenate([p[k] for p in parts]) for k in range(4))

    # ---- independent chains -------------------------------------------------------------------------------------
    @staticmethod
    def _merge(parts):
        out = {}
        for k in parts[0]:
            vals = [p[k] for p in parts]
            if vals[0] is None:
                out[k] = None
            elif k in ("chain", "chain_logp", "accepted"):
                out[k] = np.concatenate(vals, axis=1)          # [steps][W][...]
            else:
                out[k] = np.concatenate(vals, axis=0)
        return out

    def mh_run(self, state, obs, theta, scales, step_size, nsteps, seed=0, logp=None, first_chain_id=0, **kw):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        hs = self._handles(state, obs)
        if logp is not None:
            logp = np.ascontiguousarray(logp, dtype=np.float64)
            if logp.shape != (len(theta),):
                raise ValueError("logp must hold one value per chain")
        # per-chain arguments are sliced with the chains; the RNG streams are keyed by the GLOBAL chain id
        return self._merge(self._map(len(theta), lambda r, lo, hi: hs[r][0].mh_run(
            hs[r][1], theta[lo:hi], scales, step_size, nsteps, seed=seed, first_chain_id=first_chain_id + lo,
            logp=None if logp is None else logp[lo:hi], **kw)))

    def smala_run(self, state, obs, theta, eps, alpha, nsteps, seed=0, first_chain_id=0, **kw):
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        hs = self._handles(state, obs)
        return self._merge(self._map(len(theta), lambda r, lo, hi: hs[r][0].smala_run(
            hs[r][1], theta[lo:hi], eps, alpha, nsteps, seed=seed, first_chain_id=first_chain_id + lo, **kw)))


def _stretch_run(group, state, obs, theta0, nsteps, a=2.0, seed=0):
    theta0 = np.ascontiguousarray(theta0, dtype=np.float64)
    W, nv = theta0.shape
    G = len(group.ctxs)
    if W % 2 or (W // 2) % G:
        raise ValueError("walkers (%d) must split evenly into two halves of %d-GPU slices" % (W, G))
    h, n_loc = W // 2, (W // 2) // G
    hs = group._handles(state, obs)
    row = nv * 8
    d_theta = [c.dev_alloc(W * row) for c in group.ctxs]
    d_lnp = [c.dev_alloc(2 * n_loc * 8) for c in group.ctxs]
    d_st = [c.dev_alloc(2 * n_loc * 4) for c in group.ctxs]
    d_nacc = [c.dev_alloc(2 * n_loc * 8) for c in group.ctxs]
    try:
        def owned(r, half):
            lo = half * h + r * n_loc
            return lo, lo + n_loc

        def init(r):
            c, (m, oh) = group.ctxs[r], hs[r]
            c.dev_upload(d_theta[r], theta0)
            c.dev_upload(d_nacc[r], np.zeros(2 * n_loc, dtype=np.uint64))
            for half in (0, 1):
                lo, _ = owned(r, half)
                m.loglik_dev(oh, d_theta[r] + lo * row, n_loc, d_lnp[r] + half * n_loc * 8, d_st[r] + half * n_loc * 4)
            c.sync()
            lnp = np.zeros(2 * n_loc); st = np.zeros(2 * n_loc, dtype=np.int32)
            c.dev_download(lnp, d_lnp[r]); c.dev_download(st, d_st[r])
            lnp[st != 0] = -np.inf                       # lnprob(): -inf on any failure (mcmc.py:28-35)
            c.dev_upload(d_lnp[r], lnp)
        list(group.pool.map(init, range(G)))
        for k in range(nsteps):
            for half in (0, 1):
                def move(r):
                    c, (m, oh) = group.ctxs[r], hs[r]
                    lo, _ = owned(r, half)
                    comp = d_theta[r] + (h if half == 0 else 0) * row
                    m.stretch_half_dev(oh, d_theta[r] + lo * row, n_loc, lo, comp, h, d_lnp[r] + half * n_loc * 8, a, seed, k, half,
                                       d_n_accept=d_nacc[r] + half * n_loc * 8)
                    c.sync()
                list(group.pool.map(move, range(G)))

                def spread(r):                            # this GPU's updated slice -> every other GPU's copy
                    lo, _ = owned(r, half)
                    for q in range(G):
                        if q != r:
                            group.ctxs[r].dev_copy_to(group.ctxs[q], d_theta[q] + lo * row, d_theta[r] + lo * row, n_loc * row)
                list(group.pool.map(spread, range(G)))
        theta = np.zeros((W, nv)); group.ctxs[0].dev_download(theta, d_theta[0])
        lnp = np.zeros(W); nacc = np.zeros(W, dtype=np.uint64)
        for r in range(G):
            part = np.zeros(2 * n_loc); pa = np.zeros(2 * n_loc, dtype=np.uint64)
            group.ctxs[r].dev_download(part, d_lnp[r]); group.ctxs[r].dev_download(pa, d_nacc[r])
            for half in (0, 1):
                lo, hi = owned(r, half)
                lnp[lo:hi] = part[half * n_loc:(half + 1) * n_loc]
                nacc[lo:hi] = pa[half * n_loc:(half + 1) * n_loc]
        return dict(theta=theta, lnp=lnp, n_accept=nacc)
    finally:
        for r, c in enumerate(group.ctxs):
            for p in (d_theta[r], d_lnp[r], d_st[r], d_nacc[r]):
                c.dev_free(p)


DeviceGroup.stretch_run = lambda self, state, obs, theta0, nsteps, a=2.0, seed=0: _stretch_run(self, state, obs, theta0, nsteps, a, seed)


def _obs_handle(obs, ctx):
    """One rv_obs per context, from that context's content-keyed cache (closed with the context)."""
    return ctx.obs_handle(obs)
