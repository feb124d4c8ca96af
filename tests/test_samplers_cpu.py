"""CPU tests of the sampler contract: the oracle's Philox against Random123 known answers, and the multi-rank
sharding logic of the stretch move (world_size-2 gloo) against the single-rank oracle run."""
import ctypes as C
import socket
import sys

import numpy as np

import rvtest as T


def test_philox_known_answers():
    # Random123 kat_vectors: philox4x32 10 rounds
    out = (C.c_uint32 * 4)()
    T.oracle().orc_philox(C.c_uint64(0), C.c_uint64(0), 0, 0, out)
    assert [hex(x) for x in out] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    T.oracle().orc_philox(C.c_uint64(0xffffffffffffffff), C.c_uint64(0xffffffffffffffff), 0xffffffff, 0xffffffff, out)
    assert [hex(x) for x in out] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    T.oracle().orc_philox(C.c_uint64(0x299f31d0a4093822), C.c_uint64(0x85a308d3243f6a88), 0x13198a2e, 0x03707344, out)
    assert [hex(x) for x in out] == ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def _small_problem():
    rng = np.random.RandomState(8)
    obs = T.Obs()
    obs.tf = np.append([0], np.sort(rng.uniform(0, 0.85, 10))); obs.tb = np.sort(rng.uniform(0, -0.85, 10))
    planets = [{"a": 0.35, "m": 0.001965, "h": 0.02, "k": 0.01, "l": 0.3}]
    E = T.elems_from_planets(planets)
    st, rvf = T.orc_rv(E, 0.0, obs.tf); st, rvb = T.orc_rv(E, 0.0, obs.tb)
    obs.errorf = np.full(11, 3e-4); obs.errorb = np.full(10, 3e-4)
    obs.rvf = rvf + 3e-4 * rng.normal(size=11); obs.rvb = rvb + 3e-4 * rng.normal(size=10); obs.Npoints = 20
    fp = [0, 0, 0]; fe = [1, 2, 3]       # free: a, h, k
    center = np.array([0.35, 0.02, 0.01])
    return obs, E, fp, fe, center


def _orc_stretch(obs, E, fp, fe, theta, nsteps, seed, W):
    theta = np.ascontiguousarray(theta.copy()); lnp = np.zeros(W)
    acc = np.zeros((nsteps, W), dtype=np.uint8)
    fpa = np.array(fp, dtype=np.int32); fea = np.array(fe, dtype=np.int32)
    T.oracle().orc_stretch_run(E.shape[0], T.vp(E), len(fp), T.vp(fpa), T.vp(fea), C.c_double(1.0),
                               T.vp(obs.tf), T.vp(obs.rvf), T.vp(obs.errorf), len(obs.tf),
                               T.vp(obs.tb), T.vp(obs.rvb), T.vp(obs.errorb), len(obs.tb), C.c_double(obs.Npoints),
                               T.vp(theta), T.vp(lnp), 0, C.c_double(2.0), C.c_uint64(seed), 0, nsteps, C.c_long(W),
                               None, T.vp(acc), 4)
    return theta, lnp, acc


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, T.ROOT)
    from rvel_mcmc_b200.samplers import ShardedStretch
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    obs, E, fp, fe, center = _small_problem()
    W, nsteps, seed = 16, 6, 77
    theta0 = T.gaussian_ball(center, [3e-4, 0.01, 0.01], W, 1, width=1.0)
    fpa = np.array(fp, dtype=np.int32); fea = np.array(fe, dtype=np.int32)

    def half_fn(S, id0, Cfull, lnp_view, step, half):
        Sn = S.numpy(); Cn = np.ascontiguousarray(Cfull.numpy()); ln = lnp_view.numpy()
        T.oracle().orc_stretch_half(E.shape[0], T.vp(E), len(fp), T.vp(fpa), T.vp(fea), C.c_double(1.0),
                                    T.vp(obs.tf), T.vp(obs.rvf), T.vp(obs.errorf), len(obs.tf),
                                    T.vp(obs.tb), T.vp(obs.rvb), T.vp(obs.errorb), len(obs.tb), C.c_double(obs.Npoints),
                                    Sn.ctypes.data_as(C.c_void_p), C.c_long(Sn.shape[0]), C.c_uint64(id0), T.vp(Cn),
                                    C.c_long(Cn.shape[0]), ln.ctypes.data_as(C.c_void_p), C.c_double(2.0), C.c_uint64(seed),
                                    step, half, None, 1)
    sh = ShardedStretch(half_fn, W, 3, rank, world, dist)
    theta = torch.from_numpy(theta0.copy())
    # initial lnp of the owned walkers
    lnp_local = torch.zeros(2 * sh.n_loc, dtype=torch.float64)
    for half in (0, 1):
        lo, hi = sh.owned(half)
        for i in range(lo, hi):
            Ei = E.copy()
            for v in range(3):
                Ei[fp[v], fe[v]] = theta0[i, v]
            st, lp = T.orc_logp(Ei, 1.0, obs)
            lnp_local[half * sh.n_loc + (i - lo)] = lp if st == 0 else -np.inf
    for s in range(nsteps):
        sh.step(theta, lnp_local, s)
    full_lnp = sh.gather_lnp(lnp_local, torch)
    if rank == 0:
        q.put((theta.numpy().copy(), full_lnp.numpy().copy()))
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_stretch_world2_equals_single_rank():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    theta2, lnp2 = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    obs, E, fp, fe, center = _small_problem()
    theta0 = T.gaussian_ball(center, [3e-4, 0.01, 0.01], 16, 1, width=1.0)
    theta1, lnp1, acc = _orc_stretch(obs, E, fp, fe, theta0, 6, 77, 16)
    assert acc.sum() > 10
    assert np.array_equal(theta1, theta2)          # sharding changes nothing, bit for bit
    assert np.array_equal(lnp1, lnp2)


def test_chain_shard_and_ess():
    from rvel_mcmc_b200.samplers import chain_shard, integrated_autocorr_time, ess
    cover = []
    for r in range(8):
        lo, hi = chain_shard(100003, r, 8)
        cover.append((lo, hi))
    assert cover[0][0] == 0 and cover[-1][1] == 100003
    assert all(cover[i][1] == cover[i + 1][0] for i in range(7))
    rng = np.random.RandomState(0)
    x = np.zeros(20000)
    for i in range(1, len(x)):
        x[i] = 0.8 * x[i - 1] + rng.normal()
    tau = integrated_autocorr_time(x)
    assert abs(tau - 9.0) < 2.0          # (1+rho)/(1-rho) = 9
    n_eff, t = ess(x[:, None])
    assert 1500 < n_eff < 3500
