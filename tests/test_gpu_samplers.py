"""GPU parity of the propose/accept step: with fixed counter-based RNG streams the device MH and affine-stretch
chains must make IDENTICAL accept/reject decisions to the CPU oracle (north_star: over the first 10^4 steps),
and the states must agree to rounding.  Called through the C ABI (rv_mh_run / rv_stretch_run)."""
import ctypes as C

import numpy as np
import pytest

import rvtest as T
from test_samplers_cpu import _small_problem, _orc_stretch

pytestmark = pytest.mark.gpu


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


def _orc_mh(obs, E, fp, fe, hill, theta, scales, step_size, nsteps, seed, first_id=0, first_step=0):
    W = theta.shape[0]
    theta = np.ascontiguousarray(theta.copy()); logp = np.zeros(W)
    acc = np.zeros((nsteps, W), dtype=np.uint8)
    fpa = np.array(fp, dtype=np.int32); fea = np.array(fe, dtype=np.int32)
    sc = np.ascontiguousarray(scales, dtype=np.float64)
    T.oracle().orc_mh_run(E.shape[0], T.vp(E), len(fp), T.vp(fpa), T.vp(fea), C.c_double(hill),
                          T.vp(obs.tf), T.vp(obs.rvf), T.vp(obs.errorf), len(obs.tf),
                          T.vp(obs.tb), T.vp(obs.rvb), T.vp(obs.errorb), len(obs.tb), C.c_double(obs.Npoints),
                          T.vp(theta), T.vp(logp), T.vp(sc), C.c_double(step_size), C.c_uint64(seed), C.c_uint64(first_id),
                          first_step, nsteps, C.c_long(W), None, T.vp(acc), 16)
    return theta, logp, acc


def test_mh_identical_decisions_10k_steps(ctx):
    # configs[0] shape: one planet, free (a,h,k), MH with the reference's scales (mcmc / (Ex)Full Test notebook)
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0)
    W, nsteps = 16, 10000
    theta0 = np.tile(center, (W, 1))
    scales = [3e-4, 0.01, 0.01]
    r = m.mh_run(oh, theta0, scales, 5.0, nsteps, seed=2024, record_chain=False, record_accepts=True)
    th_o, lp_o, acc_o = _orc_mh(obs, E, fp, fe, 1.0, theta0, scales, 5.0, nsteps, 2024)
    assert np.array_equal(r["accepted"], acc_o)                 # every one of 160 000 decisions
    rate = acc_o.mean()
    assert 0.1 < rate < 0.9
    assert np.array_equal(r["n_accept"], acc_o.sum(axis=0).astype(np.uint64))
    assert np.abs(r["theta"] - th_o).max() < 1e-9
    assert np.abs(r["logp"] - lp_o).max() < 1e-6


def test_mh_hd155358_identical_decisions_and_sharding_invariance(ctx):
    obs = T.load_vels("HD155358.vels")
    fixed = np.zeros((2, 7))
    oh, m = _handles(ctx, obs, fixed, T.FP10, T.FE10, 1.0)
    W, nsteps = 48, 150
    theta0 = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 4)
    scales = np.array(T.HD_SCALE_VEC)
    r = m.mh_run(oh, theta0, scales, 0.3, nsteps, seed=99, record_accepts=True)
    th_o, lp_o, acc_o = _orc_mh(obs, fixed, T.FP10, T.FE10, 1.0, theta0, scales, 0.3, nsteps, 99)
    assert np.array_equal(r["accepted"], acc_o)
    assert 0.05 < acc_o.mean() < 0.95
    assert np.abs(r["theta"] - th_o).max() < 1e-9
    assert np.abs(r["logp"] - lp_o).max() < 1e-6
    assert r["chain"].shape == (nsteps, W, 10) and np.array_equal(r["chain"][-1], r["theta"])
    # chains keyed by global id: running a shard [16,32) alone, in two legs, reproduces the same chains
    r1 = m.mh_run(oh, theta0[16:32], scales, 0.3, 100, seed=99, first_chain_id=16, record_chain=False)
    r2 = m.mh_run(oh, r1["theta"], scales, 0.3, 50, seed=99, first_chain_id=16, first_step=100, logp=r1["logp"],
                  record_chain=False)
    assert np.array_equal(r2["theta"], r["theta"][16:32])
    assert np.array_equal(r2["logp"], r["logp"][16:32])


def test_stretch_identical_decisions(ctx):
    obs = T.load_vels("HD155358.vels")
    fixed = np.zeros((2, 7))
    oh, m = _handles(ctx, obs, fixed, T.FP10, T.FE10, 1.0)
    W, nsteps = 64, 160          # 10 240 decisions
    theta0 = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, W, 6)
    r = m.stretch_run(oh, theta0, nsteps, a=2.0, seed=5, record_accepts=True)
    th = np.ascontiguousarray(theta0.copy()); lnp = np.zeros(W); acc = np.zeros((nsteps, W), dtype=np.uint8)
    fpa = np.array(T.FP10, dtype=np.int32); fea = np.array(T.FE10, dtype=np.int32)
    T.oracle().orc_stretch_run(2, T.vp(fixed), 10, T.vp(fpa), T.vp(fea), C.c_double(1.0),
                               T.vp(obs.tf), T.vp(obs.rvf), T.vp(obs.errorf), len(obs.tf),
                               T.vp(obs.tb), T.vp(obs.rvb), T.vp(obs.errorb), len(obs.tb), C.c_double(obs.Npoints),
                               T.vp(th), T.vp(lnp), 0, C.c_double(2.0), C.c_uint64(5), 0, nsteps, C.c_long(W), None, T.vp(acc), 16)
    assert np.array_equal(r["accepted"], acc)
    assert 0.1 < acc.mean() < 0.9
    assert np.abs(r["theta"] - th).max() < 1e-9
    assert np.abs(r["lnp"] - lnp).max() < 1e-6


def test_stretch_small_problem_10k_decisions(ctx):
    # 10^4 accept/reject decisions (64 walkers x 160 ensemble steps); the full 10^4-ENSEMBLE-step horizon is
    # test_gpu_parity_horizon.py.  The stretch map q = c - zz (c - s) amplifies any difference in the POSITIONS by
    # ~e^{0.08} per accepted move (E[ln zz] > 0), so kernel and oracle share the proposal arithmetic bit for bit
    # (explicit fma sequence); the likelihoods then only enter through the decisions.
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0)
    W, nsteps = 64, 160
    theta0 = T.gaussian_ball(center, [3e-4, 0.01, 0.01], W, 1, width=1.0)
    r = m.stretch_run(oh, theta0, nsteps, seed=77, record_chain=False, record_accepts=True)
    th, lnp, acc = _orc_stretch(obs, E, fp, fe, theta0, nsteps, 77, W)
    assert np.array_equal(r["accepted"], acc)
    assert np.array_equal(r["theta"], th)          # positions: bit-identical functions of the decision history


def test_posterior_moments_agree_between_samplers(ctx):
    # the reference's own validation: different samplers must agree on the posterior (KS / means within MCSE)
    from rvel_mcmc_b200.samplers import ess
    obs, E, fp, fe, center = _small_problem()
    oh, m = _handles(ctx, obs, E, fp, fe, 1.0)
    mh = m.mh_run(oh, np.tile(center, (256, 1)), [3e-4, 0.01, 0.01], 5.0, 1500, seed=1, thin=5)
    st = m.stretch_run(oh, T.gaussian_ball(center, [3e-4, 0.01, 0.01], 256, 2, width=1.0), 1500, seed=2, thin=5)
    a = mh["chain"][100:].reshape(-1, 3); b = st["chain"][100:].reshape(-1, 3)
    ne_a, _ = ess(mh["chain"][100:]); ne_b, _ = ess(st["chain"][100:])
    for i in range(3):
        mcse = np.sqrt(a[:, i].var() / max(ne_a, 10) + b[:, i].var() / max(ne_b, 10))
        assert abs(a[:, i].mean() - b[:, i].mean()) < 5 * mcse + 1e-12, i
        assert abs(a[:, i].std() / b[:, i].std() - 1.0) < 0.15


def test_reference_api_samplers_run_on_gpu():
    import os
    from rvel_mcmc_b200 import observations, state, driver
    np.random.seed(12)
    true_state = state.State([{"a": 0.2275, "h": 0., "k": 0., "m": 0.001965}], ignore_vars=["m"])
    obs = observations.FakeObservation(true_state, Npoints=70, error=3.5e-4, errorVar=9e-5, tmax=1.37)
    assert len(obs.tf) == 36 and len(obs.tb) == 35 and obs.tf[0] == 0
    bundle, h = driver.run_mh("t", 60, true_state, obs, {'a': 3e-4, 'h': 0.01, 'k': 0.01}, 5, printing_every=1000)
    assert bundle.mcmc_chain.shape == (61, 3) and np.all(np.isfinite(bundle.mcmc_chainlogp))
    assert len(np.unique(bundle.mcmc_chain[:, 0])) > 5
    bundle, h = driver.run_emcee("t", 32 * 8, true_state, obs, 32, {'a': 3e-4, 'h': 0.01, 'k': 0.01}, printing_every=1000)
    assert bundle.mcmc_chain.shape == (256, 3) and bundle.mcmc_is_emcee
    act = driver.ac_times(bundle)
    assert act.shape == (3,)


def test_two_gpu_sharded_samplers_equal_single_gpu():
    """N=2 NCCL path (skipped on a 1-GPU box): sharded stretch ensemble and MH shards equal the single-GPU runs."""
    import os
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                        "--master-addr", "127.0.0.1", "--master-port", "29533",
                        os.path.join(T.ROOT, "bench_scripts", "multigpu_check.py"), "--walkers", "512", "--steps", "4"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert '"stretch_equal_single_gpu": true' in r.stdout and '"mh_shard_equal_single_gpu": true' in r.stdout


def test_fused_device_drivers_agree_with_each_other():
    """driver.run_*_gpu: whole sampling loops in one library call; the reference's cross-sampler KS check (driver.py:416-425)."""
    from rvel_mcmc_b200 import observations, state, driver
    np.random.seed(200000)
    true_state = state.State([{"a": 0.2275, "h": 0., "k": 0., "m": 0.001965}], ignore_vars=["m"])
    obs = observations.FakeObservation(true_state, Npoints=70, error=3.5e-4, errorVar=9e-5, tmax=1.37)   # (Ex)Full Test notebook
    scal = {'a': 3e-4, 'h': 0.01, 'k': 0.01}
    bm, _ = driver.run_mh_gpu("t", 1200, true_state, obs, scal, 5, nchains=64, seed=1)
    be, _ = driver.run_emcee_gpu("t", 64 * 600, true_state, obs, 64, scal, seed=2, fast=True)      # dense-output likelihood
    bs, _ = driver.run_smala_gpu("t", 300, true_state, obs, 1.2, 0.14, nchains=64, seed=3)
    ba, _ = driver.run_alsmala_gpu("t", 300, true_state, obs, 1.2, 0.14, 3.0, nchains=64, seed=4)
    assert bm.mcmc_chain.shape == (64 * 1200, 3) and be.mcmc_chain.shape == (64 * 600, 3) and bs.mcmc_chain.shape == (64 * 300, 3)

    def trimmed(b, per, burn):
        c = b.mcmc_chain.reshape(64, per, 3)[:, burn:, :]
        return c.reshape(-1, 3)
    cm, ce, cs, ca = trimmed(bm, 1200, 200), trimmed(be, 600, 150), trimmed(bs, 300, 50), trimmed(ba, 300, 50)
    for name, other in (("emcee", ce), ("smala", cs), ("alsmala", ca)):
        ks = driver.calc_kstatistic(cm[::7], other[::3])
        assert max(ks) < 0.06, (name, ks)                                        # notebook: D = 0.014-0.050 between samplers
    assert driver.ac_times(bs).max() <= 3                                       # SMALA AC 1/1/1 in the notebook
    # the truth is recovered
    assert abs(cs[:, 0].mean() - 0.2275) < 4 * cs[:, 0].std()


def test_reference_api_smala_and_alsmala_step_by_step():
    """mcmc.Smala / mcmc.Alsmala through driver.run_smala / run_alsmala (one Python iteration per step, numpy RNG)."""
    from rvel_mcmc_b200 import observations, state, driver
    np.random.seed(5)
    true_state = state.State([{"a": 0.2275, "h": 0., "k": 0., "m": 0.001965}], ignore_vars=["m"])
    obs = observations.FakeObservation(true_state, Npoints=70, error=3.5e-4, errorVar=9e-5, tmax=1.37)
    bs, _ = driver.run_smala("t", 60, true_state, obs, 1.2, 0.14, printing_every=1000)
    assert bs.mcmc_chain.shape == (61, 3) and np.isfinite(bs.mcmc_chainlogp).all()
    assert 15 < len(np.unique(bs.mcmc_chain[:, 0])) <= 61                     # acceptance ~65 % in the notebook
    ba, _ = driver.run_alsmala("t", 60, true_state, obs, 1.2, 0.14, 3.0, 0.0, printing_every=1000)
    assert ba.mcmc_chain.shape == (61, 3) and len(np.unique(ba.mcmc_chain[:, 0])) > 10
    assert ba.mcmc.state.logp_d is not None and ba.mcmc.state.logp_dd.shape == (3, 3)


def test_device_group_single_process_multi_gpu():
    """rvel_mcmc_b200.multigpu.DeviceGroup: all visible GPUs from one process (one host thread per GPU); results equal the
    single-GPU ones bit for bit.  Runs with however many GPUs the box has (1 on the default test box)."""
    import os
    from rvel_mcmc_b200 import observations, state, _abi
    from rvel_mcmc_b200.multigpu import DeviceGroup
    obs = observations.Observation_FromFile(os.path.join(T.GOLDEN, "HD155358.vels"), Npoints=100)
    st = state.State(T.planets_from_vec(T.HD_SOL)); st.hillRadiusFactor = 2.
    g = DeviceGroup()
    try:
        theta = T.gaussian_ball(T.HD_SOL, T.HD_SCALE_VEC, 1000, 3)
        lg, sg = g.loglik(st, obs, theta)
        ctx = _abi.default_context()
        l1, s1 = st._model(ctx).loglik(obs._handle(ctx), theta)
        assert np.array_equal(lg, l1) and np.array_equal(sg, s1)
        r = g.mh_run(st, obs, theta[:64], np.array(T.HD_SCALE_VEC), 0.1, 8, seed=4, record_accepts=True)
        r1 = st._model(ctx).mh_run(obs._handle(ctx), theta[:64], np.array(T.HD_SCALE_VEC), 0.1, 8, seed=4, record_accepts=True)
        assert np.array_equal(r["chain"], r1["chain"]) and np.array_equal(r["accepted"], r1["accepted"])
        lv, gv, hv, sv = g.loglik_d_dd(st, obs, theta[:16])
        assert lv.shape == (16,) and gv.shape == (16, 10) and hv.shape == (16, 10, 10) and (sv == 0).all()
        # the stretch ensemble: slices moved on their own GPUs, exchanged by peer copies
        W = 64 * len(g)
        e = g.stretch_run(st, obs, theta[:W], 5, seed=6)
        e1 = st._model(ctx).stretch_run(obs._handle(ctx), theta[:W], 5, seed=6, record_chain=False)
        assert np.array_equal(e["theta"], e1["theta"]) and np.array_equal(e["lnp"], e1["lnp"])
        assert np.array_equal(e["n_accept"], e1["n_accept"])
        # with chain rows, thinning and a restart (first_step / lnp): same rows as the single-GPU call
        ec = g.stretch_run(st, obs, theta[:W], 6, seed=6, thin=2, record_chain=True)
        e1c = st._model(ctx).stretch_run(obs._handle(ctx), theta[:W], 6, seed=6, thin=2)
        assert ec["chain"].shape == (3, W, 10) and np.array_equal(ec["chain"], e1c["chain"])
        assert np.array_equal(ec["chain_lnp"], e1c["chain_lnp"]) and np.array_equal(ec["theta"], e1c["theta"])
        er = g.stretch_run(st, obs, ec["chain"][0], 4, seed=6, first_step=2, lnp=ec["chain_lnp"][0])
        assert np.array_equal(er["theta"], e1c["theta"]) and np.array_equal(er["lnp"], e1c["lnp"])
    finally:
        g.close()
