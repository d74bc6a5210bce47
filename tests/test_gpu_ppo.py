"""GPU parity tests of the PPO hot path (policy forward in the rollout kernel, GAE, minibatch
gradient, clip + Adam) against oracle/ppo_oracle.py -- a plain-torch restatement of SB3's
published algorithm.  PARITY UNPINNED: SB3 is not in the reference tree and not installed; the
reference has no test of it (oracle/ppo_oracle.py header).

Tolerances (float32 CUDA-core path with the MUFU-based tanh, oracle in float64):
    forward mean / value, sampled action, value      |err| <= 2e-5 * max(1, |ref|)
    log-prob                                         |err| <= 1e-4
    GAE advantages / returns                         |err| <= 2e-5 * max(1, |ref|)
    minibatch gradient (sum form)                    |err| <= 2e-4 * max|ref grad|  (+ tiny abs floor)
    parameters after k Adam steps                    |err| <= 1e-5 abs (k small; Adam normalises steps to ~lr)
"""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import drone_oracle as do  # noqa: E402
from oracle import philox, ppo_oracle as po, verify  # noqa: E402


@pytest.fixture(scope="module")
def drl():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import drone_rl_b200
    import drone_rl_b200.ppo  # noqa: F401
    return drone_rl_b200


def _rand_params(seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    flat = po.init_params(seed, dtype=torch.float64)
    # perturb everything (biases and log_std are zero at init) so every term is exercised
    flat = flat + 0.1 * scale * torch.randn(po.N_PARAMS, generator=g, dtype=torch.float64)
    return flat


def _close(got, ref, rel, what, floor=1.0):
    got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
    err = np.abs(got - ref)
    bound = rel * np.maximum(np.abs(ref), floor)
    assert (err <= bound).all(), f"{what}: worst err/bound {np.max(err / bound):.3g}"
    return float(np.max(err / bound))


def test_init_matches_oracle(drl):
    from drone_rl_b200.ppo import init_policy_params, unpack_params
    a, b = init_policy_params(3), po.init_params(3)
    assert torch.equal(a, b) and a.numel() == 10697
    p = unpack_params(a)
    assert p["pi.W1"].shape == (64, 15) and torch.count_nonzero(p["pi.b1"]) == 0 and (p["log_std"] == 0).all()
    w = p["pi.W2"].double()
    assert torch.allclose(w @ w.t(), 2 * torch.eye(64, dtype=torch.float64), atol=1e-5)       # orthogonal, gain sqrt 2
    w = p["pi.W3"].double()
    assert torch.allclose(w @ w.t(), 1e-4 * torch.eye(4, dtype=torch.float64), atol=1e-8)     # gain 0.01


def test_policy_forward(drl):
    from drone_rl_b200.ppo import PPO
    model = PPO(8, n_steps=4)
    flat = _rand_params(1)
    model.params.copy_(flat.float().cuda())
    g = torch.Generator().manual_seed(0)
    obs = torch.randn(5000, 15, generator=g, dtype=torch.float64) * torch.tensor([3.0] * 3 + [4.0] * 3 + [2.0] * 3 + [10.0] * 3 + [1.0] * 3, dtype=torch.float64)
    obs32 = obs.float()
    mean, value = model.policy_forward(obs32.cuda())
    rm, rv, _ = po.forward(flat.float().double(), obs32.double())
    _close(mean.cpu().numpy(), rm.numpy(), 2e-5, "mean")
    _close(value.cpu().numpy(), rv.numpy(), 2e-5, "value")
    model.close()


def lib_rejects_bad_mode():
    from drone_rl_b200 import _lib
    return _lib.load().dronecu_set_rollout_kernel(3) != 0


@pytest.fixture
def rollout_kernel_mode():
    from drone_rl_b200 import _lib
    lib = _lib.load()
    yield lambda mode: _lib.check(lib.dronecu_set_rollout_kernel(mode))
    lib.dronecu_set_rollout_kernel(0)


@pytest.mark.parametrize("kernel", ["thread_per_env", "warp_per_env"])
def test_rollout_policy_record(drl, rollout_kernel_mode, kernel):
    """Every recorded (obs, action, logp, value, reward, done) of an in-kernel-policy rollout is
    reproduced by the oracle policy + Philox noise + the float64 env oracle, teacher-forced -- for both float32 kernels
    (one thread per env: large batches; one warp per env: the default up to 4096 envs)."""
    from drone_rl_b200.ppo import PPO
    from drone_rl_b200._lib import PolicyOut
    import ctypes as C
    from drone_rl_b200 import _lib
    rollout_kernel_mode({"thread_per_env": 1, "warp_per_env": 2}[kernel])
    assert lib_rejects_bad_mode()
    n, K, seed = 3000, 40, 21
    model = PPO(drl.DroneBatch(n, drl.EnvConfig.single(), seed=seed, env_offset=77), n_steps=K, seed=seed)
    flat = _rand_params(5, scale=0.3)
    # a policy whose mean sits near hover so episodes live long enough to matter
    with torch.no_grad():
        off = po.offsets()["pi.b3"][0]
        flat[off:off + 4] = 2.45
        off = po.offsets()["log_std"][0]
        flat[off:off + 4] = torch.tensor([-0.5, -0.2, 0.1, -1.0], dtype=torch.float64)
    model.params.copy_(flat.float().cuda())
    b = model.buf
    last_obs = torch.empty(n, 15, device="cuda")
    out = PolicyOut(b.obs.data_ptr(), b.actions.data_ptr(), b.logp.data_ptr(), b.value.data_ptr(),
                    b.reward.data_ptr(), b.done.data_ptr(), b.last_value.data_ptr(), last_obs.data_ptr())
    t0 = model.batch.global_step
    _lib.check(model.lib.dronecu_rollout_policy(model.batch._h, K, C.c_void_p(model.params.data_ptr()), 0, C.byref(out), None))
    torch.cuda.synchronize()
    obs, act = b.obs.cpu().numpy(), b.actions.cpu().numpy()
    theta = flat.float().double()
    ids = np.arange(77, 77 + n, dtype=np.uint64)
    std = np.exp(theta[-4:].numpy())
    for k in range(K):
        mean, value, log_std = po.forward(theta, torch.from_numpy(obs[k]).double())
        z = philox.noise_normals(seed, ids, t0 + k)
        _close(act[k], mean.numpy() + std * z, 2e-5, f"action k={k}")
        lp = po.log_prob(mean, log_std, torch.from_numpy(act[k]).double())
        assert np.abs(b.logp[k].cpu().numpy() - lp.numpy()).max() < 1e-4
        _close(b.value[k].cpu().numpy(), value.numpy(), 2e-5, f"value k={k}")
    _, lv, _ = po.forward(theta, last_obs.cpu().double())
    _close(b.last_value.cpu().numpy(), lv.numpy(), 2e-5, "last_value")
    # the env transitions under the CLIPPED actions
    nxt = np.concatenate([obs[1:], last_obs.cpu().numpy()[None]], 0)
    rep = verify.check_rollout(obs[0], np.clip(act, 0, np.float32(7.3575)), nxt, b.reward.cpu().numpy(),
                               b.done.cpu().numpy().astype(bool), spec=do.SINGLE, seed=seed, env_ids=ids)
    assert rep["dones"] > 0
    # deterministic mode: action == mean, bit for bit the same forward
    model.batch.reset()
    out2 = PolicyOut(b.obs.data_ptr(), b.actions.data_ptr(), b.logp.data_ptr(), None, None, None, None, None)
    _lib.check(model.lib.dronecu_rollout_policy(model.batch._h, 1, C.c_void_p(model.params.data_ptr()), 1, C.byref(out2), None))
    mean, _ = model.policy_forward(b.obs[0])
    if kernel == "thread_per_env":
        assert torch.equal(mean, b.actions[0])
    else:       # identical hidden layers; the four head sums are warp reductions instead of a 64-term fmaf chain
        assert torch.allclose(mean, b.actions[0], rtol=0, atol=1e-5)
    model.close()


def test_gae(drl):
    from drone_rl_b200 import _lib
    import ctypes as C
    K, n = 37, 1234
    g = torch.Generator().manual_seed(2)
    rew = torch.randn(K, n, generator=g)
    val = torch.randn(K, n, generator=g)
    done = (torch.rand(K, n, generator=g) < 0.05)
    last = torch.randn(n, generator=g)
    d = lambda t: t.cuda().contiguous()
    rew_d, val_d, done_d, last_d = d(rew), d(val), d(done.to(torch.uint8)), d(last)
    adv_d, ret_d = torch.empty(K, n, device="cuda"), torch.empty(K, n, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(_lib.load().dronecu_gae(0, K, n, P(rew_d), P(val_d), P(done_d), P(last_d), 0.99, 0.95, P(adv_d), P(ret_d), None))
    adv, ret = po.gae(rew.double(), val.double(), done, last.double())
    _close(adv_d.cpu().numpy(), adv.numpy(), 2e-5, "advantage")
    _close(ret_d.cpu().numpy(), ret.numpy(), 2e-5, "returns")


def _fake_buffers(model, seed, spread=1.0):
    """Fill the model's rollout buffers with a synthetic but self-consistent batch."""
    g = torch.Generator().manual_seed(seed)
    b = model.buf
    K, n = b.logp.shape
    obs = torch.randn(K, n, 15, generator=g) * 2.0
    theta = model.params.cpu().double()
    mean, value, log_std = po.forward(theta, obs.double().reshape(-1, 15))
    act = mean + torch.exp(log_std) * torch.randn(K * n, 4, generator=g, dtype=torch.float64)
    # old log-prob from a slightly different policy so that ratios spread around 1 and some clip
    old_logp = po.log_prob(mean, log_std, act) + spread * 0.15 * torch.randn(K * n, generator=g, dtype=torch.float64)
    adv = torch.randn(K * n, generator=g, dtype=torch.float64) * 3 + 0.5
    ret = value + torch.randn(K * n, generator=g, dtype=torch.float64)
    b.obs.copy_(obs.cuda()); b.actions.copy_(act.float().reshape(K, n, 4).cuda())
    b.logp.copy_(old_logp.float().reshape(K, n).cuda()); b.adv.copy_(adv.float().reshape(K, n).cuda())
    b.ret.copy_(ret.float().reshape(K, n).cuda())
    torch.cuda.synchronize()
    f64 = lambda t: t.cpu().double()
    return (f64(b.obs).reshape(-1, 15), f64(b.actions).reshape(-1, 4), f64(b.logp).reshape(-1),
            f64(b.adv).reshape(-1), f64(b.ret).reshape(-1))


def _gpu_grad(model, index, m):
    import ctypes as C
    from drone_rl_b200 import _lib
    b = model.buf
    P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    model._adv_stats.zero_()
    _lib.check(model.lib.dronecu_ppo_adv_stats(model._h, P(b.adv), P(index), 0, m, P(model._adv_stats), None))
    _lib.check(model.lib.dronecu_ppo_grad(model._h, P(model.params), P(b.obs), P(b.actions), P(b.logp), P(b.adv),
                                          P(b.ret), P(index), 0, m, 0.0, 1.0, P(model._adv_stats), P(model._grad), None))
    torch.cuda.synchronize()
    return model._grad.cpu().double().numpy().copy()


@pytest.mark.parametrize("m", [64, 1000, 128 * 148 + 37])
def test_minibatch_gradient(drl, m):
    """Sum-form gradient of the SB3 loss over a gathered minibatch vs torch autograd in float64."""
    from drone_rl_b200.ppo import PPO
    model = PPO(512, n_steps=40, ent_coef=0.01)
    model.params.copy_(_rand_params(9, 0.5).float().cuda())
    obs, act, old_logp, adv, ret = _fake_buffers(model, 4)
    B = obs.shape[0]
    idx = torch.randperm(B, generator=torch.Generator().manual_seed(1))[:m]
    g = _gpu_grad(model, idx.to(torch.int32).cuda(), m)
    theta = model.params.cpu().double().requires_grad_(True)
    loss, stats = po.ppo_loss(theta, obs[idx], act[idx], old_logp[idx], adv[idx], ret[idx], ent_coef=0.01)
    (ref,) = torch.autograd.grad(loss, theta)
    ref = ref.numpy() * m                                   # the kernel returns the SUM over samples
    scale = np.abs(ref).max()
    err = np.abs(g[:po.N_PARAMS] - ref)
    assert err.max() <= 2e-4 * scale, f"worst {err.max() / scale:.3g} at {err.argmax()}"
    # per-block check so that a wrong small block cannot hide behind a large one
    for name, (off, shape) in po.offsets().items():
        n = int(np.prod(shape))
        blk_ref, blk = ref[off:off + n], g[off:off + n]
        assert np.abs(blk - blk_ref).max() <= 5e-4 * max(np.abs(blk_ref).max(), 1e-3 * scale), name
    st = g[po.N_PARAMS:]
    assert st[4] == m
    np.testing.assert_allclose(st[0] / m, stats["policy_gradient_loss"], rtol=2e-4, atol=1e-5)
    np.testing.assert_allclose(st[1] / m, stats["value_loss"], rtol=2e-4)
    np.testing.assert_allclose(st[2] / m, stats["approx_kl"], rtol=2e-3, atol=1e-6)
    np.testing.assert_allclose(st[3] / m, stats["clip_fraction"], atol=2.0 / m)
    assert 0.05 < stats["clip_fraction"] < 0.95             # both branches of the clip were exercised
    model.close()


def test_gradient_is_shard_additive_and_deterministic(drl):
    """Data-parallel contract: with GLOBAL advantage statistics the sum-form gradient of the whole
    minibatch equals the sum of the gradients of its shards (what the NCCL all-reduce forms)."""
    import ctypes as C
    from drone_rl_b200 import _lib
    from drone_rl_b200.ppo import PPO
    model = PPO(256, n_steps=32)
    model.params.copy_(_rand_params(2, 0.5).float().cuda())
    _fake_buffers(model, 8)
    B = 256 * 32
    b = model.buf
    P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    model._adv_stats.zero_()
    _lib.check(model.lib.dronecu_ppo_adv_stats(model._h, P(b.adv), None, 0, B, P(model._adv_stats), None))

    def grad(first, m):
        _lib.check(model.lib.dronecu_ppo_grad(model._h, P(model.params), P(b.obs), P(b.actions), P(b.logp), P(b.adv),
                                              P(b.ret), None, first, m, 0.0, 1.0, P(model._adv_stats), P(model._grad), None))
        torch.cuda.synchronize()
        return model._grad.cpu().double().numpy().copy()
    whole, again = grad(0, B), grad(0, B)
    assert np.array_equal(whole, again)                       # fixed-order reduction: bit-reproducible
    parts = grad(0, B // 2) + grad(B // 2, B - B // 2)
    scale = np.abs(whole[:po.N_PARAMS]).max()
    assert np.abs(parts - whole)[:po.N_PARAMS].max() <= 1e-5 * scale
    assert parts[po.N_PARAMS + 4] == whole[po.N_PARAMS + 4] == B
    model.close()


def test_sb3_default_update_matches_oracle(drl):
    """n_steps x n_envs = 512 samples, batch_size 64 (SB3 default), 2 epochs = 16 optimiser steps with
    the same minibatch indices on both sides: parameters after clip_grad_norm_ + Adam agree."""
    from drone_rl_b200.ppo import PPO
    model = PPO(8, n_steps=64, batch_size=64)
    theta0 = _rand_params(11, 0.3).float()
    model.params.copy_(theta0.cuda())
    obs, act, old_logp, adv, ret = _fake_buffers(model, 6, spread=0.3)
    B = obs.shape[0]
    theta, st = theta0.double(), po.AdamState(po.N_PARAMS, torch.float64)
    g = torch.Generator().manual_seed(3)
    for epoch in range(2):
        perm = torch.randperm(B, generator=g)
        for s in range(0, B, 64):
            idx = perm[s:s + 64]
            model._minibatch(idx.to(torch.int32).cuda(), 0, 64)
            theta, stats, _ = po.minibatch_update(theta, st, (obs[idx], act[idx], old_logp[idx], adv[idx], ret[idx]))
    torch.cuda.synchronize()
    got = model.params.cpu().double()
    assert (got - theta0.double()).abs().max() > 1e-3          # it moved ...
    assert (got - theta).abs().max() < 1e-5                    # ... to the same place
    info = model._info.cpu().numpy()
    np.testing.assert_allclose(info[8], stats["grad_norm"], rtol=1e-3)
    np.testing.assert_allclose(info[1], stats["value_loss"], rtol=1e-3)
    assert model.n_updates == 16
    model.close()


def test_learn_improves_and_checkpoints(drl, tmp_path):
    """End to end: rollout + GAE + update loop learns.  The reference's reward (drone.py:142-148) is
    -0.01*dist everywhere outside a 5 cm bonus ball, i.e. almost signal-free; the test widens the
    bonus ball to 1 m (a config knob, every other literal as the reference) so that staying aloft near
    the target pays and PPO must find it.  A checkpoint restores policy, Adam state, env curriculum
    counters and the RNG index."""
    from drone_rl_b200.ppo import PPO
    cfg = lambda: drl.EnvConfig.single(bonus_radius=1.0)
    m2 = PPO(drl.DroneBatch(2048, cfg(), seed=1), n_steps=64, seed=1)
    m2.collect_rollouts()
    first = m2.batch.episode_stats()
    model = PPO(drl.DroneBatch(2048, cfg(), seed=1), n_steps=64, batch_size=2048 * 64 // 4, n_epochs=4, seed=1,
                learning_rate=3e-3)
    model.learn(total_timesteps=2048 * 64 * 60)
    lv = model.logger_values
    assert np.isfinite(lv["train/value_loss"]) and np.isfinite(lv["rollout/ep_rew_mean"])
    print("untrained", first["ep_rew_mean"], first["ep_len_mean"], "trained", lv["rollout/ep_rew_mean"], lv["rollout/ep_len_mean"])
    assert lv["rollout/ep_rew_mean"] > 1.3 * first["ep_rew_mean"] > 0
    assert lv["rollout/ep_len_mean"] > 1.3 * first["ep_len_mean"]
    path = str(tmp_path / "ckpt.pt")
    model.save(path)
    m3 = PPO.load(path, drl.DroneBatch(2048, cfg(), seed=1), n_steps=64)
    assert torch.equal(m3.params, model.params) and m3.n_updates == model.n_updates
    s1, s3 = model.batch.get_state(), m3.batch.get_state()
    assert np.array_equal(s1["ep_num"], s3["ep_num"]) and np.array_equal(s1["pos"], s3["pos"])
    assert m3.batch.global_step == model.batch.global_step
    act, _ = m3.predict(np.zeros(15, np.float32), deterministic=True)
    assert act.shape == (4,) and (act >= 0).all() and (act <= 7.3575 + 1e-6).all()
    for m in (model, m2, m3):
        m.close()


@pytest.mark.parametrize("n", [1, 2, 63, 64, 4097, 65536 * 32, 3_000_001])
def test_minibatch_permutation_matches_oracle(drl, n):
    """The per-epoch minibatch order (stand-in for SB3's np.random.permutation): bit-exact against the numpy
    restatement, a true permutation at every size (power of two, ragged, one element), different per epoch."""
    import ctypes as C
    from drone_rl_b200 import _lib
    from oracle import philox
    lib = _lib.load()
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    seen = []
    for epoch in (0, 1, 2 ** 33 + 5):
        _lib.check(lib.dronecu_minibatch_permutation(0, n, 12345, epoch, C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
        got = out.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, philox.minibatch_permutation(n, 12345, epoch))
        assert np.array_equal(np.sort(got), np.arange(n))
        seen.append(got)
    if n > 64:
        assert (seen[0] == seen[1]).mean() < 0.01 and abs(np.corrcoef(np.arange(n), seen[0])[0, 1]) < 0.05


@pytest.mark.parametrize("prec", ["fp32", "bf16"])
def test_cuda_graph_epoch_equals_eager(drl, prec):
    """One epoch of minibatch steps replayed as a CUDA graph == the same launches issued one by one: bit-identical
    parameters, Adam moments and step count (the Adam step lives on the device so that the sequence is capturable)."""
    from drone_rl_b200.ppo import PPO
    out = []
    for graph in (False, True):
        m = PPO(drl.DroneBatch(64, drl.EnvConfig.single(), seed=5), n_steps=32, batch_size=64, n_epochs=3, seed=5,
                update_precision=prec, cuda_graph=graph)
        for _ in range(2):
            m.collect_rollouts()
            m.train()
        torch.cuda.synchronize()
        sd = m.state_dict()
        assert m.n_updates == 2 * 3 * 32 and sd["adam_step"] == m.n_updates
        assert (m._graph is not None) == graph
        out.append((m.params.clone(), sd["adam"].clone(), dict(m.logger_values)))
        m.close()
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert out[0][2]["train/value_loss"] == out[1][2]["train/value_loss"]


def test_epoch_advantage_statistics(drl):
    """The statistics of every minibatch of an epoch in one call == the per-minibatch entry point == numpy, incl. a
    ragged last minibatch; count column exact."""
    import ctypes as C
    from drone_rl_b200 import _lib
    from drone_rl_b200.ppo import PPO
    model = PPO(drl.DroneBatch(96, drl.EnvConfig.single(), seed=2), n_steps=21, seed=2)
    B, bs = 96 * 21, 500                                   # 4 full minibatches + one of 16
    torch.manual_seed(5)
    adv = torch.randn(B, device="cuda") * 3 + 0.5
    perm = torch.randperm(B, device="cuda").to(torch.int32)
    n_mb = (B + bs - 1) // bs
    out = torch.zeros(n_mb, 3, dtype=torch.float64, device="cuda")
    P = lambda t: C.c_void_p(t.data_ptr())
    _lib.check(model.lib.dronecu_ppo_adv_stats_epoch(model._h, P(adv), P(perm), B, bs, P(out), None))
    one = torch.zeros(3, dtype=torch.float64, device="cuda")
    for k in range(n_mb):
        m = min(bs, B - k * bs)
        one.zero_()
        _lib.check(model.lib.dronecu_ppo_adv_stats(model._h, P(adv), P(perm[k * bs:]), 0, m, P(one), None))
        a = adv[perm[k * bs:k * bs + m].long()].double().cpu().numpy()
        got = out[k].cpu().numpy()
        assert got[2] == m == one[2].item()
        np.testing.assert_allclose(got[:2], [a.sum(), (a * a).sum()], rtol=1e-12)
        np.testing.assert_allclose(got[:2], one[:2].cpu().numpy(), rtol=1e-13)
    model.close()


@pytest.mark.parametrize("n,batch", [(64, 64), (2048, 64), (4097, 1000), (65536 * 32, 65536 * 8), (3_000_001, 46_876), (1 << 20, 1 << 14)])
def test_minibatch_partition_matches_oracle(drl, n, batch):
    """dronecu_minibatch_partition: the uniformly random partition "row r -> minibatch f(r) // batch" with every minibatch's
    rows in ascending order == a stable sort of the oracle's keyed permutation by minibatch id; a permutation of 0..n-1."""
    import ctypes as C
    from drone_rl_b200 import _lib
    lib = _lib.load()
    cfg = _lib.PPOConfig()
    lib.dronecu_ppo_config_default(C.byref(cfg))
    h = C.c_void_p()
    _lib.check(lib.dronecu_ppo_create(C.byref(cfg), 0, C.byref(h)))
    out = torch.empty(n, dtype=torch.int32, device="cuda")
    n_mb = (n + batch - 1) // batch
    for epoch in (0, 7):
        _lib.check(lib.dronecu_minibatch_partition(h, n, batch, 4242, epoch, C.c_void_p(out.data_ptr()), None))
        torch.cuda.synchronize()
        got = out.cpu().numpy().astype(np.int64)
        assert np.array_equal(got, philox.minibatch_partition(n, batch, 4242, epoch))
        assert np.array_equal(np.sort(got), np.arange(n))
        for b in range(min(n_mb, 8)):
            seg = got[b * batch:(b + 1) * batch]
            assert (np.diff(seg) > 0).all() and len(seg) == min(batch, n - b * batch)
    assert lib.dronecu_minibatch_partition(h, 6500, 100, 1, 0, C.c_void_p(out.data_ptr()), None) != 0      # 65 minibatches: unsupported
    lib.dronecu_ppo_destroy(h)


def test_learn_counts_timesteps_like_sb3(drl):
    """SB3's learn(total_timesteps) trains that many MORE steps: reset_num_timesteps=True (default) restarts the counter, so
    PPO.load(...).learn(T) after a resume (reference train.py:22-30, :63-68) does not return at once; False continues it."""
    from drone_rl_b200.ppo import PPO
    model = PPO(drl.DroneBatch(64, drl.EnvConfig.single(), seed=2), n_steps=16, batch_size=256, n_epochs=2, seed=2)
    per_iter = 64 * 16
    model.learn(3 * per_iter)
    assert model.num_timesteps == 3 * per_iter and model.n_updates == 3 * 2 * 4
    model.learn(2 * per_iter)                                   # counter restarts; two more iterations
    assert model.num_timesteps == 2 * per_iter and model.n_updates == 5 * 2 * 4
    model.learn(per_iter, reset_num_timesteps=False)            # continues: one more iteration on top
    assert model.num_timesteps == 3 * per_iter and model.n_updates == 6 * 2 * 4
    assert model.logger_values["time/total_timesteps"] == 3 * per_iter
    model.close()


def test_checkpoint_env_state_only_onto_the_same_shard(drl, tmp_path):
    """The env / curriculum / Philox state of an archive belongs to ONE shard of global env ids (its env_offset and world size
    travel with it): a handle over other env ids starts fresh instead of becoming a copy of shard 0; policy and Adam load
    either way."""
    from drone_rl_b200.ppo import PPO
    n = 128
    a = PPO(drl.DroneBatch(n, drl.EnvConfig.single(), seed=4, env_offset=0), n_steps=8, batch_size=256, n_epochs=1, seed=4)
    a.learn(4 * n * 8)
    path = str(tmp_path / "shard0.zip")
    a.save(path)
    same = PPO.load(path, drl.DroneBatch(n, drl.EnvConfig.single(), seed=4, env_offset=0), n_steps=8, batch_size=256, n_epochs=1)
    other = PPO.load(path, drl.DroneBatch(n, drl.EnvConfig.single(), seed=4, env_offset=n), n_steps=8, batch_size=256, n_epochs=1)
    assert same.env_state_restored and not other.env_state_restored
    assert torch.equal(same.params, a.params) and torch.equal(other.params, a.params)
    assert np.array_equal(same.batch.get_state("ep_num")["ep_num"], a.batch.get_state("ep_num")["ep_num"])
    assert (other.batch.get_state("ep_num")["ep_num"] == 2).all()          # fresh: constructor reset + PPO's reset
    for m in (a, same, other):
        m.close()


def test_train_logs_the_mean_over_all_minibatches(drl):
    """SB3 logs np.mean over every minibatch of every epoch for policy_gradient_loss / value_loss / approx_kl / clip_fraction
    and the LAST minibatch's total loss; the device-side accumulator must give exactly that."""
    from drone_rl_b200.ppo import PPO
    model = PPO(drl.DroneBatch(512, drl.EnvConfig.single(), seed=6), n_steps=16, batch_size=1024, n_epochs=3, seed=6, cuda_graph=False)
    model.collect_rollouts()
    seen, orig = [], model._minibatch

    def spy(index, first, m, stats=None):
        orig(index, first, m, stats)
        seen.append(model._info.cpu().numpy().copy())
    model._minibatch = spy
    model.train()
    seen = np.array(seen)
    assert seen.shape[0] == 3 * 8
    lv = model.logger_values
    for k, col in (("train/policy_gradient_loss", 0), ("train/value_loss", 1), ("train/approx_kl", 2), ("train/clip_fraction", 3), ("train/grad_norm", 8)):
        np.testing.assert_allclose(lv[k], seen[:, col].mean(), rtol=2e-6, atol=1e-9, err_msg=k)
    ent = -float((0.5 + 0.5 * np.log(2 * np.pi) + model.params[-4:].cpu().numpy().astype(np.float64)).sum())
    np.testing.assert_allclose(lv["train/entropy_loss"], ent, rtol=1e-6)
    np.testing.assert_allclose(lv["train/loss"], seen[-1, 0] + 0.5 * seen[-1, 1] + 0.0 * ent, rtol=1e-6, atol=1e-9)
    b = model.buf
    y, v = b.ret.cpu().double().numpy().ravel(), b.value.cpu().double().numpy().ravel()
    np.testing.assert_allclose(lv["train/explained_variance"], 1 - np.var(y - v) / np.var(y), rtol=1e-4, atol=1e-5)
    model.close()
