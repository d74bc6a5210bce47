"""Parity at the sizes BASELINE.json's configs name ("configs not tested at their own size").

* configs[1]: VectorizedDroneEnv spec, 4096 envs x 1000 steps, random actions, ONE launch -- every one of the
  4,096,000 transitions checked teacher-forced against the float64 oracle (oracle/verify.py), incl. the shared
  1000-step time limit (vectorized_drone.py:200, :211-213) and envs that crashed and keep integrating (no reset
  logic in the reference's vectorized env).
* configs[3]: one GPU's shard of the 64M-env run -- 8,388,608 envs x 32 fused steps, DroneGymEnv spec, in-kernel
  Philox actions, the LAST shard's global env ids (7 x 8,388,608 ...): a strided sample of ~4k envs, every
  transition of theirs checked (state / obs / reward within 1e-5 relative, done / reset obs / actions bit-exact).
Tolerance: |gpu - ref| <= 1e-5 * max(|ref|, 1) (oracle/verify.py TOL_REL); the worst PURE relative error is printed
and returned next to it so the unit floor is visible.
"""
import json
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import drone_oracle as do  # noqa: E402
from oracle import philox, verify  # noqa: E402

np.seterr(all="ignore")
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


@pytest.fixture(scope="module")
def drl():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import drone_rl_b200
    return drone_rl_b200


def _save(name, rep):
    try:
        os.makedirs(OUT, exist_ok=True)
        json.dump(rep, open(os.path.join(OUT, f"parity_{name}.json"), "w"), indent=1)
    except OSError:
        pass


def test_config1_vector_4096_envs_1000_steps_every_transition(drl):
    n, K = 4096, 1000
    rng = np.random.default_rng(2)
    acts = rng.uniform(0, do.MOTOR_MAX, (K, n, 4)).astype(np.float32)
    b = drl.DroneBatch(n, drl.EnvConfig.vector(), seed=0)
    obs0 = b.reset()
    nxt = b.empty(K, n, 12); rew = b.empty(K, n); done = b.empty(K, n, dtype=torch.uint8)
    b.rollout(K, torch.from_numpy(acts).cuda(), next_obs=nxt, reward=rew, done=done)      # ONE launch, 1000 fused steps
    torch.cuda.synchronize()
    d = done.cpu().numpy().astype(bool)
    rep = verify.check_rollout(obs0.cpu().numpy(), acts, nxt.cpu().numpy(), rew.cpu().numpy(), d, spec=do.VECTOR)
    assert rep["transitions"] == n * K
    assert d[K - 1].all(), "the shared step counter reaches max_steps = 1000: every env reports done (vectorized_drone.py:211-213)"
    first_all = int(np.argmax(d.all(axis=1)))
    # no reset logic in the reference's vectorized env: a crashed env (z < 0) keeps integrating and keeps reporting done
    # (vectorized_drone.py:211), so under random actions the whole batch is "done" long before the time limit
    assert (d[1:] >= d[:-1])[:, :].mean() > 0.999 and 20 < first_all < K - 1
    assert rep["borderline_done"] <= 3 * n * K // 100000 + 3
    print("configs[1] 4096 x 1000:", rep)
    _save("c2_4096x1000", rep)
    b.close()


def test_config3_shard_8M_envs_32_steps_strided_sample(drl):
    n, K, seed = 8_388_608, 32, 0
    off = 7 * n                                   # the 8th GPU's shard of the 64M-env run
    b = drl.DroneBatch(n, drl.EnvConfig.single(), seed=seed, env_offset=off)
    obs0 = b.reset()
    nxt = b.empty(K, n, 15); act = b.empty(K, n, 4); rew = b.empty(K, n); done = b.empty(K, n, dtype=torch.uint8)
    b.rollout(K, None, next_obs=nxt, out_actions=act, reward=rew, done=done)
    torch.cuda.synchronize()
    st = b.episode_stats()
    assert st["env_steps"] == n * K and st["episodes"] == int(done.sum(dtype=torch.int64).item()) > n // 4
    sel = torch.cat([torch.arange(0, n, 2053, device="cuda"), torch.tensor([n - 1], device="cuda")])
    ids = (sel.cpu().numpy().astype(np.uint64) + np.uint64(off))
    a = act[:, sel].cpu().numpy()
    for k in range(K):                             # random-policy actions: bit-exact Philox, keyed by the GLOBAL env id
        assert np.array_equal(a[k], philox.action_uniforms(seed, ids, k).astype(np.float32) * np.float32(7.3575)), k
    rep = verify.check_rollout(obs0[sel].cpu().numpy(), a, nxt[:, sel].cpu().numpy(), rew[:, sel].cpu().numpy(),
                               done[:, sel].cpu().numpy().astype(bool), spec=do.SINGLE, seed=seed, env_ids=ids)
    assert rep["dones"] > 1000 and rep["borderline_done"] <= 3
    print("configs[3] shard 8,388,608 x 32 (sample of %d envs):" % sel.numel(), rep)
    _save("c4_8Mx32_sample", rep)
    b.close()
