"""Pin the numpy oracle against vectors generated from the unmodified reference
(tests/golden/make_golden.py).  Runs anywhere -- no /root/reference needed."""
import numpy as np
import pytest

from oracle import drone_oracle as do

np.seterr(all="ignore")


def test_vector_rollout_bit_exact(golden):
    """VectorizedDroneEnv.step, 1003 steps crossing the shared 1000-step limit
    (vectorized_drone.py:135-216): bit-identical obs / reward / done."""
    g = golden("vector_rollout")
    env = do.BatchedDroneOracle(g["actions"].shape[1], do.VECTOR)
    assert np.array_equal(env.reset(), g["obs0"])
    for t in range(g["actions"].shape[0]):
        obs, rew, done, _ = env.step(g["actions"][t])
        assert np.array_equal(obs, g["obs"][t], equal_nan=True), t
        assert np.array_equal(rew, g["reward"][t], equal_nan=True), t
        assert np.array_equal(done, g["done"][t]), t
    assert g["done"][998].mean() < 1.0 and g["done"][999].all() and g["done"][1002].all()


def test_teacher_forced_vector_bit_exact(golden):
    g = golden("teacher_forced")
    n = g["pos"].shape[0]
    env = do.BatchedDroneOracle(n, do.VECTOR)
    env.set_state(g["pos"], g["vel"], g["euler"], g["omega"])
    obs, rew, done, _ = env.step(g["action"])
    for name in ("pos", "vel", "euler", "omega"):
        assert np.array_equal(getattr(env, name), g["vec_" + name], equal_nan=True), name
    assert np.array_equal(obs, g["vec_obs"], equal_nan=True)
    assert np.array_equal(rew, g["vec_reward"], equal_nan=True)
    assert np.array_equal(done, g["vec_done"])


def test_teacher_forced_single(golden):
    """DroneEnv.step (scalar numpy) vs the batched oracle: integer results exact, float
    results to a few float64 ulp (numpy's scalar vs SIMD trig differ in the last bits)."""
    g = golden("teacher_forced")
    n = g["pos"].shape[0]
    env = do.BatchedDroneOracle(n, do.SINGLE)
    env.set_state(g["pos"], g["vel"], g["euler"], g["omega"], g["target"], step_count=g["step_count"])
    spec = do.SINGLE
    # no auto-reset for the comparison: look at the terminal observation
    obs, rew, done, info = env.step(g["action"])
    assert np.array_equal(done, g["single_done"])
    ref = g["single_state"]
    got_obs = info["terminal_obs"]
    finite = np.isfinite(ref).all(1)
    assert np.array_equal(np.isnan(got_obs[:, :12]), np.isnan(ref.astype(np.float32)))
    np.testing.assert_allclose(got_obs[finite], g["single_obs"][finite], rtol=2e-7, atol=0)
    # identical float32 observation for the overwhelming majority (ulp-level f64 differences
    # only matter when they straddle a float32 rounding boundary)
    same = (got_obs[finite] == g["single_obs"][finite]).all(1).mean()
    assert same > 0.99
    # pitch within ~1e-7 of pi/2 makes sec(pitch) ~ 2e7-conditioned: leave those rows to the
    # bit-exact vectorized comparison above
    well = finite & (np.abs(np.cos(g["euler"][:, 1].astype(np.float64))) > 1e-4)
    np.testing.assert_allclose(rew[well], g["single_reward"][well], rtol=1e-12)
    assert spec.max_steps == 200 and done[g["step_count"] == 199].all()


def test_const_action_demo(golden):
    """The reference's own demo (drone.py:288-294): 2x hover thrust, straight up, done by
    |pos| > 50 at step 158."""
    g = golden("single_const_action")
    env = do.BatchedDroneOracle(1, do.SINGLE, seed=int(g["seed"]), env_offset=int(g["env_id"]))
    obs0 = env.reset()
    assert np.array_equal(obs0[0], g["obs0"])
    T = g["obs"].shape[0]
    assert T == 158
    act = np.tile(g["action"], (1, 1))
    for t in range(T):
        obs, rew, done, info = env.step(act)
        np.testing.assert_allclose(info["terminal_obs"][0], g["obs"][t], rtol=1e-6, atol=1e-12)
        np.testing.assert_allclose(rew[0], g["reward"][t], rtol=1e-12, atol=1e-12)
        assert bool(done[0]) == bool(g["done"][t])
    assert done[0] and info["terminated"][0] and not info["truncated"][0]


def test_single_rollout_autoreset(golden):
    """8 reference DroneGymEnv under the DummyVecEnv/VecMonitor restatement, 400 random-action
    steps with Philox-fed resets: done / episode length / ep_num exact, floats to 1e-9."""
    g = golden("single_rollout")
    T, n = g["actions"].shape[:2]
    env = do.BatchedDroneOracle(n, do.SINGLE, seed=int(g["seed"]), env_offset=int(g["env_offset"]))
    obs = env.reset()            # VecEnv.reset(): second reset of each env (ep_num 2)
    np.testing.assert_allclose(obs, g["obs0"], rtol=1e-6)
    assert np.array_equal(obs, g["obs0"])
    for t in range(T):
        obs, rew, done, info = env.step(g["actions"][t])
        assert np.array_equal(done, g["done"][t]), t
        np.testing.assert_allclose(obs, g["obs"][t], rtol=2e-5, atol=1e-6, err_msg=str(t))
        np.testing.assert_allclose(rew.astype(np.float32), g["reward"][t], rtol=1e-6, atol=1e-9)
        d = done
        if d.any():
            np.testing.assert_allclose(info["terminal_obs"][d], g["terminal_obs"][t][d], rtol=2e-5, atol=1e-6)
            assert np.array_equal(info["episode_l"][d], g["episode_l"][t][d])
            np.testing.assert_allclose(info["episode_r"][d], g["episode_r"][t][d], rtol=1e-5)
    assert np.array_equal(env.ep_num, g["final_ep_num"])
    assert (g["episode_l"] == 200).sum() >= 1       # the time-limit path was exercised


def test_curriculum_schedule(golden):
    """eps bumps by 0.1 (float64 accumulation) when ep_num hits a multiple of 2000, and the
    five uniforms are consumed in the order pos.x pos.y tgt.x tgt.y tgt.z (drone.py:57-73)."""
    g = golden("curriculum")
    eps = do.curriculum_eps(g["ep_num"])
    assert np.array_equal(eps, g["eps"])
    assert eps[list(g["ep_num"]).index(6000)] == 0.30000000000000004
    env = do.BatchedDroneOracle(1, do.SINGLE, seed=int(g["seed"]), env_offset=int(g["env_id"]))
    for k, ep in enumerate(g["ep_num"]):
        env.ep_num[:] = ep - 1
        env.reset()
        assert np.array_equal(env.pos[0], g["pos"][k])
        assert np.array_equal(env.target[0], g["target"][k])
