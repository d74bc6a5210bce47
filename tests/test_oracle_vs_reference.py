"""Live check of the numpy oracle against the UNMODIFIED reference modules.  Only runs where
/root/reference exists (the build container); on the GPU box the committed golden vectors
(tests/test_oracle_golden.py) carry the same pins."""
import numpy as np
import pytest

from oracle import drone_oracle as do
from oracle import philox, ref_import
from oracle.vecenv_oracle import DummyVecEnvOracle, VecMonitorOracle

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="/root/reference not present")
np.seterr(all="ignore")


def test_vectorized_env_bit_exact_fresh_seed():
    _, vd = ref_import.load()
    B = 256
    ref, orc = vd.VectorizedDroneEnv(B), do.BatchedDroneOracle(B, do.VECTOR)
    assert np.array_equal(ref.reset(), orc.reset())
    rng = np.random.default_rng(2024)
    for t in range(120):
        a = rng.uniform(0, do.MOTOR_MAX, (B, 4)).astype(np.float32)
        o1, r1, d1, _ = ref.step(a.astype(np.float64))
        o2, r2, d2, _ = orc.step(a)
        assert np.array_equal(o1, o2, equal_nan=True) and np.array_equal(r1, r2, equal_nan=True)
        assert np.array_equal(d1, d2)
        assert r1.dtype == np.float64 and o1.dtype == np.float32 and d1.dtype == np.bool_


def test_gym_env_with_autoreset_fresh_seed():
    drone, _ = ref_import.load()
    from tests.golden.make_golden import PhiloxFedEnv

    n, seed, off = 4, 31337, 5_000_000_000          # env ids beyond 2**32
    stream = ref_import.UniformStream()
    with ref_import.patched_rand(stream):
        envs = [PhiloxFedEnv(drone, stream, seed, off + i) for i in range(n)]
        venv = VecMonitorOracle(DummyVecEnvOracle(envs))
        obs_ref = venv.reset()
        orc = do.BatchedDroneOracle(n, do.SINGLE, seed=seed, env_offset=off)
        assert np.array_equal(orc.reset(), obs_ref)
        rng = np.random.default_rng(1)
        n_done = 0
        for t in range(150):
            a = rng.uniform(0, do.MOTOR_MAX, (n, 4)).astype(np.float32)
            o1, r1, d1, infos = venv.step(a.astype(np.float64))
            o2, r2, d2, info = orc.step(a)
            assert np.array_equal(d1, d2)
            np.testing.assert_allclose(o2, o1, rtol=2e-5, atol=1e-6)
            np.testing.assert_allclose(r2.astype(np.float32), r1, rtol=1e-6, atol=1e-9)
            for i in np.flatnonzero(d1):
                n_done += 1
                assert infos[i]["episode"]["l"] == info["episode_l"][i]
                np.testing.assert_allclose(info["terminal_obs"][i], infos[i]["terminal_observation"],
                                           rtol=2e-5, atol=1e-6)
        assert n_done > 5 and stream.queue == []
    assert stream.drawn == 5 * sum(e.env.ep_num for e in envs)   # exactly 5 draws per reset
