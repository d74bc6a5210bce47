"""TEST INFRASTRUCTURE: restatement of SB3's DummyVecEnv + VecMonitor semantics.

PARITY UNPINNED: the code being restated is stable-baselines3
(``stable_baselines3.common.vec_env.DummyVecEnv`` / ``VecMonitor``; PyPI, un-pinned,
reference environment.yaml:16), which is absent from /root/reference and not installed.
The reference only *calls* it (train.py:18-20, :33-35); it holds no test of it.  The
behaviour restated here is SB3's published one:

* ``step``: for each env ``obs, r, done, info = env.step(a[i])``; if done:
  ``info["terminal_observation"] = obs; obs = env.reset()``.  Buffers: obs float32,
  rewards float32, dones bool.
* ``VecMonitor``: float32 running return and int length per env; on done attaches
  ``info["episode"] = {"r", "l"}`` and zeroes them.

It drives any object with the legacy gym API -- in particular the *real* reference
``DroneGymEnv`` (via oracle/ref_import.py), which is how BatchedDroneOracle's auto-reset
path is validated.
"""
from __future__ import annotations

import numpy as np


class DummyVecEnvOracle:
    def __init__(self, envs):
        self.envs = list(envs)
        self.num_envs = len(self.envs)

    def reset(self):
        return np.stack([np.asarray(e.reset(), dtype=np.float32) for e in self.envs])

    def step(self, actions):
        obs_buf, rew_buf, done_buf, infos = [], [], [], []
        for env, act in zip(self.envs, actions):
            obs, rew, done, info = env.step(act)
            info = dict(info)
            if done:
                info["terminal_observation"] = obs
                obs = env.reset()
            obs_buf.append(np.asarray(obs, dtype=np.float32))
            rew_buf.append(rew)
            done_buf.append(bool(done))
            infos.append(info)
        return (np.stack(obs_buf), np.asarray(rew_buf, dtype=np.float32),
                np.asarray(done_buf, dtype=bool), infos)


class VecMonitorOracle:
    def __init__(self, venv):
        self.venv = venv
        self.num_envs = venv.num_envs

    def reset(self):
        obs = self.venv.reset()
        self.episode_returns = np.zeros(self.num_envs, dtype=np.float32)
        self.episode_lengths = np.zeros(self.num_envs, dtype=np.int32)
        return obs

    def step(self, actions):
        obs, rews, dones, infos = self.venv.step(actions)
        self.episode_returns += rews
        self.episode_lengths += 1
        for i in np.flatnonzero(dones):
            infos[i] = dict(infos[i])
            infos[i]["episode"] = {"r": float(self.episode_returns[i]),
                                   "l": int(self.episode_lengths[i])}
            self.episode_returns[i] = 0
            self.episode_lengths[i] = 0
        return obs, rews, dones, infos
