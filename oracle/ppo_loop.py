"""TEST INFRASTRUCTURE: the reference's training loop restated on the CPU -- ``PPO("MlpPolicy",
VecMonitor(DummyVecEnv([DroneGymEnv])), device="cpu").learn(...)`` (reference train.py:33-43, :63-68).

PARITY UNPINNED (see oracle/ppo_oracle.py): stable-baselines3 is an un-pinned PyPI dependency that is
neither in /root/reference nor installed; this file strings the restated pieces together the way SB3's
``OnPolicyAlgorithm.collect_rollouts`` / ``PPO.train`` do:

  rollout   for n_steps: obs -> policy (torch, CPU) -> sample -> clip to the Box for the env only ->
            env.step (the float64 numpy env of oracle/drone_oracle.py with DummyVecEnv auto-reset and
            VecMonitor accounting) -> buffer (obs, unclipped action, reward f32, done, value, log-prob)
  GAE       gamma 0.99, lambda 0.95, bootstrap from the value of the last observation
  update    n_epochs x random minibatches of batch_size: normalised advantages, clipped surrogate,
            0.5 x value MSE, global-norm clip 0.5, Adam(3e-4, eps 1e-5)

Used by bench.py's CPU legs (``cpu_baseline`` of the PPO workloads and ``--impl reference``) and by the
tests; never by the product.
"""
from __future__ import annotations

import time

import numpy as np
import torch

from . import drone_oracle as do
from . import ppo_oracle as po


class PPOLoopOracle:
    def __init__(self, n_envs=1, n_steps=2048, batch_size=64, n_epochs=10, seed=0, spec=do.SINGLE,
                 dtype=torch.float32, learning_rate=3e-4):
        self.n, self.K, self.batch_size, self.n_epochs = n_envs, n_steps, batch_size, n_epochs
        self.dtype, self.lr = dtype, learning_rate
        self.env = do.BatchedDroneOracle(n_envs, spec, seed=seed)
        self.theta = po.init_params(seed, dtype)
        self.adam = po.AdamState(po.N_PARAMS, dtype)
        self.gen = torch.Generator().manual_seed(seed + 1)
        self.obs = self.env.reset()                       # SB3 _setup_learn: env.reset()
        self.num_timesteps = 0
        self.ep_returns, self.ep_lengths = [], []
        self.stats = {}

    def collect_rollouts(self):
        K, n = self.K, self.n
        obs_b = torch.empty(K, n, po.OBS, dtype=self.dtype)
        act_b = torch.empty(K, n, po.ACT, dtype=self.dtype)
        logp_b = torch.empty(K, n, dtype=self.dtype)
        val_b = torch.empty(K, n, dtype=self.dtype)
        rew_b = torch.empty(K, n, dtype=self.dtype)
        done_b = torch.empty(K, n, dtype=torch.bool)
        with torch.no_grad():
            for t in range(K):
                o = torch.from_numpy(self.obs).to(self.dtype)
                mean, value, log_std = po.forward(self.theta, o)
                a = mean + torch.exp(log_std) * torch.randn(mean.shape, generator=self.gen, dtype=self.dtype)
                logp = po.log_prob(mean, log_std, a)
                clipped = np.clip(a.numpy().astype(np.float32), 0.0, do.MOTOR_MAX)
                self.obs, rew, done, info = self.env.step(clipped.astype(np.float64))
                obs_b[t], act_b[t], logp_b[t], val_b[t] = o, a, logp, value
                rew_b[t] = torch.from_numpy(rew.astype(np.float32)).to(self.dtype)
                done_b[t] = torch.from_numpy(done)
                for i in np.flatnonzero(done):
                    self.ep_returns.append(float(info["episode_r"][i]))
                    self.ep_lengths.append(int(info["episode_l"][i]))
            _, last_value, _ = po.forward(self.theta, torch.from_numpy(self.obs).to(self.dtype))
            adv, ret = po.gae(rew_b, val_b, done_b, last_value)
        self.num_timesteps += K * n
        self.buf = tuple(x.reshape(K * n, *x.shape[2:]) for x in (obs_b, act_b, logp_b, adv, ret))

    def train(self):
        B = self.K * self.n
        for _ in range(self.n_epochs):
            perm = torch.randperm(B, generator=self.gen)
            for s in range(0, B, self.batch_size):
                idx = perm[s:s + self.batch_size]
                self.theta, self.stats, _ = po.minibatch_update(self.theta, self.adam, tuple(x[idx] for x in self.buf))

    def learn(self, total_timesteps):
        while self.num_timesteps < total_timesteps:
            self.collect_rollouts()
            self.train()
        return self

    def ep_rew_mean(self, last=100):
        return float(np.mean(self.ep_returns[-last:])) if self.ep_returns else float("nan")


def time_loop(n_envs, n_steps, batch_size, n_epochs, budget_s=10.0, update=True, seed=0):
    """env-steps/s of whole iterations (rollout [+ update]) inside ~budget_s seconds of CPU work."""
    loop = PPOLoopOracle(n_envs, n_steps, batch_size, n_epochs, seed=seed)
    t0, its = time.perf_counter(), 0
    while True:
        loop.collect_rollouts()
        if update:
            loop.train()
        its += 1
        dt = time.perf_counter() - t0
        if dt > budget_s:
            break
    return its * n_envs * n_steps / dt, its, dt
