"""PPO on the GPU: the rollout + update hot path of ``PPO("MlpPolicy", env)`` (reference
train.py:36-43, :63-68) with SB3's defaults, computed by the CUDA kernels behind the C ABI.

Host side only orchestrates: ONE launch collects ``n_steps`` env steps for every env with the policy
evaluated in-kernel (``dronecu_rollout_policy``), one computes GAE, and each minibatch is
adv-stats -> grad -> [NCCL all-reduce] -> clip+Adam.  torch supplies device memory, the random
permutation and ``torch.distributed``; no torch op touches the numerics.

Data parallel (``torch.distributed`` initialised, one process per GPU): every rank owns a contiguous
shard of global env ids and its own rollout buffers; per optimiser step the flat gradient (10,705
float32 incl. statistics) and the three advantage sums are all-reduced, so the update equals the
single-GPU update over the union of the shards.

PARITY UNPINNED: SB3 is not in the reference tree; tests compare against oracle/ppo_oracle.py.
"""
from __future__ import annotations

import ctypes as C
import math
import time
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import GRAD_LEN, POLICY_PARAMS, PolicyOut, PPOConfig
from .core import DroneBatch, EnvConfig, _ptr, _stream_ptr

_SHAPES = [("pi.W1", (64, 15)), ("pi.b1", (64,)), ("pi.W2", (64, 64)), ("pi.b2", (64,)), ("pi.W3", (4, 64)),
           ("pi.b3", (4,)), ("vf.W1", (64, 15)), ("vf.b1", (64,)), ("vf.W2", (64, 64)), ("vf.b2", (64,)),
           ("vf.W3", (1, 64)), ("vf.b3", (1,)), ("log_std", (4,))]
# names of the same tensors in an SB3 ActorCriticPolicy state_dict (for zip import / export)
SB3_NAMES = {"pi.W1": "mlp_extractor.policy_net.0.weight", "pi.b1": "mlp_extractor.policy_net.0.bias",
             "pi.W2": "mlp_extractor.policy_net.2.weight", "pi.b2": "mlp_extractor.policy_net.2.bias",
             "pi.W3": "action_net.weight", "pi.b3": "action_net.bias",
             "vf.W1": "mlp_extractor.value_net.0.weight", "vf.b1": "mlp_extractor.value_net.0.bias",
             "vf.W2": "mlp_extractor.value_net.2.weight", "vf.b2": "mlp_extractor.value_net.2.bias",
             "vf.W3": "value_net.weight", "vf.b3": "value_net.bias", "log_std": "log_std"}


def init_policy_params(seed: int = 0) -> torch.Tensor:
    """SB3 ActorCriticPolicy initialisation: orthogonal weights (gain sqrt2 towers, 0.01 action head,
    1 value head), zero biases, log_std = 0.  Flat float32 CPU vector [10697]."""
    g = torch.Generator().manual_seed(seed)
    gains = {"pi.W1": math.sqrt(2), "pi.W2": math.sqrt(2), "pi.W3": 0.01,
             "vf.W1": math.sqrt(2), "vf.W2": math.sqrt(2), "vf.W3": 1.0}
    flat, off = torch.zeros(POLICY_PARAMS, dtype=torch.float64), 0
    for name, shape in _SHAPES:
        n = int(np.prod(shape))
        if name in gains:
            rows, cols = shape
            a = torch.randn((max(rows, cols), min(rows, cols)), generator=g, dtype=torch.float64)
            q, r = torch.linalg.qr(a)
            q = q * torch.sign(torch.diag(r))
            if rows < cols:
                q = q.t()
            flat[off:off + n] = (q[:rows, :cols] * gains[name]).reshape(-1)
        off += n
    return flat.to(torch.float32)


def unpack_params(flat: torch.Tensor) -> dict:
    out, off = {}, 0
    for name, shape in _SHAPES:
        n = int(np.prod(shape))
        out[name] = flat[off:off + n].reshape(shape)
        off += n
    return out


class RolloutBuffers:
    """Device-resident rollout buffers [K, n, ...] (SB3 RolloutBuffer)."""

    def __init__(self, K, n, device):
        f = dict(dtype=torch.float32, device=device)
        self.obs = torch.empty(K, n, 15, **f)
        self.actions = torch.empty(K, n, 4, **f)
        self.logp = torch.empty(K, n, **f)
        self.value = torch.empty(K, n, **f)
        self.reward = torch.empty(K, n, **f)
        self.done = torch.empty(K, n, dtype=torch.uint8, device=device)
        self.last_value = torch.empty(n, **f)
        self.adv = torch.empty(K, n, **f)
        self.ret = torch.empty(K, n, **f)


class PPO:
    """SB3-shaped trainer: ``PPO(env).learn(total_timesteps)``, ``predict``, ``save`` / ``load``.

    ``env`` is a ``DroneBatch`` (or anything with a ``.batch`` DroneBatch, e.g. ``DroneVecEnv``), or an
    int = number of envs to create.  Defaults are SB3's (n_steps 2048, batch_size 64, n_epochs 10,
    gamma 0.99, gae_lambda 0.95, clip 0.2, ent 0, vf 0.5, max_grad_norm 0.5, lr 3e-4, Adam eps 1e-5).
    ``batch_size`` counts samples PER RANK.
    """

    def __init__(self, env=1, n_steps: int = 2048, batch_size: int = 64, n_epochs: int = 10, gamma: float = 0.99,
                 gae_lambda: float = 0.95, clip_range: float = 0.2, ent_coef: float = 0.0, vf_coef: float = 0.5,
                 max_grad_norm: float = 0.5, learning_rate: float = 3e-4, normalize_advantage: bool = True,
                 seed: int = 0, device: int = 0, verbose: int = 0, policy_seed: Optional[int] = None,
                 rollout_precision: str = "fp32", update_precision: str = "fp32", cuda_graph: Optional[bool] = None):
        self.lib = _lib.load()
        self.rank, self.world = 0, 1
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            self.rank, self.world = torch.distributed.get_rank(), torch.distributed.get_world_size()
        if isinstance(env, int):
            env = DroneBatch(env, EnvConfig.single(), device=device, seed=seed, env_offset=self.rank * env)
        self.batch: DroneBatch = getattr(env, "batch", env)
        if self.batch.obs_dim != 15 or not self.batch.config.auto_reset:
            raise ValueError("PPO needs the DroneGymEnv spec (15-dim obs) with auto-reset")
        self.device = self.batch.device
        self.n_envs, self.n_steps, self.batch_size, self.n_epochs = self.batch.n, n_steps, batch_size, n_epochs
        self.gamma, self.gae_lambda, self.normalize_advantage = gamma, gae_lambda, normalize_advantage
        self.verbose, self.seed = verbose, seed
        if rollout_precision not in ("fp32", "tf32"):
            raise ValueError("rollout_precision must be 'fp32' (CUDA cores, parity path) or 'tf32' (tcgen05 tensor cores)")
        if update_precision not in ("fp32", "tf32", "bf16"):
            raise ValueError("update_precision must be 'fp32' (CUDA cores, parity path), 'tf32' (tcgen05, all products tf32) or "
                             "'bf16' (tcgen05, weight-gradient products with bf16 operands, three tiles per SM)")
        self.rollout_precision, self.update_precision = rollout_precision, update_precision
        cfg = PPOConfig()
        self.lib.dronecu_ppo_config_default(C.byref(cfg))
        cfg.learning_rate, cfg.clip_range, cfg.ent_coef = learning_rate, clip_range, ent_coef
        cfg.vf_coef, cfg.max_grad_norm = vf_coef, max_grad_norm
        self.cfg = cfg
        h = C.c_void_p()
        _lib.check(self.lib.dronecu_ppo_create(C.byref(cfg), self.device.index, C.byref(h)), "dronecu_ppo_create")
        self._h = h
        self.params = init_policy_params(seed if policy_seed is None else policy_seed).to(self.device)
        self.buf = RolloutBuffers(n_steps, self.n_envs, self.device)
        self._grad = torch.zeros(GRAD_LEN, dtype=torch.float32, device=self.device)
        self._adv_stats = torch.zeros(3, dtype=torch.float64, device=self.device)
        self._info = torch.zeros(9, dtype=torch.float32, device=self.device)
        self._gen = torch.Generator(device=self.device).manual_seed(seed + 1)
        self.num_timesteps, self.n_updates = 0, 0
        self._perm, self._epochs_done = None, 0
        self.grad_events = None                   # set to [] to collect (start, end, samples) events per gradient launch
        # One epoch's minibatch sequence (adv-stats -> grad -> reduce -> clip+Adam, x minibatches) as ONE CUDA graph: with
        # SB3's defaults (batch 64) an epoch is hundreds of 10-microsecond kernels and the host launch rate is the bound.
        # Default: on for small minibatches on a single GPU (NCCL all-reduces stay outside graphs here).
        self.cuda_graph = cuda_graph
        self._graph, self._graph_key, self._graph_launches = None, None, 0
        self._ep_stats = None
        self._max_mb = 8 * torch.cuda.get_device_properties(self.device).multi_processor_count
        self.logger_values: dict = {}
        self.launches = 0
        self.batch.reset()          # SB3 _setup_learn: env.reset()  (ep_num 1 -> 2)

    # -- lifetime ------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.dronecu_ppo_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- rollout ------------------------------------------------------------------------------------
    def collect_rollouts(self, deterministic: bool = False):
        """One launch: n_steps env steps for every env, policy evaluated in-kernel; then GAE."""
        b, K, n = self.buf, self.n_steps, self.n_envs
        out = PolicyOut(b.obs.data_ptr(), b.actions.data_ptr(), b.logp.data_ptr(), b.value.data_ptr(),
                        b.reward.data_ptr(), b.done.data_ptr(), b.last_value.data_ptr(), None)
        st = _stream_ptr(self.device)
        fn = self.lib.dronecu_rollout_policy_tc if self.rollout_precision == "tf32" else self.lib.dronecu_rollout_policy
        _lib.check(fn(self.batch._h, K, _ptr(self.params), int(deterministic), C.byref(out), st), "dronecu_rollout_policy")
        _lib.check(self.lib.dronecu_gae(self.device.index, K, n, _ptr(b.reward), _ptr(b.value), _ptr(b.done),
                                        _ptr(b.last_value), self.gamma, self.gae_lambda, _ptr(b.adv), _ptr(b.ret), st),
                   "dronecu_gae")
        self.launches += 2
        self.num_timesteps += K * n * self.world

    # -- update -------------------------------------------------------------------------------------
    def _minibatch(self, index: Optional[torch.Tensor], first: int, m: int, stats: Optional[torch.Tensor] = None):
        """One optimiser step.  `stats`: this minibatch's (already all-reduced) [sum, sumsq, count] of the advantages
        (train() computes them for the whole epoch at once); None: computed here."""
        b, st = self.buf, _stream_ptr(self.device)
        stats_ptr = None
        if self.normalize_advantage and stats is not None:
            stats_ptr = _ptr(stats)
        elif self.normalize_advantage:
            self._adv_stats.zero_()
            _lib.check(self.lib.dronecu_ppo_adv_stats(self._h, _ptr(b.adv), _ptr(index), first, m,
                                                      _ptr(self._adv_stats), st), "dronecu_ppo_adv_stats")
            if self.world > 1:
                torch.distributed.all_reduce(self._adv_stats)
            stats_ptr = _ptr(self._adv_stats)
            self.launches += 2
        grad_fn = {"fp32": self.lib.dronecu_ppo_grad, "tf32": self.lib.dronecu_ppo_grad_tc,
                   "bf16": self.lib.dronecu_ppo_grad_bf16}[self.update_precision]
        if self.grad_events is not None:          # bench.py: CUDA events around the gradient launches
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True), m)
            ev[0].record()
        _lib.check(grad_fn(self._h, _ptr(self.params), _ptr(b.obs), _ptr(b.actions), _ptr(b.logp),
                           _ptr(b.adv), _ptr(b.ret), _ptr(index), first, m, 0.0, 1.0, stats_ptr,
                           _ptr(self._grad), st), "dronecu_ppo_grad")
        if self.grad_events is not None:
            ev[1].record()
            self.grad_events.append(ev)
        if self.world > 1:
            torch.distributed.all_reduce(self._grad)      # NCCL: 42.8 KB, the only collective of the data path
        _lib.check(self.lib.dronecu_ppo_apply(self._h, _ptr(self.params), _ptr(self._grad),
                                              1.0 / (m * self.world), _ptr(self._info), st), "dronecu_ppo_apply")
        self.launches += 3
        self.n_updates += 1

    def _epoch_graph(self, perm: torch.Tensor, B: int, ep_stats: Optional[torch.Tensor] = None):
        """Replay (capture on first use) the CUDA graph of one epoch over the index buffer `perm`."""
        key = (perm.data_ptr(), B, self.batch_size, self.normalize_advantage, self.update_precision,
               None if ep_stats is None else ep_stats.data_ptr())
        if self._graph is None or self._graph_key != key:
            torch.cuda.synchronize(self.device)
            l0, u0 = self.launches, self.n_updates
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for k, start in enumerate(range(0, B, self.batch_size)):
                    m = min(self.batch_size, B - start)
                    self._minibatch(perm[start:start + m], 0, m, None if ep_stats is None else ep_stats[k])
            # capture does not execute: undo the bookkeeping of the capture pass, remember it per replay
            self._graph_launches, self._graph_updates = self.launches - l0, self.n_updates - u0
            self.launches, self.n_updates = l0, u0
            self._graph, self._graph_key = g, key
        self._graph.replay()
        self.launches += self._graph_launches
        self.n_updates += self._graph_updates

    def train(self):
        """SB3 PPO.train(): n_epochs passes over the buffer in random minibatches of batch_size."""
        B = self.n_steps * self.n_envs
        if self._perm is None or self._perm.numel() != B:
            self._perm = torch.empty(B, dtype=torch.int32, device=self.device)
        perm = self._perm
        for _ in range(self.n_epochs):
            # SB3: np.random.permutation(B) per epoch; here a keyed bijection evaluated on the device (no sort)
            _lib.check(self.lib.dronecu_minibatch_permutation(self.device.index, B, self.seed + 1, self._epochs_done,
                                                              _ptr(perm), _stream_ptr(self.device)), "dronecu_minibatch_permutation")
            self._epochs_done += 1
            self.launches += 1
            # advantage statistics of EVERY minibatch of the epoch: two launches and (data parallel) one all-reduce per epoch
            n_mb = (B + self.batch_size - 1) // self.batch_size
            ep_stats = None
            if self.normalize_advantage and n_mb <= self._max_mb:
                if self._ep_stats is None or self._ep_stats.shape[0] != n_mb:
                    self._ep_stats = torch.zeros(n_mb, 3, dtype=torch.float64, device=self.device)
                ep_stats = self._ep_stats
                _lib.check(self.lib.dronecu_ppo_adv_stats_epoch(self._h, _ptr(self.buf.adv), _ptr(perm), B, self.batch_size,
                                                                _ptr(ep_stats), _stream_ptr(self.device)), "dronecu_ppo_adv_stats_epoch")
                self.launches += 2
                if self.world > 1:
                    torch.distributed.all_reduce(ep_stats)
            use_graph = self.cuda_graph if self.cuda_graph is not None else (self.batch_size <= 16384 and B // self.batch_size >= 4)
            if use_graph and self.world == 1 and self.grad_events is None:
                self._epoch_graph(perm, B, ep_stats)
                continue
            for k, start in enumerate(range(0, B, self.batch_size)):
                m = min(self.batch_size, B - start)
                self._minibatch(perm[start:start + m], 0, m, None if ep_stats is None else ep_stats[k])
        info = self._info.cpu().numpy()
        self.logger_values.update({"train/policy_gradient_loss": float(info[0]), "train/value_loss": float(info[1]),
                                   "train/approx_kl": float(info[2]), "train/clip_fraction": float(info[3]),
                                   "train/loss": float(info[0] + self.cfg.vf_coef * info[1]),
                                   "train/grad_norm": float(info[8]), "train/n_updates": self.n_updates,
                                   "train/std": float(torch.exp(self.params[-4:]).mean())})

    def learn(self, total_timesteps: int, log_interval: int = 1, callback=None):
        t0, it = time.time(), 0
        while self.num_timesteps < total_timesteps:
            self.collect_rollouts()
            st = self.batch.episode_stats(reset=True)
            if self.world > 1:
                v = torch.tensor([st["return_sum"], float(st["length_sum"]), float(st["episodes"])],
                                 dtype=torch.float64, device=self.device)
                torch.distributed.all_reduce(v)
                st["ep_rew_mean"] = float(v[0] / v[2]) if v[2] > 0 else float("nan")
                st["ep_len_mean"] = float(v[1] / v[2]) if v[2] > 0 else float("nan")
            self.train()
            it += 1
            self.logger_values.update({"rollout/ep_rew_mean": st["ep_rew_mean"], "rollout/ep_len_mean": st["ep_len_mean"],
                                       "time/iterations": it, "time/total_timesteps": self.num_timesteps,
                                       "time/fps": int(self.num_timesteps / max(time.time() - t0, 1e-9))})
            if callback is not None and callback(self) is False:
                break
            if self.verbose and self.rank == 0 and it % log_interval == 0:
                print(" | ".join(f"{k}={v:.4g}" if isinstance(v, float) else f"{k}={v}"
                                 for k, v in self.logger_values.items()), flush=True)
        return self

    # -- inference ----------------------------------------------------------------------------------
    def policy_forward(self, obs: torch.Tensor, precision: str = "fp32", debug: bool = False):
        """(mean [B,4], value [B]) for device observations [B,15].  precision "tf32" runs the tcgen05
        kernel; with debug=True it also returns the layer-1 / layer-2 pre-activations [B,128] each."""
        obs = obs.to(self.device, torch.float32).contiguous()
        B = obs.shape[0]
        mean = torch.empty(B, 4, dtype=torch.float32, device=self.device)
        value = torch.empty(B, dtype=torch.float32, device=self.device)
        if precision == "tf32":
            d1 = torch.zeros(B, 128, dtype=torch.float32, device=self.device) if debug else None
            d2 = torch.zeros(B, 128, dtype=torch.float32, device=self.device) if debug else None
            _lib.check(self.lib.dronecu_policy_forward_tc(self.device.index, B, _ptr(self.params), _ptr(obs), _ptr(mean),
                                                          _ptr(value), _ptr(d1), _ptr(d2), _stream_ptr(self.device)),
                       "dronecu_policy_forward_tc")
            return (mean, value, d1, d2) if debug else (mean, value)
        _lib.check(self.lib.dronecu_policy_forward(self.device.index, B, _ptr(self.params), _ptr(obs), _ptr(mean),
                                                   _ptr(value), _stream_ptr(self.device)), "dronecu_policy_forward")
        return mean, value

    def predict(self, observation, state=None, episode_start=None, deterministic: bool = False):
        """SB3 ``predict``: numpy obs [15] or [n,15] -> (clipped action, None)  (train.py:48, test.py:14)."""
        obs = np.asarray(observation, dtype=np.float32)
        single = obs.ndim == 1
        mean, _ = self.policy_forward(torch.from_numpy(obs.reshape(-1, 15)))
        if not deterministic:
            std = torch.exp(self.params[-4:])
            mean = mean + std * torch.randn(mean.shape, device=self.device, generator=self._gen)
        act = torch.clamp(mean, 0.0, self.batch.config.motor_max).cpu().numpy()
        return (act[0] if single else act), None

    # -- checkpoint (policy + Adam + env curriculum / RNG state; the reference loses the env state) ------
    def state_dict(self) -> dict:
        mom = torch.empty(2 * POLICY_PARAMS, dtype=torch.float32, device=self.device)
        step = C.c_int64()
        _lib.check(self.lib.dronecu_ppo_get_state(self._h, _ptr(mom), C.byref(step), _stream_ptr(self.device)))
        torch.cuda.synchronize(self.device)
        return {"params": self.params.cpu(), "adam": mom.cpu(), "adam_step": step.value,
                "num_timesteps": self.num_timesteps, "n_updates": self.n_updates, "epochs_done": self._epochs_done,
                "env_state": self.batch.get_state(), "env_global_step": self.batch.global_step,
                "sb3_policy": {SB3_NAMES[k]: v.clone() for k, v in unpack_params(self.params.cpu()).items()}}

    def load_state_dict(self, sd: dict, load_env: bool = True):
        self.params.copy_(sd["params"].to(self.device))
        _lib.check(self.lib.dronecu_ppo_set_state(self._h, _ptr(sd["adam"].to(self.device)), int(sd["adam_step"]),
                                                  _stream_ptr(self.device)))
        torch.cuda.synchronize(self.device)
        self.num_timesteps, self.n_updates = sd["num_timesteps"], sd["n_updates"]
        self._epochs_done = int(sd.get("epochs_done", 0))
        if load_env and "env_state" in sd and sd["env_state"]["pos"].shape[0] == self.n_envs:
            self.batch.set_state(**sd["env_state"])
            self.batch.global_step = sd.get("env_global_step", self.batch.global_step)

    def save(self, path: str):
        """SB3 ``model.save(path)`` (train.py:70): no suffix -> ``path + ".zip"``, an archive in SB3's layout
        (sb3_zip.py) that also carries the env / curriculum / RNG state; ``.pt`` -> a plain torch checkpoint."""
        if path.endswith(".pt"):
            torch.save(self.state_dict(), path)
            return
        from . import sb3_zip
        sd = self.state_dict()
        hyper = {"n_steps": self.n_steps, "batch_size": self.batch_size, "n_epochs": self.n_epochs, "gamma": self.gamma,
                 "gae_lambda": self.gae_lambda, "learning_rate": float(self.cfg.learning_rate),
                 "clip_range": float(self.cfg.clip_range), "ent_coef": float(self.cfg.ent_coef),
                 "vf_coef": float(self.cfg.vf_coef), "max_grad_norm": float(self.cfg.max_grad_norm),
                 "n_envs": self.n_envs, "num_timesteps": self.num_timesteps, "_n_updates": self.n_updates, "seed": self.seed}
        extra = {k: sd[k] for k in ("num_timesteps", "n_updates", "epochs_done", "env_state", "env_global_step")}
        sb3_zip.export_zip(path if path.endswith(".zip") else path + ".zip", sd["params"], sd["adam"], sd["adam_step"],
                           hyper, extra)

    @classmethod
    def load(cls, path: str, env=1, **kw):
        """SB3 ``PPO.load(path, env, **overrides)`` (train.py:22-30, test.py:7): a ``.zip`` written by this class or by
        stable-baselines3 itself (policy + Adam state; hyper-parameters stored in the archive are defaults that
        keyword arguments override), or a ``.pt`` checkpoint."""
        import os
        if not path.endswith((".pt", ".zip")) and os.path.isfile(path + ".zip"):
            path = path + ".zip"
        if path.endswith(".pt"):
            model = cls(env, **kw)
            model.load_state_dict(torch.load(path, weights_only=False))
            return model
        from . import sb3_zip
        z = sb3_zip.import_zip(path)
        for k in ("n_steps", "batch_size", "n_epochs", "gamma", "gae_lambda", "learning_rate", "clip_range", "ent_coef",
                  "vf_coef", "max_grad_norm"):
            if k in z["hyper"] and k not in kw and isinstance(z["hyper"][k], (int, float)):
                kw[k] = z["hyper"][k]
        model = cls(env, **kw)
        sd = {"params": z["params"], "adam": z["adam"] if z["adam"] is not None else torch.zeros(2 * POLICY_PARAMS),
              "adam_step": z["adam_step"], "num_timesteps": int(z["hyper"].get("num_timesteps", 0) or 0),
              "n_updates": int(z["hyper"].get("_n_updates", 0) or 0)}
        if z["extra"]:
            sd.update(z["extra"])
        model.load_state_dict(sd)
        return model
