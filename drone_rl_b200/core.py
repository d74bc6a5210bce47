"""Device-level handle: ``DroneBatch`` owns one ``dronecu_env`` (n envs on one GPU).

This is the thin host layer over the C ABI; torch is used only for device/pinned memory and
streams.  The reference-facing classes (``DroneGymEnv``, ``VectorizedDroneGymEnv``,
``DroneVecEnv``) in envs.py are built on it.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Optional

import numpy as np
import torch

from . import _lib
from ._lib import Config, RolloutOut, StateView, Stats, StepOut


@dataclass
class EnvConfig:
    """Every literal of the reference env, 1:1 with ``dronecu_config`` (include/dronecu.h).

    ``EnvConfig.single()`` = DroneEnv/DroneGymEnv (drone.py:14-46) under SB3's auto-resetting
    DummyVecEnv (train.py:18-20); ``EnvConfig.vector()`` = VectorizedDroneEnv
    (vectorized_drone.py:13-36).
    """
    dt: float = 0.02
    mass: float = 1.0
    gravity: float = 9.81
    inertia: tuple = (0.005, 0.005, 0.01)
    arm_length: float = 0.5
    k_yaw: float = 0.01
    reward_scale: float = 0.01
    bonus_radius: float = 0.05
    bonus: float = 1.0
    z_floor: float = 0.0
    r_max: float = 50.0
    fixed_target: tuple = (0.0, 0.0, 10.0)
    fixed_start: tuple = (0.1, 0.1, 0.1)
    start_z: float = 1.0
    target_z: float = 1.0
    curriculum_step: float = 0.1
    curriculum_period: int = 2000
    max_steps: int = 200
    obs_dim: int = 15
    randomized: bool = True
    auto_reset: bool = True

    @classmethod
    def single(cls, **kw):
        return cls(**kw)

    @classmethod
    def vector(cls, **kw):
        base = dict(bonus_radius=1.0, max_steps=1000, obs_dim=12, randomized=False, auto_reset=False)
        base.update(kw)
        return cls(**base)

    def to_c(self) -> Config:
        c = Config()
        for name in ("dt", "mass", "gravity", "arm_length", "k_yaw", "reward_scale", "bonus_radius",
                     "bonus", "z_floor", "r_max", "start_z", "target_z", "curriculum_step",
                     "curriculum_period", "max_steps", "obs_dim"):
            setattr(c, name, getattr(self, name))
        c.inertia[:] = self.inertia
        c.fixed_target[:] = self.fixed_target
        c.fixed_start[:] = self.fixed_start
        c.flags = (_lib.FLAG_RANDOMIZED if self.randomized else 0) | (_lib.FLAG_AUTORESET if self.auto_reset else 0)
        return c

    @property
    def motor_max(self) -> float:
        return 3 * self.mass * self.gravity / 4.0      # drone.py:263


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream_ptr(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class DroneBatch:
    """n quadcopter envs resident on one GPU (one ``dronecu_env`` handle)."""

    def __init__(self, n_envs: int, config: Optional[EnvConfig] = None, device=0, seed: int = 0,
                 env_offset: int = 0):
        self.lib = _lib.load()
        self.config = config or EnvConfig.single()
        self.device = torch.device("cuda", device if isinstance(device, int) else torch.device(device).index or 0)
        if not torch.cuda.is_available():
            raise _lib.DronecuError("no CUDA device: drone_rl_b200 has no CPU fallback")
        torch.cuda.init()
        self.n = int(n_envs)
        self.obs_dim = int(self.config.obs_dim)
        self.seed, self.env_offset = int(seed), int(env_offset)
        handle = C.c_void_p()
        cfg = self.config.to_c()
        _lib.check(self.lib.dronecu_create(C.byref(cfg), self.device.index, self.n, self.env_offset,
                                           C.c_uint64(self.seed & (2 ** 64 - 1)), C.byref(handle)), "dronecu_create")
        self._h = handle

    # -- lifetime ----------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self.lib.dronecu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- helpers -----------------------------------------------------------------------------
    def _check_dev(self, t: torch.Tensor, shape, dtype, name):
        if t.device != self.device or t.dtype != dtype or tuple(t.shape) != tuple(shape) or not t.is_contiguous():
            raise ValueError(f"{name}: expected contiguous {dtype} {tuple(shape)} on {self.device}, got "
                             f"{t.dtype} {tuple(t.shape)} on {t.device}")
        return t

    def empty(self, *shape, dtype=torch.float32):
        return torch.empty(*shape, dtype=dtype, device=self.device)

    @property
    def global_step(self) -> int:
        return int(self.lib.dronecu_global_step(self._h))

    @global_step.setter
    def global_step(self, t: int):
        _lib.check(self.lib.dronecu_set_global_step(self._h, int(t)), "dronecu_set_global_step")

    @property
    def launch_count(self) -> int:
        return int(self.lib.dronecu_launch_count(self._h))

    # -- device API --------------------------------------------------------------------------
    def reset(self, mask: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None):
        """Reset the rows selected by ``mask`` (uint8/bool [n]; None = all) and return obs [n,D]."""
        obs = self.empty(self.n, self.obs_dim) if out is None else self._check_dev(out, (self.n, self.obs_dim), torch.float32, "out")
        if mask is not None:
            mask = self._check_dev(mask.view(torch.uint8) if mask.dtype == torch.bool else mask, (self.n,), torch.uint8, "mask")
        _lib.check(self.lib.dronecu_reset(self._h, _ptr(mask), _ptr(obs), _stream_ptr(self.device)), "dronecu_reset")
        return obs

    def step(self, actions: torch.Tensor, out=None, want_truncated=False, want_terminal_obs=False,
             want_episode=False):
        """One step for every env (device tensors).  Returns a dict with obs [n,D], reward [n],
        done [n] uint8 and, on request, truncated / terminal_obs / episode_r / episode_l."""
        n, D = self.n, self.obs_dim
        self._check_dev(actions, (n, 4), torch.float32, "actions")
        res = dict(out or {})
        res.setdefault("obs", self.empty(n, D))
        res.setdefault("reward", self.empty(n))
        res.setdefault("done", self.empty(n, dtype=torch.uint8))
        if want_truncated:
            res.setdefault("truncated", self.empty(n, dtype=torch.uint8))
        if want_terminal_obs:
            res.setdefault("terminal_obs", self.empty(n, D))
        if want_episode:
            res.setdefault("episode_r", self.empty(n))
            res.setdefault("episode_l", self.empty(n, dtype=torch.int32))
        o = StepOut()
        for key, fld in (("obs", "d_obs"), ("reward", "d_reward"), ("done", "d_done"), ("truncated", "d_truncated"),
                         ("terminal_obs", "d_terminal_obs"), ("episode_r", "d_episode_r"), ("episode_l", "d_episode_l")):
            if res.get(key) is not None:
                setattr(o, fld, res[key].data_ptr())
        _lib.check(self.lib.dronecu_step(self._h, _ptr(actions), C.byref(o), _stream_ptr(self.device)), "dronecu_step")
        return res

    def rollout(self, K: int, actions: Optional[torch.Tensor] = None, *, obs0=None, next_obs=None,
                out_actions=None, reward=None, done=None, truncated=None):
        """K fused steps in one launch.  ``actions`` [K,n,4] float32, or None for in-kernel
        Philox U[0, motor_max) actions.  Every output tensor is optional."""
        n, D = self.n, self.obs_dim
        mode = _lib.ACTIONS_UNIFORM if actions is None else _lib.ACTIONS_STREAMED
        if actions is not None:
            self._check_dev(actions, (K, n, 4), torch.float32, "actions")
        o = RolloutOut()
        if obs0 is not None:
            o.d_obs0 = self._check_dev(obs0, (n, D), torch.float32, "obs0").data_ptr()
        if next_obs is not None:
            o.d_next_obs = self._check_dev(next_obs, (K, n, D), torch.float32, "next_obs").data_ptr()
        if out_actions is not None:
            o.d_actions = self._check_dev(out_actions, (K, n, 4), torch.float32, "out_actions").data_ptr()
        if reward is not None:
            o.d_reward = self._check_dev(reward, (K, n), torch.float32, "reward").data_ptr()
        if done is not None:
            o.d_done = self._check_dev(done, (K, n), torch.uint8, "done").data_ptr()
        if truncated is not None:
            o.d_truncated = self._check_dev(truncated, (K, n), torch.uint8, "truncated").data_ptr()
        _lib.check(self.lib.dronecu_rollout(self._h, int(K), mode, _ptr(actions), C.byref(o),
                                            _stream_ptr(self.device)), "dronecu_rollout")

    # -- host API (numpy / pinned buffers; copies inside the call) --------------------------------
    def step_host(self, actions: np.ndarray, obs=None, reward=None, done=None, truncated=None,
                  terminal_obs=None, episode_r=None, episode_l=None):
        """One step with HOST (numpy, ideally pinned) buffers: H2D + kernel + D2H inside the call."""
        o = StepOut()
        for arr, fld in ((obs, "d_obs"), (reward, "d_reward"), (done, "d_done"), (truncated, "d_truncated"),
                         (terminal_obs, "d_terminal_obs"), (episode_r, "d_episode_r"), (episode_l, "d_episode_l")):
            if arr is not None:
                setattr(o, fld, arr.ctypes.data)
        _lib.check(self.lib.dronecu_step_host(self._h, C.c_void_p(actions.ctypes.data), C.byref(o)), "dronecu_step_host")

    def reset_host(self, obs: Optional[np.ndarray], mask: Optional[np.ndarray] = None):
        _lib.check(self.lib.dronecu_reset_host(
            self._h, None if mask is None else C.c_void_p(mask.ctypes.data),
            None if obs is None else C.c_void_p(obs.ctypes.data)), "dronecu_reset_host")

    _FIELDS = {"pos": ("d_pos", 3, np.float32), "vel": ("d_vel", 3, np.float32),
               "euler": ("d_euler", 3, np.float32), "omega": ("d_omega", 3, np.float32),
               "target": ("d_target", 3, np.float32), "step": ("d_step", 0, np.int32),
               "ep_num": ("d_ep_num", 0, np.int32), "ep_len": ("d_ep_len", 0, np.int32),
               "ep_ret": ("d_ep_ret", 0, np.float32)}

    def get_state(self, *names) -> dict:
        """Host copies of the named per-env attributes (default: all)."""
        names = names or tuple(self._FIELDS)
        view, out = StateView(), {}
        for name in names:
            fld, width, dt = self._FIELDS[name]
            arr = np.empty((self.n, width) if width else (self.n,), dtype=dt)
            setattr(view, fld, arr.ctypes.data)
            out[name] = arr
        _lib.check(self.lib.dronecu_get_state_host(self._h, C.byref(view)), "dronecu_get_state_host")
        return out

    def set_state(self, **arrays):
        view, keep = StateView(), []
        for name, val in arrays.items():
            fld, width, dt = self._FIELDS[name]
            arr = np.ascontiguousarray(np.broadcast_to(np.asarray(val, dtype=dt), (self.n, width) if width else (self.n,)))
            keep.append(arr)
            setattr(view, fld, arr.ctypes.data)
        _lib.check(self.lib.dronecu_set_state_host(self._h, C.byref(view)), "dronecu_set_state_host")

    def episode_stats(self, reset: bool = False) -> dict:
        s = Stats()
        _lib.check(self.lib.dronecu_episode_stats(self._h, C.byref(s), int(reset)), "dronecu_episode_stats")
        d = {k: getattr(s, k) for k, _ in Stats._fields_}
        d["ep_rew_mean"] = d["return_sum"] / d["episodes"] if d["episodes"] else float("nan")
        d["ep_len_mean"] = d["length_sum"] / d["episodes"] if d["episodes"] else float("nan")
        return d
