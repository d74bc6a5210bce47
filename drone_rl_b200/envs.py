"""Reference-facing environment classes: same names, constructor arguments, return types
and error behaviour as the reference's Python surface, computed by the CUDA env.

=========================  ======================================================================
this module                 reference
=========================  ======================================================================
``DroneGymEnv``             drone.py:254-274 (+ ``DroneEnv`` attributes, drone.py:13-75)
``VectorizedDroneGymEnv``   vectorized_drone.py:251-269 (+ ``VectorizedDroneEnv`` :12-57)
``DroneVecEnv``             what train.py:18-20 builds: ``VecMonitor(DummyVecEnv([DroneGymEnv]*n))``
``DroneGymnasiumEnv``       Gymnasium 5-tuple variant of ``DroneGymEnv`` (north-star: "Gym/Gymnasium")
=========================  ======================================================================

Like the reference, none of these validates or clips actions (clipping is SB3's job:
SURVEY.md section 8b) and none raises on NaN.  Unlike the reference, the random start / target
uniforms come from a counter-based Philox stream keyed by (seed, env id, episode number)
instead of numpy's global MT19937 (drone.py:57,73).
"""
from __future__ import annotations

import time
from typing import Optional

import numpy as np
import torch

from .core import DroneBatch, EnvConfig
from .spaces import Box


def _pinned(shape, dtype):
    return torch.empty(shape, dtype=dtype, pin_memory=True).numpy()


class _HostEnvBase:
    """Pinned host buffers + the blocking host<->device step used by the numpy-facing classes."""

    def _alloc_host(self, n, D):
        self._h_act = _pinned((n, 4), torch.float32)
        self._h_obs = _pinned((n, D), torch.float32)
        self._h_term = _pinned((n, D), torch.float32)
        self._h_rew = _pinned((n,), torch.float32)
        self._h_done = _pinned((n,), torch.uint8)
        self._h_trunc = _pinned((n,), torch.uint8)
        self._h_ep_r = _pinned((n,), torch.float32)
        self._h_ep_l = _pinned((n,), torch.int32)

    def _step_host(self, action, want_info=True, want_truncated=True):
        a = action
        if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == self._h_act.shape
                and a.flags.c_contiguous):
            # anything else (lists, float64, [4] for the single env) goes through the pinned staging buffer
            np.copyto(self._h_act, np.asarray(action).reshape(self._h_act.shape), casting="unsafe")
            a = self._h_act
        # every output is optional at the C ABI (a NULL pointer is neither computed nor copied): only what the caller's
        # protocol needs crosses PCIe
        self.batch.step_host(a, self._h_obs, self._h_rew, self._h_done, self._h_trunc if want_truncated else None,
                             self._h_term if want_info else None, self._h_ep_r if want_info else None,
                             self._h_ep_l if want_info else None)

    # attribute access the reference's callers use: env.pos, env.target, env.euler ... (traj_tb.py:34)
    def _attr(self, name):
        return self.batch.get_state(name)[name]


class DroneGymEnv(_HostEnvBase):
    """Single quadcopter env with the legacy gym API (reference: drone.py:254-274).

    ``reset() -> obs float32[15]``; ``step(a[4]) -> (obs float32[15], reward float, done bool, {})``.
    The env does not reset itself on ``done`` -- exactly like the reference, the caller (or SB3's
    DummyVecEnv) does.
    """
    metadata = {"render.modes": ["human"]}

    def __init__(self, dt: float = 0.02, seed: int = 0, device=0, env_id: int = 0):
        cfg = EnvConfig.single(dt=dt, auto_reset=False)
        self.batch = DroneBatch(1, cfg, device=device, seed=seed, env_offset=env_id)
        self.dt, self.mass, self.g = dt, cfg.mass, cfg.gravity
        self.I = np.array(cfg.inertia)
        self.arm_length, self.k_yaw, self.max_steps = cfg.arm_length, cfg.k_yaw, cfg.max_steps
        self.add = 0
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(15,), dtype=np.float32)   # drone.py:259
        motor_max = 3 * self.mass * self.g / 4.0                                                  # drone.py:263
        self.action_space = Box(low=0, high=motor_max, shape=(4,), dtype=np.float32)
        self._alloc_host(1, 15)
        self.total_steps = 1                                                                      # drone.py:19

    # reference attributes ------------------------------------------------------------------------
    pos = property(lambda self: self._attr("pos")[0].astype(np.float64))
    vel = property(lambda self: self._attr("vel")[0].astype(np.float64))
    euler = property(lambda self: self._attr("euler")[0].astype(np.float64))
    omega = property(lambda self: self._attr("omega")[0].astype(np.float64))
    target = property(lambda self: self._attr("target")[0].astype(np.float64))
    current_step = property(lambda self: int(self._attr("step")[0]))
    ep_num = property(lambda self: int(self._attr("ep_num")[0]))

    @property
    def eps(self):   # drone.py:33,68-70 -- float64 accumulation of +0.1 every 2000 episodes
        e = 0.0
        for _ in range(self.ep_num // 2000):
            e += 0.1
        return e

    def reset(self):
        self.batch.reset_host(self._h_obs)
        return self._h_obs[0].copy()

    _want_truncated = False            # the legacy gym 4-tuple has no use for it; DroneGymnasiumEnv turns it on

    def step(self, action):
        self.total_steps += 1
        self._step_host(action, want_info=False, want_truncated=self._want_truncated)
        return self._h_obs[0].copy(), float(self._h_rew[0]), bool(self._h_done[0]), {}

    def _get_obs(self):
        s = self.batch.get_state("pos", "vel", "euler", "omega", "target")
        return np.concatenate([s["pos"][0], s["vel"][0], s["euler"][0], s["omega"][0],
                               s["target"][0] - s["pos"][0]]).astype(np.float32)

    # rendering / recording (drone.py:189-248; test.py:10-22) -- Pillow instead of matplotlib, see render.py ---------------
    def start_record(self, filename="drone_run.gif", dpi=200, fps=20, bitrate=-1):
        """Call before the loop; every ``render()`` then grabs a frame (drone.py:189-197).  ``dpi`` / ``bitrate`` are
        accepted for signature compatibility; the output is an animated GIF (what the reference's PillowWriter writes)."""
        from .render import Recorder
        self._recorder = Recorder(filename, fps=fps)

    def stop_record(self):
        """Finish and save the recording (drone.py:199-203)."""
        rec = getattr(self, "_recorder", None)
        if rec is not None:
            rec.finish()
            self._recorder = None

    def render(self, mode="human", close=False):
        """Draw the scene of drone.py:205-241 (target, arms, centre, motors in the [-5,5] x [-5,5] x [0,5] box); returns the
        PIL image, and appends it to the recording when one is active (drone.py:243-248).  There is no live window."""
        from .render import FrameRenderer
        if getattr(self, "_renderer", None) is None:
            self._renderer = FrameRenderer()
        s = self.batch.get_state("pos", "euler", "target")
        img = self._renderer.draw(s["pos"][0], s["euler"][0], s["target"][0], self.arm_length)
        if getattr(self, "_recorder", None) is not None:
            self._recorder.grab(img)
        return img

    def close(self):
        self.batch.close()


class DroneGymnasiumEnv(DroneGymEnv):
    """Gymnasium API: ``reset(seed, options) -> (obs, info)``, ``step -> (obs, r, terminated, truncated, info)``
    with ``terminated = z<0 or |pos|>50`` and ``truncated = time limit only`` (their OR is the reference's done)."""
    _want_truncated = True

    def reset(self, *, seed: Optional[int] = None, options=None):
        if seed is not None:
            # a new seed re-keys the Philox stream: rebuild the handle, keep the device
            dev = self.batch.device.index
            self.batch.close()
            self.batch = DroneBatch(1, EnvConfig.single(dt=self.dt, auto_reset=False), device=dev, seed=seed)
        return super().reset(), {}

    def step(self, action):
        obs, rew, done, info = super().step(action)
        truncated = bool(self._h_trunc[0])
        return obs, rew, done and not truncated, truncated, info


class VectorizedDroneGymEnv(_HostEnvBase):
    """Batched env with the reference's numpy surface (vectorized_drone.py:251-269):
    ``reset() -> float32[B,12]``; ``step(a[B,4]) -> (float32[B,12], float64[B], bool[B], {})``.
    Fixed target [0,0,10], start [.1,.1,.1], one shared 1000-step limit, no resets on done."""

    def __init__(self, batch_size: int = 10, dt: float = 0.02, device=0):
        cfg = EnvConfig.vector(dt=dt)
        self.batch = DroneBatch(batch_size, cfg, device=device)
        self.batch_size = batch_size
        self.dt, self.mass, self.g = dt, cfg.mass, cfg.gravity
        self.I = np.array(cfg.inertia)
        self.arm_length, self.k_yaw, self.max_steps = cfg.arm_length, cfg.k_yaw, cfg.max_steps
        self.target = np.array(cfg.fixed_target)                                              # vectorized_drone.py:30
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(batch_size, 12), dtype=np.float32)
        motor_max = 3 * self.mass * self.g / 4.0
        self.action_space = Box(low=0, high=motor_max, shape=(batch_size, 4), dtype=np.float32)
        self._alloc_host(batch_size, 12)

    pos = property(lambda self: self._attr("pos").astype(np.float64))
    vel = property(lambda self: self._attr("vel").astype(np.float64))
    euler = property(lambda self: self._attr("euler").astype(np.float64))
    omega = property(lambda self: self._attr("omega").astype(np.float64))
    current_step = property(lambda self: int(self._attr("step")[0]))      # shared counter, vectorized_drone.py:56

    def reset(self):
        self.batch.reset_host(self._h_obs)
        return self._h_obs.copy()

    def step(self, action):
        self._step_host(action, want_info=False, want_truncated=False)
        return self._h_obs.copy(), self._h_rew.astype(np.float64), self._h_done.astype(bool), {}

    def _get_obs(self):
        s = self.batch.get_state("pos", "vel", "euler", "omega")
        return np.concatenate([s["pos"], s["vel"], s["euler"], s["omega"]], axis=1).astype(np.float32)

    def render(self, ax=None):
        """The scene of vectorized_drone.py:218-243 (target + every drone centre in the [-20,20] x [-20,20] x [0,20] box) as a
        PIL image (render.py); `ax` (a matplotlib axis in the reference) is ignored, there is no live window."""
        from .render import FrameRenderer
        if getattr(self, "_renderer", None) is None:
            self._renderer = FrameRenderer(xlim=(-20.0, 20.0), ylim=(-20.0, 20.0), zlim=(0.0, 20.0))
        return self._renderer.draw_batch(self._attr("pos"), self.target)

    def close(self):
        self.batch.close()


class DroneVecEnv(_HostEnvBase):
    """SB3 ``VecEnv`` duck-type equal to ``VecMonitor(DummyVecEnv([lambda: DroneGymEnv()] * n))``
    (train.py:18-20, :33-35): auto-reset inside ``step``, ``infos[i]["terminal_observation"]`` and
    ``infos[i]["episode"] = {"r","l","t"}`` on done, float32 rewards, bool dones.

    ``info_mode="sb3"`` builds the list of per-env dicts SB3 expects (sensible for n up to a few
    thousand); ``info_mode="arrays"`` returns one dict of arrays instead (for large n); ``"none"`` returns an empty
    dict and copies no ``truncated`` flags;
    ``copy=False`` returns views of the pinned staging buffers (valid until the next step);
    ``obs_device=True`` is for callers whose policy runs on the GPU: ``reset`` / ``step`` return the observations as a
    CUDA tensor [n, D] that never leaves the device (60 of the 66 bytes per env-step that would cross PCIe), rewards and
    dones still arrive as numpy arrays; actions may be numpy (copied H2D) or a CUDA tensor (no copy at all).
    """

    def __init__(self, n_envs: int = 1, seed: int = 0, device=0, env_offset: int = 0, dt: float = 0.02,
                 info_mode: str = "sb3", copy: bool = True, config: Optional[EnvConfig] = None, obs_device: bool = False):
        cfg = config or EnvConfig.single(dt=dt)
        self.batch = DroneBatch(n_envs, cfg, device=device, seed=seed, env_offset=env_offset)
        self.num_envs = n_envs
        D = cfg.obs_dim
        self.observation_space = Box(low=-np.inf, high=np.inf, shape=(D,), dtype=np.float32)
        self.action_space = Box(low=0, high=cfg.motor_max, shape=(4,), dtype=np.float32)
        if info_mode not in ("sb3", "arrays", "none"):
            raise ValueError("info_mode must be 'sb3', 'arrays' or 'none'")
        self.info_mode, self.copy, self.obs_device = info_mode, copy, obs_device
        self._alloc_host(n_envs, D)
        if obs_device:
            dev = self.batch.device
            self._d_act = torch.empty(n_envs, 4, device=dev)
            self._d_out = {"obs": torch.empty(n_envs, D, device=dev), "reward": torch.empty(n_envs, device=dev),
                           "done": torch.empty(n_envs, dtype=torch.uint8, device=dev),
                           "truncated": torch.empty(n_envs, dtype=torch.uint8, device=dev)}
            self._t_rew, self._t_done = torch.from_numpy(self._h_rew), torch.from_numpy(self._h_done)
            self._t_trunc, self._t_act = torch.from_numpy(self._h_trunc), torch.from_numpy(self._h_act)
            if info_mode == "sb3":
                raise ValueError("obs_device=True keeps terminal observations on the device: use info_mode 'arrays' or 'none'")
        self._t_start = time.time()
        self._actions = None
        self.render_mode = None
        # VecMonitor zeroes its accumulators in reset(); the device accumulators start at zero too

    # -- VecEnv protocol -------------------------------------------------------------------------
    def reset(self):
        if self.obs_device:
            return self.batch.reset(out=self._d_out["obs"])
        self.batch.reset_host(self._h_obs)
        return self._h_obs.copy() if self.copy else self._h_obs

    def step_async(self, actions):
        self._actions = actions

    def _step_obs_on_device(self):
        a = self._actions
        if isinstance(a, torch.Tensor) and a.is_cuda:
            d_act = a
        else:
            if not (isinstance(a, np.ndarray) and a.dtype == np.float32 and a.shape == self._h_act.shape):
                np.copyto(self._h_act, np.asarray(a).reshape(self._h_act.shape), casting="unsafe")
                a = self._h_act
            src = self._t_act if a is self._h_act else torch.from_numpy(a)
            self._d_act.copy_(src, non_blocking=True)
            d_act = self._d_act
        want_t = self.info_mode == "arrays"
        out = self.batch.step(d_act, out={k: v for k, v in self._d_out.items() if k != "truncated" or want_t},
                              want_truncated=want_t)
        self._t_rew.copy_(out["reward"], non_blocking=True)
        self._t_done.copy_(out["done"], non_blocking=True)
        if want_t:
            self._t_trunc.copy_(out["truncated"], non_blocking=True)
        torch.cuda.current_stream(self.batch.device).synchronize()
        rew, done = self._h_rew, self._h_done.view(np.bool_)
        if self.copy:
            rew, done = rew.copy(), done.copy()
        infos = {"truncated": self._h_trunc.view(np.bool_).copy() if self.copy else self._h_trunc.view(np.bool_)} if want_t else {}
        return out["obs"], rew, done, infos

    def step_wait(self):
        if self.obs_device:
            return self._step_obs_on_device()
        self._step_host(self._actions, want_info=self.info_mode == "sb3", want_truncated=self.info_mode == "arrays")
        obs, rew, done = self._h_obs, self._h_rew, self._h_done.view(np.bool_)
        if self.copy:
            obs, rew, done = obs.copy(), rew.copy(), done.copy()
        if self.info_mode == "sb3":
            infos = [{} for _ in range(self.num_envs)]
            for i in np.flatnonzero(done):
                infos[i]["terminal_observation"] = self._h_term[i].copy()
                infos[i]["episode"] = {"r": float(self._h_ep_r[i]), "l": int(self._h_ep_l[i]),
                                       "t": round(time.time() - self._t_start, 6)}
                # NB: no "TimeLimit.truncated" key -- the reference env reports a time-out as a plain
                # done (drone.py:156-157), so SB3 does not bootstrap on it.
        elif self.info_mode == "arrays":
            infos = {"truncated": self._h_trunc.view(np.bool_).copy() if self.copy else self._h_trunc.view(np.bool_)}
        else:
            infos = {}
        return obs, rew, done, infos

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        self.batch.close()

    def get_attr(self, attr_name, indices=None):
        idx = self._indices(indices)
        if attr_name in ("pos", "vel", "euler", "omega", "target"):
            arr = self._attr(attr_name).astype(np.float64)
            return [arr[i] for i in idx]
        if attr_name in ("current_step", "ep_num"):
            arr = self._attr("step" if attr_name == "current_step" else "ep_num")
            return [int(arr[i]) for i in idx]
        if attr_name == "render_mode":
            return [None for _ in idx]
        cfg = self.batch.config
        alias = {"g": "gravity", "I": "inertia"}
        return [getattr(cfg, alias.get(attr_name, attr_name)) for _ in idx]

    def set_attr(self, attr_name, value, indices=None):
        idx = self._indices(indices)
        if attr_name in ("pos", "vel", "euler", "omega", "target"):
            arr = self._attr(attr_name)
            arr[idx] = np.asarray(value, dtype=np.float32)
            self.batch.set_state(**{attr_name: arr})
        else:
            raise AttributeError(f"set_attr({attr_name!r}) is not supported by the CUDA env")

    def env_method(self, method_name, *args, indices=None, **kwargs):
        raise NotImplementedError("env_method: the envs live on the GPU, there are no per-env Python objects")

    def env_is_wrapped(self, wrapper_class, indices=None):
        return [False for _ in self._indices(indices)]

    def seed(self, seed=None):
        return [None] * self.num_envs

    def _indices(self, indices):
        if indices is None:
            return list(range(self.num_envs))
        if isinstance(indices, int):
            return [indices]
        return list(indices)

    def episode_stats(self, reset=False):
        return self.batch.episode_stats(reset)


def make_sb3_vec_env(n_envs: int = 1, **kwargs):
    """A real ``stable_baselines3.common.vec_env.VecEnv`` over ``DroneVecEnv`` -- what the reference's
    ``VecMonitor(DummyVecEnv([DroneGymEnv] * n))`` (train.py:18-20, :33-35) becomes when stable-baselines3 is installed
    next to this package: ``PPO("MlpPolicy", env=make_sb3_vec_env(n), device="cpu")`` then runs the stock SB3 algorithm
    on the CUDA env (SB3 insists on ``isinstance(env, VecEnv)``, and on gymnasium spaces).

    SB3 is NOT in this image: the subclass is built lazily here and exercised in the tests against
    stub modules with SB3's published abstract interface (reset / step_async / step_wait / close / get_attr / set_attr /
    env_method / env_is_wrapped; ``VecEnv.__init__(num_envs, observation_space, action_space)``) -- UNPINNED like every
    SB3-facing piece.  Raises ImportError when stable-baselines3 or gymnasium is missing.
    """
    from stable_baselines3.common.vec_env import VecEnv          # noqa: the optional dependency
    import gymnasium

    inner = DroneVecEnv(n_envs, info_mode="sb3", copy=True, **kwargs)
    obs_space = gymnasium.spaces.Box(low=-np.inf, high=np.inf, shape=inner.observation_space.shape, dtype=np.float32)
    act_space = gymnasium.spaces.Box(low=0.0, high=float(inner.batch.config.motor_max), shape=(4,), dtype=np.float32)

    class DroneSB3VecEnv(VecEnv):
        def __init__(self):
            super().__init__(inner.num_envs, obs_space, act_space)
            self.inner, self.batch = inner, inner.batch

        def reset(self):
            return inner.reset()

        def step_async(self, actions):
            inner.step_async(np.asarray(actions, dtype=np.float32))

        def step_wait(self):
            return inner.step_wait()

        def close(self):
            inner.close()

        def get_attr(self, attr_name, indices=None):
            return inner.get_attr(attr_name, indices)

        def set_attr(self, attr_name, value, indices=None):
            inner.set_attr(attr_name, value, indices)

        def env_method(self, method_name, *method_args, indices=None, **method_kwargs):
            return inner.env_method(method_name, *method_args, indices=indices, **method_kwargs)

        def env_is_wrapped(self, wrapper_class, indices=None):
            return inner.env_is_wrapped(wrapper_class, indices)

        def seed(self, seed=None):
            return inner.seed(seed)

        def get_images(self):
            return [None] * inner.num_envs

        def episode_stats(self, reset=False):
            return inner.episode_stats(reset)

    return DroneSB3VecEnv()
