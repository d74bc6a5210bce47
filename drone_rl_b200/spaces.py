"""Minimal ``Box`` space with the attributes SB3 / the reference's scripts read.

The reference builds ``gym.spaces.Box`` objects (drone.py:259,264; vectorized_drone.py:256,260).
gym / gymnasium are optional here: when one of them is importable its Box is used so that
``isinstance`` checks in third-party code pass; otherwise this stand-in is.
"""
from __future__ import annotations

import numpy as np


class _Box:
    def __init__(self, low, high, shape, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(shape)
        self.low = np.broadcast_to(np.asarray(low, dtype=self.dtype), self.shape).copy()
        self.high = np.broadcast_to(np.asarray(high, dtype=self.dtype), self.shape).copy()
        self._rng = np.random.default_rng()

    def seed(self, seed=None):
        self._rng = np.random.default_rng(seed)
        return [seed]

    def sample(self):
        lo = np.where(np.isfinite(self.low), self.low, -1.0)
        hi = np.where(np.isfinite(self.high), self.high, 1.0)
        return self._rng.uniform(lo, hi).astype(self.dtype)

    def contains(self, x):
        x = np.asarray(x)
        return x.shape == self.shape and bool(np.all(x >= self.low) and np.all(x <= self.high))

    def __repr__(self):
        return f"Box({self.low.min()}, {self.high.max()}, {self.shape}, {self.dtype})"


def _pick_box():
    for mod in ("gymnasium.spaces", "gym.spaces"):
        try:
            m = __import__(mod, fromlist=["Box"])
            if getattr(m.Box, "__module__", "").startswith(("gymnasium", "gym.")):
                return m.Box
        except Exception:
            continue
    return _Box


Box = _pick_box()
