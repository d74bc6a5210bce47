"""TEST INFRASTRUCTURE: import the *unmodified* reference env modules as a live oracle.

Usable where ``/root/reference`` exists (the build container) or where ``oracle/make_ref.py`` staged a
byte-for-byte copy under ``oracle/_ref/`` (git-ignored; travels to the GPU box with the snapshot).  It is used by
``tests/golden/make_golden.py`` to generate the committed fixtures, by the ``not gpu`` tests that pin
``oracle/drone_oracle.py`` against the real thing, and by ``bench.py``'s CPU legs (``--impl reference``,
``cpu_baseline`` kind "reference").  The ``-m gpu`` tests and ``smoke()`` never depend on it.

The reference imports two third-party packages at module top level that are absent in
this image: ``gym`` (drone.py:2-3, vectorized_drone.py:2-3 -- only ``gym.Env`` and
``spaces.Box`` are touched, drone.py:254-264) and ``matplotlib`` (drone.py:4-7,
vectorized_drone.py:4-6 -- only used for rendering).  We put inert stand-ins in
``sys.modules`` before the import; the reference files themselves are not changed.
"""
from __future__ import annotations

import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref", "reference_py.zip")   # written by oracle/make_ref.py


def _find_reference() -> str:
    if "DRONE_REFERENCE_DIR" in os.environ:
        return os.environ["DRONE_REFERENCE_DIR"]
    if os.path.isfile("/root/reference/drone.py"):
        return "/root/reference"
    return _STAGED                  # the GPU box: the byte-for-byte archive that travelled with the snapshot (zipimport)


REFERENCE_DIR = _find_reference()


def available() -> bool:
    return os.path.isfile(REFERENCE_DIR) or os.path.isfile(os.path.join(REFERENCE_DIR, "drone.py"))


def kind() -> str:
    """Where the live reference comes from: "tree" (/root/reference) or "staged" (oracle/_ref archive)."""
    return "staged" if os.path.isfile(REFERENCE_DIR) else "tree"


class _Box:
    """Stores what ``spaces.Box(...)`` is given (drone.py:259,264)."""

    def __init__(self, low, high, shape, dtype):
        self.low, self.high, self.shape, self.dtype = low, high, shape, dtype


def _install_stubs() -> None:
    if "gym" not in sys.modules:
        gym = types.ModuleType("gym")
        gym.Env = type("Env", (), {"__init__": lambda self: None})
        spaces = types.ModuleType("gym.spaces")
        spaces.Box = _Box
        gym.spaces = spaces
        sys.modules["gym"] = gym
        sys.modules["gym.spaces"] = spaces
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        mpl.use = lambda *a, **k: None
        plt = types.ModuleType("matplotlib.pyplot")
        anim = types.ModuleType("matplotlib.animation")
        anim.PillowWriter = object
        mpl.pyplot, mpl.animation = plt, anim
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
        sys.modules["matplotlib.animation"] = anim


_cache = {}


def load():
    """Return ``(drone, vectorized_drone)`` -- the reference modules, unmodified."""
    if "mods" in _cache:
        return _cache["mods"]
    if not available():
        raise RuntimeError(f"reference not present at {REFERENCE_DIR}")
    _install_stubs()
    sys.dont_write_bytecode = True  # /root/reference is read-only
    if REFERENCE_DIR not in sys.path:
        sys.path.insert(0, REFERENCE_DIR)
    import drone  # noqa: E402
    import vectorized_drone  # noqa: E402

    _cache["mods"] = (drone, vectorized_drone)
    return _cache["mods"]


class UniformStream:
    """Deterministic replacement for ``np.random.rand`` inside the reference's reset.

    ``DroneEnv.reset`` draws exactly five uniforms in the order pos.x, pos.y, tgt.x,
    tgt.y, tgt.z (drone.py:57, :73).  ``feed(values)`` queues the uniforms the next
    resets will consume so the reference can be driven by *our* Philox stream.
    """

    def __init__(self):
        self.queue = []
        self.drawn = 0

    def feed(self, values):
        self.queue.extend(float(v) for v in values)

    def __call__(self, *shape):
        assert not shape, "reference only calls np.random.rand() with no arguments"
        self.drawn += 1
        return self.queue.pop(0)


class patched_rand:
    """Context manager: route ``drone.np.random.rand`` to a UniformStream."""

    def __init__(self, stream: UniformStream):
        self.stream = stream

    def __enter__(self):
        drone, _ = load()
        self._saved = drone.np.random.rand
        drone.np.random.rand = self.stream
        return self.stream

    def __exit__(self, *exc):
        drone, _ = load()
        drone.np.random.rand = self._saved
        return False
