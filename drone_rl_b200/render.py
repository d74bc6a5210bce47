"""Host-side rendering of one drone -- the reference's ``DroneEnv.render`` / ``start_record`` / ``stop_record``
(drone.py:189-248; used by test.py:10-22) without matplotlib (absent here): Pillow draws the same scene -- target
(green), the two arms between opposite motors (purple), centre (red), motors (blue) inside the reference's axis box
x, y in [-5, 5], z in [0, 5] -- in a fixed orthographic view (matplotlib's default azimuth -60, elevation 30), and writes
the recording as an animated GIF (the reference's ``PillowWriter`` does the same through matplotlib).

Not on the hot path: one ``get_state`` per frame.  The body-to-inertial rotation is restated from drone.py:161-174."""
from __future__ import annotations

import math

import numpy as np


def rotation_matrix(euler) -> np.ndarray:
    """ZYX body-to-inertial rotation (drone.py:161-174)."""
    roll, pitch, yaw = (float(e) for e in euler)
    cr, sr, cp, sp, cy, sy = math.cos(roll), math.sin(roll), math.cos(pitch), math.sin(pitch), math.cos(yaw), math.sin(yaw)
    return np.array([[cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
                     [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
                     [-sp, cp * sr, cp * cr]])


def motor_positions(pos, euler, arm_length: float) -> np.ndarray:
    """The four motor positions (drone.py:222-229)."""
    a = arm_length / math.sqrt(2.0)
    offsets = np.array([[a, a, 0.0], [-a, a, 0.0], [-a, -a, 0.0], [a, -a, 0.0]])
    return np.asarray(pos, dtype=np.float64) + (rotation_matrix(euler) @ offsets.T).T


class FrameRenderer:
    def __init__(self, size: int = 480, xlim=(-5.0, 5.0), ylim=(-5.0, 5.0), zlim=(0.0, 5.0), azim: float = -60.0, elev: float = 30.0):
        self.size, self.lim = size, np.array([xlim, ylim, zlim], dtype=np.float64)
        az, el = math.radians(azim), math.radians(elev)
        # camera basis: right and up vectors of an orthographic view from (azim, elev)
        self.right = np.array([-math.sin(az), math.cos(az), 0.0])
        self.up = np.array([-math.sin(el) * math.cos(az), -math.sin(el) * math.sin(az), math.cos(el)])
        corners = np.array([[x, y, z] for x in self.lim[0] for y in self.lim[1] for z in self.lim[2]])
        uv = np.stack([corners @ self.right, corners @ self.up], 1)
        self.lo, self.hi = uv.min(0), uv.max(0)

    def project(self, p) -> tuple:
        p = np.asarray(p, dtype=np.float64)
        u, v = p @ self.right, p @ self.up
        pad = 0.06 * self.size
        s = (self.size - 2 * pad) / max(self.hi[0] - self.lo[0], self.hi[1] - self.lo[1])
        return (pad + (u - self.lo[0]) * s, self.size - pad - (v - self.lo[1]) * s)

    def draw(self, pos, euler, target, arm_length: float = 0.5):
        from PIL import Image, ImageDraw
        img = Image.new("RGB", (self.size, self.size), "white")
        d = ImageDraw.Draw(img)
        (x0, x1), (y0, y1), (z0, z1) = self.lim
        floor = [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0)]
        for a, b in zip(floor, floor[1:] + floor[:1]):
            d.line([self.project(a), self.project(b)], fill=(170, 170, 170), width=1)
        for c in floor:
            d.line([self.project(c), self.project((c[0], c[1], z1))], fill=(210, 210, 210), width=1)
        for name, end in (("X", (x1, y0, z0)), ("Y", (x0, y1, z0)), ("Z", (x0, y0, z1))):
            d.text(self.project(end), name, fill="black")

        def dot(p, r, col):
            u, v = self.project(p)
            d.ellipse([u - r, v - r, u + r, v + r], fill=col)
        if np.all(np.isfinite(target)):
            dot(target, 6, (0, 160, 0))                                   # ax.scatter(target, color='green', s=50)
        if np.all(np.isfinite(pos)) and np.all(np.isfinite(euler)):
            m = motor_positions(pos, euler, arm_length)
            d.line([self.project(m[0]), self.project(m[2])], fill=(128, 0, 128), width=3)     # purple arms
            d.line([self.project(m[1]), self.project(m[3])], fill=(128, 0, 128), width=3)
            dot(pos, 4, (220, 0, 0))                                      # centre, red
            for q in m:
                dot(q, 4, (0, 0, 220))                                    # motors, blue
        return img


    def draw_batch(self, pos, target):
        """vectorized_drone.py:218-243: the target (green) and every drone centre (red) -- use limits (-20, 20), (-20, 20), (0, 20)."""
        from PIL import Image, ImageDraw
        img = Image.new("RGB", (self.size, self.size), "white")
        d = ImageDraw.Draw(img)
        (x0, x1), (y0, y1), (z0, z1) = self.lim
        floor = [(x0, y0, z0), (x1, y0, z0), (x1, y1, z0), (x0, y1, z0)]
        for a, b in zip(floor, floor[1:] + floor[:1]):
            d.line([self.project(a), self.project(b)], fill=(170, 170, 170), width=1)
        for c in floor:
            d.line([self.project(c), self.project((c[0], c[1], z1))], fill=(210, 210, 210), width=1)
        u, v = self.project(target)
        d.ellipse([u - 6, v - 6, u + 6, v + 6], fill=(0, 160, 0))
        pos = np.asarray(pos, dtype=np.float64)
        for p in pos[np.all(np.isfinite(pos), axis=1)]:
            u, v = self.project(p)
            d.ellipse([u - 3, v - 3, u + 3, v + 3], fill=(220, 0, 0))
        return img


class Recorder:
    """``start_record`` / ``stop_record`` state: frames of ``render()`` calls, saved as an animated GIF."""

    def __init__(self, filename: str, fps: int = 20):
        self.filename, self.fps, self.frames = filename, fps, []

    def grab(self, img):
        self.frames.append(img)

    def finish(self):
        if not self.frames:
            return
        name = self.filename if self.filename.lower().endswith(".gif") else self.filename.rsplit(".", 1)[0] + ".gif"
        self.frames[0].save(name, save_all=True, append_images=self.frames[1:], duration=int(round(1000 / max(self.fps, 1))), loop=0)
        self.saved_as = name
