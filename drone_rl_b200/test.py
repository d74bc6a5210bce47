"""``python -m drone_rl_b200.test`` -- the reference's evaluation script (test.py:1-24): load ``./dd.zip``, run the
deterministic policy for 100 steps in a single ``DroneGymEnv`` while recording, save ``my_drone_run.gif``."""
from __future__ import annotations

import argparse
import time


def main(argv=None):
    ap = argparse.ArgumentParser(description=__doc__)
    ap.add_argument("--model", default="./dd.zip")                    # test.py:7
    ap.add_argument("--out", default="my_drone_run.gif")              # test.py:10
    ap.add_argument("--steps", type=int, default=100)                 # test.py:13
    ap.add_argument("--seed", type=int, default=0)
    args = ap.parse_args(argv)
    from . import DroneGymEnv
    from .ppo import PPO
    s = time.time()
    env = DroneGymEnv(seed=args.seed)
    model = PPO.load(args.model, env=1)
    env.start_record(args.out, dpi=200, fps=20)
    obs = env.reset()
    for _ in range(args.steps):
        action, _ = model.predict(obs, deterministic=True)
        obs, reward, done, info = env.step(action)
        env.render()
        if done:
            obs = env.reset()
    env.stop_record()
    model.close()
    env.close()
    print(time.time() - s)


if __name__ == "__main__":
    main()
