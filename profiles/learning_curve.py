#!/usr/bin/env python
"""Run the training driver twice (fp32 parity path, tensor-core path) and write every 10th logged iteration to a json:
    python profiles/learning_curve.py gpurun_out/r02_learning_curve.json      (on a B200)"""
import json, os, subprocess, sys, tempfile
out = sys.argv[1]
runs = {}
for prec in ("fp32", "tf32"):
    d = tempfile.mkdtemp()
    subprocess.check_call([sys.executable, "-m", "drone_rl_b200.train", "--n-envs", "4096", "--n-steps", "64", "--batch-size", "65536",
                           "--total-timesteps", "1e8", "--precision", prec, "--quiet", "--tensorboard-root", d, "--save", os.path.join(d, "m"),
                           "--resume", os.path.join(d, "none.zip")], cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    rows = [json.loads(l) for l in open(os.path.join(d, "drone_runs_1", "progress.jsonl"))]
    keep = ("step", "train/value_loss", "train/approx_kl", "train/clip_fraction", "train/explained_variance", "train/std",
            "rollout/ep_rew_mean", "rollout/ep_len_mean", "time/fps")
    runs[prec] = [{k: (round(r[k], 5) if isinstance(r[k], float) else r[k]) for k in keep if k in r} for r in rows[::10] + rows[-1:]]
json.dump({"_doc": "python -m drone_rl_b200.train --n-envs 4096 --n-steps 64 --batch-size 65536 --total-timesteps 1e8 --precision {fp32,tf32} "
                   "on one B200 (reference reward, curriculum on, seed 0); every 10th iteration of progress.jsonl. time/fps is wall-clock "
                   "env-steps/s of the whole loop incl. logging. Round 2: SB3-ordered torch init, sorted minibatch partition, train/* = "
                   "means over the minibatches, explained variance.", "runs": runs}, open(out, "w"), indent=1)
print({p: (r[0]["rollout/ep_rew_mean"], r[-1]["rollout/ep_rew_mean"], r[-1]["train/std"], r[-1]["time/fps"]) for p, r in runs.items()})
