"""Final ep_rew_mean / std of the training driver after 1e8 steps for several seeds and both precisions (is the gap between the
fp32 and the tensor-core learning curves systematic or run-to-run variation?)"""
import json, os, subprocess, sys, tempfile
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
res = {}
for seed in (1, 2, 3):
    for prec, up in (("fp32", "fp32"), ("tf32", "bf16"), ("tf32", "tf32")):
        d = tempfile.mkdtemp()
        subprocess.check_call([sys.executable, "-m", "drone_rl_b200.train", "--n-envs", "4096", "--n-steps", "64", "--batch-size", "65536",
                               "--total-timesteps", "1e8", "--precision", prec, "--update-precision", up, "--seed", str(seed), "--quiet",
                               "--tensorboard-root", d, "--save", os.path.join(d, "m"), "--resume", os.path.join(d, "none.zip")],
                              cwd=root, stdout=subprocess.DEVNULL)
        rows = [json.loads(l) for l in open(os.path.join(d, "drone_runs_1", "progress.jsonl"))]
        last = rows[-5:]
        res[f"seed{seed}/{prec}+{up}"] = (round(sum(r["rollout/ep_rew_mean"] for r in last) / 5, 4), round(last[-1]["train/std"], 3))
        print(f"seed{seed}/{prec}+{up}", res[f"seed{seed}/{prec}+{up}"], flush=True)
json.dump(res, open(sys.argv[1], "w"), indent=1)
