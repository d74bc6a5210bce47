"""The training driver end to end on the GPU: reference train.py flow (fresh run, save, resume from the zip)."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_train_driver_fresh_then_resume(tmp_path, monkeypatch):
    from drone_rl_b200 import train
    from drone_rl_b200.ppo import PPO
    monkeypatch.chdir(tmp_path)
    # the reference's own shape, shortened: ONE env, n_steps = 2048 // n_envs, batch 64, 10 epochs (train.py:12-16, :36-43)
    train.main(["--total-timesteps", "4096", "--quiet", "--save", "dd"])
    assert os.path.isfile("dd.zip") and os.path.isdir("tensorboard/drone_runs_1")
    rows = [json.loads(l) for l in open("tensorboard/drone_runs_1/progress.jsonl")]
    assert len(rows) == 2 and rows[-1]["step"] == 4096
    assert rows[-1]["train/n_updates"] == 2 * 10 * (2048 // 64)            # 320 Adam steps per 2048 transitions
    for k in ("rollout/ep_rew_mean", "rollout/ep_len_mean", "train/value_loss", "train/approx_kl", "time/fps"):
        assert k in rows[-1]
    m1 = PPO.load("dd.zip", 1)
    assert m1.n_updates == 640 and m1.n_steps == 2048 and m1.batch_size == 64
    # resume: ./dd.zip exists -> load, keep training (train.py:11-31), numbered run dir 2
    train.main(["--total-timesteps", "2048", "--quiet", "--save", "dd2"])
    assert os.path.isdir("tensorboard/drone_runs_2")
    m2 = PPO.load("dd2.zip", 1)
    assert m2.n_updates == 960
    assert not torch.equal(m1.params, m2.params)
    # env / curriculum state travels in the archive (the reference loses it: train.py rebuilds its envs)
    assert int(m2.batch.get_state("ep_num")["ep_num"][0]) > int(m1.batch.get_state("ep_num")["ep_num"][0]) > 2
    m1.close(); m2.close()


def test_train_driver_many_envs_and_trajectory_blocks(tmp_path, monkeypatch):
    from drone_rl_b200 import train
    monkeypatch.chdir(tmp_path)
    # 256 envs x (2048 // 256 = 8) steps per rollout, as the reference's resume branch scales n_steps (train.py:14)
    orig = train.TrajectoryCallback.__init__
    monkeypatch.setattr(train.TrajectoryCallback, "__init__",
                        lambda self, run_dir, **kw: orig(self, run_dir, record_interval=2, block_size=4))
    train.main(["--total-timesteps", str(2048 * 150), "--n-envs", "256", "--batch-size", "512", "--n-epochs", "2",
                "--precision", "tf32", "--quiet"])
    run = "tensorboard/drone_runs_1"
    rows = [json.loads(l) for l in open(f"{run}/progress.jsonl")]
    assert len(rows) == 150 and np.isfinite(rows[-1]["train/value_loss"])
    z = np.load(f"{run}/trajectory_block1.npz")
    assert "ep_2" in z.files and z["ep_2"].shape[1] == 3 and z["ep_2"].shape[0] >= 2
    assert abs(float(z["ep_2"][0, 2]) - 1.0) < 1e-6                       # episodes start at z = 1 (drone.py:57)
    assert os.path.isfile("ppo_drone_rel_obs_pos_reward.zip")            # train.py:70


def test_eval_script_records_a_gif(tmp_path, monkeypatch):
    """The reference's test.py flow: PPO.load('./dd.zip'), 100 deterministic steps in a DroneGymEnv with start_record /
    render / stop_record -> an animated GIF with one frame per step."""
    from PIL import Image
    from drone_rl_b200 import test as eval_script
    from drone_rl_b200 import train
    monkeypatch.chdir(tmp_path)
    train.main(["--total-timesteps", "2048", "--quiet", "--save", "dd"])
    eval_script.main(["--steps", "30"])
    g = Image.open("my_drone_run.gif")
    # Pillow merges identical consecutive frames (adding their durations): count the time, not the frames
    total = 0
    for k in range(g.n_frames):
        g.seek(k)
        total += g.info["duration"]
    assert g.is_animated and 20 <= g.n_frames <= 30 and total == 30 * 50 and g.size == (480, 480)
