"""Host-side pieces of the training driver (reference train.py / helper.py / traj_tb.py) -- CPU only."""
import io
import json
import os
import zipfile

import numpy as np
import torch

from drone_rl_b200 import sb3_zip
from drone_rl_b200.ppo import POLICY_PARAMS, SB3_NAMES, init_policy_params, unpack_params
from drone_rl_b200.train import RunLogger, TrajectoryCallback, make_run_dir


def test_make_run_dir_numbers_like_the_reference(tmp_path):
    """helper.py:6-21: drone_runs_1, drone_runs_2 exist -> drone_runs_3; gaps are not reused; other names ignored."""
    root = tmp_path / "tensorboard"
    assert make_run_dir(str(root)).endswith("drone_runs_1")
    assert make_run_dir(str(root)).endswith("drone_runs_2")
    os.makedirs(root / "drone_runs_7")
    os.makedirs(root / "drone_runs_x")
    os.makedirs(root / "other_3")
    assert make_run_dir(str(root)).endswith("drone_runs_8")
    assert make_run_dir(str(root), prefix="other_").endswith("other_4")


def test_run_logger_writes_sb3_keys(tmp_path, capsys):
    lg = RunLogger(str(tmp_path), stdout=True)
    lg.dump({"rollout/ep_rew_mean": -1.5, "rollout/ep_len_mean": 31.0, "train/value_loss": 0.25, "time/iterations": 1}, 2048)
    lg.dump({"rollout/ep_rew_mean": -1.2, "rollout/ep_len_mean": 33.0, "train/value_loss": 0.2, "time/iterations": 2}, 4096)
    lg.close()
    rows = open(tmp_path / "progress.csv").read().strip().splitlines()
    assert rows[0] == "step,rollout/ep_rew_mean,rollout/ep_len_mean,train/value_loss,time/iterations" and len(rows) == 3
    js = [json.loads(l) for l in open(tmp_path / "progress.jsonl")]
    assert js[1]["step"] == 4096 and js[1]["rollout/ep_rew_mean"] == -1.2
    out = capsys.readouterr().out
    assert "| rollout/" in out and "ep_rew_mean" in out and "| train/" in out


class _FakeModel:
    def __init__(self, pos, done):
        class B:
            pass
        self.buf = B()
        obs = torch.zeros(pos.shape[0], 2, 15)
        obs[:, 0, 0:3] = torch.from_numpy(pos)
        self.buf.obs, self.buf.done = obs, torch.from_numpy(done.astype(np.uint8))[:, None].repeat(1, 2)
        self.num_timesteps = 1234


def test_trajectory_callback_segments_episodes(tmp_path):
    """traj_tb.py:31-73: every record_interval-th episode is buffered, every block_size episodes a block is written."""
    K = 60
    pos = np.arange(K * 3, dtype=np.float32).reshape(K, 3)
    done = np.zeros(K, bool)
    done[[9, 19, 29, 39, 49, 59]] = True                     # six episodes of ten steps
    from drone_rl_b200.tb_events import EventFileWriter, read_events
    tbw = EventFileWriter(str(tmp_path))
    cb = TrajectoryCallback(str(tmp_path), record_interval=2, block_size=4, tb=tbw)
    assert cb(_FakeModel(pos[:25], done[:25])) is True       # rollouts may cut an episode in two
    assert cb(_FakeModel(pos[25:], done[25:])) is True
    assert cb.episode_count == 6 and cb.blocks_written == 1
    z = np.load(tmp_path / "trajectory_block1.npz")
    assert set(z.files) == {"step", "ep_2", "ep_4"}
    assert np.array_equal(z["ep_2"], pos[10:20]) and np.array_equal(z["ep_4"], pos[30:40])
    assert len(cb.buffered) == 1 and np.array_equal(cb.buffered[0], pos[50:60])   # episode 6, next block
    assert os.path.isfile(tmp_path / "trajectory_block1.png")
    tbw.close()
    imgs = {}
    for ev in read_events(tbw.path):
        imgs.update(ev.get("images", {}))
    assert set(imgs) == {"Trajectory/Overlay_XY_block1", "Trajectory/Overlay_XZ_block1", "Trajectory/Overlay_YZ_block1"}   # traj_tb.py:50-66
    one = imgs["Trajectory/Overlay_XY_block1"]
    assert one["height"] == 360 and one["width"] == 360 and one["png"][:8] == b"\x89PNG\r\n\x1a\n"


def _sb3_like_archive(path, params, m, v, step):
    """An archive laid out the way stable-baselines3 writes it: state_dict in ``parameters()`` order, Adam state
    keyed by parameter index, ``data`` json with a non-json-able member stored as a dict."""
    named = {SB3_NAMES[k]: t.clone() for k, t in unpack_params(params).items()}
    nm, nv = ({SB3_NAMES[k]: t.clone() for k, t in unpack_params(x).items()} for x in (m, v))
    order = sb3_zip.SB3_PARAM_ORDER
    policy = {k: named[k] for k in order}
    opt = {"state": {i: {"step": torch.tensor(float(step)), "exp_avg": nm[k], "exp_avg_sq": nv[k]} for i, k in enumerate(order)},
           "param_groups": [{"lr": 3e-4, "params": list(range(len(order)))}]}

    def pth(o):
        b = io.BytesIO(); torch.save(o, b); return b.getvalue()
    with zipfile.ZipFile(path, "w") as z:
        z.writestr("data", json.dumps({"n_steps": 2048, "batch_size": 64, "gamma": 0.99, "policy_class": {":type:": "x", ":serialized:": "abc"}}))
        z.writestr("policy.pth", pth(policy))
        z.writestr("policy.optimizer.pth", pth(opt))
        z.writestr("pytorch_variables.pth", pth({}))
        z.writestr("_stable_baselines3_version", "2.3.0")


def test_sb3_zip_import_and_export_round_trip(tmp_path):
    g = torch.Generator().manual_seed(3)
    params = init_policy_params(5) + 0.01 * torch.randn(POLICY_PARAMS, generator=g)
    m, v = torch.randn(POLICY_PARAMS, generator=g), torch.rand(POLICY_PARAMS, generator=g)
    p = str(tmp_path / "dd.zip")
    _sb3_like_archive(p, params, m, v, 320)
    z = sb3_zip.import_zip(p)
    assert torch.equal(z["params"], params) and z["adam_step"] == 320
    assert torch.equal(z["adam"], torch.cat([m, v]))
    assert z["hyper"] == {"n_steps": 2048, "batch_size": 64, "gamma": 0.99} and z["extra"] is None
    # ours -> archive -> back, including the env / curriculum extras
    q = str(tmp_path / "ours.zip")
    extra = {"num_timesteps": 4096, "n_updates": 640, "env_state": {"ep_num": np.arange(4)}, "env_global_step": 77}
    sb3_zip.export_zip(q, params, torch.cat([m, v]), 640, {"n_steps": 128, "learning_rate": 1e-3}, extra)
    names = set(zipfile.ZipFile(q).namelist())
    assert {"data", "policy.pth", "policy.optimizer.pth", "pytorch_variables.pth", "_stable_baselines3_version"} <= names
    z2 = sb3_zip.import_zip(q)
    assert torch.equal(z2["params"], params) and torch.equal(z2["adam"], torch.cat([m, v])) and z2["adam_step"] == 640
    assert z2["hyper"]["n_steps"] == 128 and z2["extra"]["env_global_step"] == 77
    # the policy member is what SB3's set_parameters() feeds to policy.load_state_dict: SB3 names, parameters() order
    sd = torch.load(io.BytesIO(zipfile.ZipFile(q).read("policy.pth")), weights_only=False)
    assert list(sd) == sb3_zip.SB3_PARAM_ORDER
    assert sd["action_net.weight"].shape == (4, 64) and sd["mlp_extractor.value_net.0.weight"].shape == (64, 15)
    # a torch module with SB3's structure accepts it
    class Extractor(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.policy_net = torch.nn.Sequential(torch.nn.Linear(15, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh())
            self.value_net = torch.nn.Sequential(torch.nn.Linear(15, 64), torch.nn.Tanh(), torch.nn.Linear(64, 64), torch.nn.Tanh())
    class Policy(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.mlp_extractor = Extractor()
            self.action_net = torch.nn.Linear(64, 4)
            self.log_std = torch.nn.Parameter(torch.zeros(4))
            self.value_net = torch.nn.Linear(64, 1)
    pol = Policy()
    pol.load_state_dict(sd, strict=True)
    # forward of that module == the oracle's forward of the flat vector
    from oracle import ppo_oracle as po
    x = torch.randn(7, 15, generator=g)
    mean, value, _ = po.forward(params, x)
    assert torch.allclose(pol.action_net(pol.mlp_extractor.policy_net(x)), mean, atol=1e-6)
    assert torch.allclose(pol.value_net(pol.mlp_extractor.value_net(x)).squeeze(-1), value, atol=1e-6)


def test_tensorboard_event_file_format(tmp_path):
    """The scalar event file: CRC-32C known answer, TFRecord framing, Event / Summary protobuf subset round trip."""
    import glob
    import struct
    from drone_rl_b200 import tb_events as tb
    assert tb.crc32c(b"123456789") == 0xE3069283                      # the CRC-32C check value
    assert tb.crc32c(b"") == 0 and tb.crc32c(bytes(32)) == 0x8A9136AA  # RFC 3720 B.4: 32 bytes of zeros
    lg = RunLogger(str(tmp_path), stdout=False)
    lg.dump({"rollout/ep_rew_mean": -1.5, "train/value_loss": 0.25, "time/iterations": 1, "note": "skipped"}, 2048)
    lg.dump({"rollout/ep_rew_mean": -1.25, "train/value_loss": 0.125, "time/iterations": 2}, 300000)
    lg.close()
    (path,) = glob.glob(str(tmp_path / "events.out.tfevents.*"))
    raw = open(path, "rb").read()
    (n0,) = struct.unpack("<Q", raw[:8])
    assert raw[12:12 + n0].endswith(b"brain.Event:2")                 # first record: the file-version event
    ev = tb.read_events(path)
    assert ev[0]["file_version"] == "brain.Event:2" and len(ev) == 3
    assert ev[1]["step"] == 2048 and ev[2]["step"] == 300000
    assert ev[1]["scalars"] == {"rollout/ep_rew_mean": -1.5, "train/value_loss": 0.25, "time/iterations": 1.0}
    assert ev[2]["scalars"]["train/value_loss"] == 0.125


def test_pillow_renderer_scene_and_gif(tmp_path):
    """render.py: the reference's scene (drone.py:205-241) drawn with Pillow -- target green, arms purple, centre red,
    motors blue -- and a recording saved as an animated GIF; geometry of the motors as drone.py:222-229."""
    from PIL import Image
    from drone_rl_b200 import render
    m = render.motor_positions([1.0, 2.0, 3.0], [0.0, 0.0, 0.0], 0.5)
    a = 0.5 / np.sqrt(2)
    assert np.allclose(m, [[1 + a, 2 + a, 3], [1 - a, 2 + a, 3], [1 - a, 2 - a, 3], [1 + a, 2 - a, 3]])
    R = render.rotation_matrix([0.3, -0.2, 1.1])
    assert np.allclose(R @ R.T, np.eye(3), atol=1e-12) and abs(np.linalg.det(R) - 1) < 1e-12
    assert np.allclose(R[:, 2], [np.cos(1.1) * np.sin(-0.2) * np.cos(0.3) + np.sin(1.1) * np.sin(0.3),
                                 np.sin(1.1) * np.sin(-0.2) * np.cos(0.3) - np.cos(1.1) * np.sin(0.3),
                                 np.cos(-0.2) * np.cos(0.3)])                       # the column the dynamics use (drone.py:170-172)
    fr = render.FrameRenderer(size=240)
    img = fr.draw([0.0, 0.0, 2.0], [0.1, 0.2, 0.3], [1.0, -1.0, 3.0], 2.0)          # long arms: visible between the dots
    px = np.asarray(img).reshape(-1, 3)
    has = lambda c: bool((px == np.array(c)).all(1).any())
    assert has((0, 160, 0)) and has((128, 0, 128)) and has((220, 0, 0)) and has((0, 0, 220))
    img_nan = fr.draw([np.nan, 0, 0], [0, 0, 0], [0, 0, 1], 0.5)                    # a diverged env still renders the box + target
    assert bool((np.asarray(img_nan).reshape(-1, 3) == np.array((0, 160, 0))).all(1).any())
    fb = render.FrameRenderer(size=240, xlim=(-20, 20), ylim=(-20, 20), zlim=(0, 20))
    imgb = np.asarray(fb.draw_batch(np.array([[0, 0, 1.0], [3, -2, 5.0], [np.nan, 0, 0]]), [0, 0, 10.0])).reshape(-1, 3)
    assert bool((imgb == np.array((0, 160, 0))).all(1).any()) and bool((imgb == np.array((220, 0, 0))).all(1).any())
    rec = render.Recorder(str(tmp_path / "run.mp4"), fps=20)                        # the reference's default name ends in .mp4
    for k in range(5):
        rec.grab(fr.draw([0.0, 0.0, 1.0 + 0.2 * k], [0, 0, 0.1 * k], [0, 0, 3.0], 0.5))
    rec.finish()
    g = Image.open(tmp_path / "run.gif")
    assert g.is_animated and g.n_frames == 5 and g.info["duration"] == 50


class _SB3ShapedPolicy(torch.nn.Module):
    """A genuine torch module registered the way SB3 2.x's ``ActorCriticPolicy._build`` registers its members for
    ``MlpPolicy`` on Box(15) -> Box(4): ``mlp_extractor`` (policy_net / value_net: Linear-Tanh-Linear-Tanh), then
    ``action_net`` + the ``log_std`` parameter (DiagGaussianDistribution.proba_distribution_net), then ``value_net``.
    ``parameters()`` therefore yields log_std first (the module's own parameter), then the children in that order."""

    def __init__(self):
        super().__init__()
        nn = torch.nn

        class MlpExtractor(nn.Module):
            def __init__(self):
                super().__init__()
                self.policy_net = nn.Sequential(nn.Linear(15, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())
                self.value_net = nn.Sequential(nn.Linear(15, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())
        self.mlp_extractor = MlpExtractor()
        self.action_net = nn.Linear(64, 4)
        self.log_std = nn.Parameter(torch.zeros(4))
        self.value_net = nn.Linear(64, 1)

    def forward(self, x):
        return self.action_net(self.mlp_extractor.policy_net(x)), self.value_net(self.mlp_extractor.value_net(x)).flatten()


def test_sb3_zip_from_a_genuine_module_and_optimizer(tmp_path):
    """The closest thing to a real SB3 archive that can be built without SB3: ``policy.pth`` = ``state_dict()`` of a genuine
    nn.Module with SB3's member names, ``policy.optimizer.pth`` = ``state_dict()`` of a genuine ``torch.optim.Adam`` that
    has taken steps on it (layout of stable-baselines3 2.x ``save_to_zip_file``).  import_zip must map both onto the flat
    vectors; export_zip must write members the genuine module / optimiser load back strictly."""
    from oracle import ppo_oracle as po
    torch.manual_seed(9)
    pol = _SB3ShapedPolicy()
    assert [n for n, _ in pol.named_parameters()] == sb3_zip.SB3_PARAM_ORDER          # SB3's parameters() order
    assert list(pol.state_dict()) == sb3_zip.SB3_PARAM_ORDER
    opt = torch.optim.Adam(pol.parameters(), lr=3e-4, eps=1e-5)
    x, a = torch.randn(32, 15), torch.randn(32, 4)
    for _ in range(3):                                     # three genuine Adam steps: non-trivial exp_avg / exp_avg_sq / step
        opt.zero_grad()
        mean, value = pol(x)
        (((mean - a) ** 2).mean() + (value ** 2).mean() + pol.log_std.sum() ** 2).backward()
        opt.step()

    def pth(o):
        b = io.BytesIO(); torch.save(o, b); return b.getvalue()
    p = str(tmp_path / "genuine.zip")
    with zipfile.ZipFile(p, "w") as z:
        z.writestr("data", json.dumps({"n_steps": 2048, "batch_size": 64, "n_epochs": 10, "gamma": 0.99, "gae_lambda": 0.95,
                                       "learning_rate": 0.0003, "num_timesteps": 6144, "_n_updates": 30,
                                       "policy_class": {":type:": "<class 'abc.ABCMeta'>", ":serialized:": "gAWV..."},
                                       "clip_range": {":type:": "<class 'function'>", ":serialized:": "gAWV..."}}))
        z.writestr("policy.pth", pth(pol.state_dict()))
        z.writestr("policy.optimizer.pth", pth(opt.state_dict()))
        z.writestr("pytorch_variables.pth", pth(None))
        z.writestr("_stable_baselines3_version", "2.3.2")
        z.writestr("system_info.txt", "- OS: Linux\n")
    got = sb3_zip.import_zip(p)
    named = {k: v for k, v in pol.state_dict().items()}
    flat = torch.cat([named[SB3_NAMES[k]].reshape(-1) for k, _ in po.SHAPES])
    assert torch.equal(got["params"], flat) and got["adam_step"] == 3
    st = opt.state_dict()["state"]
    for half, key in ((0, "exp_avg"), (1, "exp_avg_sq")):
        by_name = {n: st[i][key] for i, n in enumerate(sb3_zip.SB3_PARAM_ORDER)}
        want = torch.cat([by_name[SB3_NAMES[k]].reshape(-1) for k, _ in po.SHAPES])
        assert torch.equal(got["adam"][half * POLICY_PARAMS:(half + 1) * POLICY_PARAMS], want)
    assert got["hyper"]["n_steps"] == 2048 and got["hyper"]["num_timesteps"] == 6144 and "clip_range" not in got["hyper"]
    mean, value = pol(x)
    om, ov, _ = po.forward(got["params"], x)
    assert torch.allclose(om, mean, atol=1e-6) and torch.allclose(ov, value, atol=1e-6)
    # and back: what export_zip writes loads strictly into the genuine module AND the genuine optimiser, and the next Adam
    # step of that optimiser equals the oracle's clip-free Adam step on the flat vectors
    q = str(tmp_path / "back.zip")
    sb3_zip.export_zip(q, got["params"], got["adam"], got["adam_step"], {"learning_rate": 3e-4})
    pol2 = _SB3ShapedPolicy()
    opt2 = torch.optim.Adam(pol2.parameters(), lr=3e-4, eps=1e-5)
    zf = zipfile.ZipFile(q)
    pol2.load_state_dict(torch.load(io.BytesIO(zf.read("policy.pth")), weights_only=False), strict=True)
    opt2.load_state_dict(torch.load(io.BytesIO(zf.read("policy.optimizer.pth")), weights_only=False))
    for (n1, p1), (n2, p2) in zip(pol.named_parameters(), pol2.named_parameters()):
        assert n1 == n2 and torch.equal(p1, p2)
    for o in (opt, opt2):
        o.zero_grad()
    for m_ in (pol, pol2):
        mean, value = m_(x)
        (((mean - a) ** 2).mean() + (value ** 2).mean()).backward()
    opt.step(); opt2.step()
    for p1, p2 in zip(pol.parameters(), pol2.parameters()):
        assert torch.equal(p1, p2), "the re-imported optimiser state must continue the run bit for bit"
