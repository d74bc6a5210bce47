"""SB3 ``.zip`` interoperability for the policy (reference train.py:22 ``PPO.load('./dd.zip')``,
train.py:70 ``model.save(...)``, test.py:7).

A stable-baselines3 archive is a zip of
    data                    json: constructor arguments / hyper-parameters (class objects cloud-pickled)
    policy.pth              torch state_dict of ActorCriticPolicy
    policy.optimizer.pth    torch state_dict of its Adam optimiser
    pytorch_variables.pth   (empty for PPO)
    _stable_baselines3_version, system_info.txt

PARITY UNPINNED: SB3 is neither in /root/reference nor installed here; the layout above is the published
one.  What this module guarantees:
  * ``import_zip`` reads the two ``.pth`` members of a real SB3 archive (names in ``ppo.SB3_NAMES``; the
    optimiser state is mapped through ``policy.parameters()`` order) and the plain-json hyper-parameters;
  * ``export_zip`` writes an archive whose ``policy.pth`` / ``policy.optimizer.pth`` a real SB3 model takes
    with ``PPO.set_parameters(path)`` (that call skips ``data``).  ``data`` holds only plain json (no
    cloud-pickled classes), plus a ``dronecu`` member with the env / curriculum / RNG state the reference loses
    on resume (train.py:12-31 rebuilds its envs).
"""
from __future__ import annotations

import io
import json
import zipfile

import torch

from .ppo import SB3_NAMES, _SHAPES

# ActorCriticPolicy.parameters() order: the module's own parameter (log_std) first, then the sub-modules in
# registration order: mlp_extractor (policy_net, value_net), action_net, value_net
SB3_PARAM_ORDER = ["log_std",
                   "mlp_extractor.policy_net.0.weight", "mlp_extractor.policy_net.0.bias",
                   "mlp_extractor.policy_net.2.weight", "mlp_extractor.policy_net.2.bias",
                   "mlp_extractor.value_net.0.weight", "mlp_extractor.value_net.0.bias",
                   "mlp_extractor.value_net.2.weight", "mlp_extractor.value_net.2.bias",
                   "action_net.weight", "action_net.bias", "value_net.weight", "value_net.bias"]
_OURS = {v: k for k, v in SB3_NAMES.items()}
_SHAPE = dict(_SHAPES)


def _flat_from_named(named: dict) -> torch.Tensor:
    parts = []
    for name, shape in _SHAPES:
        t = named[SB3_NAMES[name]].detach().to(torch.float32).cpu()
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{SB3_NAMES[name]}: shape {tuple(t.shape)} != {tuple(shape)} (MlpPolicy 15-64-64 expected)")
        parts.append(t.reshape(-1))
    return torch.cat(parts)


def _named_from_flat(flat: torch.Tensor) -> dict:
    out, off = {}, 0
    for name, shape in _SHAPES:
        n = 1
        for d in shape:
            n *= d
        out[SB3_NAMES[name]] = flat[off:off + n].reshape(shape).clone()
        off += n
    return out


def export_zip(path: str, params: torch.Tensor, adam: torch.Tensor | None = None, adam_step: int = 0,
               hyper: dict | None = None, extra_state: dict | None = None) -> None:
    """params: flat [10697]; adam: flat [2 * 10697] (exp_avg | exp_avg_sq) or None."""
    params = params.detach().cpu().to(torch.float32)
    policy = _named_from_flat(params)
    ordered = {k: policy[k] for k in SB3_PARAM_ORDER}          # state_dict key order == parameters() order
    opt = {"state": {}, "param_groups": [{"lr": float((hyper or {}).get("learning_rate", 3e-4)), "betas": (0.9, 0.999),
                                          "eps": 1e-5, "weight_decay": 0, "amsgrad": False, "maximize": False,
                                          "foreach": None, "capturable": False, "differentiable": False,
                                          "fused": None, "params": list(range(len(SB3_PARAM_ORDER)))}]}
    if adam is not None and adam_step > 0:
        n = params.numel()
        m, v = _named_from_flat(adam[:n].cpu()), _named_from_flat(adam[n:].cpu())
        for idx, k in enumerate(SB3_PARAM_ORDER):
            opt["state"][idx] = {"step": torch.tensor(float(adam_step)), "exp_avg": m[k], "exp_avg_sq": v[k]}
    data = {"policy_class": "ActorCriticPolicy (MlpPolicy)", "net_arch": {"pi": [64, 64], "vf": [64, 64]},
            "activation_fn": "tanh", "observation_space": {"shape": [15], "dtype": "float32"},
            "action_space": {"shape": [4], "low": 0.0, "high": 7.3575, "dtype": "float32"}}
    data.update(hyper or {})

    def pth(obj):
        b = io.BytesIO()
        torch.save(obj, b)
        return b.getvalue()
    with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
        z.writestr("data", json.dumps(data, indent=1))
        z.writestr("policy.pth", pth(ordered))
        z.writestr("policy.optimizer.pth", pth(opt))
        z.writestr("pytorch_variables.pth", pth({}))
        z.writestr("_stable_baselines3_version", "dronecu (layout of stable-baselines3 2.x)")
        z.writestr("system_info.txt", "written by drone_rl_b200.sb3_zip.export_zip\n")
        if extra_state is not None:
            z.writestr("dronecu_state.pth", pth(extra_state))


def import_zip(path: str) -> dict:
    """-> {"params": flat [10697], "adam": flat [2*10697] or None, "adam_step": int, "hyper": dict, "extra": dict|None}"""
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        if "policy.pth" not in names:
            raise ValueError(f"{path}: not a stable-baselines3 archive (no policy.pth)")
        policy = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu", weights_only=False)
        params = _flat_from_named(policy)
        adam, step = None, 0
        if "policy.optimizer.pth" in names:
            opt = torch.load(io.BytesIO(z.read("policy.optimizer.pth")), map_location="cpu", weights_only=False)
            st = opt.get("state", {})
            if len(st) == len(SB3_PARAM_ORDER):
                keys = [k for k in policy.keys()]              # state_dict order == parameters() order (no buffers)
                m = {keys[i]: st[i]["exp_avg"] for i in range(len(keys))}
                v = {keys[i]: st[i]["exp_avg_sq"] for i in range(len(keys))}
                adam = torch.cat([_flat_from_named(m), _flat_from_named(v)])
                step = int(float(st[0]["step"]))
        hyper = {}
        if "data" in names:
            try:
                raw = json.loads(z.read("data"))
                hyper = {k: v for k, v in raw.items() if isinstance(v, (int, float, str, bool)) or v is None}
            except Exception:
                hyper = {}
        extra = None
        if "dronecu_state.pth" in names:
            extra = torch.load(io.BytesIO(z.read("dronecu_state.pth")), map_location="cpu", weights_only=False)
    return {"params": params, "adam": adam, "adam_step": step, "hyper": hyper, "extra": extra}
