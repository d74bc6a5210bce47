"""Minibatch gradient at the c5 scale: bf16 tensor-core kernel vs the fp32 CUDA-core kernel on the same inputs, NaN scan,
run-to-run reproducibility.  usage: grad_scale_check.py [n_envs] [K] [minibatches]"""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1048576
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
mb = int(sys.argv[3]) if len(sys.argv) > 3 else 4
model = PPO(n, n_steps=K, batch_size=n * K // mb, rollout_precision="tf32", update_precision="bf16")
model.collect_rollouts()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
b, B, m = model.buf, n * K, n * K // mb
for name in ("obs", "actions", "logp", "value", "reward", "adv", "ret"):
    t = getattr(b, name)
    print(name, "finite" if bool(torch.isfinite(t).all()) else "NON-FINITE", float(t.abs().max()))
perm = torch.empty(B, dtype=torch.int32, device="cuda")
has_part = hasattr(model.lib, "dronecu_minibatch_partition")
if has_part:
    _lib.check(model.lib.dronecu_minibatch_partition(model._h, B, m, 1, 0, P(perm), None))
else:
    perm = torch.randperm(B, device="cuda").to(torch.int32)
stats = torch.zeros(mb, 3, dtype=torch.float64, device="cuda")
_lib.check(model.lib.dronecu_ppo_adv_stats_epoch(model._h, P(b.adv), P(perm), B, m, P(stats), None))
print("adv stats", stats.cpu().numpy())
def grad(fn, k):
    g = torch.zeros(_lib.GRAD_LEN, device="cuda")
    idx = perm[k * m:(k + 1) * m]
    model.launch_grad(idx, 0, m, P(stats[k]), g, precision=fn)
    torch.cuda.synchronize()
    return g.cpu().numpy()
for k in range(mb):
    g16 = grad("bf16", k)
    g16b = grad("bf16", k)
    g32 = grad("fp32", k)
    bad = ~np.isfinite(g16)
    scale = np.abs(g32[:10697]).max()
    print(f"minibatch {k}: bf16 non-finite entries {bad.sum()} at {np.flatnonzero(bad)[:8]}, reproducible {np.array_equal(g16, g16b, equal_nan=True)}, "
          f"max |bf16 - fp32| / max|fp32| = {np.nanmax(np.abs(g16[:10697] - g32[:10697])) / scale:.3e}, stats bf16 {g16[10697:10702]}, fp32 {g32[10697:10702]}")
