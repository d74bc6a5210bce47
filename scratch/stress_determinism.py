"""Per-kernel run-to-run reproducibility at the c5 scale (1M envs x 32 steps): which kernel of the PPO iteration is racy?
usage: stress_determinism.py [reps]"""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
reps = int([a for a in sys.argv[1:] if not a.startswith('--')][0]) if len([a for a in sys.argv[1:] if not a.startswith('--')]) else 300
n, K, mb = 1048576, 32, 4
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
A = PPO(n, n_steps=K, batch_size=n * K // mb, rollout_precision="tf32", update_precision="bf16", seed=0)
Bm = PPO(n, n_steps=K, batch_size=n * K // mb, rollout_precision="tf32", update_precision="bf16", seed=0)
bad_roll = 0
for i in range(2 if "--bf16-only" in sys.argv else 20):
    A.collect_rollouts(); Bm.collect_rollouts()
    for name in ("obs", "actions", "logp", "value", "reward", "done", "adv", "ret"):
        if not torch.equal(getattr(A.buf, name), getattr(Bm.buf, name)):
            bad_roll += 1
            print(f"rollout {i}: buffer {name} differs between two identically seeded models")
print("rollout + GAE: mismatching buffers over 20 rollouts:", bad_roll, flush=True)
Bm.close()
b, B, m = A.buf, n * K, n * K // mb
perm0 = torch.empty(B, dtype=torch.int32, device="cuda"); perm = torch.empty_like(perm0)
_lib.check(A.lib.dronecu_minibatch_partition(A._h, B, m, 1, 0, P(perm0), None))
bad = 0
for i in range(50):
    _lib.check(A.lib.dronecu_minibatch_partition(A._h, B, m, 1, 0, P(perm), None))
    bad += int(not torch.equal(perm, perm0))
print("partition: mismatches over 50 runs:", bad, flush=True)
stats0 = torch.zeros(mb, 3, dtype=torch.float64, device="cuda"); stats = torch.zeros_like(stats0)
_lib.check(A.lib.dronecu_ppo_adv_stats_epoch(A._h, P(b.adv), P(perm0), B, m, P(stats0), None))
bad = 0
for i in range(50):
    _lib.check(A.lib.dronecu_ppo_adv_stats_epoch(A._h, P(b.adv), P(perm0), B, m, P(stats), None))
    bad += int(not torch.equal(stats, stats0))
print("adv stats: mismatches over 50 runs:", bad, flush=True)
kernels = (("bf16", A.lib.dronecu_ppo_grad_bf16, reps), ("tf32", A.lib.dronecu_ppo_grad_tc, reps // 3), ("fp32", A.lib.dronecu_ppo_grad, 12))
if "--bf16-only" in sys.argv:
    kernels = kernels[:1]
for name, fn, r in kernels:
    ref = [None] * mb
    bad, worst = 0, 0.0
    g = torch.zeros(_lib.GRAD_LEN, device="cuda")
    for i in range(r):
        k = i % mb
        A.launch_grad(perm0[k * m:(k + 1) * m], 0, m, P(stats0[k]), g, precision=name)
        if ref[k] is None:
            ref[k] = g.clone()
        elif not torch.equal(g, ref[k]):
            bad += 1
            d = (g - ref[k]).abs()
            worst = max(worst, float(d.max() / ref[k].abs().max()))
            if bad <= 5:
                idx = torch.nonzero(d > 0).flatten()
                print(f"  {name} launch {i} (minibatch {k}): {idx.numel()} entries differ, first {idx[:12].tolist()}, max rel-to-largest {float(d.max() / ref[k].abs().max()):.3e}, non-finite {int((~torch.isfinite(g)).sum())}")
    print(f"grad {name}: mismatching launches {bad} of {r}, worst deviation {worst:.3e}", flush=True)
