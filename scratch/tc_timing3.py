import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
tiles_per_wg = 8
m = 128 * 3 * 74 * tiles_per_wg
n = 8192; K = (m + n - 1) // n + 1
model = PPO(n, n_steps=K, update_precision="bf16"); model.collect_rollouts()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
dbg = torch.zeros(1 << 20, device='cuda'); _lib.check(model.lib.dronecu_ppo_debug_buffer(model._h, P(dbg)))
b = model.buf; idx = torch.randperm(K * n, device='cuda')[:m].to(torch.int32)
for rep in range(2):
    dbg.zero_(); model._adv_stats.zero_()
    _lib.check(model.lib.dronecu_ppo_adv_stats(model._h, P(b.adv), P(idx), 0, m, P(model._adv_stats), None))
    _lib.check(model.lib.dronecu_ppo_grad_bf16(model._h, P(model.params), P(b.obs), P(b.actions), P(b.logp), P(b.adv), P(b.ret), P(idx), 0, m, 0.0, 1.0, P(model._adv_stats), P(model._grad), None))
    torch.cuda.synchronize()
t = dbg.cpu().numpy().view(np.int64)
ev = []
for w in range(3):
    a = t[1024 + 256 * w: 1024 + 256 * w + 16 * 16].reshape(16, 16)
    for it in range(tiles_per_wg):
        for st in range(6):
            ev.append((a[it, 2 * st], a[it, 2 * st + 1], w, it, st))
ev.sort()
t0 = ev[0][0]; prev_end = t0
for s0, s1, w, it, st in ev[30:75]:
    print(f"t={s0 - t0:7d}  idle {s0 - prev_end:5d}  service {s1 - s0:5d}  WG{w} tile {it} S{st + 1}")
    prev_end = s1
