import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
np.set_printoptions(linewidth=220)
# usage: tc_timing4.py [tiles_per_wg] [n_envs] [K]   (defaults: a small L2-resident buffer; pass 1048576 32 for the c5 shape,
# where the minibatch rows are scattered over a 3 GB rollout buffer)
argv = [a for a in sys.argv if not a.startswith('--')]
tiles_per_wg = int(argv[1]) if len(argv) > 1 else 8
m = 128 * 3 * 74 * tiles_per_wg
n = int(argv[2]) if len(argv) > 2 else 8192
K = int(argv[3]) if len(argv) > 3 else (m + n - 1) // n + 1
assert m <= n * K
model = PPO(n, n_steps=K, update_precision="bf16"); model.collect_rollouts()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
dbg = torch.zeros(1 << 20, device='cuda'); _lib.check(model.lib.dronecu_ppo_debug_buffer(model._h, P(dbg)))
b = model.buf; idx = torch.randperm(K * n, device='cuda')[:m]
if '--random' not in sys.argv:
    idx = idx.sort().values          # what dronecu_minibatch_partition produces: ascending rows inside a minibatch
idx = idx.to(torch.int32)
for rep in range(2):
    dbg.zero_(); model._adv_stats.zero_()
    _lib.check(model.lib.dronecu_ppo_adv_stats(model._h, P(b.adv), P(idx), 0, m, P(model._adv_stats), None))
    model.launch_grad(idx, 0, m, P(model._adv_stats), model._grad)
    torch.cuda.synchronize()
t = dbg.cpu().numpy().view(np.int64)
names = ["top", "w(S6p)", "h1", "gath", "wS1", "h2", "wS2", "h3", "wS3", "h4", "wS4", "h5", "wS5", "h6"]
t0 = t[0]
for it in (3, 4):
    print("tile", it, "absolute stamps per warp (cycles since start), columns:", names)
    for wq in range(4):
        a = t[256 * wq: 256 * wq + 256].reshape(16, 16)[it, :14] - t0
        print(" warp", wq, a)
iss = t[1024:1024 + 256].reshape(16, 16)
comp = t[0:256].reshape(16, 16)
hs = [2, 5, 7, 9, 11, 13]; ws = [4, 6, 8, 10, 12, None]
for it in (3, 4):
    print("tile", it, " step: hand-over -> issuer wake -> issuer committed -> compute sees done   (cycles since start)")
    for k in range(6):
        dn = comp[it, ws[k]] if ws[k] is not None else comp[it + 1, 1]
        print(f"   S{k+1}: {comp[it, hs[k]] - t0:7d} -> {iss[it, 2*k] - t0:7d} -> {iss[it, 2*k+1] - t0:7d} -> {dn - t0:7d}")
# all services in order with WG ids
ev = []
for w in range(3):
    a = t[1024 + 256 * w: 1024 + 256 * w + 256].reshape(16, 16)
    for it in range(min(tiles_per_wg, 16)):
        for st in range(6):
            ev.append((a[it, 2 * st] - t0, a[it, 2 * st + 1] - t0, w, it, st))
ev.sort()
print([ (int(a), int(b), f"WG{w}.t{it}.S{st+1}") for a, b, w, it, st in ev if 52000 < a < 75000])

print("S2 of WG0: [before wait, after wait, after fence, committed] relative to previous service end")
for it in (3, 4, 5):
    print(it, [int(x - t0) for x in (iss[it, 12], iss[it, 13], iss[it, 2], iss[it, 3])])
