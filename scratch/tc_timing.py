"""Print the per-step clock64 timeline of ppo_grad_tc_kernel (needs a -DDRONECU_TC_TIMING=1 build: DRONECU_LIB)."""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import drone_rl_b200 as drl
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
np.set_printoptions(linewidth=220, suppress=True)
mode = sys.argv[1] if len(sys.argv) > 1 else "tf32"
nwg = 3 if mode == "bf16" else 2
tiles_per_wg = 12
m = 128 * nwg * 74 * tiles_per_wg
n = 8192
K = (m + n - 1) // n + 1
model = PPO(n, n_steps=K, update_precision=mode)
model.collect_rollouts()
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
dbg = torch.zeros(2 * 74 * 2 * 128 * 256, device='cuda')
_lib.check(model.lib.dronecu_ppo_debug_buffer(model._h, P(dbg)))
b = model.buf
idx = torch.randperm(K * n, device='cuda')[:m].to(torch.int32)
for rep in range(3):
    dbg.zero_()
    model._adv_stats.zero_()
    _lib.check(model.lib.dronecu_ppo_adv_stats(model._h, P(b.adv), P(idx), 0, m, P(model._adv_stats), None))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    _lib.check((model.lib.dronecu_ppo_grad_bf16 if mode == "bf16" else model.lib.dronecu_ppo_grad_tc)(model._h, P(model.params), P(b.obs), P(b.actions), P(b.logp), P(b.adv), P(b.ret), P(idx), 0, m, 0.0, 1.0, P(model._adv_stats), P(model._grad), None))
    e1.record(); torch.cuda.synchronize()
print("kernel+reduce ms", e0.elapsed_time(e1), "tiles per WG", tiles_per_wg)
t = dbg.cpu().numpy().view(np.int64)
comp = t[:16 * 16].reshape(16, 16)[:tiles_per_wg, :14]
iss = t[1024:1024 + 16 * 16].reshape(16, 16)[:tiles_per_wg, :12]
t0 = comp[0, 0]
names = ["top", "w(S6p)", "h1", "gath", "wS1", "h2", "wS2", "h3", "wS3", "h4", "wS4", "h5", "wS5", "h6"]
print("compute thread: deltas between consecutive stamps (cycles)")
print("       " + " ".join(f"{x:>6s}" for x in names[1:]) + "   tile_total")
for i in range(tiles_per_wg):
    d = np.diff(comp[i])
    nxt = (comp[i + 1, 0] - comp[i, 0]) if i + 1 < tiles_per_wg else 0
    print(f"tile{i:2d} " + " ".join(f"{x:6d}" for x in d) + f"   {nxt:8d}")
print("issuer: per step [wait-return -> committed] (cycles) and gap from compute hand-over to issuer wake")
hs = [2, 5, 7, 9, 11, 13]
for i in range(tiles_per_wg):
    row = []
    for k in range(6):
        row.append(f"{iss[i, 2 * k] - comp[i, hs[k]]:5d}/{iss[i, 2 * k + 1] - iss[i, 2 * k]:5d}")
    print(f"tile{i:2d} " + "  ".join(row))
print("compute wait-return minus issuer commit (MMA execution + wake), per step")
ws = [4, 6, 8, 10, 12]
for i in range(tiles_per_wg):
    print(f"tile{i:2d} " + " ".join(f"{comp[i, ws[k]] - iss[i, 2 * k + 1]:6d}" for k in range(5)))
