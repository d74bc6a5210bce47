"""Where does a PPO run go non-finite?  usage: nan_hunt.py [n_envs] [update_precision] [iterations]"""
import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
argv = [a for a in sys.argv if not a.startswith('--')]
n = int(argv[1]) if len(argv) > 1 else 1048576
up = argv[2] if len(argv) > 2 else "bf16"
its = int(argv[3]) if len(argv) > 3 else 10
K, mb = 32, 4
model = PPO(n, n_steps=K, batch_size=n * K // mb, rollout_precision="tf32", update_precision=up, cuda_graph=False)
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
orig = model._minibatch
state = {"k": 0, "bad": False}
def checked(index, first, m, stats=None):
    orig(index, first, m, stats)
    if not state["bad"]:
        g = model._grad.cpu().numpy(); p = model.params.cpu().numpy()
        if not (np.isfinite(g).all() and np.isfinite(p).all()):
            state["bad"] = True
            bg, bp = np.flatnonzero(~np.isfinite(g)), np.flatnonzero(~np.isfinite(p))
            print(f"update {state['k']}: non-finite grad entries {len(bg)} first {bg[:10]}, params {len(bp)} first {bp[:10]}; stats {g[10697:]}; adv stats {None if stats is None else stats.cpu().numpy()}")
    state["k"] += 1
if '--nosync' not in sys.argv:
    model._minibatch = checked
if '--events' in sys.argv:
    model.grad_events = []
for it in range(its):
    model.collect_rollouts()
    b = model.buf
    fin = {k: bool(torch.isfinite(getattr(b, k)).all()) for k in ("obs", "actions", "logp", "value", "reward", "adv", "ret")}
    model.train()
    lv = model.logger_values
    print(it, "buffers finite" if all(fin.values()) else fin, {k.split('/')[1]: v for k, v in lv.items() if k.split('/')[1] in ('policy_gradient_loss', 'value_loss', 'std')}, 'param checksum', float(model.params.double().sum()), flush=True)
    if state["bad"]:
        break
