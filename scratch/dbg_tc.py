import sys, ctypes as C, numpy as np, torch
sys.path.insert(0, '.')
import drone_rl_b200 as drl
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
from tests.test_gpu_ppo import _fake_buffers, _rand_params
from oracle import ppo_oracle as po
np.set_printoptions(linewidth=200, precision=4, suppress=True)
m = 128
model = PPO(1024, n_steps=2, ent_coef=0.01)
model.params.copy_(_rand_params(9, 0.5).float().cuda())
obs, act, old_logp, adv, ret = _fake_buffers(model, 4)
P = lambda t: None if t is None else C.c_void_p(t.data_ptr())
dbg = torch.zeros(2, 128, 256, device='cuda')
_lib.check(model.lib.dronecu_ppo_debug_buffer(model._h, P(dbg)))
b = model.buf
def grad(tc):
    model._adv_stats.zero_()
    _lib.check(model.lib.dronecu_ppo_adv_stats(model._h, P(b.adv), None, 0, m, P(model._adv_stats), None))
    fn = model.lib.dronecu_ppo_grad_tc if tc else model.lib.dronecu_ppo_grad
    _lib.check(fn(model._h, P(model.params), P(b.obs), P(b.actions), P(b.logp), P(b.adv), P(b.ret), None, 0, m, 0.0, 1.0, P(model._adv_stats), P(model._grad), None))
    torch.cuda.synchronize()
    return model._grad.cpu().numpy().copy()
g = grad(True); g32 = grad(False)
d = dbg.cpu().numpy()[0]
off = po.offsets()
print("nonzero lanes per column block (WG0):")
for c0 in range(0, 256, 8):
    blk = d[:, c0:c0+8]
    nz = np.nonzero(np.abs(blk).sum(1))[0]
    print(c0, "lanes:", (nz.min(), nz.max(), len(nz)) if len(nz) else None, "absmax", np.abs(blk).max())
print("dZ2 rows 0,1 [0:8]", d[0:2,256:264], "H2", d[0:2,384:392], "dH1", d[0:2,320:328])
th = model.params.cpu().double()
pp = po.unpack(th)
x = b.obs.reshape(-1,15)[:m].cpu().double()
h1 = torch.tanh(x @ pp["pi.W1"].t() + pp["pi.b1"]); h2 = torch.tanh(h1 @ pp["pi.W2"].t() + pp["pi.b2"])
print("ref H2 rows 0,1 [0:8]", h2[0:2,0:8].numpy())
W2 = g32[off['pi.W2'][0]:off['pi.W2'][0]+4096].reshape(64,64)
print("ref dW2[0:4,0:8]\n", W2[0:4,0:8])
print("tmem lanes 0..3 cols 64..72\n", d[0:4,64:72])
print("tmem lanes 32..35 cols 64..72\n", d[32:36,64:72])
print("ref dW2[16:20,0:8]\n", W2[16:20,0:8])
print("ref dW2^T[0:4,0:8]\n", W2.T[0:4,0:8])
W3 = g32[off['pi.W3'][0]:off['pi.W3'][0]+256].reshape(4,64)
print("ref dW3[:, 0:4]^T\n", W3[:,0:4].T, "\ntmem cols 144..152 lanes 0..3\n", d[0:4,144:152])
b2 = g32[off['pi.b2'][0]:off['pi.b2'][0]+64]
print("ref db2[0:4]", b2[0:4], "tmem cols 152..160 lanes 0..3\n", d[0:4,152:160])
W1 = g32[off['pi.W1'][0]:off['pi.W1'][0]+960].reshape(64,15)
print("ref dW1[0:2]\n", W1[0:2], "\ntmem cols 128..144 lanes 0..1\n", d[0:2,128:144])
print("tc grad W2 block [0:2,0:8]", g[off['pi.W2'][0]:off['pi.W2'][0]+16])
