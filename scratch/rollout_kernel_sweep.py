"""env-steps/s of the three policy-rollout kernels (float32 thread-per-env, float32 warp-per-env, tcgen05) over the batch size:
where should dronecu_rollout_policy / PPO switch?"""
import sys, time, ctypes as C, torch
sys.path.insert(0, '.')
import drone_rl_b200 as drl
from drone_rl_b200 import _lib
from drone_rl_b200.ppo import PPO
lib = _lib.load()
K = 64
for n in (1, 64, 256, 1024, 2048, 4096, 8192, 16384, 65536):
    row = []
    for name, mode, prec in (("thread", 1, "fp32"), ("warp", 2, "fp32"), ("tc", 0, "tf32")):
        lib.dronecu_set_rollout_kernel(mode)
        m = PPO(drl.DroneBatch(n, drl.EnvConfig.single(), seed=1), n_steps=K, batch_size=max(64, n), seed=1, rollout_precision=prec)
        m.tc_min_envs = 0 if name == "tc" else 10 ** 9
        for _ in range(2):
            m.collect_rollouts()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 5
        for _ in range(reps):
            m.collect_rollouts()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        row.append(f"{name} {n * K / dt:10.3e} ({dt / K * 1e6:7.2f} us/step)")
        m.close()
    print(f"n = {n:6d}: " + " | ".join(row), flush=True)
lib.dronecu_set_rollout_kernel(0)
