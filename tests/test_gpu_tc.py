"""tcgen05 / TMEM path of the policy MLP (tc_mlp.cuh) against the float64 oracle and the fp32 CUDA-core
path.  Stated tolerance: tf32 products (10-bit mantissa, inputs rounded to nearest) with fp32
accumulation and tanh.approx.f32 (max rel error 2^-11):
    layer-1 pre-activations  |err| <= 2^-10 * sum_k |w_k x_k|   (each product carries two 2^-11 roundings)
    layer-2 pre-activations  |err| <= 1.5e-2 * max(1, |ref|)      (adds the layer-1 error through tanh)
    mean / value             |err| <= 2e-2 * max(1, |ref|)
PARITY UNPINNED for the same reason as tests/test_gpu_ppo.py (SB3 is not in the reference tree)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import ppo_oracle as po  # noqa: E402


@pytest.fixture(scope="module")
def drl():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    import drone_rl_b200
    import drone_rl_b200.ppo  # noqa: F401
    return drone_rl_b200


def _params(seed):
    g = torch.Generator().manual_seed(seed)
    return po.init_params(seed, dtype=torch.float64) + 0.1 * torch.randn(po.N_PARAMS, generator=g, dtype=torch.float64)


def _pre_activations(theta, obs):
    p = po.unpack(theta)
    z1 = torch.cat([obs @ p["pi.W1"].t() + p["pi.b1"], obs @ p["vf.W1"].t() + p["vf.b1"]], 1)
    h1 = torch.tanh(z1)
    z2 = torch.cat([h1[:, :64] @ p["pi.W2"].t() + p["pi.b2"], h1[:, 64:] @ p["vf.W2"].t() + p["vf.b2"]], 1)
    return z1, z2


@pytest.mark.parametrize("B", [128, 1000, 128 * 300 + 5])
def test_tc_forward_matches_oracle(drl, B):
    from drone_rl_b200.ppo import PPO
    model = PPO(8, n_steps=4)
    theta = _params(3)
    model.params.copy_(theta.float().cuda())
    g = torch.Generator().manual_seed(B)
    obs = (torch.randn(B, 15, generator=g) * torch.tensor([1.0] * 3 + [2.0] * 3 + [1.0] * 3 + [4.0] * 3 + [1.0] * 3)).float()
    mean, value, d1, d2 = model.policy_forward(obs.cuda(), precision="tf32", debug=True)
    torch.cuda.synchronize()
    th = theta.float().double()
    z1, z2 = _pre_activations(th, obs.double())
    rm, rv, _ = po.forward(th, obs.double())

    def check(got, ref, rel, what):
        got, ref = got.cpu().double().numpy(), ref.numpy()
        err = np.abs(got - ref) / np.maximum(1.0, np.abs(ref))
        assert err.max() <= rel, f"{what}: max scaled err {err.max():.3g} at {np.unravel_index(err.argmax(), err.shape)}"
        return err.max()
    p = po.unpack(th)
    scale1 = torch.cat([obs.double().abs() @ p["pi.W1"].abs().t() + p["pi.b1"].abs(),
                        obs.double().abs() @ p["vf.W1"].abs().t() + p["vf.b1"].abs()], 1).numpy()
    err1 = np.abs(d1.cpu().double().numpy() - z1.numpy())
    assert (err1 <= 2.0 ** -10 * scale1 + 1e-6).all(), f"layer-1: worst err/bound {(err1 / (2.0 ** -10 * scale1 + 1e-6)).max():.3g}"
    e1 = (err1 / np.maximum(1.0, np.abs(z1.numpy()))).max()
    e2 = check(d2, z2, 1.5e-2, "layer-2 pre-activation")
    em = check(mean, rm, 2e-2, "mean")
    ev = check(value, rv, 2e-2, "value")
    # and against the fp32 CUDA-core kernel of the same library
    m32, v32 = model.policy_forward(obs.cuda())
    assert (mean - m32).abs().max().item() < 2e-2 and (value - v32).abs().max().item() < 4e-2
    print(f"B={B}: scaled errors L1 {e1:.2e} L2 {e2:.2e} mean {em:.2e} value {ev:.2e}")
    model.close()


def test_tc_rollout_matches_fp32_rollout_statistically(drl):
    """Same seed, same policy: the tf32 rollout's recorded values / log-probs agree with the fp32
    kernel's wherever the two trajectories have not yet diverged (first step exactly comparable)."""
    from drone_rl_b200.ppo import PPO
    n, K = 4096, 8
    outs = []
    for prec in ("fp32", "tf32"):
        m = PPO(drl.DroneBatch(n, drl.EnvConfig.single(), seed=5), n_steps=K, seed=5, rollout_precision=prec)
        m.params.copy_(_params(7).float().cuda())
        m.collect_rollouts()
        torch.cuda.synchronize()
        outs.append((m.buf.obs.clone(), m.buf.actions.clone(), m.buf.value.clone(), m.buf.logp.clone(), m.buf.adv.clone()))
        m.close()
    (o0, a0, v0, l0, adv0), (o1, a1, v1, l1, adv1) = outs
    assert torch.equal(o0[0], o1[0])                                  # same reset observation
    assert (a0[0] - a1[0]).abs().max().item() < 2e-2                  # same noise, mean within tf32 tolerance
    assert (v0[0] - v1[0]).abs().max().item() < 4e-2
    assert torch.equal(l0[0], l1[0])                                  # log-prob depends on the noise only
    assert torch.isfinite(adv1).all()
    assert (v0 - v1).abs().mean().item() < 2e-2
