"""GPU parity of ssb_discounted_returns / ssb_group_baselines (1) on the reference's recorded vectors and
(2) on real fused rollouts against the numpy oracle (oracle/learner.py).  Baselines are compared bit for bit
when fed the same returns; returns at 1e-12 relative (device exp vs numpy exp)."""
import numpy as np
import pytest
import torch

from test_learner_oracle import CASES, DIFF_CASES

pytestmark = pytest.mark.gpu


def pack(times, rewards, stride):
    from spark_sched_sim_b200 import _native as nat

    B = len(rewards)
    tr = np.zeros((B, stride), nat.TRANSITION_DTYPE)
    for i in range(B):
        n = len(rewards[i])
        tr["wall_time"][i, :n] = times[i][:n]
        tr["reward"][i, :n] = rewards[i]
    dev = torch.from_numpy(tr.view(np.uint8).reshape(-1)).cuda()
    num = torch.tensor([len(r) for r in rewards], dtype=torch.int32, device="cuda")
    final = torch.tensor([t[-1] for t in times], dtype=torch.float64, device="cuda")
    return dev, num, final


@pytest.mark.parametrize("name", sorted(CASES))
def test_reference_vectors(name):
    from spark_sched_sim_b200.returns import Baseline, ReturnsCalculator

    c = CASES[name]
    B, stride = len(c["lens"]), int(c["lens"].max()) + 3
    traj, num, final = pack(c["times"], c["rewards"], stride)
    ret = ReturnsCalculator(beta=c["beta"])(traj, num, final, stride)
    got = ret.cpu().numpy()
    for i in range(B):
        assert np.allclose(got[i, :c["lens"][i]], c["returns"][i], rtol=1e-12, atol=0.0), i
    # baselines from the reference's own returns: bit-exact
    ref_ret = torch.zeros_like(ret)
    for i in range(B):
        ref_ret[i, :c["lens"][i]] = torch.from_numpy(c["returns"][i]).cuda()
    base = Baseline(B // c["num_rollouts"], c["num_rollouts"])(traj, ref_ret, num).cpu().numpy()
    for i in range(B):
        assert np.array_equal(base[i, :c["lens"][i]], c["baselines"][i]), i


def test_on_fused_rollouts(bank):
    """Whole pipeline on the device: fair rollouts of 4 job sequences x 4 envs each (same seed within a group,
    as trainer.py:268-270 assigns them), returns and group baselines vs the numpy oracle."""
    from learner import discounted_returns, group_baselines
    from spark_sched_sim_b200 import _native as nat
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv
    from spark_sched_sim_b200.returns import Baseline, ReturnsCalculator

    S, R, K, beta = 4, 4, 700, 5e-3
    B = S * R
    cfg = {"num_executors": 10, "job_arrival_cap": 8, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0, "beta": beta}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank)
    seeds = (100 + np.arange(B) // R).astype(np.uint64)
    env.reset_host(seeds)
    # groups share the job sequence; FIFO vs fair partitioning alternates so that the rollouts differ
    traj = env.rollout_fair_traj(K, True, auto_reset=False)
    hdr = env.hdr()
    assert (hdr["terminated"] != 0).all() and (hdr["error"] == 0).all()
    st = env.stats_per_env()
    num = torch.from_numpy(st["decisions"].astype(np.int32)).cuda()
    final = torch.from_numpy(hdr["wall_time"].copy()).cuda()
    ret = ReturnsCalculator(beta=beta)(traj, num, final, K)
    base = Baseline(S, R)(traj, ret, num)
    tr = traj.cpu().numpy().view(nat.TRANSITION_DTYPE).reshape(B, K)
    n = st["decisions"].astype(int)
    ts = [np.concatenate([tr["wall_time"][i, :n[i]], [hdr["wall_time"][i]]]) for i in range(B)]
    o_ret = [discounted_returns(tr["reward"][i, :n[i]], ts[i], beta) for i in range(B)]
    g_ret = ret.cpu().numpy()
    for i in range(B):
        assert np.allclose(g_ret[i, :n[i]], o_ret[i], rtol=1e-12, atol=0.0), i
    o_base = group_baselines([t[:-1] for t in ts], [g_ret[i, :n[i]] for i in range(B)], R)
    g_base = base.cpu().numpy()
    for i in range(B):
        assert np.array_equal(g_base[i, :n[i]], o_base[i]), i


@pytest.mark.parametrize("name", sorted(DIFF_CASES))
def test_differential_returns_reference_vectors(name):
    """ssb_differential_returns over three consecutive calls of one calculator (the window carries over):
    avg_num_jobs and the returns equal the reference's ReturnsCalculator(buff_cap) bit for bit."""
    from spark_sched_sim_b200.returns import ReturnsCalculator

    calc = ReturnsCalculator(buff_cap=DIFF_CASES[name][0]["cap"])
    for c in DIFF_CASES[name]:
        B, stride = len(c["lens"]), int(c["lens"].max()) + 2
        traj, num, final = pack(c["times"], c["rewards"], stride)
        got = calc(traj, num, final, stride).cpu().numpy()
        assert float(calc.avg_num_jobs.item()) == c["avg_num_jobs"]
        for i in range(B):
            assert np.array_equal(got[i, :c["lens"][i]], c["returns"][i]), i


@pytest.mark.parametrize("n,clip,coeff,use_idx", [(1, 0.2, 0.04, False), (7, 0.2, 0.04, False), (5000, 0.2, 0.04, True),
                                                  (300000, 0.1, 0.0, False)])
def test_ppo_loss_head_matches_torch_autograd(n, clip, coeff, use_idx):
    """ssb_ppo_loss vs the trainer's `_compute_loss` (ppo.py:104-140) written in plain torch fp32 with autograd for
    the adjoint seeds.  Tolerance: 2e-5 relative on the four scalars (torch sums in f32, the kernel in f64),
    1e-5 relative on the per-sample gradients."""
    from spark_sched_sim_b200.ppo import PPOLoss

    g = torch.Generator(device="cuda").manual_seed(n)
    total = n * 2 if use_idx else n
    old = -3.0 * torch.rand(total, device="cuda", generator=g)
    new = (old + 0.3 * torch.randn(total, device="cuda", generator=g)).requires_grad_()
    ent = torch.rand(total, device="cuda", generator=g).requires_grad_()
    ret = -1e5 * torch.rand(total, device="cuda", generator=g, dtype=torch.float64)
    base = ret + 2e4 * torch.randn(total, device="cuda", generator=g, dtype=torch.float64)
    idx = torch.randperm(total, device="cuda", generator=g)[:n].to(torch.int32) if use_idx else None
    out, g_lp, g_en = PPOLoss(clip, coeff)(new.detach(), old, ent.detach(), ret, base, idx)
    out = out.cpu().numpy()
    if n == 1:  # std of one sample is nan in torch as well
        assert np.isnan(out[0]) and np.isnan(out[1])
        return
    sel = idx.long() if use_idx else slice(None)
    advgs = (ret - base)[sel].float()
    advgs = (advgs - advgs.mean()) / (advgs.std() + 1e-8)
    log_ratio = new[sel] - old[sel]
    ratio = log_ratio.exp()
    policy_loss = -torch.min(advgs * ratio, advgs * torch.clamp(ratio, 1 - clip, 1 + clip)).mean()
    entropy_loss = -ent[sel].mean()
    loss = policy_loss + coeff * entropy_loss
    kl = ((ratio - 1) - log_ratio).mean()
    loss.backward()
    want = np.array([loss.item(), policy_loss.item(), entropy_loss.item(), kl.item()])
    assert np.allclose(out, want, rtol=2e-5, atol=1e-7), (out, want)
    assert torch.allclose(g_lp, new.grad[sel], rtol=1e-5, atol=2e-6 / n)  # atol: f32 cancellation in torch's adv - mean
    assert torch.allclose(g_en, ent.grad[sel], rtol=1e-6, atol=0.0)


@pytest.mark.parametrize("n,max_norm", [(20802, 0.5), (20802, None), (5, 0.5), (1 << 20, 0.5)])
def test_adam_step_matches_torch(n, max_norm):
    """ssb_adam_step vs clip_grad_norm_ + torch.optim.Adam over 5 consecutive updates of one flat vector
    (20 802 = the Decima parameter count).  Tolerance 2e-6 relative + 1e-8 absolute on the parameters (torch's
    fused arithmetic order differs in the last bit), 1e-6 relative on the gradient norm."""
    from spark_sched_sim_b200.ppo import Adam

    g = torch.Generator(device="cuda").manual_seed(n)
    p0 = torch.randn(n, device="cuda", generator=g) * 0.1
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=3e-4)
    mine = p0.clone()
    adam = Adam(mine, lr=3e-4, max_grad_norm=max_norm)
    for it in range(5):
        grad = torch.randn(n, device="cuda", generator=g) * (10.0 if it % 2 == 0 else 1e-3)
        ref.grad = grad.clone()
        want_norm = torch.linalg.vector_norm(grad)
        if max_norm:
            torch.nn.utils.clip_grad_norm_([ref], max_norm)
        opt.step()
        norm = adam.step(grad)
        assert torch.allclose(norm[0], want_norm, rtol=1e-6)
        assert torch.allclose(mine, ref.detach(), rtol=2e-6, atol=1e-8), (it, (mine - ref.detach()).abs().max())
