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


def test_ppo_train_on_shuffled_samples(bank):
    """trainers/ppo.py:52-102 with the reference's mini-batches: the dataset is every (decision, env) sample of the
    store, shuffled and split into num_batches mini-batches that mix observations of different decisions
    (ssb_decima_snapshot_gather).  (1) a gathered mini-batch re-evaluates to exactly the log-probabilities stored at
    collection time; (2) its gradient equals the sum of the per-snapshot gradients with the seeds scattered to the
    samples' environments; (3) the epoch loop makes num_epochs * num_batches updates, starts at ratio 1 (approx KL 0)
    and stops early once the KL threshold is crossed."""
    import os.path as osp

    from helpers import GOLDEN_DIR
    from spark_sched_sim_b200 import ppo
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    B, K = 16, 10
    cfg = {"num_executors": 10, "job_arrival_cap": 6, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, decima_policy=True)
    z = np.load(osp.join(GOLDEN_DIR, "decima_model.npz"))
    w = {k: z[k] for k in z.files}
    env.set_decima_weights(w)
    env.reset_host(np.arange(B, dtype=np.uint64) + 400)
    for _ in range(7):  # a few decisions into the episodes
        a, n = env.decima_policy()
        env.step(a, n)
    store = ppo.RolloutStore(env, K).collect()
    assert bool(store.valid.all())
    gen = torch.Generator().manual_seed(3)
    # (1) + (2): one mini-batch of B samples drawn over all K * B
    pick = torch.randperm(K * B, generator=gen)[:B].cuda()
    ks, bs = (pick // B).int().contiguous(), (pick % B).int().contiguous()
    staging = torch.empty(env.decima_snapshot_bytes(), dtype=torch.uint8, device="cuda")
    lg, en = ppo._evaluate_slots(env, store, staging, ks, bs)
    assert torch.equal(lg, store.lgprob.reshape(-1)[pick])  # same weights, same observation: bit-identical
    g_lp = torch.randn(B, generator=gen).cuda()
    g_en = (0.1 * torch.randn(B, generator=gen)).cuda()
    grads_mb = torch.zeros(20802, dtype=torch.float32, device="cuda")
    env.decima_backward(g_lp, g_en, grads_mb)
    env.decima_snapshot_unload()
    grads_ref = torch.zeros(20802, dtype=torch.float32, device="cuda")
    for k in range(K):
        sel = torch.nonzero(ks == k).reshape(-1)
        if sel.numel() == 0:
            continue
        sl = torch.zeros(B, dtype=torch.float32, device="cuda")
        se = torch.zeros(B, dtype=torch.float32, device="cuda")
        sl[bs[sel].long()] = g_lp[sel]
        se[bs[sel].long()] = g_en[sel]
        env.decima_snapshot_load(store.snapshots[k])
        env.decima_evaluate(None, store.stage_sel[k].contiguous(), store.exec_sel[k].contiguous())
        env.decima_backward(sl, se, grads_ref)
        env.decima_snapshot_unload()
    scale = float(grads_ref.abs().max())
    assert scale > 0.1 and float((grads_mb - grads_ref).abs().max()) <= 3e-4 * scale
    # (3) the epoch loop
    flat = np.concatenate([w[k].reshape(-1) for k in env.DECIMA_PARAM_ORDER]).astype(np.float32)
    adam = ppo.Adam(torch.from_numpy(flat).cuda().contiguous(), lr=3e-4, max_grad_norm=0.5)
    loss_fn = ppo.PPOLoss(clip_range=0.2, entropy_coeff=0.04)
    returns = torch.randn(K, B, generator=gen, dtype=torch.float64).cuda()
    baselines = torch.zeros(K, B, dtype=torch.float64, device="cuda")
    res = ppo.ppo_train_samples(env, store, returns, baselines, loss_fn, adam, num_epochs=2, num_batches=3,
                                target_kl=None, generator=gen)
    assert res["num_updates"] == 6 and res["num_samples"] == K * B and res["batch_size"] == K * B // 3 + 1
    assert float((adam.params.cpu() - torch.from_numpy(flat)).abs().max()) > 1e-5
    res2 = ppo.ppo_train_samples(env, store, returns, baselines, loss_fn, adam, num_epochs=2, num_batches=3,
                                 target_kl=1e-9, generator=gen)
    assert res2["num_updates"] == 0 and res2["approx kl div"] > 1.5e-9  # the weights have moved: stops at the first check
    # a handle smaller than the mini-batch: chunks of B slots, same dataset
    res3 = ppo.ppo_train_samples(env, store, returns, baselines, loss_fn, adam, num_epochs=1, num_batches=1,
                                 target_kl=None, generator=gen)
    assert res3["num_updates"] == 1 and res3["batch_size"] == K * B + 1


def test_scheduler_plugin_evaluate_actions_and_update_parameters(bank):
    """schedulers.DecimaScheduler as a TrainableScheduler (schedulers/scheduler.py:21-54): evaluate_actions on stored
    samples, update_parameters(loss) = backward + clip_grad_norm_ + Adam, the new weights live in the policy."""
    import os.path as osp

    from helpers import GOLDEN_DIR
    from spark_sched_sim_b200 import ppo
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv
    from spark_sched_sim_b200.schedulers import DecimaScheduler

    B, K = 8, 6
    cfg = {"num_executors": 10, "job_arrival_cap": 5, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, decima_policy=True)
    sched = DecimaScheduler(num_executors=10, state_dict_path=osp.join(GOLDEN_DIR, "decima_model.npz"), opt_cls="Adam",
                            opt_kwargs={"lr": 3e-4}, max_grad_norm=0.5).bind(env)
    assert sched.device == env.device and sched.optim is not None
    with pytest.raises(ValueError):
        DecimaScheduler(num_executors=10, state_dict_path=osp.join(GOLDEN_DIR, "decima_model.npz"), opt_cls="SGD")
    env.reset_host(np.arange(B, dtype=np.uint64) + 9)
    store = ppo.RolloutStore(env, K).collect()
    gen = torch.Generator().manual_seed(1)
    idx = torch.randperm(K * B, generator=gen)[:B - 2].cuda()  # fewer samples than slots: the rest stay empty
    res = sched.evaluate_actions(store.samples(idx), store.actions(idx))
    assert torch.equal(res["lgprobs"], store.lgprob.reshape(-1)[idx])
    n = idx.numel()
    ret = torch.randn(n, generator=gen, dtype=torch.float64).cuda()
    out, g_lp, g_en = ppo.PPOLoss(0.2, 0.04)(res["lgprobs"].contiguous(), store.lgprob.reshape(-1)[idx].contiguous(),
                                              res["entropies"].contiguous(), ret, torch.zeros_like(ret))
    assert abs(float(out[3])) < 1e-6  # ratio 1 before the first update
    before = sched.optim.params.clone()
    sched.update_parameters(ppo.Loss(g_lp, g_en))
    assert float((sched.optim.params - before).abs().max()) > 1e-6 and sched.optim.num_steps == 1
    res2 = sched.evaluate_actions(store.samples(idx), store.actions(idx))
    assert float((res2["lgprobs"] - res["lgprobs"]).abs().max()) > 0  # the policy evaluates with the new weights
    sched.update_parameters(None)  # no loss: an optimiser step on zero gradients, the mini-batch is released
    a, c = env.decima_policy()     # the live observation is back in place
    env.step(a, c)
    assert (env.hdr()["error"] == 0).all()
