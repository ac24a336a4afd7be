"""Small workload for compute-sanitizer (memcheck / initcheck): Decima rollout on the tensor-core path with auto-reset,
and a two-slot (E = 50) fused fair rollout.  Usage on the GPU box:
    compute-sanitizer --tool memcheck python profiles/sanitizer_run.py
Round 1: memcheck 0 errors; initcheck 0 errors in kernels (after zeroing the weight blobs' padding) -- what it still
reports with the snapshot calls below are "cudaMemcpy source" records only: ssb_decima_snapshot copies the whole
observation slabs device-to-device, including the never-written tails beyond num_nodes / num_edges of each env."""
import sys, os.path as osp
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, torch
from spark_sched_sim_b200.bank import synthetic_bank
from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv
cfg = {"num_executors": 10, "job_arrival_cap": 6, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
B = 64
import os
R2_ONLY = os.environ.get("SAN_R2_ONLY") == "1"  # (racecheck / synccheck runs: only the round-2 part below)
z = np.load(osp.join('tests', 'golden', 'decima_model.npz'))
if not R2_ONLY:
  env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=synthetic_bank(0), decima_policy=True)
  env.set_decima_weights({k: z[k] for k in z.files})
  env.reset_host((np.arange(B) + 5).astype(np.uint64))
  env.set_autoreset(True, 64)
  tr = env.rollout_decima(60)
  # the policy's backward pass on the live observation and on a stored one
  snap = env.decima_snapshot()
  env.decima_policy()
  gw = torch.zeros(20802, device='cuda')
  env.decima_backward(torch.randn(B, device='cuda'), torch.randn(B, device='cuda'), gw)
  acts = env.pol_action.clone()
  env.decima_snapshot_load(snap)
  env.decima_evaluate(None, acts[:, 0].contiguous(), acts[:, 2].contiguous())
  env.decima_backward(torch.randn(B, device='cuda'), torch.randn(B, device='cuda'), gw)
  env.decima_snapshot_unload()
  assert torch.isfinite(gw).all() and gw.abs().sum() > 0
  torch.cuda.synchronize()
  h = env.hdr()
  print('decima ok', env.stats()['decisions'], (h['error'] != 0).sum())
  env2 = BatchedSparkSchedSimEnv({**cfg, "num_executors": 50, "job_arrival_cap": 12}, num_envs=32, bank=synthetic_bank(0))
  env2.reset_host((np.arange(32) + 9).astype(np.uint64))
  env2.rollout_fair(400, True, True, 32)
  torch.cuda.synchronize()
  print('e50 ok', env2.stats()['decisions'], (env2.hdr()['error'] != 0).sum())
  # learner-side kernels on the fused rollouts' buffers: returns (discounted, differential), group baseline, PPO loss
  # head, Adam step
  from spark_sched_sim_b200.ppo import Adam, PPOLoss
  from spark_sched_sim_b200.returns import Baseline, ReturnsCalculator
  env3 = BatchedSparkSchedSimEnv(cfg, num_envs=16, bank=synthetic_bank(0))
  env3.reset_host((np.arange(16) // 4 + 3).astype(np.uint64))
  K = 700
  traj = env3.rollout_fair_traj(K, True, auto_reset=False)
  num = torch.from_numpy(env3.stats_per_env()["decisions"].astype(np.int32)).cuda()
  final = torch.from_numpy(env3.hdr()["wall_time"].copy()).cuda()
  ret = ReturnsCalculator(beta=5e-3)(traj, num, final, K)
  diff = ReturnsCalculator(buff_cap=300)
  diff(traj, num, final, K); diff(traj, num, final, K)
  base = Baseline(4, 4)(traj, ret, num)
  n = 16 * K
  lp = torch.rand(n, device='cuda') - 2.0
  PPOLoss(0.2, 0.04)(lp, lp + 0.1, torch.rand(n, device='cuda'), ret.reshape(-1).contiguous(), base.reshape(-1).contiguous())
  prm = torch.randn(20802, device='cuda')
  Adam(prm, max_grad_norm=0.5).step(torch.randn(20802, device='cuda'))
  torch.cuda.synchronize()
  print('learner ok', int(num.sum()), float(diff.avg_num_jobs))
# ---- round 2: the three policy modes, Decima async rollouts, shuffled mini-batches (snapshot gather), the packed host
# observation and the executor history
from spark_sched_sim_b200 import ppo
for mode in ("0", "1"):
    os.environ["SSB_DECIMA_MODE"] = mode
    e = BatchedSparkSchedSimEnv(cfg, num_envs=24, bank=synthetic_bank(0), decima_policy=True, history_capacity=512)
    e.set_decima_weights({k: z[k] for k in z.files})
    e.reset_host((np.arange(24) + 70).astype(np.uint64))
    tr, num, el = e.rollout_decima_async(300, 6.0e5, 24)
    store = ppo.RolloutStore(e, 5).collect()
    flat = np.concatenate([z[k].reshape(-1) for k in e.DECIMA_PARAM_ORDER]).astype(np.float32)
    adam = ppo.Adam(torch.from_numpy(flat).cuda().contiguous(), lr=3e-4, max_grad_norm=0.5)
    res = ppo.ppo_train_samples(e, store, torch.randn(5, 24, dtype=torch.float64).cuda(),
                                torch.zeros(5, 24, dtype=torch.float64).cuda(), ppo.PPOLoss(0.2, 0.04), adam,
                                num_epochs=1, num_batches=3, target_kl=None)
    po = e.obs_host()
    hist = e.history(3)
    x = torch.randn(300, 53).cuda()
    y = e.decima_mlp_rows(5, x)
    torch.cuda.synchronize()
    print('mode', mode, 'ok', int(num.sum()), res["num_updates"], po["offsets"][-1].tolist(), len(hist["hist_t"]),
          bool(torch.isfinite(y).all()), int((e.hdr()['error'] != 0).sum()))
os.environ.pop("SSB_DECIMA_MODE")
