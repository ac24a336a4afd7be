"""Time of one PPO mini-batch on the device (ppo.ppo_minibatch_update: snapshot load -> evaluate -> loss -> backward ->
Adam -> weight upload), the mini-batch = the 4096 observations of one stored snapshot at mid-episode, C2 shape.
Usage on the GPU box: python profiles/ppo_update_time.py"""
import os.path as osp
import sys
import time

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402
from spark_sched_sim_b200.ppo import Adam, PPOLoss, ppo_minibatch_update  # noqa: E402

cfg = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
B = 4096
env = BatchedSparkSchedSimEnv(cfg, num_envs=B, decima_policy=True)
z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
w = {k: z[k] for k in z.files}
env.set_decima_weights(w)
env.reset_host((1234 + np.arange(B)).astype(np.uint64))
env.set_autoreset(True, B)
env.rollout_decima(150)
snap = env.decima_snapshot()
env.decima_policy()
act, lg = env.pol_action.clone(), env.pol_lgprob.clone()
flat = torch.from_numpy(np.concatenate([w[k].astype(np.float32).reshape(-1) for k in w])).cuda()
adam = Adam(flat, lr=3e-4, max_grad_norm=0.5)
ret = -1e4 * torch.rand(B, device="cuda", dtype=torch.float64)
base = ret + 2e3 * torch.randn(B, device="cuda", dtype=torch.float64)
args = (snap, act[:, 0].contiguous(), act[:, 2].contiguous(), lg, ret, base, PPOLoss(0.2, 0.04), adam)
e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    info, stepped = ppo_minibatch_update(env, *args)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print(f"update {rep}: {1e3 * (t1 - t0):.1f} ms for {B} samples = {B / (t1 - t0) / 1e3:.0f} k samples/s; "
          f"loss {info['loss']:.5f} kl {info['approx_kl_div']:.2e} grad norm {float(adam.grad_norm):.4f}")
# forward / backward split
env.decima_snapshot_load(snap)
e[0].record(); lg2, en2 = env.decima_evaluate(None, args[1], args[2]); e[1].record()
gw = torch.zeros(20802, device="cuda")
g1, g2 = torch.randn(B, device="cuda") / B, torch.randn(B, device="cuda") / B
torch.cuda.synchronize()
torch.cuda.profiler.start()  # ncu --profile-from-start off ... captures exactly this backward pass
e[2].record(); env.decima_backward(g1, g2, gw); e[3].record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
env.decima_snapshot_unload()
print(f"forward (evaluate) {e[0].elapsed_time(e[1]):.2f} ms, backward {e[2].elapsed_time(e[3]):.2f} ms, "
      f"depth loop bound dmax = {int(env.dec_depth.max())} levels in this batch")
