"""Config 3 (16 384 envs x 200 jobs x 50 executors, Decima): where one decision's time goes -- policy call vs the
step kernel -- and what an event budget per step call does to the step kernel's time and to the share of envs that
reach their next decision.  Usage on the GPU box: python profiles/c3_step_split.py [envs]"""
import os.path as osp
import sys

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
cfg = {"num_executors": 50, "job_arrival_cap": 200, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0,
       "warmup_delay": 1000.0, "beta": 5e-3}
env = BatchedSparkSchedSimEnv(cfg, num_envs=B, decima_policy=True)
z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
env.set_decima_weights({k: z[k] for k in z.files})
env.set_mean_time_limit(2e7)
env.set_autoreset(True, B)
env.reset_host((1234 + np.arange(B)).astype(np.uint64))
env.rollout_fair(1500, True, True, B)
env.rollout_decima(10)
torch.cuda.synchronize()


def timed(fn, n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


print(f"B = {B}")
print(f"policy call            {timed(env.decima_policy, 10):8.3f} ms")
print(f"rollout_decima / dec.  {timed(lambda: env.rollout_decima(10), 2) / 10:8.3f} ms")
for budget in (0, 1024, 512, 256, 128, 64):
    ts, reached = [], []
    for _ in range(12):
        a, n = env.decima_policy()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.step(a, n, max_events=budget)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        reached.append(1.0 - float((env.hdr()["pending"] != 0).mean()))
    print(f"step, budget {budget:5d}:   {np.mean(ts[2:]):8.3f} ms   envs at a decision afterwards {np.mean(reached[2:]):.3f}")
