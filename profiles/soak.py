"""Soak run: long auto-reset rollouts, counting environments that ever report an error (a failed invariant, the
reference's AssertionError).  Usage on the GPU box: python profiles/soak.py"""
import os.path as osp
import sys
import time

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spark_sched_sim_b200.bank import synthetic_bank  # noqa: E402
from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

bank = synthetic_bank(0)


def fair(E, J, B, launches, K, tl_mean=0.0, max_jobs=None):
    cfg = {"num_executors": E, "job_arrival_cap": J, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, max_jobs=max_jobs or J)
    if tl_mean:
        env.set_mean_time_limit(tl_mean)
    env.reset_host((7 + np.arange(B)).astype(np.uint64))
    t0 = time.perf_counter()
    for _ in range(launches):
        env.rollout_fair(K, True, True, B)
    torch.cuda.synchronize()
    st, h = env.stats(), env.hdr()
    print(f"fair  E={E:3d} J={J:3d} B={B}: {st['decisions'] / 1e6:8.1f} M decisions, {st['events'] / 1e9:6.2f} G events, "
          f"{st['episodes']:7d} episodes in {time.perf_counter() - t0:5.1f} s; envs with errors: {int((h['error'] != 0).sum())}")


def decima(E, J, B, calls, K, tl_mean):
    cfg = {"num_executors": E, "job_arrival_cap": J, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=bank, decima_policy=True)
    z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
    env.set_decima_weights({k: z[k] for k in z.files})
    env.set_mean_time_limit(tl_mean)
    env.set_autoreset(True, B)
    env.reset_host((11 + np.arange(B)).astype(np.uint64))
    t0 = time.perf_counter()
    for _ in range(calls):
        env.rollout_decima(K)
    torch.cuda.synchronize()
    st, h = env.stats(), env.hdr()
    print(f"decima E={E:3d} J={J:3d} B={B}: {st['decisions'] / 1e6:8.1f} M decisions, {st['events'] / 1e9:6.2f} G events, "
          f"{st['episodes']:7d} episodes in {time.perf_counter() - t0:5.1f} s; envs with errors: "
          f"{int(((h['error'] != 0) & (h['error'] != 9)).sum())}; lgprob finite: {bool(torch.isfinite(env.pol_lgprob).all())}")


fair(10, 50, 4096, 300, 128)
fair(50, 200, 8192, 40, 64)
# continuous arrivals with stochastic time limits and NO job cap: the number of jobs of an episode is unbounded
# (limit ~ Exp(mean)), so max_jobs must be generous -- an episode that needs more fails with SSB_ENV_CAPACITY (10)
# and its env stops (with max_jobs = 256, 12.8x the mean, every env had met such an episode after ~24 per env)
fair(10, 0, 4096, 100, 128, tl_mean=2.0e6, max_jobs=1024)
fair(64, 30, 2048, 100, 128)
decima(10, 50, 4096, 40, 25, 2.0e7)
decima(50, 200, 8192, 10, 25, 2.0e7)
