"""Where the time of one host-driven step() call goes (bench.py's e2e loop), C2, 4096 envs.
Usage on the GPU box: python profiles/e2e_breakdown.py [budget]"""
import os.path as osp
import sys
import time

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

budget = int(sys.argv[1]) if len(sys.argv) > 1 else 192
cfg = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
B = 4096
env = BatchedSparkSchedSimEnv(cfg, num_envs=B)
env.reset_host((1234 + np.arange(B)).astype(np.uint64))
env.set_autoreset(True, B)
a_pin = torch.empty(B, dtype=torch.int32).pin_memory()
n_pin = torch.empty(B, dtype=torch.int32).pin_memory()
t_pol = t_step = 0.0
dec0 = 0
calls = 300
for k in range(calls + 50):
    if k == 50:
        env.reset_stats(); torch.cuda.synchronize(); t_pol = t_step = 0.0; pend = resets = 0; t_all = time.perf_counter()
    t0 = time.perf_counter()
    a, n = env.fair_actions(True)
    a_pin.copy_(a, non_blocking=True); n_pin.copy_(n, non_blocking=True)
    torch.cuda.current_stream().synchronize()
    t1 = time.perf_counter()
    h = env.step_host(a_pin.numpy(), n_pin.numpy(), max_events=budget)
    t2 = time.perf_counter()
    t_pol += t1 - t0; t_step += t2 - t1
    if k >= 50:
        pend += int(h["pending"].sum()); resets += int(h["was_reset"].sum())
t_all = time.perf_counter() - t_all
dec = env.stats()["decisions"]
print(f"budget {budget}: {calls} calls, {1e6 * t_all / calls:.0f} us per call = policy+D2H+sync {1e6 * t_pol / calls:.0f} us + "
      f"step_host {1e6 * t_step / calls:.0f} us; decisions per call {dec / calls:.0f} of {B} (pending {pend / calls:.0f}, "
      f"resets {resets / calls:.1f}); {dec / t_all / 1e6:.2f} M decisions/s")
