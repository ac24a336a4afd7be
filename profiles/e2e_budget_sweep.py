"""bench.py's e2e loop (one ssb_step_fair_host call per decision batch, host buffers) for several event budgets per
call.  C2, 4096 envs.  Usage on the GPU box: python profiles/e2e_budget_sweep.py"""
import os.path as osp
import sys
import time

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

cfg = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
B = 4096
for budget in (0, 1024, 512, 384, 256, 192, 128, 96, 64, 32):
    e = BatchedSparkSchedSimEnv(cfg, num_envs=B)
    a_pin, n_pin, a_nxt, n_nxt = (torch.empty(B, dtype=torch.int32).pin_memory() for _ in range(4))
    e.reset_host((1234 + np.arange(B)).astype(np.uint64))
    e.set_autoreset(True, B)
    a0, n0 = e.fair_actions(True)
    a_pin.copy_(a0); n_pin.copy_(n0)
    torch.cuda.synchronize()
    a_h, n_h, a_o, n_o = a_pin.numpy(), n_pin.numpy(), a_nxt.numpy(), n_nxt.numpy()
    calls = 0
    for phase in (0, 1):
        if phase:
            e.reset_stats(); t0 = time.perf_counter(); calls = 0
        while True:
            e.step_fair_host(a_h, n_h, a_o, n_o, True, max_events=budget)
            a_h, n_h, a_o, n_o = a_o, n_o, a_h, n_h
            calls += 1
            if calls >= (100 if not phase else 600):
                break
    dt = time.perf_counter() - t0
    dec = e.stats()["decisions"]
    print(f"budget {budget:5d}: {dec / dt / 1e6:6.2f} M decisions/s   {1e6 * dt / calls:6.1f} us per call   "
          f"{dec / calls / B:.3f} decisions per env and call")
    e.close()
