"""Development aid: sampled Decima rollouts with multi-environment groups of the fused policy kernel
(SSB_FUSED_GROUP), checked for simulator errors; run under compute-sanitizer to look for out-of-bounds writes."""
import os
import os.path as osp
import sys

import numpy as np

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
from spark_sched_sim_b200.bank import synthetic_bank  # noqa: E402
from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

B = int(os.environ.get("REPRO_B", "64"))
steps = int(os.environ.get("REPRO_STEPS", "300"))
cfg = {"num_executors": 10, "job_arrival_cap": 8, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=synthetic_bank(0), decima_policy=True)
z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
env.set_decima_weights({k: z[k] for k in z.files})
env.reset_host(np.arange(B, dtype=np.uint64) + 5)
for k in range(steps):
    a, n = env.decima_policy()
    env.step(a, n)
    h = env.hdr()
    bad = (h["error"] != 0) & (h["error"] != 9)
    if bad.any():
        print("step", k, "errors", np.unique(h["error"]), "envs", np.flatnonzero(bad)[:10])
        break
    if (h["terminated"] != 0).all():
        print("all terminated at step", k)
        break
print("done", k)
