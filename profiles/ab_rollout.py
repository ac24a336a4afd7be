"""A/B timing of the fused fair rollout (C2 unless overridden) for one build of the library.

Usage on the GPU box:  SSB_LIB=/path/to/libssb_x.so python profiles/ab_rollout.py
Env: AB_SEED_GROUP (G consecutive envs share a seed = run as twins; 1), AB_B (envs, 4096), AB_E (10), AB_J (50), AB_K (decisions per launch, 128), AB_PRE (untimed decisions per env first, 0), AB_ITERS (5)."""
import os
import os.path as osp
import sys

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

B = int(os.environ.get("AB_B", "4096"))
E = int(os.environ.get("AB_E", "10"))
J = int(os.environ.get("AB_J", "50"))
K = int(os.environ.get("AB_K", "128"))
iters = int(os.environ.get("AB_ITERS", "5"))
cfg = {"num_executors": E, "job_arrival_cap": J, "job_arrival_rate": 4.0e-5,
       "moving_delay": 2000.0, "warmup_delay": 1000.0}
env = BatchedSparkSchedSimEnv(cfg, num_envs=B)
G = int(os.environ.get("AB_SEED_GROUP", "1"))
env.reset_host((1234 + np.arange(B) // G).astype(np.uint64))
if int(os.environ.get("AB_PRE", "0")) > 0:  # untimed decisions first (mid-episode state, as bench.py's C4)
    env.rollout_fair(int(os.environ["AB_PRE"]), True, True, B)
for _ in range(3):
    env.rollout_fair(K, True, True, B)
torch.cuda.synchronize()
env.reset_stats()
ts = []
for _ in range(iters):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.rollout_fair(K, True, True, B)
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
st = env.stats()
ms = float(np.mean(ts))
print(f"{os.environ.get('SSB_LIB', 'libssb.so').split('/')[-1]:24s} B={B} E={E} J={J} G={G}: {ms:8.2f} ms/launch  "
      f"{st['decisions'] / (ms * iters) / 1e3:7.2f} M decisions/s  {st['events'] / (ms * iters) / 1e3:8.1f} M events/s  "
      f"errors={int((env.hdr()['error'] != 0).sum())}")
