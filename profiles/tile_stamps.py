"""clock64() stamps of CTA 0 in each tensor-core tile kernel of one Decima policy call (-DSSB_PROFILE build).
Usage on the GPU box: python profiles/tile_stamps.py"""
import ctypes as C
import os
import os.path as osp
import sys

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402

lib = osp.join(REPO, "spark-sched-sim_b200", "_lib", "libssb_prof.so")
os.environ["SSB_LIB"] = lib
import torch  # noqa: E402

from spark_sched_sim_b200 import _native as nat  # noqa: E402
from spark_sched_sim_b200.bank import synthetic_bank  # noqa: E402
from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

cfg = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
B = 4096
env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=synthetic_bank(0), decima_policy=True)
z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
env.set_decima_weights({k: z[k] for k in z.files})
env.reset_host((1234 + np.arange(B)).astype(np.uint64))
for _ in range(60):
    a, c = env.decima_policy()
    env.step(a, c)
ptr = C.c_void_p()
nat.check(env.L.ssb_get_debug_counters(env._h, C.byref(ptr)), "dbg")
prof = env._view(ptr.value, B * 16 * 8, torch.int64).view(B, 16)
prof.zero_()
torch.cuda.synchronize()
env.decima_policy()
torch.cuda.synchronize()
pp = prof.cpu().numpy().reshape(-1)[:8 * 16].reshape(8, 16)
names = ["PREP", "SINK", "MSG(level 0)", "RCV(level 0)", "DAG", "GLOB", "STAGE", "EXEC"]
lab = ["entry->weights loaded", "tmem alloc+sync", "(loop entry)", "gather", "write A+sync", "L1 mma+wait", "epi1+sync",
       "L2 mma+wait", "rest of tile 0", "remaining tiles+dealloc"]
for st in range(8):
    t = pp[st]
    d = np.diff(t[:10]).astype(np.int64)
    print(f"{names[st]:14s} total {int(t[9] - t[0]):7d} cyc : " + "  ".join(f"{lab[i]}={int(d[i])}" for i in range(9)))

