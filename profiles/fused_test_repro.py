"""Development aid: the body of tests/test_gpu_decima_policy.py::test_policy_sampling_rollout_and_distribution with
diagnostics (which iteration / environments fail)."""
import os
import os.path as osp
import sys

import numpy as np

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
sys.path.insert(0, osp.join(REPO, "tests"))
from helpers import load_golden  # noqa: E402
from spark_sched_sim_b200.bank import synthetic_bank  # noqa: E402
from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

tr = load_golden("decima_e10_j8_s5_philox")
B = 2048
PART1 = os.environ.get("PART1", "1") == "1"
cfg = {"num_executors": 10, "job_arrival_cap": 8, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=synthetic_bank(0), max_jobs=10, tape_capacity=len(tr["tape"]) + 8,
                              decima_policy=True)
env.set_decima_weights({k: z[k] for k in z.files})
for b in range(B):
    env.load_trace(b, tr["job_t_arrival"], tr["job_template"], tr["tape"])
env.reset_host(np.arange(B, dtype=np.uint64) + 77)
if PART1:
    for k in range(60):
        stage_idx, _, num_exec = (int(x) for x in tr["pol_actions"][k])
        if k % 7 == 3:
            env.decima_policy()
        a, n = env.decima_policy(forced_stage=np.full(B, stage_idx, np.int32), forced_num_exec=np.full(B, num_exec, np.int32))
        env.step(a, n)
        h = env.hdr()
        assert (h["error"] == 0).all(), ("part 1", k, np.unique(h["error"]))
env.reset_host(np.arange(B, dtype=np.uint64) + 5)
for it in range(3000):
    hb = env.hdr().copy()
    a, n = env.decima_policy()
    act = env.pol_action.cpu().numpy()
    env.step(a, n)
    h = env.hdr()
    bad = np.flatnonzero((h["error"] != 0) & (h["error"] != 9))
    if len(bad):
        print("iteration", it, "bad envs", bad[:10], "errors", h["error"][bad[:10]])
        for b in bad[:4]:
            print(" env", b, "action", int(a[b]), int(n[b]), "pol_action", act[b], "before: N", hb["num_nodes"][b], "nsched",
                  hb["num_schedulable"][b], "ncommit", hb["num_committable_execs"][b], "term", hb["terminated"][b], "err", hb["error"][b])
        break
    if (h["terminated"] != 0).all():
        print("all terminated at", it)
        break
print("done")
