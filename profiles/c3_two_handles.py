"""Config 3 (Decima rollouts, 200 jobs x 50 executors): one handle of B envs vs TWO handles of B/2 envs on two CUDA
streams, driven by one host thread.  The step kernel (latency-bound, low issue use) of one half can run under the
policy's tile kernels (tensor / HBM-bound, with under-filled small levels) of the other half.
Usage on the GPU box: python profiles/c3_two_handles.py [envs] [handles...]"""
import os.path as osp
import sys

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
variants = [int(x) for x in sys.argv[2:]] or [1, 2, 4]
cfg = {"num_executors": 50, "job_arrival_cap": 200, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0,
       "warmup_delay": 1000.0, "beta": 5e-3}
z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
w = {k: z[k] for k in z.files}
chunk = 25
for H in variants:
    envs, streams = [], []
    for i in range(H):
        b = B // H
        e = BatchedSparkSchedSimEnv(cfg, num_envs=b, decima_policy=True)
        e.set_decima_weights(w)
        e.set_mean_time_limit(2e7)
        e.set_autoreset(True, B)
        e.reset_host((1234 + i * b + np.arange(b)).astype(np.uint64))
        e.rollout_fair(1500, True, True, B)
        envs.append(e)
        streams.append(torch.cuda.Stream())
    torch.cuda.synchronize()

    def slab():
        for e, s in zip(envs, streams):
            with torch.cuda.stream(s):
                e.rollout_decima(chunk)

    for _ in range(2):
        slab()
    torch.cuda.synchronize()
    for e in envs:
        e.reset_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for s in streams:
        s.wait_event(e0)
    K = 3
    for _ in range(K):
        slab()
    for s in streams:
        torch.cuda.current_stream().wait_stream(s)
    e1.record()
    torch.cuda.synchronize()
    dec = sum(e.stats()["decisions"] for e in envs)
    ms = e0.elapsed_time(e1)
    err = sum(int(((e.hdr()["error"] != 0) & (e.hdr()["error"] != 9)).sum()) for e in envs)
    print(f"{H} handle(s) x {B // H} envs: {dec / ms / 1e3:6.3f} M Decima decisions/s   {ms / K / chunk:7.3f} ms per decision "
          f"of all {B} envs   errors={err}")
    for e in envs:
        e.close()
    del envs
    torch.cuda.empty_cache()
