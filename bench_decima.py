"""Secondary benchmark (not the driver's bench line): Decima rollouts on the device, 1..N GPUs.

BASELINE.json configs[2]: Decima GNN policy rollouts with on-device observation construction.
Each decision = ssb_decima_policy (observation adapter + GNN + sampling) + ssb_step, both stream-
ordered on device tensors (no host round trip).  Weights: the reference's shipped model (fixture
tests/golden/decima_model.npz).  Prints one JSON line.

    python bench_decima.py [--envs 4096] [--executors 10] [--jobs 50] [--decisions 200] [--budget 0]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 bench_decima.py ...   (configs[4])

Under torchrun every rank owns `--envs` environments with disjoint seeds (weak scaling); the only exchange
is the all-reduce of the rollout statistics (NCCL), timed inside the region; time = max over ranks.
"""
from __future__ import annotations

import argparse
import json
import os.path as osp
import sys

import numpy as np

REPO = osp.dirname(osp.abspath(__file__))
sys.path.insert(0, REPO)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--envs", type=int, default=4096)
    ap.add_argument("--executors", type=int, default=10)
    ap.add_argument("--jobs", type=int, default=50)
    ap.add_argument("--decisions", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=50)
    ap.add_argument("--budget", type=int, default=0, help="max events per env per step (0 = run to decision)")
    ap.add_argument("--mean-time-limit", type=float, default=0.0,
                    help="StochasticTimeLimit mean in ms (config/decima_tpch.yaml: 2e7); 0 = episodes end by job cap only")
    args = ap.parse_args()

    import os

    import torch
    import torch.distributed as dist

    from spark_sched_sim_b200 import parallel
    from spark_sched_sim_b200.bank import synthetic_bank
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    cfg = {"num_executors": args.executors, "job_arrival_cap": args.jobs, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    B = args.envs
    env = BatchedSparkSchedSimEnv(cfg, num_envs=B, bank=synthetic_bank(0), decima_policy=True, device=dev)
    z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
    env.set_decima_weights({k: z[k] for k in z.files})
    seeds, _ = parallel.shard_seeds(1234, B, rank, world)
    env.reset_host(seeds)
    stats_vec = torch.zeros(8, dtype=torch.float64, device=dev)

    from spark_sched_sim_b200 import _native as nat

    env.set_autoreset(True, B * world)
    if args.mean_time_limit > 0:
        env.set_mean_time_limit(args.mean_time_limit)  # applies from the next (auto-)reset of every env
    chunk = 25  # decisions per ssb_rollout_decima call (one rollout-buffer slab)
    nb = B * chunk * nat.TRANSITION_DTYPE.itemsize
    traj_dev = torch.empty(nb, dtype=torch.uint8, device=dev)
    traj_pin = torch.empty(nb, dtype=torch.uint8).pin_memory()

    def decide(n, to_host=False):
        # rollout collection through the public call: policy + step on the device, transitions recorded
        for _ in range(max(1, n // chunk)):
            env.rollout_decima(chunk, max_events=args.budget, out=traj_dev, host=traj_pin if to_host else None)

    decide(args.warmup)
    env.reset_stats()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    decide(args.decisions)
    if world > 1:
        dist.all_reduce(stats_vec)  # rollout statistics: the path's only exchange
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    st = env.stats()
    hdr = env.hdr()
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        c = torch.tensor([float(st["decisions"]), float(st["events"])], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(c)
        ms = float(t.item())
        st = dict(st, decisions=int(c[0].item()), events=int(c[1].item()))
    # e2e: the same with every slab of transitions copied to pinned host memory (wall clock)
    import time
    env.reset_stats()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    decide(args.decisions, to_host=True)
    torch.cuda.synchronize()
    e2e_dt = time.perf_counter() - t0
    e2e_dec = env.stats()["decisions"]
    # time of the policy kernel alone
    p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    p0.record()
    for _ in range(20):
        env.decima_policy()
    p1.record()
    torch.cuda.synchronize()
    if world > 1:
        dist.destroy_process_group()
    if rank != 0:
        return
    print(json.dumps({
        "metric": "scheduling decisions/sec (Decima policy, batched envs)",
        "value": st["decisions"] / (ms * 1e-3), "unit": "decisions/s", "n_gpus": world, "scaling": "weak",
        "config": {"workload": f"{B} envs per GPU x ({args.jobs} jobs, {args.executors} executors), Decima policy "
                               "(shipped model.pt) on the tensor cores, ssb_decima_policy + ssb_step per decision",
                   "max_events_per_step": args.budget},
        "decisions": st["decisions"], "events": st["events"], "ms_total": ms,
        "policy_kernel_ms": p0.elapsed_time(p1) / 20,
        "e2e": {"value": e2e_dec * world / e2e_dt, "unit": "decisions/s (this rank x world)", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": nb, "path": "ssb_rollout_decima (25 decisions) -> D2H of the transition slab"},
        "env_errors": int(((hdr["error"] != 0) & (hdr["error"] != 9)).sum()),
        "finished_envs": int((hdr["terminated"] != 0).sum()),
    }))


if __name__ == "__main__":
    main()
