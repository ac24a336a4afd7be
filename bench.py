"""bench.py -- scheduling decisions/sec of the batched env on N B200s (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[1], "C2"): 4096 environments per GPU, 50 jobs x 10 executors,
fair scheduler (RoundRobinScheduler(10, dynamic_partition=True)), synthetic TPC-H-shaped bank
(SURVEY.md App. D), env seeds 1234 + i, auto-reset with seed + 4096 * reset_count.
One "step" = every environment takes DECISIONS_PER_STEP scheduling decisions (policy evaluation,
step(), observation construction, amortised resets included).

  value : decisions/s of the fused on-device rollout (state resident in HBM; one launch per step)
  e2e   : decisions/s through the host-buffer C ABI: per decision batch, actions are read back to
          pinned host memory, passed to ssb_step_host (H2D), and the B observation headers
          (reward/terminated/...) are copied D2H -- all inside the timed region
  roofline : algorithmic HBM bytes of the rollout kernel (SURVEY.md 8d formula, counters measured
          by the kernel itself) / its CUDA-event duration, against MEASURED_PEAKS.json
  cpu_baseline : the C oracle port (oracle/sim_oracle.c) on the box's host cores, same workload
  --impl reference : only that CPU arm, as its own JSON line
"""
from __future__ import annotations

import argparse
import json
import os
import os.path as osp
import subprocess
import sys
import threading
import time

import numpy as np

REPO = osp.dirname(osp.abspath(__file__))
for _p in (REPO, osp.join(REPO, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

ENVS_PER_GPU = 4096
DECISIONS_PER_STEP = 128
E2E_DECISIONS_PER_STEP = 32
ENV_CFG = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}  # examples.py:15-23
METRIC = "scheduling decisions/sec (batched envs)"
WORKLOAD = "C2: 4096 envs/GPU x (50 jobs, 10 executors), fair scheduler, synthetic TPC-H bank"


def peaks():
    path = osp.join(REPO, "MEASURED_PEAKS.json")
    if osp.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def algorithmic_bytes(st: dict) -> float:
    """B_dec = 80*eps + 20*N*sigma + (12 N + 8 M + 8 Ja + 16), summed over the
    observations / events / schedulability scans the kernel counted."""
    obs = max(st["observations"], 1)
    mean_nodes = st["sum_nodes"] / obs
    return (80.0 * st["events"] + 20.0 * mean_nodes * st["sched_scans"]
            + 12.0 * st["sum_nodes"] + 8.0 * st["sum_edges"] + 8.0 * st["sum_jobs"] + 16.0 * obs)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0])); mx.append(float(r[1]))
                for nm, v in zip(names, r[2:6]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def cpu_arm(seconds: float, threads: int | None = None) -> dict:
    """The CPU oracle port on host threads: fair C2 episodes back to back, seeds 1234 + i."""
    from concurrent.futures import ThreadPoolExecutor

    from oracle import OracleEnv
    from spark_sched_sim_b200.bank import synthetic_bank

    bank = synthetic_bank(0)
    P = threads or os.cpu_count() or 1
    envs = [OracleEnv(bank, ENV_CFG["num_executors"], ENV_CFG["job_arrival_cap"],
                      ENV_CFG["moving_delay"], ENV_CFG["warmup_delay"], ENV_CFG["job_arrival_rate"])
            for _ in range(P)]
    envs[0].run_fair_episode(1234)  # warm-up
    deadline = time.perf_counter() + seconds

    def work(rank):
        dec = ev = eps = 0
        k = 0
        while time.perf_counter() < deadline:
            d, e = envs[rank].run_fair_episode(1234 + rank + P * k)  # ctypes releases the GIL
            dec += d; ev += e; eps += 1; k += 1
        return dec, ev, eps

    t0 = time.perf_counter()
    with ThreadPoolExecutor(P) as ex:
        res = list(ex.map(work, range(P)))
    dt = time.perf_counter() - t0
    dec = sum(r[0] for r in res); ev = sum(r[1] for r in res); eps = sum(r[2] for r in res)
    return {"value": dec / dt, "unit": "decisions/s", "cores": P, "kind": "port",
            "sample": f"{eps} fair C2 episodes ({dec} decisions, {ev} events) in {dt:.1f} s on {P} "
                      f"host threads, oracle/sim_oracle.c incl. reset and observation building",
            "events_per_s": ev / dt}


def run_reference(args) -> None:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = 2.0
    vals = []
    for _ in range(args.warmup):
        cpu_arm(0.5)
    t0 = time.perf_counter()
    last = None
    for _ in range(args.steps):
        last = cpu_arm(per_step)
        vals.append(last["value"])
    dt = time.perf_counter() - t0
    v = float(np.mean(vals))
    last["value"] = v
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": v, "unit": "decisions/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64+int",
        "data": "synthetic", "config": {"workload": WORKLOAD, "step": f"{per_step} s of CPU episodes"},
        "cpu_baseline": last,
        "e2e": {"value": v, "unit": "decisions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference is pure Python and cannot travel to the GPU box; this arm times the C "
                "port of it (oracle/), which is ~250x faster per core than the Python original "
                "(BASELINE.md: 160-210 decisions/s/core)",
    }))


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--envs", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-threads", type=int, default=1,
                    help="host threads driving the e2e loop, each with its own shard of the envs")
    ap.add_argument("--e2e-budget", type=int, default=256,
                    help="max timeline events per env per ssb_step_host call (0 = run to next decision)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist

    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv
    from spark_sched_sim_b200.bank import synthetic_bank

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps
    B = args.envs
    D = DECISIONS_PER_STEP

    bank = synthetic_bank(0)
    env = BatchedSparkSchedSimEnv(ENV_CFG, num_envs=B, bank=bank, device=dev)
    from spark_sched_sim_b200 import parallel

    seeds, seed_step = parallel.shard_seeds(1234, B, rank, world)  # envs shard over ranks: disjoint seeds
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2
    stats_vec = torch.zeros(8, dtype=torch.float64, device=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def allreduce_stats():
        # the path's only exchange step: collect_stats (rollout_worker.py:122-129) of this rank's envs as sums,
        # reduced on the device and summed over ranks
        env.collect_stats(stats_vec)
        if world > 1:
            dist.all_reduce(stats_vec)

    # ---------------- device-resident rollout ("value")
    env.reset_host(seeds)
    for _ in range(W):
        env.rollout_fair(D, True, True, seed_step)
    env.reset_stats()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ev0 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    ev1 = [torch.cuda.Event(enable_timing=True) for _ in range(K)]
    barrier()
    t_wall0 = time.perf_counter()
    launches = 0
    for k in range(K):
        flush.fill_(k & 0xFF)  # evict L2 between timed iterations (untimed)
        ev0[k].record()
        env.rollout_fair(D, True, True, seed_step)
        launches += 3  # k_rollout_fair + the two ssb_collect_stats kernels
        allreduce_stats()
        ev1[k].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if rank == 0 else None
    kern_ms = sum(a.elapsed_time(b) for a, b in zip(ev0, ev1))
    st = env.stats()
    hdr = env.hdr()
    n_err = int((hdr["error"] != 0).sum())
    t = torch.tensor([kern_ms], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(st["decisions"]), float(st["events"]), float(st["episodes"]), float(n_err)],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt)
    max_ms = float(t.item())
    total_dec, total_ev, total_eps, total_err = (float(x) for x in cnt.tolist())
    value = total_dec / (max_ms * 1e-3)

    # ---------------- roofline of the rollout kernel (this rank's launches)
    peak, peak_src = peaks()
    alg_bytes = algorithmic_bytes(st)
    achieved = alg_bytes / (kern_ms * 1e-3) / 1e9
    traffic = None
    tpath = osp.join(REPO, "profiles", "rollout_traffic.json")
    if osp.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f).get("dram_bytes_per_launch")
    obs = max(st["observations"], 1)
    roofline = {
        "bound": "hbm", "kernel": "k_rollout_fair", "achieved": achieved, "peak": peak, "unit": "GB/s",
        "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
        "algorithmic_bytes_per_launch": alg_bytes / max(launches, 1),
        "bytes_per_decision": alg_bytes / max(st["decisions"], 1),
        "events_per_decision": st["events"] / max(st["decisions"], 1),
        "sched_scans_per_decision": st["sched_scans"] / max(st["decisions"], 1),
        "mean_nodes": st["sum_nodes"] / obs, "mean_edges": st["sum_edges"] / obs,
        "mean_active_jobs": st["sum_jobs"] / obs,
        "note": "latency/dependency-bound event chains: the HBM fraction is expected to be small",
    }

    # ---------------- e2e through the host-buffer C ABI
    # The caller-facing loop of a vector env.  --e2e-threads T > 1 splits this GPU's envs over T handles driven by
    # T host threads (one shard's host round trip overlaps the other shards' kernels); measured on B200 it does
    # not help (7.9 / 7.7 / 7.1 / 6.4 M decisions/s at T = 1 / 2 / 3 / 4): the loop is bound by the GPU, not the host.
    e2e = None
    if not args.no_e2e:
        from concurrent.futures import ThreadPoolExecutor

        De = E2E_DECISIONS_PER_STEP
        T = max(1, args.e2e_threads)
        shards = np.array_split(np.arange(B), T)
        envs = [BatchedSparkSchedSimEnv(ENV_CFG, num_envs=len(ix), bank=bank, device=dev) for ix in shards]
        pins = [(torch.empty(len(ix), dtype=torch.int32).pin_memory(),
                 torch.empty(len(ix), dtype=torch.int32).pin_memory()) for ix in shards]

        nexts = [(torch.empty(len(ix), dtype=torch.int32).pin_memory(),
                  torch.empty(len(ix), dtype=torch.int32).pin_memory()) for ix in shards]

        def e2e_steps(t, n_steps):
            torch.cuda.set_device(dev)
            e, (a_pin, n_pin), (a_nxt, n_nxt) = envs[t], pins[t], nexts[t]
            a_h, n_h, a_o, n_o = a_pin.numpy(), n_pin.numpy(), a_nxt.numpy(), n_nxt.numpy()
            for _ in range(n_steps * De):
                # one call per decision batch: H2D of the host's actions, bounded step (envs still simulating come
                # back "pending" and continue next call; finished envs re-seed themselves, ssb_set_autoreset), the
                # fair scheduler's action for the new observation evaluated inside the step kernel, D2H of the B
                # observation headers and of those actions, one synchronisation
                e.step_fair_host(a_h, n_h, a_o, n_o, True, max_events=args.e2e_budget)
                a_h, n_h, a_o, n_o = a_o, n_o, a_h, n_h   # the host feeds the actions it received back in

        for t in range(T):
            envs[t].reset_host(seeds[shards[t]])
            envs[t].set_autoreset(True, seed_step)
            a0, n0 = envs[t].fair_actions(True)   # the first actions; afterwards the step call returns them
            pins[t][0].copy_(a0); pins[t][1].copy_(n0)
            torch.cuda.synchronize()
        Ke = max(2, min(K, 5))
        with ThreadPoolExecutor(T) as ex:
            list(ex.map(lambda t: e2e_steps(t, 2), range(T)))
            for e in envs:
                e.reset_stats()
            barrier()
            t0 = time.perf_counter()
            list(ex.map(lambda t: e2e_steps(t, Ke), range(T)))
            barrier()
            dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dd = torch.tensor([float(sum(e.stats()["decisions"] for e in envs))], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dd)
        e2e = {"value": float(dd.item()) / float(tt.item()), "unit": "decisions/s",
               "h2d_bytes_per_step": De * 2 * 4 * B, "d2h_bytes_per_step": De * (2 * 4 + 48) * B,
               "steps": Ke, "calls_per_step": De, "max_events_per_call": args.e2e_budget,
               "host_threads": T,
               "path": "ssb_step_fair_host per decision batch: H2D of the host's actions -> bounded step with auto-reset "
                       "of finished envs + the fair scheduler's next action (in the step kernel) -> D2H of the "
                       "observation headers and the next actions (pinned) -> host feeds them back",
               "gpu_launches": Ke * De * T}
        launches_e2e = Ke * De * T
        # the rollout-collection call (rollout_worker.py:135-157 as ONE call): fused rollout on the device,
        # every transition (wall time, action, reward, flags) copied to pinned host memory per step
        from spark_sched_sim_b200 import _native as nat
        nb = B * D * nat.TRANSITION_DTYPE.itemsize
        traj_dev = torch.empty(nb, dtype=torch.uint8, device=dev)
        traj_pin = torch.empty(nb, dtype=torch.uint8).pin_memory()
        env.reset_host(seeds)
        for _ in range(2):
            env.rollout_fair_traj(D, True, True, seed_step, out=traj_dev, host=traj_pin)
        env.reset_stats()
        barrier()
        t0 = time.perf_counter()
        for _ in range(Ke):
            tr = env.rollout_fair_traj(D, True, True, seed_step, out=traj_dev, host=traj_pin)
        barrier()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        dd = torch.tensor([float(env.stats()["decisions"])], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(dd)
        e2e_rollout = {"value": float(dd.item()) / float(tt.item()), "unit": "decisions/s",
                       "h2d_bytes_per_step": 0, "d2h_bytes_per_step": nb, "steps": Ke,
                       "path": "ssb_rollout_fair_traj (policy + step fused on the device) -> D2H of the "
                               "B x 128 transition records to pinned memory",
                       "mean_reward_check": float(tr["reward"].mean())}
        launches_e2e += Ke
    else:
        launches_e2e = 0
        e2e_rollout = None

    cpu = None
    if rank == 0 and world == 1 and args.cpu_seconds > 0:
        cpu = cpu_arm(args.cpu_seconds)

    if rank == 0:
        out = {
            "metric": METRIC, "value": value, "unit": "decisions/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64 event times + integer state", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": B, "decisions_per_env_per_step": D,
                       "policy": "fair (on-device, fused)", "auto_reset": True,
                       "l2": "256 MiB flush write between timed iterations; workspace "
                             f"{env.workspace_bytes >> 20} MiB per GPU",
                       "parallelism": f"envs sharded over {world} GPU(s); all-reduce of rollout stats only"},
            "events_per_s": total_ev / (max_ms * 1e-3), "episodes": total_eps, "env_errors": total_err,
            "rollout_stats": parallel.stats_from_sums(stats_vec),
            "wall_s_timed_region": t_wall,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "e2e_rollout": e2e_rollout, "clocks": clocks,
            "gpu_launches": launches + launches_e2e,
        }
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
