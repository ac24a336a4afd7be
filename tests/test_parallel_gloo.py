"""N > 1 host logic on CPU: world_size-2 gloo processes shard the environments, run their share on
the CPU oracle (stand-in for the per-GPU env), and all-reduce the rollout statistics -- the same
helpers bench.py uses with NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from spark_sched_sim_b200 import parallel


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, envs_per_rank, out):
    import sys

    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), os.path.join(os.path.dirname(here), "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    from oracle import OracleEnv
    from spark_sched_sim_b200.bank import synthetic_bank

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    seeds, step = parallel.shard_seeds(1234, envs_per_rank, rank, world)
    bank = synthetic_bank(0)
    env = OracleEnv(bank, 10, 4, 2000.0, 1000.0, 4.0e-5)
    stats = {k: 0 for k in parallel.STAT_KEYS}
    for s in seeds:
        dec, ev = env.run_fair_episode(int(s), True)
        stats["decisions"] += dec
        stats["events"] += ev
        stats["episodes"] += 1
    total = parallel.allreduce_stats(stats)
    gathered = [None] * world
    dist.all_gather_object(gathered, (seeds.tolist(), step, stats))
    if rank == 0:
        out.put((total, gathered))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce():
    world, envs_per_rank = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, envs_per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    seeds = [s for g in gathered for s in g[0]]
    assert sorted(seeds) == list(range(1234, 1234 + world * envs_per_rank))  # disjoint, contiguous
    assert all(g[1] == world * envs_per_rank for g in gathered)
    assert total["decisions"] == sum(g[2]["decisions"] for g in gathered) > 0
    assert total["events"] == sum(g[2]["events"] for g in gathered)
    assert total["episodes"] == world * envs_per_rank


def _grad_worker(rank, world, port, out):
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.Generator().manual_seed(7)
    per_sample = torch.randn(10, 33, generator=g)  # every rank draws the same table, owns a slice of it
    lo, hi = parallel.shard_range(10, rank, world)
    mine = per_sample[lo:hi].sum(0)
    avg, total = parallel.allreduce_gradients(mine.clone(), hi - lo)
    if rank == 0:
        out.put((avg, total, per_sample.mean(0)))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce():
    """Each rank contributes the summed gradient of its own samples and their count; every rank ends up with the
    gradient of the mean over ALL samples."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_grad_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    avg, total, want = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert total == 10
    assert torch.allclose(avg, want, rtol=1e-6, atol=1e-7)


def test_shard_range_and_reference_seeds():
    for total in (1, 7, 4096, 65536):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    seeds, step = parallel.reference_worker_seeds(42, num_sequences=4, num_rollouts=4)
    assert step == 4 and seeds.tolist() == [42] * 4 + [43] * 4 + [44] * 4 + [45] * 4
    s = parallel.rollout_summary(10.0, 5.0, 2, 3, 4000.0)
    assert s["avg_num_jobs"] == 2.0 and s["avg_job_duration"] == 2.0
    assert parallel.allreduce_stats({k: 1 for k in parallel.STAT_KEYS})["events"] == 1.0


class _StubEnv:
    """Duck-typed stand-in for the batched env: what ppo_minibatch_update calls, on CPU tensors."""

    num_envs = 4

    def __init__(self):
        self.loaded = 0
        self.weights_set = 0

    def decima_snapshot_load(self, snapshot):
        self.loaded += 1

    def decima_snapshot_unload(self):
        self.loaded -= 1

    def decima_evaluate(self, snapshot, stage_sel, exec_sel, for_backward=False):
        return torch.zeros(4), torch.zeros(4)

    def decima_backward(self, g_lp, g_en, grads):
        grads += 1.0

    def set_decima_weights(self, params):
        self.weights_set += 1


class _StubAdam:
    def __init__(self):
        self.params = torch.zeros(5)
        self.steps = 0

    def step(self, grads):
        self.steps += 1
        self.last = grads.clone()


def _kl_worker(rank, world, port, kls, out):
    from spark_sched_sim_b200 import ppo

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    results = []
    for kl_pair in kls:
        env, adam = _StubEnv(), _StubAdam()
        loss_fn = lambda lg, old, en, ret, base, kl=kl_pair[rank]: (  # noqa: E731
            torch.tensor([0.5, 0.4, 0.1, kl]), torch.zeros(4), torch.zeros(4))
        info, stepped = ppo.ppo_minibatch_update(env, None, None, None, None, None, None, loss_fn, adam,
                                                 target_kl=0.01, allreduce=parallel.allreduce_gradients)
        results.append((stepped, info["approx_kl_div"], info["approx_kl_div_local"], adam.steps, env.loaded))
    gathered = [None] * world
    dist.all_gather_object(gathered, results)
    if rank == 0:
        out.put(gathered)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_kl_early_stop_is_collective():
    """The ranks' local approximate KLs straddle 1.5 * target_kl: the decision is taken on their sample-weighted
    mean, both ranks take the same branch (no rank is left waiting in the gradient all-reduce), and the stored
    observation is unloaded on both branches."""
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    # (rank 0, rank 1): mean 0.0125 < 0.015 -> both step; mean 0.0175 > 0.015 -> both stop; both low; both high
    kls = [(0.020, 0.005), (0.030, 0.005), (0.001, 0.002), (0.05, 0.06)]
    procs = [ctx.Process(target=_kl_worker, args=(r, world, port, kls, q)) for r in range(world)]
    for p in procs:
        p.start()
    gathered = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [True, False, True, False]
    for r in range(world):
        for i, (stepped, kl, kl_local, steps, loaded) in enumerate(gathered[r]):
            assert stepped == want[i] and steps == int(want[i]) and loaded == 0
            assert abs(kl - sum(kls[i]) / 2) < 1e-7 and abs(kl_local - kls[i][r]) < 1e-7
