"""Data-parallel PPO mini-batches over NCCL: every rank owns its envs and stored observations, the flat gradient
vector (+ sample count) is all-reduced (parallel.allreduce_gradients), every rank steps Adam on the same gradient.
Checks that the ranks' weights stay bit-identical.  Round 1, 2 x B200: "weights identical across ranks: True" after 6
mini-batches of 1024 samples per rank (the printed time includes NCCL's communicator set-up in the first all-reduce).  Usage on a 2-GPU box:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 profiles/ppo_update_2gpu.py"""
import os
import os.path as osp
import sys
import time

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
sys.path.insert(0, REPO)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from spark_sched_sim_b200 import parallel  # noqa: E402
from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv  # noqa: E402
from spark_sched_sim_b200.ppo import Adam, PPOLoss, ppo_minibatch_update  # noqa: E402

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl")
cfg = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5, "moving_delay": 2000.0, "warmup_delay": 1000.0}
B = 1024
env = BatchedSparkSchedSimEnv(cfg, num_envs=B, device=f"cuda:{local}", decima_policy=True)
z = np.load(osp.join(REPO, "tests", "golden", "decima_model.npz"))
w = {k: z[k] for k in z.files}
env.set_decima_weights(w)
seeds, step = parallel.shard_seeds(1234, B, rank, world)
env.reset_host(seeds)
env.set_autoreset(True, step)
env.rollout_decima(60)
batches = []
for _ in range(3):
    snap = env.decima_snapshot()
    a, n = env.decima_policy()
    batches.append((snap, env.pol_action[:, 0].contiguous(), env.pol_action[:, 2].contiguous(), env.pol_lgprob.clone()))
    env.step(a, n)
flat = torch.from_numpy(np.concatenate([w[k].astype(np.float32).reshape(-1) for k in w])).cuda()
adam = Adam(flat, lr=3e-4, max_grad_norm=0.5)
g = torch.Generator(device="cuda").manual_seed(100 + rank)
ret = -1e4 * torch.rand(B, device="cuda", generator=g, dtype=torch.float64)
base = ret + 2e3 * torch.randn(B, device="cuda", generator=g, dtype=torch.float64)
torch.cuda.synchronize(); t0 = time.perf_counter()
for it in range(6):
    snap, ss, es, lg = batches[it % 3]
    info, stepped = ppo_minibatch_update(env, snap, ss, es, lg, ret, base, PPOLoss(0.2, 0.04), adam,
                                         allreduce=parallel.allreduce_gradients if world > 1 else None)
torch.cuda.synchronize(); dt = time.perf_counter() - t0
digest = torch.stack([adam.params.double().sum(), adam.params.double().abs().sum(), adam.grad_norm.double()[0]])
if world > 1:
    all_d = [torch.zeros_like(digest) for _ in range(world)]
    dist.all_gather(all_d, digest)
    same = all(torch.equal(all_d[0], d) for d in all_d)
else:
    same = True
if rank == 0:
    print(f"{world} GPU(s): 6 data-parallel PPO mini-batches of {B} samples per rank in {1e3 * dt:.1f} ms "
          f"({6 * B * world / dt / 1e3:.0f} k samples/s); weights identical across ranks: {same}; "
          f"grad norm {float(adam.grad_norm):.4f}, loss {info['loss']:.5f}")
assert same
if world > 1:
    dist.barrier(); dist.destroy_process_group()
