"""Runs the UNMODIFIED reference (`/root/reference`) and records traces -- TEST INFRASTRUCTURE ONLY.

Works only in the build container (the reference is not present on the GPU box).  The reference is
imported from where it lies, with `oracle/refshim` standing in for gymnasium, the synthetic TPC-H
bank written in the reference's own on-disk layout, and instance-level
instrumentation (no source edits) that records

* the job sequence the reference sampled (arrival time, template, per-stage `num_tasks` and
  `rough_task_duration` -- used to pin `bank.py`'s restatement of tpch.py:135-187),
* the duration tape: every value `TPCHDataSampler.task_duration` returned, in call order,
* every popped event `(t, type, job, stage, task, executor, t_accepted)` (spark_sched_sim.py:326-329),
* every action and everything `step()` returned, including the whole observation dict,
* final per-job completion times,
* every executor's `history` list (executor.py:25-44) as the episode left it.

`tests/golden/gen_golden.py` serialises these as fixtures; the C oracle (oracle/sim_oracle.c) is
pinned against them, and the CUDA path is compared with the oracle.
"""
from __future__ import annotations

import hashlib
import importlib
import os
import os.path as osp
import sys
import types

import numpy as np

HERE = osp.dirname(osp.abspath(__file__))
REPO = osp.dirname(HERE)
REFERENCE = os.environ.get("SSB_REFERENCE", "/root/reference")

EV_JOB_ARRIVAL, EV_TASK_FINISHED, EV_EXECUTOR_READY = 0, 1, 2  # event.py:11-14 order


def reference_available() -> bool:
    return osp.isdir(osp.join(REFERENCE, "spark_sched_sim"))


def _import_product_bank():
    sys.path.insert(0, REPO) if REPO not in sys.path else None
    import spark_sched_sim_b200.bank as bank  # data preparation only (no simulation code)

    return bank


_SETUP_DONE = {}


def setup(bank_seed: int = 0, bank_kind: str = "appd") -> str:
    """Puts the shim + reference on sys.path, writes the synthetic dataset, chdirs next to it."""
    key = (bank_seed, bank_kind)
    if key in _SETUP_DONE:
        os.chdir(_SETUP_DONE[key])
        return _SETUP_DONE[key]
    if not reference_available():
        raise RuntimeError(f"reference not found at {REFERENCE}")
    for p in (REFERENCE, osp.join(HERE, "refshim")):
        if p not in sys.path:
            sys.path.insert(0, p)
    # `schedulers/__init__.py` eagerly imports Decima (needs PyG); stub the package so that
    # `schedulers.heuristics` can be imported on its own.
    if "schedulers" not in sys.modules:
        pkg = types.ModuleType("schedulers")
        pkg.__path__ = [osp.join(REFERENCE, "schedulers")]
        sys.modules["schedulers"] = pkg
    # the Decima *observation wrapper* needs only numpy + networkx, but its package imports PyG at
    # module level; empty stand-ins are enough to import it (the GNN itself is not executed here)
    class _Stub(types.ModuleType):
        def __getattr__(self, item):  # pyg.data.Batch etc. in annotations evaluated at import time
            if item.startswith("__"):
                raise AttributeError(item)
            return _Stub(f"{self.__name__}.{item}")

        def __or__(self, other):
            return self

        __ror__ = __or__

    for name in ("torch_geometric", "torch_sparse", "torch_scatter"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = _Stub(name)
    bank = _import_product_bank()
    root = f"/tmp/ssb_refdata_{bank_kind}_seed{bank_seed}"
    if not osp.isdir(osp.join(root, "data", "tpch", "100g")):
        bank.write_tpch_dir(root, bank.make_synthetic_tpch(bank_seed, bank_kind))
    os.chdir(root)
    _SETUP_DONE[key] = root
    return root


def make_policy(name: str, num_executors: int, seed: int = 42):
    rr = importlib.import_module("schedulers.heuristics.round_robin")
    rnd = importlib.import_module("schedulers.heuristics.random_scheduler")
    if name == "fair":
        return rr.RoundRobinScheduler(num_executors, dynamic_partition=True)
    if name == "fifo":
        return rr.RoundRobinScheduler(num_executors, dynamic_partition=False)
    if name == "random":
        return rnd.RandomScheduler(seed=seed)
    if name == "decima":
        # the reference's DecimaScheduler with the shipped weights and the agent section of
        # config/decima_tpch.yaml:66-77 (examples.py:64-72); PyG comes from oracle/refshim
        import random

        import torch

        random.seed(seed)  # utils.sample draws with python's `random` (decima/utils.py:19-23)
        torch.manual_seed(seed)
        dec = importlib.import_module("schedulers.decima.scheduler")
        sched = dec.DecimaScheduler(
            num_executors=num_executors, embed_dim=16,
            gnn_mlp_kwargs={"hid_dims": [32, 16], "act_cls": "LeakyReLU",
                            "act_kwargs": {"inplace": True, "negative_slope": 0.2}},
            policy_mlp_kwargs={"hid_dims": [64, 64], "act_cls": "Tanh"},
            state_dict_path=osp.join(REFERENCE, "models", "decima", "model.pt"))
        sched.eval()
        return sched
    raise ValueError(name)


def obs_digest(nodes, edge_links, dag_ptr, supplies, ncommit, src) -> int:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(nodes, dtype=np.float32).tobytes())
    h.update(np.ascontiguousarray(edge_links, dtype=np.int32).tobytes())
    h.update(np.asarray(dag_ptr, dtype=np.int32).tobytes())
    h.update(np.asarray(supplies, dtype=np.int32).tobytes())
    h.update(np.asarray([ncommit, src], dtype=np.int32).tobytes())
    return int.from_bytes(h.digest()[:8], "little")


def events_digest(ev: dict, lo: int, hi: int) -> int:
    h = hashlib.sha256()
    for k in ("ev_t", "ev_type", "ev_job", "ev_stage", "ev_task", "ev_exec", "ev_tacc"):
        h.update(np.ascontiguousarray(ev[k][lo:hi]).tobytes())
    return int.from_bytes(h.digest()[:8], "little")


def run_episode(
    env_cfg: dict,
    policy: str = "fair",
    seed: int = 1234,
    rng: str = "pcg64",
    policy_seed: int = 42,
    time_limit: float | None = None,
    max_steps: int | None = None,
    bank_seed: int = 0,
    decima: bool = False,
    bank_kind: str = "appd",
) -> dict:
    """One reference episode -> trace dict of numpy arrays (see module docstring)."""
    setup(bank_seed, bank_kind)
    import gymnasium
    from philox_ref import PhiloxNpRandom

    gymnasium.NP_RANDOM_FACTORY = PhiloxNpRandom if rng == "philox" else None
    env_cfg = dict(env_cfg)
    env_cfg.setdefault("data_sampler_cls", "TPCHDataSampler")
    env = gymnasium.make("spark_sched_sim:SparkSchedSimEnv-v0", env_cfg=env_cfg)
    sched = make_policy(policy, env_cfg["num_executors"], policy_seed)
    dec_wrapper = None
    if decima:  # the reference's own DecimaObsWrapper, evaluated on every base observation
        dec_wrapper = importlib.import_module("schedulers.decima.env_wrapper").DecimaObsWrapper(env)
    from spark_sched_sim.components.event import Event

    type_map = {
        Event.Type.JOB_ARRIVAL: EV_JOB_ARRIVAL,
        Event.Type.TASK_FINISHED: EV_TASK_FINISHED,
        Event.Type.EXECUTOR_READY: EV_EXECUTOR_READY,
    }

    tape, tape_meta, events = [], [], []
    orig_td = env.data_sampler.task_duration

    def task_duration(job, stage, task, executor):
        d = orig_td(job, stage, task, executor)
        tape.append(float(d))
        tape_meta.append((job.id_, stage.id_, task.id_, executor.id_))
        return d

    env.data_sampler.task_duration = task_duration
    orig_he = env._handle_event

    def handle_event(event):
        d = event.data
        if event.type == Event.Type.JOB_ARRIVAL:
            row = (env.wall_time, EV_JOB_ARRIVAL, d["job"].id_, -1, -1, -1, np.inf)
        elif event.type == Event.Type.TASK_FINISHED:
            t = d["task"]
            row = (env.wall_time, EV_TASK_FINISHED, d["stage"].job_id, d["stage"].id_,
                   t.id_, t.executor_id, float(t.t_accepted))
        else:
            row = (env.wall_time, EV_EXECUTOR_READY, d["stage"].job_id, d["stage"].id_,
                   -1, d["executor"].id_, np.inf)
        events.append(row)
        return orig_he(event)

    env._handle_event = handle_event

    # A policy with an env wrapper (Decima) steps the reference's own wrapper stack; the base
    # observation underneath is captured at `_observe` so every trace has the same layout.
    last_base = {}
    orig_observe = env._observe

    def observe():
        o = orig_observe()
        last_base["obs"] = o
        return o

    env._observe = observe
    wenv = sched.env_wrapper_cls(env) if getattr(sched, "env_wrapper_cls", None) else env
    pol_rec = {"stage_logits": [], "exec_logits": [], "dec_actions": [], "lgprob": []}
    if policy == "decima":
        dutils = importlib.import_module("schedulers.decima.utils")
        orig_sample = dutils.sample
        calls = []

        def sample(logits):  # records the scores the policy networks produced for this decision
            calls.append(np.asarray(logits.detach().cpu().numpy(), np.float32).copy())
            return orig_sample(logits)

        dutils.sample = sample

    options = {"time_limit": time_limit} if time_limit is not None else None
    obs, info = wenv.reset(seed=seed, options=options)

    jobs = list(env.jobs.values())
    sizes = ["2g", "5g", "10g", "20g", "50g", "80g", "100g"]
    job_template = [sizes.index(str(j.query_size)) * 22 + int(j.query_num) - 1 for j in jobs]
    st_num_tasks = [s.num_tasks for j in jobs for s in j.stages]
    st_rough = [float(s.most_recent_duration) for j in jobs for s in j.stages]

    rec = {k: [] for k in (
        "actions", "reward", "wall", "term", "ncommit", "src", "N", "M", "Ja", "ev_count",
        "nodes", "edges", "dag_ptr", "supplies", "obs_digest",
        "dec_feat", "dec_caps", "dec_depth", "dec_edge_bits", "dec_stage_mask")}

    def record_base(o):
        g = o["dag_batch"]
        nodes = np.asarray(g.nodes, np.float32).reshape(-1, 3)
        el = np.asarray(g.edge_links, np.int64).reshape(-1, 2)
        rec["N"].append(nodes.shape[0])
        rec["M"].append(el.shape[0])
        rec["Ja"].append(len(o["exec_supplies"]))
        rec["ncommit"].append(int(o["num_committable_execs"]))
        rec["src"].append(int(o["source_job_idx"]))
        rec["nodes"].append(nodes.copy())
        rec["edges"].append(el.astype(np.int32))
        rec["dag_ptr"].append(np.asarray(o["dag_ptr"], np.int32))
        rec["supplies"].append(np.asarray(o["exec_supplies"], np.int32))
        rec["obs_digest"].append(
            obs_digest(nodes, el, o["dag_ptr"], o["exec_supplies"],
                       o["num_committable_execs"], o["source_job_idx"]))

    def record_decima(o):
        d = dec_wrapper.observation(o)
        feat = np.asarray(d["dag_batch"].nodes, np.float32).reshape(-1, 5)
        em = np.asarray(d["edge_masks"], bool)  # (depth - 1, M)
        bits = np.zeros(em.shape[1], np.uint64)
        for k in range(em.shape[0]):
            bits |= em[k].astype(np.uint64) << np.uint64(k)
        rec["dec_feat"].append(feat.copy())
        rec["dec_caps"].append(np.asarray(d["exec_mask"], bool).sum(1).astype(np.int32))
        rec["dec_depth"].append(em.shape[0])
        rec["dec_edge_bits"].append(bits)
        rec["dec_stage_mask"].append(np.asarray(d["stage_mask"], np.uint8))

    def record_obs():
        o = last_base["obs"]
        record_base(o)
        if dec_wrapper is not None:
            record_decima(o)

    record_obs()  # observation 0 = reset
    rec["ev_count"].append(len(events))
    terminated = truncated = False
    steps = 0
    while not (terminated or truncated):
        action, pinfo = sched.schedule(obs)
        if policy == "decima":
            # env action as DecimaActWrapper forms it (env_wrapper.py:33-34)
            a = (int(action["stage_idx"]), 1 + int(action["num_exec"]))
            pol_rec["stage_logits"].append(calls[-2])
            pol_rec["exec_logits"].append(calls[-1])
            pol_rec["dec_actions"].append((int(action["stage_idx"]), int(action["job_idx"]),
                                           int(action["num_exec"])))
            pol_rec["lgprob"].append(float(pinfo["lgprob"]))
            obs, reward, terminated, truncated, info = wenv.step(action)
        else:
            a = (int(action["stage_idx"]), int(action["num_exec"]))
            obs, reward, terminated, truncated, info = wenv.step(
                {"stage_idx": a[0], "num_exec": a[1]})
        rec["actions"].append(a)
        rec["reward"].append(float(reward))
        rec["wall"].append(float(info["wall_time"]))
        rec["term"].append(bool(terminated))
        record_obs()
        rec["ev_count"].append(len(events))
        steps += 1
        if time_limit is not None and info["wall_time"] >= time_limit:
            truncated = True  # what StochasticTimeLimit.step does (stochastic_time_limit.py:29-30)
        if max_steps is not None and steps >= max_steps:
            break
    if policy == "decima":
        dutils.sample = orig_sample

    evarr = np.array(events, dtype=np.float64).reshape(-1, 7)
    # executor.history (executor.py:25-44) as the reference left it: per executor the list [[t, job_id], ...] whose
    # last entry has t = None; stored as its add_history calls (t_k = history[k-1][0], job_k = history[k][1])
    hist_ptr, hist_t, hist_job = [0], [], []
    for ex in env.executors:
        h = ex.history
        assert h[0][1] == -1 and h[-1][0] is None
        for k in range(1, len(h)):
            hist_t.append(float(h[k - 1][0]))
            hist_job.append(int(h[k][1]))
        hist_ptr.append(len(hist_t))
    trace = {
        "num_executors": env_cfg["num_executors"],
        "moving_delay": float(env_cfg["moving_delay"]),
        "warmup_delay": float(env_cfg["warmup_delay"]),
        "job_arrival_rate": float(env_cfg["job_arrival_rate"]),
        "job_arrival_cap": -1 if env_cfg.get("job_arrival_cap") is None else int(env_cfg["job_arrival_cap"]),
        "beta": float(env_cfg.get("beta", 0.0)),
        "time_limit": np.inf if time_limit is None else float(time_limit),
        "seed": seed,
        "rng": rng,
        "policy": policy,
        "policy_seed": policy_seed,
        "bank_seed": bank_seed,
        "bank_kind": bank_kind,
        "job_t_arrival": np.array([float(j.t_arrival) for j in jobs], np.float64),
        "job_template": np.array(job_template, np.int32),
        "job_t_completed": np.array([float(j.t_completed) for j in jobs], np.float64),
        "st_num_tasks": np.array(st_num_tasks, np.int32),
        "st_rough": np.array(st_rough, np.float64),
        "tape": np.array(tape, np.float64),
        "tape_meta": np.array(tape_meta, np.int32).reshape(-1, 4),
        "actions": np.array(rec["actions"], np.int32).reshape(-1, 2),
        "reward": np.array(rec["reward"], np.float64),
        "wall": np.array(rec["wall"], np.float64),
        "term": np.array(rec["term"], np.uint8),
        "ncommit": np.array(rec["ncommit"], np.int32),
        "src": np.array(rec["src"], np.int32),
        "N": np.array(rec["N"], np.int32),
        "M": np.array(rec["M"], np.int32),
        "Ja": np.array(rec["Ja"], np.int32),
        "ev_count": np.array(rec["ev_count"], np.int64),
        "obs_digest": np.array(rec["obs_digest"], np.uint64),
        "nodes": np.concatenate(rec["nodes"], 0) if rec["nodes"] else np.zeros((0, 3), np.float32),
        "edges": np.concatenate(rec["edges"], 0) if rec["edges"] else np.zeros((0, 2), np.int32),
        "dag_ptr": np.concatenate(rec["dag_ptr"]),
        "supplies": np.concatenate(rec["supplies"]) if rec["supplies"] else np.zeros(0, np.int32),
        "dec_feat": np.concatenate(rec["dec_feat"], 0) if rec["dec_feat"] else np.zeros((0, 5), np.float32),
        "dec_caps": np.concatenate(rec["dec_caps"]) if rec["dec_caps"] else np.zeros(0, np.int32),
        "dec_depth": np.array(rec["dec_depth"], np.int32),
        "dec_edge_bits": np.concatenate(rec["dec_edge_bits"]) if rec["dec_edge_bits"] else np.zeros(0, np.uint64),
        "dec_stage_mask": np.concatenate(rec["dec_stage_mask"]) if rec["dec_stage_mask"] else np.zeros(0, np.uint8),
        "pol_stage_logits": (np.concatenate(pol_rec["stage_logits"]) if pol_rec["stage_logits"]
                             else np.zeros(0, np.float32)),
        "pol_stage_count": np.array([len(x) for x in pol_rec["stage_logits"]], np.int32),
        "pol_exec_logits": (np.concatenate(pol_rec["exec_logits"]) if pol_rec["exec_logits"]
                            else np.zeros(0, np.float32)),
        "pol_exec_count": np.array([len(x) for x in pol_rec["exec_logits"]], np.int32),
        "pol_actions": np.array(pol_rec["dec_actions"], np.int32).reshape(-1, 3),
        "pol_lgprob": np.array(pol_rec["lgprob"], np.float64),
        "ev_t": evarr[:, 0].copy(),
        "ev_type": evarr[:, 1].astype(np.uint8),
        "ev_job": evarr[:, 2].astype(np.int16),
        "ev_stage": evarr[:, 3].astype(np.int16),
        "ev_task": evarr[:, 4].astype(np.int32),
        "ev_exec": evarr[:, 5].astype(np.int16),
        "ev_tacc": evarr[:, 6].copy(),
        "hist_ptr": np.array(hist_ptr, np.int32),
        "hist_t": np.array(hist_t, np.float64),
        "hist_job": np.array(hist_job, np.int32),
        "final_wall": float(env.wall_time),
        "avg_job_duration_s": float(
            np.mean([min(j.t_completed, env.wall_time) - j.t_arrival for j in jobs]) * 1e-3),
    }
    return trace


def decima_digest(feat, caps, depth, edge_bits, stage_mask) -> int:
    """sha256 over what DecimaObsWrapper adds to one observation (features, exec_mask as caps, edge masks as per-edge
    level bits, stage mask)."""
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(feat, dtype=np.float32).tobytes())
    h.update(np.asarray(caps, dtype=np.int32).tobytes())
    h.update(np.asarray([depth], dtype=np.int32).tobytes())
    h.update(np.ascontiguousarray(edge_bits, dtype=np.uint64).tobytes())
    h.update(np.ascontiguousarray(stage_mask, dtype=np.uint8).tobytes())
    return int.from_bytes(h.digest()[:8], "little")


def slim(trace: dict, logit_stride: int = 1) -> dict:
    """Replaces the bulky per-step observations and per-event rows by digests (big episodes).  Decima-driven
    episodes keep the policy's recorded scores of every `logit_stride`-th decision (`pol_kept` = their indices) and
    the log-probability of every decision."""
    t = dict(trace)
    if len(t["dec_depth"]):  # per-observation digest of the Decima adapter's outputs
        n = e = s = 0
        dig = []
        for k in range(len(t["N"])):
            N, M, Ja = int(t["N"][k]), int(t["M"][k]), int(t["Ja"][k])
            dig.append(decima_digest(t["dec_feat"][n:n + N], t["dec_caps"][s:s + Ja], int(t["dec_depth"][k]),
                                     t["dec_edge_bits"][e:e + M], t["dec_stage_mask"][n:n + N]))
            n += N; e += M; s += Ja
        t["dec_digest"] = np.array(dig, np.uint64)
    if len(t["pol_stage_count"]) and logit_stride > 1:
        keep = np.arange(0, len(t["pol_stage_count"]), logit_stride)
        so = np.concatenate([[0], np.cumsum(t["pol_stage_count"])])
        eo = np.concatenate([[0], np.cumsum(t["pol_exec_count"])])
        t["pol_stage_logits"] = np.concatenate([t["pol_stage_logits"][so[k]:so[k + 1]] for k in keep])
        t["pol_exec_logits"] = np.concatenate([t["pol_exec_logits"][eo[k]:eo[k + 1]] for k in keep])
        t["pol_kept"] = keep.astype(np.int32)
    ec = t["ev_count"]
    lo = np.concatenate([[0], ec[:-1]])
    t["ev_digest"] = np.array(
        [events_digest(trace, int(a), int(b)) for a, b in zip(lo, ec)], np.uint64)
    for k in ("nodes", "edges", "dag_ptr", "supplies", "ev_t", "ev_type", "ev_job", "ev_stage",
              "ev_task", "ev_exec", "ev_tacc", "tape_meta", "dec_feat", "dec_caps",
              "dec_edge_bits", "dec_stage_mask"):
        t.pop(k)
    t["slim"] = True
    return t


if __name__ == "__main__":
    import time

    cfg = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    t0 = time.time()
    tr = run_episode(cfg, "fair", 1234, rng=sys.argv[1] if len(sys.argv) > 1 else "pcg64")
    dt = time.time() - t0
    print(f"steps={len(tr['actions'])} events={len(tr['ev_t'])} tasks={len(tr['tape'])} "
          f"final_wall={tr['final_wall']} avg_jct_s={tr['avg_job_duration_s']:.3f} "
          f"stages={len(tr['st_num_tasks'])} time={dt:.2f}s")
