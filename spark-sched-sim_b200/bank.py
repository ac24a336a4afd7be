"""Template bank: TPC-H-shaped job templates flattened for the device.

The reference samples a job by picking one of 22 queries x 7 input sizes and loading two pickled
files per job (`spark_sched_sim/data_samplers/tpch.py:118-132,176-206`).  Here every template is
loaded ONCE, pre-processed exactly as the reference does per job
(`_pre_process_task_duration` `tpch.py:135-159`, `_rough_task_duration` `:162-174`, `num_tasks`
`:185-187`) and flattened into a handful of arrays ("CSR bank") that live in HBM for the whole run:

    template t = size_idx * 22 + (query_num - 1)
    stage_base[t] .. stage_base[t+1]      -> template-stage rows ("ts")
    edge_base[t]  .. edge_base[t+1]       -> edges (u, v), row-major order of the adjacency matrix
                                             (= `nx.from_numpy_array(...).edges`, `tpch.py:199`)
    per ts: num_tasks, rough_duration (f64), parent_mask / child_mask (u64 bitmasks over the
            job's stages), and a 3 x 8 table (wave x executor level) of (offset, count) into
            `dur_values` (f64, ms), plus per-wave "level key present" bitmasks.

Waves: 0 = fresh_durations, 1 = first_wave (cleaned), 2 = rest_wave.  Executor levels are the
reference's `exec_levels` (`tpch.py:238`).

The real dataset cannot be downloaded here (no network), so `make_synthetic_tpch` regenerates the
synthetic workload of SURVEY.md App. D verbatim; `load_tpch_dir` reads the reference's on-disk
layout when a `data/tpch` directory is supplied; `write_tpch_dir` writes that layout (used to feed
the unmodified reference when golden traces are generated).
"""
from __future__ import annotations

import hashlib
import os
import os.path as osp
from dataclasses import dataclass, field

import numpy as np

QUERY_SIZES = ["2g", "5g", "10g", "20g", "50g", "80g", "100g"]  # tpch.py:14
NUM_QUERIES = 22  # tpch.py:15
EXEC_LEVELS = [5, 10, 20, 40, 50, 60, 80, 100]  # tpch.py:238
WAVES = ["fresh_durations", "first_wave", "rest_wave"]
WAVE_FRESH, WAVE_FIRST, WAVE_REST = 0, 1, 2
MAX_STAGES_PER_JOB = 64  # parent/child/active/frontier sets are u64 bitmasks on the device


def make_synthetic_tpch(seed: int = 0, kind: str = "appd") -> dict:
    """SURVEY.md App. D generator, verbatim: {(size, q): (adj int[n,n], td dict)}.
    kind "wide": the same generator with 30..64 stages per template and few tasks per stage -- a test workload
    for the code paths that only jobs of more than 32 stages reach (the u64 halves of the stage bitmasks)."""
    assert kind in ("appd", "wide")
    rng = np.random.default_rng(seed if kind == "appd" else (seed, 64))
    wide = kind == "wide"
    data = {}
    for si, size in enumerate(QUERY_SIZES):
        for q in range(1, NUM_QUERIES + 1):
            n = int(rng.integers(30, 65)) if wide else int(rng.integers(2, 19))
            adj = np.zeros((n, n), dtype=int)
            for v in range(1, n):
                k = int(rng.integers(1, min(3, v) + 1))
                for u in rng.choice(v, size=k, replace=False):
                    adj[u, v] = 1
            td = {}
            for s in range(n):
                ntasks = int(rng.integers(1, (6 if wide else 60) * (1 + si)))
                base = float(rng.uniform(200, 4000)) * (1 + 0.3 * si)
                d = {"fresh_durations": {}, "first_wave": {}, "rest_wave": {}}
                for e in EXEC_LEVELS:
                    nf = min(e, ntasks)
                    d["first_wave"][e] = [
                        int(x) for x in rng.normal(base * 1.5, base * 0.2, nf).clip(10)
                    ]
                    d["rest_wave"][e] = [
                        int(x)
                        for x in rng.normal(base, base * 0.2, ntasks - nf).clip(10)
                    ]
                    d["fresh_durations"][e] = [
                        int(x)
                        for x in rng.normal(base * 2.0, base * 0.3, max(1, nf // 2)).clip(10)
                    ]
                td[s] = d
            data[(size, q)] = (adj, td)
    return data


def write_tpch_dir(root: str, data: dict) -> None:
    """Writes `data` under `<root>/data/tpch/<size>/{adj_mat,task_duration}_<q>.npy`
    (the layout `tpch.py:118-127` reads, relative to the CWD)."""
    for (size, q), (adj, td) in data.items():
        d = osp.join(root, "data", "tpch", size)
        os.makedirs(d, exist_ok=True)
        np.save(osp.join(d, f"adj_mat_{q}.npy"), adj)
        np.save(osp.join(d, f"task_duration_{q}.npy"), td, allow_pickle=True)


def load_tpch_dir(root: str) -> dict:
    """Reads the reference's on-disk layout from `<root>/data/tpch` (or `<root>` itself)."""
    base = osp.join(root, "data", "tpch")
    if not osp.isdir(base):
        base = root
    data = {}
    for size in QUERY_SIZES:
        for q in range(1, NUM_QUERIES + 1):
            adj = np.load(osp.join(base, size, f"adj_mat_{q}.npy"), allow_pickle=True)
            td = np.load(
                osp.join(base, size, f"task_duration_{q}.npy"), allow_pickle=True
            ).item()
            assert adj.shape[0] == adj.shape[1] == len(td)  # tpch.py:129-130
            data[(size, q)] = (adj, td)
    return data


def _clean_first_wave(td_stage: dict) -> dict:
    """`_pre_process_task_duration` (tpch.py:135-159): drop from each first-wave list the values
    that also occur in the fresh list of the same level (multiset semantics), then give empty
    levels the list of the nearest lower level."""
    clean: dict = {}
    for e in td_stage["first_wave"]:
        clean[e] = []
        fresh: dict = {}
        for d in td_stage["fresh_durations"][e]:
            fresh[d] = fresh.get(d, 0) + 1
        for d in td_stage["first_wave"][e]:
            if d not in fresh:
                clean[e].append(d)
            else:
                fresh[d] -= 1
                if fresh[d] == 0:
                    del fresh[d]
    last: list = []
    for e in sorted(clean.keys()):
        if len(clean[e]) == 0:
            clean[e] = last
        last = clean[e]
    return clean


@dataclass
class TemplateBank:
    """Flattened, read-only workload description shared by every environment."""

    num_stages: np.ndarray  # i32[T]
    stage_base: np.ndarray  # i32[T+1]
    edge_base: np.ndarray  # i32[T+1]
    edges: np.ndarray  # i32[sum M, 2]   (u, v) local stage ids
    num_tasks: np.ndarray  # i32[TS]
    rough_duration: np.ndarray  # f64[TS]
    parent_mask: np.ndarray  # u64[TS]
    child_mask: np.ndarray  # u64[TS]
    present: np.ndarray  # u8[TS, 3]  bit l set <=> level EXEC_LEVELS[l] is a key of that wave
    dur_off: np.ndarray  # u32[TS, 3, 8]
    dur_cnt: np.ndarray  # u32[TS, 3, 8]
    dur_values: np.ndarray  # f64[sum]
    meta: dict = field(default_factory=dict)

    @property
    def num_templates(self) -> int:
        return int(self.num_stages.shape[0])

    @property
    def max_stages(self) -> int:
        return int(self.num_stages.max())

    @property
    def max_edges(self) -> int:
        return int(np.diff(self.edge_base).max())

    def checksum(self) -> str:
        h = hashlib.sha256()
        for name in (
            "num_stages", "edges", "num_tasks", "rough_duration", "parent_mask",
            "present", "dur_off", "dur_cnt", "dur_values",
        ):
            h.update(np.ascontiguousarray(getattr(self, name)).tobytes())
        return h.hexdigest()[:16]


def build_bank(data: dict) -> TemplateBank:
    """Flattens `{(size, q): (adj, td)}` (all 7 x 22 templates required) into a TemplateBank."""
    T = len(QUERY_SIZES) * NUM_QUERIES
    num_stages = np.zeros(T, np.int32)
    stage_base = np.zeros(T + 1, np.int32)
    edge_base = np.zeros(T + 1, np.int32)
    edges, num_tasks, rough, pmask, cmask = [], [], [], [], []
    present, dur_off, dur_cnt, values = [], [], [], []
    nvals = 0
    for si, size in enumerate(QUERY_SIZES):
        for q in range(1, NUM_QUERIES + 1):
            t = si * NUM_QUERIES + (q - 1)
            adj, td = data[(size, q)]
            adj = np.asarray(adj)
            n = adj.shape[0]
            if n > MAX_STAGES_PER_JOB:
                raise ValueError(f"template {size}/{q}: {n} stages > {MAX_STAGES_PER_JOB}")
            num_stages[t] = n
            stage_base[t + 1] = stage_base[t] + n
            us, vs = np.nonzero(adj)  # row-major == nx.from_numpy_array edge order
            if len(us) == 0:
                # the reference crashes on edge-less DAGs (np.vstack([]) at spark_sched_sim.py:254)
                raise ValueError(f"template {size}/{q} has no edges")
            edge_base[t + 1] = edge_base[t] + len(us)
            edges.append(np.stack([us, vs], 1).astype(np.int32))
            for s in range(n):
                d = td[s]
                e0 = next(iter(d["first_wave"]))
                num_tasks.append(len(d["first_wave"][e0]) + len(d["rest_wave"][e0]))
                waves = {
                    "fresh_durations": d["fresh_durations"],
                    "first_wave": _clean_first_wave(d),
                    "rest_wave": d["rest_wave"],
                }
                allv = []
                for w in WAVES:  # tpch.py:168-172 order
                    for ts_ in waves[w].values():
                        allv += list(ts_)
                rough.append(np.mean(allv))
                pm = 0
                for u in np.nonzero(adj[:, s])[0]:
                    pm |= 1 << int(u)
                cm = 0
                for v in np.nonzero(adj[s, :])[0]:
                    cm |= 1 << int(v)
                pmask.append(pm)
                cmask.append(cm)
                pres = [0, 0, 0]
                off = np.zeros((3, 8), np.uint32)
                cnt = np.zeros((3, 8), np.uint32)
                for wi, w in enumerate(WAVES):
                    for key, lst in waves[w].items():
                        if key not in EXEC_LEVELS:
                            raise ValueError(f"unexpected executor level {key!r}")
                        li = EXEC_LEVELS.index(key)
                        pres[wi] |= 1 << li
                        off[wi, li] = nvals
                        cnt[wi, li] = len(lst)
                        values.append(np.asarray(lst, dtype=np.float64))
                        nvals += len(lst)
                present.append(pres)
                dur_off.append(off)
                dur_cnt.append(cnt)
    return TemplateBank(
        num_stages=num_stages,
        stage_base=stage_base,
        edge_base=edge_base,
        edges=np.concatenate(edges, 0),
        num_tasks=np.asarray(num_tasks, np.int32),
        rough_duration=np.asarray(rough, np.float64),
        parent_mask=np.asarray(pmask, np.uint64),
        child_mask=np.asarray(cmask, np.uint64),
        present=np.asarray(present, np.uint8),
        dur_off=np.stack(dur_off, 0),
        dur_cnt=np.stack(dur_cnt, 0),
        dur_values=np.concatenate(values) if values else np.zeros(0, np.float64),
    )


_SYNTH_CACHE: dict = {}


def synthetic_bank(seed: int = 0, kind: str = "appd") -> TemplateBank:
    """The App. D workload (or its "wide" test variant, see make_synthetic_tpch) as a TemplateBank (cached)."""
    key = (seed, kind)
    if key not in _SYNTH_CACHE:
        bank = build_bank(make_synthetic_tpch(seed, kind))
        bank.meta = {"source": f"synthetic({kind}, default_rng({seed}))"}
        _SYNTH_CACHE[key] = bank
    return _SYNTH_CACHE[key]


def executor_intervals(exec_cap: int) -> np.ndarray:
    """`_init_executor_intervals` (tpch.py:237-262): (exec_cap+1, 2) table of the two data levels
    bracketing each local-executor count."""
    lv = EXEC_LEVELS
    iv = np.zeros((exec_cap + 1, 2))
    iv[: lv[0] + 1] = lv[0]
    for i in range(len(lv) - 1):
        iv[lv[i] + 1 : lv[i + 1]] = (lv[i], lv[i + 1])
        if lv[i + 1] > exec_cap:
            break
        iv[lv[i + 1]] = lv[i + 1]
    if exec_cap > lv[-1]:
        iv[lv[-1] + 1 : exec_cap] = lv[-1]
    return iv
