"""CPU restatement of Decima's observation adapter -- TEST INFRASTRUCTURE ONLY.

Follows `DecimaObsWrapper.observation/_build_node_features` (schedulers/decima/env_wrapper.py:69-143)
and `make_dag_layer_edge_masks` (schedulers/decima/utils.py:238-267) in plain numpy, without
networkx.  Pinned against the reference's own wrapper through the `dec_*` arrays of tests/golden
(recorded by oracle/refrun.py with the unmodified wrapper class).

Edge masks are returned transposed and bit-packed: `edge_bits[e]` has bit k set iff edge e belongs to
`edge_masks[k]` of the reference (k = 0 .. depth-1, depth = number of topological generations - 1).
"""
from __future__ import annotations

import numpy as np

NUM_TASKS_SCALE = 200  # env_wrapper.py:41
WORK_SCALE = 1e5


def topological_generations(num_nodes: int, edge_links: np.ndarray) -> list[list[int]]:
    """networkx.topological_generations: level 0 = nodes of in-degree 0, then peel."""
    indeg = np.zeros(num_nodes, np.int64)
    succ: list[list[int]] = [[] for _ in range(num_nodes)]
    for u, v in edge_links:
        indeg[v] += 1
        succ[u].append(int(v))
    level = [n for n in range(num_nodes) if indeg[n] == 0]
    out = []
    while level:
        out.append(level)
        nxt = []
        for n in level:
            for c in succ[n]:
                indeg[c] -= 1
                if indeg[c] == 0:
                    nxt.append(c)
        level = nxt
    return out


def layer_edge_bits(edge_links: np.ndarray, num_nodes: int) -> tuple[np.ndarray, int]:
    """(bits u64[M], depth): mask k = edges with both ends in level_k U successors(level_k)
    (utils.py:259-265), k over all levels but the last; depth 0 when there is at most one level."""
    M = edge_links.shape[0]
    levels = topological_generations(num_nodes, edge_links)
    bits = np.zeros(M, np.uint64)
    if len(levels) <= 1:
        return bits, 0
    succ: list[set[int]] = [set() for _ in range(num_nodes)]
    for u, v in edge_links:
        succ[int(u)].add(int(v))
    for k, level in enumerate(levels[:-1]):
        mask = np.zeros(num_nodes, bool)
        mask[level] = True
        for n in level:
            for c in succ[n]:
                mask[c] = True
        em = mask[edge_links[:, 0]] & mask[edge_links[:, 1]]
        bits |= em.astype(np.uint64) << np.uint64(k)
    return bits, len(levels) - 1


def decima_observation(obs: dict, num_executors: int) -> dict:
    """obs: {"nodes" f32[N,3], "edge_links" int[M,2], "dag_ptr", "exec_supplies",
    "num_committable_execs", "source_job_idx"} -> Decima features / masks."""
    nodes = np.asarray(obs["nodes"], np.float32)
    ptr = np.asarray(obs["dag_ptr"])
    counts = ptr[1:] - ptr[:-1]
    supplies = np.asarray(obs["exec_supplies"])
    ncommit = int(obs["num_committable_execs"])
    j_src = int(obs["source_job_idx"])
    caps = np.minimum(np.maximum(num_executors - supplies, 0), ncommit)  # :74-77
    if j_src < supplies.size:
        caps[j_src] = ncommit  # :81-82
    N = nodes.shape[0]
    feat = np.zeros((N, 5), np.float32)
    feat[:, 0] = np.repeat(caps, counts) / num_executors  # f64 expression stored as f32 (:124)
    feat[:, 1] = -1
    if j_src < supplies.size:
        feat[ptr[j_src]:ptr[j_src + 1], 1] = 1
    feat[:, 2] = np.repeat(supplies, counts) / num_executors
    feat[:, 3] = nodes[:, 0] / NUM_TASKS_SCALE  # float32 arithmetic (:137)
    feat[:, 4] = nodes[:, 0] * nodes[:, 1] / WORK_SCALE  # float32 arithmetic (:141)
    bits, depth = layer_edge_bits(np.asarray(obs["edge_links"]).reshape(-1, 2), N)
    return {"features": feat, "stage_mask": nodes[:, 2].astype(bool), "commit_caps": caps.astype(np.int32),
            "edge_bits": bits, "depth": depth}


def exec_mask_from_caps(caps: np.ndarray, num_executors: int) -> np.ndarray:
    """exec_mask bool[Ja, E]: the first `cap` entries of each row (env_wrapper.py:92-94)."""
    return np.arange(num_executors)[None, :] < np.asarray(caps)[:, None]


def edge_masks_from_bits(bits: np.ndarray, depth: int) -> np.ndarray:
    """bool[depth, M] as the reference returns it."""
    k = np.arange(depth, dtype=np.uint64)[:, None]
    return ((np.asarray(bits, np.uint64)[None, :] >> k) & np.uint64(1)).astype(bool)
