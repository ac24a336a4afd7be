"""CPU restatement of the Decima policy forward pass -- TEST INFRASTRUCTURE ONLY.

Follows schedulers/decima/scheduler.py:71-99 (schedule), :142-276 (encoders), :279-385 (policy
networks) and utils.py:45-64 (make_mlp) in float32 numpy; the literal sequential loop over the
per-level edge masks with the "overwrite" semantics of NodeEncoder.forward (:191-234, SURVEY.md
App. E).  Pinned against the logits the reference's DecimaScheduler (shipped models/decima/model.pt)
produced on every decision of the recorded episodes (tests/golden/decima_*.npz `pol_*` arrays).
Weights: tests/golden/decima_model.npz (exported from the reference's model.pt by gen_golden.py).
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


def _mlp(x, w, prefix, act):
    """make_mlp(in, [h1, h2], out): Linear/act/Linear/act/Linear (utils.py:45-64), in x's dtype."""
    dt = x.dtype.type
    for i, k in enumerate((0, 2, 4)):
        W, b = w[f"{prefix}.{k}.weight"].astype(dt), w[f"{prefix}.{k}.bias"].astype(dt)
        x = (x @ W.T + b).astype(dt)
        if i < 2:
            x = act(x)
    return x


def _leaky(x):
    return np.where(x > 0, x, x.dtype.type(0.2) * x).astype(x.dtype)


def _tanh(x):
    return np.tanh(x).astype(x.dtype)


def encode(w, x, edge_links, edge_bits, depth, dag_ptr):
    """-> (h_node [N,16], h_dag [Ja,16], h_glob [16]).  Computes in x's dtype: float32 = the reference's arithmetic;
    pass the features as float64 for the "exact" evaluation the accuracy tests measure both sides against."""
    F32 = x.dtype.type  # noqa: N806 (shadows the module constant on purpose: everything below follows x)
    N = x.shape[0]
    pre = "encoder.node_encoder"
    h_init = _mlp(x, w, f"{pre}.mlp_prep", _leaky)
    if depth == 0:
        h = h_init  # _forward_no_mp (:236-241)
    else:
        h = np.zeros_like(h_init)
        u, v = edge_links[:, 0], edge_links[:, 1]
        is_src = np.zeros(N, bool)
        is_src[u] = True  # nodes that are the tail of some edge
        sinks = ~is_src
        h[sinks] = _mlp(h_init[sinks], w, f"{pre}.mlp_update", _leaky)  # :208-212
        for k in reversed(range(depth)):  # reverse_flow: children -> parents, deepest level first
            m = ((edge_bits >> np.uint64(k)) & np.uint64(1)).astype(bool)
            uk, vk = u[m], v[m]
            senders = np.zeros(N, bool); senders[vk] = True
            receivers = np.zeros(N, bool); receivers[uk] = True
            msg = np.zeros_like(h)
            msg[senders] = _mlp(h[senders], w, f"{pre}.mlp_msg", _leaky)
            agg = np.zeros_like(h)
            np.add.at(agg, uk, msg[vk])  # adj @ msg, sequential in edge order
            h[receivers] = h_init[receivers] + _mlp(agg[receivers], w, f"{pre}.mlp_update", _leaky)
    z = _mlp(np.concatenate([x, h], 1), w, "encoder.dag_encoder.mlp", _leaky)
    Ja = len(dag_ptr) - 1
    h_dag = np.zeros((Ja, 16), F32)
    for j in range(Ja):
        for n in range(dag_ptr[j], dag_ptr[j + 1]):
            h_dag[j] += z[n]
    g = _mlp(h_dag, w, "encoder.global_encoder.mlp", _leaky)
    h_glob = np.zeros(16, F32)
    for j in range(Ja):
        h_glob += g[j]
    return h, h_dag, h_glob


def stage_scores(w, x, h, h_dag, h_glob, dag_ptr, stage_mask):
    """scores over the schedulable nodes, in node order (:293-320)."""
    F32 = x.dtype.type  # noqa: N806
    idx = np.flatnonzero(stage_mask)
    job = np.searchsorted(np.asarray(dag_ptr), idx, side="right") - 1
    inp = np.concatenate([x[idx], h[idx], h_dag[job], np.broadcast_to(h_glob, (len(idx), 16))], 1).astype(F32)
    return _mlp(inp, w, "stage_policy_network.mlp_score", _tanh)[:, 0], job


def exec_scores(w, x, h_dag, h_glob, dag_ptr, job, cap, num_executors):
    """scores over num_exec = 0 .. cap-1 for the chosen job (:338-385)."""
    F32 = x.dtype.type  # noqa: N806
    x_dag = x[dag_ptr[job], :3]
    # (torch.arange(E) / E is a float32 tensor in the reference: the input itself is rounded to float32)
    c = (np.arange(cap, dtype=np.int64) / num_executors).astype(np.float32).astype(F32)[:, None]
    inp = np.concatenate([np.broadcast_to(x_dag, (cap, 3)), np.broadcast_to(h_dag[job], (cap, 16)),
                          np.broadcast_to(h_glob, (cap, 16)), c], 1).astype(F32)
    return _mlp(inp, w, "exec_policy_network.mlp_score", _tanh)[:, 0]


def log_softmax_at(scores, idx):
    """lgprob as utils.sample computes it: log(softmax(scores)[idx]) (utils.py:19-23)."""
    s = scores.astype(np.float32)
    e = np.exp(s - s.max())
    return float(np.log((e / e.sum())[idx]))


def load_weights(path):
    z = np.load(path)
    return {k: z[k].astype(F32) for k in z.files}
