"""Pins oracle/decima_policy.py (numpy float32 restatement of the Decima GNN + policy heads) to the
scores the reference's DecimaScheduler (shipped model.pt) produced for every decision of the recorded
Decima-driven episodes.  Tolerance: 2e-5 absolute on scores of magnitude ~10 (float32 matmuls with a
different summation order than torch's)."""
import os.path as osp

import numpy as np
import pytest

import decima_policy
from helpers import GOLDEN_DIR, golden_names, load_golden

TOL = 2e-5


def iter_decisions(tr):
    n = e = d = s = 0
    so = eo = 0
    for k in range(len(tr["actions"])):
        N, M, Ja = int(tr["N"][k]), int(tr["M"][k]), int(tr["Ja"][k])
        ns, ne = int(tr["pol_stage_count"][k]), int(tr["pol_exec_count"][k])
        yield k, {
            "features": tr["dec_feat"][n:n + N], "edge_links": tr["edges"][e:e + M],
            "edge_bits": tr["dec_edge_bits"][e:e + M], "depth": int(tr["dec_depth"][k]),
            "dag_ptr": tr["dag_ptr"][d:d + Ja + 1], "stage_mask": tr["dec_stage_mask"][n:n + N].astype(bool),
            "caps": tr["dec_caps"][s:s + Ja],
            "stage_logits": tr["pol_stage_logits"][so:so + ns], "exec_logits": tr["pol_exec_logits"][eo:eo + ne],
            "action": tuple(int(x) for x in tr["pol_actions"][k]), "lgprob": float(tr["pol_lgprob"][k]),
        }
        n += N; e += M; d += Ja + 1; s += Ja; so += ns; eo += ne


@pytest.mark.parametrize("name", [n for n in golden_names() if n.startswith("decima_")])
def test_policy_oracle_matches_reference_scores(name):
    tr = load_golden(name)
    w = decima_policy.load_weights(osp.join(GOLDEN_DIR, "decima_model.npz"))
    assert sum(v.size for v in w.values()) == 20802
    E = tr["num_executors"]
    worst = 0.0
    for k, o in iter_decisions(tr):
        h, h_dag, h_glob = decima_policy.encode(w, o["features"], o["edge_links"], o["edge_bits"],
                                                o["depth"], o["dag_ptr"])
        ss, jobs = decima_policy.stage_scores(w, o["features"], h, h_dag, h_glob, o["dag_ptr"], o["stage_mask"])
        assert ss.shape == o["stage_logits"].shape, k
        worst = max(worst, float(np.abs(ss - o["stage_logits"]).max()))
        stage_idx, job_idx, num_exec = o["action"]
        assert jobs[stage_idx] == job_idx  # scheduler.py:87-88
        cap = int(o["caps"][job_idx])
        es = decima_policy.exec_scores(w, o["features"], h_dag, h_glob, o["dag_ptr"], job_idx, cap, E)
        assert es.shape == o["exec_logits"].shape, k
        worst = max(worst, float(np.abs(es - o["exec_logits"]).max()))
        lg = decima_policy.log_softmax_at(ss, stage_idx) + decima_policy.log_softmax_at(es, num_exec)
        assert abs(lg - o["lgprob"]) < 1e-4, (k, lg, o["lgprob"])
    assert worst < TOL, worst
