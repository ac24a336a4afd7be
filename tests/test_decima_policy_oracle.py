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


@pytest.mark.parametrize("name", [n for n in golden_names(slim=False) if n.startswith("decima_")])
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


@pytest.mark.parametrize("name", [n for n in golden_names(slim=True) if n.startswith("decima_")])
def test_oracles_at_the_configured_scale(name):
    """Decima-driven reference episodes at the headline shape (50 jobs x 10 executors) and at
    config/decima_tpch.yaml:80-87 (200 jobs x 50 executors).  The fixtures hold digests instead of observations, so
    the chain is: C oracle replays the recorded actions -> its observation (digest == the reference's) -> numpy
    restatement of DecimaObsWrapper (digest == the reference wrapper's output on EVERY observation) -> numpy policy
    on the decisions whose scores were kept == the reference DecimaScheduler's scores."""
    import decima_obs
    from helpers import bank_for, decima_digest, obs_digest
    from oracle import OracleEnv

    tr = load_golden(name)
    E = tr["num_executors"]
    w = decima_policy.load_weights(osp.join(GOLDEN_DIR, "decima_model.npz"))
    env = OracleEnv(bank_for(tr), E, tr["job_arrival_cap"], tr["moving_delay"], tr["warmup_delay"],
                    tr["job_arrival_rate"], tr["beta"])
    obs = env.reset_seed(tr["seed"], tr["time_limit"])
    kept = {int(k): i for i, k in enumerate(tr["pol_kept"])} if "pol_kept" in tr else None
    so = np.concatenate([[0], np.cumsum(tr["pol_stage_count"])])
    eo = np.concatenate([[0], np.cumsum(tr["pol_exec_count"])])
    if kept is not None:  # offsets into the kept subset
        ks = tr["pol_kept"]
        so_k = np.concatenate([[0], np.cumsum(tr["pol_stage_count"][ks])])
        eo_k = np.concatenate([[0], np.cumsum(tr["pol_exec_count"][ks])])
    worst, checked = 0.0, 0
    for k in range(len(tr["actions"]) + 1):
        assert obs_digest(obs) == int(tr["obs_digest"][k]), (k, "base observation")
        d = decima_obs.decima_observation(obs, E)
        assert decima_digest(d["features"], d["commit_caps"], d["depth"], d["edge_bits"],
                             d["stage_mask"].astype(np.uint8)) == int(tr["dec_digest"][k]), (k, "decima observation")
        assert d["depth"] == int(tr["dec_depth"][k])
        if k == len(tr["actions"]):
            break
        if kept is None or k in kept:
            if kept is None:
                s0, s1, e0, e1 = so[k], so[k + 1], eo[k], eo[k + 1]
            else:
                i = kept[k]
                s0, s1, e0, e1 = so_k[i], so_k[i + 1], eo_k[i], eo_k[i + 1]
            h, h_dag, h_glob = decima_policy.encode(w, d["features"], obs["edge_links"], d["edge_bits"], d["depth"],
                                                    obs["dag_ptr"])
            ss, jobs = decima_policy.stage_scores(w, d["features"], h, h_dag, h_glob, obs["dag_ptr"], d["stage_mask"])
            stage_idx, job_idx, num_exec = (int(x) for x in tr["pol_actions"][k])
            assert ss.shape[0] == s1 - s0 and jobs[stage_idx] == job_idx, k
            es = decima_policy.exec_scores(w, d["features"], h_dag, h_glob, obs["dag_ptr"], job_idx,
                                           int(d["commit_caps"][job_idx]), E)
            assert es.shape[0] == e1 - e0, k
            worst = max(worst, float(np.abs(ss - tr["pol_stage_logits"][s0:s1]).max()),
                        float(np.abs(es - tr["pol_exec_logits"][e0:e1]).max()))
            lg = decima_policy.log_softmax_at(ss, stage_idx) + decima_policy.log_softmax_at(es, num_exec)
            assert abs(lg - float(tr["pol_lgprob"][k])) < 1e-4, (k, lg)
            checked += 1
        a, n = tr["actions"][k]
        rc, r, term = env.step(int(a), int(n))
        assert rc == 0 and r == tr["reward"][k] and env.wall_time == tr["wall"][k], k
        obs = env.obs()
    assert term and checked >= 100 and worst < TOL, (checked, worst)
