"""GPU parity of the Decima observation adapter (ssb_decima_obs) against what the reference's own
DecimaObsWrapper produced on every observation of the recorded episodes (tests/golden dec_* arrays)
and against the numpy oracle: features bit-exact in float32, commit caps / exec_mask, stage mask,
per-level edge masks, message-passing depth."""
import numpy as np
import pytest

import decima_obs as oracle_dec
from helpers import bank_for, golden_names, load_golden

pytestmark = pytest.mark.gpu


def env_cfg_of(tr):
    return {"num_executors": tr["num_executors"],
            "job_arrival_cap": tr["job_arrival_cap"] if tr["job_arrival_cap"] > 0 else None,
            "job_arrival_rate": tr["job_arrival_rate"], "moving_delay": tr["moving_delay"],
            "warmup_delay": tr["warmup_delay"], "beta": tr["beta"]}


@pytest.mark.parametrize("name", golden_names(slim=False))
def test_decima_obs_matches_reference_wrapper(name):
    from spark_sched_sim_b200.batched_env import BatchedSparkSchedSimEnv

    tr = load_golden(name)
    bank = bank_for(tr)
    E = tr["num_executors"]
    B, slot = 2, 1
    env = BatchedSparkSchedSimEnv(env_cfg_of(tr), num_envs=B, bank=bank,
                                  max_jobs=len(tr["job_template"]) + 2,
                                  tape_capacity=len(tr["tape"]) + 8, decima_obs=True)
    for b in range(B):
        env.load_trace(b, tr["job_t_arrival"], tr["job_template"], tr["tape"])
    hdr = env.reset_host(np.full(B, tr["seed"], np.uint64)).copy()
    n = e = s = 0
    for k in range(len(tr["N"])):
        N, M, Ja = int(tr["N"][k]), int(tr["M"][k]), int(tr["Ja"][k])
        env.decima_obs()
        d = env.decima_obs_host(slot, hdr)
        assert np.array_equal(d["features"], tr["dec_feat"][n:n + N]), (k, "features")
        assert np.array_equal(d["stage_mask"], tr["dec_stage_mask"][n:n + N].astype(bool)), (k, "stage_mask")
        # frontier (heuristics/utils.py:5-14): nodes of the observed graph without an incoming edge
        want = np.ones(N, bool)
        want[env.obs(slot, hdr)["edge_links"][:, 1]] = False
        assert np.array_equal(d["frontier_mask"], want), (k, "frontier_mask")
        assert np.array_equal(d["commit_caps"], tr["dec_caps"][s:s + Ja]), (k, "caps")
        assert d["depth"] == tr["dec_depth"][k], (k, "depth")
        assert np.array_equal(d["edge_bits"], tr["dec_edge_bits"][e:e + M]), (k, "edge masks")
        assert d["exec_mask"].shape == (Ja, E) and d["edge_masks"].shape == (d["depth"], M)
        if k % 25 == 0:  # and the oracle agrees on the observation the GPU itself produced
            o = oracle_dec.decima_observation(env.obs(slot, hdr), E)
            assert np.array_equal(o["features"], d["features"]) and np.array_equal(o["edge_bits"], d["edge_bits"])
        n += N; e += M; s += Ja
        if k < len(tr["actions"]):
            a, c = tr["actions"][k]
            hdr = env.step_host(np.full(B, a, np.int32), np.full(B, c, np.int32)).copy()
            assert hdr[slot]["error"] == 0


def test_decima_wrappers_on_facade(bank):
    """DecimaEnvWrapper over the gym-style facade: observation keys/shapes of the reference wrapper
    and the +1 shift of num_exec (env_wrapper.py:33-34)."""
    from spark_sched_sim_b200.decima import DecimaEnvWrapper
    from spark_sched_sim_b200.env import SparkSchedSimEnv

    cfg = {"num_executors": 10, "job_arrival_cap": 6, "job_arrival_rate": 4.0e-5,
           "moving_delay": 2000.0, "warmup_delay": 1000.0}
    with pytest.raises(ValueError):
        DecimaEnvWrapper(SparkSchedSimEnv(cfg, bank=bank, decima_obs=False))  # built without the Decima buffers
    env = DecimaEnvWrapper(SparkSchedSimEnv(cfg, bank=bank))  # default: any env can be wrapped (examples.py:84-88)
    obs, _ = env.reset(seed=3)
    steps = 0
    done = False
    while not done and steps < 400:
        N = obs["dag_batch"].nodes.shape[0]
        Ja = len(obs["dag_ptr"]) - 1
        assert obs["dag_batch"].nodes.shape == (N, 5) and obs["dag_batch"].nodes.dtype == np.float32
        assert obs["stage_mask"].shape == (N,) and obs["exec_mask"].shape == (Ja, 10)
        assert obs["edge_masks"].shape[1] == obs["dag_batch"].edge_links.shape[0]
        sched = np.flatnonzero(obs["stage_mask"])
        assert sched.size > 0
        # first schedulable stage, as many executors as its job's cap allows
        node = int(sched[0])
        job = int(np.searchsorted(np.asarray(obs["dag_ptr"]), node, side="right") - 1)
        cap = int(obs["exec_mask"][job].sum())
        assert cap >= 1
        obs, reward, done, trunc, info = env.step({"stage_idx": 0, "job_idx": job, "num_exec": cap - 1})
        steps += 1
    assert done and env.unwrapped.all_jobs_complete
