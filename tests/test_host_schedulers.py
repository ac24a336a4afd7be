"""Host-side scheduler plug-ins (same interface as schedulers/heuristics in the reference) against
the observations and actions recorded from the reference run (no GPU needed)."""
import numpy as np
import pytest

from helpers import golden_names, load_golden
from spark_sched_sim_b200.gym_compat import GraphInstance
from spark_sched_sim_b200.schedulers import RandomScheduler, RoundRobinScheduler, make_scheduler


def recorded_obs(tr):
    n = e = d = s = 0
    for k in range(len(tr["N"])):
        N, M, Ja = int(tr["N"][k]), int(tr["M"][k]), int(tr["Ja"][k])
        yield {
            "dag_batch": GraphInstance(tr["nodes"][n:n + N], np.zeros(M, int),
                                       tr["edges"][e:e + M].astype(np.int64)),
            "dag_ptr": tr["dag_ptr"][d:d + Ja + 1].tolist(),
            "num_committable_execs": int(tr["ncommit"][k]),
            "source_job_idx": int(tr["src"][k]),
            "exec_supplies": tr["supplies"][s:s + Ja].tolist(),
        }
        n += N; e += M; d += Ja + 1; s += Ja


@pytest.mark.parametrize("name", [n for n in golden_names(slim=False) if not n.startswith("decima_")])
def test_host_policies_reproduce_recorded_actions(name):
    tr = load_golden(name)
    if tr["policy"] in ("fair", "fifo"):
        sched = RoundRobinScheduler(tr["num_executors"], dynamic_partition=(tr["policy"] == "fair"))
    else:
        sched = RandomScheduler(seed=tr["policy_seed"])
    assert sched.env_wrapper_cls is None and sched.name in ("Fair", "FIFO", "Random")
    for k, obs in enumerate(recorded_obs(tr)):
        if k == len(tr["actions"]):
            break
        action, info = sched.schedule(obs)
        assert (int(action["stage_idx"]), int(action["num_exec"])) == tuple(tr["actions"][k]), k
        assert info == {}


def _reference_policy(tr):
    """The reference's own scheduler object for a recorded trace, or None outside the build container."""
    import os.path as osp
    import sys

    sys.path.insert(0, osp.join(osp.dirname(osp.dirname(osp.abspath(__file__))), "oracle"))
    import refrun

    if not refrun.reference_available():
        return None
    refrun.setup()
    return refrun.make_policy(tr["policy"], tr["num_executors"], tr["policy_seed"])


@pytest.mark.parametrize("name", [n for n in golden_names(slim=False) if not n.startswith("decima_")][:8])
def test_reference_schedulers_agree_on_the_same_observations(name):
    """The UNMODIFIED reference schedulers (imported from /root/reference, build container only) and this package's
    classes, side by side on every recorded observation: same action, and the same keys left in the observation
    dict.  Together with the GPU facade test (the facade's observations equal the recorded ones bit for bit) this is
    "the reference's schedulers run on the facade unchanged"."""
    import copy

    tr = load_golden(name)
    ref = _reference_policy(tr)
    if ref is None:
        pytest.skip("reference not present (GPU box)")
    if tr["policy"] in ("fair", "fifo"):
        ours = RoundRobinScheduler(tr["num_executors"], dynamic_partition=(tr["policy"] == "fair"))
    else:
        ours = RandomScheduler(seed=tr["policy_seed"])
    assert ours.name == ref.name
    for k, obs in enumerate(recorded_obs(tr)):
        if k == len(tr["actions"]):
            break
        o_ref, o_ours = copy.deepcopy(obs), copy.deepcopy(obs)
        a_ref, i_ref = ref.schedule(o_ref)
        a_ours, i_ours = ours.schedule(o_ours)
        assert {k_: int(v) for k_, v in a_ref.items()} == {k_: int(v) for k_, v in a_ours.items()}, k
        assert i_ref == i_ours == {}
        assert {int(x) for x in o_ref["frontier_stages"]} == o_ours["frontier_stages"]
        assert {int(a): int(b) for a, b in o_ref["schedulable_stages"].items()} == o_ours["schedulable_stages"]


def test_make_scheduler_factory():
    s = make_scheduler({"agent_cls": "RoundRobinScheduler", "num_executors": 10, "dynamic_partition": False})
    assert s.name == "FIFO"
    with pytest.raises(AssertionError):
        make_scheduler({"agent_cls": "Nope"})


def test_decima_scheduler_constructor_contract():
    """schedulers.DecimaScheduler keeps the reference's constructor keywords (schedulers/decima/scheduler.py:22-69)
    and refuses what the device policy does not implement -- no GPU needed for that."""
    import os.path as osp

    from helpers import GOLDEN_DIR
    from spark_sched_sim_b200.decima import DecimaEnvWrapper
    from spark_sched_sim_b200.schedulers import DecimaScheduler, make_scheduler

    agent_cfg = {"agent_cls": "DecimaScheduler", "embed_dim": 16, "gnn_mlp_kwargs": {"hid_dims": [32, 16], "act_cls": "LeakyReLU"},
                 "policy_mlp_kwargs": {"hid_dims": [64, 64], "act_cls": "Tanh"}, "num_executors": 10,
                 "state_dict_path": osp.join(GOLDEN_DIR, "decima_model.npz"), "training_mode": False}
    s = make_scheduler(agent_cfg)
    assert isinstance(s, DecimaScheduler) and s.env_wrapper_cls is DecimaEnvWrapper and s.name.startswith("Decima:")
    assert len(s._state_dict) == 42 and sum(v.size for v in s._state_dict.values()) == 20802
    with pytest.raises(ValueError):
        DecimaScheduler(10, embed_dim=32, state_dict_path=agent_cfg["state_dict_path"])
    with pytest.raises(ValueError):
        DecimaScheduler(10)  # no weights
    with pytest.raises(ValueError):
        s.schedule({"dag_ptr": [0]})  # not an observation of DecimaEnvWrapper


def test_stats_from_sums():
    from spark_sched_sim_b200 import parallel

    d = parallel.stats_from_sums([50.0, 2.0, 30.0, 44.0, 600000.0, 9.0, 0.0, 0.0])
    assert d == {"avg_num_jobs": 25.0, "num_completed_jobs": 15.0, "num_job_arrivals": 22.0, "avg_job_duration": 20.0}
