"""Pins the CPU oracle (oracle/sim_oracle.c) to traces of the unmodified reference (tests/golden):
event order, executor assignments, observations, rewards and f64 job completion times, bit-exact.
Rewards with beta > 0 go through exp() and are compared at 1e-12 relative."""
import numpy as np
import pytest

from helpers import bank_for, golden_names, load_golden, replay_and_compare
from oracle import OracleEnv


def make_oracle(bank, tr):
    return OracleEnv(bank, tr["num_executors"], tr["job_arrival_cap"], tr["moving_delay"],
                     tr["warmup_delay"], tr["job_arrival_rate"], tr["beta"], log=True)


@pytest.mark.parametrize("name", golden_names())
def test_oracle_replays_reference_tape(name):
    tr = load_golden(name)
    env = make_oracle(bank_for(tr), tr)  # (checks the bank's checksum against the fixture's)
    replay_and_compare(env, tr, "tape", check_policy=env.fair_action,
                       reward_rtol=1e-12 if tr["beta"] > 0 else 0.0)


@pytest.mark.parametrize("name", [n for n in golden_names() if "philox" in n])
def test_oracle_replays_reference_from_seed(name):
    """Philox-plugged reference runs: the oracle samples jobs and durations itself (tpch.py logic)."""
    tr = load_golden(name)
    env = make_oracle(bank_for(tr), tr)
    replay_and_compare(env, tr, "seed", check_policy=env.fair_action,
                       reward_rtol=1e-12 if tr["beta"] > 0 else 0.0)


def test_bank_matches_reference_sampler():
    """num_tasks / rough_task_duration per stage as the reference computed them (tpch.py:162-196)."""
    for name in golden_names():
        tr = load_golden(name)
        bank = bank_for(tr)
        ts = np.concatenate([np.arange(bank.stage_base[t], bank.stage_base[t + 1])
                             for t in tr["job_template"]])
        assert np.array_equal(bank.num_tasks[ts], tr["st_num_tasks"])
        assert np.array_equal(bank.rough_duration[ts], tr["st_rough"])


def test_oracle_invalid_actions(bank):
    tr = load_golden("e10_j8_fair_s2_philox")
    env = make_oracle(bank, tr)
    obs = env.reset_seed(tr["seed"])
    N = obs["nodes"].shape[0]
    assert env.step(N, 1)[0] == 1       # outside the action space
    env = make_oracle(bank, tr); env.reset_seed(tr["seed"])
    assert env.step(0, 0)[0] == 1       # num_exec outside Discrete(E, start=1)
    env = make_oracle(bank, tr); env.reset_seed(tr["seed"])
    assert env.step(0, 11)[0] == 1
    env = make_oracle(bank, tr); obs = env.reset_seed(tr["seed"])
    nsched = int(obs["nodes"][:, 2].sum())
    if nsched < N:
        assert env.step(nsched, 1)[0] == 2  # KeyError on stage_selection_map
