"""Pins oracle/learner.py (discounted returns, interpolated group baseline) to the outputs of the reference's
own ReturnsCalculator / Baseline classes recorded in tests/golden/learner_vectors.npz."""
import os.path as osp

import numpy as np
import pytest

from helpers import GOLDEN_DIR


def load_cases():
    z = np.load(osp.join(GOLDEN_DIR, "learner_vectors.npz"))
    cases = {}
    for name in sorted({k.rsplit("_", 1)[0] for k in z.files if not k.startswith("diff_")}):
        num_seq, num_roll, beta = z[name + "_meta"]
        lens = z[name + "_len"].astype(int)
        t_off = np.concatenate([[0], np.cumsum(lens + 1)])
        r_off = np.concatenate([[0], np.cumsum(lens)])
        cases[name] = dict(
            num_rollouts=int(num_roll), beta=float(beta), lens=lens,
            times=[z[name + "_times"][t_off[i]:t_off[i + 1]] for i in range(len(lens))],
            rewards=[z[name + "_rewards"][r_off[i]:r_off[i + 1]] for i in range(len(lens))],
            returns=[z[name + "_returns"][r_off[i]:r_off[i + 1]] for i in range(len(lens))],
            baselines=[z[name + "_baselines"][r_off[i]:r_off[i + 1]] for i in range(len(lens))])
    return cases


def load_diff_cases():
    """Differential returns: per case three consecutive calls of ONE reference ReturnsCalculator(buff_cap)."""
    z = np.load(osp.join(GOLDEN_DIR, "learner_vectors.npz"))
    cases = {}
    for name in ("diff_small", "diff_large"):
        calls = []
        for c in range(3):
            key = f"{name}_c{c}"
            cap, avg = z[key + "_meta"]
            lens = z[key + "_len"].astype(int)
            t_off = np.concatenate([[0], np.cumsum(lens + 1)])
            r_off = np.concatenate([[0], np.cumsum(lens)])
            calls.append(dict(
                cap=int(cap), avg_num_jobs=float(avg), lens=lens,
                times=[z[key + "_times"][t_off[i]:t_off[i + 1]] for i in range(len(lens))],
                rewards=[z[key + "_rewards"][r_off[i]:r_off[i + 1]] for i in range(len(lens))],
                returns=[z[key + "_returns"][r_off[i]:r_off[i + 1]] for i in range(len(lens))]))
        cases[name] = calls
    return cases


CASES = load_cases()
DIFF_CASES = load_diff_cases()


@pytest.mark.parametrize("name", sorted(CASES))
def test_returns_and_baselines_match_reference(name):
    from learner import discounted_returns, group_baselines  # oracle/learner.py (oracle/ is on sys.path)

    c = CASES[name]
    rets = [discounted_returns(r, t, c["beta"]) for r, t in zip(c["rewards"], c["times"])]
    for got, want in zip(rets, c["returns"]):
        # math.exp vs numpy's exp: last-bit differences, accumulated over the recurrence
        assert np.allclose(got, want, rtol=1e-12, atol=0.0)
    # the baseline from the REFERENCE's returns must be reproduced bit for bit (interp + pairwise mean)
    base = group_baselines([t[:-1] for t in c["times"]], c["returns"], c["num_rollouts"])
    for got, want in zip(base, c["baselines"]):
        assert np.array_equal(got, want)


@pytest.mark.parametrize("name", sorted(DIFF_CASES))
def test_differential_returns_match_reference(name):
    """The moving average of the number of jobs and the differential returns, bit for bit over three calls."""
    from learner import DifferentialReturns

    calc = DifferentialReturns(DIFF_CASES[name][0]["cap"])
    for c in DIFF_CASES[name]:
        rets = calc(c["rewards"], c["times"])
        assert calc.avg_num_jobs == c["avg_num_jobs"]
        for got, want in zip(rets, c["returns"]):
            assert np.array_equal(got, want)
