"""Pins oracle/learner.py (discounted returns, interpolated group baseline) to the outputs of the reference's
own ReturnsCalculator / Baseline classes recorded in tests/golden/learner_vectors.npz."""
import os.path as osp

import numpy as np
import pytest

from helpers import GOLDEN_DIR


def load_cases():
    z = np.load(osp.join(GOLDEN_DIR, "learner_vectors.npz"))
    cases = {}
    for name in sorted({k.rsplit("_", 1)[0] for k in z.files}):
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


CASES = load_cases()


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
