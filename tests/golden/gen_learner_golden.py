"""Generates tests/golden/learner_vectors.npz by running the REFERENCE's own ReturnsCalculator and Baseline
(pure numpy: /root/reference/trainers/utils/{returns_calculator,baselines}.py, imported unmodified) on seeded
rollout-shaped inputs: non-decreasing wall times with repeats (same-round decisions), ragged lengths.

    python tests/golden/gen_learner_golden.py        (build container only: needs /root/reference)
"""
import importlib.util
import os.path as osp

import numpy as np

REF = "/root/reference/trainers/utils"


def load(name):
    spec = importlib.util.spec_from_file_location(name, osp.join(REF, name + ".py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


def main():
    rc, bl = load("returns_calculator"), load("baselines")
    rng = np.random.default_rng(20260101)
    out = {}
    cases = [("r4", 2, 4, 5e-3), ("r8", 1, 8, 5e-3), ("r12", 1, 12, 1e-3), ("r3", 3, 3, 2e-2)]
    for name, num_seq, num_roll, beta in cases:
        n = num_seq * num_roll
        times_list, rewards_list = [], []
        for i in range(n):
            K = int(rng.integers(5, 160))
            # wall times: whole-ms event times mixed with fractional arrival times, ~35 % repeats
            steps = np.where(rng.random(K) < 0.35, 0.0, np.round(rng.exponential(4000.0, K)))
            steps += np.where(rng.random(K) < 0.1, rng.random(K), 0.0) * (steps > 0)
            ts = np.concatenate([[0.0], np.cumsum(steps)])
            times_list.append(ts.tolist())
            rewards_list.append((-(ts[1:] - ts[:-1]) * rng.integers(1, 30, K)).tolist())
        calc = rc.ReturnsCalculator(beta=beta)
        returns_list = calc(rewards_list, times_list, [set()] * n)
        base = bl.Baseline(num_seq, num_roll)
        baselines_list = base([ts[:-1] for ts in times_list], returns_list)
        out[f"{name}_meta"] = np.array([num_seq, num_roll, beta])
        out[f"{name}_len"] = np.array([len(r) for r in rewards_list])
        out[f"{name}_times"] = np.concatenate([np.asarray(t) for t in times_list])
        out[f"{name}_rewards"] = np.concatenate([np.asarray(r) for r in rewards_list])
        out[f"{name}_returns"] = np.concatenate(returns_list)
        out[f"{name}_baselines"] = np.concatenate(baselines_list)
    # differential returns: one calculator, three consecutive calls (the window carries over); caps below and above
    # the number of rows per call
    for name, cap in (("diff_small", 150), ("diff_large", 5000)):
        calc = rc.ReturnsCalculator(buff_cap=cap)
        for call in range(3):
            n = 6
            times_list, rewards_list = [], []
            for i in range(n):
                K = int(rng.integers(20, 120))
                steps = np.where(rng.random(K) < 0.35, 0.0, np.round(rng.exponential(4000.0, K)))
                ts = np.concatenate([[0.0], np.cumsum(steps)])
                times_list.append(ts.tolist())
                rewards_list.append((-(ts[1:] - ts[:-1]) * rng.integers(1, 30, K)).tolist())
            returns_list = calc(rewards_list, times_list, [set()] * n)
            key = f"{name}_c{call}"
            out[f"{key}_meta"] = np.array([cap, calc.avg_num_jobs])
            out[f"{key}_len"] = np.array([len(r) for r in rewards_list])
            out[f"{key}_times"] = np.concatenate([np.asarray(t) for t in times_list])
            out[f"{key}_rewards"] = np.concatenate([np.asarray(r) for r in rewards_list])
            out[f"{key}_returns"] = np.concatenate(returns_list)
    np.savez_compressed(osp.join(osp.dirname(osp.abspath(__file__)), "learner_vectors.npz"), **out)
    print({k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    main()
