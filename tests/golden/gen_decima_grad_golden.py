"""Regenerates tests/golden/decima_grads_*.npz: gradients of the UNMODIFIED reference DecimaScheduler (shipped
models/decima/model.pt, PyG stand-ins of oracle/refshim) through its own evaluate_actions
(schedulers/decima/scheduler.py:101-139) + loss.backward(), on observations of the recorded Decima-driven episodes.
Build container only (needs /root/reference).  Usage: python tests/golden/gen_decima_grad_golden.py

For a fixture's episode (same env config, seeds and policy seed as the golden trace of that name, so the trajectory
is the recorded one) the observations of a few decisions and the actions taken there are stored by the reference's
wrapper stack; evaluate_actions re-evaluates them in ONE batch; the scalar  sum_i c_i lgprob_i + e_i entropy_i  (fixed
coefficients, stored) is differentiated; the 42 parameter gradients are stored flattened in state_dict order."""
from __future__ import annotations

import importlib
import os.path as osp
import sys

import numpy as np

HERE = osp.dirname(osp.abspath(__file__))
REPO = osp.dirname(osp.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, osp.join(REPO, "oracle"))
sys.path.insert(0, osp.join(REPO, "tests"))

CASES = {
    # fixture name: (golden trace it follows, decisions whose observations enter the batch)
    "decima_grads_e10_j8_s5": ("decima_e10_j8_s5_philox", [3, 17, 40, 75, 110, 150]),
    "decima_grads_e50_j14_s4": ("decima_e50_j14_s4_philox", [5, 60, 130, 222, 301, 377, 399]),
}


def main():
    import refrun
    from helpers import load_golden

    refrun.setup()  # puts the gymnasium / PyG stand-ins and the reference on sys.path
    import gymnasium
    import torch
    from philox_ref import PhiloxNpRandom

    for name, (trace_name, picks) in CASES.items():
        tr = load_golden(trace_name)
        refrun.setup(tr["bank_seed"], tr["bank_kind"])
        gymnasium.NP_RANDOM_FACTORY = PhiloxNpRandom
        env_cfg = {"num_executors": tr["num_executors"], "job_arrival_cap": tr["job_arrival_cap"],
                   "job_arrival_rate": tr["job_arrival_rate"], "moving_delay": tr["moving_delay"],
                   "warmup_delay": tr["warmup_delay"], "data_sampler_cls": "TPCHDataSampler"}
        env = gymnasium.make("spark_sched_sim:SparkSchedSimEnv-v0", env_cfg=env_cfg)
        sched = refrun.make_policy("decima", tr["num_executors"], tr["policy_seed"])
        wenv = sched.env_wrapper_cls(env)
        obs, _ = wenv.reset(seed=tr["seed"])
        kept_obs, kept_act = [], []
        for k in range(max(picks) + 1):
            action, _ = sched.schedule(obs)
            assert (int(action["stage_idx"]), int(action["job_idx"]), int(action["num_exec"])) == tuple(
                int(x) for x in tr["pol_actions"][k]), (k, "the episode must be the recorded one")
            if k in picks:
                kept_obs.append(obs)
                kept_act.append(tuple(action.values()))  # RolloutBuffer.add(obs, wall_time, tuple(action.values()), ...)
            obs, *_ = wenv.step(action)
        n = len(picks)
        rng = np.random.default_rng(7)
        c = rng.uniform(-1.5, 1.5, n).astype(np.float32)
        e = rng.uniform(-0.6, 0.6, n).astype(np.float32)
        sched.train()
        sched.zero_grad()
        res = sched.evaluate_actions(kept_obs, kept_act)
        loss = (torch.from_numpy(c) * res["lgprobs"]).sum() + (torch.from_numpy(e) * res["entropies"]).sum()
        loss.backward()
        order = list(sched.state_dict().keys())
        params = dict(sched.named_parameters())
        grad = np.concatenate([params[k].grad.detach().numpy().reshape(-1) for k in order]).astype(np.float32)
        assert grad.size == 20802
        np.savez_compressed(
            osp.join(HERE, name + ".npz"), trace=trace_name, picks=np.array(picks, np.int32), coef_lgprob=c,
            coef_entropy=e, lgprobs=res["lgprobs"].detach().numpy().astype(np.float32),
            entropies=res["entropies"].detach().numpy().astype(np.float32), loss=np.float64(loss.item()), grad=grad,
            param_order=np.array(order))
        print(name, "loss", float(loss.item()), "|grad|", float(np.abs(grad).max()), "n", n)


if __name__ == "__main__":
    main()
