"""Regenerates tests/golden/*.npz by running the UNMODIFIED reference (needs /root/reference; build
container only).  Usage: python tests/golden/gen_golden.py [case ...]

Each fixture is one reference episode recorded by oracle/refrun.py: job sequence, duration tape,
actions, everything step() returned, every popped event and the final job completion times.
"small" cases keep full observations and event rows; "slim" cases (the 50-job headline config) keep
per-step digests instead.  `rng`: "pcg64" = the reference's own numpy Generator (replayed through the
duration tape); "philox" = the counter-based stream of oracle/philox_ref.py plugged into the
reference's sampler through gymnasium's seeding hook (replayed from the seed alone).
"""
from __future__ import annotations

import os.path as osp
import sys

import numpy as np

HERE = osp.dirname(osp.abspath(__file__))
REPO = osp.dirname(osp.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, osp.join(REPO, "oracle"))

C2 = {"num_executors": 10, "job_arrival_cap": 50, "job_arrival_rate": 4.0e-5,
      "moving_delay": 2000.0, "warmup_delay": 1000.0}  # examples.py:15-23


def cfg(E, J, **kw):
    c = {"num_executors": E, "job_arrival_cap": J, "job_arrival_rate": 4.0e-5,
         "moving_delay": 2000.0, "warmup_delay": 1000.0}
    c.update(kw)
    return c


# name: (env_cfg, policy, seed, rng, time_limit, slim)
CASES = {
    "c2_fair_s1234_pcg64": (C2, "fair", 1234, "pcg64", None, True),
    "c2_fair_s1234_philox": (C2, "fair", 1234, "philox", None, True),
    "c2_random_s7_philox": (C2, "random", 7, "philox", None, True),
    # BASELINE config 4's episode shape: 200 jobs x 50 executors (two executor slots per lane on the device)
    "c4_fair_s21_philox": (cfg(50, 200), "fair", 21, "philox", None, True),
    # continuous Poisson arrivals until a time limit (no job cap), 50 executors
    "c4_tl_fifo_s22_philox": (cfg(50, None), "fifo", 22, "philox", 3.0e6, True),
    "e10_j8_fair_s1_pcg64": (cfg(10, 8), "fair", 1, "pcg64", None, False),
    "e10_j8_fair_s2_philox": (cfg(10, 8), "fair", 2, "philox", None, False),
    "e10_j8_fifo_s3_philox": (cfg(10, 8), "fifo", 3, "philox", None, False),
    "e10_j8_random_s4_pcg64": (cfg(10, 8), "random", 4, "pcg64", None, False),
    "e10_j8_random_s5_philox": (cfg(10, 8), "random", 5, "philox", None, False),
    "e10_j6_random_s6_philox_beta": (cfg(10, 6, beta=5e-3), "random", 6, "philox", None, False),
    "e50_j8_fair_s7_philox": (cfg(50, 8), "fair", 7, "philox", None, False),
    "e50_j8_random_s8_philox": (cfg(50, 8), "random", 8, "philox", None, False),
    "e50_j8_random_s9_pcg64": (cfg(50, 8), "random", 9, "pcg64", None, False),
    "e10_tl_fair_s10_philox": (cfg(10, None), "fair", 10, "philox", 4.0e5, False),
    "e10_tl_random_s11_philox": (cfg(10, None), "random", 11, "philox", 5.0e5, False),
    "e3_j5_random_s12_philox": (cfg(3, 5), "random", 12, "philox", None, False),
    "e1_j3_fair_s13_philox": (cfg(1, 3), "fair", 13, "philox", None, False),
    # driven by the reference's DecimaScheduler with the shipped models/decima/model.pt; these also
    # hold the policy's stage / executor-count scores for every decision (`pol_*`)
    "decima_e10_j8_s5_philox": (cfg(10, 8), "decima", 5, "philox", None, False),
    "decima_e50_j6_s3_philox": (cfg(50, 6), "decima", 3, "philox", None, False),
    "decima_e50_j14_s4_philox": (cfg(50, 14), "decima", 4, "philox", None, False),
    # Decima at the configured scales (slim: per-observation digests of the adapter's outputs, the recorded scores of
    # every (8th) decision): the headline shape, and config/decima_tpch.yaml:80-87 (50 executors, 200 jobs)
    "decima_c2_s1234_philox": (C2, "decima", 1234, "philox", None, True),
    "decima_c3_s42_philox": (cfg(50, 200), "decima", 42, "philox", None, True, {"logit_stride": 8}),
    # templates of 30..64 stages ("wide" bank): the 64-bit halves of every stage bitmask, adapter path for ns > 32
    "wide_e10_j6_fair_s31_philox": (cfg(10, 6), "fair", 31, "philox", None, False, {"bank_kind": "wide"}),
    "wide_e50_j5_random_s32_philox": (cfg(50, 5), "random", 32, "philox", None, False, {"bank_kind": "wide"}),
    "decima_wide_e10_j5_s33_philox": (cfg(10, 5), "decima", 33, "philox", None, False, {"bank_kind": "wide"}),
}


def export_decima_weights():
    """models/decima/model.pt (42 tensors, 20 802 fp32 parameters) -> decima_model.npz fixture."""
    import torch

    sd = torch.load(osp.join("/root/reference", "models", "decima", "model.pt"), map_location="cpu")
    np.savez_compressed(osp.join(HERE, "decima_model.npz"), **{k: v.numpy() for k, v in sd.items()})


def main(argv):
    import refrun
    import spark_sched_sim_b200.bank as bankmod

    names = argv or list(CASES)
    if any(n.startswith("decima_") for n in names):
        export_decima_weights()
    for name in names:
        env_cfg, policy, seed, rng, tl, slim = CASES[name][:6]
        extra = CASES[name][6] if len(CASES[name]) > 6 else {}
        kind = extra.get("bank_kind", "appd")
        # small cases (and every Decima-driven one) also record what the reference's DecimaObsWrapper makes of
        # every observation
        tr = refrun.run_episode(env_cfg, policy, seed, rng=rng, time_limit=tl,
                                decima=(not slim) or policy == "decima", bank_kind=kind)
        if slim:
            tr = refrun.slim(tr, extra.get("logit_stride", 1))
        tr["bank_checksum"] = bankmod.synthetic_bank(0, kind).checksum()
        path = osp.join(HERE, name + ".npz")
        np.savez_compressed(path, **tr)
        print(f"{name}: steps={len(tr['actions'])} launches={len(tr['tape'])} "
              f"jobs={len(tr['job_template'])} -> {osp.getsize(path)/1024:.0f} KiB")


if __name__ == "__main__":
    main(sys.argv[1:])
