"""Pins oracle/decima_obs.py (numpy restatement of DecimaObsWrapper + make_dag_layer_edge_masks)
to what the reference's own wrapper produced on every observation of the small golden episodes:
features bit-exact in float32, commit caps (= exec_mask row sums), stage mask, per-level edge masks."""
import numpy as np
import pytest

import decima_obs
from helpers import golden_names, load_golden


def iter_obs(tr):
    n = e = d = s = 0
    for k in range(len(tr["N"])):
        N, M, Ja = int(tr["N"][k]), int(tr["M"][k]), int(tr["Ja"][k])
        yield k, {"nodes": tr["nodes"][n:n + N], "edge_links": tr["edges"][e:e + M],
                  "dag_ptr": tr["dag_ptr"][d:d + Ja + 1], "exec_supplies": tr["supplies"][s:s + Ja],
                  "num_committable_execs": int(tr["ncommit"][k]), "source_job_idx": int(tr["src"][k])}, (n, e, s)
        n += N; e += M; d += Ja + 1; s += Ja


@pytest.mark.parametrize("name", golden_names(slim=False))
def test_decima_obs_oracle_matches_reference_wrapper(name):
    tr = load_golden(name)
    E = tr["num_executors"]
    for k, obs, (n, e, s) in iter_obs(tr):
        N, M, Ja = obs["nodes"].shape[0], obs["edge_links"].shape[0], len(obs["exec_supplies"])
        d = decima_obs.decima_observation(obs, E)
        assert np.array_equal(d["features"], tr["dec_feat"][n:n + N]), (k, "features")
        assert np.array_equal(d["stage_mask"], tr["dec_stage_mask"][n:n + N].astype(bool))
        assert np.array_equal(d["commit_caps"], tr["dec_caps"][s:s + Ja]), (k, "caps")
        assert d["depth"] == tr["dec_depth"][k], (k, "depth")
        assert np.array_equal(d["edge_bits"], tr["dec_edge_bits"][e:e + M]), (k, "edge masks")


def test_mask_helpers_roundtrip():
    caps = np.array([0, 3, 10])
    m = decima_obs.exec_mask_from_caps(caps, 10)
    assert m.shape == (3, 10) and m.sum(1).tolist() == [0, 3, 10] and m[1, :3].all()
    bits = np.array([0b101, 0b010, 0], np.uint64)
    em = decima_obs.edge_masks_from_bits(bits, 3)
    assert em.tolist() == [[True, False, False], [False, True, False], [True, False, False]]
    assert decima_obs.edge_masks_from_bits(bits, 0).shape == (0, 3)


def test_float32_division_equals_the_reference_double_then_float32():
    """The CTA adapter (Sim::decima_obs_job_w) computes cap / E and supply / E with one float32 division; the
    reference divides Python floats (double) and stores into a float32 array (env_wrapper.py:110-127).  Same bits
    for every integer pair the adapter can see (0 <= x <= E <= 128)."""
    for E in range(1, 129):
        x = np.arange(0, E + 1)
        via_double = (x.astype(np.float64) / np.float64(E)).astype(np.float32)
        direct = x.astype(np.float32) / np.float32(E)
        assert (via_double == direct).all(), E
