"""Template bank: the reference's on-disk TPC-H layout round-trips through load_tpch_dir, and the
executor-interval table matches tpch.py:237-262."""
import numpy as np

import spark_sched_sim_b200.bank as bankmod


def test_on_disk_layout_roundtrip(tmp_path, bank):
    data = bankmod.make_synthetic_tpch(0)
    bankmod.write_tpch_dir(str(tmp_path), data)
    loaded = bankmod.load_tpch_dir(str(tmp_path))
    assert loaded.keys() == data.keys()
    b2 = bankmod.build_bank(loaded)
    assert b2.checksum() == bank.checksum()
    assert b2.num_templates == 154 and b2.max_stages <= 64


def test_clean_first_wave_multiset_semantics():
    stage = {"first_wave": {5: [10, 10, 20, 30], 10: [7], 20: []},
             "fresh_durations": {5: [10, 30, 30], 10: [7], 20: []}, "rest_wave": {5: [], 10: [], 20: []}}
    clean = bankmod._clean_first_wave(stage)
    assert clean[5] == [10, 20]          # one 10 and the 30 are "fresh" (multiset removal)
    assert clean[10] == [10, 20]         # emptied level inherits the nearest lower level's list
    assert clean[20] == [10, 20]


def test_executor_intervals():
    iv = bankmod.executor_intervals(10)
    assert iv.shape == (11, 2)
    assert (iv[:6] == 5).all() and (iv[6:10] == (5, 10)).all() and (iv[10] == 10).all()
    iv = bankmod.executor_intervals(50)
    assert (iv[11:20] == (10, 20)).all() and (iv[20] == 20).all() and (iv[41:50] == (40, 50)).all()
    assert (iv[50] == 50).all()
