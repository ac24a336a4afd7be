import os
import os.path as osp
import sys

import pytest

REPO = osp.dirname(osp.dirname(osp.abspath(__file__)))
for p in (REPO, osp.join(REPO, "oracle"), osp.dirname(osp.abspath(__file__))):
    if p not in sys.path:
        sys.path.insert(0, p)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session")
def bank():
    import spark_sched_sim_b200.bank as bankmod

    return bankmod.synthetic_bank(0)
