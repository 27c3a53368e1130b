import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden():
    with open(os.path.join(GOLDEN_DIR, "golden.json")) as f:
        g = json.load(f)
    z = np.load(os.path.join(GOLDEN_DIR, "testfa.npz"))
    g["seqs"] = {"seq1": z["seq1"], "seq2": z["seq2"]}
    return g


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.build(ref=True)       # compiles the checker; oracle/_ref only where /root/reference exists
    return O.Oracle()


@pytest.fixture(scope="session")
def reflib():
    from oracle import oracle as O
    if not O.RefLib.available():
        pytest.skip("oracle/_ref/libssw.so not built (no /root/reference here)")
    return O.RefLib()
