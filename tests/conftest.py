import glob
import json
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

GOLDEN_DIR = os.path.join(REPO, "tests", "golden")
GOLDEN_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


class Golden:
    """One golden case produced by tests/golden/make_golden.py from the unmodified reference."""

    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.name = name
        self.z = z
        self.cfg = json.loads(str(z["cfg"]))
        self.model_params = {k[len("model."):]: z[k] for k in z.files if k.startswith("model.")}
        self.score_params = {k[len("score."):]: z[k] for k in z.files if k.startswith("score.")}
        self.n = z["x"].shape[0]

    def __getitem__(self, k):
        return self.z[k]

    def sets(self):
        return {t: (self.z[f"set_{t}_ix"], self.z[f"set_{t}_src"], self.z[f"set_{t}_tgt"])
                for t in ("cn", "1hop", "non1hop") if f"set_{t}_ix" in self.z.files}


@pytest.fixture(params=GOLDEN_CASES)
def golden(request):
    return Golden(request.param)
