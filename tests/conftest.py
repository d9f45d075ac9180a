import glob
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


def golden_names():
    return sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


class Golden:
    """One reference-generated fixture (see oracle/make_golden.py)."""

    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.cm = z["cm"].astype(np.int64)
        self.n, self.m = self.cm.shape
        self.algo = str(z["algo"])
        self.missing = int(z["missing"])
        self.prior_transformation = str(z["prior_transformation"])
        self.priors = z["priors"] if z["priors"].size else None
        self.order = z["order"]
        self.joins = z["joins"]
        self.D = z["D"] if "D" in z.files else None

    @property
    def rooted_cm(self):
        """Character matrix rows as the reference's map sees them (root row appended for NJ)."""
        cm = self.cm
        if self.algo == "nj":
            cm = np.vstack([cm, np.zeros((1, self.m), dtype=np.int64)])
        return cm

    @property
    def ordered_cm(self):
        return self.rooted_cm[self.order]

    @property
    def weights(self):
        """Dense [m, S+1] weight table with the reference's numpy arithmetic (NaN = no prior)."""
        if self.priors is None:
            return None
        p = self.priors
        with np.errstate(divide="ignore", invalid="ignore"):
            if self.prior_transformation == "negative_log":
                return -np.log(p)
            if self.prior_transformation == "inverse":
                return 1 / p
            return np.sqrt(1 / p)

    def prior_dict(self):
        if self.priors is None:
            return None
        out = {}
        for c in range(self.priors.shape[0]):
            out[c] = {s: float(self.priors[c, s]) for s in range(self.priors.shape[1]) if not np.isnan(self.priors[c, s])}
        return out

    def clades(self, collapsed):
        tag = "collapsed" if collapsed else "plain"
        return (self.z[f"clades_{tag}_hash"], self.z[f"clades_{tag}_size"], self.z[f"clades_{tag}_depth"])


@pytest.fixture(scope="session")
def goldens():
    return {n: Golden(n) for n in golden_names()}
