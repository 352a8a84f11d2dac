import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


@pytest.fixture(scope="session")
def golden():
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = dict(np.load(os.path.join(GOLDEN, name + ".npz")))
        return cache[name]

    return load


def rel_linf(a, b):
    """max|a-b| / max|b| -- the relative L-infinity norm used by every parity test."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def expected_index_boundary_distance(prob):
    """prob [..., D, H, W] (the reference's softmax over depth, any float type) -> |e - rint(e)| with e = sum_d p_d * d the
    expected hypothesis index in float64.  The photometric confidences truncate / window around e (MVSNet model.py:211-215:
    `.long()`; Vis nn_utils.py:463-465: |d - e| <= 2), i.e. INDEX work: the only pixels where two correct fp32
    evaluations may disagree are those whose e sits on an integer."""
    prob = np.asarray(prob, np.float64)
    D = prob.shape[-3]
    e = (prob * np.arange(D, dtype=np.float64).reshape((D, 1, 1))).sum(-3)
    return np.abs(e - np.rint(e))


def assert_mismatches_on_boundary(got, want, boundary_dist, tol, eps, what):
    """Every element where |got - want| > tol must have boundary_dist < eps (it sits on a truncation / threshold boundary);
    returns the number of such elements (reported by the caller)."""
    got, want = np.asarray(got, np.float64), np.asarray(want, np.float64)
    bad = np.abs(got - want) > tol
    off = bad & ~(np.asarray(boundary_dist) < eps)
    assert not off.any(), "%s: %d of %d mismatching elements are NOT on a boundary (worst distance %.3g, worst diff %.3g)" % (
        what, int(off.sum()), int(bad.sum()), float(np.asarray(boundary_dist)[off].max()), float(np.abs(got - want)[off].max()))
    return int(bad.sum())
