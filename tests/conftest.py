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


def golden_files():
    return sorted(f for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith(("jacx_", "densex_", "zebra_", "select_", "eval_", "init_", "cov2d_")))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name))
    d = {k: z[k] for k in z.files}
    d["valid"] = d["in_valid"] if bool(d["has_valid"]) else None
    return d


def rel_err(a, b):
    """max over samples of ||a_b - b_b|| / ||b_b|| (first dim = sample)."""
    a = np.asarray(a, dtype=np.float64).reshape(len(a), -1)
    b = np.asarray(b, dtype=np.float64).reshape(len(b), -1)
    den = np.linalg.norm(b, axis=1)
    den = np.where(den > 0, den, 1.0)
    return float((np.linalg.norm(a - b, axis=1) / den).max())


def quat_angle(qa, qb):
    """geodesic angle (rad) between unit-ish wxyz quaternions, per sample."""
    qa = qa / np.linalg.norm(qa, axis=-1, keepdims=True)
    qb = qb / np.linalg.norm(qb, axis=-1, keepdims=True)
    d = np.abs((qa * qb).sum(-1)).clip(max=1.0)
    # 2*acos(d) loses precision near d=1; use the chord instead
    diff = np.minimum(np.linalg.norm(qa - qb, axis=-1), np.linalg.norm(qa + qb, axis=-1))
    return 2.0 * np.arcsin(np.clip(diff / 2.0, 0, 1))


@pytest.fixture(scope="session")
def oracle():
    from oracle import cpu_oracle
    cpu_oracle.build()
    return cpu_oracle
