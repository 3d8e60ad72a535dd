import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    """GPU tests must never silently pass on a machine without a GPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc
    orc.build()
    return orc


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name), allow_pickle=False)


def golden_noise_tkn(seed, K, T, nu, sigma):
    """Same recipe as oracle/make_golden.py:golden_noise, in the native [T][K][nu] layout."""
    import torch
    g = torch.Generator().manual_seed(int(seed))
    n = torch.randn(K, T, nu, generator=g, dtype=torch.float32) * float(sigma)
    return n.permute(1, 0, 2).contiguous().numpy()


def fixture_noise(gold, i, nu):
    """Noise of step i of a fixture: stored, or regenerated from its seed and checked."""
    key = f"noise_{i}"
    if key in gold.files:
        return gold[key]
    K, T = int(gold["K"]), int(gold["T"])
    n = golden_noise_tkn(gold["seeds"][i], K, T, nu, float(gold["sigma"]))
    s, a = n.astype(np.float64).sum(), np.abs(n.astype(np.float64)).sum()
    assert abs(s - float(gold[f"noise_sum_{i}"])) < 1e-6 * a, "torch CPU generator drifted"
    assert abs(a - float(gold[f"noise_abs_sum_{i}"])) < 1e-9 * a, "torch CPU generator drifted"
    return n


def rel_inf(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))
