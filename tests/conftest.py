import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, f"{name}.npz")) as z:
        return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def unflatten(d, prefix):
    return {k[len(prefix):]: v for k, v in d.items() if k.startswith(prefix)}


@pytest.fixture(scope="session")
def small_params():
    w = load_golden("small_weights")
    return dict(sdf=unflatten(w, "sdf."), color=unflatten(w, "color."), variance=unflatten(w, "variance."))


def assert_close(a, b, tol, name=""):
    a, b = torch.as_tensor(a).double().cpu(), torch.as_tensor(b).double().cpu()
    assert a.shape == b.shape, f"{name}: shape {tuple(a.shape)} vs {tuple(b.shape)}"
    err = (a - b).abs().max().item() if a.numel() else 0.0
    scale = max(1.0, b.abs().max().item() if b.numel() else 1.0)
    assert err <= tol * scale, f"{name}: max abs err {err:.3e} > {tol:.1e} * {scale:.3g}"


def rel_err(a, b):
    a, b = torch.as_tensor(a).double().cpu().flatten(), torch.as_tensor(b).double().cpu().flatten()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cos_sim(a, b):
    a, b = torch.as_tensor(a).double().cpu().flatten(), torch.as_tensor(b).double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()
