import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")
OUT_KEYS = ("pred_logits", "center_unnormalized", "size_unnormalized", "ortho6d", "sem_cls_prob", "coord_pos")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100) GPU; run with `pytest -m gpu` on the GPU box")


def pytest_collection_modifyitems(config, items):
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def regenerate_case(gold):
    """Re-create the inputs a fixture was generated from and verify their checksums."""
    from parq_b200 import inputs as I
    B, T, H, W, Nq, seed, wild, smooth = [int(x) for x in gold["shape"]]
    tokens = I.make_tokens(B, T, H, W, seed=seed, smooth=bool(smooth))
    cam, Tcp, Twp, Twl = I.make_geometry(B, T, H, W, seed=seed, wild=bool(wild))
    case = dict(B=B, T=T, H=H, W=W, Nq=Nq, seed=seed, tokens=tokens, camera=cam._data, T_cp=Tcp._data, T_wp=Twp._data, T_wl=Twl._data)
    if "points" in gold:
        case["points"] = torch.from_numpy(gold["points"])
        assert I.tensor_checksum(tokens, cam._data, Tcp._data, Twp._data, Twl._data, case["points"]) == str(gold["inputs_sum"])
    else:
        assert I.tensor_checksum(tokens, cam._data, Tcp._data, Twp._data, Twl._data) == str(gold["inputs_sum"]), \
            "synthetic input generator drifted from the one the fixture was made with"
        sd = I.make_weights(seed, Nq)
        assert I.tensor_checksum(*[sd[k] for k in sorted(sd)]) == str(gold["weights_sum"]), "weight generator drifted"
        case["sd"] = sd
    return case


def relerr(a, b):
    """max|a-b| / max|b|: the "1e-3 relative" metric of the parity bar."""
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return ((a - b).abs().max() / b.abs().max().clamp(min=1e-30)).item()


def bit_equal(a, b):
    a = np.ascontiguousarray(torch.as_tensor(a).cpu().numpy())
    b = np.ascontiguousarray(torch.as_tensor(b).cpu().numpy())
    return a.shape == b.shape and np.array_equal(a.view(np.uint8), b.view(np.uint8))
