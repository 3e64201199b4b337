import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def unpack_instances(z, prefix):
    import numpy as np
    labels = z[prefix + "labels"]
    lens = z[prefix + "lens"]
    offs = np.concatenate([[0], np.cumsum(lens)])
    out = {}
    for i, l in enumerate(labels):
        out[int(l)] = {"box": tuple(int(v) for v in z[prefix + "boxes"][i]),
                       "starts": z[prefix + "starts"][offs[i]:offs[i + 1]],
                       "runs": z[prefix + "runs"][offs[i]:offs[i + 1]]}
    return out


def assert_instances_equal(a, b):
    import numpy as np
    assert list(a.keys()) == list(b.keys()), (list(a.keys()), list(b.keys()))
    for k in a:
        assert tuple(int(v) for v in a[k]["box"]) == tuple(int(v) for v in b[k]["box"]), k
        assert np.array_equal(np.asarray(a[k]["starts"]), np.asarray(b[k]["starts"])), k
        assert np.array_equal(np.asarray(a[k]["runs"]), np.asarray(b[k]["runs"])), k
