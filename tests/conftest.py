import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")
REFERENCE = "/root/reference"   # only present in the build container; never read by -m gpu tests


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (runs on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """`gpu` tests need a CUDA device: on a CPU-only box they are skipped, not failed (plain `pytest tests` stays green)."""
    try:
        import torch
        have = torch.cuda.is_available()
    except Exception:
        have = False
    if have:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (run on the B200 box with -m gpu)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """Unpack one tests/golden/*.npz written by tests/golden/make_golden.py."""
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    shape = tuple(int(v) for v in z["shape"])
    bits = np.unpackbits(z["copies_bits"])[: int(np.prod(shape))].reshape(shape)
    copies = bits.astype(np.float32) * np.float32(z["value"])
    out = {k: z[k] for k in z.files}
    out["copies"] = copies
    return out


def golden_params(g):
    """The SolveParams repr stored with the golden -> kwargs dict."""
    import re
    txt = str(g["params"])
    body = txt[txt.index("(") + 1: txt.rindex(")")]
    kw = {}
    for part in re.split(r",\s*(?=[a-z_0-9]+=)", body):
        k, v = part.split("=", 1)
        kw[k.strip()] = eval(v)
    return kw


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as O
    O.lib()
    return O
