import json
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
if str(ROOT) not in sys.path:
    sys.path.insert(0, str(ROOT))

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (run with -m gpu on the GPU box)")


def _has_gpu():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    """(iq bytes, DemodResult of the reference, meta dict) of tests/golden/<name>.npz"""
    from readsb_protobuf_b200.results import DemodResult
    z = np.load(GOLDEN / f"{name}.npz")
    meta = json.loads(bytes(z["meta"]).decode())
    iq = z["iq"]
    return iq, DemodResult(z["msgs"], z["stats"][0], z["blocks"], iq.size // (2 if meta["fmt"] == "uc8" else 4)), meta


GOLDEN_NAMES = ["uc8_fix1", "uc8_fix2_aggressive", "uc8_nofix_thr75", "uc8_whole_blocks", "sc16", "sc16q11", "kat_frame", "uc8_modeac", "uc8_df18", "uc8_all_df", "uc8_all_df_nofix",
                "uc8_dcfilter", "sc16_dcfilter", "sc16q11_table8"]
