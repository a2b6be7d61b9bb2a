import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: test needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    import torch

    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device visible")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    import numpy as np
    import torch

    data = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    # numeric arrays as tensors; string arrays (log lines, names) stay numpy
    return {k: (torch.from_numpy(data[k]) if data[k].dtype.kind in "fiub" else data[k]) for k in data.files}


EPS_GOLDEN_CASES = [
    "eps_c2_k2_single_pixel",
    "eps_c1_k3_two_pixels",
    "eps_c1_k2_q2",
    "eps_c1_k2_q3",
    "eps_c2_k2_q2_4x5",
    "eps_c3_k1_q2",
    "eps_c1_k4_q2",
    "eps_c1_k2_q6",
    "eps_c1_k3_q3",
]
LME_GOLDEN_CASES = ["lme_small", "lme_64", "lme_ragged", "lme_scale150"]
CONVSBS_CASES = ["convsbs_as_eps_perm0", "convsbs_as_eps_perm7", "convsbs_as_eps_perm23"]
CONVSBS_LOG_CASES = ["convsbs_log_2x2_ring", "convsbs_log_2x2_perm", "convsbs_log_3x3_snake_ring", "convsbs_log_3x3_snake_open"]
WINDOW_STATS_CASES = ["stats_windows_c1_k3", "stats_windows_c2_k2", "stats_windows_c1_k4_ragged"]
LME_BATCHED_CASES = ["lme_batched_small", "lme_batched_r8", "lme_batched_scale150"]


def convsbs_log_case(g):
    """(log_cores, positions, log_x) of a convsbs_log_* golden file."""
    n = len(g["outs"])
    return [g[f"log_core{i}"] for i in range(n)], [tuple(int(v) for v in p) for p in g["positions"]], g["log_x"]
