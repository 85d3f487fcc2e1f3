"""Shared helpers for the test-suite (golden loader, CUDA probe, synthetic generators)."""

import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name)) as data:
        return {key: data[key] for key in data.files}


def has_cuda():
    try:
        import torch

        return torch.cuda.is_available()
    except Exception:
        return False
