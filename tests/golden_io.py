"""Loader for tests/golden/*.npz (written by tests/golden/make_golden.py)."""
import glob
import json
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def list_cases(pattern="*"):
    return sorted(os.path.splitext(os.path.basename(p))[0]
                  for p in glob.glob(os.path.join(GOLDEN_DIR, pattern + ".npz")))


def load(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    meta = json.loads(bytes(z["__meta__"]).decode())
    arrays = {k: torch.from_numpy(np.array(z[k])) for k in z.files if k != "__meta__"}
    return meta, arrays


def model_state(arrays, key):
    pre = f"model.{key}."
    return {k[len(pre):]: v for k, v in arrays.items() if k.startswith(pre)}


def expected(arrays):
    return {k[4:]: v for k, v in arrays.items() if k.startswith("out.")}
