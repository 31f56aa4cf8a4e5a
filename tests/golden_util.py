"""Shared loading of tests/golden/*.npz (written by tests/golden/make_golden.py from the unmodified reference)."""
import glob
import os

import numpy as np
import torch

from femasr_b200.spec import random_state_dict

GOLDEN = sorted(glob.glob(os.path.join(os.path.dirname(__file__), "golden", "*.npz")))
IDS = [os.path.basename(p)[:-4] for p in GOLDEN]


def codebooks_of(g):
    """codebook_params rows [[scale, n_e, e_dim], ...] of a fixture (older fixtures: one codebook at 32)."""
    if "codebooks" in g.files:
        return [[int(v) for v in row] for row in g["codebooks"]]
    return [[32, 1024, int(g["e_dim"])]]


def gt_indices_of(g):
    ks = sorted(k for k in g.files if k.startswith("gt_indices"))
    return [torch.from_numpy(g[k]) for k in ks] if ks else None


def indices_of(g):
    out = [g["indices"]]
    k = 1
    while f"indices{k}" in g.files:
        out.append(g[f"indices{k}"])
        k += 1
    return out


def load_case(path):
    g = np.load(path)
    cbs = codebooks_of(g)
    sd = random_state_dict(int(g["scale"]), int(g["e_dim"]), seed=int(g["seed"]), init=str(g["init"]), codebooks=cbs)
    return g, sd, cbs
