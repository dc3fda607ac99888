"""Shared test helpers: golden loading and tolerances (tests only)."""
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# north_star: logits / attention outputs / gradients within 1e-3 relative error.  Element-wise
# relative error is meaningless next to zero (SURVEY.md section 7 "hard parts"), so the bound used
# everywhere is  |a-b| <= RTOL * max|ref|  over the compared tensor (row-max for logits).
RTOL = 1e-3


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd.")}
    rest = {k: z[k] for k in z.files if not k.startswith("sd.")}
    return sd, rest


def max_rel(a, b):
    a = torch.as_tensor(a, dtype=torch.float64)
    b = torch.as_tensor(b, dtype=torch.float64)
    denom = b.abs().max().clamp_min(1e-30)
    return float((a - b).abs().max() / denom)


def assert_close_rel(a, b, rtol=RTOL, what=""):
    r = max_rel(a, b)
    assert r <= rtol, f"{what}: max|a-b|/max|ref| = {r:.3e} > {rtol:g}"
