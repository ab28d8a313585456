"""The SURVEY.md 8d synthetic inputs generated with torch (scripts/spec_inputs.py) are the same numbers as the numpy hash the parity
tests use, and the K1 / K2 fields have the stated structure (CPU only)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
import spec_inputs as si  # noqa: E402
from util import hash_uniform  # noqa: E402


def test_torch_hash_matches_numpy_hash():
    for seed in (1, 2, 12345):
        a = si.hash_uniform(seed, (7, 5, 3)).numpy()
        b = hash_uniform(seed, (7, 5, 3))
        assert np.array_equal(a, b)
        assert -1.0 <= a.min() and a.max() < 1.0


def test_k1_k2_fields():
    k1 = si.k1_gsrb(16)
    rho = k1["rho"].numpy()
    assert abs(rho.mean() - 1.0) < 1e-14 and rho.min() > 0.5 - 1e-12 and rho.max() < 1.5 + 1e-12
    bx = k1["beta"][0].numpy()
    i, j, k = 3, 4, 5
    assert abs(bx[k, j, i] - 0.5 * k1["dt"] / (0.5 * (rho[k, j, i] + rho[k, j, i - 1]))) < 1e-16
    k2 = si.k2_advection(16)
    u, um = k2["vel"][0].numpy(), k2["umac"].numpy()
    # discretely divergence-free MAC field up to truncation: TG is solenoidal
    div = (np.roll(um, -1, 2) - um) + (np.roll(k2["vmac"].numpy(), -1, 1) - k2["vmac"].numpy()) + (np.roll(k2["wmac"].numpy(), -1, 0) - k2["wmac"].numpy())
    assert np.abs(div).max() < 1e-12
    assert abs(np.abs(u).max() - np.abs(um).max()) < 0.05


def test_ghosted_matches_test_helper():
    import iamr_b200 as ix
    from util import to_fab
    n = (6, 5, 4)
    dense = hash_uniform(3, (2, n[2], n[1], n[0]))
    import torch
    for ixtype, ext in ((ix.CELL, (0, 0, 0)), (ix.XFACE, (1, 0, 0)), (ix.NODE, (1, 1, 1))):
        t, _ = to_fab(dense, ((0, 0, 0), (n[0] - 1, n[1] - 1, n[2] - 1)), 2, ixtype, "cpu")
        g = si.ghosted(torch.from_numpy(dense), 2, ext)
        assert np.array_equal(t.numpy(), g.numpy())
