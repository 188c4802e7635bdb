"""drn_pool_proposals (proposal feature pooling + padding, reference dataset.py:105-155,180-206) through the C ABI: BIT-EXACT
against the reference-generated fixture (tests/golden/pool_props.npz) and against the numpy oracle on random ragged batches at
the real feature width (D = 4096): ragged proposal counts, short feature files (clamp), single-window proposals."""
import os

import numpy as np
import pytest
import torch

from drn_b200.pooling import pool_proposals
from oracle import pooling as P

pytestmark = pytest.mark.gpu


def test_matches_reference_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "pool_props.npz"))
    B = len(g["props_num"])
    feats = [torch.from_numpy(g["feat_%d" % k]).cuda() for k in range(B)]
    ps = [g["p_start_%d" % k] for k in range(B)]
    pe = [g["p_end_%d" % k] for k in range(B)]
    out, pse = pool_proposals(feats, ps, pe, g["num_frames"].tolist(), window=16, overlap=0.5)
    assert np.array_equal(out.cpu().numpy(), g["props_features"])
    assert np.array_equal(pse.cpu().numpy(), g["props_s_e"])


@pytest.mark.parametrize("seed", [0, 1])
def test_random_ragged_batches_bit_exact(seed):
    rng = np.random.default_rng(seed)
    B, D = 9, 4096
    feats, ps, pe, nfr = [], [], [], []
    for b in range(B):
        num_frames = int(rng.integers(40, 900))
        full = max(1, (num_frames - 16) // 8 + 1)
        n_win = full if b % 3 else max(1, full // 3)       # some files are much shorter than the video (clamp)
        feats.append(rng.standard_normal((n_win, D)).astype(np.float32))
        n = int(rng.integers(1, 33))                       # ragged proposal counts
        s = np.sort(rng.uniform(0, num_frames - 1, size=n))
        e = np.minimum((s + rng.uniform(1, num_frames, size=n)).astype(np.int64), num_frames)
        e = np.maximum(e, s.astype(np.int64) + 1)
        if b == 0:
            s[0], e[0] = 3.0, 10                           # duration <= window: single feature window
        ps.append(s)
        pe.append(e)
        nfr.append(num_frames)
    ref, ref_pse = P.pool_and_pad(feats, ps, pe, nfr, window=16, interval=8)
    out, pse = pool_proposals([torch.from_numpy(f).cuda() for f in feats], ps, pe, nfr, window=16, overlap=0.5)
    assert np.array_equal(out.cpu().numpy(), ref)
    assert np.array_equal(pse.cpu().numpy(), ref_pse)


def test_rejects_bad_arguments():
    import ctypes as C
    from drn_b200 import lib as L
    assert L.load().drn_pool_proposals(None, None, None, None, None, None, 1, 1, 6, 16, 8, None, None, None) == -1
    assert L.load().drn_pool_proposals(None, None, None, None, None, None, 1, 1, 8, 16, 0, None, None, None) == -1
