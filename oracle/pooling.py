"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's proposal feature pooling + padding
(dataset.py:105-155 `CharadesSTA.get_data`, dataset.py:180-206 `collate_data`), the step immediately before the hot path
(SURVEY.md section 8f-2).  numpy, integer / max arithmetic only: the CUDA kernel (drn_pool_proposals) must match bit for bit.

Pinned by tests/golden/pool_props.npz, produced by the UNMODIFIED reference dataset code on synthetic feature files
(oracle/make_pool_goldens.py)."""
import numpy as np


def window_range(p_start, p_end, n_win, window, interval):
    """dataset.py:120-146: inclusive range [lo, hi] of feature-window rows a proposal max-pools over.
    p_start: float start frame; p_end: int end frame (already min(int(end), num_frames), dataset.py:16); n_win = len(vid_feature)."""
    ft_start_index = (int(p_start) // interval) * interval          # dataset.py:126
    lo = ft_start_index // interval
    if p_end - p_start <= window:                                   # dataset.py:128-131
        hi = lo
    else:                                                           # dataset.py:133-139: range(ft_start_index, p_end, interval)
        count = len(range(ft_start_index, int(p_end), interval))
        hi = lo + count - 1
    return min(n_win - 1, lo), min(n_win - 1, hi)                   # dataset.py:145


def pool_and_pad(vid_feats, p_starts, p_ends, num_frames, window, interval):
    """vid_feats: list of [n_win_b, D] float32; p_starts[b]: float array [P_b]; p_ends[b]: int array [P_b];
    num_frames[b]: int.  Returns props_features [B, Pmax, D] float32 (zero padded, dataset.py:188,203) and
    props_s_e [B, Pmax, 2] float64 = (start / num_frames, end / num_frames) (dataset.py:124,189,205)."""
    B = len(vid_feats)
    D = vid_feats[0].shape[1]
    Pmax = max(len(s) for s in p_starts)
    out = np.zeros((B, Pmax, D), dtype=np.float32)
    pse = np.zeros((B, Pmax, 2), dtype=np.float64)
    for b in range(B):
        f = vid_feats[b]
        for i, (s, e) in enumerate(zip(p_starts[b], p_ends[b])):
            lo, hi = window_range(float(s), int(e), f.shape[0], window, interval)
            out[b, i] = f[lo:hi + 1].max(axis=0)                    # dataset.py:151
            pse[b, i] = (float(s) / num_frames[b], int(e) / num_frames[b])
    return out, pse
