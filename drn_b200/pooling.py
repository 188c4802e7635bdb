"""Proposal feature pooling on the GPU (reference dataset.py:105-155 `CharadesSTA.get_data` + the padding of `collate_data`,
dataset.py:180-206): turns the per-video window features of a batch into the `props_features` / `props_start_end` tensors
`mainModel.forward` consumes, without a CPU DataLoader round trip.  Thin marshalling over `drn_pool_proposals`."""
import ctypes as C

import torch

from . import lib as L


def pool_proposals(vid_feats, p_starts, p_ends, num_frames, window, overlap):
    """vid_feats: list of CUDA fp32 tensors [n_win_b, D]; p_starts[b]: sequence of float start frames; p_ends[b]: sequence of
    int end frames (already min(int(end), num_frames), dataset.py:16); num_frames[b]: int.  window / overlap: the feature
    type's `ft_window_size` / `ft_overlap` (default_config.yaml).  Returns (props_features [B, Pmax, D] fp32, props_start_end
    [B, Pmax, 2] f64) on the device, zero padded like collate_data."""
    dev = vid_feats[0].device
    if dev.type != "cuda":
        raise RuntimeError("pool_proposals runs on a B200 through libdrn_sm100.so only (no CPU fallback)")
    B, D = len(vid_feats), vid_feats[0].shape[1]
    P = max(len(s) for s in p_starts)
    interval = int(window * (1 - overlap))  # dataset.py:125
    feats = torch.cat([f.contiguous() for f in vid_feats]) if B > 1 else vid_feats[0].contiguous()
    off = [0]
    for f in vid_feats:
        off.append(off[-1] + f.shape[0])
    ps = torch.zeros(B, P, dtype=torch.float64)
    pe = torch.zeros(B, P, dtype=torch.int32)
    for b in range(B):
        n = len(p_starts[b])
        ps[b, :n] = torch.as_tensor(p_starts[b], dtype=torch.float64)
        pe[b, :n] = torch.as_tensor(p_ends[b], dtype=torch.int32)
    win_off = torch.tensor(off, dtype=torch.int64, device=dev)
    nprops = torch.tensor([len(s) for s in p_starts], dtype=torch.int32, device=dev)
    nfr = torch.tensor([int(x) for x in num_frames], dtype=torch.int32, device=dev)
    ps, pe = ps.to(dev), pe.to(dev)
    out = torch.empty(B, P, D, device=dev)
    pse = torch.empty(B, P, 2, dtype=torch.float64, device=dev)
    L.check(L.load().drn_pool_proposals(L.ptr(feats), L.ptr(win_off), L.ptr(ps), L.ptr(pe), L.ptr(nprops), L.ptr(nfr), B, P, D,
                                        int(window), interval, L.ptr(out), L.ptr(pse), L.stream_ptr()), "pool_proposals")
    return out, pse
