"""Query encoder parameter container (reference model/language_module.py:9-62, model/ops.py:16-25).

Holds the parameters under the reference's names (including the never-called `textualAttention`, 6.3 M parameters
that live in every checkpoint) with the reference's initialisers.  The math -- embedding, packed BiLSTM, the three
attention "commands" (language_module.py:27-62), forward and backward -- runs in drn_qe_forward / drn_qe_backward
(drn_b200/csrc/query.cu) as part of the path's CUDA-graph; there is no torch forward.
"""
import numpy as np
import torch.nn as nn


def _no_forward(self, *a, **k):
    raise RuntimeError("parameter container: the query encoder runs in libdrn_sm100 (drn_qe_forward), "
                       "there is no torch forward / CPU fallback")


class XavierLinear(nn.Linear):
    """ops.Linear of the reference (ops.py:16-25): uniform(+-sqrt(3 / fan_avg)), zero bias."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        bound = np.sqrt(3.0 / ((self.in_features + self.out_features) / 2.0))
        nn.init.uniform_(self.weight, -bound, bound)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0.0)
    forward = _no_forward


class TextualAttention(nn.Module):
    """Parameter holder only: instantiated by the reference (language_module.py:17) and never called."""

    def __init__(self, hidden_dim=1024):
        super().__init__()
        self.W1 = nn.Linear(hidden_dim, 1)
        self.W2 = nn.Linear(hidden_dim * 2, hidden_dim)
        self.W3 = nn.Linear(hidden_dim * 2, hidden_dim * 2)
    forward = _no_forward


class QueryEncoder(nn.Module):
    def __init__(self, vocab_size, hidden_dim=512, embed_dim=300, num_layers=1, bidirection=True):
        super().__init__()
        if num_layers != 1 or not bidirection:
            raise NotImplementedError("the B200 query-encoder kernels implement the reference configuration: one "
                                      "bidirectional LSTM layer (default_config.yaml lstm_layers: 1, main.py:89-90)")
        self.hidden_dim, self.embed_dim = hidden_dim, embed_dim
        self.embedding = nn.Embedding(vocab_size + 1, embed_dim, padding_idx=0)
        self.biLSTM = nn.LSTM(embed_dim, hidden_dim, num_layers, dropout=0.0, batch_first=True, bidirectional=bidirection)
        self.textualAttention = TextualAttention()
        self.qInput = XavierLinear(hidden_dim * 4, hidden_dim)
        for t in range(3):
            setattr(self, "qInput%d" % t, XavierLinear(hidden_dim, hidden_dim * 2))
        self.cmd_inter2logits = XavierLinear(hidden_dim * 2, 1)
    forward = _no_forward
