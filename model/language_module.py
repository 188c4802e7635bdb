"""Query encoder (reference model/language_module.py:9-62, model/ops.py:16-25,74-85).

0.6 % of the path's FLOPs, latency-bound.  The module keeps the reference's parameter
names (including the never-called `textualAttention`, 6.3 M parameters that live in every checkpoint).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

# The parity contract of the path is fp32 (1e-3 on outputs AND gradients).  cuDNN's RNN otherwise runs in TF32, whose
# 1e-4 error in the three query commands is amplified ~100x by the train-mode BatchNorm backward of the dense path.
torch.backends.cudnn.allow_tf32 = False


class XavierLinear(nn.Linear):
    """ops.Linear of the reference (ops.py:16-25): uniform(+-sqrt(3 / fan_avg)), zero bias."""

    def __init__(self, *a, **k):
        super().__init__(*a, **k)
        bound = np.sqrt(3.0 / ((self.in_features + self.out_features) / 2.0))
        nn.init.uniform_(self.weight, -bound, bound)
        if self.bias is not None:
            nn.init.constant_(self.bias, 0.0)


class TextualAttention(nn.Module):
    """Parameter holder only: instantiated by the reference (language_module.py:17) and never called."""

    def __init__(self, hidden_dim=1024):
        super().__init__()
        self.W1 = nn.Linear(hidden_dim, 1)
        self.W2 = nn.Linear(hidden_dim * 2, hidden_dim)
        self.W3 = nn.Linear(hidden_dim * 2, hidden_dim * 2)


class QueryEncoder(nn.Module):
    def __init__(self, vocab_size, hidden_dim=512, embed_dim=300, num_layers=1, bidirection=True):
        super().__init__()
        self.hidden_dim, self.embed_dim = hidden_dim, embed_dim
        self.embedding = nn.Embedding(vocab_size + 1, embed_dim, padding_idx=0)
        self.biLSTM = nn.LSTM(embed_dim, hidden_dim, num_layers, dropout=0.0, batch_first=True, bidirectional=bidirection)
        self.textualAttention = TextualAttention()
        self.qInput = XavierLinear(hidden_dim * 4, hidden_dim)
        for t in range(3):
            setattr(self, "qInput%d" % t, XavierLinear(hidden_dim, hidden_dim * 2))
        self.cmd_inter2logits = XavierLinear(hidden_dim * 2, 1)

    def forward(self, query_tokens, query_length):
        """tokens [B,L] int64 (device), lengths [B] int64 sorted descending -> 3 x [B, 2*hidden]."""
        lengths_cpu = query_length.detach().to("cpu")
        emb = self.embedding(query_tokens)
        packed = pack_padded_sequence(emb, lengths_cpu, batch_first=True)
        self.biLSTM.flatten_parameters()
        out, _ = self.biLSTM(packed)
        out, _ = pad_packed_sequence(out, batch_first=True)  # [B, Lmax, 2H], zero after each length
        B, Lm = out.shape[0], out.shape[1]
        lengths = query_length.to(out.device)
        last = out[torch.arange(B, device=out.device), lengths - 1]
        q_vector = torch.cat((out[:, 0], last), dim=-1)
        hid = F.relu(self.qInput(q_vector))
        mask = torch.arange(Lm, device=out.device)[None, :] >= lengths[:, None]
        cmds = []
        for t in range(3):
            q_cmd = getattr(self, "qInput%d" % t)(hid)
            raw = self.cmd_inter2logits(q_cmd[:, None, :] * out).squeeze(-1)
            att = F.softmax(raw.masked_fill(mask, -1e30), dim=-1)
            cmds.append(torch.bmm(att[:, None, :], out).squeeze(1))
        return cmds
