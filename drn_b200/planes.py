"""Split-plane activations / weights: x ~= hi + lo with hi = bf16(x), lo = bf16(x - hi), stored [2][B][T][C] bf16.

This is the operand format of the tensor-core contraction `drn_gemm` (include/drn_b200.h).  The kernels of the path
write planes themselves; `from_float` exists for inputs, weights (until packed by the library) and tests.
"""
import ctypes as C

import torch

from . import lib as L


class Planes:
    """bf16 tensor `data` of shape [2, B, T, C] (contiguous)."""

    __slots__ = ("data",)

    def __init__(self, data):
        assert data.dtype == torch.bfloat16 and data.dim() == 4 and data.shape[0] == 2 and data.is_contiguous()
        self.data = data

    @staticmethod
    def empty(B, T, Cn, device):
        return Planes(torch.empty(2, B, T, Cn, dtype=torch.bfloat16, device=device))

    @staticmethod
    def zeros(B, T, Cn, device):
        return Planes(torch.zeros(2, B, T, Cn, dtype=torch.bfloat16, device=device))

    @staticmethod
    def from_float(x):
        """x fp32 [B, T, C] -> planes (torch ops; test / setup helper, not on the hot path)."""
        x = x.float()
        hi = x.to(torch.bfloat16)
        lo = (x - hi.float()).to(torch.bfloat16)
        return Planes(torch.stack([hi, lo]).contiguous())

    def to_float(self):
        return self.data[0].float() + self.data[1].float()

    @property
    def B(self):
        return self.data.shape[1]

    @property
    def T(self):
        return self.data.shape[2]

    @property
    def C(self):
        return self.data.shape[3]

    @property
    def plane_stride(self):
        return self.data.stride(0)

    def desc(self, parity=1):
        """C descriptor; parity=2 views the time axis as [T/2][2] (stride-2 convs)."""
        assert self.T % parity == 0
        return L.Planes(C.c_void_p(self.data.data_ptr()), self.plane_stride, self.B, self.T // parity, parity, self.C)
