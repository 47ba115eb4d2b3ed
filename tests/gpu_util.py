"""Helpers shared by the GPU parity tests."""
import torch


def to_tm(x):
    """[B,C,T] channel-major (reference layout) -> ragged time-major [B*T, C] + lens."""
    B, C, T = x.shape
    return x.transpose(1, 2).reshape(B * T, C).contiguous(), [T] * B


def from_tm(y, B):
    """[B*T, C] -> [B,C,T]"""
    return y.view(B, -1, y.shape[-1]).transpose(1, 2).contiguous()


def ragged_tm(xs):
    """list of [C,T_b] -> [sum T_b, C], lens"""
    return torch.cat([x.transpose(0, 1) for x in xs], 0).contiguous(), [x.shape[1] for x in xs]


_ctx = {}


def ctx(precision):
    from cosyvoice_b200 import cvk
    if precision not in _ctx:
        _ctx[precision] = cvk.Context(0, precision, workspace_gb=6.0)
    return _ctx[precision]


def maxdiff(a, b):
    return (a.detach().float().cpu() - b.detach().float().cpu()).abs().max().item()
