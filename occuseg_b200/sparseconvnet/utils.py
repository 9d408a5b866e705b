"""Small helpers with the reference's names (sparseconvnet/utils.py:13-28)."""
import torch


def toLongTensor(dimension, x):
    if isinstance(x, torch.Tensor) and x.dtype == torch.int64 and not x.is_cuda:
        return x
    if isinstance(x, (list, tuple)):
        assert len(x) == dimension
        return torch.LongTensor(list(x))
    return torch.LongTensor(dimension).fill_(int(x))


def optionalTensor(obj, name):
    return getattr(obj, name) if hasattr(obj, name) else torch.Tensor()


def optionalTensorReturn(t):
    return t if t.numel() else None
