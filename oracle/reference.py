"""TEST INFRASTRUCTURE ONLY -- thin driver around oracle/_ref (the reference's own CPU code, compiled
unmodified by oracle/build_ref.py) that feeds it rulebooks from oracle/rulebook.py.

Reference entry points exercised (all under /root/reference/sparseconvnet/SCN/):
  cpu_SubmanifoldConvolution_updateOutput/backward   CPU/Convolution.cpp:114-189
  cpu_Convolution_updateOutput/backward              CPU/Convolution.cpp:35-112
  cpu_Deconvolution_updateOutput/backward            CPU/Deconvolution.cpp:7-88
  cpu_BatchNormalization_updateOutput/backward       CPU/BatchNormalization.cpp:109-157
"""
from __future__ import annotations

import numpy as np
import torch

from . import build_ref

_mod = None


def module():
    global _mod
    if _mod is None:
        _mod = build_ref.load()
    return _mod


def available():
    return module() is not None


def _lt(v):
    return torch.LongTensor([int(v)] * 3)


def _t(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32))


def _lists(rules):
    return [torch.from_numpy(np.ascontiguousarray(r, dtype=np.int32)) for r in rules]


class Ref:
    """One reference Metadata stub loaded with rulebooks for a chain of scales."""

    def __init__(self):
        self.m = module().Metadata_3()

    def load_submanifold(self, size, rules, n_active):
        self.m.load_submanifold(int(size), _lists(rules))
        self.m.set_nactive(int(size), int(n_active))

    def load_strided(self, fine_size, coarse_size, rules, n_fine, n_coarse):
        self.m.load_strided(int(fine_size), _lists(rules))
        self.m.set_nactive(int(fine_size), int(n_fine))
        self.m.set_nactive(int(coarse_size), int(n_coarse))

    # -- submanifold ---------------------------------------------------------------------------
    def subm_forward(self, size, x, w):
        out = torch.empty(0)
        macs = module().SubmanifoldConvolution_updateOutput(_lt(size), _lt(3), self.m, _t(x), out, _t(w), torch.empty(0))
        return out.numpy(), macs

    def subm_backward(self, size, x, d_out, w):
        din, dw = torch.empty(0), torch.zeros(w.shape, dtype=torch.float32)
        module().SubmanifoldConvolution_backward(_lt(size), _lt(3), self.m, _t(x), din, _t(d_out), _t(w), dw, torch.empty(0))
        return din.numpy(), dw.numpy()

    # -- strided -------------------------------------------------------------------------------
    def conv_forward(self, fine, coarse, x, w):
        out = torch.empty(0)
        macs = module().Convolution_updateOutput(_lt(fine), _lt(coarse), _lt(2), _lt(2), self.m, _t(x), out, _t(w), torch.empty(0))
        return out.numpy(), macs

    def conv_backward(self, fine, coarse, x, d_out, w):
        din, dw = torch.empty(0), torch.zeros(w.shape, dtype=torch.float32)
        module().Convolution_backward(_lt(fine), _lt(coarse), _lt(2), _lt(2), self.m, _t(x), din, _t(d_out), _t(w), dw, torch.empty(0))
        return din.numpy(), dw.numpy()

    def deconv_forward(self, coarse, fine, x, w):
        out = torch.empty(0)
        macs = module().Deconvolution_updateOutput(_lt(coarse), _lt(fine), _lt(2), _lt(2), self.m, _t(x), out, _t(w), torch.empty(0))
        return out.numpy(), macs

    def deconv_backward(self, coarse, fine, x, d_out, w):
        din, dw = torch.empty(0), torch.zeros(w.shape, dtype=torch.float32)
        module().Deconvolution_backward(_lt(coarse), _lt(fine), _lt(2), _lt(2), self.m, _t(x), din, _t(d_out), _t(w), dw, torch.empty(0))
        return din.numpy(), dw.numpy()


def bn_forward(x, gamma, beta, rm, rv, eps=1e-4, momentum=0.9, train=True, leakiness=0.0):
    x = _t(x)
    out, sm, si = torch.empty(0), torch.empty(x.shape[1]), torch.empty(x.shape[1])
    rm, rv = _t(rm).clone(), _t(rv).clone()
    module().BatchNormalization_updateOutput(x, out, sm, si, rm, rv, _t(gamma), _t(beta), eps, momentum, train, leakiness)
    return out.numpy(), sm.numpy(), si.numpy(), rm.numpy(), rv.numpy()


def bn_backward(x, y, d_y, gamma, beta, sm, si, leakiness=0.0):
    x = _t(x)
    din = torch.empty(0)
    dw, db = torch.zeros(x.shape[1]), torch.zeros(x.shape[1])
    d_y = _t(d_y).clone()                       # the reference masks d_output in place
    module().BatchNormalization_backward(x, din, _t(y), d_y, _t(sm), _t(si), torch.empty(0), torch.empty(0),
                                         _t(gamma), _t(beta), dw, db, leakiness)
    return din.numpy(), dw.numpy(), db.numpy()
