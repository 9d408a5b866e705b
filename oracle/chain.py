"""TEST INFRASTRUCTURE ONLY -- replays a whole sparse network, layer by layer, through the reference's CPU
arithmetic (oracle/ref_scn.py -> oracle/_ref) so that network-level parity tests compare the CUDA path with the
ORACLE instead of with itself.

`replay(net, x)` walks the module tree of a network built from `occuseg_b200.sparseconvnet` layers (or the reference's
own -- the walk goes by class NAME and the reference attribute names: nIn/nOut/filter_size/weight/bias/eps/momentum/
leakiness/running_mean/...) on CPU copies of its parameters and evaluates
    InputLayer, SubmanifoldConvolution, Convolution, Deconvolution, BatchNormalization(+ReLU/LeakyReLU),
    NetworkInNetwork, OutputLayer, Sequential, ConcatTable, AddTable, JoinTable, Identity
with torch CPU autograd around the stand-in SCN entry points -- the same division of labour as the reference's
Function classes (submanifoldConvolution.py:76-128, convolution.py:72-127, deconvolution.py:87-155,
batchNormalization.py:90-161, networkInNetwork.py:14-59, ioLayers.py:157-223).  That this walker computes what the
reference's own Python package computes is itself tested (tests/test_chain_oracle.py, authoring container only).

Every leaf records a tape entry {name, kind, module, x, y, gy, gx, grads...} so a test can also feed ONE CUDA layer
the oracle's input of that layer and compare its output (per-layer tolerance for the tensor-core precisions)."""
from __future__ import annotations

import torch
from torch.autograd import Function

from . import ref_scn as R


def _lt(v, like=None):
    return torch.LongTensor([int(v)] * 3)


class _Subm(Function):
    @staticmethod
    def forward(ctx, x, w, b, meta, size, rec):
        ctx.meta, ctx.size, ctx.rec = meta, size, rec
        ctx.save_for_backward(x, w, b)
        out = torch.empty(0)
        rec["macs"] = R.SubmanifoldConvolution_updateOutput(_lt(size), _lt(3), meta, x, out, w, b, 1)
        rec["x"], rec["y"] = x.detach(), out.detach()
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, b = ctx.saved_tensors
        gx, gw, gb = torch.empty(0), torch.zeros_like(w), torch.zeros_like(b)
        R.SubmanifoldConvolution_backward(_lt(ctx.size), _lt(3), ctx.meta, x, gx, g.contiguous(), w, gw, gb, 1)
        ctx.rec.update(gy=g.detach().clone(), gx=gx, gw=gw)
        return gx, gw, (gb if gb.numel() else None), None, None, None


class _Strided(Function):
    @staticmethod
    def forward(ctx, x, w, b, meta, in_size, out_size, deconv, rec):
        ctx.meta, ctx.sizes, ctx.deconv, ctx.rec = meta, (in_size, out_size), deconv, rec
        ctx.save_for_backward(x, w, b)
        out = torch.empty(0)
        f = R.Deconvolution_updateOutput if deconv else R.Convolution_updateOutput
        rec["macs"] = f(_lt(in_size), _lt(out_size), _lt(2), _lt(2), meta, x, out, w, b)
        rec["x"], rec["y"] = x.detach(), out.detach()
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, b = ctx.saved_tensors
        gx, gw, gb = torch.empty(0), torch.zeros_like(w), torch.zeros_like(b)
        f = R.Deconvolution_backward if ctx.deconv else R.Convolution_backward
        f(_lt(ctx.sizes[0]), _lt(ctx.sizes[1]), _lt(2), _lt(2), ctx.meta, x, gx, g.contiguous(), w, gw, gb)
        ctx.rec.update(gy=g.detach().clone(), gx=gx, gw=gw)
        return gx, gw, (gb if gb.numel() else None), None, None, None, None, None


class _BN(Function):
    @staticmethod
    def forward(ctx, x, w, b, rm, rv, eps, momentum, train, leak, rec):
        c = rm.numel()
        out, sm, si = torch.empty(0), torch.empty(c), torch.empty(c)
        R.BatchNormalization_updateOutput(x, out, sm, si, rm, rv, w, b, eps, momentum, train, leak)
        ctx.save_for_backward(x, out, w, b, rm, rv, sm, si)
        ctx.leak, ctx.rec = leak, rec
        rec.update(x=x.detach(), y=out.detach(), save_mean=sm, save_invstd=si)
        return out

    @staticmethod
    def backward(ctx, g):
        x, out, w, b, rm, rv, sm, si = ctx.saved_tensors
        gx, gw, gb = torch.empty(0), torch.zeros_like(w), torch.zeros_like(b)
        R.BatchNormalization_backward(x, gx, out, g.contiguous(), sm, si, rm, rv, w, b, gw, gb, ctx.leak)
        ctx.rec.update(gy=g.detach().clone(), gx=gx, gw=gw, gb=gb)
        return gx, (gw if gw.numel() else None), (gb if gb.numel() else None), None, None, None, None, None, None, None


class _NiN(Function):
    @staticmethod
    def forward(ctx, x, w, b, rec):
        out = torch.empty(0)
        ctx.save_for_backward(x, w, b)
        ctx.rec = rec
        R.NetworkInNetwork_updateOutput(x, out, w, b)
        rec["x"], rec["y"] = x.detach(), out.detach()
        return out

    @staticmethod
    def backward(ctx, g):
        x, w, b = ctx.saved_tensors
        g = g.contiguous()
        gx, gw, gb = torch.empty(0), torch.zeros_like(w), torch.zeros_like(b)
        R.NetworkInNetwork_updateGradInput(gx, g, w)
        R.NetworkInNetwork_accGradParameters(x, g, gw, gb)
        ctx.rec.update(gy=g.detach().clone(), gx=gx, gw=gw)
        return gx, gw, (gb if gb.numel() else None), None


class _Input(Function):
    @staticmethod
    def forward(ctx, feats, coords, meta, size, batch, mode, rec):
        ctx.meta = meta
        out = torch.empty(0)
        R.InputLayer_updateOutput(meta, _lt(size), coords, feats, out, batch, mode, None)
        rec["x"], rec["y"] = feats.detach(), out.detach()
        return out

    @staticmethod
    def backward(ctx, g):
        gx = torch.empty(0)
        R.InputLayer_updateGradInput(ctx.meta, gx, g.contiguous())
        return gx, None, None, None, None, None, None


class _Output(Function):
    @staticmethod
    def forward(ctx, x, meta, rec):
        ctx.meta, ctx.rec = meta, rec
        out = torch.empty(0)
        R.OutputLayer_updateOutput(meta, x, out)
        rec["x"], rec["y"] = x.detach(), out.detach()
        return out

    @staticmethod
    def backward(ctx, g):
        gx = torch.empty(0)
        R.OutputLayer_updateGradInput(ctx.meta, gx, g.contiguous())
        ctx.rec.update(gy=g.detach().clone(), gx=gx)
        return gx, None, None


class _T:
    """features + metadata + spatial size (the reference's SparseConvNetTensor, sparseConvNetTensor.py:13-66)"""

    def __init__(self, features, meta, size):
        self.features, self.meta, self.size = features, meta, size


def _param(mod, name, params):
    t = getattr(mod, name, None)
    if t is None:
        return torch.empty(0)
    key = id(t)
    if key not in params:
        params[key] = t.detach().cpu().float().clone().requires_grad_(isinstance(t, torch.nn.Parameter))
    return params[key]


class Replay:
    def __init__(self, net):
        self.net = net
        self.params = {}          # id(original tensor) -> CPU copy (leaf with grad for Parameters)
        self.tape = []
        self.names = {id(m): n for n, m in net.named_modules()}

    def grad_of(self, p):
        """gradient the oracle computed for the network's parameter p (after backward())"""
        return self.params[id(p)].grad

    def _rec(self, mod, kind):
        rec = {"name": self.names.get(id(mod), "?"), "kind": kind, "module": mod}
        self.tape.append(rec)
        return rec

    def run(self, mod, t):
        cls = type(mod).__name__
        P = lambda n: _param(mod, n, self.params)  # noqa: E731
        if cls in ("Sequential",) or (isinstance(mod, torch.nn.Sequential) and cls not in
                                      ("ConcatTable", "ResidualConcatTable", "AddTable", "JoinTable")):
            for child in mod._modules.values():
                t = self.run(child, t)
            return t
        if cls in ("ConcatTable", "ResidualConcatTable"):
            return [self.run(child, t) for child in mod._modules.values()]
        if cls == "AddTable":
            total = t[0].features
            for u in t[1:]:
                total = total + u.features
            return _T(total, t[0].meta, t[0].size)
        if cls == "JoinTable":
            return _T(torch.cat([u.features for u in t], 1), t[0].meta, t[0].size)
        if cls == "Identity":
            return t
        if cls == "InputLayer":
            coords, feats = t[0], t[1]
            batch = t[3]
            meta = R.Metadata_3()
            size = int(mod.spatial_size[0])
            f = _Input.apply(feats.detach().cpu().float(), coords.cpu().long() if coords.dtype != torch.int64 else coords.cpu(),
                             meta, size, batch, mod.mode, self._rec(mod, "input"))
            return _T(f, meta, size)
        if cls == "OutputLayer":
            return _Output.apply(t.features, t.meta, self._rec(mod, "output"))
        if cls in ("SubmanifoldConvolution", "ValidConvolution"):
            f = _Subm.apply(t.features, P("weight"), P("bias"), t.meta, t.size, self._rec(mod, "subm"))
            return _T(f, t.meta, t.size)
        if cls == "Convolution":
            out_size = t.size // 2
            f = _Strided.apply(t.features, P("weight"), P("bias"), t.meta, t.size, out_size, False, self._rec(mod, "conv"))
            return _T(f, t.meta, out_size)
        if cls == "Deconvolution":
            out_size = t.size * 2
            f = _Strided.apply(t.features, P("weight"), P("bias"), t.meta, t.size, out_size, True, self._rec(mod, "deconv"))
            return _T(f, t.meta, out_size)
        if cls in ("BatchNormalization", "BatchNormReLU", "BatchNormLeakyReLU"):
            f = _BN.apply(t.features, P("weight"), P("bias"), P("running_mean"), P("running_var"), mod.eps, mod.momentum,
                          mod.training, mod.leakiness, self._rec(mod, "bn"))
            return _T(f, t.meta, t.size)
        if cls == "NetworkInNetwork":
            f = _NiN.apply(t.features, P("weight"), P("bias"), self._rec(mod, "nin"))
            return _T(f, t.meta, t.size)
        if cls in ("Linear", "Sigmoid", "Softplus", "ReLU"):        # dense heads: plain torch on CPU copies
            if cls == "Linear":
                return torch.nn.functional.linear(t, P("weight"), P("bias") if mod.bias is not None else None)
            return mod(t)
        raise NotImplementedError(f"oracle/chain.py: no replay rule for {cls}")


def replay(net, x):
    """x = [coords [P,4], feats [P,C], normals-or-None, batch_size] (any device).  Returns (Replay, output tensor on CPU
    with the autograd graph attached: call .backward() on a loss of it, then read Replay.grad_of(param) / Replay.tape)."""
    r = Replay(net)
    out = r.run(net, x)
    # name -> CPU copy (parameters: .grad holds the oracle's gradient after backward(); buffers: updated running
    # statistics).  Keyed by NAME so it survives net.cuda(), which replaces the buffer tensors.
    r.named = {n: r.params[id(t)] for n, t in list(net.named_parameters()) + list(net.named_buffers()) if id(t) in r.params}
    return r, out
