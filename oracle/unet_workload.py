"""TEST / BENCH INFRASTRUCTURE ONLY -- the UNet-m hot-path workload replayed layer by layer through the
reference's own CPU arithmetic (oracle/_ref, kind "reference") or, if that cannot be loaded, the numpy
port (oracle/arith.py, kind "port").  Used by bench.py's cpu_baseline leg and by `bench.py --impl reference`.

Every sparse layer the UNet instantiates (sparseconvnet/networkArchitectures.py:202-306, reps=1, residual
blocks) is run forward + backward once on tensors of the right shape, with the rulebooks of a real
synthetic scene: SubmanifoldConvolution (CPU/Convolution.cpp:114-189), Convolution 2/2 (:35-112),
Deconvolution 2/2 (CPU/Deconvolution.cpp:7-88), BatchNorm+ReLU (CPU/BatchNormalization.cpp:109-157) and the
1x1 NetworkInNetwork shortcut (CPU/NetworkInNetwork.cpp).  Element-wise adds / concatenations and the
optimizer are not replayed (negligible next to the convolutions).  Rulebooks come from oracle/rulebook.py
and are built outside the timed region (the reference's builder needs CUDA + cudpp)."""
from __future__ import annotations

import time

import numpy as np

from . import arith, rulebook as rb


def unet_layers(m, levels, in_ch=3):
    """[(kind, level, c_in, c_out)] in forward order.  kind in {subm, conv, deconv, bn, nin}."""
    planes = [m * (i + 1) for i in range(levels)]
    out = [("subm", 0, in_ch, m)]

    def block(l, a, b):
        if a != b:
            out.append(("nin", l, a, b))
        out.extend([("bn", l, a, a), ("subm", l, a, b), ("bn", l, b, b), ("subm", l, b, b)])

    def level(l):
        c = planes[l]
        block(l, c, c)
        if l + 1 < levels:
            out.extend([("bn", l, c, c), ("conv", l, c, planes[l + 1])])
            level(l + 1)
            out.extend([("bn", l + 1, planes[l + 1], planes[l + 1]), ("deconv", l, planes[l + 1], c)])
            block(l, 2 * c, c)

    level(0)
    out.append(("bn", 0, m, m))
    return out


class Workload:
    def __init__(self, coords, batch, m=64, levels=6, seed=0):
        self.m, self.levels = m, levels
        vox = rb.voxelize(coords, batch)
        locs = vox["locs"]
        self.n, self.subm, self.strided = [], [], []
        for l in range(levels):
            self.n.append(len(locs))
            self.subm.append(rb.submanifold_rules(locs, batch))
            if l + 1 < levels:
                locs, s = rb.strided_rules(locs, batch)
                self.strided.append(s)
        self.layers = unet_layers(m, levels)
        self.rng = np.random.default_rng(seed)
        self.kind = "port"
        self.ref = None
        try:
            from . import reference
            if reference.available():
                self.kind = "reference"
                self.reference = reference
                self.ref = reference.Ref()
                for l in range(levels):
                    self.ref.load_submanifold(4096 >> l, self.subm[l], self.n[l])
                    if l + 1 < levels:
                        self.ref.load_strided(4096 >> l, 4096 >> (l + 1), self.strided[l], self.n[l], self.n[l + 1])
        except Exception:
            self.kind = "port"
        # inputs and weights are created once, outside the timed region
        self.data = []
        for kind, l, a, b in self.layers:
            n_in = self.n[l + 1] if kind == "deconv" else self.n[l]
            n_out = self.n[l + 1] if kind == "conv" else self.n[l]
            x = self.rng.standard_normal((n_in, a), dtype=np.float32)
            g = self.rng.standard_normal((n_out, b), dtype=np.float32)
            v = {"subm": 27, "conv": 8, "deconv": 8}.get(kind, 0)
            w = (self.rng.standard_normal((v, a, b), dtype=np.float32) * 0.05) if v else \
                (self.rng.standard_normal((a, b), dtype=np.float32) * 0.05)
            self.data.append((x, g, w))

    @property
    def voxels(self):
        return self.n[0]

    def set_threads(self, n):
        if self.kind == "reference":
            self.reference.module().set_threads(int(n))
        import torch
        torch.set_num_threads(int(n))

    def step(self):
        """one forward+backward pass over every layer; returns seconds"""
        t0 = time.perf_counter()
        for (kind, l, a, b), (x, g, w) in zip(self.layers, self.data):
            size, csize = 4096 >> l, 4096 >> (l + 1)
            if self.kind == "reference":
                R = self.ref
                if kind == "subm":
                    R.subm_forward(size, x, w); R.subm_backward(size, x, g, w)
                elif kind == "conv":
                    R.conv_forward(size, csize, x, w); R.conv_backward(size, csize, x, g, w)
                elif kind == "deconv":
                    R.deconv_forward(csize, size, x, w); R.deconv_backward(csize, size, x, g, w)
                elif kind == "bn":
                    ones, zeros = np.ones(a, np.float32), np.zeros(a, np.float32)
                    y, sm, si, _, _ = self.reference.bn_forward(x, ones, zeros, zeros, ones)
                    self.reference.bn_backward(x, y, g, ones, zeros, sm, si)
                else:
                    import torch
                    xt, gt, wt = torch.from_numpy(x), torch.from_numpy(g), torch.from_numpy(w)
                    out, din, dw = torch.empty(0), torch.empty(0), torch.zeros_like(wt)
                    mod = self.reference.module()
                    mod.NetworkInNetwork_updateOutput(xt, out, wt, torch.empty(0))
                    mod.NetworkInNetwork_updateGradInput(din, gt, wt)
                    mod.NetworkInNetwork_accGradParameters(xt, gt, dw, torch.empty(0))
            else:
                if kind == "subm":
                    arith.rule_conv_forward(x, w, self.subm[l], len(g)); arith.rule_conv_backward(x, g, w, self.subm[l])
                elif kind == "conv":
                    arith.rule_conv_forward(x, w, self.strided[l], len(g)); arith.rule_conv_backward(x, g, w, self.strided[l])
                elif kind == "deconv":
                    arith.rule_conv_forward(x, w, self.strided[l], len(g), 1, 0)
                    arith.rule_conv_backward(x, g, w, self.strided[l], 1, 0)
                elif kind == "bn":
                    ones, zeros = np.ones(a, np.float32), np.zeros(a, np.float32)
                    y, sm, si, _, _ = arith.batchnorm_forward(x, ones, zeros, zeros, ones)
                    arith.batchnorm_backward(x, y, g, ones, sm, si)
                else:
                    _ = x @ w; _ = g @ w.T; _ = x.T @ g
        return time.perf_counter() - t0
