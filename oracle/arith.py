"""TEST INFRASTRUCTURE ONLY -- numpy port of the reference's CPU arithmetic (kind "port").

Follows /root/reference/sparseconvnet/SCN/CPU/Convolution.cpp:8-33,114-189 (gather rows by rule,
at::mm, scatter-add by rule), CPU/Deconvolution.cpp:7-88 (same with the rule columns swapped) and
CPU/BatchNormalization.cpp:12-107.  It is validated against the reference's own compiled code
(oracle/_ref, see tests/test_oracle.py) and exists so that the checker still works on a machine
where oracle/_ref cannot be loaded.  All accumulation is fp32 like the reference.
"""
from __future__ import annotations

import numpy as np


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def rule_conv_forward(inp, weight, rules, n_out, in_col=0, out_col=1):
    """Convolution.cpp:36-72 / :114-145.  rules[k]: int32 [n,2]; returns (out, macs)."""
    inp, weight = _f32(inp), _f32(weight)
    out = np.zeros((n_out, weight.shape[2]), np.float32)
    macs = 0.0
    for k, r in enumerate(rules):
        if len(r):
            rows = inp[r[:, in_col]] @ weight[k]
            np.add.at(out, r[:, out_col], rows.astype(np.float32))
            macs += float(len(r)) * weight.shape[1] * weight.shape[2]
    return out, macs


def rule_conv_backward(inp, d_out, weight, rules, in_col=0, out_col=1):
    """Convolution.cpp:74-112 / :147-189.  Returns (d_inp, d_weight)."""
    inp, d_out, weight = _f32(inp), _f32(d_out), _f32(weight)
    d_inp = np.zeros_like(inp)
    d_w = np.zeros_like(weight)
    for k, r in enumerate(rules):
        if len(r):
            a = inp[r[:, in_col]]
            g = d_out[r[:, out_col]]
            d_w[k] = a.T @ g
            np.add.at(d_inp, r[:, in_col], (g @ weight[k].T).astype(np.float32))
    return d_inp, d_w


def batchnorm_forward(x, gamma, beta, running_mean, running_var, eps=1e-4, momentum=0.9, train=True, leakiness=0.0):
    """BatchNormalization.cpp:12-61.  Returns (y, save_mean, save_invstd, new_running_mean, new_running_var).
    Sums are taken in float64 and rounded once (the reference's serial fp32 running sum is itself
    only accurate to ~sqrt(N)*eps; see DESIGN.md, tolerance section)."""
    x = _f32(x)
    n = x.shape[0]
    if train:
        mean = x.astype(np.float64).mean(0)
        var_sum = ((x.astype(np.float64) - mean) ** 2).sum(0)
        rm = momentum * running_mean + (1 - momentum) * mean
        rv = momentum * running_var + (1 - momentum) * var_sum / max(n - 1, 1)
        invstd = (var_sum / n + eps) ** -0.5
    else:
        mean, rm, rv = running_mean.astype(np.float64), running_mean, running_var
        invstd = (running_var.astype(np.float64) + eps) ** -0.5
    w = invstd * (gamma if gamma is not None else 1.0)
    b = -mean * w + (beta if beta is not None else 0.0)
    y = x * w.astype(np.float32) + b.astype(np.float32)
    y = np.where(y > 0, y, y * np.float32(leakiness)).astype(np.float32)
    return y, mean.astype(np.float32), invstd.astype(np.float32), np.asarray(rm, np.float32), np.asarray(rv, np.float32)


def batchnorm_backward(x, y, d_y, gamma, save_mean, save_invstd, leakiness=0.0):
    """BatchNormalization.cpp:63-107.  Returns (d_x, d_gamma, d_beta)."""
    x, y, d_y = _f32(x), _f32(y), _f32(d_y)
    n = x.shape[0]
    d = np.where(y > 0, d_y, d_y * np.float32(leakiness)).astype(np.float64)
    xc = x.astype(np.float64) - save_mean
    grad_sum = d.sum(0)
    dotp = (xc * d).sum(0)
    k = dotp * save_invstd.astype(np.float64) ** 2 / n
    g = gamma.astype(np.float64) if gamma is not None else 1.0
    d_x = (d - grad_sum / n - xc * k) * save_invstd * g
    return d_x.astype(np.float32), (dotp * save_invstd).astype(np.float32), grad_sum.astype(np.float32)
