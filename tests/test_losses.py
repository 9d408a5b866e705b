"""CPU tests of occuseg_b200.losses: the segment reductions against direct loops (the definitions torch_scatter documents)
and the vectorised per-instance regression losses against a line-by-line restatement of the reference's double loop
(examples/ScanNet/train_instance.py:198-243)."""
import torch

from occuseg_b200 import losses


def _data(seed=0, B=3, P=600, n_inst=7):
    g = torch.Generator().manual_seed(seed)
    sample = torch.sort(torch.randint(0, B, (P,), generator=g))[0]
    inst = torch.randint(0, n_inst, (P,), generator=g)
    # make instance ids dense per sample like np.unique(..., return_inverse) does (scannet.py:172)
    for b in range(B):
        m = sample == b
        inst[m] = torch.unique(inst[m], return_inverse=True)[1]
    sem = torch.randint(0, 5, (P,), generator=g)
    d, dg = torch.randn(P, 3, generator=g), torch.randn(P, 3, generator=g)
    o, og = torch.randn(P, 1, generator=g), torch.randn(P, 1, generator=g)
    return sample, inst, sem, d, dg, o, og, B


def test_segment_reductions_match_loops():
    sample, inst, sem, d, dg, o, og, B = _data()
    n = int(inst.max()) + 1
    s, m, sd = losses.segment_sum(d, inst, n), losses.segment_mean(d, inst, n), losses.segment_std(o, inst, n)
    mx, arg = losses.segment_max(o[:, 0], inst, n)
    for k in range(n):
        sel = inst == k
        assert torch.allclose(s[k], d[sel].sum(0), atol=1e-5) and torch.allclose(m[k], d[sel].mean(0), atol=1e-6)
        assert torch.allclose(sd[k], o[sel].std(0, unbiased=True), atol=1e-6)
        assert mx[k] == o[sel, 0].max() and o[arg[k], 0] == mx[k] and inst[arg[k]] == k
    # empty segment and single-element segment
    ids = torch.tensor([0, 0, 2])
    v = torch.tensor([1.0, 3.0, 5.0])
    assert losses.segment_mean(v, ids, 4).tolist() == [2.0, 0.0, 5.0, 0.0]
    assert torch.allclose(losses.segment_std(v, ids, 4), torch.tensor([2.0 ** 0.5, 0.0, 0.0, 0.0]))
    mx, arg = losses.segment_max(v, ids, 4)
    assert mx.tolist() == [3.0, 0.0, 5.0, 0.0] and arg.tolist() == [1, -1, 2, -1]


def _reference_loop(sample, inst, sem, d, dg, o, og, B):
    """train_instance.py:198-243 with scatter_mean / scatter_std written out"""
    D = torch.zeros(1)
    O = torch.zeros(1)
    for b in range(B):
        idx = sample == b
        im = inst[idx]
        ps = sem[idx]
        de = (d[idx] - dg[idx]).norm(dim=1)
        oe = (o[idx] - og[idx]).norm(dim=1)
        derr, oerr, n = torch.zeros(1), torch.zeros(1), 0
        for mid in range(int(im.max()) + 1):
            sel = im == mid
            if ps[sel][0] > 1:
                derr += de[sel].mean()
                std = o[idx][sel].std(0, unbiased=True) if sel.sum() > 1 else torch.zeros(1)
                oerr += oe[sel].mean() + std.sum()
                n += 1
        if n > 0:
            D += derr / n
            O += oerr / n
    return D / B, O / B


def test_cluster_regression_losses_match_the_reference_loop():
    for seed in range(3):
        sample, inst, sem, d, dg, o, og, B = _data(seed)
        d.requires_grad_(True)
        got_d, got_o = losses.cluster_regression_losses(d, dg, o, og, inst, sample, sem, B)
        want_d, want_o = _reference_loop(sample, inst, sem, d.detach(), dg, o, og, B)
        assert torch.allclose(got_d, want_d[0], atol=1e-5) and torch.allclose(got_o, want_o[0], atol=1e-5)
        got_d.backward()
        assert torch.isfinite(d.grad).all() and d.grad.abs().sum() > 0
        d.requires_grad_(False)
