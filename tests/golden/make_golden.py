"""Generates tests/golden/*.npz.  Run in the authoring container only (needs /root/reference to build
oracle/_ref):   python tests/golden/make_golden.py

Every floating-point array in the fixture is an OUTPUT OF THE REFERENCE'S OWN CPU CODE
(sparseconvnet/SCN/CPU/*.cpp compiled unmodified, oracle/ref_shim.cpp) on the seeded inputs stored
beside it.  The integer rulebooks are in the GPU builders' convention (row = sorted key rank, taps
x-outermost; oracle/rulebook.py) and are checked here, before they are written, against the reference's own
CPU rule builders compiled from the reference tree (oracle/_ref/scn_rules_ref.so, oracle/rules_ref.py):
same relation {(tap, in xyz, out xyz)}, same point->voxel grouping, same coarse voxel set.  They are what the
reference arithmetic above consumed.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from occuseg_b200 import scenes  # noqa: E402
from oracle import reference, rulebook as rb, rules_ref as rr  # noqa: E402


def pack(lists):
    """list of [n,2] -> (flat int32 [sum n,2], offsets int64 [V+1])"""
    off = np.concatenate([[0], np.cumsum([len(r) for r in lists])]).astype(np.int64)
    flat = np.concatenate(lists, 0).astype(np.int32) if off[-1] else np.zeros((0, 2), np.int32)
    return flat, off


def make(name, preset, seeds, cin, cout, seed):
    assert reference.available(), "oracle/_ref not built"
    rng = np.random.default_rng(seed)
    coords, feats = scenes.make_batch(preset, seeds)
    B = len(seeds)
    vox = rb.voxelize(coords, B)
    locs = vox["locs"]
    N = len(locs)
    subm = rb.submanifold_rules(locs, B)
    clocs, strided = rb.strided_rules(locs, B)
    Nc = len(clocs)
    subm_c = rb.submanifold_rules(clocs, B)
    # pin: the reference's compiled CPU builders give the same relation through coordinates
    assert rr.available(), "oracle/_ref/scn_rules_ref.so not built"
    sc = rr.Scene(coords, B, 4)
    assert sc.n == N and np.array_equal(sc.locs[sc.row_of_point], locs[vox["row_of_point"]])
    assert np.array_equal(sc.submanifold(0), rr.relation_of_lists(subm, locs))
    rel, cl, _ = sc.strided()
    assert np.array_equal(rel, rr.relation_of_lists(strided, locs, clocs)) and len(cl) == Nc
    sc_c = rr.Scene(clocs, B, 4)
    assert np.array_equal(sc_c.submanifold(0), rr.relation_of_lists(subm_c, clocs))
    R = reference.Ref()
    R.load_submanifold(4096, subm, N)
    R.load_strided(4096, 2048, strided, N, Nc)

    x = rng.standard_normal((N, cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) * (2.0 / cin / 27) ** 0.5).astype(np.float32)
    g = rng.standard_normal((N, cout)).astype(np.float32)
    y, macs = R.subm_forward(4096, x, w)
    dx, dw = R.subm_backward(4096, x, g, w)

    w8 = (rng.standard_normal((8, cin, cout)) * (2.0 / cin / 8) ** 0.5).astype(np.float32)
    gc = rng.standard_normal((Nc, cout)).astype(np.float32)
    yc, _ = R.conv_forward(4096, 2048, x, w8)
    dxc, dw8 = R.conv_backward(4096, 2048, x, gc, w8)

    xd = rng.standard_normal((Nc, cout)).astype(np.float32)
    wd = (rng.standard_normal((8, cout, cin)) * (2.0 / cout / 8) ** 0.5).astype(np.float32)
    gd = rng.standard_normal((N, cin)).astype(np.float32)
    yd, _ = R.deconv_forward(2048, 4096, xd, wd)
    dxd, dwd = R.deconv_backward(2048, 4096, xd, gd, wd)

    gamma = rng.uniform(0.5, 1.5, cin).astype(np.float32)
    beta = rng.standard_normal(cin).astype(np.float32)
    bn_y, bn_mean, bn_invstd, bn_rm, bn_rv = reference.bn_forward(
        x, gamma, beta, np.zeros(cin, np.float32), np.ones(cin, np.float32), 1e-4, 0.9, True, 0.0)
    gb = rng.standard_normal((N, cin)).astype(np.float32)
    bn_dx, bn_dgamma, bn_dbeta = reference.bn_backward(x, bn_y, gb, gamma, beta, bn_mean, bn_invstd, 0.0)

    inp_mean = rb.input_layer_mean(feats, vox, True)
    sf, so = pack(rb.canonical(subm))
    tf, to = pack(rb.canonical(strided))
    cf, co = pack(rb.canonical(subm_c))
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(
        out, coords=coords.astype(np.int32), feats=feats, batch=np.int64(B), locs=locs.astype(np.int32),
        row_of_point=vox["row_of_point"], coarse_locs=clocs.astype(np.int32),
        subm_flat=sf, subm_off=so, strided_flat=tf, strided_off=to, subm_coarse_flat=cf, subm_coarse_off=co,
        input_mean=inp_mean, rules_pinned=np.int64(1),
        x=x, w=w, g=g, y=y, macs=np.float64(macs), dx=dx, dw=dw,
        w8=w8, gc=gc, yc=yc, dxc=dxc, dw8=dw8,
        xd=xd, wd=wd, gd=gd, yd=yd, dxd=dxd, dwd=dwd,
        gamma=gamma, beta=beta, bn_y=bn_y, bn_mean=bn_mean, bn_invstd=bn_invstd, bn_rm=bn_rm, bn_rv=bn_rv,
        gb=gb, bn_dx=bn_dx, bn_dgamma=bn_dgamma, bn_dbeta=bn_dbeta,
    )
    print(name, "N", N, "Nc", Nc, "rules", len(sf), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    make("room_b2_c16", (0.36, 0.30, 0.24, 1), (3, 4), 16, 16, seed=11)
    make("room_b1_c3_8", (0.30, 0.24, 0.20, 1), (5,), 3, 8, seed=12)
    make("room_b3_c32_64", (0.30, 0.24, 0.20, 1), (6, 7, 8), 32, 64, seed=13)
