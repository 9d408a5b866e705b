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


def make_guided(name, preset, seeds, c, seed):
    """Normal-guided rules (OccuSeg's use_normal): the lists are oracle/rulebook.py's restatement, checked here against the
    reference's compiled builders (remap_rules_with_normal on GPU-ordered lists, the CPU normal-guided builder, the normal-guided
    2/2 builder) before they are written; y / dx / dw and the strided results are outputs of the reference's CPU arithmetic on
    those lists -- run SINGLE-THREADED: the reference's rule_index_add_ (CPU/Convolution.cpp:21-33) is an OpenMP loop without
    atomics that relies on a target row appearing at most once per list, which guided lists do not guarantee for d_input (an
    input row can feed two outputs of different orientation classes through the same permuted tap); with several threads its
    d_input loses updates."""
    assert os.environ.get("OMP_NUM_THREADS") == "1", "generate the guided fixture with OMP_NUM_THREADS=1 (see the docstring)"
    assert reference.available() and rr.available(), "oracle/_ref not built"
    rng = np.random.default_rng(seed)
    coords, _ = scenes.make_batch(preset, seeds)
    coords = np.concatenate([coords, coords[:37]], 0)
    coords = coords[np.argsort(coords[:, 3], kind="stable")]
    B = len(seeds)
    pn = rng.standard_normal((len(coords), 3)).astype(np.float32)
    pn /= np.linalg.norm(pn, axis=1, keepdims=True)
    pn[::19] = 0
    pn[3::29] = np.float32([[0.5, 0.0, 0.5]])
    vox = rb.voxelize(coords, B)
    locs = vox["locs"]
    N = len(locs)
    vn = rb.voxel_normals(pn, vox)
    ori = rb.oriented_filter(vn)
    guided = rb.guided_submanifold_rules(rb.submanifold_rules(locs, B), vn)
    clocs, strided, cn = rb.guided_strided_rules(locs, vn, B)
    Nc = len(clocs)
    # pin against the reference-compiled builders (rows matched through coordinates)
    sc = rr.Scene(coords, B, 4)
    row_of = {tuple(l): i for i, l in enumerate(locs.tolist())}
    mine_of_ref = np.array([row_of[tuple(l)] for l in sc.locs.tolist()])
    mine = rr.relation_of_lists(guided, locs)
    assert np.array_equal(sc.submanifold(3, normals=vn[mine_of_ref]), mine)
    assert np.array_equal(sc.submanifold(1, normals=vn[mine_of_ref]), mine)
    rel, rcl, rcn = sc.strided(normals=vn[mine_of_ref])
    crow = {tuple(l): i for i, l in enumerate(clocs.tolist())}
    order = np.array([crow[tuple(l)] for l in rcl.tolist()])
    assert np.allclose(rcn, cn[order], rtol=0, atol=1e-6)
    a = np.sort(np.abs(cn), 1)
    tie = (a[:, 2] - a[:, 1]) < 1e-5                 # class unspecified in the reference itself (hash-map summation order)
    key = lambda r: r[np.lexsort(r[:, 1:].T[::-1])]
    ra, rm = key(rel), key(rr.relation_of_lists(strided, locs, clocs))
    assert np.array_equal(ra[:, 1:], rm[:, 1:])
    crows = np.array([crow[tuple(l)] for l in ra[:, 4:8].tolist()])
    assert np.array_equal(ra[~tie[crows], 0], rm[~tie[crows], 0])
    R = reference.Ref()
    R.load_submanifold(4096, guided, N)
    R.load_strided(4096, 2048, strided, N, Nc)
    x = rng.standard_normal((N, c)).astype(np.float32)
    w = (rng.standard_normal((27, c, c)) * (2.0 / c / 27) ** 0.5).astype(np.float32)
    g = rng.standard_normal((N, c)).astype(np.float32)
    y, _ = R.subm_forward(4096, x, w)
    dx, dw = R.subm_backward(4096, x, g, w)
    w8 = (rng.standard_normal((8, c, c)) * (2.0 / c / 8) ** 0.5).astype(np.float32)
    gc = rng.standard_normal((Nc, c)).astype(np.float32)
    yc, _ = R.conv_forward(4096, 2048, x, w8)
    dxc, dw8 = R.conv_backward(4096, 2048, x, gc, w8)
    gf, go = pack(rb.canonical(guided))
    tf, to = pack(rb.canonical(strided))
    out = os.path.join(ROOT, "tests", "golden", name + ".npz")
    np.savez_compressed(
        out, coords=coords.astype(np.int32), batch=np.int64(B), point_normals=pn, locs=locs.astype(np.int32),
        voxel_normals=vn, ori=ori, guided_flat=gf, guided_off=go, coarse_locs=clocs.astype(np.int32), coarse_normals=cn,
        strided_flat=tf, strided_off=to, coarse_ties=tie, x=x, w=w, g=g, y=y, dx=dx, dw=dw, w8=w8, gc=gc, yc=yc, dxc=dxc, dw8=dw8)
    print(name, "N", N, "Nc", Nc, "ties", int(tie.sum()), os.path.getsize(out) // 1024, "KiB")


if __name__ == "__main__":
    if os.environ.get("OMP_NUM_THREADS") != "1":          # the guided fixture needs the serial reference arithmetic
        os.environ["OMP_NUM_THREADS"] = "1"
        os.execv(sys.executable, [sys.executable] + sys.argv)
    if "--guided-only" in sys.argv:
        make_guided("guided_b2_c16", (0.30, 0.24, 0.20, 1), (9, 10), 16, seed=14)
        sys.exit(0)
    make("room_b2_c16", (0.36, 0.30, 0.24, 1), (3, 4), 16, 16, seed=11)
    make("room_b1_c3_8", (0.30, 0.24, 0.20, 1), (5,), 3, 8, seed=12)
    make("room_b3_c32_64", (0.30, 0.24, 0.20, 1), (6, 7, 8), 32, 64, seed=13)
    make_guided("guided_b2_c16", (0.30, 0.24, 0.20, 1), (9, 10), 16, seed=14)
