"""CPU tests of the checker itself: the numpy restatement against the committed golden vectors (which
were produced by the reference's own compiled CPU code), the compiled reference (oracle/_ref) when present,
and the edge cases the domain has (duplicates, single voxels, empty samples, sample boundaries, key aliasing)."""
import numpy as np
import pytest

from conftest import load_golden, rel_err, unpack
from oracle import arith, rulebook as rb


def _canon_equal(a, b):
    a, b = rb.canonical(a), rb.canonical(b)
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def test_rulebook_restatement_matches_golden(golden):
    g = golden
    B = int(g["batch"])
    vox = rb.voxelize(g["coords"].astype(np.int64), B)
    assert np.array_equal(vox["locs"], g["locs"])
    assert np.array_equal(vox["row_of_point"], g["row_of_point"])
    assert _canon_equal(rb.submanifold_rules(vox["locs"], B), unpack(g["subm_flat"], g["subm_off"]))
    clocs, strided = rb.strided_rules(vox["locs"], B)
    assert np.array_equal(clocs, g["coarse_locs"])
    assert _canon_equal(strided, unpack(g["strided_flat"], g["strided_off"]))
    assert _canon_equal(rb.submanifold_rules(clocs, B), unpack(g["subm_coarse_flat"], g["subm_coarse_off"]))


def test_rulebook_pinned_to_reference_compiled_builders(golden):
    """The pin: the reference's own CPU rule builders (Metadata/IOLayersRules.h:19-130,
    SubmanifoldConvolutionRules.h:114-209, ConvolutionRules.h:95-119), compiled from the reference tree into
    oracle/_ref/scn_rules_ref.so, produce exactly the relation {(tap, in xyz, out xyz)} and the point->voxel grouping
    that the golden lists (= the numpy restatement of the GPU builders) hold.  Rows are compared through coordinates
    because the CPU builders number rows in first-appearance order and the GPU ones by sorted key rank."""
    from oracle import rules_ref as rr
    if not rr.available():
        pytest.skip("oracle/_ref/scn_rules_ref.so not available")
    g = golden
    B = int(g["batch"])
    sc = rr.Scene(g["coords"].astype(np.int64), B, 4)
    locs = g["locs"].astype(np.int64)
    assert sc.n == len(locs) and set(map(tuple, sc.locs.tolist())) == set(map(tuple, locs.tolist()))
    # InputLayer: same voxel for every point, same points (same order) for every voxel
    assert np.array_equal(sc.locs[sc.row_of_point], locs[g["row_of_point"]])
    vox = rb.voxelize(g["coords"].astype(np.int64), B)
    pov = sc.points_of_voxels()
    for r in range(len(locs)):
        assert pov[tuple(locs[r].tolist())] == vox["rule_pts"][vox["rule_ptr"][r]:vox["rule_ptr"][r + 1]].tolist()
    # submanifold 3x3x3: plain builder (tap order converted z-major -> x-major) and the normal-guided builder with
    # the identity orientation (which enumerates x-major itself, RectangularRegions.h:77-92)
    mine = rr.relation_of_lists(unpack(g["subm_flat"], g["subm_off"]), locs)
    assert np.array_equal(sc.submanifold(0), mine)
    ident = np.tile(np.array([[1, 0, 0]], np.float32), (sc.n, 1))
    assert np.array_equal(sc.submanifold(1, normals=ident), mine)
    # size-2 / stride-2 convolution: same (tap, fine, coarse) relation, same coarse voxel set
    rel, cl, _ = sc.strided()
    clocs = g["coarse_locs"].astype(np.int64)
    assert np.array_equal(rel, rr.relation_of_lists(unpack(g["strided_flat"], g["strided_off"]), locs, clocs))
    assert set(map(tuple, cl.tolist())) == set(map(tuple, clocs.tolist()))


def test_restatement_pinned_on_a_scene_with_duplicates_and_empty_sample():
    from oracle import rules_ref as rr
    if not rr.available():
        pytest.skip("oracle/_ref/scn_rules_ref.so not available")
    from occuseg_b200 import scenes
    c, _ = scenes.make_batch("tiny", (4, 5))
    c[c[:, 3] == 1, 3] = 2                                   # sample 1 is empty
    c = np.concatenate([c, c[:7]], 0)                        # exact duplicates, out of batch order -> re-sort
    c = c[np.argsort(c[:, 3], kind="stable")]
    v = rb.voxelize(c, 3)
    sc = rr.Scene(c, 3, 4)
    assert sc.n == len(v["locs"])
    assert np.array_equal(sc.submanifold(0), rr.relation_of_lists(rb.submanifold_rules(v["locs"], 3), v["locs"]))
    cl, lists = rb.strided_rules(v["locs"], 3)
    assert np.array_equal(sc.strided()[0], rr.relation_of_lists(lists, v["locs"], cl))


def test_arith_port_matches_reference_outputs(golden):
    """golden y/dx/dw... are outputs of sparseconvnet/SCN/CPU/*.cpp; tolerance 1e-5 (fp32, GEMM summation order)."""
    g = golden
    N, Nc = len(g["locs"]), len(g["coarse_locs"])
    subm = unpack(g["subm_flat"], g["subm_off"])
    strided = unpack(g["strided_flat"], g["strided_off"])
    y, macs = arith.rule_conv_forward(g["x"], g["w"], subm, N)
    assert rel_err(y, g["y"]) < 1e-5 and macs == float(g["macs"])
    dx, dw = arith.rule_conv_backward(g["x"], g["g"], g["w"], subm)
    assert rel_err(dx, g["dx"]) < 1e-5 and rel_err(dw, g["dw"]) < 1e-5
    yc, _ = arith.rule_conv_forward(g["x"], g["w8"], strided, Nc)
    assert rel_err(yc, g["yc"]) < 1e-5
    dxc, dw8 = arith.rule_conv_backward(g["x"], g["gc"], g["w8"], strided)
    assert rel_err(dxc, g["dxc"]) < 1e-5 and rel_err(dw8, g["dw8"]) < 1e-5
    yd, _ = arith.rule_conv_forward(g["xd"], g["wd"], strided, N, in_col=1, out_col=0)
    assert rel_err(yd, g["yd"]) < 1e-5
    dxd, dwd = arith.rule_conv_backward(g["xd"], g["gd"], g["wd"], strided, in_col=1, out_col=0)
    assert rel_err(dxd, g["dxd"]) < 1e-5 and rel_err(dwd, g["dwd"]) < 1e-5
    C = g["x"].shape[1]
    out = arith.batchnorm_forward(g["x"], g["gamma"], g["beta"], np.zeros(C, np.float32), np.ones(C, np.float32))
    for got, key in zip(out, ["bn_y", "bn_mean", "bn_invstd", "bn_rm", "bn_rv"]):
        assert rel_err(got, g[key]) < 1e-5, key
    bdx, bdg, bdb = arith.batchnorm_backward(g["x"], g["bn_y"], g["gb"], g["gamma"], g["bn_mean"], g["bn_invstd"])
    assert rel_err(bdx, g["bn_dx"]) < 1e-5 and rel_err(bdg, g["bn_dgamma"]) < 1e-5 and rel_err(bdb, g["bn_dbeta"]) < 1e-5


def test_compiled_reference_reproduces_golden(golden):
    """When oracle/_ref is loadable (authoring container, or the prebuilt .so on the GPU box) the reference code must
    reproduce its own fixtures bit for bit (same binary, same inputs)."""
    from oracle import reference
    if not reference.available():
        pytest.skip("oracle/_ref not available")
    g = golden
    N, Nc = len(g["locs"]), len(g["coarse_locs"])
    R = reference.Ref()
    R.load_submanifold(4096, unpack(g["subm_flat"], g["subm_off"]), N)
    R.load_strided(4096, 2048, unpack(g["strided_flat"], g["strided_off"]), N, Nc)
    y, macs = R.subm_forward(4096, g["x"], g["w"])
    assert rel_err(y, g["y"]) < 1e-6 and macs == float(g["macs"])
    dx, dw = R.subm_backward(4096, g["x"], g["g"], g["w"])
    assert rel_err(dx, g["dx"]) < 1e-6 and rel_err(dw, g["dw"]) < 1e-6
    yc, _ = R.conv_forward(4096, 2048, g["x"], g["w8"])
    assert rel_err(yc, g["yc"]) < 1e-6


def test_identity_offset_known_answer():
    """Known-answer test: a rulebook holding only the centre offset turns the convolution into x @ W[13]."""
    rng = np.random.default_rng(0)
    x = rng.standard_normal((50, 5)).astype(np.float32)
    w = rng.standard_normal((27, 5, 7)).astype(np.float32)
    rules = [np.zeros((0, 2), np.int32) for _ in range(27)]
    rules[13] = np.stack([np.arange(50), np.arange(50)], 1).astype(np.int32)
    y, macs = arith.rule_conv_forward(x, w, rules, 50)
    assert np.allclose(y, x @ w[13], rtol=1e-6, atol=1e-6) and macs == 50 * 5 * 7


def test_offset_convention_is_x_major():
    """Two voxels differing by +1 in x only: the rule must land in offset (dx+1)*9+(0+1)*3+(0+1) (GPU builder
    convention, SubmanifoldRules_cuda.cu:63-73), i.e. 22 for dx=+1 and 4 for dx=-1."""
    locs = np.array([[10, 10, 10, 0], [11, 10, 10, 0]], np.int64)
    rules = rb.submanifold_rules(locs, 1)
    assert rules[22].tolist() == [[1, 0]] and rules[4].tolist() == [[0, 1]]
    assert rules[13].tolist() == [[0, 0], [1, 1]]
    assert sum(len(r) for r in rules) == 4
    # +1 in z -> offset 14
    locs = np.array([[10, 10, 10, 0], [10, 10, 11, 0]], np.int64)
    assert rb.submanifold_rules(locs, 1)[14].tolist() == [[1, 0]]


def test_voxelize_duplicates_order_and_batches():
    coords = np.array([[5, 5, 5, 0], [4, 5, 5, 0], [5, 5, 5, 0], [5, 5, 4, 0], [5, 5, 5, 1], [5, 5, 5, 1]], np.int64)
    v = rb.voxelize(coords, 2)
    # sample 0 sorted by (z,y,x): (5,5,4) < (4,5,5) < (5,5,5); sample 1 continues the numbering
    assert v["locs"].tolist() == [[5, 5, 4, 0], [4, 5, 5, 0], [5, 5, 5, 0], [5, 5, 5, 1]]
    assert v["row_of_point"].tolist() == [2, 1, 2, 0, 3, 3]
    assert v["max_repeat"] == 2 and v["sample_ctr"].tolist() == [0, 3, 4]
    # points of row 2 keep their original order (stable sort)
    assert v["rule_pts"][v["rule_ptr"][2]:v["rule_ptr"][3]].tolist() == [0, 2]
    feats = np.arange(12, dtype=np.float32).reshape(6, 2)
    mean = rb.input_layer_mean(feats, v, True)
    assert np.allclose(mean[2], (feats[0] + feats[2]) / 2) and np.allclose(mean[3], (feats[4] + feats[5]) / 2)


def test_no_rules_across_samples():
    locs = np.array([[10, 10, 10, 0], [11, 10, 10, 1]], np.int64)
    rules = rb.submanifold_rules(locs, 2)
    assert sum(len(r) for r in rules) == 2 and len(rules[13]) == 2


def test_empty_sample_in_batch():
    coords = np.array([[3, 3, 3, 0], [4, 3, 3, 2]], np.int64)   # sample 1 is empty
    v = rb.voxelize(coords, 3)
    assert v["sample_ctr"].tolist() == [0, 1, 1, 2]
    rules = rb.submanifold_rules(v["locs"], 3)
    assert sum(len(r) for r in rules) == 2
    clocs, strided = rb.strided_rules(v["locs"], 3)
    assert clocs.tolist() == [[1, 1, 1, 0], [2, 1, 1, 2]]


def test_strided_offsets_and_parent():
    locs = np.array([[2, 2, 2, 0], [3, 2, 2, 0], [2, 3, 2, 0], [2, 2, 3, 0], [4, 2, 2, 0]], np.int64)
    v = rb.voxelize(locs, 1)
    clocs, rules = rb.strided_rules(v["locs"], 1)
    assert clocs.tolist() == [[1, 1, 1, 0], [2, 1, 1, 0]]
    got = {}
    for k, r in enumerate(rules):
        for i, o in r.tolist():
            got[tuple(v["locs"][i, :3])] = (k, o)
    assert got[(2, 2, 2)] == (0, 0) and got[(3, 2, 2)] == (4, 0) and got[(2, 3, 2)] == (2, 0)
    assert got[(2, 2, 3)] == (1, 0) and got[(4, 2, 2)] == (0, 1)


def test_key31_aliasing_is_reproduced():
    """x=-1 wraps to an all-ones word exactly as on the device (never matches a voxel inside the valid range)."""
    assert int(rb.key31(-1, 5, 5)) == 0x7FFFFFFF
    assert int(rb.key31(3, 2, 1)) == (1 << 21) | (2 << 10) | 3


def test_total_rules_symmetry():
    """Property: the rule relation is symmetric -- list k reversed is list 26-k."""
    from occuseg_b200 import scenes
    c, _ = scenes.make_batch("tiny", (1, 2))
    v = rb.voxelize(c, 2)
    rules = rb.submanifold_rules(v["locs"], 2)
    for k in range(27):
        assert _canon_equal([rules[k][:, ::-1]], [rules[26 - k]])


@pytest.mark.parametrize("rate", [2, 3])
def test_dilated_rules_pinned_to_reference_compiled_builder(rate):
    """SubmanifoldConvolution(dilated_rate): the restatement against SubmanifoldConvolution_SgToRules(grid, rules, size, rate)
    compiled from the reference tree (Metadata/SubmanifoldConvolutionRules.h:114-153)."""
    from oracle import rules_ref as rr
    if not rr.available():
        pytest.skip("oracle/_ref/scn_rules_ref.so not available")
    from occuseg_b200 import scenes
    c, _ = scenes.make_batch("tiny", (7, 8))
    v = rb.voxelize(c, 2)
    sc = rr.Scene(c, 2, 4)
    mine = rr.relation_of_lists(rb.submanifold_rules(v["locs"], 2, rate), v["locs"])
    ref = sc.submanifold(0, dilated_rate=rate)
    assert len(ref) > len(v["locs"]) and np.array_equal(ref, mine)
    # the dilated relation really differs from the plain one, and taps sit `rate` voxels away
    plain = rr.relation_of_lists(rb.submanifold_rules(v["locs"], 2, 1), v["locs"])
    assert not np.array_equal(plain, mine)
    off = mine[:, 1:4] - mine[:, 4:7]
    assert set(np.unique(np.abs(off)).tolist()) <= {0, rate}


def _dyadic_normals(rng, n):
    """Normals with small integer components: every partial sum is exact in fp32, so the summation order (hash-map iteration
    order in the reference's strided builder) cannot matter, and ties |x| == |y| are frequent (the tie rule is exercised)."""
    v = rng.integers(-3, 4, (n, 3)).astype(np.float32)
    v[(v == 0).all(1)] = (0, 0, 1)
    return v


def test_normal_guided_restatement_pinned_to_reference_builders():
    """oracle/rulebook.py's normal-guided rules == the reference's remap_rules_with_normal applied to GPU-ordered lists
    (SubmanifoldConvolutionRules.h:213-245,486-490), == its CPU normal-guided builder (:159-209), and the guided 2/2 rules ==
    Convolution_InputSgToRulesAndOutputSg with normals (ConvolutionRules.h:18-92), coarse normals included."""
    from oracle import rules_ref as rr
    if not rr.available():
        pytest.skip("oracle/_ref/scn_rules_ref.so not available")
    from occuseg_b200 import scenes
    rng = np.random.default_rng(5)
    c, _ = scenes.make_batch("tiny", (2, 3))
    c = np.concatenate([c, c[:11]], 0)
    c = c[np.argsort(c[:, 3], kind="stable")]
    B = 2
    v = rb.voxelize(c, B)
    locs = v["locs"]
    pn = _dyadic_normals(rng, len(c))
    vn = rb.voxel_normals(pn, v)
    sc = rr.Scene(c, B, 4)
    row_of = {tuple(l): i for i, l in enumerate(locs.tolist())}
    mine_of_ref = np.array([row_of[tuple(l)] for l in sc.locs.tolist()])
    vn_ref_order = vn[mine_of_ref]
    # the orientation classes agree with the reference's OrientedFilter
    ori = rb.oriented_filter(vn)
    assert all(rr.oriented_filter(vn[i]) == ori[i] for i in range(0, len(vn), 7))
    assert set(np.unique(ori).tolist()) == {0, 2, 4}
    guided = rb.guided_submanifold_rules(rb.submanifold_rules(locs, B), vn)
    mine = rr.relation_of_lists(guided, locs)
    assert np.array_equal(sc.submanifold(3, normals=vn_ref_order), mine)
    assert np.array_equal(sc.submanifold(1, normals=vn_ref_order), mine)
    assert not np.array_equal(sc.submanifold(0), mine)              # the permutation does something
    # strided.  The reference sums the children's normals in hash-map iteration order, so with general (normalised) fine
    # normals the last bits of a coarse normal -- and with them the class of a coarse voxel whose two largest components tie
    # -- depend on that order: (a) integer-valued fine normals (every sum exact) must agree bit for bit, (b) general ones agree
    # to 1e-6 and on every coarse voxel that is not within 1e-5 of a tie
    crow = None
    for fine_normals, exact in ((_dyadic_normals(rng, len(locs)), True), (vn, False)):
        cl, lists, cn = rb.guided_strided_rules(locs, fine_normals, B)
        rel, rcl, rcn = sc.strided(normals=fine_normals[mine_of_ref])
        mine = rr.relation_of_lists(lists, locs, cl)
        crow = {tuple(l): i for i, l in enumerate(cl.tolist())}
        order = np.array([crow[tuple(l)] for l in rcl.tolist()])
        if exact:
            assert np.array_equal(rel, mine)
            assert np.array_equal(rcn, cn[order])                    # coarse normals bit-identical
            assert len(np.unique(rb.oriented_filter(cn))) == 3
        else:
            assert np.allclose(rcn, cn[order], rtol=0, atol=1e-6)
            a = np.sort(np.abs(cn), 1)
            tie = (a[:, 2] - a[:, 1]) < 1e-5
            key = lambda r: r[np.lexsort(r[:, 1:].T[::-1])]          # the (fine, coarse) pairs are the same; only taps can differ
            ra, rm = key(rel), key(mine)
            assert np.array_equal(ra[:, 1:], rm[:, 1:])
            coarse_row = np.array([crow[tuple(l)] for l in ra[:, 4:8].tolist()])
            assert np.array_equal(ra[~tie[coarse_row], 0], rm[~tie[coarse_row], 0])
            assert tie.sum() < 0.1 * len(cn)     # integer point normals make exact ties common (5 %)


def test_normal_guided_golden_fixture():
    """tests/golden/guided_b2_c16.npz (lists checked against the reference's compiled builders when it was written, floats =
    outputs of the reference's CPU arithmetic on those lists): the restatement reproduces the integer part bit for bit, the
    numpy port of the arithmetic the floating-point part within 1e-5 -- on a box without the reference tree too."""
    g = load_golden("guided_b2_c16")
    B = int(g["batch"])
    coords = g["coords"].astype(np.int64)
    vox = rb.voxelize(coords, B)
    assert np.array_equal(vox["locs"], g["locs"])
    vn = rb.voxel_normals(g["point_normals"], vox)
    assert np.array_equal(vn, g["voxel_normals"]) and np.array_equal(rb.oriented_filter(vn), g["ori"])
    guided = rb.guided_submanifold_rules(rb.submanifold_rules(vox["locs"], B), vn)
    want = unpack(g["guided_flat"], g["guided_off"])
    assert all(np.array_equal(a, b) for a, b in zip(rb.canonical(guided), want))
    cl, strided, cn = rb.guided_strided_rules(vox["locs"], vn, B)
    assert np.array_equal(cl, g["coarse_locs"]) and np.array_equal(cn, g["coarse_normals"])
    assert all(np.array_equal(a, b) for a, b in zip(rb.canonical(strided), unpack(g["strided_flat"], g["strided_off"])))
    N, Nc = len(vox["locs"]), len(cl)
    y, _ = arith.rule_conv_forward(g["x"], g["w"], guided, N)
    dx, dw = arith.rule_conv_backward(g["x"], g["g"], g["w"], guided)
    assert rel_err(y, g["y"]) < 1e-5 and rel_err(dx, g["dx"]) < 1e-5 and rel_err(dw, g["dw"]) < 1e-5
    yc, _ = arith.rule_conv_forward(g["x"], g["w8"], strided, Nc)
    dxc, dw8 = arith.rule_conv_backward(g["x"], g["gc"], g["w8"], strided)
    assert rel_err(yc, g["yc"]) < 1e-5 and rel_err(dxc, g["dxc"]) < 1e-5 and rel_err(dw8, g["dw8"]) < 1e-5


def test_tile_order_key_puts_rare_taps_on_top():
    """The design claim behind the sort key of the tile order (occuseg_b200/csrc/meta.cu: tap_key_bit), re-derived on the CPU
    with tools/sim_tile_order.py: sorting the rows of a scene by a key whose most significant bits are the rarely present taps
    (corners, then edges) and whose least significant bits are the frequent ones (faces) leaves fewer non-empty (tile group,
    tap) pairs -- pipeline items of the convolution kernel -- than the plain tap order, and both beat the natural row order."""
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
    import sim_tile_order as sim
    from occuseg_b200 import scenes
    coords, _ = scenes.make_batch("S100k", (0,))
    locs = rb.voxelize(coords, 1)["locs"]
    pat = sim.patterns_of(locs, 1)

    def cls(t):
        return (t // 9 != 1) + ((t // 3) % 3 != 1) + (t % 3 != 1)        # 0 centre, 1 face, 2 edge, 3 corner
    shipped = sorted(range(27), key=lambda t: ({0: 0, 3: 1, 2: 2, 1: 3}[cls(t)], t))     # most significant first
    items = {name: sim.cost(pat if order is None else sim.sorted_by(pat, sim.permute_bits(pat, order), 262144))["items"]
             for name, order in (("natural", None), ("tap order", list(range(26, -1, -1))), ("shipped", shipped))}
    assert items["shipped"] < 0.95 * items["tap order"] < 0.95 * items["natural"], items
