"""GPU parity tests (-m gpu) of the normal-guided rules (OccuSeg's `use_normal`): per-voxel normals, orientation classes, the
tap-permuted submanifold table and the tap-permuted 2/2 rules bit-exact against oracle/rulebook.py's restatement -- which
tests/test_oracle.py pins to the reference's compiled builders (remap_rules_with_normal, SubmanifoldConvolutionRules.h:213-245;
Convolution_InputSgToRulesAndOutputSg with normals, ConvolutionRules.h:18-92) -- and the three products of SubmanifoldConvolution
/ Convolution / Deconvolution on those rules against the reference's CPU arithmetic (fp32 1e-5, tensor-core tiles 2e-2)."""
import numpy as np
import pytest

from conftest import have_cuda, rel_err

pytestmark = pytest.mark.gpu

if have_cuda():
    import torch
    import occuseg_b200.sparseconvnet as scn
    from occuseg_b200.sparseconvnet import SCN
    from occuseg_b200 import scenes
from oracle import arith, rulebook as rb

SIZE = 4096
FP32_TOL, TC_TOL = 1e-5, 2e-2


@pytest.fixture(autouse=True)
def _default_precision():
    scn.set_precision("fp32")
    yield
    scn.set_precision("fp32")


def lt(v):
    return torch.LongTensor([v, v, v])


def cu(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def _table(rules, n):
    t = np.full((len(rules), n), -1, np.int32)
    for k, r in enumerate(rules):
        assert len(np.unique(r[:, 1])) == len(r)
        t[k, r[:, 1]] = r[:, 0]
    return t


def _scene(seed=0):
    rng = np.random.default_rng(seed)
    coords, _ = scenes.make_batch("small", (3 + seed, 4 + seed))
    coords = np.concatenate([coords, coords[:53]], 0)               # duplicate points: normals are averaged per voxel
    coords = coords[np.argsort(coords[:, 3], kind="stable")]
    pn = rng.standard_normal((len(coords), 3)).astype(np.float32)
    pn /= np.linalg.norm(pn, axis=1, keepdims=True)
    pn[::17] = 0                                                     # degenerate normals (|n| < 1e-8 stays unnormalised, class 0)
    pn[5::23] = np.float32([[0.5, 0.5, 0.0]])                        # exact ties: the x >= y >= z preference order decides
    return coords, pn


def _guided_meta(coords, pn, batch, guide_scale):
    m = SCN.Metadata_3()
    m.setNormalGuideScale(guide_scale)
    out = torch.empty(0, device="cuda")
    SCN.InputLayer_updateOutput(m, lt(SIZE), torch.from_numpy(coords), torch.zeros(len(coords), 1, device="cuda"), out, batch, 4,
                                torch.from_numpy(pn))
    return m


def _strided_lists(parent, off):
    rows = np.arange(len(parent), dtype=np.int32)
    return [np.stack([rows[off == k], parent[off == k]], 1) for k in range(8)]


def test_guided_rules_are_bit_exact():
    coords, pn = _scene()
    vox = rb.voxelize(coords, 2)
    locs, N = vox["locs"], len(vox["locs"])
    vn = rb.voxel_normals(pn, vox)
    m = _guided_meta(coords, pn, 2, SIZE)             # strided layers from the input scale on are guided
    assert m.guided(lt(SIZE))
    assert np.array_equal(m.normalsOf(lt(SIZE)).numpy(), vn)
    tbl, ori = m.submanifoldGuidedTable(lt(SIZE))
    assert np.array_equal(ori.numpy(), rb.oriented_filter(vn)) and set(np.unique(ori.numpy()).tolist()) == {0, 2, 4}
    plain = rb.submanifold_rules(locs, 2)
    guided = rb.guided_submanifold_rules(plain, vn)
    assert np.array_equal(tbl.numpy(), _table(guided, N))
    assert not np.array_equal(tbl.numpy(), _table(plain, N))
    assert np.array_equal(m.submanifoldNeighbourTable(lt(SIZE))[0].numpy(), _table(plain, N))     # the plain table is still there
    # 2/2 rules: taps permuted by the class of the coarse voxel, coarse normals = normalised mean of the children's
    cl, lists, cn = rb.guided_strided_rules(locs, vn, 2)
    parent, off, nc = m.stridedTable(lt(SIZE), lt(SIZE // 2))
    assert nc == len(cl) and np.array_equal(m.getSpatialLocations(lt(SIZE // 2)).numpy(), cl)
    got = _strided_lists(parent.numpy(), off.numpy())
    assert all(np.array_equal(a, b) for a, b in zip(rb.canonical(got), rb.canonical(lists)))
    _, plain8 = rb.strided_rules(locs, 2)
    assert not all(np.array_equal(a, b) for a, b in zip(rb.canonical(got), rb.canonical(plain8)))
    assert m.guided(lt(SIZE // 2))
    assert np.array_equal(m.normalsOf(lt(SIZE // 2)).numpy(), cn)
    tbl2, ori2 = m.submanifoldGuidedTable(lt(SIZE // 2))
    assert np.array_equal(ori2.numpy(), rb.oriented_filter(cn))
    assert np.array_equal(tbl2.numpy(), _table(rb.guided_submanifold_rules(rb.submanifold_rules(cl, 2), cn), len(cl)))
    # the next strided layer starts below the guide scale: plain rules, no normals on its output (ConvolutionRules.h:774)
    p2, o2, nc2 = m.stridedTable(lt(SIZE // 2), lt(SIZE // 4))
    cl2, plain_c = rb.strided_rules(cl, 2)
    assert nc2 == len(cl2)
    assert all(np.array_equal(a, b) for a, b in zip(rb.canonical(_strided_lists(p2.numpy(), o2.numpy())), rb.canonical(plain_c)))
    assert not m.guided(lt(SIZE // 4))
    # reference default (guide scale above every spatial size): only the submanifold rules of the input scale are guided
    m3 = _guided_meta(coords, pn, 2, 102400)
    p3, o3, _ = m3.stridedTable(lt(SIZE), lt(SIZE // 2))
    assert all(np.array_equal(a, b) for a, b in zip(rb.canonical(_strided_lists(p3.numpy(), o3.numpy())), rb.canonical(plain8)))
    assert m3.guided(lt(SIZE)) and not m3.guided(lt(SIZE // 2))
    # no normals: nothing is guided
    m4 = SCN.Metadata_3()
    SCN.InputLayer_updateOutput(m4, lt(SIZE), torch.from_numpy(coords), torch.zeros(len(coords), 1, device="cuda"),
                                torch.empty(0, device="cuda"), 2, 4, None)
    assert not m4.guided(lt(SIZE))
    with pytest.raises(Exception):
        m4.submanifoldGuidedTable(lt(SIZE))


@pytest.mark.parametrize("precision,c_in,c_out", [("fp32", 16, 24), ("bf16", 64, 64), ("bf16", 128, 64), ("tf32", 32, 96)])
def test_guided_convolutions_match_the_reference_arithmetic(precision, c_in, c_out):
    coords, pn = _scene(1)
    vox = rb.voxelize(coords, 2)
    locs, N = vox["locs"], len(vox["locs"])
    vn = rb.voxel_normals(pn, vox)
    rules = rb.guided_submanifold_rules(rb.submanifold_rules(locs, 2), vn)
    cl, lists, cn = rb.guided_strided_rules(locs, vn, 2)
    Nc = len(cl)
    tol = FP32_TOL if precision == "fp32" else TC_TOL
    rng = np.random.default_rng(7)
    m = _guided_meta(coords, pn, 2, SIZE)
    scn.set_precision(precision)

    def run(layer, x, g, size):
        xin = cu(x).requires_grad_(True)
        y = layer(scn.SparseConvNetTensor(xin, m, lt(size))).features
        y.backward(cu(g))
        return y.detach().cpu().numpy(), xin.grad.cpu().numpy(), layer.weight.grad.cpu().numpy()

    # SubmanifoldConvolution: forward on the guided table, dgrad as one pass per orientation class, wgrad on the guided lists
    conv = scn.SubmanifoldConvolution(3, c_in, c_out, 3, False).cuda()
    x = rng.standard_normal((N, c_in), dtype=np.float32)
    g = rng.standard_normal((N, c_out), dtype=np.float32)
    w = conv.weight.detach().cpu().numpy()
    y, gx, gw = run(conv, x, g, SIZE)
    y0, _ = arith.rule_conv_forward(x, w, rules, N)
    gx0, gw0 = arith.rule_conv_backward(x, g, w, rules)
    assert rel_err(y, y0) < tol and rel_err(gx, gx0) < tol and rel_err(gw, gw0) < tol
    # and it is NOT the plain convolution
    yp, _ = arith.rule_conv_forward(x, w, rb.submanifold_rules(locs, 2), N)
    assert rel_err(y, yp) > 10 * tol
    # Convolution 2/2 (fine -> coarse) and Deconvolution 2/2 (coarse -> fine) on the guided lists
    down = scn.Convolution(3, c_in, c_out, 2, 2, False).cuda()
    gc = rng.standard_normal((Nc, c_out), dtype=np.float32)
    wd = down.weight.detach().cpu().numpy()
    y, gx, gw = run(down, x, gc, SIZE)
    y0, _ = arith.rule_conv_forward(x, wd, lists, Nc)
    gx0, gw0 = arith.rule_conv_backward(x, gc, wd, lists)
    assert rel_err(y, y0) < tol and rel_err(gx, gx0) < tol and rel_err(gw, gw0) < tol
    up = scn.Deconvolution(3, c_out, c_in, 2, 2, False).cuda()
    xc = rng.standard_normal((Nc, c_out), dtype=np.float32)
    gf = rng.standard_normal((N, c_in), dtype=np.float32)
    wu = up.weight.detach().cpu().numpy()
    y, gx, gw = run(up, xc, gf, SIZE // 2)
    y0, _ = arith.rule_conv_forward(xc, wu, lists, N, in_col=1, out_col=0)
    gx0, gw0 = arith.rule_conv_backward(xc, gf, wu, lists, in_col=1, out_col=0)
    assert rel_err(y, y0) < tol and rel_err(gx, gx0) < tol and rel_err(gw, gw0) < tol


@pytest.mark.parametrize("precision", ["bf16", "fp32"])
def test_guided_network_step_runs_through_the_public_api(precision):
    """InputLayer with a normals tensor (input[2], as examples/ScanNet/model.py passes it) -> residual UNet -> OutputLayer: a
    training step on guided rules; the fused BatchNorm -> convolution nodes step aside for guided submanifold layers, so the
    bf16 run tracks the exact-fp32 run of the same network."""
    coords, pn = _scene(2)
    feats = np.random.default_rng(3).standard_normal((len(coords), 3)).astype(np.float32)

    def step(prec):
        scn.set_precision(prec)
        torch.manual_seed(5)
        net = scn.Sequential().add(scn.InputLayer(3, SIZE, mode=4, normal_guide_scale=SIZE // 2)) \
            .add(scn.SubmanifoldConvolution(3, 3, 64, 3, False)).add(scn.UNet(3, 1, [64, 128, 192], True)) \
            .add(scn.BatchNormReLU(64)).add(scn.OutputLayer(3)).cuda()
        out = net([torch.from_numpy(coords), cu(feats), torch.from_numpy(pn), 2])
        out.square().mean().backward()
        return out.detach().cpu().numpy(), [p.grad.detach().cpu().numpy() for p in net.parameters()]

    out, grads = step(precision)
    assert np.isfinite(out).all() and all(np.isfinite(g).all() for g in grads)
    if precision == "bf16":
        out0, grads0 = step("fp32")
        l2 = lambda a, b: np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30)
        assert l2(out, out0) < 5e-2


def test_guided_golden_fixture():
    """The committed fixture (reference-checked lists, reference CPU arithmetic): CUDA normals / classes / guided tables / guided
    2/2 rules bit-exact, fp32 products within 1e-5 of the reference outputs."""
    from conftest import load_golden, unpack
    g = load_golden("guided_b2_c16")
    B = int(g["batch"])
    coords = g["coords"].astype(np.int64)
    m = _guided_meta(coords, g["point_normals"], B, SIZE)
    N = len(g["locs"])
    assert np.array_equal(m.getSpatialLocations(lt(SIZE)).numpy(), g["locs"])
    assert np.array_equal(m.normalsOf(lt(SIZE)).numpy(), g["voxel_normals"])
    tbl, ori = m.submanifoldGuidedTable(lt(SIZE))
    assert np.array_equal(ori.numpy(), g["ori"])
    assert np.array_equal(tbl.numpy(), _table(unpack(g["guided_flat"], g["guided_off"]), N))
    parent, off, nc = m.stridedTable(lt(SIZE), lt(SIZE // 2))
    assert nc == len(g["coarse_locs"]) and np.array_equal(m.normalsOf(lt(SIZE // 2)).numpy(), g["coarse_normals"])
    got = _strided_lists(parent.numpy(), off.numpy())
    assert all(np.array_equal(a, b) for a, b in zip(rb.canonical(got), unpack(g["strided_flat"], g["strided_off"])))
    scn.set_precision("fp32")
    c = g["x"].shape[1]
    conv = scn.SubmanifoldConvolution(3, c, c, 3, False).cuda()
    down = scn.Convolution(3, c, c, 2, 2, False).cuda()
    with torch.no_grad():
        conv.weight.copy_(cu(g["w"]))
        down.weight.copy_(cu(g["w8"]))
    for layer, gout, ky, kdx, kdw in ((conv, g["g"], "y", "dx", "dw"), (down, g["gc"], "yc", "dxc", "dw8")):
        xin = cu(g["x"]).requires_grad_(True)
        y = layer(scn.SparseConvNetTensor(xin, m, lt(SIZE))).features
        y.backward(cu(gout))
        assert rel_err(y.detach().cpu().numpy(), g[ky]) < FP32_TOL
        assert rel_err(xin.grad.cpu().numpy(), g[kdx]) < FP32_TOL
        assert rel_err(layer.weight.grad.cpu().numpy(), g[kdw]) < FP32_TOL
