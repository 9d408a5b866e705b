"""GPU parity tests proper (-m gpu): the CUDA path, called through the C ABI (ctypes -> libscn_b200.so), against
 (a) the committed golden vectors = outputs of the reference's own CPU code (tests/golden/make_golden.py),
 (b) the oracle on seeded scenes at sizes it finishes in seconds,
 (c) size-independent properties at BASELINE.json's full sizes.
Tolerances (max-norm relative, conftest.rel_err): fp32 path 1e-5; tf32 / bf16 tensor-core tiles 2e-2 (north_star).
Integer results (row order, rulebooks, MAC counts) must be bit-exact."""
import numpy as np
import pytest

from conftest import GOLDEN_NAMES, have_cuda, load_golden, rel_err, unpack

pytestmark = pytest.mark.gpu

if have_cuda():
    import torch
    import occuseg_b200.sparseconvnet as scn
    from occuseg_b200.sparseconvnet import SCN
    from occuseg_b200 import scenes, _lib
from oracle import arith, rulebook as rb

FP32_TOL = 1e-5
TF32_TOL = 2e-2
SIZE = 4096


@pytest.fixture(autouse=True)
def _tensor_core_mode():
    """The package default is the exact fp32 path; the tests in this file opt into bf16 tiles unless they say otherwise."""
    scn.set_precision("bf16")
    yield
    scn.set_precision("fp32")


def lt(v):
    return torch.LongTensor([v, v, v])


def cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def build_meta(coords, batch, feats=None, mode=4):
    m = SCN.Metadata_3()
    c = torch.from_numpy(np.ascontiguousarray(coords, dtype=np.int64))
    f = cu(feats if feats is not None else np.zeros((len(coords), 1), np.float32))
    out = torch.empty(0, device="cuda")
    SCN.InputLayer_updateOutput(m, lt(SIZE), c, f, out, batch, mode, None)
    return m, out


def canon_equal(a, b):
    a, b = rb.canonical(a), rb.canonical(b)
    return len(a) == len(b) and all(np.array_equal(x, y) for x, y in zip(a, b))


def strided_lists(parent, off):
    rows = np.arange(len(parent), dtype=np.int32)
    return [np.stack([rows[off == k], parent[off == k]], 1) for k in range(8)]


# ---------------------------------------------------------------------------------------- rulebooks (bit-exact)
@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_rulebooks_match_golden(name):
    g = load_golden(name)
    B = int(g["batch"])
    m, feats = build_meta(g["coords"], B, g["feats"])
    assert np.array_equal(m.getSpatialLocations(lt(SIZE)).numpy(), g["locs"].astype(np.int64))
    nbr, n_rules = m.submanifoldNeighbourTable(lt(SIZE))
    assert n_rules == len(g["subm_flat"])
    assert canon_equal(rb.rules_from_neighbour_table(nbr.numpy()), unpack(g["subm_flat"], g["subm_off"]))
    parent, off, nc = m.stridedTable(lt(SIZE), lt(SIZE // 2))
    assert nc == len(g["coarse_locs"])
    assert canon_equal(strided_lists(parent.numpy(), off.numpy()), unpack(g["strided_flat"], g["strided_off"]))
    assert np.array_equal(m.getSpatialLocations(lt(SIZE // 2)).numpy(), g["coarse_locs"].astype(np.int64))
    nbr_c, _ = m.submanifoldNeighbourTable(lt(SIZE // 2))
    assert canon_equal(rb.rules_from_neighbour_table(nbr_c.numpy()), unpack(g["subm_coarse_flat"], g["subm_coarse_off"]))
    # InputLayer mode 4 mean (fp32, same summation order as the reference kernel)
    assert rel_err(feats.cpu().numpy(), g["input_mean"]) < FP32_TOL


@pytest.mark.parametrize("preset,seeds", [("small", (0, 1, 2)), ("S100k", (0,))])
def test_rulebooks_match_oracle_on_scenes(preset, seeds):
    coords, feats = scenes.make_batch(preset, seeds)
    B = len(seeds)
    vox = rb.voxelize(coords, B)
    m, _ = build_meta(coords, B)
    locs = vox["locs"]
    size = SIZE
    for level in range(4):
        assert np.array_equal(m.getSpatialLocations(lt(size)).numpy(), locs)
        nbr, n_rules = m.submanifoldNeighbourTable(lt(size))
        want = rb.submanifold_rules(locs, B)
        assert n_rules == sum(len(r) for r in want)
        assert canon_equal(rb.rules_from_neighbour_table(nbr.numpy()), want)
        parent, off, nc = m.stridedTable(lt(size), lt(size // 2))
        locs, want_s = rb.strided_rules(locs, B)
        assert nc == len(locs)
        assert canon_equal(strided_lists(parent.numpy(), off.numpy()), want_s)
        size //= 2


def test_rulebook_edge_cases():
    # duplicates, an empty sample in the middle, a single-voxel sample, unsorted points inside a sample
    coords = np.array([[5, 5, 5, 0], [4, 5, 5, 0], [5, 5, 5, 0], [5, 5, 4, 0], [0, 0, 0, 2], [1, 0, 0, 2], [0, 0, 0, 2],
                       [7, 7, 7, 3]], np.int64)
    feats = np.arange(16, dtype=np.float32).reshape(8, 2)
    m, out = build_meta(coords, 4, feats)
    vox = rb.voxelize(coords, 4)
    assert np.array_equal(m.getSpatialLocations(lt(SIZE)).numpy(), vox["locs"])
    assert rel_err(out.cpu().numpy(), rb.input_layer_mean(feats, vox, True)) < FP32_TOL
    nbr, n_rules = m.submanifoldNeighbourTable(lt(SIZE))
    assert canon_equal(rb.rules_from_neighbour_table(nbr.numpy()), rb.submanifold_rules(vox["locs"], 4))
    parent, off, nc = m.stridedTable(lt(SIZE), lt(SIZE // 2))
    clocs, want = rb.strided_rules(vox["locs"], 4)
    assert nc == len(clocs) and canon_equal(strided_lists(parent.numpy(), off.numpy()), want)


def test_bad_inputs_raise():
    from occuseg_b200._lib import ScnError
    with pytest.raises(ScnError):
        build_meta(np.array([[1, 2, -3, 0]], np.int64), 1)            # negative coordinate
    with pytest.raises(ScnError):
        build_meta(np.array([[1, 2, 3, 5]], np.int64), 2)             # batch index out of range
    m, _ = build_meta(np.array([[1, 2, 3, 0]], np.int64), 1)
    with pytest.raises(ScnError):
        m.submanifoldNeighbourTable(lt(123))                          # unknown scale
    with pytest.raises(TypeError):
        SCN.BatchNormalization_updateOutput(torch.zeros(4, 4), torch.empty(0), None, None, None, None, None, None,
                                            1e-4, 0.9, True, 0.0)     # CPU tensor: no CPU path


# ---------------------------------------------------------------------------------------- arithmetic vs golden
def _subm(m, x, w, g, size=SIZE):
    y = torch.empty(0, device="cuda")
    macs = SCN.SubmanifoldConvolution_updateOutput(lt(size), lt(3), m, x, y, w, torch.empty(0), 1)
    dx, dw = torch.empty(0, device="cuda"), torch.zeros_like(w)
    SCN.SubmanifoldConvolution_backward(lt(size), lt(3), m, x, dx, g, w, dw, torch.empty(0), 1)
    return y, macs, dx, dw


@pytest.mark.parametrize("name", GOLDEN_NAMES)
@pytest.mark.parametrize("precision", ["fp32", "tf32", "bf16"])
def test_convolutions_match_reference_outputs(name, precision):
    g = load_golden(name)
    tol = FP32_TOL if precision == "fp32" else TF32_TOL
    scn.set_precision(precision)
    try:
        m, _ = build_meta(g["coords"], int(g["batch"]))
        x, w, go = cu(g["x"]), cu(g["w"]), cu(g["g"])
        y, macs, dx, dw = _subm(m, x, w, go)
        assert macs == float(g["macs"])
        assert rel_err(y.cpu().numpy(), g["y"]) < tol
        assert rel_err(dx.cpu().numpy(), g["dx"]) < tol
        assert rel_err(dw.cpu().numpy(), g["dw"]) < tol
        # strided convolution
        w8, gc = cu(g["w8"]), cu(g["gc"])
        yc = torch.empty(0, device="cuda")
        SCN.Convolution_updateOutput(lt(SIZE), lt(SIZE // 2), lt(2), lt(2), m, x, yc, w8, torch.empty(0))
        assert rel_err(yc.cpu().numpy(), g["yc"]) < tol
        dxc, dw8 = torch.empty(0, device="cuda"), torch.zeros_like(w8)
        SCN.Convolution_backward(lt(SIZE), lt(SIZE // 2), lt(2), lt(2), m, x, dxc, gc, w8, dw8, torch.empty(0))
        assert rel_err(dxc.cpu().numpy(), g["dxc"]) < tol
        assert rel_err(dw8.cpu().numpy(), g["dw8"]) < tol
        # deconvolution over the same rulebook
        xd, wd, gd = cu(g["xd"]), cu(g["wd"]), cu(g["gd"])
        yd = torch.empty(0, device="cuda")
        SCN.Deconvolution_updateOutput(lt(SIZE // 2), lt(SIZE), lt(2), lt(2), m, xd, yd, wd, torch.empty(0))
        assert rel_err(yd.cpu().numpy(), g["yd"]) < tol
        dxd, dwd = torch.empty(0, device="cuda"), torch.zeros_like(wd)
        SCN.Deconvolution_backward(lt(SIZE // 2), lt(SIZE), lt(2), lt(2), m, xd, dxd, gd, wd, dwd, torch.empty(0))
        assert rel_err(dxd.cpu().numpy(), g["dxd"]) < tol
        assert rel_err(dwd.cpu().numpy(), g["dwd"]) < tol
    finally:
        scn.set_precision("bf16")


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_batchnorm_matches_reference_outputs(name):
    g = load_golden(name)
    C = g["x"].shape[1]
    x, gamma, beta = cu(g["x"]), cu(g["gamma"]), cu(g["beta"])
    rm, rv = torch.zeros(C, device="cuda"), torch.ones(C, device="cuda")
    y, sm, si = torch.empty(0, device="cuda"), torch.empty(C, device="cuda"), torch.empty(C, device="cuda")
    SCN.BatchNormalization_updateOutput(x, y, sm, si, rm, rv, gamma, beta, 1e-4, 0.9, True, 0.0)
    for got, key in [(y, "bn_y"), (sm, "bn_mean"), (si, "bn_invstd"), (rm, "bn_rm"), (rv, "bn_rv")]:
        assert rel_err(got.cpu().numpy(), g[key]) < FP32_TOL, key
    gb = cu(g["gb"])
    gb_before = gb.clone()
    dx, dg, db = torch.empty(0, device="cuda"), torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    SCN.BatchNormalization_backward(x, dx, y, gb, sm, si, rm, rv, gamma, beta, dg, db, 0.0)
    assert torch.equal(gb, gb_before)       # d_output is not mutated (documented deviation)
    assert rel_err(dx.cpu().numpy(), g["bn_dx"]) < FP32_TOL
    assert rel_err(dg.cpu().numpy(), g["bn_dgamma"]) < FP32_TOL
    assert rel_err(db.cpu().numpy(), g["bn_dbeta"]) < FP32_TOL
    # eval mode uses the running statistics
    y2 = torch.empty(0, device="cuda")
    SCN.BatchNormalization_updateOutput(x, y2, sm, si, rm, rv, gamma, beta, 1e-4, 0.9, False, 0.333)
    want = arith.batchnorm_forward(g["x"], g["gamma"], g["beta"], rm.cpu().numpy(), rv.cpu().numpy(), train=False,
                                   leakiness=0.333)[0]
    assert rel_err(y2.cpu().numpy(), want) < FP32_TOL


def test_io_layers_roundtrip_and_grads():
    coords, feats = scenes.make_batch("tiny", (3, 4))
    vox = rb.voxelize(coords, 2)
    N, P = len(vox["locs"]), len(coords)
    m, out = build_meta(coords, 2, feats, mode=4)
    assert out.shape == (N, 3)
    rng = np.random.default_rng(0)
    cnt = np.diff(vox["rule_ptr"]).astype(np.float32)
    # InputLayer backward: d_point = (1/n) * d_row
    gr = rng.standard_normal((N, 5)).astype(np.float32)
    gp = torch.empty(0, device="cuda")
    SCN.InputLayer_updateGradInput(m, gp, cu(gr))
    want = gr[vox["row_of_point"]] / cnt[vox["row_of_point"], None]
    assert gp.shape == (P, 5) and rel_err(gp.cpu().numpy(), want) < FP32_TOL
    # OutputLayer forward = copy of the voxel row; backward = sum over the voxel's points
    xr = rng.standard_normal((N, 8)).astype(np.float32)
    op = torch.empty(0, device="cuda")
    SCN.OutputLayer_updateOutput(m, cu(xr), op)
    assert np.array_equal(op.cpu().numpy(), xr[vox["row_of_point"]])
    gpo = rng.standard_normal((P, 8)).astype(np.float32)
    gi = torch.empty(0, device="cuda")
    SCN.OutputLayer_updateGradInput(m, gi, cu(gpo))
    want = np.zeros((N, 8), np.float32)
    np.add.at(want, vox["row_of_point"], gpo)
    assert rel_err(gi.cpu().numpy(), want) < FP32_TOL
    # mode 3 sums
    m3, out3 = build_meta(coords, 2, feats, mode=3)
    assert rel_err(out3.cpu().numpy(), rb.input_layer_mean(feats, vox, False)) < FP32_TOL


# ---------------------------------------------------------------------------------------- config 1 + properties
@pytest.mark.parametrize("precision,cin,cout", [("fp32", 16, 16), ("fp32", 3, 32), ("tf32", 32, 32), ("tf32", 64, 64),
                                                ("tf32", 128, 64), ("tf32", 64, 96), ("bf16", 64, 64), ("bf16", 128, 64),
                                                ("bf16", 64, 192), ("bf16", 320, 128), ("bf16", 64, 96)])
def test_submanifold_on_scene_vs_oracle(precision, cin, cout):
    """BASELINE.json config 1 shape (100k-voxel scene, 3x3x3, 16->16) plus the tensor-core channel shapes."""
    tol = FP32_TOL if precision == "fp32" else TF32_TOL
    coords, _ = scenes.make_batch("S100k", (0,))
    vox = rb.voxelize(coords, 1)
    rules = rb.submanifold_rules(vox["locs"], 1)
    N = len(vox["locs"])
    rng = np.random.default_rng(1)
    x = rng.standard_normal((N, cin)).astype(np.float32)
    w = (rng.standard_normal((27, cin, cout)) * (2.0 / cin / 27) ** 0.5).astype(np.float32)
    go = rng.standard_normal((N, cout)).astype(np.float32)
    scn.set_precision(precision)
    try:
        m, _ = build_meta(coords, 1)
        _lib.profile(True)
        y, macs, dx, dw = _subm(m, cu(x), cu(w), cu(go))
    finally:
        scn.set_precision("bf16")
    y0, macs0 = arith.rule_conv_forward(x, w, rules, N)
    dx0, dw0 = arith.rule_conv_backward(x, go, w, rules)
    assert macs == macs0
    prof = _lib.profile_read()
    _lib.profile(False)
    if precision != "fp32":      # the tcgen05 kernels really ran (forward + dgrad), no silent fp32 substitute
        assert prof["conv_tc"]["launches"] >= 2 and prof["conv_fp32"]["launches"] == 0, prof
        assert prof["wgrad_tc"]["launches"] >= 1 and prof["wgrad_fp32"]["launches"] == 0, prof
    else:
        assert prof["conv_tc"]["launches"] == 0 and prof["conv_fp32"]["launches"] >= 2, prof
        assert prof["wgrad_tc"]["launches"] == 0 and prof["wgrad_fp32"]["launches"] >= 1, prof
    assert rel_err(y.cpu().numpy(), y0) < tol
    assert rel_err(dx.cpu().numpy(), dx0) < tol
    assert rel_err(dw.cpu().numpy(), dw0) < tol


def test_full_size_properties():
    """BASELINE.json config 3 size (8 x S250k ~ 2 M voxels): properties that need no oracle run.
    linearity in the input, centre-only weights == x @ W[13], adjointness <conv(x),g> == <x,dgrad(g)> and
    == <W, wgrad> (the three passes describe one bilinear form)."""
    coords, _ = scenes.make_batch("S250k", tuple(range(8)))
    m, _ = build_meta(coords, 8)
    N = m.getNActive(lt(SIZE))
    assert N > 1_800_000
    C = 32
    gen = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(N, C, device="cuda", generator=gen)
    g = torch.randn(N, C, device="cuda", generator=gen)
    w = torch.randn(27, C, C, device="cuda", generator=gen) * 0.05
    for precision, tol, C in (("fp32", 1e-4, 32), ("tf32", TF32_TOL, 32), ("bf16", TF32_TOL, 64)):
        if x.shape[1] != C:
            x = torch.randn(N, C, device="cuda", generator=gen)
            g = torch.randn(N, C, device="cuda", generator=gen)
            w = torch.randn(27, C, C, device="cuda", generator=gen) * 0.05
        scn.set_precision(precision)
        try:
            y, macs, dx, dw = _subm(m, x, w, g)
            y2 = _subm(m, 2 * x, w, g)[0]
            assert rel_err(y2.cpu().numpy(), (2 * y).cpu().numpy()) < 1e-6          # exact: scaling by 2
            wc = torch.zeros_like(w)
            wc[13] = w[13]
            yc = _subm(m, x, wc, g)[0]
            assert rel_err(yc.cpu().numpy(), (x @ w[13]).cpu().numpy()) < max(tol, 2e-3)  # torch.mm may use tf32
            if precision == "bf16":
                assert rel_err(y2.cpu().numpy(), (2 * y).cpu().numpy()) == 0.0
            a = (y.double() * g.double()).sum().item()
            b = (x.double() * dx.double()).sum().item()
            c = (w.double() * dw.double()).sum().item()
            assert abs(a - b) / abs(a) < tol and abs(a - c) / abs(a) < tol
        finally:
            scn.set_precision("bf16")
    nbr, n_rules = m.submanifoldNeighbourTable(lt(SIZE))
    nbr = nbr.numpy()
    # symmetry of the rule relation: nbr[26-k][nbr[k][o]] == o
    for k in (0, 4, 13, 22):
        o = np.flatnonzero(nbr[k] >= 0)
        assert np.array_equal(nbr[26 - k][nbr[k][o]], o)
    assert n_rules == int((nbr >= 0).sum())


def test_bf16_operand_copies_are_equivalent():
    """precision 'bf16': the copy BatchNormReLU attaches to its output, and the copy a convolution keeps for its
    weight gradient, must give bit-identical results to the copies the library makes on its own."""
    coords, _ = scenes.make_batch("small", (0, 1))
    scn.set_precision("bf16")
    torch.manual_seed(3)
    inp = scn.InputLayer(3, SIZE, mode=4)
    x0 = [torch.from_numpy(coords).float(), torch.randn(len(coords), 64, device="cuda"), None, 2]
    bn = scn.BatchNormReLU(64).cuda()
    conv = scn.SubmanifoldConvolution(3, 64, 64, 3, False).cuda()
    t = bn(inp(x0))
    held = t.features._scn_bf16
    assert torch.equal(held[2], t.features.to(torch.bfloat16))           # round-to-nearest-even, like torch
    y_a = conv(t).features
    y_a.square().sum().backward()
    g_a = conv.weight.grad.clone()
    conv.weight.grad = None
    del t.features._scn_bf16                                             # library makes its own copies
    t2 = scn.SparseConvNetTensor(t.features.detach().clone().requires_grad_(True), t.metadata, t.spatial_size)
    y_b = conv(t2).features
    y_b.square().sum().backward()
    assert torch.equal(y_a, y_b) and torch.equal(g_a, conv.weight.grad)
    # an in-place change invalidates the attached copy (version counter)
    t3 = bn(inp(x0))
    with torch.no_grad():
        t3.features.mul_(2.0)
    y_c = conv(t3).features
    assert rel_err(y_c.detach().cpu().numpy(), (2 * y_a).detach().cpu().numpy()) < 1e-6


def _small_unet(m=64):
    torch.manual_seed(11)
    return scn.Sequential().add(scn.InputLayer(3, SIZE, mode=4)).add(scn.SubmanifoldConvolution(3, 3, m, 3, False)) \
        .add(scn.UNet(3, 1, [m, 2 * m, 3 * m], True)).add(scn.BatchNormReLU(m)).add(scn.OutputLayer(3)).cuda()


def _l2_err(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / np.linalg.norm(b))


def test_unet_tensor_core_precisions_track_the_fp32_path():
    """Three-level residual UNet (m=64), forward + backward: the bf16 and tf32 tensor-core paths against the exact
    fp32 path of this library (itself pinned to the reference at 1e-5 per layer).  Errors of ~30 layers compound and
    every rounding can flip a ReLU mask on this small scene, which makes the max-norm of single elements erratic, so
    the end-to-end check uses the relative L2 error: output 5e-2, gradient of the FIRST layer (it has crossed every
    layer twice) 2e-1 (measured: bf16 1.2e-1).  The per-layer budget stays 2e-2 max-norm (north_star) and is what the
    other tests assert."""
    coords, feats = scenes.make_batch("small", (0, 1))
    x = [torch.from_numpy(coords).float(), torch.from_numpy(feats).cuda(), None, 2]
    res = {}
    for precision in ("fp32", "tf32", "bf16"):
        scn.set_precision(precision)
        try:
            net = _small_unet()
            out = net(x)
            out.square().mean().backward()
            res[precision] = (out.detach().cpu().numpy(), net[1].weight.grad.detach().cpu().numpy())
        finally:
            scn.set_precision("bf16")
    for precision in ("tf32", "bf16"):
        assert _l2_err(res[precision][0], res["fp32"][0]) < 5e-2, precision
        assert _l2_err(res[precision][1], res["fp32"][1]) < 2e-1, precision


def test_prebuilt_scale_chain_equals_lazy_build():
    """Once the hierarchy depth is known the library builds every scale inside the InputLayer call; the rulebooks
    must be the ones the lazy path builds."""
    coords, _ = scenes.make_batch("small", (0, 1, 2))
    tabs = []
    for _ in range(2):                      # first handle may build lazily, second one is prebuilt
        m, _ = build_meta(coords, 3)
        gen = torch.Generator(device="cuda").manual_seed(5)
        x = torch.randn(m.getNActive(lt(SIZE)), 32, device="cuda", generator=gen)
        w8 = torch.randn(8, 32, 32, device="cuda", generator=gen)
        yc = torch.empty(0, device="cuda")
        SCN.Convolution_updateOutput(lt(SIZE), lt(SIZE // 2), lt(2), lt(2), m, x, yc, w8, torch.empty(0))
        yc2 = torch.empty(0, device="cuda")
        SCN.Convolution_updateOutput(lt(SIZE // 2), lt(SIZE // 4), lt(2), lt(2), m, yc, yc2, w8, torch.empty(0))
        nbr0, r0 = m.submanifoldNeighbourTable(lt(SIZE))
        nbr1, r1 = m.submanifoldNeighbourTable(lt(SIZE // 2))
        nbr2, r2 = m.submanifoldNeighbourTable(lt(SIZE // 4))
        tabs.append((nbr0.numpy(), nbr1.numpy(), nbr2.numpy(), r0, r1, r2, yc2.cpu().numpy()))
    for a, b in zip(tabs[0][:3], tabs[1][:3]):
        assert np.array_equal(a, b)
    assert tabs[0][3:6] == tabs[1][3:6]
    assert np.array_equal(tabs[0][6], tabs[1][6])


def test_fused_residual_block_equals_unfused():
    """The fusions inside scn.UNet against the plain composition of the same modules: ResidualConcatTable folds the
    shortcut add into the last convolution's epilogue and the shortcut's gradient into the first BatchNorm's backward
    kernel (the same single fp32 additions as the AddTable / autograd accumulation they replace), and the convolution
    epilogue hands its column statistics to the BatchNorm that follows (same sums, different summation order).
    The epilogue statistics differ from the reduction pass in the last bits (summation order: 5e-7 on one BatchNorm
    output, test_epilogue_statistics_and_residual_match_torch pins that).  With bf16 operand copies any such
    perturbation moves a few operand roundings (1 bf16 ulp = 4e-3) in the next layer, more in the one after, and within
    a few layers the two runs sit one bf16-rounding noise floor apart (measured 2.3e-3 relative L2 at the output), so
    the comparison is the relative L2 error at 1e-2 for the output and 5e-2 for the gradients.  Covers both shortcuts."""
    from occuseg_b200.sparseconvnet import SCN as scn_SCN
    coords, _ = scenes.make_batch("small", (0, 1))
    feats = torch.randn(len(coords), 3, device="cuda", generator=torch.Generator(device="cuda").manual_seed(2))
    x = [torch.from_numpy(coords).float(), feats, None, 2]
    res = []
    real, real_bn = scn_SCN.fuses_residual, scn_SCN.fuses_bn_conv
    for fused in (True, False):
        scn_SCN.fuses_residual = real if fused else (lambda a, b: False)
        scn_SCN.fuses_bn_conv = real_bn if fused else (lambda a, b: False)
        try:
            net = _small_unet()
            out = net(x)
            out.square().mean().backward()
            res.append((out.detach().clone(), [p.grad.detach().clone() for p in net.parameters()]))
        finally:
            scn_SCN.fuses_residual, scn_SCN.fuses_bn_conv = real, real_bn
    assert _l2_err(res[0][0].cpu().numpy(), res[1][0].cpu().numpy()) < 1e-2
    ga = torch.cat([g.flatten() for g in res[0][1]]).cpu().numpy()      # all parameter gradients as one vector
    gb = torch.cat([g.flatten() for g in res[1][1]]).cpu().numpy()
    assert _l2_err(ga, gb) < 5e-2


def test_epilogue_statistics_and_residual_match_torch():
    """scn_subm_fwd(residual, stats): out == conv(x) + residual bit for bit, stats == column sums / sums of squares."""
    coords, _ = scenes.make_batch("small", (0, 1))
    m, _ = build_meta(coords, 2)
    N = m.getNActive(lt(SIZE))
    gen = torch.Generator(device="cuda").manual_seed(7)
    for C in (64, 192, 320):                    # 320: two column tiles of 160
        x = torch.randn(N, C, device="cuda", generator=gen)
        w = torch.randn(27, C, C, device="cuda", generator=gen) * 0.05
        r = torch.randn(N, C, device="cuda", generator=gen)
        y0, y1 = torch.empty(0, device="cuda"), torch.empty(0, device="cuda")
        st = torch.empty(2, C, dtype=torch.float64, device="cuda")
        SCN.SubmanifoldConvolution_updateOutput(lt(SIZE), lt(3), m, x, y0, w, torch.empty(0), 1)
        SCN.SubmanifoldConvolution_updateOutput(lt(SIZE), lt(3), m, x, y1, w, torch.empty(0), 1, r, st)
        assert torch.equal(y1, y0 + r)
        s0, s1 = y1.double().sum(0), (y1.double() ** 2).sum(0)
        # fp32 partial sums per CTA (about 50 additions per column), merged in fp64: 5e-6
        assert float((st[0] - s0).abs().max() / s0.abs().max()) < 5e-6
        assert float((st[1] - s1).abs().max() / s1.abs().max()) < 5e-6


def test_runs_on_a_non_default_stream():
    """All work is enqueued on the caller's stream and a handle's buffers are freed in that stream's order: the same
    network on a side stream (handles created and destroyed there, no device-wide synchronisation in between) must
    reproduce the default-stream result."""
    coords, feats = scenes.make_batch("small", (0, 1))
    x = [torch.from_numpy(coords).float(), torch.from_numpy(feats).cuda(), None, 2]
    scn.set_precision("fp32")                       # deterministic forward: bit-exact comparison
    try:
        net = _small_unet()
        with torch.no_grad():
            ref = net(x).clone()
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        outs = []
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(4):                      # handles of earlier iterations die while later ones are enqueued
                outs.append(net(x))
        side.synchronize()
        for o in outs:
            assert torch.equal(o, ref)
    finally:
        scn.set_precision("bf16")


def test_inference_bn_relu_epilogue_is_bit_identical():
    """Inference: SubmanifoldConvolution followed by BatchNormReLU runs as one kernel (BatchNorm+ReLU in the convolution
    epilogue, north_star item 3).  It evaluates the same fp32 expressions as the two separate kernels, so the whole
    network output must be bit-identical to the unfused evaluation, in every tensor-core precision."""
    from occuseg_b200.sparseconvnet import layers as L
    coords, feats = scenes.make_batch("small", (0, 1))
    x = [torch.from_numpy(coords).float(), torch.from_numpy(feats).cuda(), None, 2]
    real = L._fusable_inference_pair
    for precision in ("bf16", "tf32"):
        scn.set_precision(precision)
        try:
            net = _small_unet()
            with torch.no_grad():
                net(x)                                   # one training-mode pass so the running statistics are not trivial
            net.eval()
            calls = []
            L._fusable_inference_pair = lambda a, b, i: (calls.append(1) or True) if real(a, b, i) else False
            with torch.no_grad():
                fused = net(x).clone()
            assert len(calls) >= 5                        # one conv+BN pair per residual block
            L._fusable_inference_pair = lambda a, b, i: False
            with torch.no_grad():
                plain = net(x).clone()
            assert torch.equal(fused, plain)
        finally:
            L._fusable_inference_pair = real
            scn.set_precision("bf16")


def test_occuseg_networks_run_end_to_end():
    """LearningBWDenseUNet (sparse backbone + the seven dense heads) forward and backward on the CUDA path."""
    from occuseg_b200 import models
    coords, feats = scenes.make_batch("small", (0, 1))
    net = models.LearningBWDenseUNet(models.default_config(m=64, levels=3)).cuda()
    outs = net([torch.from_numpy(coords).float(), torch.from_numpy(feats).cuda(), None, 2])
    P = len(coords)
    assert [tuple(o.shape) for o in outs] == [(P, 20), (P, 64), (P, 64), (P, 1), (P, 3), (P, 2), (P, 1)]
    sum(o.square().mean() for o in outs).backward()
    assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in net.parameters())


@pytest.mark.parametrize("precision,cin,cout", [("bf16", 64, 64), ("bf16", 128, 64), ("tf32", 32, 96)])
def test_pattern_sorted_tile_order_is_bit_identical(precision, cin, cout):
    """The tensor-core kernels walk every level in a tile order sorted by neighbourhood pattern (scn_tile_sort) and write
    their rows back through the permutation.  Every output row still accumulates its taps in the same order, so forward
    and dgrad results must be bit-identical to the natural row order, for any block size."""
    coords, _ = scenes.make_batch("S100k", (2,))
    gen = torch.Generator(device="cuda").manual_seed(9)
    res = []
    scn.set_precision(precision)
    prev = _lib.tile_sort(0)
    try:
        for block in (0, 1000, 32768, 1 << 30):
            _lib.tile_sort(block)
            m, _ = build_meta(coords, 1)
            N = m.getNActive(lt(SIZE))
            if not res:
                x = torch.randn(N, cin, device="cuda", generator=gen)
                w = torch.randn(27, cin, cout, device="cuda", generator=gen) * 0.05
                g = torch.randn(N, cout, device="cuda", generator=gen)
                r = torch.randn(N, cout, device="cuda", generator=gen)
            y = torch.empty(0, device="cuda")
            st = torch.empty(2, cout, dtype=torch.float64, device="cuda")
            SCN.SubmanifoldConvolution_updateOutput(lt(SIZE), lt(3), m, x, y, w, torch.empty(0), 1, r, st)
            dx, dw = torch.empty(0, device="cuda"), torch.zeros_like(w)
            SCN.SubmanifoldConvolution_backward(lt(SIZE), lt(3), m, x, dx, g, w, dw, torch.empty(0), 1)
            res.append((y, dx, st))
    finally:
        _lib.tile_sort(prev)
    for y, dx, st in res[1:]:
        assert torch.equal(y, res[0][0]) and torch.equal(dx, res[0][1])
        assert float((st - res[0][2]).abs().max() / res[0][2].abs().max()) < 5e-6      # fp32 partial sums regroup


@pytest.mark.parametrize("kind", ["subm", "conv", "deconv"])
def test_fused_batchnorm_conv_pair_equals_the_two_layers(kind):
    """Training-mode BatchNormReLU -> convolution as one autograd node (bf16-only activation, BatchNorm backward sums in
    the dgrad epilogue, functions.BatchNormConvFunction) against the same two modules run separately.  The forward
    arithmetic is identical (same bf16 operand, same mask expression): outputs must be bit-identical; the backward differs
    only in the fp32 grouping of the two BatchNorm column sums: 1e-5."""
    from occuseg_b200.sparseconvnet import SCN as scn_SCN
    coords, _ = scenes.make_batch("small", (0, 1))
    torch.manual_seed(21)
    cin, cout = (128, 64) if kind != "conv" else (64, 128)
    inp = scn.InputLayer(3, SIZE, mode=4)
    pre = scn.Convolution(3, 64, 64, 2, 2, False).cuda()           # creates the coarse scale the deconvolution needs
    bn = scn.BatchNormLeakyReLU(cin, leakiness=0.2).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5)
        bn.bias.normal_()
    conv = {"subm": lambda: scn.SubmanifoldConvolution(3, cin, cout, 3, False),
            "conv": lambda: scn.Convolution(3, cin, cout, 2, 2, False),
            "deconv": lambda: scn.Deconvolution(3, cin, cout, 2, 2, False)}[kind]().cuda()
    seq = scn.Sequential().add(bn).add(conv)
    feats0 = torch.randn(len(coords), 64, device="cuda")
    real = scn_SCN.fuses_bn_conv
    res = []
    for fused in (True, False):
        scn_SCN.fuses_bn_conv = real if fused else (lambda a, b: False)
        try:
            t0 = inp([torch.from_numpy(coords).float(), feats0, None, 2])
            tc = pre(t0)
            base = tc if kind == "deconv" else t0
            n = base.features.size(0)
            gen = torch.Generator(device="cuda").manual_seed(5)
            xin = torch.randn(n, cin, device="cuda", generator=gen).requires_grad_(True)
            bn.running_mean.zero_(); bn.running_var.fill_(1.0)
            for p_ in list(bn.parameters()) + list(conv.parameters()):
                p_.grad = None
            out = seq(scn.SparseConvNetTensor(xin, base.metadata, base.spatial_size))
            g = torch.randn(out.features.shape, device="cuda", generator=gen)
            out.features.backward(g)
            res.append((out.features.detach().clone(), xin.grad.clone(), bn.weight.grad.clone(), bn.bias.grad.clone(),
                        conv.weight.grad.clone(), bn.running_mean.clone(), bn.running_var.clone()))
        finally:
            scn_SCN.fuses_bn_conv = real
    a, b = res
    assert torch.equal(a[0], b[0]) and torch.equal(a[5], b[5]) and torch.equal(a[6], b[6])
    for i, name in ((1, "d_x"), (2, "d_gamma"), (3, "d_beta"), (4, "d_weight")):
        assert rel_err(a[i].cpu().numpy(), b[i].cpu().numpy()) < 1e-5, (kind, name)


def test_bf16_gradient_copies_replace_the_cast_passes(monkeypatch):
    """The BatchNorm backward of a fused BatchNorm -> convolution node leaves a bf16 copy of its d_x; the convolution that
    receives that gradient as d_out reads the copy (scn_grad_bf16) instead of casting.  Same bits as the cast (both round to
    nearest even), fewer cast launches."""
    coords, feats = scenes.make_batch("small", (5, 6))

    def step(use_copies):
        if not use_copies:
            monkeypatch.setattr(SCN, "held_bf16", lambda t: None)
        torch.manual_seed(3)
        net = scn.Sequential().add(scn.InputLayer(3, SIZE, mode=4)).add(scn.SubmanifoldConvolution(3, 3, 64, 3, False)) \
            .add(scn.UNet(3, 2, [64, 128], True)).add(scn.BatchNormReLU(64)).add(scn.OutputLayer(3)).cuda()
        _lib.profile(True)
        out = net([torch.from_numpy(coords), cu(feats), None, 2])
        out.square().mean().backward()
        torch.cuda.synchronize()
        prof = _lib.profile_read()
        _lib.profile(False)
        monkeypatch.undo()
        return [p.grad.clone() for p in net.parameters()], prof["cast"]["launches"]

    g1, casts1 = step(True)
    g0, casts0 = step(False)
    assert casts1 < casts0, (casts1, casts0)
    # the weight gradients merge partial sums with floating-point atomics (run-to-run order), so equality is up to that noise
    for a, b in zip(g1, g0):
        assert rel_err(a.cpu().numpy(), b.cpu().numpy()) < 1e-5
