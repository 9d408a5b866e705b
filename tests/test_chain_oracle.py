"""CPU tests of the network-level oracle (oracle/chain.py) and of drop-in compatibility with the reference's OWN
Python code.  Both need the reference tree (/root/reference) and therefore run in the authoring container only; on
the GPU box they skip (the GPU parity tests use oracle/chain.py, validated here)."""
import os
import sys

import numpy as np
import pytest
import torch

from conftest import rel_err

REF_ROOT = "/root/reference"
needs_ref = pytest.mark.skipif(not os.path.isdir(os.path.join(REF_ROOT, "sparseconvnet")),
                               reason="reference tree not present (GPU box)")


def _restore_modules(saved):
    for k in [k for k in sys.modules if k == "sparseconvnet" or k.startswith("sparseconvnet.")]:
        del sys.modules[k]
    sys.modules.update(saved)


@pytest.fixture
def ref_pkg():
    """the reference's own `sparseconvnet` Python package imported over the CPU stand-in for its pybind module"""
    from oracle import ref_scn, reference
    if not reference.available():
        pytest.skip("oracle/_ref not available")
    saved = {k: v for k, v in sys.modules.items() if k == "sparseconvnet" or k.startswith("sparseconvnet.")}
    pkg = ref_scn.import_reference_package(REF_ROOT)
    yield pkg
    _restore_modules(saved)


def _net(scn, planes, m_in=3):
    return scn.Sequential().add(scn.InputLayer(3, 4096, mode=4)).add(scn.SubmanifoldConvolution(3, m_in, planes[0], 3, False)) \
        .add(scn.UNet(3, 1, planes, True)).add(scn.BatchNormReLU(planes[0])).add(scn.OutputLayer(3))


@needs_ref
def test_chain_oracle_equals_the_reference_python_package(ref_pkg):
    """The reference's sparseconvnet/*.py (networkArchitectures.UNet, its autograd Functions, tables, ...) run
    unmodified on the reference's compiled CPU arithmetic; oracle/chain.py walking THIS repo's module tree with the
    same weights must reproduce its output and every parameter gradient.  Also proves state_dict compatibility:
    the reference network's state_dict loads strictly into the mirror built from occuseg_b200.sparseconvnet."""
    import occuseg_b200.sparseconvnet as ours_scn
    from occuseg_b200 import scenes
    from oracle import chain
    coords, feats = scenes.make_batch("tiny", (0, 1))
    x = [torch.from_numpy(coords).float(), torch.from_numpy(feats), None, 2]
    planes = [16, 32, 48]
    torch.manual_seed(5)
    ref_net = _net(ref_pkg, planes)
    ours = _net(ours_scn, planes)
    assert list(ours.state_dict().keys()) == list(ref_net.state_dict().keys())
    ours.load_state_dict(ref_net.state_dict(), strict=True)
    for (ka, va), (kb, vb) in zip(ours.state_dict().items(), ref_net.state_dict().items()):
        assert va.shape == vb.shape, ka

    out_ref = ref_net(x)
    out_ref.square().mean().backward()
    rp, out = chain.replay(ours, x)
    out.square().mean().backward()
    assert rel_err(out.detach().numpy(), out_ref.detach().numpy()) < 1e-6
    ref_params = dict(ref_net.named_parameters())
    n = 0
    for name, p in ours.named_parameters():
        assert rel_err(rp.grad_of(p).numpy(), ref_params[name].grad.numpy()) < 1e-5, name
        n += 1
    assert n == len(ref_params) and n > 30
    # running statistics were updated identically
    for (name, b), (_, rb_) in zip(ours.named_buffers(), ref_net.named_buffers()):
        assert rel_err(rp.params[id(b)].numpy(), rb_.numpy()) < 1e-6, name
    kinds = {r["kind"] for r in rp.tape}
    assert kinds == {"input", "subm", "bn", "conv", "deconv", "nin", "output"}


@needs_ref
def test_reference_model_py_constructs_on_this_package():
    """examples/ScanNet/model.py of the reference, UNMODIFIED, imported with `sparseconvnet` resolving to this package
    (occuseg_b200.install_as_sparseconvnet): LearningBWDenseUNet / InstanceDenseUNet construct, and their state_dict
    (names, shapes) is the one occuseg_b200.models ships -- so reference checkpoints load either way.  tensorboardX and
    torch_scatter (absent here, used only by the training scripts / ClusterSegNet) are stubbed."""
    import importlib.util
    import types
    import occuseg_b200
    from occuseg_b200 import models
    saved = {k: v for k, v in sys.modules.items() if k == "sparseconvnet" or k.startswith("sparseconvnet.")}
    stubs = {}
    try:
        occuseg_b200.install_as_sparseconvnet()
        for name, attrs in (("tensorboardX", ["SummaryWriter"]),
                            ("torch_scatter", ["scatter_max", "scatter_mean", "scatter_std", "scatter_sub", "scatter_min",
                                               "scatter_add", "scatter_div"])):
            if name not in sys.modules:
                mod = types.ModuleType(name)
                for a in attrs:
                    setattr(mod, a, None)
                sys.modules[name] = stubs[name] = mod
        spec = importlib.util.spec_from_file_location("occuseg_ref_model", os.path.join(REF_ROOT, "examples", "ScanNet", "model.py"))
        ref_model = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref_model)
        cfg = models.default_config(m=32, levels=4)
        torch.manual_seed(0)
        ref_net = ref_model.LearningBWDenseUNet(cfg)
        mine = models.LearningBWDenseUNet(cfg)
        sd_ref, sd_mine = ref_net.state_dict(), mine.state_dict()
        assert list(sd_ref.keys()) == list(sd_mine.keys())
        assert all(sd_ref[k].shape == sd_mine[k].shape for k in sd_ref)
        mine.load_state_dict(sd_ref, strict=True)
        # the sparse layers the reference file instantiated ARE this package's
        import occuseg_b200.sparseconvnet as scn
        assert isinstance(ref_net.backbone.sub, scn.SubmanifoldConvolution)
        assert isinstance(ref_net.backbone.input, scn.InputLayer) and ref_net.backbone.input.mode == 4
        assert ref_net.backbone.input.normal_guide_scale == (cfg["full_scale"] >> cfg["rotation_guide_level"]) + 1
    finally:
        for name in stubs:
            sys.modules.pop(name, None)
        _restore_modules(saved)
