"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol include/scn_b200.h
declares (no compute without a GPU), the ctypes table agrees with the header, the Python surface mirrors the
reference's class names / parameter shapes, and the product path fails loudly without CUDA."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "scn_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(scn_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from occuseg_b200 import _lib
    from occuseg_b200.csrc import build
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = header_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/scn_b200.h but not exported"
    assert set(names) == set(_lib.PROTOTYPES), set(names) ^ set(_lib.PROTOTYPES)
    assert _lib.lib().scn_version() == 100


def test_product_does_not_import_the_oracle():
    """The oracle is test infrastructure: nothing under occuseg_b200/ may reference it."""
    for dp, _, files in os.walk(os.path.join(ROOT, "occuseg_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dp, f)).read()
                assert "oracle" not in text.replace("oracle/", "").replace("the oracle", "") or f == "__init__.py", (dp, f)
                assert "import oracle" not in text and "from oracle" not in text, (dp, f)


def test_no_cpu_fallback():
    import occuseg_b200.sparseconvnet as scn
    bn = scn.BatchNormReLU(8)
    t = scn.SparseConvNetTensor(torch.zeros(4, 8), None, None)
    with pytest.raises(TypeError, match="no CPU path"):
        bn(t)


def test_python_surface_matches_reference_names():
    import occuseg_b200.sparseconvnet as scn
    for name in ["InputLayer", "OutputLayer", "SubmanifoldConvolution", "ValidConvolution", "Convolution",
                 "Deconvolution", "BatchNormalization", "BatchNormReLU", "BatchNormLeakyReLU", "NetworkInNetwork",
                 "Sequential", "ConcatTable", "AddTable", "JoinTable", "Identity", "UNet", "Metadata",
                 "SparseConvNetTensor"]:
        assert hasattr(scn, name), name
    for fn in ["InputLayer_updateOutput", "InputLayer_updateGradInput", "OutputLayer_updateOutput",
               "OutputLayer_updateGradInput", "SubmanifoldConvolution_updateOutput", "SubmanifoldConvolution_backward",
               "Convolution_updateOutput", "Convolution_backward", "Deconvolution_updateOutput",
               "Deconvolution_backward", "BatchNormalization_updateOutput", "BatchNormalization_backward",
               "NetworkInNetwork_updateOutput", "NetworkInNetwork_updateGradInput",
               "NetworkInNetwork_accGradParameters", "Metadata_3"]:
        assert hasattr(scn.SCN, fn), fn
    c = scn.SubmanifoldConvolution(3, 16, 32, 3, False)
    assert tuple(c.weight.shape) == (27, 16, 32) and not hasattr(c, "bias")
    d = scn.Convolution(3, 16, 32, 2, 2, True)
    assert tuple(d.weight.shape) == (8, 16, 32) and tuple(d.bias.shape) == (32,)
    bn = scn.BatchNormReLU(16)
    assert set(bn.state_dict()) == {"weight", "bias", "running_mean", "running_var"}
    assert bn.eps == 1e-4 and bn.momentum == 0.9 and bn.leakiness == 0


def test_unet_state_dict_layout():
    """Key layout of scn.UNet(3, 1, [m..6m], residual) -- the checkpoint format of baseline_m64."""
    import occuseg_b200.sparseconvnet as scn
    u = scn.UNet(3, 1, [64, 128, 192, 256, 320, 384], True)
    sd = u.state_dict()
    assert len(sd) == 165 and sum(v.numel() for k, v in sd.items() if "running" not in k) == 43429120
    assert tuple(sd["0.1.1.weight"].shape) == (27, 64, 64)          # first block, first SubmConv
    assert tuple(sd["2.1.1.weight"].shape) == (8, 64, 128)          # down-convolution 64->128
    assert tuple(sd["2.1.4.weight"].shape) == (8, 128, 64)          # matching deconvolution
    assert tuple(sd["4.0.weight"].shape) == (128, 64)               # NiN shortcut of the 2c->c block
    ref_pkg = "/root/reference/sparseconvnet"
    if os.path.isdir(ref_pkg):                                      # authoring container: compare with the real thing
        import importlib
        import sys
        import types
        stub = types.ModuleType("sparseconvnet.SCN")
        for d in range(1, 7):
            setattr(stub, f"Metadata_{d}", object)
        saved = {k: sys.modules.get(k) for k in ("sparseconvnet", "sparseconvnet.SCN")}
        sys.path.insert(0, "/root/reference")
        sys.modules["sparseconvnet.SCN"] = stub
        try:
            sys.modules.pop("sparseconvnet", None)
            ref = importlib.import_module("sparseconvnet")
            r = ref.UNet(3, 1, [64, 128, 192, 256, 320, 384], True)
            assert {k: tuple(v.shape) for k, v in r.state_dict().items()} == {k: tuple(v.shape) for k, v in sd.items()}
        finally:
            sys.path.remove("/root/reference")
            for k in [k for k in sys.modules if k == "sparseconvnet" or k.startswith("sparseconvnet.")]:
                del sys.modules[k]
            for k, v in saved.items():
                if v is not None:
                    sys.modules[k] = v


def test_scene_generator_is_seeded_and_shaped():
    import numpy as np
    from occuseg_b200 import scenes
    a, fa = scenes.make_scene("small", 3)
    b, fb = scenes.make_scene("small", 3)
    assert np.array_equal(a, b) and np.array_equal(fa, fb)
    assert a.min() == 10 and fa.dtype == np.float32 and fa.shape == (len(a), 3)
    c, f = scenes.make_batch("tiny", (0, 1, 2))
    assert c.shape[1] == 4 and np.all(np.diff(c[:, 3]) >= 0) and c[:, 3].max() == 2


def test_occuseg_model_heads_follow_the_reference_names():
    """occuseg_b200.models mirrors examples/ScanNet/model.py:657-717: same attribute names and head shapes, so the
    tensors of a reference checkpoint load by name (construction only: no CUDA needed)."""
    from occuseg_b200 import models
    cfg = models.default_config(m=32, levels=3, class_num=20)
    net = models.LearningBWDenseUNet(cfg)
    sd = net.state_dict()
    for name, shape in {"backbone.linear.weight": (20, 32), "backbone.fc_regress.weight": (32, 32),
                        "backbone.linear_regress.weight": (1, 32), "backbone.fc_embedding.weight": (32, 32),
                        "backbone.linear_embedding.weight": (32, 32), "backbone.fc_displacement.weight": (32, 32),
                        "backbone.linear_displacement.weight": (3, 32), "fc_bw.weight": (32, 32), "linear_bw.weight": (2, 32),
                        "fc_occupancy.weight": (32, 32), "linear_occupancy.weight": (1, 32),
                        "backbone.sub.weight": (27, 3, 32), "backbone.bn.running_mean": (32,)}.items():
        assert tuple(sd[name].shape) == shape, name
    assert any(k.startswith("backbone.unet.") for k in sd)
