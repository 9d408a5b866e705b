"""TEST INFRASTRUCTURE ONLY -- a CPU stand-in for the reference's pybind11 module `sparseconvnet.SCN`
(sparseconvnet/SCN/pybind.cpp:11-239), for the entry points the OccuSeg UNet calls.

Same names, same argument order as the reference's C++ entries (sparseconvnet/SCN/sparseconvnet.h), so that
  * the REFERENCE'S OWN PYTHON PACKAGE (/root/reference/sparseconvnet/*.py, unmodified) runs on top of it
    (`install()` below puts it at sys.modules['sparseconvnet.SCN'] before that package is imported), and
  * oracle/chain.py can replay a whole network layer by layer.
Arithmetic = the reference's CPU code compiled unmodified (oracle/_ref/scn_cpu_ref.so: CPU/Convolution.cpp,
Deconvolution.cpp, BatchNormalization.cpp, NetworkInNetwork.cpp).  Rulebooks = oracle/rulebook.py, i.e. the GPU
builders' convention, pinned to the reference's compiled CPU builders (tests/test_oracle.py).  InputLayer /
OutputLayer arithmetic follows CUDA/IOLayers.cu:16-75 (restated in oracle/rulebook.py:input_layer_mean).
"""
from __future__ import annotations

import sys
import types

import numpy as np
import torch

from . import reference, rulebook as rb


def _size(t):
    return int(t.reshape(-1)[0].item()) if torch.is_tensor(t) else int(t)


class Metadata_3:
    """Replaces Metadata<3> for the CPU replay: holds the voxelisation of one batch and the rulebooks of its scales."""

    def __init__(self):
        self.ref = reference.Ref()
        self.vox = None
        self.batch = 0
        self.mode = 4
        self.locs = {}            # spatial size -> int64 [N,4]
        self.subm = set()
        self.strided = set()
        self.normal_guide_scale = None

    def setNormalGuideScale(self, v):
        self.normal_guide_scale = v

    def getNActive(self, spatial_size):
        return len(self.locs[_size(spatial_size)])

    def getSpatialLocations(self, spatial_size):
        return torch.from_numpy(self.locs[_size(spatial_size)].copy())

    def clear(self):
        pass

    # -- rulebooks on demand, as Metadata::getSubmanifoldRuleBook / getRuleBook cache them (Metadata.cpp:503-529,597-625)
    def need_subm(self, size):
        if size not in self.subm:
            self.ref.load_submanifold(size, rb.submanifold_rules(self.locs[size], self.batch), len(self.locs[size]))
            self.subm.add(size)

    def need_strided(self, fine, coarse):
        if fine not in self.strided:
            clocs, rules = rb.strided_rules(self.locs[fine], self.batch)
            assert coarse not in self.locs or np.array_equal(self.locs[coarse], clocs)
            self.locs[coarse] = clocs
            self.ref.load_strided(fine, coarse, rules, len(self.locs[fine]), len(clocs))
            self.strided.add(fine)


def _fill(dst, src):
    src = torch.as_tensor(src)
    dst.resize_(src.shape)
    dst.copy_(src)


def InputLayer_updateOutput(m, spatial_size, input_coords, input_features, output_features, batch_size, mode,
                            input_normal=None):
    assert mode in (3, 4), "only modes 3/4 exist on the reference GPU path (IOLayersRules.h:143)"
    coords = input_coords.numpy().astype(np.int64)
    m.batch, m.mode = int(batch_size), int(mode)
    m.vox = rb.voxelize(coords, m.batch)
    m.input_size = _size(spatial_size)
    m.locs[m.input_size] = m.vox["locs"]
    _fill(output_features, rb.input_layer_mean(input_features.detach().numpy(), m.vox, mode == 4))


def InputLayer_updateGradInput(m, d_input_features, d_output_features):
    """CUDA/IOLayers.cu:45-58 with the roles swapped: d_point = mult * d_row (mult = 1/n in mode 4)."""
    g = d_output_features.numpy()
    rows = m.vox["row_of_point"]
    cnt = np.diff(m.vox["rule_ptr"]).astype(np.float32)
    mult = (np.float32(1) / cnt) if m.mode == 4 else np.ones_like(cnt)
    _fill(d_input_features, g[rows] * mult[rows, None])


def OutputLayer_updateOutput(m, input_features, output_features):
    _fill(output_features, input_features.detach().numpy()[m.vox["row_of_point"]])


def OutputLayer_updateGradInput(m, d_input_features, d_output_features):
    g = d_output_features.numpy()
    out = np.zeros((len(m.vox["locs"]), g.shape[1]), np.float32)
    np.add.at(out, m.vox["row_of_point"], g)
    _fill(d_input_features, out)


def _l(t):
    """size tensors as LongTensor: under torch >= 1.5 the reference's `(size - f) / s + 1` (convolution.py:35) is a float tensor"""
    return t.long() if torch.is_tensor(t) else torch.LongTensor([int(t)] * 3)


def _e():
    return torch.empty(0)


def _opt(t):
    return t if (t is not None and t.numel()) else _e()


def SubmanifoldConvolution_updateOutput(spatial_size, filter_size, m, input_features, output_features, weight, bias,
                                        dilated_rate=1):
    assert int(dilated_rate) == 1
    m.need_subm(_size(spatial_size))
    return reference.module().SubmanifoldConvolution_updateOutput(_l(spatial_size), _l(filter_size), m.ref.m,
                                                                  input_features.detach().contiguous(), output_features,
                                                                  weight.detach(), _opt(bias))


def SubmanifoldConvolution_backward(spatial_size, filter_size, m, input_features, d_input_features, d_output_features,
                                    weight, d_weight, d_bias, dilated_rate=1):
    m.need_subm(_size(spatial_size))
    reference.module().SubmanifoldConvolution_backward(_l(spatial_size), _l(filter_size), m.ref.m, input_features.detach().contiguous(),
                                                       d_input_features, d_output_features.contiguous(), weight.detach(),
                                                       d_weight, _opt(d_bias))


def Convolution_updateOutput(in_size, out_size, filter_size, filter_stride, m, input_features, output_features, weight,
                             bias):
    m.need_strided(_size(in_size), _size(out_size))
    return reference.module().Convolution_updateOutput(_l(in_size), _l(out_size), _l(filter_size), _l(filter_stride), m.ref.m,
                                                       input_features.detach().contiguous(), output_features,
                                                       weight.detach(), _opt(bias))


def Convolution_backward(in_size, out_size, filter_size, filter_stride, m, input_features, d_input_features,
                         d_output_features, weight, d_weight, d_bias):
    reference.module().Convolution_backward(_l(in_size), _l(out_size), _l(filter_size), _l(filter_stride), m.ref.m,
                                            input_features.detach().contiguous(), d_input_features,
                                            d_output_features.contiguous(), weight.detach(), d_weight, _opt(d_bias))


def Deconvolution_updateOutput(in_size, out_size, filter_size, filter_stride, m, input_features, output_features,
                               weight, bias):
    # the reference keys the shared rulebook by the FINE size (CUDA/Deconvolution.cpp:28-29)
    return reference.module().Deconvolution_updateOutput(_l(in_size), _l(out_size), _l(filter_size), _l(filter_stride), m.ref.m,
                                                         input_features.detach().contiguous(), output_features,
                                                         weight.detach(), _opt(bias))


def Deconvolution_backward(in_size, out_size, filter_size, filter_stride, m, input_features, d_input_features,
                           d_output_features, weight, d_weight, d_bias):
    reference.module().Deconvolution_backward(_l(in_size), _l(out_size), _l(filter_size), _l(filter_stride), m.ref.m,
                                              input_features.detach().contiguous(), d_input_features,
                                              d_output_features.contiguous(), weight.detach(), d_weight, _opt(d_bias))


def BatchNormalization_updateOutput(input_features, output_features, saveMean, saveInvStd, runningMean, runningVar,
                                    weight, bias, eps, momentum, train, leakiness):
    reference.module().BatchNormalization_updateOutput(input_features.detach().contiguous(), output_features, saveMean,
                                                       saveInvStd, runningMean, runningVar, _opt(weight).detach(),
                                                       _opt(bias).detach(), eps, momentum, train, leakiness)


def BatchNormalization_backward(input_features, d_input_features, output_features, d_output_features, saveMean,
                                saveInvStd, runningMean, runningVar, weight, bias, d_weight, d_bias, leakiness):
    # the reference masks d_output in place (CPU/BatchNormalization.cpp:76-80); autograd owns that tensor, so it gets a copy
    reference.module().BatchNormalization_backward(input_features.detach().contiguous(), d_input_features,
                                                   output_features.detach().contiguous(), d_output_features.clone(),
                                                   saveMean, saveInvStd, runningMean, runningVar, _opt(weight).detach(),
                                                   _opt(bias).detach(), _opt(d_weight), _opt(d_bias), leakiness)


def NetworkInNetwork_updateOutput(input_features, output_features, weight, bias):
    return reference.module().NetworkInNetwork_updateOutput(input_features.detach().contiguous(), output_features,
                                                            weight.detach(), _opt(bias))


def NetworkInNetwork_updateGradInput(d_input_features, d_output_features, weight):
    reference.module().NetworkInNetwork_updateGradInput(d_input_features, d_output_features.contiguous(), weight.detach())


def NetworkInNetwork_accGradParameters(input_features, d_output_features, d_weight, d_bias):
    reference.module().NetworkInNetwork_accGradParameters(input_features.detach().contiguous(),
                                                          d_output_features.contiguous(), d_weight, _opt(d_bias))


def n_rulebook_bits():
    return 32


def as_module():
    """This file's entry points packaged as a module object named `sparseconvnet.SCN`."""
    mod = types.ModuleType("sparseconvnet.SCN")
    for k, v in globals().items():
        if k[0].isupper() or k == "n_rulebook_bits":
            setattr(mod, k, v)
    return mod


def import_reference_package(ref_root="/root/reference"):
    """Import the reference's own Python package `sparseconvnet` (unmodified, from ref_root) on top of this stand-in.
    Returns the package, or None when the tree is absent (GPU box).  Leaves sys.modules['sparseconvnet'] pointing at
    the reference package; callers that also use occuseg_b200.install_as_sparseconvnet() must restore it."""
    import os
    if not os.path.isdir(os.path.join(ref_root, "sparseconvnet")) or not reference.available():
        return None
    for k in [k for k in sys.modules if k == "sparseconvnet" or k.startswith("sparseconvnet.")]:
        del sys.modules[k]
    sys.modules["sparseconvnet.SCN"] = as_module()
    sys.path.insert(0, ref_root)
    try:
        import sparseconvnet  # noqa: F401  (the reference's)
        pkg = sys.modules["sparseconvnet"]
        pkg.SCN = sys.modules["sparseconvnet.SCN"]
    finally:
        sys.path.remove(ref_root)
    return pkg
