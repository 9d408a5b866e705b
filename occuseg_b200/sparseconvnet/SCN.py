"""Host-side mirror of the reference's pybind11 module `sparseconvnet.SCN` for the hot path
(reference: sparseconvnet/SCN/pybind.cpp:11-239, sparseconvnet/SCN/sparseconvnet.h).

Same names, same argument order and meaning, same ownership rule (the caller passes EMPTY output
tensors, the callee resizes and fills them; sparseconvnet/SCN/CUDA/Convolution.cpp:122), but every
call goes straight to libscn_b200.so through the C ABI in include/scn_b200.h with raw device pointers
and the current torch stream.  Errors are Python exceptions instead of exit()/abort().
There is no CPU path: tensors must be CUDA float32.
"""
from __future__ import annotations

import ctypes as C
import os

import torch

from .. import _lib

_PRECISIONS = {"fp32": _lib.FP32, "tf32": _lib.TF32, "bf16": _lib.BF16}
# The package mirrors an fp32 reference, so the DEFAULT is the exact fp32 path (rel 1e-5 per layer vs the reference's CPU
# code).  The tensor-core modes are an explicit opt-in (set_precision / SCN_B200_PRECISION) with a stated budget of
# rel 2e-2 per layer (measured ~3e-3 for bf16 operands, ~5e-4 for tf32, fp32 accumulation in both); README.md.
_precision = _PRECISIONS[os.environ.get("SCN_B200_PRECISION", "fp32").lower()]


def set_deterministic(on=True):
    """Run-to-run bit-identical results: weight gradients and column statistics are summed in a fixed order instead of with
    floating-point atomics (off by default, as in the reference; also SCN_DETERMINISTIC=1).  Returns the previous setting."""
    return _lib.deterministic(on)


def set_precision(name):
    """'fp32' (default) = exact FMA path everywhere; 'tf32' / 'bf16' = tcgen05 tiles (tf32 operands / bf16 copies of
    the operands, fp32 accumulate and fp32 results) where the channel counts allow, exact fp32 elsewhere."""
    global _precision
    _precision = _PRECISIONS[name.lower()]


def get_precision():
    return {v: k for k, v in _PRECISIONS.items()}[_precision]


def _stream():
    # raw handle of the current stream of the current device (torch.cuda.current_stream() builds a Stream object: ~15 us)
    return C.c_void_p(torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice()))


class _NoSwitch:
    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False


_NO_SWITCH = _NoSwitch()


def _on(device):
    """`with _on(device)`, skipped when that device is already current (the usual case: one process per GPU)"""
    if device.index is None or device.index == torch._C._cuda_getDevice():
        return _NO_SWITCH
    return torch.cuda.device(device)


# ---- bf16 operand copies (precision 'bf16') ---------------------------------------------------------------------
# A producer (BatchNormReLU) may attach the bf16 copy of its output to the tensor; a convolution that finds a valid
# copy (same storage, unchanged version counter) hands it to the library instead of paying for a cast pass, and
# keeps it for the weight gradient.  Purely an optimisation: without a copy the library makes its own.
def fuses_residual(c_in, c_out):
    """True when SubmanifoldConvolution_updateOutput(nIn=c_in, nOut=c_out) can add a residual in its epilogue."""
    return bool(_lib.lib().scn_fuses_residual(int(c_in), int(c_out), _precision))


def wants_bf16(channels):
    return _precision == _lib.BF16 and channels % 64 == 0


def attach_bf16(t, copy):
    t._scn_bf16 = (t._version, t.data_ptr(), copy)


def attach_stats(t, stats):
    t._scn_stats = (t._version, t.data_ptr(), stats)


def held_stats(t):
    """The column statistics the producer of `t` attached, if `t` has not changed since."""
    held = getattr(t, "_scn_stats", None)
    if held is not None and held[0] == t._version and held[1] == t.data_ptr() and held[2].size(1) == t.size(1):
        return held[2]
    return None


def bf16_operand(m, x, c_in, c_out):
    """Register the bf16 copy of `x` for the next convolution entry on handle m.  Returns the copy (kept by the
    caller for the backward pass) or None when this layer does not run on bf16 tiles."""
    if not _lib.lib().scn_bf16_plan(int(c_in), int(c_out), _precision) or x.numel() == 0:
        return None
    held = getattr(x, "_scn_bf16", None)
    if held is not None and held[0] == x._version and held[1] == x.data_ptr() and held[2].shape == x.shape:
        copy, ready = held[2], 1
    else:
        copy, ready = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device), 0
    _lib.check(_lib.lib().scn_bf16_operand(m._handle(), _ptr(x), _ptr(copy), ready))
    return copy


def bf16_operand_again(m, x, copy):
    """Backward pass: `copy` was filled during the forward call."""
    if copy is not None:
        _lib.check(_lib.lib().scn_bf16_operand(m._handle(), _ptr(x), _ptr(copy), 1))


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None and t.numel() else C.c_void_p(0)


def _cuda_f32(t, name):
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{name}: expected a CUDA float32 tensor (there is no CPU path), got {t.device} {t.dtype}")
    return t.contiguous()


def _rows_f32(t, name):
    """A CUDA float32 matrix whose rows may be a column slice of a wider one (stride(1) == 1): returned as it is with its row
    stride, anything else is made contiguous.  -> (tensor, ld)"""
    if not (t.is_cuda and t.dtype == torch.float32):
        raise TypeError(f"{name}: expected a CUDA float32 tensor (there is no CPU path), got {t.device} {t.dtype}")
    if t.dim() == 2 and t.stride(1) == 1 and t.stride(0) > t.size(1) and t.stride(0) % 4 == 0 and t.data_ptr() % 16 == 0 \
            and t.size(1) % 8 == 0:
        return t, t.stride(0)
    t = t.contiguous()
    return t, 0


def held_bf16(t):
    """The bf16 copy the producer of `t` attached (attach_bf16), if `t` has not changed since."""
    held = getattr(t, "_scn_bf16", None)
    if held is not None and held[0] == t._version and held[1] == t.data_ptr() and held[2].shape == t.shape:
        return held[2]
    return None


def _grad_operand(m, d_output_features, strided_ok):
    """d_out of a backward entry; a row-strided slice is registered with scn_grad_stride when the entry can read it in place,
    and a bf16 copy left by the kernel that produced a dense d_out (the BatchNorm backward behind this layer) is registered
    with scn_grad_bf16 so that the entry skips its own cast pass"""
    if not strided_ok:
        return _cuda_f32(d_output_features, "grad")
    g, ld = _rows_f32(d_output_features, "grad")
    if ld:
        _lib.check(_lib.lib().scn_grad_stride(m._handle(), int(ld)))
    elif _precision == _lib.BF16:
        copy = held_bf16(d_output_features)
        if copy is not None and g.data_ptr() == d_output_features.data_ptr():
            _lib.check(_lib.lib().scn_grad_bf16(m._handle(), _ptr(g), _ptr(copy)))
    return g


def _opt(t):
    """optionalTensor convention of the reference (utils.py:23-24): an empty tensor means 'absent'."""
    return t if (t is not None and t.numel()) else None


def _conv_input(m, input_features, input_bf16):
    """The `in` operand of a convolution entry.  input_bf16 (extension): the bf16 operand produced by a fused
    BatchNorm+ReLU (BatchNormalization_updateOutput(output_features=None, output_bf16=...)); there is no fp32 activation,
    so the bf16 buffer itself is handed in as `in` with a ready copy registered for it (include/scn_b200.h)."""
    if input_bf16 is None:
        return _cuda_f32(input_features, "input")
    if not (input_bf16.is_cuda and input_bf16.dtype == torch.bfloat16 and input_bf16.is_contiguous()):
        raise TypeError("input_bf16: expected a contiguous CUDA bfloat16 tensor")
    _lib.check(_lib.lib().scn_bf16_operand(m._handle(), _ptr(input_bf16), _ptr(input_bf16), 1))
    return input_bf16


def fuses_bn_conv(c_in, c_out):
    """True when a training-mode BatchNorm(+ReLU) -> convolution [c_in -> c_out] pair can run fused: the BatchNorm writes
    only the bf16 operand (forward product and weight gradient both read bf16), and the convolution's dgrad epilogue does
    the BatchNorm's backward reduction."""
    return (_precision == _lib.BF16 and _lib.lib().scn_bf16_plan(int(c_in), int(c_out), _precision) == 3
            and bool(_lib.lib().scn_bn_bwd_fusable(int(c_in), int(c_out), _precision)))


class Metadata_3:
    """Replaces Metadata<3> (Metadata/Metadata.h:218-364): one handle per batch, owns every scale's voxel
    keys, hash, neighbour tables and stride-2 links on the device; freed with the last Python reference."""

    def __init__(self):
        self._h = None
        self._device = None
        self.normal_guide_scale = None

    # the handle is created lazily so that it lands on the device of the first tensor it sees
    def _handle(self, device_index=None):
        if self._h is None:
            if device_index is None:
                device_index = torch.cuda.current_device()
            h = _lib.lib().scn_meta_create(int(device_index))
            if not h:
                raise _lib.ScnError(_lib.lib().scn_last_error().decode())
            self._h, self._device = C.c_void_p(h), device_index
        return self._h

    def __del__(self):
        try:
            if self._h is not None:
                _lib.lib().scn_meta_destroy(self._h)
                self._h = None
        except Exception:
            pass

    def clear(self):
        self.__del__()

    def setNormalGuideScale(self, v):          # Metadata::setNormalGuideScale (Metadata.cpp:87, ConvolutionRules.h:774)
        self.normal_guide_scale = v

    def guided(self, spatial_size):
        """True when the scale carries per-voxel normals, i.e. its rules are tap-permuted by orientation class."""
        return self._h is not None and bool(_lib.lib().scn_guided(self._h, _lib.size3(spatial_size)))

    def normalsOf(self, spatial_size):
        """(parity) float32 [N,3] CPU tensor: the per-voxel normals of a guided scale (Metadata::normals)."""
        n = self.getNActive(spatial_size)
        out = torch.empty((n, 3), dtype=torch.float32)
        _lib.check(_lib.lib().scn_normals(self._handle(), _lib.size3(spatial_size), _stream(), _ptr(out)))
        return out

    def submanifoldGuidedTable(self, spatial_size):
        """(parity) int32 [27,N] CPU tensor = the forward table with every output row's taps permuted by its orientation class,
        and the classes uint8 [N]."""
        n = self.getNActive(spatial_size)
        out = torch.empty((27, n), dtype=torch.int32)
        ori = torch.empty(n, dtype=torch.uint8)
        _lib.check(_lib.lib().scn_subm_guided_table(self._handle(), _lib.size3(spatial_size), _stream(), _ptr(out), _ptr(ori)))
        return out, ori

    def getNActive(self, spatial_size):
        return int(_lib.lib().scn_nactive(self._handle(), _lib.size3(spatial_size)))

    def getSpatialLocations(self, spatial_size):
        """CPU LongTensor [N,4] = (x,y,z,batch) in row order (Metadata.cpp:724-748)."""
        n = self.getNActive(spatial_size)
        if n < 0:
            raise _lib.ScnError("getSpatialLocations: no such scale")
        out = torch.empty((n, 4), dtype=torch.int64)
        _lib.check(_lib.lib().scn_spatial_locations(self._handle(), _lib.size3(spatial_size), _ptr(out)))
        return out

    # --- rulebook access for parity tests ------------------------------------------------------------
    def submanifoldNeighbourTable(self, spatial_size, dilated_rate=1):
        """int32 [27,N] CPU tensor: input row feeding output row o at offset k (times dilated_rate), or -1."""
        sz = _lib.size3(spatial_size)
        nr = C.c_int64(0)
        _dilation(self, dilated_rate)
        _lib.check(_lib.lib().scn_subm_rulebook(self._handle(), sz, _stream(), C.byref(nr)))
        n = self.getNActive(spatial_size)
        out = torch.empty((27, n), dtype=torch.int32)
        _dilation(self, dilated_rate)
        _lib.check(_lib.lib().scn_subm_neighbour_table(self._handle(), sz, _ptr(out)))
        return out, int(nr.value)

    def stridedTable(self, fine_size, coarse_size):
        """(parent int32 [Nfine], offset uint8 [Nfine], nCoarse)."""
        nc = C.c_int64(0)
        _lib.check(_lib.lib().scn_strided_rulebook(self._handle(), _lib.size3(fine_size), _lib.size3(coarse_size),
                                                   _stream(), C.byref(nc)))
        n = self.getNActive(fine_size)
        parent = torch.empty(n, dtype=torch.int32)
        off = torch.empty(n, dtype=torch.uint8)
        _lib.check(_lib.lib().scn_strided_table(self._handle(), _lib.size3(fine_size), _ptr(parent), _ptr(off)))
        return parent, off, int(nc.value)


def n_rulebook_bits():
    return 32


def ResolutionBasedScattering(m, points_lr, points_hr, stride):
    """Same entry as the reference's (pybind.cpp:33-35, sparseconvnet_cuda.cpp:203-209): points_lr [Nl,3], points_hr [Nh,3]
    CUDA int tensors; returns int32 [Nh] = row of the low-resolution voxel points_hr // stride falls into (its rank among the
    sorted unique lr voxels = the index into points_lr for a sample's spatial locations), -1 where there is none."""
    lr, hr = points_lr, points_hr
    if not (lr.is_cuda and hr.is_cuda):
        raise TypeError("ResolutionBasedScattering: expected CUDA tensors (there is no CPU path)")
    lr = lr[:, :3].to(torch.int32).contiguous()
    hr = hr[:, :3].to(torch.int32).contiguous()
    out = torch.empty(hr.size(0), dtype=torch.int32, device=hr.device)
    with _on(hr.device):
        _lib.check(_lib.lib().scn_resolution_scatter(_ptr(lr), lr.size(0), _ptr(hr), hr.size(0), int(stride), _ptr(out),
                                                     _stream()))
    return out


# ---- IO layers (sparseconvnet.h:151-179) -------------------------------------------------------------
def InputLayer_updateOutput(m, spatial_size, input_coords, input_features, output_features, batch_size, mode,
                            input_normal=None):
    feats = _cuda_f32(input_features, "InputLayer features")
    coords = input_coords
    if coords.dtype != torch.int64:
        coords = coords.long()
    coords = coords.contiguous()
    if coords.dim() != 2 or coords.size(1) != 4:
        raise ValueError("InputLayer: coords must be [P,4] = (x,y,z,batch) on this (3-D, batched) path")
    if coords.size(0) != feats.size(0):
        raise ValueError("InputLayer: coords and features disagree on the number of points")
    on_dev = coords.is_cuda
    n = C.c_int64(0)
    h = m._handle(feats.device.index)
    # use_normal of CUDA/IOLayers.cpp:41-42: a 2-D tensor with one row per point switches the normal-guided rules on
    nrm = None
    if torch.is_tensor(input_normal) and input_normal.dim() == 2 and input_normal.size(0) == feats.size(0):
        if input_normal.size(1) != 3:
            raise ValueError("InputLayer: normals must be [P,3]")
        nrm = input_normal.to(device=feats.device, dtype=torch.float32).contiguous()
    with _on(feats.device):
        if nrm is not None:
            scale = m.normal_guide_scale if m.normal_guide_scale is not None else 0
            _lib.check(_lib.lib().scn_input_normals(h, _ptr(nrm), int(scale)))
        _lib.check(_lib.lib().scn_input_layer_build(h, _lib.size3(spatial_size), _ptr(coords), int(on_dev),
                                                    coords.size(0), int(batch_size), int(mode), _stream(), C.byref(n)))
        m._input_size = _lib.size3(spatial_size)[:]
        output_features.resize_(n.value, feats.size(1))
        _lib.check(_lib.lib().scn_input_layer_fwd(h, _ptr(feats), feats.size(1), _ptr(output_features), _stream()))


def InputLayer_updateGradInput(m, d_input_features, d_output_features):
    g = _cuda_f32(d_output_features, "InputLayer grad")
    with _on(g.device):
        d_input_features.resize_(int(_lib.lib().scn_n_points(m._handle())), g.size(1))
        _lib.check(_lib.lib().scn_input_layer_bwd(m._handle(), _ptr(g), g.size(1), _ptr(d_input_features), _stream()))


def OutputLayer_updateOutput(m, input_features, output_features):
    x = _cuda_f32(input_features, "OutputLayer input")
    with _on(x.device):
        output_features.resize_(int(_lib.lib().scn_n_points(m._handle())), x.size(1))
        _lib.check(_lib.lib().scn_output_layer_fwd(m._handle(), _ptr(x), x.size(1), _ptr(output_features), _stream()))


def OutputLayer_updateGradInput(m, d_input_features, d_output_features):
    g = _cuda_f32(d_output_features, "OutputLayer grad")
    with _on(g.device):
        n = int(_lib.lib().scn_nactive(m._handle(), _lib.size3(m._input_size)))
        d_input_features.resize_(n, g.size(1))
        _lib.check(_lib.lib().scn_output_layer_bwd(m._handle(), _ptr(g), g.size(1), _ptr(d_input_features), _stream()))


# ---- convolutions (sparseconvnet.h:50-61, 89-116) ------------------------------------------------------
def _dilation(m, rate):
    """dilated_rate of the reference entries: taps at offsets rate*(dx,dy,dz) for the next submanifold entry on this handle"""
    if int(rate) != 1:
        _lib.check(_lib.lib().scn_subm_dilation(m._handle(), int(rate)))


def _check_weight(weight, v):
    w = _cuda_f32(weight, "weight")
    if w.dim() != 3 or w.size(0) != v:
        raise ValueError(f"weight must be [{v}, nIn, nOut]")
    return w


def SubmanifoldConvolution_updateOutput(spatial_size, filter_size, m, input_features, output_features, weight, bias,
                                        dilated_rate=1, residual=None, stats=None, input_bf16=None):
    """residual (extension): [N, nOut] tensor added to the result inside the kernel (see fuses_residual).
    stats (extension): float64 [2, nOut] tensor that receives the column sums / sums of squares of the result.
    input_bf16 (extension): see _conv_input."""
    if any(int(f) != 3 for f in filter_size.tolist()):
        raise NotImplementedError("SubmanifoldConvolution: only 3x3x3 filters are on this path")
    x, w, b = _conv_input(m, input_features, input_bf16), _check_weight(weight, 27), _opt(bias)
    if residual is not None:
        residual = _cuda_f32(residual, "residual")
        if residual.shape != (x.size(0), w.size(2)):
            raise ValueError(f"SubmanifoldConvolution: residual is {tuple(residual.shape)}")
    macs = C.c_double(0.0)
    with _on(x.device):
        n = m.getNActive(spatial_size)
        if x.size(0) != n or x.size(1) != w.size(1):
            raise ValueError(f"SubmanifoldConvolution: input is {tuple(x.shape)}, scale has {n} rows, nIn={w.size(1)}")
        output_features.resize_(n, w.size(2))
        _dilation(m, dilated_rate)
        _lib.check(_lib.lib().scn_subm_fwd(m._handle(), _lib.size3(spatial_size), _ptr(x), _ptr(w), _ptr(b),
                                           _ptr(residual), _ptr(stats), _ptr(output_features), w.size(1), w.size(2),
                                           _precision, _stream(), C.byref(macs)))
    return macs.value


def BatchNormalization_evalCoefficients(runningMean, runningVar, weight, bias, eps):
    """(extension) scale, shift [C] of the inference BatchNorm y = scale*x + shift, evaluated exactly as
    BatchNormalization_updateOutput(train=False) evaluates them."""
    c = runningMean.numel()
    scale, shift = torch.empty_like(runningMean), torch.empty_like(runningMean)
    with _on(runningMean.device):
        _lib.check(_lib.lib().scn_bn_eval_coeffs(_ptr(runningMean), _ptr(runningVar), _ptr(_opt(weight)), _ptr(_opt(bias)), c,
                                                 float(eps), _ptr(scale), _ptr(shift), _stream()))
    return scale, shift


def SubmanifoldConvolutionBN_updateOutput(spatial_size, filter_size, m, input_features, output_features, weight, bias,
                                          bn_scale, bn_shift, leakiness, residual=None, output_bf16=None):
    """(extension, inference) SubmanifoldConvolution + BatchNorm(running statistics) + (leaky) ReLU in one kernel: the
    BatchNorm+ReLU runs in the convolution epilogue.  Bit-identical to the two separate entries."""
    if any(int(f) != 3 for f in filter_size.tolist()):
        raise NotImplementedError("SubmanifoldConvolution: only 3x3x3 is on this path")
    x, w, b = _cuda_f32(input_features, "input"), _check_weight(weight, 27), _opt(bias)
    macs = C.c_double(0.0)
    with _on(x.device):
        n = m.getNActive(spatial_size)
        if x.size(0) != n or x.size(1) != w.size(1):
            raise ValueError(f"SubmanifoldConvolution: input is {tuple(x.shape)}, scale has {n} rows, nIn={w.size(1)}")
        output_features.resize_(n, w.size(2))
        if output_bf16 is not None:
            output_bf16.resize_(n, w.size(2))
        _lib.check(_lib.lib().scn_subm_fwd_bn(m._handle(), _lib.size3(spatial_size), _ptr(x), _ptr(w), _ptr(b), _ptr(residual),
                                              _ptr(bn_scale), _ptr(bn_shift), float(leakiness), _ptr(output_features),
                                              _ptr(output_bf16), w.size(1), w.size(2), _precision, _stream(),
                                              C.byref(macs)))
    return macs.value


def SubmanifoldConvolution_backward(spatial_size, filter_size, m, input_features, d_input_features, d_output_features,
                                    weight, d_weight, d_bias, dilated_rate=1, input_bf16=None):
    x, w = _conv_input(m, input_features, input_bf16), _check_weight(weight, 27)
    g = _grad_operand(m, d_output_features, input_bf16 is not None and _opt(d_bias) is None)
    with _on(x.device):
        if d_input_features is not None:      # None: the caller does not need the input gradient (first layer)
            d_input_features.resize_(x.size(0), x.size(1))
        _dilation(m, dilated_rate)
        _lib.check(_lib.lib().scn_subm_bwd(m._handle(), _lib.size3(spatial_size), _ptr(x), _ptr(g), _ptr(w),
                                           _ptr(d_input_features), _ptr(d_weight), _ptr(_opt(d_bias)), w.size(1),
                                           w.size(2), _precision, _stream()))


def _check_2s2(filter_size, filter_stride):
    if any(int(f) != 2 for f in filter_size.tolist()) or any(int(f) != 2 for f in filter_stride.tolist()):
        raise NotImplementedError("Convolution/Deconvolution: only size 2 / stride 2 is on this path "
                                  "(the reference GPU builder asserts the same, ConvolutionRules.h:354-358)")


def Convolution_updateOutput(in_size, out_size, filter_size, filter_stride, m, input_features, output_features, weight,
                             bias, input_bf16=None, stats=None):
    _check_2s2(filter_size, filter_stride)
    w, b = _check_weight(weight, 8), _opt(bias)
    macs, nc = C.c_double(0.0), C.c_int64(0)
    with _on(weight.device):
        _lib.check(_lib.lib().scn_strided_rulebook(m._handle(), _lib.size3(in_size), _lib.size3(out_size), _stream(),
                                                   C.byref(nc)))
        x = _conv_input(m, input_features, input_bf16)
        output_features.resize_(nc.value, w.size(2))
        if stats is not None:
            _lib.check(_lib.lib().scn_out_stats(m._handle(), _ptr(stats)))
        _lib.check(_lib.lib().scn_conv_fwd(m._handle(), _lib.size3(in_size), _lib.size3(out_size), _ptr(x), _ptr(w),
                                           _ptr(b), _ptr(output_features), w.size(1), w.size(2), _precision, _stream(),
                                           C.byref(macs)))
    return macs.value


def Convolution_backward(in_size, out_size, filter_size, filter_stride, m, input_features, d_input_features,
                         d_output_features, weight, d_weight, d_bias, input_bf16=None):
    x, w = _conv_input(m, input_features, input_bf16), _check_weight(weight, 8)
    g = _grad_operand(m, d_output_features, input_bf16 is not None and _opt(d_bias) is None)
    with _on(x.device):
        d_input_features.resize_(x.size(0), x.size(1))
        _lib.check(_lib.lib().scn_conv_bwd(m._handle(), _lib.size3(in_size), _lib.size3(out_size), _ptr(x), _ptr(g),
                                           _ptr(w), _ptr(d_input_features), _ptr(d_weight), _ptr(_opt(d_bias)),
                                           w.size(1), w.size(2), _precision, _stream()))


def Deconvolution_updateOutput(in_size, out_size, filter_size, filter_stride, m, input_features, output_features,
                               weight, bias, input_bf16=None, stats=None):
    _check_2s2(filter_size, filter_stride)
    x, w, b = _conv_input(m, input_features, input_bf16), _check_weight(weight, 8), _opt(bias)
    macs = C.c_double(0.0)
    with _on(x.device):
        n = m.getNActive(out_size)
        if n < 0:
            raise _lib.ScnError("Deconvolution: output scale does not exist (no matching Convolution ran on this batch)")
        output_features.resize_(n, w.size(2))
        if stats is not None:
            _lib.check(_lib.lib().scn_out_stats(m._handle(), _ptr(stats)))
        _lib.check(_lib.lib().scn_deconv_fwd(m._handle(), _lib.size3(in_size), _lib.size3(out_size), _ptr(x), _ptr(w),
                                             _ptr(b), _ptr(output_features), w.size(1), w.size(2), _precision,
                                             _stream(), C.byref(macs)))
    return macs.value


def Deconvolution_backward(in_size, out_size, filter_size, filter_stride, m, input_features, d_input_features,
                           d_output_features, weight, d_weight, d_bias, input_bf16=None):
    x, w = _conv_input(m, input_features, input_bf16), _check_weight(weight, 8)
    g = _grad_operand(m, d_output_features, input_bf16 is not None and _opt(d_bias) is None)
    with _on(x.device):
        d_input_features.resize_(x.size(0), x.size(1))
        _lib.check(_lib.lib().scn_deconv_bwd(m._handle(), _lib.size3(in_size), _lib.size3(out_size), _ptr(x), _ptr(g),
                                             _ptr(w), _ptr(d_input_features), _ptr(d_weight), _ptr(_opt(d_bias)),
                                             w.size(1), w.size(2), _precision, _stream()))


# ---- batch norm (sparseconvnet.h:21-33) ------------------------------------------------------------------
def BatchNormalization_updateOutput(input_features, output_features, saveMean, saveInvStd, runningMean, runningVar,
                                    weight, bias, eps, momentum, train, leakiness, output_bf16=None, stats=None):
    """output_bf16 (extension): an EMPTY bfloat16 tensor that receives the bf16 copy of the output for the
    tensor-core convolution that follows (see attach_bf16 / bf16_operand).
    stats (extension): float64 [2, C] column sums / sums of squares of the input, made by the convolution that
    produced it (attach_stats); the training-mode reduction pass is skipped."""
    x = _cuda_f32(input_features, "input")
    with _on(x.device):
        if output_features is not None:        # None (extension): only the bf16 operand is written
            output_features.resize_(x.size(0), x.size(1))
        if output_bf16 is not None:
            output_bf16.resize_(x.size(0), x.size(1))
        _lib.check(_lib.lib().scn_bn_fwd(_ptr(x), _ptr(output_features), _ptr(output_bf16), _ptr(stats), _ptr(saveMean),
                                         _ptr(saveInvStd),
                                         _ptr(runningMean), _ptr(runningVar), _ptr(_opt(weight)), _ptr(_opt(bias)),
                                         x.size(0), x.size(1), float(eps), float(momentum), int(bool(train)),
                                         float(leakiness), _stream()))


def BatchNormalization_backward(input_features, d_input_features, output_features, d_output_features, saveMean,
                                saveInvStd, runningMean, runningVar, weight, bias, d_weight, d_bias, leakiness,
                                d_input_add=None):
    """d_input_add (extension): a gradient of the same input that arrived through a residual shortcut; it is added to
    d_input in the same pass."""
    x, g = _cuda_f32(input_features, "input"), _cuda_f32(d_output_features, "grad")
    if d_input_add is not None:
        d_input_add = _cuda_f32(d_input_add, "d_input_add")
    with _on(x.device):
        d_input_features.resize_(x.size(0), x.size(1))
        _lib.check(_lib.lib().scn_bn_bwd(_ptr(x), _ptr(output_features), _ptr(g), _ptr(saveMean), _ptr(saveInvStd),
                                         _ptr(_opt(weight)), _ptr(_opt(bias)), _ptr(d_input_add), _ptr(d_input_features),
                                         _ptr(_opt(d_weight)),
                                         _ptr(_opt(d_bias)), x.size(0), x.size(1), float(leakiness), _stream()))


def BatchNormalization_backwardFusion(m, input_features, saveMean, saveInvStd, weight, bias, leakiness, acc):
    """(extension) register this BatchNorm for the NEXT *_backward entry on handle m: that entry's d_input comes back
    already multiplied by the activation mask, and acc (float64 [2, C]) holds sum d' and sum d'*x (scn_bn_bwd_fusion)."""
    x = _cuda_f32(input_features, "input")
    _lib.check(_lib.lib().scn_bn_bwd_fusion(m._handle(), _ptr(x), _ptr(saveMean), _ptr(saveInvStd), _ptr(_opt(weight)),
                                            _ptr(_opt(bias)), float(leakiness), _ptr(acc)))


def BatchNormalization_backwardApply(input_features, d_masked, acc, saveMean, saveInvStd, weight, d_input_features, d_weight,
                                     d_bias, d_input_add=None, d_input_bf16=None):
    """(extension) second half of BatchNormalization_backward after a fused dgrad epilogue (scn_bn_bwd_apply).
    d_input_bf16: optional bfloat16 tensor that receives a copy of d_input_features (see _grad_operand)."""
    x, g = _cuda_f32(input_features, "input"), _cuda_f32(d_masked, "grad")
    ld_add = 0
    if d_input_add is not None:
        d_input_add, ld_add = _rows_f32(d_input_add, "d_input_add")      # may be a column slice of a joined gradient
    with _on(x.device):
        d_input_features.resize_(x.size(0), x.size(1))
        if d_input_bf16 is not None:
            d_input_bf16.resize_(x.size(0), x.size(1))
        _lib.check(_lib.lib().scn_bn_bwd_apply(_ptr(x), _ptr(g), _ptr(acc), _ptr(saveMean), _ptr(saveInvStd), _ptr(_opt(weight)),
                                               _ptr(d_input_add), int(ld_add), _ptr(d_input_features), _ptr(d_input_bf16),
                                               _ptr(_opt(d_weight)), _ptr(_opt(d_bias)), x.size(0), x.size(1), _stream()))


# ---- 1x1 "NetworkInNetwork": the reference itself calls ATen's GEMM here (CUDA/NetworkInNetwork.cpp:9-50).
# Plain library GEMMs; in 'tf32' precision they run on the tensor cores like the convolutions around them.
class _gemm_precision:
    def __enter__(self):
        self.prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = _precision != _lib.FP32

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32 = self.prev


def NetworkInNetwork_updateOutput(input_features, output_features, weight, bias):
    n = input_features.size(0)
    output_features.resize_(n, weight.size(1))
    with _gemm_precision():
        if _opt(bias) is not None:
            torch.addmm(bias, input_features, weight, out=output_features)
        else:
            torch.mm(input_features, weight, out=output_features)
    return float(n * weight.size(0) * weight.size(1))


def NetworkInNetwork_updateGradInput(d_input_features, d_output_features, weight):
    d_input_features.resize_(d_output_features.size(0), weight.size(0))
    with _gemm_precision():
        torch.mm(d_output_features, weight.t(), out=d_input_features)


def NetworkInNetwork_accGradParameters(input_features, d_output_features, d_weight, d_bias):
    if input_features.size(0):
        if d_bias is not None and d_bias.numel():
            torch.sum(d_output_features, 0, out=d_bias)
        with _gemm_precision():
            torch.mm(input_features.t(), d_output_features, out=d_weight)
