"""nn.Module surface of the hot path, constructor-compatible with the reference classes so that
examples/ScanNet/model.py (InstanceDenseUNet) and reference state_dicts work unchanged:

  InputLayer / OutputLayer        sparseconvnet/ioLayers.py:15-87
  SubmanifoldConvolution          sparseconvnet/submanifoldConvolution.py:18-69   weight [27, nIn, nOut]
  Convolution / Deconvolution     sparseconvnet/convolution.py:14-70, deconvolution.py:13-85   weight [8, nIn, nOut]
  BatchNormalization (+ReLU...)   sparseconvnet/batchNormalization.py:13-88       weight, bias, running_mean, running_var
  NetworkInNetwork                sparseconvnet/networkInNetwork.py:62-88         weight [nIn, nOut]
  Sequential / ConcatTable / AddTable / JoinTable / Identity   sequential.py, tables.py, identity.py

Parameter names, shapes and initialisation (normal(0, sqrt(2/(nIn*volume)))) follow the reference.
"""
import torch
from torch.nn import Module, Parameter

from . import SCN
from . import functions as F
from .SCN import Metadata_3
from .tensor import SparseConvNetTensor
from .utils import optionalTensor, toLongTensor


def Metadata(dim):
    if dim != 3:
        raise NotImplementedError("only 3-D grids are on the B200 hot path")
    return Metadata_3()


def _same(ref, features, spatial_size=None):
    t = SparseConvNetTensor(features, ref.metadata, ref.spatial_size if spatial_size is None else spatial_size)
    return t


def _size_str(v):
    v = v.tolist()
    return str(v[0]) if min(v) == max(v) else "(" + ",".join(map(str, v)) + ")"


# ---- containers -------------------------------------------------------------------------------------------
def _conv_bn_inference(conv, bn, input, residual=None):
    """Inference only: SubmanifoldConvolution `conv` followed by BatchNormalization `bn` as ONE kernel (the BatchNorm +
    ReLU runs in the convolution epilogue and, in bf16 mode, also leaves the bf16 copy for the next convolution)."""
    conv._check(input)
    x = input.features
    # scale / shift depend only on the layer's parameters and running statistics: cached until one of them changes
    tensors = [bn.running_mean, bn.running_var, optionalTensor(bn, "weight"), optionalTensor(bn, "bias")]
    key = tuple((t.data_ptr(), t._version) for t in tensors)
    held = getattr(bn, "_scn_eval_coeffs", None)
    if held is None or held[0] != key:
        held = (key,) + SCN.BatchNormalization_evalCoefficients(*tensors, bn.eps)
        bn._scn_eval_coeffs = held
    scale, shift = held[1], held[2]
    out = x.new_empty(0)
    out16 = torch.empty(0, dtype=torch.bfloat16, device=x.device) if SCN.wants_bf16(conv.nOut) else None
    x16 = SCN.bf16_operand(input.metadata, x, conv.nIn, conv.nOut)
    macs = SCN.SubmanifoldConvolutionBN_updateOutput(input.spatial_size, conv.filter_size, input.metadata, x, out, conv.weight,
                                                     optionalTensor(conv, "bias"), scale, shift, bn.leakiness, residual, out16)
    del x16
    F._count(macs, out)
    if out16 is not None:
        SCN.attach_bf16(out, out16)
    return _same(input, out)


def _has_hooks(m):
    return bool(m._forward_hooks or m._forward_pre_hooks or m._backward_hooks or m._backward_pre_hooks)


def _fusable_inference_pair(a, b, input):
    # the fused entry is the 3x3x3, dilation-1 kernel and bypasses the modules' __call__: anything else (dilated or
    # other filter sizes, hooks registered on either module) takes the two separate layers
    return (isinstance(a, SubmanifoldConvolution) and isinstance(b, BatchNormalization) and not b.training
            and not torch.is_grad_enabled() and input.features.is_cuda and a.nOut == b.nPlanes
            and getattr(a, "dilated_rate", 1) == 1 and a.filter_volume == 27
            and not _has_hooks(a) and not _has_hooks(b)
            and SCN.fuses_residual(a.nIn, a.nOut))


def _fusable_training_pair(bn, conv, input):
    """Training: BatchNorm(+ReLU) `bn` directly followed by a convolution `conv` runs as one autograd node
    (functions.BatchNormConvFunction) when both of the convolution's bf16 products and its dgrad epilogue allow it."""
    return (isinstance(bn, BatchNormalization) and isinstance(conv, (SubmanifoldConvolution, Convolution, Deconvolution))
            and bn.training and torch.is_grad_enabled() and input.features.is_cuda and input.features.requires_grad
            and bn.nPlanes == conv.nIn and getattr(conv, "dilated_rate", 1) == 1
            and conv.filter_volume == (27 if isinstance(conv, SubmanifoldConvolution) else 8)
            and not _has_hooks(bn) and not _has_hooks(conv) and input.features.size(0) > 1
            and SCN.fuses_bn_conv(conv.nIn, conv.nOut)
            # a normal-guided submanifold dgrad is three passes (one per orientation class): no single epilogue to fuse into
            and not (isinstance(conv, SubmanifoldConvolution) and input.metadata.guided(input.spatial_size)))


def _bn_conv_training(bn, conv, input, residual=None, with_alias=False):
    conv._check(input)
    if isinstance(conv, SubmanifoldConvolution):
        kind, in_size, out_size, stride = "subm", input.spatial_size, input.spatial_size, conv.filter_size
    elif isinstance(conv, Convolution):
        kind, in_size, stride = "conv", input.spatial_size, conv.filter_stride
        out_size = (in_size - conv.filter_size) // stride + 1
        assert ((out_size - 1) * stride + conv.filter_size == in_size).all(), (in_size, out_size)
    else:
        kind, in_size, stride = "deconv", input.spatial_size, conv.filter_stride
        out_size = (in_size - 1) * stride + conv.filter_size
    want_stats = SCN.fuses_residual(conv.nIn, conv.nOut)      # the epilogue also leaves the statistics for the next BatchNorm
    out, stats, alias = F.BatchNormConvFunction.apply(
        input.features, optionalTensor(bn, "weight"), optionalTensor(bn, "bias"), bn.running_mean, bn.running_var, bn.eps,
        bn.momentum, bn.leakiness, conv.weight, optionalTensor(conv, "bias"), input.metadata, kind, in_size, out_size,
        conv.filter_size, stride, residual, want_stats, with_alias)
    if stats.numel():
        SCN.attach_stats(out, stats)
    t = _same(input, out, out_size)
    return (t, _same(input, alias)) if with_alias else t


class Sequential(torch.nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, input, start=0):
        # same as torch.nn.Sequential, except that at inference a SubmanifoldConvolution directly followed by a
        # BatchNorm(+ReLU) runs as one kernel (fused BatchNorm+ReLU epilogue), and in training a BatchNorm(+ReLU) directly
        # followed by a convolution runs as one autograd node; `start` skips modules a caller has already applied
        mods = list(self._modules.values())
        i = start
        while i < len(mods):
            if i + 1 < len(mods) and isinstance(input, SparseConvNetTensor) and _fusable_inference_pair(mods[i], mods[i + 1], input):
                input = _conv_bn_inference(mods[i], mods[i + 1], input)
                i += 2
            elif i + 1 < len(mods) and isinstance(input, SparseConvNetTensor) and _fusable_training_pair(mods[i], mods[i + 1], input):
                input = _bn_conv_training(mods[i], mods[i + 1], input)      # training: BN(+ReLU) -> conv as one node
                i += 2
            else:
                input = mods[i](input)
                i += 1
        return input

    def input_spatial_size(self, out_size):
        for m in reversed(list(self._modules.values())):
            out_size = m.input_spatial_size(out_size)
        return out_size


class ConcatTable(torch.nn.Sequential):
    def add(self, module):
        self._modules[str(len(self._modules))] = module
        return self

    def forward(self, input):
        mods = list(self._modules.values())
        # The UNet's skip connection, ConcatTable(Identity, Sequential(BN, Convolution, ...)) (networkArchitectures.py:246-260):
        # the input feeds both branches, so autograd would add the two gradients in a separate pass.  When the inner
        # branch starts with a fusable BatchNorm -> convolution pair, the Identity branch receives the input re-issued by that
        # node (a view), and whatever gradient returns through it is added inside the BatchNorm's backward kernel.
        if (len(mods) == 2 and isinstance(mods[0], Identity) and isinstance(mods[1], Sequential)
                and isinstance(input, SparseConvNetTensor) and len(mods[1]._modules) >= 2):
            inner = list(mods[1]._modules.values())
            if _fusable_training_pair(inner[0], inner[1], input):
                held = SCN.held_stats(input.features)
                t, alias = _bn_conv_training(inner[0], inner[1], input, with_alias=True)
                if held is not None:
                    SCN.attach_stats(alias.features, held)       # the re-issued view carries its producer's column statistics on
                return [alias, mods[1].forward(t, start=2)]
        return [m(input) for m in mods]

    def input_spatial_size(self, out_size):
        return self._modules["0"].input_spatial_size(out_size)


class ResidualConcatTable(ConcatTable):
    """ConcatTable([shortcut, Sequential(BN, SubmConv, BN, SubmConv)]) of a residual block -- same children, hence the
    same state_dict keys as the reference (networkArchitectures.py:225-240) -- whose forward fuses the two elementwise
    passes of the block into its kernels: the shortcut is added in the epilogue of the last convolution, and the
    gradient that returns through the shortcut is added inside the first BatchNorm's backward kernel.  Returns a
    one-element list, so the AddTable that follows passes it through."""

    def forward(self, input):
        shortcut, body = self._modules["0"], self._modules["1"]
        mods = list(body._modules.values())
        last = mods[-1]
        fusable = (torch.is_grad_enabled() and len(mods) == 4 and isinstance(mods[0], BatchNormalization)
                   and isinstance(last, SubmanifoldConvolution) and input.features.is_cuda
                   and SCN.fuses_residual(last.nIn, last.nOut) and mods[0].training)
        if not fusable:
            return [shortcut(input), body(input)]
        if _fusable_training_pair(mods[0], mods[1], input) and isinstance(mods[1], SubmanifoldConvolution) \
                and _fusable_training_pair(mods[2], last, input):
            # both BN -> conv pairs as single nodes: bf16-only activations, BatchNorm backward reductions in the dgrad epilogues
            t, alias = _bn_conv_training(mods[0], mods[1], input, with_alias=True)
            s = shortcut(alias)
            return [_bn_conv_training(mods[2], last, t, residual=s.features)]
        t, alias = mods[0](input, with_alias=True)
        s = shortcut(alias)
        for m in mods[1:-1]:
            t = m(t)
        return [last(t, residual=s.features)]


class AddTable(torch.nn.Sequential):
    def forward(self, input):
        total = input[0].features
        for t in input[1:]:
            total = total + t.features
        return _same(input[0], total)

    def input_spatial_size(self, out_size):
        return out_size


class JoinTable(torch.nn.Sequential):
    def forward(self, input):
        feats = torch.cat([t.features for t in input], 1)
        # column statistics concatenate like the columns do: the BatchNorm behind the join skips its reduction pass
        st = [SCN.held_stats(t.features) if t.features.is_cuda else None for t in input]
        if all(s is not None for s in st):
            SCN.attach_stats(feats, torch.cat(st, 1))
        return _same(input[0], feats)

    def input_spatial_size(self, out_size):
        return out_size


class Identity(Module):
    def forward(self, input):
        return input

    def input_spatial_size(self, out_size):
        return out_size


# ---- IO -------------------------------------------------------------------------------------------------
class InputLayer(Module):
    """input = [coords [P,4] (x,y,z,batch; float or long), features [P,C] CUDA float, normals ([P,3] float: switches the
    normal-guided rules on, as in the reference; None or any other shape: plain rules), batch_size].  mode 3 sums, mode 4 averages the features of points sharing a voxel."""

    def __init__(self, dimension, spatial_size, mode=3, normal_guide_scale=10240):
        super().__init__()
        self.dimension = dimension
        self.spatial_size = toLongTensor(dimension, spatial_size)
        self.mode = mode
        self.normal_guide_scale = normal_guide_scale

    def forward(self, input):
        coords = input[0]
        if not coords.is_cuda:
            coords = coords.type(torch.LongTensor)     # truncation toward zero, as ioLayers.py:56
        out = SparseConvNetTensor(metadata=Metadata(self.dimension), spatial_size=self.spatial_size)
        batch_size = input[3] if len(input) > 3 else int(coords[:, -1].max().item()) + 1
        normals = input[2] if len(input) > 2 else None
        out.features = F.InputLayerFunction.apply(self.dimension, out.metadata, self.spatial_size, coords, input[1],
                                                  batch_size, self.mode, normals, self.normal_guide_scale)
        return out


class OutputLayer(Module):
    def __init__(self, dimension):
        super().__init__()
        self.dimension = dimension

    def forward(self, input):
        return F.OutputLayerFunction.apply(self.dimension, input.metadata, input.features)


# ---- convolutions -----------------------------------------------------------------------------------------
class _ConvBase(Module):
    def _init_weight(self, dimension, nIn, nOut, filter_size, bias):
        self.dimension, self.nIn, self.nOut = dimension, nIn, nOut
        self.filter_size = toLongTensor(dimension, filter_size)
        self.filter_volume = int(self.filter_size.prod().item())
        std = (2.0 / nIn / self.filter_volume) ** 0.5
        self.weight = Parameter(torch.empty(self.filter_volume, nIn, nOut).normal_(0, std))
        if bias:
            self.bias = Parameter(torch.zeros(nOut))

    def _check(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn, (self.nIn, self.nOut, input)


class SubmanifoldConvolution(_ConvBase):
    def __init__(self, dimension, nIn, nOut, filter_size, bias, dilated_rate=1):
        super().__init__()
        self._init_weight(dimension, nIn, nOut, filter_size, bias)
        self.dilated_rate = dilated_rate

    def forward(self, input, residual=None):
        """residual (extension): features of the shortcut branch, added inside the kernel (SCN.fuses_residual)."""
        self._check(input)
        # training on the tensor-core path: the kernel epilogue also accumulates the column statistics of the result,
        # which the BatchNorm that consumes it picks up instead of running its own reduction pass
        want_stats = self.training and torch.is_grad_enabled() and input.features.is_cuda and \
            SCN.fuses_residual(self.nIn, self.nOut)
        feats, stats = F.SubmanifoldConvolutionFunction.apply(input.features, self.weight, optionalTensor(self, "bias"),
                                                              input.metadata, input.spatial_size, self.dimension,
                                                              self.filter_size, self.dilated_rate, residual, want_stats)
        if stats.numel():
            SCN.attach_stats(feats, stats)
        return _same(input, feats)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return f"SubmanifoldConvolution {self.nIn}->{self.nOut} C{_size_str(self.filter_size)}"


ValidConvolution = SubmanifoldConvolution


class Convolution(_ConvBase):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias):
        super().__init__()
        self._init_weight(dimension, nIn, nOut, filter_size, bias)
        self.filter_stride = toLongTensor(dimension, filter_stride)

    def forward(self, input):
        self._check(input)
        out_size = (input.spatial_size - self.filter_size) // self.filter_stride + 1
        assert ((out_size - 1) * self.filter_stride + self.filter_size == input.spatial_size).all(), \
            (input.spatial_size, out_size, self.filter_size, self.filter_stride)
        feats = F.ConvolutionFunction.apply(input.features, self.weight, optionalTensor(self, "bias"), input.metadata,
                                            input.spatial_size, out_size, self.dimension, self.filter_size,
                                            self.filter_stride)
        return _same(input, feats, out_size)

    def input_spatial_size(self, out_size):
        return (out_size - 1) * self.filter_stride + self.filter_size

    def __repr__(self):
        return f"Convolution {self.nIn}->{self.nOut} C{_size_str(self.filter_size)}/{_size_str(self.filter_stride)}"


class Deconvolution(_ConvBase):
    def __init__(self, dimension, nIn, nOut, filter_size, filter_stride, bias):
        super().__init__()
        self._init_weight(dimension, nIn, nOut, filter_size, bias)
        self.filter_stride = toLongTensor(dimension, filter_stride)

    def forward(self, input):
        self._check(input)
        out_size = (input.spatial_size - 1) * self.filter_stride + self.filter_size
        feats = F.DeconvolutionFunction.apply(input.features, self.weight, optionalTensor(self, "bias"), input.metadata,
                                              input.spatial_size, out_size, self.dimension, self.filter_size,
                                              self.filter_stride)
        return _same(input, feats, out_size)

    def input_spatial_size(self, out_size):
        in_size = (out_size - self.filter_size) // self.filter_stride + 1
        assert ((in_size - 1) * self.filter_stride + self.filter_size == out_size).all()
        return in_size

    def __repr__(self):
        return f"Deconvolution {self.nIn}->{self.nOut} C{_size_str(self.filter_size)}/{_size_str(self.filter_stride)}"


class NetworkInNetwork(Module):
    def __init__(self, nIn, nOut, bias):
        super().__init__()
        self.nIn, self.nOut = nIn, nOut
        self.weight = Parameter(torch.empty(nIn, nOut).normal_(0, (2.0 / nIn) ** 0.5))
        if bias:
            self.bias = Parameter(torch.zeros(nOut))

    def forward(self, input):
        assert input.features.nelement() == 0 or input.features.size(1) == self.nIn, (self.nIn, input.features.shape)
        return _same(input, F.NetworkInNetworkFunction.apply(input.features, self.weight, optionalTensor(self, "bias")))

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        return f"NetworkInNetwork{self.nIn}->{self.nOut}"


# ---- batch norm ---------------------------------------------------------------------------------------------
class BatchNormalization(Module):
    """eps 1e-4, momentum 0.9 (running = 0.9*running + 0.1*batch), leakiness: 0 ReLU, (0,1) leaky, 1 none."""

    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, affine=True, leakiness=1):
        super().__init__()
        self.nPlanes, self.eps, self.momentum, self.affine, self.leakiness = nPlanes, eps, momentum, affine, leakiness
        self.register_buffer("running_mean", torch.zeros(nPlanes))
        self.register_buffer("running_var", torch.ones(nPlanes))
        if affine:
            self.weight = Parameter(torch.ones(nPlanes))
            self.bias = Parameter(torch.zeros(nPlanes))

    def forward(self, input, with_alias=False):
        """with_alias (extension): also return the input re-issued as a view whose gradient is folded into this
        layer's backward kernel (used for the shortcut of a residual block)."""
        assert input.features.nelement() == 0 or input.features.size(1) == self.nPlanes, \
            (self.nPlanes, input.features.shape)
        feats, feats16, alias = F.BatchNormalizationFunction.apply(
            input.features, optionalTensor(self, "weight"), optionalTensor(self, "bias"), self.running_mean,
            self.running_var, self.eps, self.momentum, self.training, self.leakiness, with_alias)
        if feats16.numel():
            SCN.attach_bf16(feats, feats16)
        if with_alias:
            return _same(input, feats), _same(input, alias)
        return _same(input, feats)

    def input_spatial_size(self, out_size):
        return out_size

    def __repr__(self):
        s = f"BatchNorm({self.nPlanes},eps={self.eps},momentum={self.momentum},affine={self.affine}"
        return s + (f",leakiness={self.leakiness})" if self.leakiness > 0 else ")")


class BatchNormReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9):
        super().__init__(nPlanes, eps, momentum, True, 0)


class BatchNormLeakyReLU(BatchNormalization):
    def __init__(self, nPlanes, eps=1e-4, momentum=0.9, leakiness=0.333):
        super().__init__(nPlanes, eps, momentum, True, leakiness)
