"""autograd glue between the nn.Modules in layers.py and the SCN entry points.

Same division of labour as the reference Function classes (submanifoldConvolution.py:76-128,
convolution.py:72-127, deconvolution.py:87-155, batchNormalization.py:90-161, ioLayers.py:157-223,
networkInNetwork.py:14-59): forward allocates an empty output, calls SCN.*_updateOutput, keeps the
Metadata alive on ctx; backward hands zero-initialised parameter gradients to SCN.*_backward.
"""
import torch
from torch.autograd import Function

from . import SCN
from .utils import optionalTensorReturn

# the reference keeps two global counters on the package (sparseconvnet/__init__.py:7-8)
counters = {"multiplyAdd": 0.0, "hidden_states": 0}


def _count(macs, out):
    counters["multiplyAdd"] += macs
    counters["hidden_states"] += out.nelement()


class SubmanifoldConvolutionFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, metadata, spatial_size, dimension, filter_size, dilated_rate=1, residual=None,
                want_stats=False):
        """Returns (out, stats).  residual (extension): added to the result in the kernel epilogue; its gradient is
        grad_out itself.  stats: float64 [2, nOut] column sums / sums of squares of out when want_stats, else empty."""
        ctx.scn_meta, ctx.dilated_rate = metadata, dilated_rate
        ctx.save_for_backward(x, spatial_size, weight, bias, filter_size)
        out = x.new_empty(0)
        stats = torch.empty((2, weight.size(2)) if want_stats else 0, dtype=torch.float64, device=x.device)
        ctx.x16 = SCN.bf16_operand(metadata, x, weight.size(1), weight.size(2))
        _count(SCN.SubmanifoldConvolution_updateOutput(spatial_size, filter_size, metadata, x, out, weight, bias,
                                                       dilated_rate, residual, stats if want_stats else None), out)
        ctx.mark_non_differentiable(stats)
        return out, stats

    @staticmethod
    def backward(ctx, grad_out, _grad_stats=None):
        x, spatial_size, weight, bias, filter_size = ctx.saved_tensors
        gw, gb = torch.empty_like(weight), torch.zeros_like(bias)      # d_weight is overwritten by the entry (it zeroes it itself)
        # the layer behind the InputLayer has no use for d_input (point features are data): skip that product
        gx = grad_out.new_empty(0) if ctx.needs_input_grad[0] else None
        SCN.bf16_operand_again(ctx.scn_meta, x, ctx.x16)
        SCN.SubmanifoldConvolution_backward(spatial_size, filter_size, ctx.scn_meta, x, gx, grad_out.contiguous(),
                                            weight, gw, gb, ctx.dilated_rate)
        del ctx.scn_meta, ctx.x16
        return (gx, gw, optionalTensorReturn(gb), None, None, None, None, None,
                (grad_out if ctx.needs_input_grad[8] else None), None)


class _StridedFunction(Function):
    """Shared body of Convolution / Deconvolution (they differ only in the SCN entry they call)."""
    fwd = bwd = None

    @classmethod
    def _forward(cls, ctx, x, weight, bias, metadata, in_size, out_size, dimension, filter_size, filter_stride):
        ctx.scn_meta = metadata
        ctx.save_for_backward(x, in_size, weight, bias, out_size, filter_size, filter_stride)
        out = x.new_empty(0)
        ctx.x16 = SCN.bf16_operand(metadata, x, weight.size(1), weight.size(2))
        _count(cls.fwd(in_size, out_size, filter_size, filter_stride, metadata, x, out, weight, bias), out)
        return out

    @classmethod
    def _backward(cls, ctx, grad_out):
        x, in_size, weight, bias, out_size, filter_size, filter_stride = ctx.saved_tensors
        gx, gw, gb = grad_out.new_empty(0), torch.empty_like(weight), torch.zeros_like(bias)
        SCN.bf16_operand_again(ctx.scn_meta, x, ctx.x16)
        cls.bwd(in_size, out_size, filter_size, filter_stride, ctx.scn_meta, x, gx, grad_out.contiguous(), weight, gw, gb)
        del ctx.scn_meta, ctx.x16
        return gx, gw, optionalTensorReturn(gb), None, None, None, None, None, None


class ConvolutionFunction(_StridedFunction):
    fwd, bwd = staticmethod(SCN.Convolution_updateOutput), staticmethod(SCN.Convolution_backward)

    @staticmethod
    def forward(ctx, *a):
        return ConvolutionFunction._forward(ctx, *a)

    @staticmethod
    def backward(ctx, g):
        return ConvolutionFunction._backward(ctx, g)


class DeconvolutionFunction(_StridedFunction):
    fwd, bwd = staticmethod(SCN.Deconvolution_updateOutput), staticmethod(SCN.Deconvolution_backward)

    @staticmethod
    def forward(ctx, *a):
        return DeconvolutionFunction._forward(ctx, *a)

    @staticmethod
    def backward(ctx, g):
        return DeconvolutionFunction._backward(ctx, g)


class BatchNormalizationFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias, running_mean, running_var, eps, momentum, train, leakiness, with_alias=False):
        """Returns (out, out16, alias).  out16: bf16 copy of out (or empty).  alias (with_alias): a view of x for the
        shortcut branch of a residual block -- whatever gradient comes back through it is added to d_x inside the
        backward kernel instead of by a separate accumulation pass."""
        ctx.train, ctx.leakiness = train, leakiness
        n_planes = running_mean.shape[0]
        out = x.new_empty(0)
        save_mean, save_invstd = x.new_empty(n_planes), x.new_empty(n_planes)
        # second, non-differentiable output: the bf16 copy of `out` for the tensor-core convolution that follows
        out16 = torch.empty(0, dtype=torch.bfloat16, device=x.device)
        SCN.BatchNormalization_updateOutput(x, out, save_mean, save_invstd, running_mean, running_var, weight, bias,
                                            eps, momentum, train, leakiness,
                                            out16 if SCN.wants_bf16(n_planes) and x.is_cuda else None,
                                            SCN.held_stats(x) if train else None)
        # the backward pass recomputes the activation mask from x, so `out` is not kept alive for it
        ctx.save_for_backward(x, weight, bias, running_mean, running_var, save_mean, save_invstd)
        ctx.mark_non_differentiable(out16)
        return out, out16, (x.view_as(x) if with_alias else x.new_empty(0))

    @staticmethod
    def backward(ctx, grad_out, _grad_out16=None, grad_alias=None):
        x, weight, bias, running_mean, running_var, save_mean, save_invstd = ctx.saved_tensors
        assert ctx.train, "BatchNormalization backward is only defined in training mode (as in the reference)"
        # d_gamma / d_beta are written for every channel by the entry (unless there are no rows at all)
        fresh = torch.empty_like if x.size(0) else torch.zeros_like
        gx, gw, gb = grad_out.new_empty(0), fresh(weight), fresh(bias)
        add = grad_alias.contiguous() if grad_alias is not None and grad_alias.numel() == x.numel() else None
        SCN.BatchNormalization_backward(x, gx, None, grad_out.contiguous(), save_mean, save_invstd, running_mean,
                                        running_var, weight, bias, gw, gb, ctx.leakiness, add)
        return gx, optionalTensorReturn(gw), optionalTensorReturn(gb), None, None, None, None, None, None, None


class BatchNormConvFunction(Function):
    """Training-mode BatchNorm(+ReLU) followed by a convolution (SubmanifoldConvolution / Convolution / Deconvolution) as
    ONE autograd node -- the UNet's BN -> ReLU -> conv pattern (networkArchitectures.py:225-229).

    forward : the BatchNorm writes only the bf16 operand the convolution gathers (the fp32 activation is never
              materialised: 6 instead of 10 bytes per element), the convolution runs on it (+ residual / statistics
              epilogue as in SubmanifoldConvolutionFunction);
    backward: the convolution's dgrad epilogue recomputes the activation mask from the BatchNorm's input, writes the masked
              gradient and accumulates the BatchNorm's two backward column sums (its reduction pass disappears); one
              apply pass finishes d_x (+ the gradient arriving through a residual shortcut), d_gamma, d_beta.
    Same arithmetic as the two separate layers up to fp32 summation order of the column sums."""

    KINDS = {
        "subm": (SCN.SubmanifoldConvolution_updateOutput, SCN.SubmanifoldConvolution_backward),
        "conv": (SCN.Convolution_updateOutput, SCN.Convolution_backward),
        "deconv": (SCN.Deconvolution_updateOutput, SCN.Deconvolution_backward),
    }

    @staticmethod
    def forward(ctx, x, bn_weight, bn_bias, running_mean, running_var, eps, momentum, leakiness, conv_weight, conv_bias,
                metadata, kind, in_size, out_size, filter_size, filter_stride, residual=None, want_stats=False,
                with_alias=False):
        """Returns (out, stats, alias) -- see SubmanifoldConvolutionFunction / BatchNormalizationFunction."""
        n_planes = running_mean.shape[0]
        save_mean, save_invstd = x.new_empty(n_planes), x.new_empty(n_planes)
        y16 = torch.empty(0, dtype=torch.bfloat16, device=x.device)
        SCN.BatchNormalization_updateOutput(x, None, save_mean, save_invstd, running_mean, running_var, bn_weight, bn_bias,
                                            eps, momentum, True, leakiness, y16, SCN.held_stats(x))
        out = x.new_empty(0)
        n_out = conv_weight.size(2)
        stats = torch.empty((2, n_out) if want_stats else 0, dtype=torch.float64, device=x.device)
        fwd = BatchNormConvFunction.KINDS[kind][0]
        if kind == "subm":
            macs = fwd(in_size, filter_size, metadata, None, out, conv_weight, conv_bias, 1, residual,
                       stats if want_stats else None, input_bf16=y16)
        else:
            macs = fwd(in_size, out_size, filter_size, filter_stride, metadata, None, out, conv_weight, conv_bias,
                       input_bf16=y16, stats=stats if want_stats else None)
        _count(macs, out)
        ctx.scn_meta, ctx.kind, ctx.leakiness = metadata, kind, leakiness
        ctx.save_for_backward(x, bn_weight, bn_bias, save_mean, save_invstd, y16, conv_weight, conv_bias, in_size, out_size,
                              filter_size, filter_stride)
        ctx.mark_non_differentiable(stats)
        return out, stats, (x.view_as(x) if with_alias else x.new_empty(0))

    @staticmethod
    def backward(ctx, grad_out, _grad_stats=None, grad_alias=None):
        (x, bn_weight, bn_bias, save_mean, save_invstd, y16, conv_weight, conv_bias, in_size, out_size, filter_size,
         filter_stride) = ctx.saved_tensors
        m, kind = ctx.scn_meta, ctx.kind
        g = grad_out          # may be a column slice of a JoinTable's gradient: read in place (SCN._grad_operand)
        acc = torch.empty((2, x.size(1)), dtype=torch.float64, device=x.device)      # zeroed by the convolution entry
        d_masked = g.new_empty(0)
        gw, gb = torch.empty_like(conv_weight), torch.zeros_like(conv_bias)
        SCN.BatchNormalization_backwardFusion(m, x, save_mean, save_invstd, bn_weight, bn_bias, ctx.leakiness, acc)
        bwd = BatchNormConvFunction.KINDS[kind][1]
        if kind == "subm":
            bwd(in_size, filter_size, m, None, d_masked, g, conv_weight, gw, gb, 1, input_bf16=y16)
        else:
            bwd(in_size, out_size, filter_size, filter_stride, m, None, d_masked, g, conv_weight, gw, gb, input_bf16=y16)
        gx = g.new_empty(0)
        fresh = torch.empty_like if x.size(0) else torch.zeros_like      # written for every channel by the entry
        g_gamma, g_beta = fresh(bn_weight), fresh(bn_bias)
        add = grad_alias if grad_alias is not None and grad_alias.numel() == x.numel() else None
        # the gradient leaves with a bf16 copy attached: the convolution that produced x receives it as its d_out and reads the
        # copy instead of casting (SCN._grad_operand); dropped silently if autograd adds another gradient to it on the way
        gx16 = torch.empty(0, dtype=torch.bfloat16, device=x.device) if SCN.wants_bf16(x.size(1)) else None
        SCN.BatchNormalization_backwardApply(x, d_masked, acc, save_mean, save_invstd, bn_weight, gx, g_gamma, g_beta, add, gx16)
        if gx16 is not None:
            SCN.attach_bf16(gx, gx16)
        del ctx.scn_meta
        return (gx, optionalTensorReturn(g_gamma), optionalTensorReturn(g_beta), None, None, None, None, None, gw,
                optionalTensorReturn(gb), None, None, None, None, None, None,
                (grad_out if ctx.needs_input_grad[16] else None), None, None)


class NetworkInNetworkFunction(Function):
    @staticmethod
    def forward(ctx, x, weight, bias):
        out = x.new_empty(0)
        ctx.save_for_backward(x, weight, bias)
        _count(SCN.NetworkInNetwork_updateOutput(x, out, weight, bias), out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        x, weight, bias = ctx.saved_tensors
        if grad_out.stride(-1) != 1:
            grad_out = grad_out.contiguous()       # row-strided slices are fine for the GEMMs as they are
        gx, gw = grad_out.new_empty(0), torch.zeros_like(weight)
        gb = torch.zeros_like(bias) if bias is not None and bias.numel() else None
        SCN.NetworkInNetwork_updateGradInput(gx, grad_out, weight)
        SCN.NetworkInNetwork_accGradParameters(x, grad_out, gw, gb)
        return gx, gw, gb


class InputLayerFunction(Function):
    @staticmethod
    def forward(ctx, dimension, metadata, spatial_size, coords, features, batch_size, mode, normals,
                normal_guide_scale=10241):
        ctx.scn_meta = metadata
        metadata.setNormalGuideScale(normal_guide_scale)
        out = features.new_empty(0)
        SCN.InputLayer_updateOutput(metadata, spatial_size, coords, features.contiguous(), out, batch_size, mode,
                                    normals)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        gx = grad_out.new_empty(0)
        SCN.InputLayer_updateGradInput(ctx.scn_meta, gx, grad_out.contiguous())
        del ctx.scn_meta
        return None, None, None, None, gx, None, None, None, None


class OutputLayerFunction(Function):
    @staticmethod
    def forward(ctx, dimension, metadata, features):
        ctx.scn_meta = metadata
        out = features.new_empty(0)
        SCN.OutputLayer_updateOutput(metadata, features.contiguous(), out)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        gx = grad_out.new_empty(0)
        SCN.OutputLayer_updateGradInput(ctx.scn_meta, gx, grad_out.contiguous())
        del ctx.scn_meta
        return None, None, gx
