"""scn.UNet builder with the same module tree (hence the same state_dict keys) as the reference
(sparseconvnet/networkArchitectures.py:202-306): per level  [BN-ReLU, SubmConv, BN-ReLU, SubmConv (+ identity or
1x1 NiN shortcut)] x reps, then BN-ReLU -> Convolution 2/2 -> U(next level) -> BN-ReLU -> Deconvolution 2/2,
joined with the skip path and folded back from 2c to c channels by the first block on the way up."""
from . import layers as L


def _block(seq, dimension, a, b, residual, leakiness):
    def bn_conv(cin, cout):
        return [L.BatchNormLeakyReLU(cin, leakiness=leakiness), L.SubmanifoldConvolution(dimension, cin, cout, 3, False)]

    if residual:
        body = L.Sequential()
        for m in bn_conv(a, b) + bn_conv(b, b):
            body.add(m)
        shortcut = L.Identity() if a == b else L.NetworkInNetwork(a, b, False)
        seq.add(L.ResidualConcatTable().add(shortcut).add(body)).add(L.AddTable())
    else:
        body = L.Sequential()
        for m in bn_conv(a, b):
            body.add(m)
        seq.add(body)


def UNet(dimension, reps, nPlanes, residual_blocks=False, downsample=[2, 2], leakiness=0):
    def level(planes):
        seq = L.Sequential()
        c = planes[0]
        for _ in range(reps):
            _block(seq, dimension, c, c, residual_blocks, leakiness)
        if len(planes) > 1:
            inner = (L.Sequential()
                     .add(L.BatchNormLeakyReLU(c, leakiness=leakiness))
                     .add(L.Convolution(dimension, c, planes[1], downsample[0], downsample[1], False))
                     .add(level(planes[1:]))
                     .add(L.BatchNormLeakyReLU(planes[1], leakiness=leakiness))
                     .add(L.Deconvolution(dimension, planes[1], c, downsample[0], downsample[1], False)))
            seq.add(L.ConcatTable().add(L.Identity()).add(inner))
            seq.add(L.JoinTable())
            for i in range(reps):
                _block(seq, dimension, c * (2 if i == 0 else 1), c, residual_blocks, leakiness)
        return seq

    return level(list(nPlanes))
