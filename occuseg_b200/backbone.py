"""The sparse backbone of OccuSeg's InstanceDenseUNet (examples/ScanNet/model.py:657-691) without the dense
nn.Linear heads (SURVEY.md section 8f.1 -- next scope):

    InputLayer(3, full_scale, mode=4) -> SubmanifoldConvolution(3, in_ch -> m, 3, bias=False)
      -> scn.UNet(3, block_reps, [m, 2m, ..., levels*m], residual_blocks) -> BatchNormReLU(m) -> OutputLayer

Attribute names (`input`, `sub`, `unet`, `bn`, `output`) match the reference module so a reference
state_dict's backbone entries load by name."""
import torch.nn as nn

from . import sparseconvnet as scn


class SparseBackbone(nn.Module):
    def __init__(self, m=64, levels=6, block_reps=1, residual_blocks=True, input_channels=3, full_scale=4096,
                 dimension=3):
        super().__init__()
        planes = [m * (i + 1) for i in range(levels)]
        self.input = scn.InputLayer(dimension, full_scale, mode=4)
        self.sub = scn.SubmanifoldConvolution(dimension, input_channels, m, 3, False)
        self.unet = scn.UNet(dimension, block_reps, planes, residual_blocks)
        self.bn = scn.BatchNormReLU(m)
        self.output = scn.OutputLayer(dimension)

    def forward(self, x):
        """x = [coords [P,4], feats [P,C] cuda, normals-or-None, batch_size]; returns point features [P, m]."""
        return self.output(self.bn(self.unet(self.sub(self.input(x)))))
