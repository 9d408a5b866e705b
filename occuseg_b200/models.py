"""OccuSeg's networks on top of the B200 sparse-convolution path (SURVEY.md section 8f, row 1 -- first step of the
widening: the dense heads; the losses of train_instance.py are not part of this round).

Attribute names and layer shapes follow examples/ScanNet/model.py:657-717 (`InstanceDenseUNet`, `LearningBWDenseUNet`), so
a reference checkpoint's tensors load by name.  The heads are plain `nn.Linear` layers on the [P, m] point features
(library GEMMs, as in the reference); everything sparse goes through `occuseg_b200.sparseconvnet`."""
import torch.nn as nn

from . import sparseconvnet as scn


class InstanceDenseUNet(nn.Module):
    """config keys used (same names as the reference config dict): dimension, full_scale, rotation_guide_level,
    input_feature_number, unet_structure, block_reps, residual_blocks, class_num."""

    def __init__(self, config):
        super().__init__()
        self.config = config
        d, m = config["dimension"], config["unet_structure"][0]
        self.input = scn.InputLayer(d, config["full_scale"], mode=4,
                                    normal_guide_scale=(config["full_scale"] >> config.get("rotation_guide_level", 0)) + 1)
        self.sub = scn.SubmanifoldConvolution(d, config["input_feature_number"], m, 3, False)
        self.unet = scn.UNet(d, config["block_reps"], config["unet_structure"], config["residual_blocks"])
        self.output_feature_dim = m
        self.bn = scn.BatchNormReLU(m)
        self.output = scn.OutputLayer(d)
        self.linear = nn.Linear(m, config["class_num"])
        self.fc_regress = nn.Linear(m, m)
        self.linear_regress = nn.Linear(m, 1)
        self.sigmoid_regress = nn.Sigmoid()
        self.fc_embedding = nn.Linear(m, m)
        self.linear_embedding = nn.Linear(m, m)
        self.fc_displacement = nn.Linear(m, m)
        self.linear_displacement = nn.Linear(m, d)

    def forward(self, x):
        feature = self.output(self.bn(self.unet(self.sub(self.input(x)))))
        y = self.linear(feature)
        embedding = self.linear_embedding(self.fc_embedding(feature))
        offset = self.sigmoid_regress(self.linear_regress(self.fc_regress(feature)))
        displacement = self.linear_displacement(self.fc_displacement(feature))
        return y, feature, embedding, offset, displacement


class LearningBWDenseUNet(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.config = config
        self.backbone = InstanceDenseUNet(config)
        m = self.backbone.output_feature_dim
        self.fc_bw = nn.Linear(m, m)
        self.linear_bw = nn.Linear(m, 2)
        self.relu_bw = nn.Softplus()
        self.fc_occupancy = nn.Linear(m, m)
        self.linear_occupancy = nn.Linear(m, 1)
        self.relu_occupancy = nn.Softplus()

    def forward(self, x):
        semantics, feature, embedding, offset, displacement = self.backbone(x)
        bw = self.relu_bw(self.linear_bw(self.fc_bw(feature)))
        occupancy = self.relu_occupancy(self.linear_occupancy(self.fc_occupancy(feature)))
        return semantics, feature, embedding, offset, displacement, bw, occupancy


def default_config(m=64, levels=6, class_num=20):
    """The shipped ScanNet configuration (examples/ScanNet/config: baseline_m64) reduced to the keys the models read."""
    return {"dimension": 3, "full_scale": 4096, "rotation_guide_level": 0, "input_feature_number": 3,
            "unet_structure": [m * (i + 1) for i in range(levels)], "block_reps": 1, "residual_blocks": True,
            "class_num": class_num}
