"""`sparseconvnet` surface of the B200 hot path (mirror of sparseconvnet/__init__.py:9-37, hot subset only).

    import occuseg_b200.sparseconvnet as scn          # or occuseg_b200.install_as_sparseconvnet()
"""
from . import SCN
from .SCN import set_precision, get_precision, set_deterministic
from .architectures import UNet
from .functions import counters
from .layers import (AddTable, BatchNormalization, BatchNormLeakyReLU, BatchNormReLU, ConcatTable, Convolution,
                     Deconvolution, Identity, InputLayer, JoinTable, Metadata, NetworkInNetwork, OutputLayer,
                     Sequential, SubmanifoldConvolution, ValidConvolution)
from .tensor import SparseConvNetTensor
from .utils import optionalTensor, optionalTensorReturn, toLongTensor, upsample_feature

forward_pass_multiplyAdd_count = 0
forward_pass_hidden_states = 0
