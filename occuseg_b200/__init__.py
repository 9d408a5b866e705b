"""occuseg_b200: B200-native (sm_100a) implementation of OccuSeg's SparseConvNet hot path --
rulebook construction, submanifold / strided sparse convolution fwd+dgrad+wgrad, BatchNorm+ReLU --
behind the reference's `sparseconvnet` Python API.  See DESIGN.md."""
import sys

__version__ = "0.1.0"


def install_as_sparseconvnet():
    """Make `import sparseconvnet` resolve to this package's mirror, so reference model code
    (examples/ScanNet/model.py) runs unchanged."""
    from . import sparseconvnet as scn
    sys.modules["sparseconvnet"] = scn
    sys.modules["sparseconvnet.SCN"] = scn.SCN
    return scn
