"""SparseConvNetTensor: features [N,C] + the batch's Metadata handle + the spatial size of the scale
(mirror of the reference container, sparseconvnet/sparseConvNetTensor.py:13-66)."""
import torch


class SparseConvNetTensor(object):
    def __init__(self, features=None, metadata=None, spatial_size=None):
        self.features = features
        self.metadata = metadata
        self.spatial_size = spatial_size

    def get_spatial_locations(self, spatial_size=None):
        """Coordinates and batch index of the active rows, LongTensor [N,4]."""
        return self.metadata.getSpatialLocations(self.spatial_size if spatial_size is None else spatial_size)

    def type(self, t=None):
        if t:
            self.features = self.features.type(t)
            return self
        return self.features.type()

    def cuda(self):
        self.features = self.features.cuda()
        return self

    def cpu(self):
        self.features = self.features.cpu()
        return self

    def __repr__(self):
        shape = None if self.features is None else tuple(self.features.shape)
        return f"SparseConvNetTensor<<features.shape={shape}, spatial size={self.spatial_size}>>"
