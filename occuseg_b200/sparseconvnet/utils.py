"""Helpers with the reference's names: sparseconvnet/utils.py:13-28 (tensor plumbing) and :72-132 (`upsample_feature`, the
caller of SCN.ResolutionBasedScattering used by the DenseUNet variants of examples/ScanNet/model.py)."""
import torch


def toLongTensor(dimension, x):
    if isinstance(x, torch.Tensor) and x.dtype == torch.int64 and not x.is_cuda:
        return x
    if isinstance(x, (list, tuple)):
        assert len(x) == dimension
        return torch.LongTensor(list(x))
    return torch.LongTensor(dimension).fill_(int(x))


def optionalTensor(obj, name):
    return getattr(obj, name) if hasattr(obj, name) else torch.Tensor()


def optionalTensorReturn(t):
    return t if t.numel() else None


def upsample_feature(lr, hr, stride, bilinear=False):
    """Features of the low-resolution tensor `lr` carried to the active sites of the high-resolution tensor `hr`
    (sparseconvnet/utils.py:72-132).  Nearest mode: every hr voxel takes the row of lr voxel (hr // stride) -- rows without a
    parent take lr row 0 of their sample, as the reference's `correspondence[correspondence < 0] = 0` does.  Bilinear mode:
    the 8 lr voxels around (hr - (stride-1)/2) / stride, weighted trilinearly and renormalised over the ones that exist."""
    from . import SCN
    from .tensor import SparseConvNetTensor
    dev = lr.features.device
    loc_lr = lr.get_spatial_locations().to(dev).int()
    loc_hr = hr.get_spatial_locations().to(dev).int()
    batch_size = int(loc_hr[:, 3].max().item()) + 1
    out = SparseConvNetTensor(metadata=hr.metadata, spatial_size=hr.spatial_size)
    pieces, lr_start = [], 0
    for k in range(batch_size):
        sel_lr, sel_hr = loc_lr[:, 3] == k, loc_hr[:, 3] == k
        p_lr, p_hr = loc_lr[sel_lr, 0:3], loc_hr[sel_hr, 0:3]
        if not bilinear:
            corr = SCN.ResolutionBasedScattering(lr.metadata, p_lr, p_hr, stride).long() + lr_start
            corr[corr < 0] = 0                                   # as the reference (utils.py:89): applied AFTER the offset
            pieces.append(torch.index_select(lr.features, 0, corr))
        else:
            cand = (p_hr.float() - (stride - 1) / 2) / stride
            anchors = [torch.ceil(cand), torch.floor(cand)]
            diff = [anchors[0] - cand, cand - anchors[1]]
            weights, queries = [], []
            for x in (0, 1):
                for y in (0, 1):
                    for z in (0, 1):
                        weights.append((1 - diff[x][:, 0]) * (1 - diff[y][:, 1]) * (1 - diff[z][:, 2]))
                        queries.append(torch.stack([anchors[x][:, 0], anchors[y][:, 1], anchors[z][:, 2]], 1).int())
            weight, query = torch.cat(weights, 0), torch.cat(queries, 0)
            corr = SCN.ResolutionBasedScattering(lr.metadata, p_lr, query, 1).long()
            weight = torch.where(corr < 0, torch.zeros_like(weight), weight)
            corr = corr.clamp_min(0) + lr_start
            n = p_hr.shape[0]
            feats = (torch.index_select(lr.features, 0, corr) * weight[:, None]).view(8, n, -1).sum(0)
            pieces.append(feats / weight.view(8, n).sum(0)[:, None])
        lr_start += int(sel_lr.sum().item())
    out.features = torch.cat(pieces, 0)
    return out
