"""torch_scatter-free pieces of the OccuSeg training loss.

The reference's `calculate_cost` (examples/ScanNet/train_instance.py:186-255) needs `torch_scatter`'s scatter_mean /
scatter_std / scatter_add / scatter_max (absent here, and a Python loop over samples and instances on top).  This module
provides the segment reductions on plain torch index ops (one pass each, any device) and the two per-instance regression
terms of that function -- DisplacementLoss and OccupancyLoss -- vectorised over the whole batch with the same result as the
reference's double loop.  The semantic NLL and the foreground L1 terms are single torch calls in the reference and stay so;
the discriminative embedding loss (discriminative.py) and ClassificationLoss are not ported."""
import torch


def segment_sum(values, ids, n):
    """out[k] = sum of values[i] with ids[i] == k   (torch_scatter.scatter_add along dim 0)"""
    out = values.new_zeros((n,) + tuple(values.shape[1:]))
    return out.index_add_(0, ids, values)


def segment_count(ids, n, dtype=torch.float32):
    return torch.zeros(n, dtype=dtype, device=ids.device).index_add_(0, ids, torch.ones_like(ids, dtype=dtype))


def segment_mean(values, ids, n):
    """torch_scatter.scatter_mean: empty segments give 0"""
    cnt = segment_count(ids, n, values.dtype).clamp_min(1)
    return segment_sum(values, ids, n) / cnt.view((-1,) + (1,) * (values.dim() - 1))


def segment_std(values, ids, n, unbiased=True):
    """torch_scatter.scatter_std (unbiased by default): sqrt(sum (v - mean)^2 / (count - 1)); segments with one element give 0"""
    cnt = segment_count(ids, n, values.dtype)
    shape = (-1,) + (1,) * (values.dim() - 1)
    mean = segment_sum(values, ids, n) / cnt.clamp_min(1).view(shape)
    var = segment_sum((values - mean[ids]) ** 2, ids, n) / (cnt - (1 if unbiased else 0)).clamp_min(1).view(shape)
    return var.sqrt()


def segment_max(values, ids, n):
    """(max, argmax) per segment like torch_scatter.scatter_max along dim 0; empty segments: (0, -1)"""
    out = values.new_full((n,) + tuple(values.shape[1:]), float("-inf"))
    out = out.scatter_reduce(0, ids.view((-1,) + (1,) * (values.dim() - 1)).expand_as(values), values, "amax", include_self=True)
    hit = values == out[ids]
    idx = torch.arange(values.size(0), device=values.device).view((-1,) + (1,) * (values.dim() - 1)).expand_as(values)
    arg = torch.full_like(out, -1, dtype=torch.long).scatter_reduce(0, ids.view((-1,) + (1,) * (values.dim() - 1)).expand_as(values),
                                                                    torch.where(hit, idx, torch.full_like(idx, -1)), "amax",
                                                                    include_self=True)
    out = torch.where(torch.isinf(out), torch.zeros_like(out), out)
    return out, arg


def cluster_regression_losses(displacements, displacements_gt, occupancy, occupancy_gt, instance_ids, sample_ids, semantics,
                              batch_size):
    """DisplacementLoss and OccupancyLoss of calculate_cost (train_instance.py:198-236), whole batch at once.

    displacements [P,3], occupancy [P,1] predictions and their targets; instance_ids [P] (per-sample instance index, 0-based,
    `batch['instance_masks']`); sample_ids [P] (4th coordinate column); semantics [P] (`batch['y'][:,0]`).
    Per sample: for every instance whose FIRST point has class > 1 (:219-221), the mean displacement error norm, the mean
    |occupancy error| and the unbiased std of the predicted occupancy are summed and divided by the number of such instances;
    the per-sample values are averaged over `batch_size` (:241-243)."""
    dev = displacements.device
    n_inst = int(instance_ids.max().item()) + 1 if instance_ids.numel() else 0
    seg = sample_ids.long() * n_inst + instance_ids.long()          # one segment per (sample, instance)
    n_seg = batch_size * n_inst
    disp_err = segment_mean((displacements - displacements_gt).norm(dim=1), seg, n_seg)
    occ_err = segment_mean((occupancy - occupancy_gt).norm(dim=1), seg, n_seg)
    occ_std = segment_std(occupancy, seg, n_seg).view(n_seg, -1).sum(1)
    # class of the first point of every segment
    first = torch.full((n_seg,), semantics.numel(), dtype=torch.long, device=dev)
    first = first.scatter_reduce(0, seg, torch.arange(semantics.numel(), device=dev), "amin", include_self=True)
    present = first < semantics.numel()
    cls = torch.zeros(n_seg, dtype=semantics.dtype, device=dev)
    cls[present] = semantics[first[present]]
    fg = (present & (cls > 1)).to(disp_err.dtype).view(batch_size, n_inst)
    k = fg.sum(1)
    per_sample_d = (disp_err.view(batch_size, n_inst) * fg).sum(1) / k.clamp_min(1)
    per_sample_o = ((occ_err + occ_std).view(batch_size, n_inst) * fg).sum(1) / k.clamp_min(1)
    has = (k > 0).to(disp_err.dtype)
    return (per_sample_d * has).sum() / batch_size, (per_sample_o * has).sum() / batch_size
