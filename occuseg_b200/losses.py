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


# ---- the embedding losses of examples/ScanNet/discriminative.py, torch_scatter-free and without per-instance Python loops -----
class DiscriminativeLoss(torch.nn.Module):
    """discriminative.py:117-226 (the paths `forward` takes: _new_centroids, _new_variance, _distance, _regularization).
    forward(embedded [B,P,E], instance_mask [B,P] long, dense 0-based ids) -> alpha*L_v + beta*L_d + gamma*L_r.
    Same constructor and the same value as the reference; centroids come from one index_add per sample instead of
    torch_scatter.scatter_mean."""

    def __init__(self, delta_d, delta_v, alpha=1.0, beta=1.0, gamma=0.001, reduction='mean'):
        super().__init__()
        self.alpha, self.beta, self.gamma = alpha, beta, gamma
        self.delta_d, self.delta_v = delta_d, delta_v

    def forward(self, embedded, instance_mask):
        B = embedded.size(0)
        l_v = embedded.new_zeros(())
        l_d = embedded.new_zeros(())
        l_r = embedded.new_zeros(())
        k_all = int(instance_mask.max().item()) + 1          # scatter_mean sizes every sample's centroid table alike (torch.stack)
        for i in range(B):
            e, ids = embedded[i], instance_mask[i].view(-1)
            mu = segment_mean(e, ids, k_all)
            n = int(ids.max().item()) + 1                    # size = max(instance_mask) + 1, :134
            dev = (e - mu[ids]).norm(2, dim=1)
            l_v = l_v + (torch.clamp(dev - self.delta_v, min=0.0) ** 2).mean()                       # :174-181
            c = mu[:n]
            if n > 1:                                                                                # :200-216
                norm = (c.unsqueeze(1) - c.unsqueeze(0)).norm(2, dim=2)
                margin = 2 * self.delta_d * (1.0 - torch.eye(n, device=c.device, dtype=c.dtype))
                l_d = l_d + torch.sum(torch.clamp(margin - norm, min=0.0) ** 2) / float(n * (n - 1))
            l_r = l_r + c.norm(2, dim=1).mean()                                                      # :218-226
        return self.alpha * l_v + self.beta * (l_d / B) + self.gamma * (l_r / B)


def ClassificationLoss(embedded, bw, regressed_pose, pose, instance_mask, pred_semantics, min_points=30):
    """discriminative.py:41-114, same arguments ([B,P,*] tensors, instance_mask [B,P], pred_semantics [P] = the class of every
    point of the -- single -- sample the caller passes), returns (loss [1], mean instance IoU).

    Per instance with at least 30 points whose first point's class is > -1: the points closer to the instance's mean position than
    4x its radius are classified as inside / outside by exp(-(|e - mu| * s1)^2 - (|q - c| * s2)^2) (mu, c = instance means of the
    embedding / the position, s1, s2 = instance means of the two bandwidth channels, q = regressed position) with a BCE loss;
    loss = 10 * mean over instances.  The reference loops over instances in Python; here every sample is one [K,P] problem."""
    B = embedded.size(0)
    dev = embedded.device
    total = embedded.new_zeros(1)
    miou = 0.0
    count = 0
    E = embedded.size(2)
    for i in range(B):
        ids = instance_mask[i].view(-1)
        K = int(ids.max().item()) + 1
        P = ids.numel()
        volume = torch.cat((embedded[i], pose[i], bw[i], regressed_pose[i]), dim=1)
        mean = segment_mean(volume, ids, K)
        mu, centre, s1, s2 = mean[:, :E], mean[:, E:E + 3], mean[:, E + 3], mean[:, E + 4]
        n_pts = segment_count(ids, K)
        first = torch.full((K,), P, dtype=torch.long, device=dev).scatter_reduce(0, ids, torch.arange(P, device=dev), "amin")
        cls = pred_semantics[first.clamp_max(P - 1)]
        valid = (n_pts >= min_points) & (cls > -1) & (first < P)
        if not bool(valid.any()):
            continue
        vk = torch.nonzero(valid).view(-1)
        exact = "donot_use_mm_for_euclid_dist"
        dist = torch.cdist(centre[vk], pose[i], compute_mode=exact)                                   # [k,P] |pose - mean_pose|
        own = ids.view(1, -1) == vk.view(-1, 1)                                                       # [k,P] instance_indices
        radius = torch.where(own, dist, dist.new_full((), float("-inf"))).max(dim=1).values
        samples = dist < (radius * 4).view(-1, 1)
        d1 = torch.cdist(mu[vk], embedded[i], compute_mode=exact) * s1[vk].view(-1, 1)
        d2 = torch.cdist(centre[vk], regressed_pose[i], compute_mode=exact) * s2[vk].view(-1, 1)
        prob = torch.exp(-d1 * d1 - d2 * d2)
        bce = torch.nn.functional.binary_cross_entropy(prob, own.to(prob.dtype), reduction="none")
        w = samples.to(prob.dtype)
        total = total + ((bce * w).sum(1) / w.sum(1)).sum()
        with torch.no_grad():
            u = (prob > 0.5) & samples
            tp = (u & own).sum(1).to(torch.float64)
            fp = (u & ~own).sum(1).to(torch.float64)
            tot = (own & samples).sum(1).to(torch.float64)
            miou += float((tp / (tot + fp)).sum().item())
        count += int(vk.numel())
    if count > 0:
        total = total / count * 10
        miou = miou / count
    return total, miou


def calculate_cost(predictions, embeddings, offsets, displacements, bw, criterion, batch, occupancy, config):
    """calculate_cost of examples/ScanNet/train_instance.py:186-255 on this module's pieces: same arguments (+ `config` for
    'scale', 'dimension', 'regress_weight', 'displacement_weight', a module global there), same dictionary of losses.
    `batch` as the reference's data loader builds it: 'x' = [coords [P,4], feats], 'y' [P,2] (semantic, instance), 'id',
    'instance_masks' [P], 'instance_sizes' [P], 'displacements' [P,3], 'offsets' [P,...]."""
    dev = predictions.device
    y = batch['y'].to(dev)
    sem = y[:, 0]
    coords = batch['x'][0].to(dev)
    sample = coords[:, config['dimension']].long()
    B = len(batch['id'])
    inst = batch['instance_masks'].to(dev).long()
    pose = coords[:, 0:3].to(displacements.dtype) / config['scale']
    regressed_pose = pose - displacements
    displacements_gt = batch['displacements'].to(dev)
    occupancy_gt = batch['instance_sizes'].to(dev).view(-1, 1)
    out = {}
    out['semantic_loss'] = criterion['nll'](predictions, sem)
    emb = predictions.new_zeros(1)
    cls_loss = predictions.new_zeros(1)
    iou = predictions.new_zeros(1)
    for b in range(B):
        idx = sample == b
        e = embeddings[idx].unsqueeze(0)
        im = inst[idx].view(1, -1)
        emb = emb + criterion['discriminative'](e, im)
        lc, ii = ClassificationLoss(e, bw[idx].view(1, -1, 2), regressed_pose[idx].unsqueeze(0), pose[idx].unsqueeze(0), im, sem[idx])
        cls_loss = cls_loss + lc
        iou = iou + ii
    d_loss, o_loss = cluster_regression_losses(displacements, displacements_gt, occupancy, occupancy_gt, inst, sample, sem, B)
    fg = sem > 1
    out['embedding_loss'] = emb / B
    out['regression_loss'] = criterion['regression'](offsets[fg], batch['offsets'].to(dev)[fg]) * config['regress_weight']
    out['displacement_loss'] = d_loss.view(1)
    out['classification_loss'] = cls_loss / B
    out['drift_loss'] = predictions.new_zeros(1)
    out['instance_iou'] = iou / B
    out['occupancy_loss'] = o_loss.view(1)
    return out
