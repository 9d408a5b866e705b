"""The step before the hot path, on the device: augmentation of a float point cloud and its
conversion to the InputLayer's coordinate list without the reference's host round trip.

Mirrors examples/ScanNet/datasets/scannet.py:49-160: random affine (rotation noise, x flip, scale, rotation about z,
:103-110), two elastic distortions (blurred noise grids interpolated at the points, :49-70, :121-124), shift so the cloud
starts at (10,10,10) plus a random sub-voxel offset, crop to [0, full_scale) (:133-137,160).  The affine and the elastic
field are small library-op pipelines on the GPU (matmul, three separable 3-tap convolutions, trilinear grid_sample); the
final shift / crop / truncation to int64 is one kernel of libscn_b200.so (scn_float_coords) whose output is what
`scn.InputLayer` takes as a device-resident coordinate list -- the points never visit the host."""
import ctypes as C
import math

import torch
import torch.nn.functional as F

from . import _lib


def random_affine(scale, rotation_noise=True, generator=None, device="cpu"):
    """m of scannet.py:103-110: (I + 0.1*N(0,1)) with a random x flip, times `scale`, times a random rotation about z."""
    g = generator
    m = torch.eye(3, dtype=torch.float64)
    if rotation_noise:
        m = m + torch.randn(3, 3, generator=g, dtype=torch.float64) * 0.1
    m[0, 0] *= float(torch.randint(0, 2, (1,), generator=g).item() * 2 - 1)
    m = m * scale
    theta = float(torch.rand(1, generator=g).item()) * 2 * math.pi
    rot = torch.tensor([[math.cos(theta), math.sin(theta), 0.0], [-math.sin(theta), math.cos(theta), 0.0], [0.0, 0.0, 1.0]],
                       dtype=torch.float64)
    return (m @ rot).to(torch.float32).to(device)


def elastic(x, gran, mag, noise=None, generator=None):
    """x + g(x) * mag with g = a smooth random displacement field (scannet.py:49-70): per axis a noise grid of
    bb = |x|.max(0) // gran + 3 cells, blurred six times with the 3-tap box filters blur0/1/2 (zero padding), sampled at
    the points by trilinear interpolation over the axes linspace(-(b-1)*gran, (b-1)*gran, b) (0 outside).
    `noise` (3 tensors of shape bb, optional) pins the random field for tests."""
    assert x.dim() == 2 and x.size(1) == 3
    dev = x.device
    bb = (x.abs().amax(0).to(torch.int64) // int(gran) + 3).tolist()
    if noise is None:
        noise = [torch.randn(bb, generator=generator, device=dev, dtype=torch.float32) for _ in range(3)]
    field = torch.stack([n.to(dev, torch.float32) for n in noise], 0)[:, None]            # [3,1,b0,b1,b2]
    box = torch.full((1, 1, 3), 1.0 / 3, device=dev)
    for _ in range(2):                                   # blur0, blur1, blur2, twice (:57-62)
        field = F.conv3d(field, box.view(1, 1, 3, 1, 1), padding=(1, 0, 0))
        field = F.conv3d(field, box.view(1, 1, 1, 3, 1), padding=(0, 1, 0))
        field = F.conv3d(field, box.view(1, 1, 1, 1, 3), padding=(0, 0, 1))
    # grid_sample wants normalised coordinates in (x=last axis, y, z=first axis) order; align_corners=True puts -1/+1 on
    # the first/last grid node, i.e. on -(b-1)*gran / +(b-1)*gran
    half = torch.tensor([(b - 1) * gran for b in bb], device=dev, dtype=torch.float32)
    q = (x / half).flip(1).view(1, 1, 1, -1, 3)
    g = F.grid_sample(field.view(1, 3, *bb), q, mode="bilinear", padding_mode="zeros", align_corners=True)
    return x + g.view(3, -1).t() * mag


def to_input_coords(xyz, batch_index, full_scale=4096, offset_rand=None, generator=None):
    """Shift / crop / truncate on the device (scannet.py:133-137,160,210; ioLayers.py:56).
    Returns (coords int64 [P',4] on the device -- the kept points only --, keep bool [P])."""
    xyz = xyz.contiguous().float()
    if offset_rand is None:
        offset_rand = torch.rand(3, generator=generator)
    off = (xyz.amin(0).cpu() - 10.0 + torch.as_tensor(offset_rand, dtype=torch.float32)).tolist()
    n = xyz.size(0)
    coords = torch.empty((n, 4), dtype=torch.int64, device=xyz.device)
    keep = torch.empty(n, dtype=torch.uint8, device=xyz.device)
    with torch.cuda.device(xyz.device):
        _lib.check(_lib.lib().scn_float_coords(C.c_void_p(xyz.data_ptr()), n, (C.c_float * 3)(*off), int(batch_index),
                                               float(full_scale), C.c_void_p(coords.data_ptr()), C.c_void_p(keep.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream().cuda_stream)))
    keep = keep.bool()
    return coords[keep], keep


def prepare_scene(points, colors, batch_index, scale=50, full_scale=4096, use_elastic=True, rotation_noise=True,
                  generator=None):
    """One scene of trainMerge (scannet.py:72-160) on the device: returns (coords int64 [P',4], colours [P',3], keep [P])."""
    dev = points.device
    a = points.float() @ random_affine(scale, rotation_noise, generator, dev)
    if use_elastic:
        r = float(torch.rand(1, generator=generator).item())
        a = elastic(a, 6 * scale // 50, r * 40 * scale / 50, generator=None)
        r = float(torch.rand(1, generator=generator).item())
        a = elastic(a, 20 * scale // 50, r * 160 * scale / 50, generator=None)
    coords, keep = to_input_coords(a, batch_index, full_scale, generator=generator)
    col = (colors.float() + torch.randn(3, generator=generator).to(dev) * 0.1).clamp(-1, 1)
    return coords, col[keep], keep
