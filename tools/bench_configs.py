"""BASELINE.json configs 2 and 5 (the ones bench.py does not time):
   config 2: UNet m=32 inference on one S250k scene;  config 5: one S1M scene, UNet m=64 fwd+bwd + single-layer channel sweep.
   python tools/bench_configs.py"""
import os, sys, time, torch, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200.sparseconvnet import SCN
from occuseg_b200 import scenes, _lib
from occuseg_b200.backbone import SparseBackbone

def lt(v): return torch.LongTensor([v, v, v])
def ev(fn, reps=3, warm=2):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# ---- config 2
torch.manual_seed(0)
coords, feats = scenes.make_batch("S250k", (0,))
c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
net = SparseBackbone(m=32, levels=6).cuda().eval()
with torch.no_grad():
    t = ev(lambda: net([c, f, None, 1]))
    n = net.input([c, f, None, 1]).features.shape[0]
    out_b = net([c, f, None, 1])
    scn.set_precision("fp32"); out_f = net([c, f, None, 1]); scn.set_precision("bf16")
err = float((out_b - out_f).abs().max() / out_f.abs().max())
print(f"config 2: UNet m=32 inference, 1 x S250k ({n} voxels): {t:.2f} ms, {n / t / 1e3:.2f} M voxels/s; default-precision vs fp32 path rel err {err:.2e}")

# ---- config 5
coords, feats = scenes.make_batch("S1M", (0,))
c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
net = SparseBackbone(m=64, levels=6).cuda()
def step():
    out = net([c, f, None, 1]); out.square().mean().backward()
t = ev(step)
with torch.no_grad():
    n = net.input([c, f, None, 1]).features.shape[0]
print(f"config 5: UNet m=64 fwd+bwd, 1 x S1M ({n} voxels): {t:.2f} ms, {n / t / 1e3:.2f} M voxels/s")
m = SCN.Metadata_3(); out = torch.empty(0, device='cuda')
SCN.InputLayer_updateOutput(m, lt(4096), torch.from_numpy(coords), f, out, 1, 4, None)
N = m.getNActive(lt(4096)); _, R = m.submanifoldNeighbourTable(lt(4096))
for C in (16, 32, 64, 128, 256):
    x = torch.randn(N, C, device='cuda'); g = torch.randn(N, C, device='cuda'); w = torch.randn(27, C, C, device='cuda') * 0.05
    y = torch.empty(0, device='cuda'); dx = torch.empty(0, device='cuda'); dw = torch.zeros_like(w)
    tf = ev(lambda: SCN.SubmanifoldConvolution_updateOutput(lt(4096), lt(3), m, x, y, w, torch.empty(0), 1))
    tb = ev(lambda: SCN.SubmanifoldConvolution_backward(lt(4096), lt(3), m, x, dx, g, w, dw, torch.empty(0), 1))
    fl = 2.0 * R * C * C
    print(f"  sweep C={C:3d} (N={N}, R={R}): fwd {tf:.3f} ms ({fl / tf / 1e9:.0f} TFLOP/s)  bwd(dgrad+wgrad) {tb:.3f} ms ({2 * fl / tb / 1e9:.0f} TFLOP/s)")
