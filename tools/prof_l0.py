"""One level-0 submanifold layer (fwd, dgrad, wgrad) for ncu captures: python tools/prof_l0.py [C] [scenes]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200.sparseconvnet import SCN
from occuseg_b200 import scenes
def lt(v): return torch.LongTensor([v, v, v])
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
nsc = int(sys.argv[2]) if len(sys.argv) > 2 else 8
coords, feats = scenes.make_batch("S250k", tuple(range(nsc)))
m = SCN.Metadata_3(); out = torch.empty(0, device='cuda')
SCN.InputLayer_updateOutput(m, lt(4096), torch.from_numpy(coords), torch.from_numpy(feats).cuda(), out, nsc, 4, None)
scn.set_precision(sys.argv[3] if len(sys.argv) > 3 else 'bf16')
N = m.getNActive(lt(4096))
x = torch.randn(N, C, device='cuda'); g = torch.randn(N, C, device='cuda'); w = torch.randn(27, C, C, device='cuda') * 0.05
y = torch.empty(0, device='cuda'); dx = torch.empty(0, device='cuda'); dw = torch.zeros_like(w)
for _ in range(2):
    SCN.SubmanifoldConvolution_updateOutput(lt(4096), lt(3), m, x, y, w, torch.empty(0), 1)
    SCN.SubmanifoldConvolution_backward(lt(4096), lt(3), m, x, dx, g, w, dw, torch.empty(0), 1)
torch.cuda.synchronize()
print("done", N)
