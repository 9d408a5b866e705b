"""Two training steps of the bench workload (UNet m=64, 8 x S250k, bf16) for ncu captures: python tools/prof_step.py [steps]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200 import scenes
from occuseg_b200.backbone import SparseBackbone
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scn.set_precision("bf16")
torch.manual_seed(1234)
net = SparseBackbone(m=64, levels=6).cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
coords, feats = scenes.make_batch("S250k", tuple(range(8)))
c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
for _ in range(steps):
    out = net([c, f, None, 8])
    out.square().mean().backward()
    opt.step()
    opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("done")
