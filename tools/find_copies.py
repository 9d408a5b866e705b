"""Which torch ops still run inside one training step (outside libscn_b200.so): python tools/find_copies.py"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200 import scenes
from occuseg_b200.backbone import SparseBackbone
from torch.profiler import profile, ProfilerActivity
scn.set_precision("bf16")
torch.manual_seed(1234)
net = SparseBackbone(m=64, levels=6).cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
coords, feats = scenes.make_batch("S250k", tuple(range(8)))
c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
def step():
    out = net([c, f, None, 8]); out.square().mean().backward(); opt.step(); opt.zero_grad(set_to_none=False)
step(); step(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA], with_stack=True) as prof:
    step(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
for e in prof.key_averages(group_by_stack_n=6):
    if e.key in ("aten::copy_", "aten::contiguous", "aten::cat", "aten::clone") and e.device_time_total > 100:
        print(e.key, e.count, f"{e.device_time_total/1e3:.3f} ms")
        for s in e.stack[:6]: print("     ", s)
