"""Where the HOST time of one training step goes (cProfile over 5 steps, no device sync inside): python tools/host_profile.py"""
import cProfile, pstats, os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200 import scenes
from occuseg_b200.backbone import SparseBackbone
scn.set_precision("bf16")
torch.manual_seed(1234)
net = SparseBackbone(m=64, levels=6).cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
coords, feats = scenes.make_batch("S250k", tuple(range(8)))
c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
def step():
    out = net([c, f, None, 8]); out.square().mean().backward(); opt.step(); opt.zero_grad(set_to_none=True)
for _ in range(3): step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5): step()
t1 = time.perf_counter(); torch.cuda.synchronize()
print(f"host enqueue {(t1 - t0) * 200:.2f} ms/step, with device {(time.perf_counter() - t0) * 200:.2f} ms/step")
pr = cProfile.Profile(); pr.enable()
for _ in range(5): step()
pr.disable(); torch.cuda.synchronize()
st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(28)
