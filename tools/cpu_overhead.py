"""How long does the host take to enqueue one training step, against the device time?  python tools/cpu_overhead.py"""
import os, sys, time, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200 import scenes
from occuseg_b200.backbone import SparseBackbone
torch.manual_seed(0)
net = SparseBackbone(m=64, levels=6).cuda()
opt = torch.optim.Adam(net.parameters(), lr=1e-3, fused=True)
coords, feats = scenes.make_batch("S250k", tuple(range(8)))
c = torch.from_numpy(coords).cuda(); f = torch.from_numpy(feats).cuda()
def step():
    out = net([c, f, None, 8]); loss = out.square().mean(); loss.backward(); opt.step(); opt.zero_grad(set_to_none=False)
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(3):
    t0 = time.perf_counter(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
    step()
    t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host enqueue {1e3*(t1-t0):.1f} ms, device {e0.elapsed_time(e1):.1f} ms, wall {1e3*(t2-t0):.1f} ms")
if len(sys.argv) > 1:
    import cProfile, pstats
    pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(35)
