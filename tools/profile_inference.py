"""Host-side profile of BASELINE.json config 2 (UNet m=32 inference, one S250k scene): python tools/profile_inference.py"""
import os, sys, time, torch, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from occuseg_b200 import scenes
from occuseg_b200.backbone import SparseBackbone
coords, feats = scenes.make_batch("S250k", (0,))
c, f = torch.from_numpy(coords).cuda(), torch.from_numpy(feats).cuda()
net = SparseBackbone(m=32, levels=6).cuda().eval()
with torch.no_grad():
    for _ in range(3): net([c, f, None, 1])
    torch.cuda.synchronize()
    for _ in range(3):
        t0 = time.perf_counter(); e0 = torch.cuda.Event(True); e1 = torch.cuda.Event(True); e0.record()
        net([c, f, None, 1]); t1 = time.perf_counter(); e1.record(); torch.cuda.synchronize()
        print(f"host enqueue {1e3*(t1-t0):.2f} ms, device {e0.elapsed_time(e1):.2f} ms")
    pr = cProfile.Profile(); pr.enable(); net([c, f, None, 1]); torch.cuda.synchronize(); pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(14)
