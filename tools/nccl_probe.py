"""Does an NCCL communicator change the speed of single kernels?  torchrun --nproc-per-node 2 tools/nccl_probe.py"""
import os, sys, torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200.sparseconvnet import SCN
from occuseg_b200 import scenes
rank, local = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
def lt(v): return torch.LongTensor([v, v, v])
coords, feats = scenes.make_batch("S250k", tuple(range(8)))
m = SCN.Metadata_3(); out = torch.empty(0, device='cuda')
SCN.InputLayer_updateOutput(m, lt(4096), torch.from_numpy(coords), torch.from_numpy(feats).cuda(), out, 8, 4, None)
scn.set_precision('bf16')
N = m.getNActive(lt(4096)); C = 64
x = torch.randn(N, C, device='cuda'); w = torch.randn(27, C, C, device='cuda') * 0.05
r = torch.randn(N, C, device='cuda'); g = torch.randn(N, C, device='cuda')
st = torch.empty(2, C, dtype=torch.float64, device='cuda')
def ev(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def measure(tag):
    y = torch.empty(0, device='cuda'); dx = torch.empty(0, device='cuda'); dw = torch.zeros_like(w)
    a = ev(lambda: SCN.SubmanifoldConvolution_updateOutput(lt(4096), lt(3), m, x, y, w, torch.empty(0), 1))
    b = ev(lambda: SCN.SubmanifoldConvolution_updateOutput(lt(4096), lt(3), m, x, y, w, torch.empty(0), 1, r, st))
    c = ev(lambda: SCN.SubmanifoldConvolution_updateOutput(lt(4096), lt(3), m, x, y, w, torch.empty(0), 1, r, None))
    d = ev(lambda: SCN.SubmanifoldConvolution_backward(lt(4096), lt(3), m, x, dx, g, w, dw, torch.empty(0), 1))
    e = ev(lambda: torch.mm(x.t(), g))
    print(f"[rank {rank}] {tag}: fwd plain {a:.3f}  fwd+residual+stats {b:.3f}  fwd+residual {c:.3f}  bwd(dgrad+wgrad+casts) {d:.3f}  mm {e:.3f} ms", flush=True)
measure("before init")
if int(os.environ.get("WORLD_SIZE", 1)) > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dist.barrier()
    measure("after NCCL init")
    t = torch.ones(1 << 20, device="cuda"); dist.all_reduce(t); torch.cuda.synchronize()
    measure("after an all-reduce")
    dist.destroy_process_group()
    measure("after destroy")
