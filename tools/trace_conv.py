"""clock64 breakdown of k_conv_tma (needs a -DSCN_TRACE_BUILD build and SCN_TRACE=1): python tools/trace_conv.py [levels]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200.sparseconvnet import SCN
from occuseg_b200 import scenes
def lt(v): return torch.LongTensor([v, v, v])
levels = int(sys.argv[1]) if len(sys.argv) > 1 else 2
coords, feats = scenes.make_batch("S250k", tuple(range(8)))
m = SCN.Metadata_3(); out = torch.empty(0, device='cuda')
SCN.InputLayer_updateOutput(m, lt(4096), torch.from_numpy(coords), torch.from_numpy(feats).cuda(), out, 8, 4, None)
scn.set_precision('bf16')
size = 4096
for lvl in range(levels):
    C = 64 * (lvl + 1)
    N = m.getNActive(lt(size))
    x = torch.randn(N, C, device='cuda'); w = torch.randn(27, C, C, device='cuda') * 0.05
    y = torch.empty(0, device='cuda')
    SCN.SubmanifoldConvolution_updateOutput(lt(size), lt(3), m, x, y, w, torch.empty(0), 1)
    torch.cuda.synchronize()
    if lvl + 1 < levels:
        m.stridedTable(lt(size), lt(size // 2))
    size //= 2
