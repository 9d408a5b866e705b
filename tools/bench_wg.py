import sys, numpy as np, torch, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import occuseg_b200.sparseconvnet as scn
from occuseg_b200.sparseconvnet import SCN
from occuseg_b200 import _lib, scenes
def lt(v): return torch.LongTensor([v,v,v])
nsc=8
coords,feats=scenes.make_batch("S250k",tuple(range(nsc)))
m=SCN.Metadata_3(); out=torch.empty(0,device='cuda')
SCN.InputLayer_updateOutput(m, lt(4096), torch.from_numpy(coords), torch.from_numpy(feats).cuda(), out, nsc, 4, None)
scn.set_precision('tf32')
size=4096; C=int(sys.argv[1]) if len(sys.argv)>1 else 64
N=m.getNActive(lt(size)); nbr,R=m.submanifoldNeighbourTable(lt(size))
x=torch.randn(N,C,device='cuda'); g=torch.randn(N,C,device='cuda'); w=torch.randn(27,C,C,device='cuda')*0.05
dx=torch.empty(0,device='cuda'); dw=torch.zeros_like(w)
for _ in range(2): SCN.SubmanifoldConvolution_backward(lt(size),lt(3),m,x,dx,g,w,dw,torch.empty(0),1)
_lib.profile(True)
for _ in range(3): SCN.SubmanifoldConvolution_backward(lt(size),lt(3),m,x,dx,g,w,dw,torch.empty(0),1)
torch.cuda.synchronize(); pr=_lib.profile_read(); _lib.profile(False)
print("DBG",os.environ.get("SCN_WG_DBG"),"C",C,"wgrad ms",pr['wgrad_tc']['ms']/3, "dgrad", pr['conv_tc']['ms']/3)
